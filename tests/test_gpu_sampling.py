"""Multinomial sampling (sample_max = 0) and the self-critical loss on the CUDA path against the CPU oracle
(models/AttModel.py:231-239, misc/criterion.py:104-124, trainer.py:166-173)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import unpaired_image_captioning_b200 as uic  # noqa: E402
from oracle import decoder_oracle as O  # noqa: E402
from unpaired_image_captioning_b200 import _lib, synth  # noqa: E402
from unpaired_image_captioning_b200._lib import check, ptr, stream  # noqa: E402
from parity import compare_greedy  # noqa: E402

DEV = "cuda"


def test_sampling_epilogue_matches_the_oracle_noise():
    """uic_logit_stats(temperature, seed) + uic_greedy_advance == argmax(logits / T + oracle Gumbel noise), and the
    reported value is the log-prob of the unperturbed logit."""
    lib = _lib.load()
    _lib.require_device()
    R, V, H, T, temperature, seed, step = 90, 3000, 128, 5, 0.8, 424242, 2
    g = torch.Generator().manual_seed(0)
    h = torch.randn(R, H, generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn(V, H, generator=g) * 0.05).to(DEV).to(torch.bfloat16)
    bias = (torch.randn(V, generator=g) * 0.1).to(DEV)
    logits = h.float() @ w.float().t() + bias
    parts = lib.uic_logit_stats_parts(R, V)
    stats = torch.empty(R, parts, 4, device=DEV)
    seed_t = torch.tensor([seed], dtype=torch.int64, device=DEV)
    seq, lp = torch.zeros(R, T, dtype=torch.int64, device=DEV), torch.zeros(R, T, device=DEV)
    unf, tok = torch.ones(R, dtype=torch.uint8, device=DEV), torch.zeros(R, dtype=torch.int64, device=DEV)
    nunf = torch.zeros(T, dtype=torch.int32, device=DEV)
    nunf[step - 1] = R
    check(lib.uic_logit_stats(ptr(h), H, ptr(w), H, ptr(bias), None, 0, ptr(stats), R, V, H, 1, 0, temperature, ptr(seed_t), step,
                              stream()))
    check(lib.uic_greedy_advance(ptr(stats), parts, ptr(seq), ptr(lp), ptr(unf), ptr(tok), ptr(nunf), step, T, R, None, 0, None, 0,
                                 0, V, temperature, ptr(seed_t), stream()))
    keys = logits.cpu() / temperature + O.gumbel_noise(seed, step, R, V)
    top2 = keys.topk(2, dim=1)
    clear = (top2.values[:, 0] - top2.values[:, 1]) > 1e-2
    assert float(clear.float().mean()) > 0.9
    want = top2.indices[:, 0]
    assert torch.equal(tok.cpu()[clear], want[clear])
    ref_lp = torch.log_softmax(logits.cpu(), 1).gather(1, tok.cpu()[:, None]).squeeze(1)
    torch.testing.assert_close(lp[:, step].cpu(), ref_lp, rtol=1e-3, atol=2e-3)


def test_sampling_frequencies_follow_softmax():
    """20k rows with the SAME logits: the sampled tokens follow softmax(logits / T) (chi-square)."""
    lib = _lib.load()
    R, V, H, temperature = 20480, 24, 32, 1.3
    g = torch.Generator().manual_seed(1)
    h1 = torch.randn(1, H, generator=g)
    h = h1.expand(R, H).contiguous().to(DEV).to(torch.bfloat16)
    w = (torch.randn(V, H, generator=g) * 0.3).to(DEV).to(torch.bfloat16)
    bias = torch.zeros(V, device=DEV)
    logits = (h[:1].float() @ w.float().t()).cpu()[0]
    parts = lib.uic_logit_stats_parts(R, V)
    stats = torch.empty(R, parts, 4, device=DEV)
    seed_t = torch.tensor([77], dtype=torch.int64, device=DEV)
    T = 2
    seq, lp = torch.zeros(R, T, dtype=torch.int64, device=DEV), torch.zeros(R, T, device=DEV)
    unf, tok = torch.ones(R, dtype=torch.uint8, device=DEV), torch.zeros(R, dtype=torch.int64, device=DEV)
    nunf = torch.zeros(T, dtype=torch.int32, device=DEV)
    check(lib.uic_logit_stats(ptr(h), H, ptr(w), H, ptr(bias), None, 0, ptr(stats), R, V, H, 1, 0, temperature, ptr(seed_t), 0, stream()))
    check(lib.uic_greedy_advance(ptr(stats), parts, ptr(seq), ptr(lp), ptr(unf), ptr(tok), ptr(nunf), 0, T, R, None, 0, None, 0, 0, V,
                                 temperature, ptr(seed_t), stream()))
    counts = torch.bincount(seq[:, 0].cpu(), minlength=V).double()
    p = torch.softmax(logits.double() / temperature, 0)
    chi2 = float(((counts - R * p) ** 2 / (R * p)).sum())
    assert chi2 < 65.0, chi2        # 23 degrees of freedom: P(chi2 > 65) ~ 1e-5


@pytest.mark.parametrize("kind,L", [("att2in2", 49), ("topdown", 36)])
def test_model_sampling_matches_oracle(kind, L):
    opt = synth.make_opt(caption_model=kind, vocab_size=9999, rnn_size=512, input_encoding_size=512, att_hid_size=512, seq_length=16)
    sd = synth.init_state_dict(opt, seed=21, eos_bias=3.0)
    fc, att = synth.make_features(12, L, 2048, seed=21)
    model = uic.setup(opt)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    o = {"sample_max": 0, "temperature": 0.9, "seed": 31337}
    ref_seq, ref_lp, margins = O.sample_multinomial(sd, kind, fc, att, 16, temperature=0.9, seed=31337, return_margins=True)
    with torch.no_grad():
        seq, lp = model(fc.cuda(), None, att.cuda(), None, opt=o, mode="sample")
        seq_b, _ = model(fc.cuda(), None, att.cuda(), None, opt=o, mode="sample")
        seq_c, _ = model(fc.cuda(), None, att.cuda(), None, opt=dict(o, seed=4), mode="sample")
    assert torch.equal(seq, seq_b) and not torch.equal(seq, seq_c)          # a function of the seed only
    exact, exempt, failures = compare_greedy(seq.cpu(), ref_seq, margins, tol=1e-2)
    assert not failures, failures
    assert exact >= 3
    rows = (seq.cpu() == ref_seq).all(1)
    torch.testing.assert_close(lp.cpu()[rows], ref_lp[rows], rtol=1e-2, atol=2e-2)
    assert int((ref_seq == 0).sum()) > 0 and int((ref_seq > 0).sum()) > 0


@pytest.mark.parametrize("kind,L,B", [("att2in2", 49, 6), ("topdown", 36, 8)])
def test_self_critical_gradients_match_oracle(kind, L, B):
    """trainer.py:166-173: sample, reward, RewardCriterion, backward -- gradients against the oracle's autograd for the
    SAME sampled tokens and rewards."""
    opt = synth.make_opt(caption_model=kind, vocab_size=9999, rnn_size=512, input_encoding_size=512, att_hid_size=512, seq_length=16)
    sd = synth.init_state_dict(opt, seed=13, eos_bias=3.0)
    fc, att = synth.make_features(B, L, 2048, seed=13)
    model = uic.setup(opt)
    model.load_state_dict(sd)
    model = model.cuda().train()
    gen, sample_lp = model(fc.cuda(), None, att.cuda(), None, opt={"sample_max": 0, "seed": 5}, mode="sample")
    assert sample_lp.requires_grad
    g = torch.Generator().manual_seed(2)
    reward = torch.randn(B, 1, generator=g).expand(B, 16).contiguous()     # misc/rewards.py:80 repeats one reward per row
    loss = uic.RewardCriterion()(sample_lp, gen, reward.cuda())
    loss.backward()
    ref_loss, ref_grads, ref_lp = O.rl_loss_and_grads(sd, kind, fc, att, gen.cpu(), reward)
    written = sample_lp.detach().cpu() != 0
    torch.testing.assert_close(sample_lp.detach().cpu()[written], ref_lp[written], rtol=1e-2, atol=1e-2)
    assert abs(float(loss) - float(ref_loss)) < 2e-2 * max(1.0, abs(float(ref_loss)))
    errs = {}
    for name, p in model.named_parameters():
        ref = ref_grads[name].cuda()
        gr = p.grad if p.grad is not None else torch.zeros_like(p)
        errs[name] = float(gr.abs().max()) if float(ref.abs().max()) < 1e-7 else float((gr - ref).norm()) / float(ref.norm())
    bad = {k: v for k, v in errs.items() if v > 6e-2}
    assert not bad, bad


@pytest.mark.parametrize("kind,L,B", [("att2in2", 49, 12), ("topdown", 36, 12), ("att2in2", 33, 1)])
def test_scheduled_sampling_matches_oracle(kind, L, B):
    """Training-mode forward with ss_prob > 0 (models/AttModel.py:130-143): the input tokens actually used, the loss and
    the gradients against the oracle's restatement with the same counter-based draws."""
    opt = synth.make_opt(caption_model=kind, vocab_size=9999, rnn_size=512, input_encoding_size=512, att_hid_size=512, seq_length=16)
    sd = synth.init_state_dict(opt, seed=17)
    fc, att = synth.make_features(B, L, 2048, seed=17)
    labels, masks = synth.make_captions(B, 16, 9999, seed=17)
    model = uic.setup(opt)
    model.load_state_dict(sd)
    model = model.cuda().train()
    model.ss_prob, model.ss_seed = 0.5, 909
    _, ref_used, ref_margins = O.teacher_forced(sd, kind, fc, att, labels, ss_prob=0.5, ss_seed=909, return_tokens=True)
    # the product path, with the tokens it fed kept for inspection
    from unpaired_image_captioning_b200 import autograd as AG
    ss = model._scheduled_sampling(torch.device("cuda"))
    r = AG.teacher_forced_run(model, fc.cuda(), att.cuda(), labels.cuda(), None, all_steps=True, ss=ss)
    used = r.tokens.t().cpu()                                            # (B, T)
    n = min(used.size(1), ref_used.size(1))
    exact, exempt, failures = compare_greedy(used[:, :n], ref_used[:, :n], ref_margins[:, :n], tol=1e-2)
    assert not failures, failures
    assert exact >= B // 2
    sampled = torch.isfinite(ref_margins[:, 1:n])
    assert (0.3 if B > 1 else 0.1) < float(sampled.float().mean()) < (0.7 if B > 1 else 0.9)
    assert int((ref_used[:, 1:n] != labels[:, 1:n])[sampled].sum()) > 0   # the draws really replace ground truth
    labels_dev = labels.cuda()
    AG.teacher_forced_run(model, fc.cuda(), att.cuda(), labels_dev, None, all_steps=True, ss=ss)
    assert torch.equal(labels_dev.cpu(), labels)                          # the caller's labels are never written (B == 1 view aliasing)
    # loss + gradients through the public fast path, against the oracle fed with the draws the product path made
    # (identical to its own wherever the margins are clear)
    loss = model(fc.cuda(), None, att.cuda(), labels.cuda(), masks.cuda(), None, mode="forward_loss")
    loss.backward()
    ref_loss, ref_grads = O.loss_and_grads(sd, kind, fc, att, labels, masks, inputs=used)
    assert abs(float(loss.detach()) - float(ref_loss)) < 2e-3 * float(ref_loss)
    errs = {}
    for name, p in model.named_parameters():
        ref = ref_grads[name].cuda()
        gr = p.grad if p.grad is not None else torch.zeros_like(p)
        errs[name] = float(gr.abs().max()) if float(ref.abs().max()) < 1e-7 else float((gr - ref).norm()) / float(ref.norm())
    bad = {k: v for k, v in errs.items() if v > 6e-2}
    assert not bad, bad


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dropout_kernel_matches_oracle_mask(dtype):
    seed_t = torch.tensor([123456789], dtype=torch.int64, device=DEV)
    drop = (0.3, seed_t)
    x = torch.randn(300, 72, device=DEV).to(dtype)
    y = x.clone()
    view = y[:, 8:8 + 40]                                      # pitched view: ld 72, 40 columns
    _lib.dropout(view, drop, _lib.DROP_XT, row0=17, row_stride=3)
    mask = O.dropout_mask((0.3, 123456789), O.DROP_XT, 17 + 3 * torch.arange(300).numpy(), 40).to(DEV)
    want = (x[:, 8:48].float() * mask).to(dtype)
    torch.testing.assert_close(y[:, 8:48].float(), want.float(), rtol=1e-2 if dtype == torch.bfloat16 else 1e-6, atol=1e-6)
    assert torch.equal(y[:, :8], x[:, :8]) and torch.equal(y[:, 48:], x[:, 48:])
    assert torch.equal((y[:, 8:48] == 0), (mask == 0) | (x[:, 8:48] == 0))


@pytest.mark.parametrize("kind,L,B", [("att2in2", 49, 6), ("topdown", 36, 8)])
def test_training_with_dropout_matches_oracle(kind, L, B):
    """drop_prob_lm = 0.5 in training mode (the reference's default): loss and every gradient against the oracle with the
    same counter-based masks on embed / fc_embed / att_embed / core output."""
    opt = synth.make_opt(caption_model=kind, vocab_size=9999, rnn_size=512, input_encoding_size=512, att_hid_size=512, seq_length=16,
                         drop_prob_lm=0.5)
    sd = synth.init_state_dict(opt, seed=23)
    fc, att = synth.make_features(B, L, 2048, seed=23)
    labels, masks = synth.make_captions(B, 16, 9999, seed=23)
    model = uic.setup(opt)
    model.load_state_dict(sd)
    model = model.cuda().train()
    model.dropout_seed = 4242
    loss = model(fc.cuda(), None, att.cuda(), labels.cuda(), masks.cuda(), None, mode="forward_loss")
    loss.backward()
    ref_loss, ref_grads = O.loss_and_grads(sd, kind, fc, att, labels, masks, drop=(0.5, 4242))
    plain_loss = O.train_loss(sd, kind, fc, att, labels, masks)
    assert abs(float(ref_loss) - float(plain_loss)) > 1e-4 * float(plain_loss)        # the masks matter
    assert abs(float(loss.detach()) - float(ref_loss)) < 2e-3 * float(ref_loss)
    errs = {}
    for name, p in model.named_parameters():
        ref = ref_grads[name].cuda()
        gr = p.grad if p.grad is not None else torch.zeros_like(p)
        errs[name] = float(gr.abs().max()) if float(ref.abs().max()) < 1e-7 else float((gr - ref).norm()) / float(ref.norm())
    # (kept activations are doubled at p = 0.5 and so is their bf16 rounding: the 8-row fc_embed gradients are the noisiest)
    bad = {k: v for k, v in errs.items() if v > 9e-2}
    assert not bad, bad
    # eval mode ignores dropout
    model.eval()
    with torch.no_grad():
        out = model(fc.cuda(), None, att.cuda(), labels.cuda(), None)
    ref = O.teacher_forced(sd, kind, fc, att, labels)
    sel = masks[:, 1:].bool()
    assert float(((out.cpu() - ref).abs() / ref.abs().clamp_min(1.0))[sel].max()) < 2e-3


@pytest.mark.parametrize("kind,L,B", [("att2in2", 49, 8), ("topdown", 36, 8)])
def test_self_critical_step_in_train_mode_with_dropout(kind, L, B):
    """trainer.py:166-173 as the reference runs it: model.train() with drop_prob_lm = 0.5, so the roll-out and the
    log-probs that are differentiated both see the dropout masks.  Roll-out against the oracle (same masks, same
    Gumbel noise), roll-out log-probs == differentiable log-probs, loss and gradients against the oracle's autograd."""
    opt = synth.make_opt(caption_model=kind, vocab_size=9999, rnn_size=512, input_encoding_size=512, att_hid_size=512, seq_length=16,
                         drop_prob_lm=0.5)
    sd = synth.init_state_dict(opt, seed=29, eos_bias=3.0)
    fc, att = synth.make_features(B, L, 2048, seed=29)
    model = uic.setup(opt)
    model.load_state_dict(sd)
    model = model.cuda().train()
    model.dropout_seed = 777
    o = {"sample_max": 0, "seed": 55}
    ref_seq, ref_lp, margins = O.sample_multinomial(sd, kind, fc, att, 16, seed=55, return_margins=True, drop=(0.5, 777))
    with torch.no_grad():
        seq_ng, lp_ng = model(fc.cuda(), None, att.cuda(), None, opt=o, mode="sample")
    exact, exempt, failures = compare_greedy(seq_ng.cpu(), ref_seq, margins, tol=1e-2)
    assert not failures, failures
    assert exact >= 2
    plain_seq, _ = O.sample_multinomial(sd, kind, fc, att, 16, seed=55)
    assert not torch.equal(ref_seq, plain_seq)                         # the masks change the roll-out
    gen, sample_lp = model(fc.cuda(), None, att.cuda(), None, opt=o, mode="sample")
    assert torch.equal(gen, seq_ng) and sample_lp.requires_grad
    written = lp_ng != 0
    torch.testing.assert_close(sample_lp.detach()[written], lp_ng[written], rtol=2e-2, atol=2e-2)
    g = torch.Generator().manual_seed(2)
    reward = torch.randn(B, 1, generator=g).expand(B, 16).contiguous()
    loss = uic.RewardCriterion()(sample_lp, gen, reward.cuda())
    loss.backward()
    ref_loss, ref_grads, _ = O.rl_loss_and_grads(sd, kind, fc, att, gen.cpu(), reward, drop=(0.5, 777))
    assert abs(float(loss.detach()) - float(ref_loss)) < 2e-2 * max(1.0, abs(float(ref_loss)))
    errs = {}
    for name, p in model.named_parameters():
        ref = ref_grads[name].cuda()
        gr = p.grad if p.grad is not None else torch.zeros_like(p)
        errs[name] = float(gr.abs().max()) if float(ref.abs().max()) < 1e-7 else float((gr - ref).norm()) / float(ref.norm())
    bad = {k: v for k, v in errs.items() if v > 9e-2}
    assert not bad, bad
