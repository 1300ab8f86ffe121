"""Randomised shapes for the public model API against the oracle (used by tests/test_gpu_shapes.py and scripts/shape_sweep.py):
odd region counts, ragged masks, batch sizes down to 1, one-step sequences, every beam width, several layer widths."""
import random

import torch

import unpaired_image_captioning_b200 as uic
from oracle import decoder_oracle as O
from unpaired_image_captioning_b200 import synth
from parity import compare_beam, compare_greedy


def draw(rng):
    return dict(kind=rng.choice(["att2in2", "att2all2", "topdown"]), H=rng.choice([32, 64, 96, 128, 256, 512]),
                E=rng.choice([32, 64, 128, 512]), A=rng.choice([32, 64, 128, 512]), D=rng.choice([64, 128, 264]),
                V=rng.choice([51, 99, 500, 1237]), L=rng.choice([1, 2, 3, 5, 7, 15, 16, 17, 31, 33, 36, 49, 64, 100]),
                B=rng.choice([1, 2, 3, 5, 8, 13]), T=rng.choice([1, 2, 5, 9, 16]), beam=rng.choice([1, 2, 3, 4, 5, 7, 8, 10]),
                use_masks=rng.random() < 0.5)


def draw_big(rng):
    """Benchmark-sized shapes: several logit chunks in training, whole-job and segmented attention plans, big GEMM grids."""
    return dict(kind=rng.choice(["att2in2", "att2all2", "topdown"]), H=rng.choice([512, 1024]), E=512, A=512, D=2048,
                V=rng.choice([9999, 9486]), L=rng.choice([36, 196]), B=rng.choice([64, 150, 256, 300]), T=16,
                beam=rng.choice([3, 5]), use_masks=rng.random() < 0.5)


def cases(n, seed, big=False):
    rng = random.Random(seed)
    return [(draw_big if big else draw)(rng) for _ in range(n)]


def run_case(case, c, strict_sampling=True):
    """Returns (list of failure messages, summary string).  Tolerance checks use plain weights; sampling uses the
    wide-margin variant.  strict_sampling=False only requires the sampling calls to run and to be well-formed (beam
    search over a near-tie legitimately returns another row)."""
    kind, H, E, A, D, V, L, B, T, beam, use_masks = (c[k] for k in ("kind", "H", "E", "A", "D", "V", "L", "B", "T", "beam", "use_masks"))
    opt = synth.make_opt(caption_model=kind, vocab_size=V, rnn_size=H, input_encoding_size=E, att_hid_size=A, seq_length=T,
                         fc_feat_size=D, att_feat_size=D)
    sd_peaked = synth.init_state_dict(opt, seed=100 + case, peaked=40.0, eos_bias=1.0)
    sd = synth.init_state_dict(opt, seed=100 + case)
    fc, att = synth.make_features(B, L, D, seed=100 + case)
    labels, masks = synth.make_captions(B, T, V, seed=100 + case, min_len=1)
    am = synth.make_att_masks(B, L, seed=100 + case) if use_masks else None
    model = uic.setup(opt)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    cu = lambda t: None if t is None else t.cuda()
    msg = []
    ref = O.teacher_forced(sd, kind, fc, att, labels, am)
    with torch.no_grad():
        out = model(cu(fc), None, cu(att), cu(labels), cu(am))
    sel = masks[:, 1:].bool()
    rel = float(((out.cpu() - ref).abs() / ref.abs().clamp_min(1.0))[sel].max())
    if not (rel < 3e-3):
        msg.append(f"teacher-forced rel err {rel:.3e}")
    model.load_state_dict(sd_peaked)
    g_ref, g_lp, margins = O.sample_greedy(sd_peaked, kind, fc, att, T, am, return_margins=True, relative_margins=True)
    g_seq, g_lpc = model(cu(fc), None, cu(att), cu(am), opt={"beam_size": 1}, mode="sample")
    assert g_seq.shape == (B, T) and g_seq.dtype == torch.int64 and g_lpc.shape == (B, T)
    exact, exempt, failures = compare_greedy(g_seq.cpu(), g_ref, margins, tol=3e-3)   # (the tolerance of the log-probs above: widths down to 32 average the bf16 rounding over few terms)
    if failures:
        msg.append(f"greedy mismatches {failures[:2]}")
    if 1 < beam <= V:
        b_seq, b_lpc = model(cu(fc), None, cu(att), cu(am), opt={"beam_size": beam}, mode="sample")
        assert b_seq.shape == (B, T) and len(model.done_beams) == B
        for k in range(B):
            ps = [d["p"] for d in model.done_beams[k]]
            assert 1 <= len(ps) <= beam and all(ps[i] >= ps[i + 1] for i in range(len(ps) - 1))
        if strict_sampling:
            b_ref, b_lp, _, b_margins = O.sample_beam(sd_peaked, kind, fc, att, T, beam, am, return_margins=True)
            rows = (b_seq == b_ref).all(1)
            _, _, b_fail = compare_beam(b_seq, b_ref, b_margins, tol=3e-3)
            if b_fail:
                msg.append(f"beam mismatches not at a near-tie {b_fail[:2]}")
            elif rows.any() and float((b_lpc[rows] - b_lp[rows]).abs().max()) > 5e-2:
                msg.append(f"beam logprob diff {float((b_lpc[rows] - b_lp[rows]).abs().max()):.3e}")
    model.load_state_dict(sd)
    model.train()
    ref_loss, ref_grads = O.loss_and_grads(sd, kind, fc, att, labels, masks, am)
    loss = model(cu(fc), None, cu(att), cu(labels), cu(masks), cu(am), mode="forward_loss")
    loss.backward()
    if abs(float(loss.detach()) - float(ref_loss)) > 2e-3 * max(1.0, abs(float(ref_loss))):
        msg.append(f"loss {float(loss.detach()):.5f} vs {float(ref_loss):.5f}")
    worst = ("", 0.0)
    for name, p in model.named_parameters():
        r = ref_grads[name]
        if float(r.abs().max()) < 1e-7:
            continue
        e = float((p.grad.cpu() - r).norm() / r.norm())
        if e > worst[1]:
            worst = (name, e)
    if worst[1] > 0.2:    # a handful of rows: one ReLU / maxout unit within bf16 rounding of its kink moves a whole gradient row
        msg.append(f"grad {worst[0]} rel {worst[1]:.3f}")
    return msg, f"tf {rel:.1e} grad {worst[1]:.3f}"


VARIANTS = ("dropout", "scheduled_sampling", "self_critical", "use_bn", "diverse_beam")


def _grad_worst(model, ref_grads):
    worst = ("", 0.0)
    for name, p in model.named_parameters():
        r = ref_grads[name]
        if float(r.abs().max()) < 1e-7:
            continue
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        e = float((g.cpu() - r).norm() / r.norm())
        if e > worst[1]:
            worst = (name, e)
    return worst


def run_variant(case, c, variant):
    """One of the optional paths (VARIANTS) at the drawn shape, against the oracle; returns (failure messages, summary)."""
    kind, H, E, A, D, V, L, B, T, beam, use_masks = (c[k] for k in ("kind", "H", "E", "A", "D", "V", "L", "B", "T", "beam", "use_masks"))
    use_masks = use_masks or variant == "use_bn"          # BatchNorm1d only runs on the packed (masked) form
    kw = dict(caption_model=kind, vocab_size=V, rnn_size=H, input_encoding_size=E, att_hid_size=A, seq_length=T, fc_feat_size=D,
              att_feat_size=D)
    if variant == "dropout":
        kw["drop_prob_lm"] = 0.5
    if variant == "use_bn":
        kw["use_bn"] = 1
    if variant == "use_bn" and B * L == 1:
        return [], "skipped (one value per channel: torch's BatchNorm1d raises in train mode, and so does this path)"
    opt = synth.make_opt(**kw)
    sd = synth.init_state_dict(opt, seed=300 + case, eos_bias=2.0 if variant == "self_critical" else 0.0,
                               peaked=40.0 if variant == "diverse_beam" else 0.0)
    fc, att = synth.make_features(B, L, D, seed=300 + case)
    labels, masks = synth.make_captions(B, T, V, seed=300 + case, min_len=1)
    am = synth.make_att_masks(B, L, seed=300 + case) if use_masks else None
    model = uic.setup(opt)
    model.load_state_dict(sd)
    model = model.cuda().train()
    cu = lambda t: None if t is None else t.cuda()
    msg, worst = [], ("", 0.0)
    if variant == "diverse_beam":
        model.eval()
        groups = [g for g in (2, 3, 4, 5) if beam % g == 0 and beam <= V]
        if not groups:
            return [], "skipped (beam width has no group divisor)"
        G = groups[case % len(groups)]
        o = {"beam_size": beam, "group_size": G, "diversity_lambda": 0.5}
        seq, lp = model(cu(fc), None, cu(att), cu(am), opt=o, mode="sample")
        ref_seq, ref_lp, ref_done = O.sample_beam(sd, kind, fc, att, T, beam, am, group_size=G, diversity_lambda=0.5)
        assert seq.shape == (B, T)
        for k in range(B):
            assert len(model.done_beams[k]) == len(ref_done[k])
        first = (seq[:, 0] == ref_seq[:, 0]).float().mean()
        if float(first) < 0.5:
            msg.append(f"diverse beam: first tokens equal {float(first):.2f}")
        return msg, f"G={G} rows equal {float((seq == ref_seq).all(1).float().mean()):.2f}"
    if variant == "self_critical":
        gen, sample_lp = model(cu(fc), None, cu(att), cu(am), opt={"sample_max": 0, "seed": 5 + case}, mode="sample")
        g = torch.Generator().manual_seed(case)
        reward = torch.randn(B, 1, generator=g).expand(B, T).contiguous()
        loss = uic.RewardCriterion()(sample_lp, gen, reward.cuda())
        loss.backward()
        ref_loss, ref_grads, ref_lp = O.rl_loss_and_grads(sd, kind, fc, att, gen.cpu(), reward, am)
        written = sample_lp.detach().cpu() != 0
        d = float((sample_lp.detach().cpu()[written] - ref_lp[written]).abs().max()) if written.any() else 0.0
        if d > 2e-2:
            msg.append(f"sample log-prob diff {d:.3e}")
    else:
        ref_kw = {}
        if variant == "dropout":
            model.dropout_seed = 4242 + case
            ref_kw["drop"] = (0.5, 4242 + case)
        if variant == "scheduled_sampling":
            model.ss_prob, model.ss_seed = 0.5, 909 + case
            from unpaired_image_captioning_b200 import autograd as AG
            r = AG.teacher_forced_run(model, cu(fc), cu(att), cu(labels), cu(am), all_steps=True, ss=model._scheduled_sampling(torch.device("cuda")))
            ref_kw["inputs"] = r.tokens.t().cpu()        # the oracle is fed the draws the product path made
        loss = model(cu(fc), None, cu(att), cu(labels), cu(masks), cu(am), mode="forward_loss")
        loss.backward()
        if variant == "use_bn":
            with O.bn_training():
                ref_loss, ref_grads = O.loss_and_grads(sd, kind, fc, att, labels, masks, am)
        else:
            ref_loss, ref_grads = O.loss_and_grads(sd, kind, fc, att, labels, masks, am, **ref_kw)
    if abs(float(loss.detach()) - float(ref_loss)) > 5e-3 * max(1.0, abs(float(ref_loss))):
        msg.append(f"{variant}: loss {float(loss.detach()):.5f} vs {float(ref_loss):.5f}")
    worst = _grad_worst(model, ref_grads)
    if worst[1] > 0.25:
        msg.append(f"{variant}: grad {worst[0]} rel {worst[1]:.3f}")
    return msg, f"loss {float(loss.detach()):.4f}/{float(ref_loss):.4f} grad {worst[1]:.3f} ({worst[0]})"
