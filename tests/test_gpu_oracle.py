"""CUDA path vs the CPU oracle on seeded synthetic inputs at the real layer sizes of configs 1-3
(rnn 512, vocab 10k, 196 / 36 regions), sized so the oracle finishes in seconds."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import unpaired_image_captioning_b200 as uic  # noqa: E402
from oracle import decoder_oracle as O  # noqa: E402
from unpaired_image_captioning_b200 import synth  # noqa: E402
from parity import compare_beam, compare_greedy  # noqa: E402

REL = 1e-3


def _case(kind, B, L, seed, peaked=0.0, eos_bias=0.0, masks=False):
    opt = synth.make_opt(caption_model=kind, vocab_size=9999, rnn_size=512, input_encoding_size=512, att_hid_size=512,
                         seq_length=16)
    sd = synth.init_state_dict(opt, seed=seed, peaked=peaked, eos_bias=eos_bias)
    fc, att = synth.make_features(B, L, 2048, seed=seed)
    labels, lmasks = synth.make_captions(B, 16, 9999, seed=seed)
    am = synth.make_att_masks(B, L, seed=seed) if masks else None
    model = uic.setup(opt)
    model.load_state_dict(sd)
    return opt, sd, model.cuda().eval(), fc, att, labels, lmasks, am


@pytest.mark.parametrize("kind,L,masks", [("att2in2", 196, False), ("att2all2", 100, True), ("topdown", 36, False), ("topdown", 36, True),
                                          ("stackatt", 36, False), ("denseatt", 49, True)])
def test_teacher_forced_and_loss(kind, L, masks):
    opt, sd, model, fc, att, labels, lmasks, am = _case(kind, 8, L, seed=1234, masks=masks)
    ref = O.teacher_forced(sd, kind, fc, att, labels, am)
    ref_loss = O.xe_loss(ref, labels[:, 1:], lmasks[:, 1:])
    cu = lambda t: None if t is None else t.cuda()
    with torch.no_grad():
        out = model(cu(fc), None, cu(att), cu(labels), cu(am))
        loss = uic.LanguageModelCriterion(opt)(out, cu(labels)[:, 1:], cu(lmasks)[:, 1:])
    sel = lmasks[:, 1:].bool()
    rel = ((out.cpu() - ref).abs() / ref.abs().clamp_min(1.0))[sel]
    assert float(rel.max()) < REL, float(rel.max())
    assert abs(float(loss) - float(ref_loss)) < REL * float(ref_loss)


@pytest.mark.parametrize("kind,shift", [("att2in2", 8.0), ("topdown", 8.0), ("att2in2", -9.0), ("topdown", 15.0)])
def test_attention_operand_range(kind, shift):
    """p_att far from zero that att_h cancels (ctx2att.bias + shift, h2att.bias - shift on half of the units): the sum
    p_att + att_h is unchanged, so the log-probs must still match the oracle at the north-star tolerance.  The first
    (fp16) operand tile saturated for p_att > 6.9 and flushed below -8.3; the bf16 tile covers |p_att| < 20.8."""
    opt, sd, model, fc, att, labels, lmasks, am = _case(kind, 6, 36, seed=11)
    sd = {k: v.clone() for k, v in sd.items()}
    half = sd["ctx2att.bias"].numel() // 2
    sd["ctx2att.bias"][:half] += shift
    sd["core.attention.h2att.bias"][:half] -= shift
    model.load_state_dict(sd)
    ref = O.teacher_forced(sd, kind, fc, att, labels)
    with torch.no_grad():
        out = model(fc.cuda(), None, att.cuda(), labels.cuda())
    sel = lmasks[:, 1:].bool()
    rel = ((out.cpu() - ref).abs() / ref.abs().clamp_min(1.0))[sel]
    assert float(rel.max()) < REL, float(rel.max())
    # and the gradients flow through the same tile (BPTT recomputes tanh from it)
    model.train()
    ref_loss, ref_grads = O.loss_and_grads(sd, kind, fc, att, labels, lmasks)
    loss = model(fc.cuda(), None, att.cuda(), labels.cuda(), lmasks.cuda(), None, mode="forward_loss")
    loss.backward()
    assert abs(float(loss.detach()) - float(ref_loss)) < REL * float(ref_loss)
    for name in ("ctx2att.weight", "core.attention.h2att.weight", "core.attention.alpha_net.weight"):
        g, r = dict(model.named_parameters())[name].grad.cpu(), ref_grads[name]
        assert float((g - r).norm() / r.norm()) < 5e-2, name


@pytest.mark.parametrize("kind,L", [("att2in2", 196), ("att2all2", 64), ("topdown", 36), ("stackatt", 36), ("denseatt", 36)])
def test_greedy_with_margin_exemption(kind, L):
    opt, sd, model, fc, att, *_ = _case(kind, 16, L, seed=77)
    ref_seq, ref_lp, margins = O.sample_greedy(sd, kind, fc, att, 16, return_margins=True, relative_margins=True)
    seq, lp = model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": 1}, mode="sample")
    exact, exempt, failures = compare_greedy(seq.cpu(), ref_seq, margins, tol=REL)
    assert not failures, failures
    assert exact >= 1
    first = (seq.cpu() == ref_seq).all(1)
    torch.testing.assert_close(lp.cpu()[first], ref_lp[first], rtol=REL, atol=REL * 10)


@pytest.mark.parametrize("kind,L,beam", [("att2in2", 196, 3), ("topdown", 36, 3), ("att2in2", 49, 5), ("att2all2", 49, 3),
                                         ("stackatt", 36, 3), ("denseatt", 49, 3), ("denseatt", 36, 5)])
def test_beam_peaked_exact(kind, L, beam):
    """Wide-margin variant (scaled logit weights, raised EOS bias): ids must be identical -- a differing row must sit at an
    oracle decision margin inside the north-star tolerance."""
    opt, sd, model, fc, att, *_ = _case(kind, 6, L, seed=5, peaked=40.0, eos_bias=2.0)
    ref_seq, ref_lp, ref_done, margins = O.sample_beam(sd, kind, fc, att, 16, beam, return_margins=True)
    seq, lp = model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": beam}, mode="sample")
    exact, exempt, failures = compare_beam(seq, ref_seq, margins, tol=REL)
    assert not failures, (failures, seq, ref_seq)
    rows = (seq == ref_seq).all(1)
    torch.testing.assert_close(lp[rows], ref_lp[rows], rtol=2e-2, atol=2e-2)


def test_config5_layer_sizes_beam5_and_greedy():
    """BASELINE.json configs[4] at a reduced batch: att2in2 with rnn 1024, vocab 30k, 20 steps, beam 5 (the wide
    attention / H > 512 kernel variants, two beam groups per image, kslots = 5 statistics)."""
    opt = synth.make_opt(caption_model="att2in2", vocab_size=29999, rnn_size=1024, input_encoding_size=512, att_hid_size=512,
                         seq_length=20)
    sd = synth.init_state_dict(opt, seed=9, peaked=40.0, eos_bias=2.0)
    fc, att = synth.make_features(3, 196, 2048, seed=9)
    model = uic.setup(opt)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    ref_seq, ref_lp, _, bmargins = O.sample_beam(sd, "att2in2", fc, att, 20, 5, return_margins=True)
    seq, lp = model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": 5}, mode="sample")
    exact, exempt, failures = compare_beam(seq, ref_seq, bmargins, tol=REL)
    assert not failures, (failures, seq, ref_seq)
    rows = (seq == ref_seq).all(1)
    torch.testing.assert_close(lp[rows], ref_lp[rows], rtol=2e-2, atol=2e-2)
    g_ref, g_lp, margins = O.sample_greedy(sd, "att2in2", fc, att, 20, return_margins=True, relative_margins=True)
    g_seq, g_lpc = model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": 1}, mode="sample")
    exact, exempt, failures = compare_greedy(g_seq.cpu(), g_ref, margins, tol=REL)
    assert not failures, failures
    assert exact >= 1


@pytest.mark.parametrize("kind,over", [("att2in2", dict(use_bn=2)), ("topdown", dict(logit_layers=2)),
                                       ("denseatt", dict(use_bn=2, logit_layers=3)), ("att2in2", dict(logit_layers=2))])
def test_second_batchnorm_and_hidden_logit_layers(kind, over):
    """use_bn = 2 (BatchNorm1d behind att_embed, models/AttModel.py:84) and logit_layers > 1 (:89-91) in eval mode, at real
    widths: teacher-forced log-probs, greedy and beam-3 against the oracle (itself pinned against the live reference for
    these options, tests/test_oracle_vs_reference.py).  Training calls of these configurations fail loudly."""
    opt = synth.make_opt(caption_model=kind, vocab_size=9999, rnn_size=512, input_encoding_size=512, att_hid_size=512, seq_length=16, **over)
    sd = synth.init_state_dict(opt, seed=21)
    B, L = 6, 36
    fc, att = synth.make_features(B, L, 2048, seed=21)
    labels, lmasks = synth.make_captions(B, 16, 9999, seed=21)
    am = synth.make_att_masks(B, L, seed=21) if opt.use_bn else None
    model = uic.setup(opt)
    assert set(model.state_dict().keys()) == set(sd.keys())
    model.load_state_dict(sd)
    model = model.cuda().eval()
    cu = lambda t: None if t is None else t.cuda()
    ref = O.teacher_forced(sd, kind, fc, att, labels, am)
    with torch.no_grad():
        out = model(cu(fc), None, cu(att), cu(labels), cu(am))
    sel = lmasks[:, 1:].bool()
    rel = ((out.cpu() - ref).abs() / ref.abs().clamp_min(1.0))[sel]
    assert float(rel.max()) < REL, float(rel.max())
    ref_seq, _, margins = O.sample_greedy(sd, kind, fc, att, 16, am, return_margins=True, relative_margins=True)
    seq, _ = model(cu(fc), None, cu(att), cu(am), opt={"beam_size": 1}, mode="sample")
    assert not compare_greedy(seq.cpu(), ref_seq, margins, tol=REL)[2]
    b_ref, _, _, b_margins = O.sample_beam(sd, kind, fc, att, 16, 3, am, return_margins=True)
    b_seq, _ = model(cu(fc), None, cu(att), cu(am), opt={"beam_size": 3}, mode="sample")
    assert not compare_beam(b_seq, b_ref, b_margins, tol=REL)[2]
    model.train()
    with pytest.raises(NotImplementedError):
        model(cu(fc), None, cu(att), cu(labels), cu(lmasks), cu(am), mode="forward_loss")


@pytest.mark.parametrize("kind", ["att2in2", "att2all2"])
def test_gate_table_decode_equals_the_full_gemm(kind):
    """Decode loops take the input-word gate term from the (V, 5H) table (engine.use_gate_table): same captions and
    log-probs as contracting the E embedding columns in every step's GEMM (two fp32 sums of the same bf16 products)."""
    opt, sd, model, fc, att, *_ = _case(kind, 7, 49, seed=5, peaked=40.0, eos_bias=2.0)
    outs = {}
    for use in (True, False):
        model.engine.use_gate_table = use
        outs[use] = [model(fc.cuda(), None, att.cuda(), None, opt=o, mode="sample")
                     for o in ({"beam_size": 3}, {"beam_size": 1}, {"beam_size": 4, "group_size": 2})]
    model.engine.use_gate_table = True
    for (s_a, lp_a), (s_b, lp_b) in zip(outs[True], outs[False]):
        assert torch.equal(s_a.cpu(), s_b.cpu())
        # (x40 logit weights: an fp32 last-bit difference in a gate sum that flips the bf16 rounding of one h element moves a
        #  log-prob by ~1e-3; the tolerance of the other wide-margin tests)
        torch.testing.assert_close(lp_a.cpu(), lp_b.cpu(), rtol=2e-2, atol=2e-2)


def test_bf16_feature_cache_is_consumed_directly():
    """Features handed over as bf16 (a host/device feature cache in the operand precision) give the same captions as the
    fp32 tensors they were rounded from, up to that rounding."""
    opt, sd, model, fc, att, *_ = _case("att2in2", 6, 49, seed=5, peaked=40.0, eos_bias=2.0)
    att_bf = att.cuda().to(torch.bfloat16)
    seq_a, lp_a = model(fc.cuda(), None, att_bf.float(), None, opt={"beam_size": 3}, mode="sample")
    seq_b, lp_b = model(fc.cuda(), None, att_bf, None, opt={"beam_size": 3}, mode="sample")
    assert torch.equal(seq_a, seq_b)
    torch.testing.assert_close(lp_a, lp_b, rtol=0, atol=0)


def test_beam10_default_and_unfused_paths():
    """beam_size 10 is the reference's default (models/AttModel.py:170) and runs through the logit-materialising kernels
    (the fused statistics keep at most 8 candidates per part); the same kernels must also reproduce the fused path's
    beam-3 and greedy results when the fusion is switched off."""
    opt, sd, model, fc, att, *_ = _case("att2in2", 5, 49, seed=5, peaked=40.0, eos_bias=2.0)
    ref_seq, ref_lp, _, margins = O.sample_beam(sd, "att2in2", fc, att, 16, 10, return_margins=True)
    seq, lp = model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": 10}, mode="sample")
    exact, exempt, failures = compare_beam(seq, ref_seq, margins, tol=REL)
    assert not failures, (failures, seq, ref_seq)
    rows = (seq == ref_seq).all(1)
    torch.testing.assert_close(lp[rows], ref_lp[rows], rtol=2e-2, atol=2e-2)
    assert len(model.done_beams[0]) <= 10
    fused = [model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": b}, mode="sample") for b in (3, 1)]
    model.engine.fused_vocab = False
    try:
        plain = [model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": b}, mode="sample") for b in (3, 1)]
    finally:
        model.engine.fused_vocab = True
    for (s_f, lp_f), (s_p, lp_p) in zip(fused, plain):
        assert torch.equal(s_f.cpu(), s_p.cpu())
        torch.testing.assert_close(lp_f.cpu(), lp_p.cpu(), rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("kind,beam,groups,lam,dc,ppl", [("att2in2", 6, 3, 0.5, 0, 0), ("topdown", 4, 2, 1.5, 1, 1),
                                                          ("att2in2", 4, 4, 0.5, 0, 0)])
def test_diverse_beam_search(kind, beam, groups, lam, dc, ppl):
    """group_size > 1 (models/CaptionModel.py:36-45,100-177): groups run staggered, later groups are penalised for
    repeating the earlier groups' tokens; done_beams is the groups' lists one after the other."""
    opt, sd, model, fc, att, *_ = _case(kind, 5, 36, seed=7, peaked=40.0, eos_bias=2.0)
    ref_seq, ref_lp, ref_done, bmargins = O.sample_beam(sd, kind, fc, att, 16, beam, decoding_constraint=dc, max_ppl=ppl,
                                                        group_size=groups, diversity_lambda=lam, return_margins=True)
    o = {"beam_size": beam, "group_size": groups, "diversity_lambda": lam, "decoding_constraint": dc, "max_ppl": ppl}
    seq, lp = model(fc.cuda(), None, att.cuda(), None, opt=o, mode="sample")
    rows = (seq == ref_seq).all(1)
    if beam == groups:
        # one beam per group: the first group IS greedy decoding, which cannot recover from a flipped near-tie the way a wider
        # beam does -- exempt near-ties like the greedy test, and require identity with the device's own greedy path
        g_ref, _, margins = O.sample_greedy(sd, kind, fc, att, 16, return_margins=True, relative_margins=True)
        assert torch.equal(g_ref, ref_seq)
        exact, exempt, failures = compare_greedy(seq.cpu(), ref_seq, margins, tol=REL)
        assert not failures, failures
        g_seq, _ = model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": 1}, mode="sample")
        assert torch.equal(g_seq.cpu(), seq.cpu())
    else:
        exact, exempt, failures = compare_beam(seq, ref_seq, bmargins, tol=REL)
        assert not failures, (failures, seq, ref_seq)
    torch.testing.assert_close(lp[rows], ref_lp[rows], rtol=2e-2, atol=2e-2)
    same = total = 0
    for k in range(5):
        mine, ref = model.done_beams[k], ref_done[k]
        assert len(mine) == len(ref)
        for a, b in zip(mine, ref):
            total += 1
            if torch.equal(a["seq"], b["seq"]):
                same += 1
                assert abs(a["p"] - float(b["p"])) <= 2e-2 * max(1.0, abs(float(b["p"])))
                assert abs(a["unaug_p"] - float(b["unaug_p"])) <= 2e-2 * max(1.0, abs(float(b["unaug_p"])))
    # (hypotheses of an image whose search took a different branch at an exempted near-tie differ as a whole list)
    assert same >= int(rows.sum()) * len(ref_done[0]) * 0.5, (same, total)
    # a replay of the captured loop gives the same tables
    seq2, lp2 = model(fc.cuda(), None, att.cuda(), None, opt=o, mode="sample")
    assert torch.equal(seq, seq2) and torch.equal(lp, lp2)


def test_decode_graph_cache_is_bounded():
    """Masked batches are clipped to their longest region count (AttModel.py:99-105), so every new length is a new decode
    graph with its own static tile buffers: the engine keeps only the most recently used few."""
    opt, sd, model, fc, att, *_ = _case("att2in2", 4, 24, seed=5, peaked=40.0, eos_bias=2.0)
    eng = model.engine
    eng.max_graphs = 3
    first = None
    for keep in (24, 20, 16, 12, 8, 24):
        masks = torch.zeros(4, 24)
        masks[:, :keep] = 1.0
        seq, _ = model(fc.cuda(), None, att.cuda(), masks.cuda(), opt={"beam_size": 1}, mode="sample")
        if keep == 24:
            first = seq.clone() if first is None else first
            assert torch.equal(seq, first)            # an evicted shape is simply captured again
        assert len(eng._graphs) <= 3
