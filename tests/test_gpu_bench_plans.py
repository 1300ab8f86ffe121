"""Parity at the launch plans bench.py times (BASELINE.json configs[1] and configs[2] at their full batch sizes): the
attention kernel's whole-job and several-jobs-per-CTA plans, the chunked logit stage (B*T > 2048 rows), the decode graphs at
256 images and the graph-captured training step.  Rows are independent, so the full batch is decoded on the device and a
fixed subset of the SAME batch is compared with the CPU oracle (sized so the oracle finishes in seconds)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import unpaired_image_captioning_b200 as uic  # noqa: E402
from oracle import decoder_oracle as O  # noqa: E402
from unpaired_image_captioning_b200 import _lib, synth  # noqa: E402
from parity import compare_beam, compare_greedy  # noqa: E402
from test_gpu_kernels import _att_reference, _rand_bf16  # noqa: E402

REL = 1e-3   # north_star tolerance (relative, log-probs / losses / decision margins)
DEV = "cuda"


def _bench_model(cfg_name, seed=1234, **sd_kw):
    opt, cfg = synth.opt_for(cfg_name)
    sd = synth.init_state_dict(opt, seed=seed, **sd_kw)           # the weights bench.py uses
    model = uic.setup(opt)
    model.load_state_dict(sd)
    return opt, cfg, sd, model.cuda()


@pytest.fixture(scope="module")
def cfg2():
    opt, cfg, sd, model = _bench_model("cfg2")
    fc, att = synth.make_features(cfg["batch"], cfg["att_size"], opt.att_feat_size, seed=1234)   # bench.py's rank-0 batch
    return opt, cfg, sd, model.eval(), fc, att


SUBSET = list(range(0, 256, 17))   # 16 images spread over the batch (first and last included)


def test_cfg2_beam3_at_bench_batch(cfg2):
    """configs[1]: att2in2, 256 images x 196 regions, beam 3 -- the plan `value` and `e2e` of bench.py time."""
    opt, cfg, sd, model, fc, att = cfg2
    assert cfg["batch"] == 256 and SUBSET[-1] == 255
    seq, lp = model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": cfg["beam_size"]}, mode="sample")
    idx = torch.tensor(SUBSET)
    ref_seq, ref_lp, ref_done, margins = O.sample_beam(sd, "att2in2", fc[idx], att[idx], opt.seq_length, cfg["beam_size"],
                                                       return_margins=True)
    exact, exempt, failures = compare_beam(seq[idx], ref_seq, margins, tol=REL)
    assert not failures, (failures, exact, exempt)
    assert exact >= 1
    rows = (seq[idx] == ref_seq).all(1)
    torch.testing.assert_close(lp[idx][rows], ref_lp[rows], rtol=REL, atol=10 * REL)
    for j, k in enumerate(SUBSET):                               # scores of the kept hypotheses
        if rows[j]:
            assert abs(model.done_beams[k][0]["p"] - ref_done[j][0]["p"]) <= REL * abs(ref_done[j][0]["p"])
    # a second call replays the captured graph on the same inputs
    seq2, lp2 = model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": cfg["beam_size"]}, mode="sample")
    assert torch.equal(seq, seq2) and torch.equal(lp, lp2)


def test_cfg2_greedy_at_bench_batch(cfg2):
    opt, cfg, sd, model, fc, att = cfg2
    seq, lp = model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": 1}, mode="sample")
    idx = torch.tensor(SUBSET)
    ref_seq, ref_lp, margins = O.sample_greedy(sd, "att2in2", fc[idx], att[idx], opt.seq_length, return_margins=True,
                                               relative_margins=True)
    exact, exempt, failures = compare_greedy(seq.cpu()[idx], ref_seq, margins, tol=REL)
    assert not failures, (failures, exact, exempt)
    assert exact >= 1
    rows = (seq.cpu()[idx] == ref_seq).all(1)
    torch.testing.assert_close(lp.cpu()[idx][rows], ref_lp[rows], rtol=REL, atol=10 * REL)


def test_cfg2_teacher_forced_at_bench_batch(cfg2):
    """256 rows x 17 steps = 4352 rows of logits: three chunks of the logit stage; 256 whole jobs in the attention kernel."""
    opt, cfg, sd, model, fc, att = cfg2
    labels, masks = synth.make_captions(cfg["batch"], opt.seq_length, opt.vocab_size, seed=1234)
    with torch.no_grad():
        out = model(fc.cuda(), None, att.cuda(), labels.cuda())
    idx = torch.tensor(SUBSET)
    ref = O.teacher_forced(sd, "att2in2", fc[idx], att[idx], labels[idx])
    sel = masks[idx][:, 1:].bool()
    rel = ((out.cpu()[idx] - ref).abs() / ref.abs().clamp_min(1.0))[sel]
    assert float(rel.max()) < REL, float(rel.max())


def _grad_errors(model, ref_grads):
    errs = {}
    for name, p in model.named_parameters():
        r = ref_grads[name]
        if float(r.abs().max()) < 1e-9:
            continue
        errs[name] = float((p.grad.detach().cpu() - r).norm() / r.norm())
    return errs


def test_cfg3_train_step_at_bench_batch():
    """configs[2]: TopDown, 512 rows x 36 regions: fused loss + every parameter gradient against the oracle's autograd on the
    WHOLE batch (512 x 17 = 8704 logit rows: five chunks; 512 jobs on 296 attention CTAs)."""
    opt, cfg, sd, model = _bench_model("cfg3")
    model.train()
    B = cfg["batch"]
    assert B == 512
    fc, att = synth.make_features(B, cfg["att_size"], opt.att_feat_size, seed=4321)
    labels, masks = synth.make_captions(B, opt.seq_length, opt.vocab_size, seed=4321)
    ref_loss, ref_grads = O.loss_and_grads(sd, "topdown", fc, att, labels, masks)
    loss = model(fc.cuda(), None, att.cuda(), labels.cuda(), masks.cuda(), None, mode="forward_loss")
    loss.backward()
    assert abs(float(loss.detach()) - float(ref_loss)) <= REL * abs(float(ref_loss))
    errs = _grad_errors(model, ref_grads)
    worst = max(errs.items(), key=lambda kv: kv[1])
    assert worst[1] <= 5e-2, sorted(errs.items(), key=lambda kv: -kv[1])[:5]    # relative Frobenius error per parameter


def test_cfg3_graphed_train_step_matches_eager():
    """The graph-captured step bench.py times (zero grads, fwd + loss, BPTT, clip, Adam in ONE CUDA graph) produces the
    same loss trajectory and parameters as the same step issued eagerly."""
    from unpaired_image_captioning_b200 import train_bench

    def fresh():
        return train_bench.make_state(0, 0, 1, rows=128)          # (four steps each way: keep it short)

    st_e = fresh()
    eager = [float(train_bench.one_train_step(st=st_e)) for _ in range(3 + 3)]   # the graphed step warms up 3x (the capture itself executes nothing)
    st_g = fresh()
    step = train_bench.GraphedTrainStep(st_g)
    graphed = [float(step()) for _ in range(3)]
    torch.testing.assert_close(torch.tensor(graphed), torch.tensor(eager[3:]), rtol=2e-3, atol=2e-3)
    assert eager[-1] < eager[0]                                    # and the loss goes down
    for (n, a), (_, b) in zip(st_e["model"].named_parameters(), st_g["model"].named_parameters()):
        assert float((a - b).norm() / a.norm().clamp_min(1e-6)) < 2e-2, n


@pytest.mark.parametrize("n_img,beams,L", [(148, 1, 196), (148, 3, 196), (256, 3, 196), (256, 1, 196), (300, 3, 100), (300, 1, 36),
                                           (512, 1, 36), (512, 3, 36), (1000, 1, 36), (37, 3, 196), (296, 2, 50)])
def test_att_step_fwd_launch_plans(n_img, beams, L):
    """uic_att_step_fwd at job counts around the CTA-slot count (296): one whole job per CTA, fewer jobs than slots, more jobs
    than slots (several jobs per CTA, att_h buffer rotation), and the segmented plan for comparison."""
    A = H = 512
    R = n_img * beams
    p_att = _rand_bf16(n_img, L, A, seed=21)
    att = _rand_bf16(n_img, L, H, seed=22).abs()
    att_h = torch.randn(R, A, device=DEV)
    w = torch.randn(A, device=DEV) * 0.2
    ctx_f = torch.empty(R, H, device=DEV)
    alpha = torch.empty(R, L, device=DEV)
    e_tile = _lib.exp_tile(p_att)
    f = (torch.exp(2.0 * att_h) * _lib.ATT_F_SCALE).contiguous()
    for _ in range(2):
        ctx_f.zero_()
        _lib.att_step(f, A, e_tile, att, w, None, None, 0, ctx_f, H, alpha, n_img, beams, L, A, H)
    p_eff = _lib.tile_value(e_tile)
    for r0 in range(0, n_img, 64):                                  # reference in slabs (the tanh tensor is R x L x A fp32)
        r1 = min(n_img, r0 + 64)
        ref_ctx, ref_alpha = _att_reference(att_h[r0 * beams:r1 * beams], p_eff[r0:r1], att[r0:r1], w, None, beams)
        torch.testing.assert_close(alpha[r0 * beams:r1 * beams], ref_alpha, rtol=5e-3, atol=2e-5)
        torch.testing.assert_close(ctx_f[r0 * beams:r1 * beams], ref_ctx, rtol=5e-3, atol=5e-4)


@pytest.mark.parametrize("n_img,beams,L,H", [(150, 5, 52, 1024), (150, 4, 36, 1024), (500, 5, 196, 1024), (80, 5, 100, 768), (150, 7, 36, 1024),
                                             (60, 5, 36, 1024)])
def test_att_step_fwd_wide_beam_groups(n_img, beams, L, H):
    """configs[4] shapes (rnn 1024, beam 5): four or five beams of an image share ONE pass over its tiles in the v7 kernel
    (att_v7_beam_cap); 7 beams -> groups of 4 + 3; 60 images -> below the v7 threshold, v6 with groups of <= 3."""
    A = 512
    R = n_img * beams
    p_att = _rand_bf16(n_img, L, A, seed=31)
    att = _rand_bf16(n_img, L, H, seed=32).abs()
    att_h = torch.randn(R, A, device=DEV)
    w = torch.randn(A, device=DEV) * 0.2
    ctx_f = torch.empty(R, H, device=DEV)
    ctx_b = torch.empty(R, H, device=DEV, dtype=torch.bfloat16)
    alpha = torch.empty(R, L, device=DEV)
    e_tile = _lib.exp_tile(p_att)
    f = (torch.exp(2.0 * att_h) * _lib.ATT_F_SCALE).contiguous()
    for _ in range(2):
        ctx_f.zero_()
        _lib.att_step(f, A, e_tile, att, w, None, None, 0, ctx_f, H, alpha, n_img, beams, L, A, H)
        _lib.att_step(f, A, e_tile, att, w, None, ctx_b, H, None, 0, None, n_img, beams, L, A, H)
    p_eff = _lib.tile_value(e_tile)
    for r0 in range(0, n_img, 32):
        r1 = min(n_img, r0 + 32)
        ref_ctx, ref_alpha = _att_reference(att_h[r0 * beams:r1 * beams], p_eff[r0:r1], att[r0:r1], w, None, beams)
        torch.testing.assert_close(alpha[r0 * beams:r1 * beams], ref_alpha, rtol=5e-3, atol=2e-5)
        torch.testing.assert_close(ctx_f[r0 * beams:r1 * beams], ref_ctx, rtol=5e-3, atol=5e-4)
        torch.testing.assert_close(ctx_b[r0 * beams:r1 * beams].float(), ref_ctx, rtol=2e-2, atol=5e-3)


@pytest.mark.parametrize("n_img,beams,L,H", [(200, 2, 52, 512), (160, 3, 100, 512), (150, 5, 40, 1024), (300, 1, 36, 512)])
def test_att_step_fwd_v7_with_region_masks(n_img, beams, L, H):
    """The batch-balanced kernel with ragged region masks and the alpha output (its AUX instantiations): masked regions get
    zero weight and the rest is renormalised (models/AttModel.py:552-555), also for jobs cut between two CTAs."""
    A = 512
    R = n_img * beams
    g = torch.Generator().manual_seed(41)
    lens = torch.randint(1, L + 1, (n_img,), generator=g)
    lens[0], lens[-1] = L, 1
    masks = (torch.arange(L)[None, :] < lens[:, None]).float().to(DEV)
    p_att = _rand_bf16(n_img, L, A, seed=42)
    att = _rand_bf16(n_img, L, H, seed=43).abs()
    att_h = torch.randn(R, A, device=DEV)
    w = torch.randn(A, device=DEV) * 0.2
    ctx_f = torch.empty(R, H, device=DEV)
    alpha = torch.empty(R, L, device=DEV)
    e_tile = _lib.exp_tile(p_att)
    f = (torch.exp(2.0 * att_h) * _lib.ATT_F_SCALE).contiguous()
    for _ in range(2):
        _lib.att_step(f, A, e_tile, att, w, masks, None, 0, ctx_f, H, alpha, n_img, beams, L, A, H)
    p_eff = _lib.tile_value(e_tile)
    for r0 in range(0, n_img, 50):
        r1 = min(n_img, r0 + 50)
        ref_ctx, ref_alpha = _att_reference(att_h[r0 * beams:r1 * beams], p_eff[r0:r1], att[r0:r1], w, masks[r0:r1], beams)
        torch.testing.assert_close(alpha[r0 * beams:r1 * beams], ref_alpha, rtol=5e-3, atol=2e-5)
        torch.testing.assert_close(ctx_f[r0 * beams:r1 * beams], ref_ctx, rtol=5e-3, atol=5e-4)


def test_cfg5_beam5_single_pass_plan_matches_oracle():
    """configs[4] layer sizes (rnn 1024, vocab 30k, beam 5, 20 steps) at a batch that takes the plan `bench.py --workload cfg5`
    times: >= 74 images, so the five beams of an image share ONE pass of the batch-balanced attention kernel; the oracle
    decodes three images of the same batch (first, middle, last)."""
    opt = synth.make_opt(caption_model="att2in2", vocab_size=29999, rnn_size=1024, input_encoding_size=512, att_hid_size=512,
                         seq_length=20)
    sd = synth.init_state_dict(opt, seed=9, peaked=40.0, eos_bias=2.0)
    B = 96
    fc, att = synth.make_features(B, 196, 2048, seed=9)
    model = uic.setup(opt)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    seq, lp = model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": 5}, mode="sample")
    idx = torch.tensor([0, B // 2, B - 1])
    ref_seq, ref_lp, _, margins = O.sample_beam(sd, "att2in2", fc[idx], att[idx], 20, 5, return_margins=True)
    exact, exempt, failures = compare_beam(seq[idx], ref_seq, margins, tol=REL)
    assert not failures, (failures, seq[idx], ref_seq)
    rows = (seq[idx] == ref_seq).all(1)
    torch.testing.assert_close(lp[idx][rows], ref_lp[rows], rtol=2e-2, atol=2e-2)

