"""Per-kernel parity on the B200: each C-ABI entry point against a plain fp32 PyTorch evaluation of
the same formula on the same (bf16-rounded) operands.  All calls go through libuic_b200.so."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from unpaired_image_captioning_b200 import _lib  # noqa: E402
from unpaired_image_captioning_b200._lib import check, ptr, stream  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _device():
    _lib.require_device()
    yield
    _lib.load().uic_set_gemm_impl(0)


def _rand_bf16(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(torch.bfloat16)


# ---------------------------------------------------------------------------------------------------
# GEMM (tcgen05) -- shapes cover full tiles, M/N/K tails, the per-step and prologue sizes
# ---------------------------------------------------------------------------------------------------
GEMM_SHAPES = [(128, 128, 64), (128, 128, 512), (256, 384, 192), (5, 52, 32), (35, 52, 64), (300, 200, 72), (2048, 2040, 328),
               (768, 3072, 1024), (256, 10000, 512), (16 * 17, 10000, 512), (6272, 512, 2048), (48, 2560, 1024)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_tn_matches_fp32(M, N, K):
    a, b = _rand_bf16(M, K, seed=1), _rand_bf16(N, K, seed=2, scale=0.05)
    bias = torch.randn(N, device=DEV)
    out = torch.empty(M, N, device=DEV)
    _lib.gemm(a, b, bias, out_f32=out)
    ref = a.float() @ b.float().t() + bias
    torch.testing.assert_close(out, ref, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 200, 72), (768, 3072, 1024)])
def test_gemm_tcgen05_equals_simt_kernel(M, N, K):
    a, b = _rand_bf16(M, K, seed=3), _rand_bf16(N, K, seed=4, scale=0.05)
    lib = _lib.load()
    o_tc, o_simt = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    _lib.gemm(a, b, out_f32=o_tc)
    check(lib.uic_set_gemm_impl(1))
    try:
        _lib.gemm(a, b, out_f32=o_simt)
    finally:
        check(lib.uic_set_gemm_impl(0))
    torch.testing.assert_close(o_tc, o_simt, rtol=1e-4, atol=1e-4)


def test_gemm_epilogues_and_strided_operands():
    M, N, K = 200, 328, 256
    big_a = _rand_bf16(M, K + 128, seed=5)
    a = big_a[:, 64:64 + K]                      # column slice of a wider activation matrix (pitch K+128)
    b = _rand_bf16(N, K, seed=6, scale=0.05)
    bias = torch.randn(N, device=DEV)
    ref = a.float() @ b.float().t() + bias
    out_f, out_b = torch.empty(M, N, device=DEV), torch.empty(M, N + 8, device=DEV, dtype=torch.bfloat16)[:, :N]
    _lib.gemm(a, b, bias, out_f32=out_f, out_bf16=out_b, relu=True)
    torch.testing.assert_close(out_f, ref.relu(), rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(out_b.float(), ref.relu(), rtol=1e-2, atol=1e-2)
    acc = torch.full((M, N), 0.5, device=DEV)
    _lib.gemm(a, b, None, out_f32=acc, accumulate=True)
    torch.testing.assert_close(acc, ref - bias + 0.5, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 328, 136), (512, 2560, 8704 // 8), (2048, 2048, 520), (1104, 1000, 264)])
def test_gemm_mn_major_operands(a_mn, b_mn, M, N, K):
    """dgrad (B stored [K,N]) and wgrad (A stored [K,M], B stored [K,N]) forms."""
    a, b = _rand_bf16(M, K, seed=7), _rand_bf16(N, K, seed=8, scale=0.05)
    ref = a.float() @ b.float().t()
    a_arg = a.t().contiguous() if a_mn else a
    b_arg = b.t().contiguous() if b_mn else b
    out = torch.empty(M, N, device=DEV)
    _lib.gemm(a_arg, b_arg, out_f32=out, a_mn=a_mn, b_mn=b_mn)
    torch.testing.assert_close(out, ref, rtol=2e-4, atol=2e-4)


def test_gemm_exponential_attention_epilogue():
    """Columns >= exp_col0 come out as scale * exp(2x), capped at 2^60 (65504 for an fp16 output buffer)."""
    M, N, K = 70, 96, 64
    a, b = _rand_bf16(M, K, seed=31, scale=0.3), _rand_bf16(N, K, seed=32, scale=0.3)
    bias = torch.randn(N, device=DEV) * 0.1
    ref = a.float() @ b.float().t() + bias
    out = torch.empty(M, N, device=DEV)
    _lib.gemm(a, b, bias, out_f32=out, exp_col0=40, exp_scale=16.0)
    torch.testing.assert_close(out[:, :40], ref[:, :40], rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(out[:, 40:], 16.0 * torch.exp(2 * ref[:, 40:]), rtol=2e-3, atol=1e-4)
    out16 = torch.empty(M, N, device=DEV, dtype=torch.float16)
    _lib.gemm(a, b, bias + 6.0, out_bf16=out16, exp_col0=0, exp_scale=1 / 16)
    want = (torch.exp(2 * (ref + 6.0)) / 16).clamp(max=65504.0)
    torch.testing.assert_close(out16.float(), want, rtol=3e-3, atol=1e-3)
    assert torch.isfinite(out16.float()).all()
    # the attention operand tile: bf16, E = exp(2 p) for |p| far outside the old fp16 window, capped at 2^60
    tile = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    for shift in (-15.0, 12.0, 30.0):
        _lib.gemm(a, b, bias + shift, out_bf16=tile, exp_col0=0, exp_scale=_lib.ATT_E_SCALE)
        want = torch.exp(2 * (ref.double() + shift)).clamp(max=2.0 ** 60).float()
        torch.testing.assert_close(tile.float(), want, rtol=6e-3, atol=0)
        assert torch.isfinite(tile.float()).all() and bool((tile.float() > 0).all())


def test_gemm_rejects_bad_arguments():
    a, b = _rand_bf16(16, 24), _rand_bf16(8, 24)
    with pytest.raises(ValueError):
        _lib.gemm(a.float(), b, out_f32=torch.empty(16, 8, device=DEV))
    with pytest.raises(_lib.UicError):  # pitch 20 elements = 40 bytes is not TMA-addressable
        _lib.gemm(_rand_bf16(16, 20), _rand_bf16(8, 20), out_f32=torch.empty(16, 8, device=DEV))


# ---------------------------------------------------------------------------------------------------
# fused attention step
# ---------------------------------------------------------------------------------------------------
def _att_reference(att_h, p_att, att, w, masks, beams):
    B, L, A = p_att.shape
    p = p_att.float().repeat_interleave(beams, 0)
    a = att.float().repeat_interleave(beams, 0)
    e = torch.tanh(p + att_h[:, None, :]) @ w
    alpha = torch.softmax(e, 1)
    if masks is not None:
        m = masks.repeat_interleave(beams, 0)
        alpha = alpha * m
        alpha = alpha / alpha.sum(1, keepdim=True)
    return torch.bmm(alpha[:, None, :], a).squeeze(1), alpha


@pytest.mark.parametrize("B,beams,L,A,H,use_masks", [
    (5, 1, 7, 32, 32, False), (5, 3, 7, 32, 32, True), (4, 5, 36, 512, 512, False), (3, 2, 196, 512, 512, True),
    (2, 3, 196, 512, 1024, False), (3, 10, 20, 64, 40, False), (16, 1, 196, 512, 512, False), (2, 1, 5, 264, 776, True),
    (3, 3, 196, 512, 512, True), (2, 1, 400, 256, 512, False), (4, 1, 3, 32, 64, False), (2, 5, 600, 512, 512, True)])
def test_att_step_fwd(B, beams, L, A, H, use_masks):
    lib = _lib.load()
    R = B * beams
    p_att, att = _rand_bf16(B, L, A, seed=11), _rand_bf16(B, L, H, seed=12).abs()
    att_h_full = torch.randn(R, A + 24, device=DEV)   # pitch larger than A, like the fused gate GEMM output
    att_h = att_h_full[:, 8:8 + A]
    w = torch.randn(A, device=DEV) * 0.2
    masks = None
    if use_masks:
        n = torch.randint(1, L + 1, (B,))
        masks = (torch.arange(L)[None, :] < n[:, None]).float().to(DEV).contiguous()
    ctx_b = torch.empty(R, H, device=DEV, dtype=torch.bfloat16)
    ctx_f = torch.empty(R, H, device=DEV)
    alpha = torch.empty(R, L, device=DEV)
    # operands in the exponential form the GEMM epilogues produce: E = exp(2 p) (bf16 tile), F = exp(2 att_h)
    e_tile = _lib.exp_tile(p_att)
    f_full = (torch.exp(2.0 * att_h_full) * _lib.ATT_F_SCALE).contiguous()
    f_view = f_full[:, 8:8 + A]
    for _ in range(2):   # twice: the split-merge arrival counters must be left at zero by the kernel
        ctx_f.zero_()
        _lib.att_step(f_view, f_full.stride(0), e_tile, att, w, masks, ctx_b, H, ctx_f, H, alpha, B, beams, L, A, H)
    p_eff = _lib.tile_value(e_tile)                         # the value the bf16 tile actually encodes
    ref_ctx, ref_alpha = _att_reference(att_h, p_eff, att, w, masks, beams)
    torch.testing.assert_close(alpha, ref_alpha, rtol=5e-3, atol=2e-5)     # tanh.approx.f32 inside the score
    torch.testing.assert_close(ctx_f, ref_ctx, rtol=5e-3, atol=5e-4)
    torch.testing.assert_close(ctx_b.float(), ref_ctx, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("scale_p,scale_h", [(4.0, 4.0), (8.0, 8.0), (12.0, 3.0)])
def test_att_step_fwd_wide_operand_range(scale_p, scale_h):
    """|p_att| and |att_h| up to ~3 sigma x scale, including pairs of opposite sign: the bf16 tile and fp32 F keep fp32's
    exponent range, a saturated unit contributes its limit and products that overflow contribute 0, never NaN."""
    B, beams, L, A, H = 3, 3, 50, 512, 512
    R = B * beams
    p_att = (_rand_bf16(B, L, A, seed=41).float() * scale_p).to(torch.bfloat16)
    att = _rand_bf16(B, L, H, seed=42).abs()
    att_h = torch.randn(R, A, device=DEV) * scale_h
    att_h[:, :64] = -p_att.float()[:, 0, :64].repeat_interleave(beams, 0)       # exact cancellation against region 0
    w = torch.randn(A, device=DEV) * 0.2
    e_tile = _lib.exp_tile(p_att)
    f = (torch.exp(2.0 * att_h) * _lib.ATT_F_SCALE).clamp_(max=_lib.ATT_EXP_CAP).contiguous()
    ctx_f, alpha = torch.empty(R, H, device=DEV), torch.empty(R, L, device=DEV)
    _lib.att_step(f, A, e_tile, att, w, None, None, 0, ctx_f, H, alpha, B, beams, L, A, H)
    assert torch.isfinite(ctx_f).all() and torch.isfinite(alpha).all()
    # (the reference sees what the operands encode: both are capped at 2^60, i.e. |p_att|, |att_h| <= 20.8)
    ref_ctx, ref_alpha = _att_reference(0.5 * torch.log(f), _lib.tile_value(e_tile), att, w, None, beams)
    torch.testing.assert_close(alpha, ref_alpha, rtol=5e-3, atol=2e-5)
    torch.testing.assert_close(ctx_f, ref_ctx, rtol=5e-3, atol=5e-4)


# ---------------------------------------------------------------------------------------------------
# LSTM pointwise
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("R,H", [(5, 32), (48, 512), (7, 1024)])
def test_lstm_maxout_and_cell_fwd(R, H):
    lib = _lib.load()
    sums, a2c, c_prev = torch.randn(R, 5 * H, device=DEV), torch.randn(R, 2 * H, device=DEV), torch.randn(R, H, device=DEV)
    c_out, h_f = torch.empty(R, H, device=DEV), torch.empty(R, H, device=DEV)
    X = torch.zeros(R, 3 * H, device=DEV, dtype=torch.bfloat16)
    check(lib.uic_lstm_maxout_fwd(ptr(sums), 5 * H, ptr(a2c), 2 * H, ptr(c_prev), ptr(c_out), ptr(h_f), ptr(X[:, H:]), 3 * H,
                                  ptr(X[:, 2 * H:]), 3 * H, R, H, stream()))
    sig = torch.sigmoid(sums[:, :3 * H])
    g = torch.maximum(sums[:, 3 * H:4 * H] + a2c[:, :H], sums[:, 4 * H:] + a2c[:, H:])
    c_ref = sig[:, H:2 * H] * c_prev + sig[:, :H] * g
    h_ref = sig[:, 2 * H:] * torch.tanh(c_ref)
    torch.testing.assert_close(c_out, c_ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(h_f, h_ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(X[:, H:2 * H].float(), h_ref, rtol=1e-2, atol=1e-2)
    assert torch.equal(X[:, H:2 * H], X[:, 2 * H:]) and float(X[:, :H].abs().max()) == 0.0

    # att2all2 form: no separate context addend (models/AttModel.py:639-648)
    check(lib.uic_lstm_maxout_fwd(ptr(sums), 5 * H, None, 0, ptr(c_prev), ptr(c_out), ptr(h_f), None, 0, None, 0, R, H, stream()))
    c_all = sig[:, H:2 * H] * c_prev + sig[:, :H] * torch.maximum(sums[:, 3 * H:4 * H], sums[:, 4 * H:])
    torch.testing.assert_close(c_out, c_all, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(h_f, sig[:, 2 * H:] * torch.tanh(c_all), rtol=1e-5, atol=1e-5)

    gates = torch.randn(R, 4 * H, device=DEV)
    check(lib.uic_lstm_cell_fwd(ptr(gates), 4 * H, ptr(c_prev), ptr(c_out), ptr(h_f), None, 0, None, 0, R, H, stream()))
    i, f, g, o = gates.chunk(4, 1)
    c_ref = torch.sigmoid(f) * c_prev + torch.sigmoid(i) * torch.tanh(g)
    torch.testing.assert_close(c_out, c_ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(h_f, torch.sigmoid(o) * torch.tanh(c_ref), rtol=1e-5, atol=1e-5)
    # zero initial state through the NULL c_prev form
    check(lib.uic_lstm_cell_fwd(ptr(gates), 4 * H, None, ptr(c_out), ptr(h_f), None, 0, None, 0, R, H, stream()))
    torch.testing.assert_close(c_out, torch.sigmoid(i) * torch.tanh(g), rtol=1e-5, atol=1e-5)


def test_cast_embed_and_zero_padding():
    lib = _lib.load()
    src = torch.randn(37, 72, device=DEV)
    assert torch.equal(_lib.cast_bf16(src), src.to(torch.bfloat16))
    assert torch.equal(_lib.cast_bf16(src, relu=True), src.relu().to(torch.bfloat16))
    odd = torch.randn(5, 13, device=DEV)      # non-vector path
    assert torch.equal(_lib.cast_bf16(odd), odd.to(torch.bfloat16))
    table = _rand_bf16(50, 40, seed=21)
    tok = torch.tensor([0, 49, 7, 7, 3], device=DEV)
    out = torch.zeros(5, 100, device=DEV, dtype=torch.bfloat16)
    check(lib.uic_embed_rows(ptr(table), 40, ptr(tok), ptr(out[:, 16:]), 100, 5, 40, 50, stream()))
    assert torch.equal(out[:, 16:56], table[tok]) and float(out[:, :16].abs().max()) == 0.0
    x = _rand_bf16(3 * 4, 16, seed=22)
    masks = torch.tensor([[1, 1, 1, 1], [1, 0, 0, 0], [1, 1, 0, 0]], device=DEV, dtype=torch.float32)
    ref = x.clone().view(3, 4, 16) * masks[:, :, None].to(torch.bfloat16)
    check(lib.uic_zero_padded_rows(ptr(x), ptr(masks), 3, 4, 16, stream()))
    assert torch.equal(x.view(3, 4, 16), ref)


# ---------------------------------------------------------------------------------------------------
# vocabulary kernels
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("R,V", [(5, 52), (33, 10000), (7, 30001), (4, 1000)])
def test_log_softmax_and_xent(R, V):
    lib = _lib.load()
    ld = V + 3
    buf = torch.randn(R, ld, device=DEV) * 3
    logits = buf[:, :V]
    out = torch.empty(R, V, device=DEV)
    check(lib.uic_log_softmax_rows(ptr(logits), ld, ptr(out), V, R, V, stream()))
    ref = torch.log_softmax(logits, 1)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=2e-5)
    target = torch.randint(0, V, (R,), device=DEV)
    mask = (torch.rand(R, device=DEV) > 0.3).float()
    lse, nll = torch.empty(R, device=DEV), torch.empty(R, device=DEV)
    check(lib.uic_lse_xent_fwd(ptr(logits), ld, ptr(target), ptr(mask), ptr(lse), ptr(nll), R, V, stream()))
    torch.testing.assert_close(lse, torch.logsumexp(logits, 1), rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(nll, -ref.gather(1, target[:, None]).squeeze(1) * mask, rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("R,V,k", [(6, 52, 3), (40, 10000, 3), (9, 10000, 5), (5, 30000, 10), (3, 20, 16)])
def test_row_topk(R, V, k):
    lib = _lib.load()
    logits = torch.randn(R, V, device=DEV) * 2
    logits[:, V - 1] += 20.0                      # UNK would win without the -1000 edit
    prev = torch.randint(0, V - 1, (R,), device=DEV)
    logits[torch.arange(R), prev] += 30.0         # the banned token would win without the constraint
    val, idx = torch.empty(R, k, device=DEV), torch.empty(R, k, device=DEV, dtype=torch.int32)
    for flags in (0, _lib.SAMPLE_DECODING_CONSTRAINT):
        check(lib.uic_row_topk(ptr(logits), V, ptr(prev), ptr(val), ptr(idx), R, V, k, flags, stream()))
        lp = torch.log_softmax(logits, 1)
        if flags:
            lp[torch.arange(R), prev] = float("-inf")
        lp[:, V - 1] -= 1000.0
        ref_val, ref_idx = torch.sort(lp, dim=1, descending=True, stable=True)
        assert torch.equal(idx.long(), ref_idx[:, :k])
        torch.testing.assert_close(val, ref_val[:, :k], rtol=1e-5, atol=3e-5)


def test_greedy_step_sequence_semantics():
    """Three steps on hand-made logits: finished rows emit 0, log-probs are not masked, and nothing is
    written once every row has finished (models/AttModel.py:242-251)."""
    lib = _lib.load()
    R, V, T = 4, 50, 5
    seq, lp = torch.zeros(R, T, dtype=torch.int64, device=DEV), torch.zeros(R, T, device=DEV)
    unf, tok = torch.zeros(R, dtype=torch.uint8, device=DEV), torch.zeros(R, dtype=torch.int64, device=DEV)
    nunf = torch.zeros(T, dtype=torch.int32, device=DEV)
    want = [[3, 0, 7, 9], [5, 4, 0, 0], [0, 0, 0, 0], [8, 8, 8, 8]]   # argmax per step/row; step 3 must be skipped
    refs = []
    for t in range(4):
        logits = torch.randn(R, V, device=DEV)
        logits[torch.arange(R), torch.tensor(want[t], device=DEV)] += 10.0
        refs.append(torch.log_softmax(logits, 1).max(1).values)
        check(lib.uic_greedy_step(ptr(logits), V, ptr(seq), ptr(lp), ptr(unf), ptr(tok), ptr(nunf), t, T, R, V, 0, stream()))
    assert seq[:, :3].t().tolist() == [[3, 0, 7, 9], [5, 0, 0, 0], [0, 0, 0, 0]]
    assert nunf.tolist() == [3, 1, 0, 0, 0]
    for t in range(3):
        torch.testing.assert_close(lp[:, t], refs[t], rtol=1e-5, atol=2e-5)
    assert float(lp[:, 3:].abs().max()) == 0.0 and int(seq[:, 3:].abs().max()) == 0


# ---------------------------------------------------------------------------------------------------
# fused logit statistics (logit GEMM whose epilogue keeps max / sum-exp / best keys instead of logits)
# ---------------------------------------------------------------------------------------------------
def _stats_inputs(R, V, H, seed):
    h = _rand_bf16(R, H, seed=seed)
    w = _rand_bf16(V, H, seed=seed + 1, scale=0.5)
    bias = torch.randn(V, device=DEV) * 0.5
    logits = h.float() @ w.float().t() + bias
    return h, w, bias, logits


@pytest.mark.parametrize("R,V,H,k,kslots", [(6, 52, 64, 3, 3), (300, 10000, 512, 3, 3), (129, 10000, 512, 5, 5), (50, 10000, 512, 2, 3),
                                            (40, 9489, 512, 8, 8), (17, 300, 128, 1, 1), (768, 10000, 512, 3, 3)])
def test_logit_stats_topk_equals_reference_order(R, V, H, k, kslots):
    """uic_logit_stats + uic_beam_topk_merge == log_softmax, UNK - 1000, banned -inf, stable sort
    (models/CaptionModel.py:130-140), without materialising the logits."""
    lib = _lib.load()
    h, w, bias, logits = _stats_inputs(R, V, H, seed=R + V)
    bias[V - 1] += 30.0                                   # UNK would win without the -1000 edit
    prev = torch.randint(0, V - 1, (R,), device=DEV)
    logits = h.float() @ w.float().t() + bias
    parts = lib.uic_logit_stats_parts(R, V)
    stats = torch.empty(R, parts, lib.uic_logit_stats_entry_floats(kslots), device=DEV)
    val, idx = torch.empty(R, k, device=DEV), torch.empty(R, k, device=DEV, dtype=torch.int32)
    for constrained in (False, True):
        if constrained:   # ban each row's current best token: the constraint must change the answer
            prev = logits[:, :V - 1].argmax(1)
        check(lib.uic_logit_stats(ptr(h), H, ptr(w), H, ptr(bias), ptr(prev) if constrained else None, 1, ptr(stats),
                                  R, V, H, kslots, 1, 0.0, None, 0, stream()))
        check(lib.uic_beam_topk_merge(ptr(stats), parts, kslots, ptr(val), ptr(idx), R, k, stream()))
        lp = torch.log_softmax(logits, 1)
        if constrained:
            lp[torch.arange(R), prev] = float("-inf")
        lp[:, V - 1] -= 1000.0
        ref_val, ref_idx = torch.sort(lp, dim=1, descending=True, stable=True)
        torch.testing.assert_close(val, ref_val[:, :k], rtol=1e-4, atol=2e-4)
        gap = (ref_val[:, :k] - ref_val[:, 1:k + 1]).abs()           # an order flip needs a near-tie
        prev_gap = torch.cat([torch.full((R, 1), 1.0, device=DEV), gap[:, :-1]], 1)
        clear = (gap > 1e-3) & (prev_gap > 1e-3)
        assert bool(clear.float().mean() > 0.9)
        assert torch.equal(idx.long()[clear], ref_idx[:, :k][clear])
        assert int((idx == V - 1).sum()) == 0


def test_logit_stats_without_unk_suppression_keeps_unk():
    lib = _lib.load()
    R, V, H = 33, 1000, 128
    h, w, bias, _ = _stats_inputs(R, V, H, seed=5)
    bias[V - 1] += 50.0
    logits = h.float() @ w.float().t() + bias
    parts = lib.uic_logit_stats_parts(R, V)
    stats = torch.empty(R, parts, 4, device=DEV)
    val, idx = torch.empty(R, 1, device=DEV), torch.empty(R, 1, device=DEV, dtype=torch.int32)
    check(lib.uic_logit_stats(ptr(h), H, ptr(w), H, ptr(bias), None, 0, ptr(stats), R, V, H, 1, 0, 0.0, None, 0, stream()))
    check(lib.uic_beam_topk_merge(ptr(stats), parts, 1, ptr(val), ptr(idx), R, 1, stream()))
    assert bool((idx == V - 1).all())
    torch.testing.assert_close(val[:, 0], torch.log_softmax(logits, 1)[:, V - 1], rtol=1e-4, atol=2e-4)


def test_greedy_merge_equals_greedy_step():
    """The fused greedy path writes the same tokens / log-probs / counters as logits + uic_greedy_step."""
    lib = _lib.load()
    R, V, H, T = 70, 10000, 512, 4
    parts = lib.uic_logit_stats_parts(R, V)
    stats = torch.empty(R, parts, 4, device=DEV)

    def state():
        return (torch.zeros(R, T, dtype=torch.int64, device=DEV), torch.zeros(R, T, device=DEV),
                torch.zeros(R, dtype=torch.uint8, device=DEV), torch.zeros(R, dtype=torch.int64, device=DEV),
                torch.zeros(T, dtype=torch.int32, device=DEV))

    a, b = state(), state()
    logits = torch.empty(R, V, device=DEV)
    for t in range(3):
        h, w = _rand_bf16(R, H, seed=10 + t), _rand_bf16(V, H, seed=20 + t, scale=0.05)
        bias = torch.randn(V, device=DEV) * 0.1
        bias[0] += 4.0                                     # some rows finish (token 0) at every step
        _lib.gemm(h, w, bias=bias, out_f32=logits)
        flags = _lib.SAMPLE_DECODING_CONSTRAINT if t > 0 else 0
        check(lib.uic_greedy_step(ptr(logits), V, ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(a[3]), ptr(a[4]), t, T, R, V, flags, stream()))
        banned = b[0][:, t - 1:] if t > 0 else None
        check(lib.uic_logit_stats(ptr(h), H, ptr(w), H, ptr(bias), ptr(banned), T, ptr(stats), R, V, H, 1, 0, 0.0, None, 0, stream()))
        check(lib.uic_greedy_merge(ptr(stats), parts, ptr(b[0]), ptr(b[1]), ptr(b[2]), ptr(b[3]), ptr(b[4]), t, T, R, stream()))
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
    assert 0 < int(a[4][0]) < R
    torch.testing.assert_close(a[1], b[1], rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("compact_first", [False, True])
@pytest.mark.parametrize("B,b,kslots", [(5, 3, 3), (9, 2, 3), (4, 5, 5), (3, 8, 8), (7, 1, 1)])
def test_beam_advance_equals_the_four_step_chain(B, b, kslots, compact_first):
    """uic_beam_advance == uic_beam_topk_merge + uic_beam_step + uic_beam_gather + uic_embed_rows, step by step.
    compact_first: the first step's source buffers hold ONE row per image (src_beams = 1): only beam 0 is read at t = 0."""
    lib = _lib.load()
    V, Hh, E, T = 300, 64, 32, 6
    R = B * b
    parts = lib.uic_logit_stats_parts(R, V)
    es = lib.uic_logit_stats_entry_floats(kslots)
    table = _rand_bf16(V, E, seed=3)
    ld_x = E + 2 * Hh + 8
    ga, na, gb, nb_ = E, Hh, E + Hh, Hh      # two state column ranges after the embedding columns

    def state():
        g = torch.Generator(device="cpu").manual_seed(11)
        return {"beam_seq": torch.zeros(B, b, T, dtype=torch.int32, device=DEV), "beam_lp": torch.zeros(B, b, T, device=DEV),
                "beam_sum": torch.zeros(B, b, device=DEV), "done_seq": torch.zeros(B, b, T, dtype=torch.int32, device=DEV),
                "done_lp": torch.zeros(B, b, T, device=DEV), "done_p": torch.zeros(B, b, dtype=torch.float64, device=DEV),
                "done_unaug": torch.zeros(B, b, device=DEV), "done_cnt": torch.zeros(B, dtype=torch.int32, device=DEV),
                "parent": torch.zeros(R, dtype=torch.int32, device=DEV), "tok": torch.zeros(R, dtype=torch.int64, device=DEV),
                "X": [torch.randn(R, ld_x, generator=g).to(DEV).to(torch.bfloat16), torch.zeros(R, ld_x, device=DEV, dtype=torch.bfloat16)],
                "c": [torch.randn(2, R, Hh, generator=g).to(DEV), torch.zeros(2, R, Hh, device=DEV)]}

    u, f = state(), state()
    tkv, tki = torch.empty(R, b, device=DEV), torch.empty(R, b, dtype=torch.int32, device=DEV)
    for t in range(T):
        h, w, bias, _ = _stats_inputs(R, V, Hh, seed=40 + t)
        bias[0] += 3.0                                   # some beams finish early
        stats = torch.empty(R, parts, es, device=DEV)
        check(lib.uic_logit_stats(ptr(h), Hh, ptr(w), Hh, ptr(bias), None, 0, ptr(stats), R, V, Hh, kslots, 1, 0.0, None, 0, stream()))
        src, dst = t % 2, (t + 1) % 2
        move = int(t + 1 < T)
        # unfused chain
        check(lib.uic_beam_topk_merge(ptr(stats), parts, kslots, ptr(tkv), ptr(tki), R, b, stream()))
        check(lib.uic_beam_step(ptr(tkv), ptr(tki), None, ptr(u["beam_seq"]), ptr(u["beam_lp"]), ptr(u["beam_sum"]), ptr(u["done_seq"]),
                                ptr(u["done_lp"]), ptr(u["done_p"]), ptr(u["done_unaug"]), ptr(u["done_cnt"]), ptr(u["parent"]),
                                ptr(u["tok"]), t, T, B, b, 0, stream()))
        if move:
            check(lib.uic_beam_gather(ptr(u["parent"]), ptr(u["X"][src]), ptr(u["X"][dst]), ld_x, ga, na, gb, nb_, ptr(u["c"][src]),
                                      ptr(u["c"][dst]), 2, R, Hh, stream()))
            check(lib.uic_embed_rows(ptr(table), E, ptr(u["tok"]), ptr(u["X"][dst]), ld_x, R, E, V, stream()))
        # fused
        st_f, x_f, c_f, sb = stats, f["X"][src], f["c"][src], b
        if compact_first and t == 0:
            st_f, x_f, c_f, sb = stats[::b].contiguous(), f["X"][src][::b].contiguous(), f["c"][src][:, ::b].contiguous(), 1
        check(lib.uic_beam_advance(ptr(st_f), parts, kslots, ptr(f["beam_seq"]), ptr(f["beam_lp"]), ptr(f["beam_sum"]),
                                   ptr(f["done_seq"]), ptr(f["done_lp"]), ptr(f["done_p"]), ptr(f["done_unaug"]), ptr(f["done_cnt"]),
                                   ptr(f["parent"]), ptr(f["tok"]), t, T, B, b, 0, move, ptr(x_f), ptr(f["X"][dst]), ld_x,
                                   ga, na, gb, nb_, ptr(c_f), ptr(f["c"][dst]), 2, Hh, ptr(table), E, 0, E, V, sb, stream()))
        for k in ("beam_seq", "beam_lp", "beam_sum", "done_seq", "done_lp", "done_p", "done_unaug", "done_cnt", "parent", "tok"):
            assert torch.equal(u[k], f[k]), (k, t)
        if move:
            assert torch.equal(u["X"][dst][:, :E + 2 * Hh], f["X"][dst][:, :E + 2 * Hh]) and torch.equal(u["c"][dst], f["c"][dst])
    assert int(u["done_cnt"].sum()) > 0


def test_greedy_advance_equals_merge_plus_embed():
    lib = _lib.load()
    R, V, H, T, E = 70, 1000, 128, 4, 48
    parts = lib.uic_logit_stats_parts(R, V)
    stats = torch.empty(R, parts, 4, device=DEV)
    table = _rand_bf16(V, E, seed=9)

    def state():
        return (torch.zeros(R, T, dtype=torch.int64, device=DEV), torch.zeros(R, T, device=DEV),
                torch.zeros(R, dtype=torch.uint8, device=DEV), torch.zeros(R, dtype=torch.int64, device=DEV),
                torch.zeros(T, dtype=torch.int32, device=DEV), torch.zeros(R, E + 16, device=DEV, dtype=torch.bfloat16))

    a, b = state(), state()
    for t in range(3):
        h, w = _rand_bf16(R, H, seed=10 + t), _rand_bf16(V, H, seed=20 + t, scale=0.05)
        bias = torch.randn(V, device=DEV) * 0.1
        bias[0] += 3.0
        check(lib.uic_logit_stats(ptr(h), H, ptr(w), H, ptr(bias), None, 0, ptr(stats), R, V, H, 1, 0, 0.0, None, 0, stream()))
        check(lib.uic_greedy_merge(ptr(stats), parts, ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(a[3]), ptr(a[4]), t, T, R, stream()))
        check(lib.uic_embed_rows(ptr(table), E, ptr(a[3]), ptr(a[5][:, 8:]), E + 16, R, E, V, stream()))
        check(lib.uic_greedy_advance(ptr(stats), parts, ptr(b[0]), ptr(b[1]), ptr(b[2]), ptr(b[3]), ptr(b[4]), t, T, R, ptr(table), E,
                                     ptr(b[5][:, 8:]), E + 16, E, V, 0.0, None, stream()))
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    assert 0 < int(a[4][0]) < R


@pytest.mark.parametrize("B,b,group,V", [(5, 2, 0, 50), (5, 2, 2, 50), (3, 4, 3, 40), (7, 1, 5, 30)])
def test_diverse_select(B, b, group, V):
    """uic_row_topk(k = b (group + 1)) + uic_diverse_select == sort of the fully penalised row (CaptionModel.py:36-45,61)."""
    lib = _lib.load()
    T, lt, lam = 6, 2, 0.75
    g = torch.Generator().manual_seed(B * 100 + b * 10 + group)
    R = B * b
    logits = torch.randn(R, V, generator=g).cuda()
    tables = torch.randint(0, 12, (group + 1, B, b, T), generator=g, dtype=torch.int32).cuda()   # few distinct tokens: repeats
    kp = b * (group + 1)
    cv, ci = torch.empty(R, kp, device=DEV), torch.empty(R, kp, dtype=torch.int32, device=DEV)
    check(lib.uic_row_topk(ptr(logits), V, None, ptr(cv), ptr(ci), R, V, kp, 0, stream()))
    tv, tu, ti = torch.empty(R, b, device=DEV), torch.empty(R, b, device=DEV), torch.empty(R, b, dtype=torch.int32, device=DEV)
    check(lib.uic_diverse_select(ptr(cv), ptr(ci), kp, ptr(tables), group, B, b, T, lt, lam, ptr(tv), ptr(tu), ptr(ti), stream()))
    lp = torch.log_softmax(logits.double(), 1).float().cpu()
    lp[:, V - 1] -= 1000.0
    aug = lp.clone()
    for r in range(R):
        for gg in range(group):
            for j in range(b):
                aug[r, int(tables[gg, r // b, j, lt])] -= lam
    ys, ix = torch.sort(aug, dim=1, descending=True, stable=True)
    assert torch.equal(ti.cpu().long(), ix[:, :b])
    torch.testing.assert_close(tv.cpu(), ys[:, :b], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(tu.cpu(), torch.gather(lp, 1, ix[:, :b]), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,L,D,bf16", [(5, 7, 64, True), (37, 50, 2048, True), (9, 13, 200, False)])
def test_col_moments(B, L, D, bf16):
    """Column sums / sums of squares over the packed valid regions (BatchNorm1d statistics of att_embed, AttModel.py:44-53,80)."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(B + L)
    x = torch.randn(B, L, D, generator=g).clamp_(min=0).cuda()
    x = x.bfloat16() if bf16 else x
    lens = torch.randint(1, L + 1, (B,), generator=g, dtype=torch.int32).cuda()
    mom = torch.zeros(2, D, dtype=torch.float64, device=DEV)
    check(lib.uic_col_moments(ptr(x), int(bf16), D, ptr(lens), B, L, D, ptr(mom[0]), ptr(mom[1]), stream()))
    valid = (torch.arange(L, device=DEV)[None, :] < lens[:, None])
    xv = x[valid].double()
    torch.testing.assert_close(mom[0], xv.sum(0), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(mom[1], (xv * xv).sum(0), rtol=1e-6, atol=1e-6)
    mom.zero_()
    check(lib.uic_col_moments(ptr(x), int(bf16), D, None, B, L, D, ptr(mom[0]), ptr(mom[1]), stream()))
    torch.testing.assert_close(mom[0], x.double().sum((0, 1)), rtol=1e-6, atol=1e-6)
