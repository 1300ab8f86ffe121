"""ctypes binding of libuic_b200.so (the C ABI declared in include/uic_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libuic_b200.so")

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol of include/uic_b200.h (tests check this)
SIGNATURES = {
    "uic_last_error": (C.c_char_p, []),
    "uic_version": (_i, []),
    "uic_launch_count": (_i64, []),
    "uic_set_gemm_impl": (_i, [_i]),
    "uic_check_device": (_i, []),
    "uic_gemm_set_trace": (_i, [_p]),
    "uic_profile_enable": (_i, [_i]),
    "uic_profile_dump": (_i64, [C.c_char_p, _i64]),
    "uic_gemm_bf16": (_i, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i, _i, _i, _i, _p]),
    "uic_gemm_bf16_ex": (_i, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i, _i, _i, _i, _i, _f, _p]),
    "uic_gemm_bf16_affine": (_i, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _p, _p, _i, _i, _i, _i, _p]),
    "uic_cast_f32_bf16": (_i, [_p, _i64, _p, _i64, _i64, _i64, _i, _p]),
    "uic_embed_rows": (_i, [_p, _i64, _p, _p, _i64, _i, _i, _i, _p]),
    "uic_zero_padded_rows": (_i, [_p, _p, _i, _i, _i, _p]),
    "uic_att_step_fwd": (_i, [_p, _i64, _p, _p, _p, _p, _p, _i64, _p, _i64, _p, _p, _i64, _i, _i, _i, _i, _i, _p]),
    "uic_att_step_workspace_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "uic_lstm_maxout_fwd": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _p, _i64, _p, _i64, _i, _i, _p]),
    "uic_lstm_cell_fwd": (_i, [_p, _i64, _p, _p, _p, _p, _i64, _p, _i64, _i, _i, _p]),
    "uic_lstm_maxout_fwd_add": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _p, _i64, _p, _i64, _i, _i, _p, _i64, _p, _i, _p, _i64, _i, _p]),
    "uic_lstm_cell_fwd_add": (_i, [_p, _i64, _p, _p, _p, _p, _i64, _p, _i64, _i, _i, _p, _i64, _p, _i, _p, _i64, _i, _p]),
    "uic_log_softmax_rows": (_i, [_p, _i64, _p, _i64, _i, _i, _p]),
    "uic_lse_xent_fwd": (_i, [_p, _i64, _p, _p, _p, _p, _i, _i, _p]),
    "uic_greedy_step": (_i, [_p, _i64, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "uic_row_topk": (_i, [_p, _i64, _p, _p, _p, _i, _i, _i, _i, _p]),
    "uic_logit_stats_parts": (_i, [_i, _i]),
    "uic_logit_stats_entry_floats": (_i, [_i]),
    "uic_logit_stats": (_i, [_p, _i64, _p, _i64, _p, _p, _i64, _p, _i, _i, _i, _i, _i, _f, _p, _i, _p]),
    "uic_beam_topk_merge": (_i, [_p, _i, _i, _p, _p, _i, _i, _p]),
    "uic_greedy_merge": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "uic_beam_advance": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _i64, _i, _i, _i, _i, _p, _p,
                              _i, _i, _p, _i64, _i, _i, _i, _i, _p]),
    "uic_greedy_advance": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p, _i64, _p, _i64, _i, _i, _f, _p, _p]),
    "uic_dropout": (_i, [_p, _i, _i64, _i64, _i, _f, _p, _i, _i64, _i64, _p]),
    "uic_ss_advance": (_i, [_p, _i, _p, _i64, _f, _p, _i, _p, _i, _p, _i64, _p, _i64, _i, _i, _p]),
    "uic_beam_step": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "uic_col_moments": (_i, [_p, _i, _i64, _p, _i, _i, _i, _p, _p, _p]),
    "uic_diverse_select": (_i, [_p, _p, _i, _p, _i, _i, _i, _i, _i, _f, _p, _p, _p, _p]),
    "uic_beam_gather": (_i, [_p, _p, _p, _i64, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p]),
    "uic_lstm_cell_bwd": (_i, [_p, _i64, _p, _p, _p, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _p, _i, _i, _p]),
    "uic_lstm_maxout_bwd": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _i64, _p, _i64, _p, _p, _i64, _p, _i64, _p, _i, _i, _p]),
    "uic_att_step_bwd": (_i, [_p, _i64, _p, _p, _p, _p, _i64, _p, _p, _p, _i64, _i, _i, _i, _i, _p]),
    "uic_att_tiles_bwd": (_i, [_p, _p, _p, _i64, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "uic_lse_xent_bwd": (_i, [_p, _i64, _p, _p, _p, _p, _f, _p, _i64, _i, _i, _p]),
    "uic_log_softmax_bwd": (_i, [_p, _i64, _p, _i64, _p, _i64, _i, _i, _p]),
    "uic_col_sum": (_i, [_p, _i, _i64, _p, _i, _i, _p]),
    "uic_embed_bwd": (_i, [_p, _i64, _p, _p, _p, _i64, _i, _i, _p]),
    "uic_relu_bwd_cast": (_i, [_p, _p, _p, _i64, _p]),
    "uic_reduce_time": (_i, [_p, _i64, _i64, _i, _p, _i, _i, _i, _p]),
}

GEMM_RELU, GEMM_ACCUMULATE, GEMM_A_MN, GEMM_B_MN, GEMM_OUT_F16, GEMM_A_STREAM, GEMM_B_STREAM = 1, 2, 4, 8, 16, 32, 64
SAMPLE_DECODING_CONSTRAINT, BEAM_MAX_PPL = 1, 2

_lib = None


class UicError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UicError(f"{LIB_PATH} not found: build it with `python -m unpaired_image_captioning_b200.build` "
                       "(there is no CPU fallback for the decoder kernels)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise UicError(f"uic error {rc}: {load().uic_last_error().decode(errors='replace')}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_device():
    if not torch.cuda.is_available():
        raise UicError("no CUDA device: the decoder kernels are sm_100a-only and have no CPU fallback")
    check(load().uic_check_device())


def launch_count():
    return int(load().uic_launch_count())


def profile(on):
    """Enable/disable (and clear) the library's live per-kernel event timing."""
    check(load().uic_profile_enable(int(on)))


def profile_dump():
    """{label: (launches, total_ms)} since profile(True); synchronises the recorded events."""
    buf = C.create_string_buffer(1 << 16)
    n = load().uic_profile_dump(buf, len(buf))
    if n < 0:
        check(int(n))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(" ", 2)
        out[name] = (int(cnt), float(ms))
    return out


# ---- typed convenience wrappers (validate dtype/contiguity like torch would) -----------------------
def _is_bf16(t):
    return t.dtype == torch.bfloat16 and t.is_cuda and t.stride(-1) == 1


# Exponential operand form of the additive attention: E = exp(2 p_att) is stored once per image as a bf16 tile and
# F = exp(2 att_h) is produced per step in fp32; tanh(p + a) = 1 - 2 / (E F + 1).  bf16 keeps fp32's exponent range, so any
# |p_att| < 20.8 (the 2^60 cap of the GEMM epilogues) is represented with the same relative precision (2^-9 on E, i.e.
# 2^-10 absolute on p_att) -- a first version used an fp16 tile, which saturated for p_att > 6.9 and went subnormal below -3.5.
ATT_E_SCALE, ATT_F_SCALE = 1.0, 1.0
ATT_EXP_CAP = float(2 ** 60)
TILE_DTYPE = torch.bfloat16


def exp_tile(p_att):
    """Raw ctx2att output -> the bf16 operand tile E = exp(2 p) (API-compat path only; the engine gets it straight from
    the ctx2att GEMM epilogue)."""
    return (torch.exp(2.0 * p_att.float()) * ATT_E_SCALE).clamp_(max=ATT_EXP_CAP).to(TILE_DTYPE)


def tile_value(e_tile):
    """The p_att values an operand tile encodes."""
    return 0.5 * torch.log(e_tile.float() / ATT_E_SCALE)


def gemm(a, b, bias=None, out_f32=None, out_bf16=None, relu=False, accumulate=False, a_mn=False, b_mn=False,
         exp_col0=0, exp_scale=0.0, a_stream=False, b_stream=False, post_scale=None, post_shift=None):
    """D[M,N] = act(A @ B^T + bias).  `a` is (M,K) [or (K,M) if a_mn], `b` is (N,K) [or (K,N) if b_mn];
    both 2-D bf16 views whose last stride is 1 (row pitch arbitrary).  a_stream / b_stream: the operand is read once
    (L2 evict-first hint), e.g. the raw feature matrix of the prologue."""
    if not (_is_bf16(a) and _is_bf16(b) and a.dim() == 2 and b.dim() == 2):
        raise ValueError("gemm: operands must be 2-D CUDA bfloat16 tensors with unit inner stride")
    M, K = (a.shape[1], a.shape[0]) if a_mn else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else b.shape
    if K != Kb:
        raise ValueError(f"gemm: K mismatch {K} vs {Kb}")
    out_f16 = out_bf16 is not None and out_bf16.dtype == torch.float16   # 16-bit output buffer may be fp16 (p_att tiles)
    for o, dts in ((out_f32, (torch.float32,)), (out_bf16, (torch.bfloat16, torch.float16))):
        if o is not None and (o.dtype not in dts or o.shape != (M, N) or o.stride(1) != 1):
            raise ValueError("gemm: bad output tensor")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("gemm: bias must be contiguous fp32 of length N")
    flags = ((GEMM_RELU if relu else 0) | (GEMM_ACCUMULATE if accumulate else 0) | (GEMM_A_MN if a_mn else 0) |
             (GEMM_B_MN if b_mn else 0) | (GEMM_OUT_F16 if out_f16 else 0) | (GEMM_A_STREAM if a_stream else 0) |
             (GEMM_B_STREAM if b_stream else 0))
    if post_scale is not None:   # act(.) * post_scale + post_shift per column (eval-mode BatchNorm behind the layer)
        if exp_scale or post_shift is None or post_scale.numel() != N or post_shift.numel() != N:
            raise ValueError("gemm: post_scale / post_shift must be fp32 vectors of length N (not combinable with the exp epilogue)")
        check(load().uic_gemm_bf16_affine(ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
                                          ptr(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0, ptr(bias),
                                          ptr(post_scale.float().contiguous()), ptr(post_shift.float().contiguous()), M, N, K, flags, stream()))
        return
    check(load().uic_gemm_bf16_ex(ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
                                  ptr(out_bf16), out_bf16.stride(0) if out_bf16 is not None else 0, ptr(bias), M, N, K, flags,
                                  int(exp_col0), float(exp_scale), stream()))


_att_ws = {}


def att_workspace(n_img, beams, L, A, H, device):
    """Cached, zero-initialised workspace of uic_att_step_fwd for this shape (stream-ordered reuse)."""
    key = (n_img, beams, L, A, H, str(device))
    ws = _att_ws.get(key)
    if ws is None:
        nbytes = int(load().uic_att_step_workspace_bytes(n_img, beams, L, A, H))
        ws = torch.zeros(max(nbytes, 16), dtype=torch.uint8, device=device)
        # (never evicted: captured CUDA graphs hold raw pointers into these tensors; they are a few KB for large batches and
        # about a megabyte when small batches are cut into segments)
        _att_ws[key] = ws
    return ws


def att_step(att_h, ld_att_h, p_att, att, w_alpha, masks, ctx_bf16, ld_ctx_bf16, ctx_f32, ld_ctx_f32, alpha, n_img, beams, L, A, H):
    """uic_att_step_fwd with dtype checks and the cached workspace.  att_h / ctx_* may be views
    (pass their row pitch); p_att is the bf16 operand tile E (n_img, L, A), att is bf16 (n_img, L, H), both contiguous."""
    if p_att.dtype != TILE_DTYPE or att.dtype != torch.bfloat16 or not p_att.is_contiguous() or not att.is_contiguous():
        raise ValueError("att_step: p_att (operand tile) and att must be contiguous bf16")
    ws = att_workspace(n_img, beams, L, A, H, att.device)
    check(load().uic_att_step_fwd(ptr(att_h), ld_att_h, ptr(p_att), ptr(att), ptr(w_alpha), ptr(masks), ptr(ctx_bf16), ld_ctx_bf16,
                                  ptr(ctx_f32), ld_ctx_f32, ptr(alpha), ptr(ws), ws.numel(), n_img, beams, L, A, H, stream()))


def cast_bf16(src, dst=None, relu=False):
    """fp32 (rows, cols) -> bf16, optional ReLU."""
    if src.dtype != torch.float32 or src.dim() != 2 or src.stride(1) != 1:
        raise ValueError("cast_bf16: src must be a 2-D fp32 tensor with unit inner stride")
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    check(load().uic_cast_f32_bf16(ptr(src), src.stride(0), ptr(dst), dst.stride(0), src.shape[0], src.shape[1], int(relu), stream()))
    return dst


DROP_XT, DROP_ATT, DROP_FC, DROP_OUT = 0, 1, 2, 3   # dropout sites (row ids: t*B+b | b*L+l | b | b*T+t)


def dropout(x, drop, site, row0=0, row_stride=1):
    """In-place training-mode dropout on a 2-D view (fp32 or bf16, unit inner stride); drop = (p, device seed tensor).
    The same call on the gradient of that tensor is its backward."""
    if x.dim() != 2 or x.stride(1) != 1 or x.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("dropout: 2-D fp32/bf16 view with unit inner stride expected")
    check(load().uic_dropout(ptr(x), int(x.dtype == torch.bfloat16), x.stride(0), x.shape[0], x.shape[1], float(drop[0]), ptr(drop[1]),
                             int(site), int(row0), int(row_stride), stream()))
