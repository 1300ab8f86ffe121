"""Teacher-forced training path: forward with saved step operands, fused masked cross-entropy and
hand-written BPTT, exposed as torch.autograd.Functions so that the reference's training call
pattern (trainer.py:164-173) works unchanged:

    out  = model(fc, attri, att, labels, att_masks)          # (B, T, V) log-probs, differentiable
    loss = crit(out, labels[:, 1:], masks[:, 1:]); loss.backward()

and, without ever materialising (B, T, V) (SURVEY.md §8b "additive fast path"):

    loss = model(fc, attri, att, labels, masks, att_masks, mode='forward_loss'); loss.backward()

Everything between the Python loop and the GPU is a C-ABI call into libuic_b200.so; torch is used
for allocation, views and the final slicing of packed weight gradients back into the reference's
parameter layout.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, gemm, ptr, stream
from .engine import BF16, Slots, _nvtx

LOGIT_CHUNK_ROWS = 2048  # rows of the (B*T, V) logit matrix processed per fused fwd+bwd pass (stays L2-resident)


def _round8(n):
    return (n + 7) // 8 * 8


class _Run:
    """Saved state of one teacher-forced pass."""


# ----------------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------------
def teacher_forced_run(model, fc_feats, att_feats, seq, att_masks, save=True, all_steps=False, ss=None, drop=None):
    """ss = (ss_prob, seed tensor) turns on scheduled sampling (models/AttModel.py:130-143): from step 1 on, each row's
    input token is replaced with probability ss_prob by a draw from the model's distribution of the previous step
    (Gumbel-max in the logit GEMM's statistics epilogue; no gradient flows through the draw, as in the reference).
    drop = (p, seed tensor) applies the training-mode nn.Dropout layers of the reference (embed, fc_embed, att_embed,
    core output) with the library's counter-based masks; bptt re-applies the same masks to the gradients."""
    eng = model.engine
    w, lib, st = eng.w, eng.lib, stream()
    kind, H, E, A = eng.kind, w.H, w.E, w.A
    feats = eng.prepare(fc_feats, att_feats, att_masks, keep_inputs=True, drop=drop)
    B, L, dev = feats.B, feats.L, feats.att.device
    T_total = seq.size(1) - 1
    # The reference stops at the first all-zero token column (AttModel.py:148-151).  Those steps carry mask 0,
    # so the fused loss (and its gradients) are identical if they are simply run: no host sync needed there.
    T = T_total if all_steps else model._active_steps(seq)
    sl = Slots(kind, E, H)
    r = _Run()
    r.model, r.feats, r.B, r.L, r.T, r.T_total, r.sl, r.drop = model, feats, B, L, T, T_total, sl, drop
    r.X = torch.zeros(T + 1, B, w.Kx, dtype=BF16, device=dev)
    r.c = torch.zeros(T + 1, sl.n_state, B, H, dtype=torch.float32, device=dev)
    r.alpha = torch.empty(T, B, L, dtype=torch.float32, device=dev)
    r.h_all = torch.zeros(B, T_total, H, dtype=BF16, device=dev)
    # (T, B) step-major like X.  An explicit copy: for B == 1 the transposed view already counts as contiguous, and the
    # scheduled-sampling draws written into r.tokens would land in the caller's label tensor
    r.tokens = seq[:, :T].t().long().clone(memory_format=torch.contiguous_format)
    X2d = r.X.view((T + 1) * B, w.Kx)
    check(lib.uic_embed_rows(ptr(w.emb_relu), E, ptr(r.tokens), ptr(X2d[:, sl.xt[0]:]), w.Kx, T * B, E, w.V, st))
    if drop is not None:   # self.embed's nn.Dropout; rows are t * B + b
        _lib.dropout(X2d[:T * B, sl.xt[0]:sl.xt[0] + E], drop, _lib.DROP_XT)
    if kind == "att2in2":
        r.S = torch.empty(T, B, 5 * H + A, dtype=torch.float32, device=dev)
        r.ctx = torch.empty(T, B, H, dtype=BF16, device=dev)
        r.a2c = None if w.all_gates else torch.empty(T, B, 2 * H, dtype=torch.float32, device=dev)
    else:
        r.G1 = torch.empty(T, B, 4 * H, dtype=torch.float32, device=dev)
        r.G2 = torch.empty(T, B, 4 * H, dtype=torch.float32, device=dev)
        r.ah = torch.empty(T, B, A, dtype=torch.float32, device=dev)
        r.X[:T, :, sl.fc[0]:sl.fc[1]] = feats.fc                       # fc is re-fed at every step (:432)
    if ss is not None:
        ss_prob, ss_seed = float(ss[0]), ss[1]
        parts = int(lib.uic_logit_stats_parts(B, w.V))
        ss_stats = torch.empty(B, parts, 4, dtype=torch.float32, device=dev)
        seq_l = seq.long().contiguous()
    for t in range(T):
        if ss is not None and t >= 1:
            h_prev = r.h_all[:, t - 1]                                   # (B, H) view, row pitch T_total * H
            check(lib.uic_logit_stats(ptr(h_prev), h_prev.stride(0), ptr(w.w_logit), H, ptr(w.b_logit), None, 0, ptr(ss_stats),
                                      B, w.V, H, 1, 0, 1.0, ptr(ss_seed), t, st))
            check(lib.uic_ss_advance(ptr(ss_stats), parts, ptr(seq_l[:, t]), seq_l.stride(0), ss_prob, ptr(ss_seed), t,
                                     ptr(r.tokens[t]), B, ptr(w.emb_relu), E, ptr(r.X[t][:, sl.xt[0]:]), w.Kx, E, w.V, st))
            if drop is not None:
                _lib.dropout(r.X[t][:, sl.xt[0]:sl.xt[0] + E], drop, _lib.DROP_XT, row0=t * B)
        if kind == "att2in2":
            ws = {"S": r.S[t], "ctx": r.ctx[t], "a2c": None if r.a2c is None else r.a2c[t]}
        else:
            ws = {"G": r.G1[t], "G2": r.G2[t], "att_h": r.ah[t]}
        eng.core_step(r.X[t], r.c[t], feats, ws, X_next=r.X[t + 1], c_out=r.c[t + 1], h_all=r.h_all[:, t], alpha=r.alpha[t])
        if drop is not None:   # the core's output dropout feeds the logit layer only (AttModel.py:431,599); rows are b * T + t
            _lib.dropout(r.h_all[:, t], drop, _lib.DROP_OUT, row0=t, row_stride=T_total)
    return r


def _logit_stage(r, target, mask, inv_norm, dlogprobs=None, logprobs_out=None, want_grad=True):
    """Time-batched logit GEMM + log-softmax/XE, and (fused, chunk by chunk while the logits are still
    in L2) the backward of the logit layer: d h, d W_logit, d b_logit.

    Two modes: fused loss (target/mask/inv_norm given) or dense log-probs (logprobs_out given on the
    forward call; dlogprobs given on the backward call)."""
    eng = r.model.engine
    w, lib, st = eng.w, eng.lib, stream()
    B, T_total, H, V = r.B, r.T_total, w.H, w.V
    rows = B * T_total
    dev = r.h_all.device
    h2d = r.h_all.view(rows, H)
    Vp = _round8(V)
    chunk = min(rows, LOGIT_CHUNK_ROWS)
    logits = torch.empty(chunk, V, dtype=torch.float32, device=dev)
    out = {}
    if want_grad:
        dlogits = torch.empty(chunk, Vp, dtype=BF16, device=dev)
        out["dh"] = torch.empty(rows, H, dtype=torch.float32, device=dev)
        out["dW"] = torch.zeros(V, H, dtype=torch.float32, device=dev)
        out["db"] = torch.zeros(V, dtype=torch.float32, device=dev)
    if target is not None:
        out["lse"] = torch.empty(rows, dtype=torch.float32, device=dev)
        out["nll"] = torch.empty(rows, dtype=torch.float32, device=dev)
    for r0 in range(0, rows, chunk):
        n = min(chunk, rows - r0)
        hs = h2d[r0:r0 + n]
        if dlogprobs is None:  # forward (or fused forward+backward)
            gemm(hs, w.w_logit, w.b_logit, out_f32=logits[:n])
        if target is not None:
            check(lib.uic_lse_xent_fwd(ptr(logits), V, ptr(target[r0:]), ptr(mask[r0:]), ptr(out["lse"][r0:]), ptr(out["nll"][r0:]),
                                       n, V, st))
            if want_grad:
                check(lib.uic_lse_xent_bwd(ptr(logits), V, ptr(out["lse"][r0:]), ptr(target[r0:]), ptr(mask[r0:]), ptr(inv_norm), 1.0,
                                           ptr(dlogits), Vp, n, V, st))
        elif dlogprobs is None:
            check(lib.uic_log_softmax_rows(ptr(logits), V, ptr(logprobs_out[r0:]), V, n, V, st))
            continue
        else:
            check(lib.uic_log_softmax_bwd(ptr(dlogprobs[r0:]), V, ptr(logprobs_out[r0:]), V, ptr(dlogits), Vp, n, V, st))
        if want_grad:
            dl = dlogits[:n]
            # K = V: the TMA zero-fills the K tail of both operands, the row pitch Vp keeps rows 16-byte aligned
            gemm(dl[:, :V], w.w_logit, out_f32=out["dh"][r0:r0 + n], b_mn=True)
            gemm(dl[:, :V], hs, out_f32=out["dW"], a_mn=True, b_mn=True, accumulate=True)
            check(lib.uic_col_sum(ptr(dl), 1, Vp, ptr(out["db"]), n, V, st))
    return out


# ----------------------------------------------------------------------------------------------------
# backward through time
# ----------------------------------------------------------------------------------------------------
def bptt(r, dh_all, on_ready=None):
    """dh_all: (B*T_total, H) fp32 gradient w.r.t. the step outputs (from the logit layer).
    Returns {reference parameter name: gradient}.  on_ready(dict): called with each group of gradients as soon as it is
    final (recurrent core + fc_embed after the time-batched wgrads, then the embedding table; the feature-side layers
    come last, in the returned dict) so that a data-parallel caller can start their all-reduce while the rest still runs."""
    eng = r.model.engine
    w, lib, st = eng.w, eng.lib, stream()
    kind, H, E, A, V = eng.kind, w.H, w.E, w.A, w.V
    B, L, T, T_total, sl, feats = r.B, r.L, r.T, r.T_total, r.sl, r.feats
    dev = dh_all.device
    f32 = dict(dtype=torch.float32, device=dev)
    if r.drop is not None:   # output dropout: the same mask on the gradient (rows b * T + t, like the forward)
        _lib.dropout(dh_all, r.drop, _lib.DROP_OUT)
    dh3 = dh_all.view(B, T_total, H)
    ld_dh = T_total * H
    de = torch.empty(T, B, L, **f32)
    dc = [torch.empty(sl.n_state, B, H, **f32) for _ in range(2)]
    X2d = r.X.view((T + 1) * B, w.Kx)[:T * B]
    g = {}

    def emit(names):
        if on_ready is not None:
            on_ready({n: g[n] for n in names})

    if kind == "att2in2":
        NS = 5 * H + A
        dS = torch.empty(T, B, NS, dtype=BF16, device=dev)
        da2c = None if w.all_gates else torch.empty(T, B, 2 * H, dtype=BF16, device=dev)
        dX = torch.empty(T, B, w.Kx, **f32)
        dctx = torch.empty(T, B, H, **f32)
        for t in reversed(range(T)):
            last = t + 1 == T
            check(lib.uic_lstm_maxout_bwd(ptr(r.S[t]), NS, None if da2c is None else ptr(r.a2c[t]), 2 * H, ptr(r.c[t][0]), ptr(r.c[t + 1][0]),
                                          ptr(dh3[:, t]), ld_dh, None if last else ptr(dX[t + 1][:, E:]), w.Kx,
                                          None if last else ptr(dc[(t + 1) % 2][0]), ptr(dS[t]), NS, None if da2c is None else ptr(da2c[t]), 2 * H,
                                          ptr(dc[t % 2][0]), B, H, st))
            # d ctx through a2c (2H maxout inputs) or a2h (att2all2: all five gate sums)
            gemm(dS[t][:, :5 * H] if da2c is None else da2c[t], w.w_a2c, out_f32=dctx[t], b_mn=True)
            check(lib.uic_att_step_bwd(ptr(dctx[t]), H, ptr(r.alpha[t]), ptr(feats.p_att), ptr(feats.att), ptr(r.S[t][:, 5 * H:]), NS,
                                       ptr(w.w_alpha), ptr(de[t]), ptr(dS[t][:, 5 * H:]), NS, B, L, A, H, st))
            gemm(dS[t], w.w1, out_f32=dX[t], b_mn=True)
        dS2 = dS.view(T * B, NS)
        da2 = dS2[:, :5 * H] if da2c is None else da2c.view(T * B, 2 * H)
        Na = da2.shape[1]
        dW1 = torch.empty(NS, w.Kx, **f32)
        gemm(dS2, X2d, out_f32=dW1, a_mn=True, b_mn=True)
        db1 = torch.zeros(NS, **f32)
        check(lib.uic_col_sum(ptr(dS2), 1, NS, ptr(db1), T * B, NS, st))
        dWa = torch.empty(Na, H, **f32)
        gemm(da2, r.ctx.view(T * B, H), out_f32=dWa, a_mn=True, b_mn=True)
        dba = torch.zeros(Na, **f32)
        check(lib.uic_col_sum(ptr(da2), 1, da2.stride(0), ptr(dba), T * B, Na, st))
        g["core.i2h.weight"], g["core.h2h.weight"] = dW1[:5 * H, :E], dW1[:5 * H, E:]
        g["core.attention.h2att.weight"] = dW1[5 * H:, E:]
        g["core.i2h.bias"] = g["core.h2h.bias"] = db1[:5 * H]
        g["core.attention.h2att.bias"] = db1[5 * H:]
        g["core.a2h.weight" if da2c is None else "core.a2c.weight"], g["core.a2h.bias" if da2c is None else "core.a2c.bias"] = dWa, dba
        emit([n for n in g if n.startswith("core.")])
        dxt2d, ld_dxt = dX.view(T * B, w.Kx), w.Kx
        dctx_ptr, dctx_stride, dctx_ld = dctx, B * H, H
        ah_ptr, ah_stride, ah_ld = r.S[0][:, 5 * H:], B * NS, NS
    else:
        K1 = E + 3 * H
        dG1 = torch.empty(T, B, 4 * H, dtype=BF16, device=dev)
        dG2 = torch.empty(T, B, 4 * H, dtype=BF16, device=dev)
        dXa = torch.empty(T, B, K1, **f32)          # grads of [h_att_prev | xt | fc | h_lang_prev]
        dXb = torch.empty(T, B, 3 * H, **f32)       # grads of [h_lang_prev | h_att | ctx]
        dah = torch.empty(T, B, A, dtype=BF16, device=dev)
        for t in reversed(range(T)):
            last = t + 1 == T
            check(lib.uic_lstm_cell_bwd(ptr(r.G2[t]), 4 * H, ptr(r.c[t][1]), ptr(r.c[t + 1][1]), ptr(dh3[:, t]), ld_dh,
                                        None if last else ptr(dXa[t + 1][:, E + 2 * H:]), K1,
                                        None if last else ptr(dXb[t + 1]), 3 * H, None if last else ptr(dc[(t + 1) % 2][1]),
                                        ptr(dG2[t]), 4 * H, ptr(dc[t % 2][1]), B, H, st))
            gemm(dG2[t], w.w2, out_f32=dXb[t], b_mn=True)
            check(lib.uic_att_step_bwd(ptr(dXb[t][:, 2 * H:]), 3 * H, ptr(r.alpha[t]), ptr(feats.p_att), ptr(feats.att), ptr(r.ah[t]), A,
                                       ptr(w.w_alpha), ptr(de[t]), ptr(dah[t]), A, B, L, A, H, st))
            gemm(dah[t], w.w_h2att, out_f32=dXb[t][:, H:2 * H], b_mn=True, accumulate=True)
            check(lib.uic_lstm_cell_bwd(ptr(r.G1[t]), 4 * H, ptr(r.c[t][0]), ptr(r.c[t + 1][0]), ptr(dXb[t][:, H:]), 3 * H,
                                        None if last else ptr(dXa[t + 1]), K1, None, 0, None if last else ptr(dc[(t + 1) % 2][0]),
                                        ptr(dG1[t]), 4 * H, ptr(dc[t % 2][0]), B, H, st))
            gemm(dG1[t], w.w1, out_f32=dXa[t], b_mn=True)
        d1, d2, da = dG1.view(T * B, 4 * H), dG2.view(T * B, 4 * H), dah.view(T * B, A)
        dW1 = torch.empty(4 * H, K1, **f32)
        gemm(d1, X2d[:, :K1], out_f32=dW1, a_mn=True, b_mn=True)
        dW2 = torch.empty(4 * H, 3 * H, **f32)
        gemm(d2, X2d[:, E + 2 * H:], out_f32=dW2, a_mn=True, b_mn=True)
        dWh = torch.empty(A, H, **f32)
        gemm(da, X2d[:, sl.h_att[0]:sl.h_att[1]], out_f32=dWh, a_mn=True, b_mn=True)
        db1, db2, dbh = torch.zeros(4 * H, **f32), torch.zeros(4 * H, **f32), torch.zeros(A, **f32)
        check(lib.uic_col_sum(ptr(d1), 1, 4 * H, ptr(db1), T * B, 4 * H, st))
        check(lib.uic_col_sum(ptr(d2), 1, 4 * H, ptr(db2), T * B, 4 * H, st))
        check(lib.uic_col_sum(ptr(da), 1, A, ptr(dbh), T * B, A, st))
        g["core.att_lstm.weight_hh"] = dW1[:, :H]
        g["core.att_lstm.weight_ih"] = torch.cat([dW1[:, E + 2 * H:], dW1[:, H + E:2 * H + E], dW1[:, H:H + E]], 1)
        g["core.att_lstm.bias_ih"] = g["core.att_lstm.bias_hh"] = db1
        g["core.lang_lstm.weight_hh"] = dW2[:, :H]
        g["core.lang_lstm.weight_ih"] = torch.cat([dW2[:, 2 * H:], dW2[:, H:2 * H]], 1)
        g["core.lang_lstm.bias_ih"] = g["core.lang_lstm.bias_hh"] = db2
        g["core.attention.h2att.weight"], g["core.attention.h2att.bias"] = dWh, dbh
        # fc_embed: the embedded fc vector is an input of every step
        dfc = torch.empty(B, H, **f32)
        check(lib.uic_reduce_time(ptr(dXa), B * K1, K1, H + E, ptr(dfc), T, B, H, st))
        if r.drop is not None:
            _lib.dropout(dfc, r.drop, _lib.DROP_FC)
        dfc_pre = torch.empty(B, H, dtype=BF16, device=dev)
        check(lib.uic_relu_bwd_cast(ptr(dfc), ptr(feats.fc), ptr(dfc_pre), B * H, st))
        dWfc = torch.empty(H, feats.fc_in.size(1), **f32)
        gemm(dfc_pre, feats.fc_in, out_f32=dWfc, a_mn=True, b_mn=True)
        dbfc = torch.zeros(H, **f32)
        check(lib.uic_col_sum(ptr(dfc_pre), 1, H, ptr(dbfc), B, H, st))
        g["fc_embed.0.weight"], g["fc_embed.0.bias"] = dWfc, dbfc
        emit([n for n in g if n.startswith("core.") or n.startswith("fc_embed.")])
        dxt2d, ld_dxt = dXa.view(T * B, K1)[:, H:], K1
        dctx_ptr, dctx_stride, dctx_ld = dXb[0][:, 2 * H:], B * 3 * H, 3 * H
        ah_ptr, ah_stride, ah_ld = r.ah, B * A, A

    # ---- embedding ------------------------------------------------------------------------------------
    if r.drop is not None:
        _lib.dropout(dxt2d[:, :E], r.drop, _lib.DROP_XT)
    demb = torch.zeros(V, E, **f32)
    check(lib.uic_embed_bwd(ptr(dxt2d), ld_dxt, ptr(r.tokens), ptr(w.emb_relu), ptr(demb), T * B, E, V, st))
    g["embed.0.weight"] = demb
    emit(["embed.0.weight"])

    # ---- feature tiles and the prologue layers -----------------------------------------------------------
    datt = torch.empty(B * L, H, **f32)
    dp_att = torch.empty(B * L, A, dtype=BF16, device=dev)
    dw_alpha = torch.zeros(2 * A, **f32)                                  # [d w_alpha | d bias of ctx2att]
    check(lib.uic_att_tiles_bwd(ptr(de), ptr(r.alpha), ptr(dctx_ptr), dctx_stride, dctx_ld, ptr(ah_ptr), ah_stride, ah_ld,
                                ptr(feats.p_att), ptr(w.w_alpha), ptr(datt), ptr(dp_att), ptr(dw_alpha), T, B, L, A, H, st))
    g["core.attention.alpha_net.weight"] = dw_alpha[:A].view(1, A)
    g["core.attention.alpha_net.bias"] = torch.zeros(1, **f32)            # the bias cancels inside the softmax
    att2d = feats.att.view(B * L, H)
    dWc = torch.empty(A, H, **f32)
    gemm(dp_att, att2d, out_f32=dWc, a_mn=True, b_mn=True)
    g["ctx2att.weight"], g["ctx2att.bias"] = dWc, dw_alpha[A:]
    gemm(dp_att, w.w_ctx2att, out_f32=datt, b_mn=True, accumulate=True)   # d att += d p_att @ W_ctx2att
    if r.drop is not None:
        _lib.dropout(datt, r.drop, _lib.DROP_ATT)
    d_pre = torch.empty(B * L, H, dtype=BF16, device=dev)
    check(lib.uic_relu_bwd_cast(ptr(datt), ptr(att2d), ptr(d_pre), B * L * H, st))
    dWe = torch.empty(H, feats.x_in.size(1), **f32)
    gemm(d_pre, feats.x_in, out_f32=dWe, a_mn=True, b_mn=True)
    dbe = torch.zeros(H, **f32)
    check(lib.uic_col_sum(ptr(datt), 0, H, ptr(dbe), B * L, H, st))       # fp32 masked copy left by relu_bwd_cast
    if w.use_bn:
        # the Linear ran on the folded operand W' = W diag(s), b' = b + W t (engine._fold_bn); the statistics depend on the
        # (constant) features only, so the chain rule back to W, b, gamma, beta is column-wise arithmetic on (H, D) matrices
        bn, W = feats.bn, w.w_att_f32
        # z = W' (x - mean) + b + W beta: the (x - mean) terms take their column sums from the same bf16 d_pre that the wgrad
        # GEMM read, so that d_pre's rounding cancels between d W' and mean * colsum even for low-variance columns (large s)
        dbe_r = torch.zeros(H, **f32)
        check(lib.uic_col_sum(ptr(d_pre), 1, H, ptr(dbe_r), B * L, H, st))
        dWc = dWe - dbe_r[:, None] * bn["mean"][None, :]                  # sum_rows d_pre (x - mean)^T
        g["att_embed.0.bias"] = torch.mv(W.t(), dbe)                      # d beta = W^T d b
        g["att_embed.0.weight"] = (dWc * W).sum(0) * bn["inv"]
        g["att_embed.1.weight"] = dWc * bn["s"][None, :] + dbe[:, None] * bn["beta"][None, :]
        g["att_embed.1.bias"] = dbe
    else:
        g["att_embed.0.weight"], g["att_embed.0.bias"] = dWe, dbe
    return g


# ----------------------------------------------------------------------------------------------------
# autograd.Functions
# ----------------------------------------------------------------------------------------------------
def _param_list(model):
    names, params = zip(*model.named_parameters())
    return list(names), list(params)


def _finish(r, grads, grad_scale, names, params):
    out = []
    for n, p in zip(names, params):
        gr = grads.get(n)
        if gr is None or not p.requires_grad:
            out.append(None)
            continue
        gr = gr.reshape(p.shape).to(p.dtype)
        out.append(gr * grad_scale if grad_scale is not None else gr.contiguous())
    return out


class _DecoderLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, fc_feats, att_feats, labels, masks, att_masks, global_mask_sum, ss, drop, *params):
        r = teacher_forced_run(model, fc_feats, att_feats, labels, att_masks, all_steps=True, ss=ss, drop=drop)
        T_total = r.T_total
        target = labels[:, 1:T_total + 1].contiguous().view(-1).long()
        mask = masks[:, 1:T_total + 1].contiguous().view(-1).float()
        norm = mask.sum() if global_mask_sum is None else global_mask_sum.to(mask.device).float()
        inv_norm = (1.0 / norm).reshape(1).contiguous()
        o = _logit_stage(r, target, mask, inv_norm, want_grad=True)
        ctx.run, ctx.logit = r, o
        return o["nll"].sum() * inv_norm[0]

    @staticmethod
    def backward(ctx, grad_out):
        r, o = ctx.run, ctx.logit
        ctx.run = ctx.logit = None   # the saved step operands die with this backward
        names, params = _param_list(r.model)
        # every gradient is linear in (d h, d W_logit, d b_logit): scaling these three by the incoming gradient
        # replaces one strided multiply per parameter (24 launches) by three
        g = bptt(r, o["dh"] * grad_out)
        g["logit.weight"], g["logit.bias"] = o["dW"] * grad_out, o["db"] * grad_out
        return (None,) * 9 + tuple(_finish(r, g, None, names, params))


class _DecoderLogprobsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, fc_feats, att_feats, seq, att_masks, ss, drop, *params):
        r = teacher_forced_run(model, fc_feats, att_feats, seq, att_masks, ss=ss, drop=drop)
        B, T_total, V = r.B, r.T_total, r.model.engine.w.V
        out = torch.zeros(B, T_total, V, dtype=torch.float32, device=r.h_all.device)
        _logit_stage(r, None, None, None, logprobs_out=out.view(B * T_total, V), want_grad=False)
        if r.T < T_total:
            out[:, r.T:] = 0.0   # steps after the all-zero-column break stay zero (AttModel.py:123,148-151)
        ctx.run = r
        ctx.save_for_backward(out)   # (an output kept as a plain ctx attribute would form the cycle out -> grad_fn -> ctx -> out:
        return out                   #  330 MB at 512 x 17 x 10k freed only when the cyclic GC happens to run)

    @staticmethod
    def backward(ctx, dlp):
        r, (out,) = ctx.run, ctx.saved_tensors
        ctx.run = None               # the saved step operands die with this backward
        B, T_total, V = out.shape
        dlp = dlp.contiguous().float()
        if r.T < T_total:
            dlp = dlp.clone()
            dlp[:, r.T:] = 0.0
        o = _logit_stage(r, None, None, None, dlogprobs=dlp.view(B * T_total, V), logprobs_out=out.view(B * T_total, V))
        names, params = _param_list(r.model)
        g = bptt(r, o["dh"])
        g["logit.weight"], g["logit.bias"] = o["dW"], o["db"]
        return (None,) * 7 + tuple(_finish(r, g, None, names, params))


class _DecoderTokenLogprobsFn(torch.autograd.Function):
    """Log-probs of given tokens under teacher forcing, (B, T_total), differentiable: the policy-gradient path of
    self-critical training (trainer.py:166-173 back-propagates through the sampled roll-out, which is the same
    computation as teacher forcing on the sampled tokens).  The (B, T, V) log-probs are never materialised."""

    @staticmethod
    def forward(ctx, model, fc_feats, att_feats, labels, att_masks, drop, *params):
        r = teacher_forced_run(model, fc_feats, att_feats, labels, att_masks, all_steps=True, drop=drop)
        T_total = r.T_total
        target = labels[:, 1:T_total + 1].contiguous().view(-1).long()
        ones = torch.ones(target.numel(), dtype=torch.float32, device=target.device)
        one = torch.ones(1, dtype=torch.float32, device=target.device)
        o = _logit_stage(r, target, ones, one, want_grad=False)
        ctx.run, ctx.target, ctx.one = r, target, one
        return -o["nll"].view(r.B, T_total)

    @staticmethod
    def backward(ctx, dlp):
        r = ctx.run
        ctx.run = None
        # d loss / d logits = (softmax - onehot) * (-d loss / d logprob): the fused XE backward with per-token weights
        weights = (-dlp).contiguous().float().view(-1)
        o = _logit_stage(r, ctx.target, weights, ctx.one, want_grad=True)
        names, params = _param_list(r.model)
        g = bptt(r, o["dh"])
        g["logit.weight"], g["logit.bias"] = o["dW"], o["db"]
        return (None,) * 6 + tuple(_finish(r, g, None, names, params))


def xe_sum_and_grads(model, fc_feats, att_feats, labels, masks, att_masks=None, ss=None, drop=None, on_ready=None):
    """The training computation without the autograd plumbing, for data-parallel trainers (dp.DataParallelStep):
    returns (sum over rows and steps of the masked NLL, sum(mask), {parameter name: gradient of that SUM}).
    Dividing both by the mask sum of the WHOLE batch -- summed over ranks -- gives the reference's loss and gradients
    (misc/criterion.py:149), so no collective has to run before the forward pass.  on_ready(dict) receives groups of
    gradients the moment they are final: logit.* right after the fused logit stage (before BPTT starts), then the
    groups bptt() announces; whatever is left comes with the returned dict."""
    with torch.no_grad():
        with _nvtx("uic.train.forward"):
            r = teacher_forced_run(model, fc_feats, att_feats, labels, att_masks, all_steps=True, ss=ss, drop=drop)
        T_total = r.T_total
        target = labels[:, 1:T_total + 1].contiguous().view(-1).long()
        mask = masks[:, 1:T_total + 1].contiguous().view(-1).float()
        one = torch.ones(1, dtype=torch.float32, device=mask.device)
        with _nvtx("uic.train.logit_stage"):
            o = _logit_stage(r, target, mask, one, want_grad=True)
        head = {"logit.weight": o["dW"], "logit.bias": o["db"]}
        if on_ready is not None:
            on_ready(head)
        with _nvtx("uic.train.bptt"):
            g = bptt(r, o["dh"], on_ready=on_ready)
        g.update(head)
        return o["nll"].sum(), mask.sum(), g


def decoder_token_logprobs(model, fc_feats, att_feats, labels, att_masks=None, drop=None):
    """labels: (B, T_total + 1) int64 with the BOS column in front (token t is the target of step t - 1)."""
    _, params = _param_list(model)
    return _DecoderTokenLogprobsFn.apply(model, fc_feats, att_feats, labels, att_masks, drop, *params)


def decoder_loss(model, fc_feats, att_feats, labels, masks, att_masks=None, global_mask_sum=None, ss=None, drop=None):
    _, params = _param_list(model)
    return _DecoderLossFn.apply(model, fc_feats, att_feats, labels, masks, att_masks, global_mask_sum, ss, drop, *params)


def decoder_logprobs(model, fc_feats, att_feats, seq, att_masks=None, ss=None, drop=None):
    _, params = _param_list(model)
    return _DecoderLogprobsFn.apply(model, fc_feats, att_feats, seq, att_masks, ss, drop, *params)
