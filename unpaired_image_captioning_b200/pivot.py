"""Pivot translator of BASELINE.json configs[3] (the "unpaired" joint step: caption decoder + zh->en NMT in one iteration).

north_star keeps the onmt translator in PyTorch "unless profiling shows it matters"; round 1 measured that it does -- run
eagerly it was 32.7 ms of a 35.5 ms joint step, almost all of it launch latency of its per-token python loop (~4 000 small
launches per step).  This module removes that cost WITHOUT new arithmetic kernels: the translator's training step is written
shape-statically (no packed sequences, no host-side lengths) so that forward + NLL + backward + all-reduce + clip + Adam are
captured into ONE CUDA graph and replayed; the arithmetic stays cuBLAS / ATen.  It is the "first move" of SURVEY.md §8f rank 4;
hand-written input-feed LSTM / Luong attention kernels would be the second.

Structure restated from the reference (which is not importable here and whose joint step is broken as shipped, SURVEY.md
F2/F3 -- parity unpinned; the masked bi-LSTM below is pinned against torch's packed nn.LSTM instead, tests/test_pivot.py):
  Embeddings      models/NMT_Models.py:27-72     lookup -> Linear -> ReLU on the encoder side, plain lookup in the decoder
  Encoder         models/NMT_Models.py:75-135    bi-LSTM over the (packed) source sentences
  Decoder         models/NMT_Models.py:137-271   input-feed stacked LSTM, one step per target token
  GlobalAttention misc/OpenNMT-py-dalegebit/onmt/modules/GlobalAttention.py:84-177   "general" score, softmax over the source, tanh(W [c; h])
  loss            misc/criterion.py:126-136,181-205   generator + NLL summed over the non-pad target tokens
Sizes (models/nmt/readme.md): rnn 512, word vectors 512, 2 layers, src / tgt vocab ~12k / 8.6k.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

PAD, BOS, EOS = 0, 2, 3   # misc/constants.py:2-5


class MaskedBiLSTM(nn.Module):
    """nn.LSTM(bidirectional=True) over padded (S, B, D) input with per-row lengths, computed step by step on sequences
    left-aligned per direction (the states after a row's last valid step are gathered, what the extra steps compute is
    never read): same outputs (zeros at padded positions) and final states as the packed cuDNN call, but with no
    host-side lengths -- the loop is identical for every batch of the same (S, B), so it can live in a CUDA graph.
    Holds an nn.LSTM for its parameters (state_dict compatible with the reference's encoder.rnn)."""

    def __init__(self, input_size, hidden_size, num_layers, dropout):
        super().__init__()
        self.rnn = nn.LSTM(input_size, hidden_size, num_layers=num_layers, dropout=dropout, bidirectional=True)
        self.hidden_size, self.num_layers, self.dropout = hidden_size, num_layers, dropout

    def _layer(self, x, rev_idx, layer):
        """One bidirectional layer over padded (S, B, D) input.  The captured step is bound by its NUMBER of kernels, so
        (i) the input projections of all time steps and BOTH directions are one GEMM, (ii) the two directions advance
        together -- one batched matmul with the two recurrent weights and one fused cell over 2B rows per step -- and
        (iii) nothing is masked per step: the backward direction's inputs are gathered so that row b's sentence runs
        len_b-1 .. 0 over the steps 0 .. len_b-1 (left-aligned like the forward one), both directions run unmasked past a
        row's end, and the caller zeroes / re-orders the outputs once.  2 kernels per step where the per-direction torch
        LSTM cell + freezing `where`s took 12."""
        rnn, Hh = self.rnn, self.hidden_size
        S, B, _ = x.shape
        p = lambda name, rev: getattr(rnn, f"{name}_l{layer}" + ("_reverse" if rev else ""))
        w_ih = torch.cat([p("weight_ih", False), p("weight_ih", True)], 0)                    # (8 Hh, D)
        bias = torch.cat([p("bias_ih", False) + p("bias_hh", False), p("bias_ih", True) + p("bias_hh", True)], 0)
        xw = F.linear(x, w_ih, bias).view(S, B, 2, 4 * Hh)
        xw_rev = xw[:, :, 1].gather(0, rev_idx.expand(S, B, 4 * Hh))                          # step s of row b = position len_b-1-s
        ig = torch.stack([xw[:, :, 0], xw_rev], 1)                                            # (S, 2, B, 4 Hh)
        w_hh = torch.stack([p("weight_hh", False).t(), p("weight_hh", True).t()], 0).to(ig.dtype)   # (2, Hh, 4 Hh)
        return _BiLstmLayerFn.apply(ig.reshape(S, 2 * B, 4 * Hh), w_hh)                       # (S, 2, B, Hh) each

    def forward(self, x, lengths):
        """x (S, B, D); lengths (B,) on x's device.  Returns memory (S, B, 2 hidden), (h_n, c_n) each (2 layers, B, hidden)."""
        S, B = x.size(0), x.size(1)
        steps = torch.arange(S, device=x.device)[:, None]
        valid = (steps < lengths[None, :]).unsqueeze(2)                                        # (S, B, 1)
        rev_idx = (lengths[None, :] - 1 - steps).clamp_min(0).unsqueeze(2)                     # (S, B, 1): an involution on a row's valid steps
        last = (lengths - 1).clamp_min(0).view(1, B, 1)
        h_n, c_n = [], []
        for layer in range(self.num_layers):
            hs, cs = self._layer(x, rev_idx, layer)
            Hh = hs.size(3)
            idx = last.expand(1, B, Hh)
            # final states of both directions: the state after a row's last valid step
            h_n += [hs[:, 0].gather(0, idx).squeeze(0), hs[:, 1].gather(0, idx).squeeze(0)]
            c_n += [cs[:, 0].gather(0, idx).squeeze(0), cs[:, 1].gather(0, idx).squeeze(0)]
            bw = hs[:, 1].gather(0, rev_idx.expand(S, B, Hh))                                  # back to position order
            x = torch.cat([hs[:, 0], bw], 2) * valid.to(hs.dtype)
            if layer + 1 < self.num_layers and self.dropout > 0:
                x = F.dropout(x, self.dropout, self.training)
        return x, (torch.stack(h_n), torch.stack(c_n))


class _BiLstmLayerFn(torch.autograd.Function):
    """The recurrence of one bidirectional layer (both directions per step, MaskedBiLSTM._layer) with a hand-written backward
    pass: autograd's per-step bookkeeping (a weight-gradient matmul + accumulation add per step, fan-out adds) was ~10
    kernels per step; here the backward step is the fused cell backward and one batched matmul for the state gradient, and
    the recurrent weights' gradient is ONE batched matmul over all S steps at the end.  No masks: both directions arrive
    LEFT-aligned (a row's sentence occupies steps 0 .. len-1 of its direction), so the steps past a row's end compute
    garbage that nothing reads.  ig (S, 2B, 4Hh) input gate sums (biases included), w_hh (2, Hh, 4Hh); returns the states
    after every step, h and c (S, 2, B, Hh)."""

    @staticmethod
    def forward(ctx, ig, w_hh):
        S, B2, G = ig.shape
        B, Hh = B2 // 2, w_hh.size(1)
        h = ig.new_zeros(2, B, Hh)
        c = ig.new_zeros(B2, Hh)
        hs, cs, work = [], [], []
        ig_s = ig.unbind(0)
        for s in range(S):
            h_new, c_new, wk = _cell_fwd(ig_s[s], torch.bmm(h, w_hh).view(B2, G), c)
            work.append(wk)
            h, c = h_new.view(2, B, Hh), c_new
            hs.append(h)
            cs.append(c)
        hs, cs = torch.stack(hs), torch.stack(cs)                                              # (S, 2, B, Hh), (S, 2B, Hh)
        ctx.save_for_backward(w_hh, hs, cs, torch.stack(work))
        return hs, cs.view(S, 2, B, Hh)

    @staticmethod
    def backward(ctx, dhs, dcs):
        w_hh, hs, cs, work = ctx.saved_tensors
        S, _, B, Hh = hs.shape
        B2, G = 2 * B, 4 * Hh
        w_t = w_hh.transpose(1, 2)
        dc_next = None
        dgs = [None] * S
        dhs_s, dcs_s = dhs.reshape(S, B2, Hh).unbind(0), dcs.reshape(S, B2, Hh).unbind(0)
        c0 = cs.new_zeros(B2, Hh)
        dh = dhs_s[S - 1]
        for s in range(S - 1, -1, -1):
            dc = dcs_s[s] if dc_next is None else dcs_s[s] + dc_next
            dg, dc_next = _cell_bwd(dh, dc, cs[s - 1] if s > 0 else c0, cs[s], work[s])
            dgs[s] = dg
            if s > 0:   # gradient of the previous step's state: its own output gradient + what flows back through the recurrence
                dh = torch.baddbmm(dhs_s[s - 1].view(2, B, Hh), dg.view(2, B, G), w_t).view(B2, Hh)
        dig = torch.stack(dgs)                                                                 # (S, 2B, 4Hh)
        # d w_hh[d] = sum over steps and rows of h_{s-1}^T dgates_s (h_{-1} = 0): one batched matmul with K = (S-1) B
        hp = hs[:S - 1].permute(1, 3, 0, 2).reshape(2, Hh, (S - 1) * B)
        dw = torch.bmm(hp, dig[1:].view(S - 1, 2, B, G).permute(1, 0, 2, 3).reshape(2, (S - 1) * B, G)) if S > 1 else torch.zeros_like(w_hh)
        return dig, dw.to(w_hh.dtype)


def _cell_fwd(ig, hg, c):
    """LSTM cell from gate sums -> (h, c_new, workspace for _cell_bwd)."""
    if ig.is_cuda:
        return torch.ops.aten._thnn_fused_lstm_cell(ig, hg, c)
    i, f, g, o = (ig + hg).chunk(4, 1)
    i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
    c_new = f * c + i * g
    return o * torch.tanh(c_new), c_new, torch.stack([i, f, g, o])


def _cell_bwd(dh, dc, c_prev, c_new, wk):
    """-> (d gate sums (B, 4H), d c_prev); dc may be None."""
    if dh.is_cuda:
        dg, dc_prev, _ = torch.ops.aten._thnn_fused_lstm_cell_backward_impl(dh, dc, c_prev, c_new, wk, False)
        return dg, dc_prev
    i, f, g, o = wk.unbind(0)
    tc = torch.tanh(c_new)
    dcn = dh * o * (1 - tc * tc)
    if dc is not None:
        dcn = dcn + dc
    dg = torch.cat([dcn * g * i * (1 - i), dcn * c_prev * f * (1 - f), dcn * i * (1 - g * g), dh * tc * o * (1 - o)], 1)
    return dg, dcn * f


class _InputFeedDecoderFn(torch.autograd.Function):
    """The per-token loop of the input-feed decoder (models/NMT_Models.py:209-262: stacked LSTM cells on [emb_t | feed_{t-1}],
    "general" global attention, feed_t = dropout(tanh(W_out [ctx_t | h_t]))) with a hand-written backward pass.  Left to
    autograd, every token step re-derives a weight gradient per weight use (five small GEMMs) and accumulates it (thirteen
    adds per step with the fan-out sums); here the backward step only propagates the state gradients, and the gradients of
    the weights, of the attention keys and of the memory bank are ONE (batched) matmul each over all T steps at the end.
    Inputs are already in the compute dtype.  eg (T, B, 4d): word half of layer 0's gate sums (biases in); w_fh (4d, 2d) on
    [feed | h0]; keys / memory (B, S, d); neg (B, S) additive score mask; w_out (d, 2d); h0s / c0s (L, B, d) initial
    states; then (w_cat_i (4d, 2d) on [x | h_i], b_i (4d)) for the layers i >= 1.  Returns feeds (T, B, d)."""

    @staticmethod
    def forward(ctx, p_drop, training, eg, w_fh, keys, memory, neg, w_out, h0s, c0s, *upper):
        T, B, _ = eg.shape
        L = h0s.size(0)
        drop = training and p_drop > 0
        h, c = list(h0s.unbind(0)), list(c0s.unbind(0))
        feed = eg.new_zeros(B, h0s.size(2))
        zero_gates = eg.new_zeros(B, eg.size(2))
        sv = {k: [] for k in ("cat0", "c0p", "c0n", "wk0", "a", "x", "cato", "y", "mo")}
        sv_up = [{k: [] for k in ("cat", "cp", "cn", "wk", "m")} for _ in range(L - 1)]
        outs = []
        eg_s = eg.unbind(0)
        for t in range(T):
            cat0 = torch.cat([feed, h[0]], 1)
            h_new, c_new, wk = _cell_fwd(eg_s[t], cat0 @ w_fh.t(), c[0])
            sv["cat0"].append(cat0); sv["c0p"].append(c[0]); sv["c0n"].append(c_new); sv["wk0"].append(wk)
            h[0], c[0] = h_new, c_new
            x = h_new
            for i in range(1, L):
                w_cat, b = upper[2 * (i - 1)], upper[2 * (i - 1) + 1]
                m = None
                if drop:
                    x, m = torch.ops.aten.native_dropout(x, p_drop, True)
                cat_i = torch.cat([x, h[i]], 1)
                h_new, c_new, wk = _cell_fwd(torch.addmm(b, cat_i, w_cat.t()), zero_gates, c[i])   # gate sums complete in the input half
                u = sv_up[i - 1]
                u["cat"].append(cat_i); u["cp"].append(c[i]); u["cn"].append(c_new); u["wk"].append(wk); u["m"].append(m)
                h[i], c[i] = h_new, c_new
                x = h_new
            a = torch.softmax(torch.baddbmm(neg.unsqueeze(2), keys, x.unsqueeze(2)).squeeze(2), 1)
            cx = torch.bmm(a.unsqueeze(1), memory).squeeze(1)
            cato = torch.cat([cx, x], 1)
            y = torch.tanh(cato @ w_out.t())
            mo = None
            feed = y
            if drop:
                feed, mo = torch.ops.aten.native_dropout(y, p_drop, True)
            sv["a"].append(a); sv["x"].append(x); sv["cato"].append(cato); sv["y"].append(y); sv["mo"].append(mo)
            outs.append(feed)
        ctx.p_drop, ctx.drop, ctx.L = p_drop, drop, L
        ctx.sv, ctx.sv_up = sv, sv_up
        ctx.save_for_backward(w_fh, keys, memory, w_out, *upper)
        return torch.stack(outs)

    @staticmethod
    def backward(ctx, d_outs):
        w_fh, keys, memory, w_out, *upper = ctx.saved_tensors
        sv, sv_up, L, drop = ctx.sv, ctx.sv_up, ctx.L, ctx.drop
        scale = 1.0 / (1.0 - ctx.p_drop) if drop else 1.0
        T = d_outs.size(0)
        d = w_out.size(0)
        d_feed = None
        dh = [None] * L
        dc = [None] * L
        dg0_l, dy_l, ds_l, dcx_l = [None] * T, [None] * T, [None] * T, [None] * T
        dg_up = [[None] * T for _ in range(L - 1)]
        do_s = d_outs.unbind(0)
        for t in range(T - 1, -1, -1):
            df = do_s[t] if d_feed is None else do_s[t] + d_feed
            if drop:
                df = torch.ops.aten.native_dropout_backward(df, sv["mo"][t], scale)
            dy = torch.ops.aten.tanh_backward(df, sv["y"][t])
            dy_l[t] = dy
            dcat = dy @ w_out                                            # (B, 2d): [d ctx | d x]
            dcx, dx = dcat[:, :d], dcat[:, d:]
            dcx_l[t] = dcx
            a = sv["a"][t]
            da = torch.bmm(memory, dcx.unsqueeze(2)).squeeze(2)         # (B, S)
            ds = torch.ops.aten._softmax_backward_data(da, a, 1, a.dtype)
            ds_l[t] = ds
            dx = torch.baddbmm(dx.unsqueeze(1), ds.unsqueeze(1), keys).squeeze(1)
            for i in range(L - 1, 0, -1):
                u = sv_up[i - 1]
                dhi = dx if dh[i] is None else dx + dh[i]
                dg, dc[i] = _cell_bwd(dhi, dc[i], u["cp"][t], u["cn"][t], u["wk"][t])
                dg_up[i - 1][t] = dg
                dcat_i = dg @ upper[2 * (i - 1)]                         # (B, 2d): [d x | d h_i]
                dx, dh[i] = dcat_i[:, :d], dcat_i[:, d:]
                if drop:
                    dx = torch.ops.aten.native_dropout_backward(dx, u["m"][t], scale)
            dh0 = dx if dh[0] is None else dx + dh[0]
            dg0, dc[0] = _cell_bwd(dh0, dc[0], sv["c0p"][t], sv["c0n"][t], sv["wk0"][t])
            dg0_l[t] = dg0
            dcat0 = dg0 @ w_fh                                           # (B, 2d): [d feed | d h0]
            d_feed, dh[0] = dcat0[:, :d], dcat0[:, d:]
        B = d_outs.size(1)
        d_eg = torch.stack(dg0_l)                                        # (T, B, 4d)
        flat = lambda lst: torch.stack(lst).reshape(T * B, -1)
        d_w_fh = d_eg.reshape(T * B, -1).t() @ flat(sv["cat0"])
        d_w_out = flat(dy_l).t() @ flat(sv["cato"])
        ds_all = torch.stack(ds_l, 2)                                    # (B, S, T)
        d_keys = torch.bmm(ds_all, torch.stack(sv["x"], 1))              # (B, S, T) x (B, T, d)
        d_memory = torch.bmm(torch.stack(sv["a"], 2), torch.stack(dcx_l, 1))
        zero = lambda t_: torch.zeros_like(t_)
        d_h0s = torch.stack([dh[i] if dh[i] is not None else zero(sv["c0p"][0]) for i in range(L)])
        d_c0s = torch.stack([dc[i] if dc[i] is not None else zero(sv["c0p"][0]) for i in range(L)])
        d_upper = []
        for i in range(1, L):
            dg_all = torch.stack(dg_up[i - 1]).reshape(T * B, -1)
            d_upper += [dg_all.t() @ flat(sv_up[i - 1]["cat"]), dg_all.sum(0)]
        return (None, None, d_eg, d_w_fh, d_keys, d_memory, None, d_w_out, d_h0s, d_c0s, *d_upper)


def _lstm_from_gates(igates, hgates, c):
    """LSTM cell from pre-computed input / hidden gate sums (biases included): torch's fused CUDA cell (one pointwise kernel,
    forward and backward), the gate formulas spelled out on the CPU (tests)."""
    if igates.is_cuda:
        h, c2, _ = torch.ops.aten._thnn_fused_lstm_cell(igates, hgates, c)
        return h, c2
    i, f, g, o = (igates + hgates).chunk(4, 1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c2), c2


class PivotNMT(nn.Module):
    def __init__(self, src_vocab=12000, tgt_vocab=8600, dim=512, layers=2, dropout=0.3):
        super().__init__()
        self.dim, self.layers = dim, layers
        self.src_lut = nn.Embedding(src_vocab, dim, padding_idx=PAD)
        self.src_mlp = nn.Linear(dim, dim)
        self.encoder = MaskedBiLSTM(dim, dim // 2, layers, dropout)
        self.tgt_lut = nn.Embedding(tgt_vocab, dim, padding_idx=PAD)
        self.cells = nn.ModuleList([nn.LSTMCell(2 * dim if i == 0 else dim, dim) for i in range(layers)])
        self.attn_in = nn.Linear(dim, dim, bias=False)
        self.attn_out = nn.Linear(2 * dim, dim, bias=False)
        self.drop = nn.Dropout(dropout)
        self.generator = nn.Linear(dim, tgt_vocab)

    def forward(self, src, src_len, tgt):
        """src (S, B), src_len (B,) on the device, tgt (T, B) with BOS first; returns (summed NLL, number of target tokens)."""
        emb = F.relu(self.src_mlp(self.src_lut(src)))
        memory, (h, c) = self.encoder(emb, src_len)
        memory = memory.transpose(0, 1)                                                    # (B, S, dim)
        fix = lambda s: torch.cat([s[0::2], s[1::2]], 2)                                  # _fix_enc_hidden :284-288
        h, c = list(fix(h)), list(fix(c))
        mask = torch.arange(memory.size(1), device=src.device)[None, :] >= src_len[:, None]
        feed = memory.new_zeros(src.size(1), self.dim)                                     # zero input feed :289-295
        keys = self.attn_in(memory)                                                        # "general" score h^T W m
        tgt_emb = self.tgt_lut(tgt[:-1])
        # The per-token loop replays from a CUDA graph and is launch-bound, so: the word half of the first layer's input
        # projection is one GEMM over all T steps; per step the input-feed half and the recurrent projection are ONE GEMM
        # over [feed | h0] (layers above: over [x | h_i]); the padding mask enters the scores as the additive term of
        # baddbmm; and the loop runs under _InputFeedDecoderFn, whose backward pass keeps every weight gradient out of it.
        c0 = self.cells[0]
        e_gates = F.linear(tgt_emb, c0.weight_ih[:, :self.dim], c0.bias_ih + c0.bias_hh)   # (T-1, B, 4 dim)
        dt = e_gates.dtype                                                                 # compute dtype (bf16 under autocast)
        w_fh = torch.cat([c0.weight_ih[:, self.dim:], c0.weight_hh], 1).to(dt)             # (4 dim, 2 dim): [feed | h0]
        upper = []
        for i in range(1, self.layers):
            ci = self.cells[i]
            upper += [torch.cat([ci.weight_ih, ci.weight_hh], 1).to(dt), (ci.bias_ih + ci.bias_hh).to(dt)]
        neg = torch.zeros(mask.shape, dtype=dt, device=src.device).masked_fill(mask, float("-inf"))
        with torch.autocast(src.device.type, enabled=False):
            feeds = _InputFeedDecoderFn.apply(float(self.drop.p), self.training, e_gates, w_fh, keys.to(dt), memory.to(dt).contiguous(), neg,
                                              self.attn_out.weight.to(dt), torch.stack(h).to(dt), torch.stack(c).to(dt), *upper)
        outs = feeds.unbind(0)
        logp = F.log_softmax(self.generator(torch.stack(outs)), -1)
        gold = tgt[1:]
        nll = F.nll_loss(logp.view(-1, logp.size(-1)), gold.reshape(-1), ignore_index=PAD, reduction="sum")
        return nll, (gold != PAD).sum()


def sentences(batch, vocab, gen, lo=5, hi=30, bos=None, max_len=None):
    """Synthetic sentences, lengths ~U{lo..hi} sorted descending (the loader sorts by source length); (T, batch) int64 padded
    to max_len (default: the longest) + the lengths."""
    n = torch.randint(lo, hi + 1, (batch,), generator=gen)
    n, _ = torch.sort(n, descending=True)
    T = (int(n.max()) if max_len is None else max_len) + (2 if bos is not None else 0)
    x = torch.full((T, batch), PAD, dtype=torch.int64)
    for b in range(batch):
        words = torch.randint(4, vocab, (int(n[b]),), generator=gen)
        if bos is not None:
            x[0, b], x[1:1 + int(n[b]), b], x[1 + int(n[b]), b] = bos, words, EOS
        else:
            x[:int(n[b]), b] = words
    return x, n


class PivotTrainStep:
    """One translator training step (forward + NLL / tokens + backward + gradient all-reduce + clip 5.0 + Adam) on static
    device buffers `src (S, B)`, `src_len (B,)`, `tgt (T, B)`; `graph=True` captures it into one CUDA graph (call
    `load(src, src_len, tgt)` to refresh the buffers, then `step()`); `amp=True` runs the matmuls in bf16 autocast."""

    def __init__(self, model, S, T, B, lr=1e-3, clip=5.0, graph=True, device="cuda", batch=None, amp=True):
        import torch.distributed as dist
        self.model, self.clip, self.amp = model, clip, amp
        self.src = torch.zeros(S, B, dtype=torch.int64, device=device)
        self.src_len = torch.ones(B, dtype=torch.int64, device=device)
        self.tgt = torch.zeros(T, B, dtype=torch.int64, device=device)
        params = [p for p in model.parameters() if p.requires_grad]
        self.flat = torch.zeros(sum(p.numel() for p in params), device=device)
        off = 0
        for p in params:          # gradients live in one flat buffer: a single all-reduce, a single norm
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.optim = torch.optim.Adam(params, lr=lr, fused=True, capturable=True)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.loss = torch.zeros((), device=device)
        self.graph = None
        if batch is not None:     # (the warm-up steps of the capture below run on this batch: they are real updates)
            self.load(*batch)
        if graph:
            if batch is None:
                raise ValueError("PivotTrainStep: pass the first batch (batch=(src, src_len, tgt)) -- the capture warms up on it")
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._eager()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            from .engine import _no_gc
            self.graph = torch.cuda.CUDAGraph()
            with _no_gc():
                with torch.cuda.graph(self.graph):
                    self._eager()

    def _eager(self):
        import torch.distributed as dist
        self.flat.zero_()
        # amp: bf16 operands on the tensor cores with fp32 accumulation (the precision of the decoder's own GEMMs); in fp32 the
        # step is bound by SIMT matmuls (24 M parameters, ~600 GFLOP per 256-sentence step)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
            nll, n = self.model(self.src, self.src_len, self.tgt)
        nll = nll.float()
        if self.world > 1:      # normalise by the token count of the global batch
            n = n.float()
            dist.all_reduce(n)
        loss = nll / n
        loss.backward()
        if self.world > 1:
            dist.all_reduce(self.flat)
        norm = self.flat.norm()
        self.flat.mul_(torch.clamp(self.clip / (norm + 1e-6), max=1.0))
        self.optim.step()
        self.loss.copy_(loss.detach())

    def load(self, src, src_len, tgt):
        self.src.copy_(src, non_blocking=True)
        self.src_len.copy_(src_len, non_blocking=True)
        self.tgt.copy_(tgt, non_blocking=True)

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._eager()
        return self.loss
