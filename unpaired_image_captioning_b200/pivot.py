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
    """nn.LSTM(bidirectional=True) over padded (S, B, D) input with per-row lengths, computed step by step with frozen
    states past a row's end: same outputs (zeros at padded positions) and final states as the packed cuDNN call, but with no
    host-side lengths -- the loop is identical for every batch of the same (S, B), so it can live in a CUDA graph.
    Holds an nn.LSTM for its parameters (state_dict compatible with the reference's encoder.rnn)."""

    def __init__(self, input_size, hidden_size, num_layers, dropout):
        super().__init__()
        self.rnn = nn.LSTM(input_size, hidden_size, num_layers=num_layers, dropout=dropout, bidirectional=True)
        self.hidden_size, self.num_layers, self.dropout = hidden_size, num_layers, dropout

    def _layer(self, x, keep, layer):
        """One bidirectional layer over padded (S, B, D) input.  The captured step is bound by its NUMBER of kernels, so
        (i) the input projections of all time steps and BOTH directions are one GEMM, (ii) the two directions advance
        together -- step s handles position s forward and S-1-s backward: one batched matmul with the two recurrent
        weights and one fused cell over 2B rows -- and (iii) only the backward direction is masked per step (its state
        must stay zero until its sentence starts); the forward direction runs unmasked past a row's end: its outputs
        there are zeroed once afterwards and its final state is gathered at position len-1.  4 kernels per step where
        the per-direction torch LSTM cell + freezing `where`s took 12."""
        rnn, Hh = self.rnn, self.hidden_size
        S, B, _ = x.shape
        p = lambda name, rev: getattr(rnn, f"{name}_l{layer}" + ("_reverse" if rev else ""))
        w_ih = torch.cat([p("weight_ih", False), p("weight_ih", True)], 0)                    # (8 Hh, D)
        bias = torch.cat([p("bias_ih", False) + p("bias_hh", False), p("bias_ih", True) + p("bias_hh", True)], 0)
        xw = F.linear(x, w_ih, bias).view(S, B, 2, 4 * Hh)
        ig = torch.stack([xw[:, :, 0], xw[:, :, 1].flip(0)], 1)                               # (S, 2, B, 4 Hh): backward time-flipped
        w_hh = torch.stack([p("weight_hh", False).t(), p("weight_hh", True).t()], 0).to(ig.dtype)   # (2, Hh, 4 Hh)
        return _BiLstmLayerFn.apply(ig.reshape(S, 2 * B, 4 * Hh), w_hh, keep.to(ig.dtype))       # (S, 2, B, Hh) each

    def forward(self, x, lengths):
        """x (S, B, D); lengths (B,) on x's device.  Returns memory (S, B, 2 hidden), (h_n, c_n) each (2 layers, B, hidden)."""
        S, B = x.size(0), x.size(1)
        valid = (torch.arange(S, device=x.device)[:, None] < lengths[None, :]).unsqueeze(2)    # (S, B, 1)
        keep = torch.stack([torch.ones_like(valid), valid.flip(0)], 1)                         # (S, 2, B, 1)
        last = (lengths - 1).clamp_min(0).view(1, B, 1)
        h_n, c_n = [], []
        for layer in range(self.num_layers):
            hs, cs = self._layer(x, keep, layer)
            idx = last.expand(1, B, hs.size(3))
            # final states: forward = the state at a row's last valid position, backward = the state after the last step
            h_n += [hs[:, 0].gather(0, idx).squeeze(0), hs[S - 1, 1]]
            c_n += [cs[:, 0].gather(0, idx).squeeze(0), cs[S - 1, 1]]
            x = torch.cat([hs[:, 0] * valid.to(hs.dtype), hs[:, 1].flip(0)], 2)
            if layer + 1 < self.num_layers and self.dropout > 0:
                x = F.dropout(x, self.dropout, self.training)
        return x, (torch.stack(h_n), torch.stack(c_n))


class _BiLstmLayerFn(torch.autograd.Function):
    """The recurrence of one bidirectional layer (both directions per step, MaskedBiLSTM._layer) with a hand-written backward
    pass: autograd's per-step bookkeeping (a weight-gradient matmul + accumulation add per step, fan-out adds, mask
    multiplies and their gradients) was ~10 kernels per step; here the backward step is mask, fused cell backward, one batched
    matmul for the state gradient, and the recurrent weights' gradient is ONE batched matmul over all S steps at the end.
    ig (S, 2B, 4Hh) input gate sums (biases included), w_hh (2, Hh, 4Hh), keep (S, 2, B, 1); returns h, c (S, 2, B, Hh)."""

    @staticmethod
    def forward(ctx, ig, w_hh, keep):
        S, B2, G = ig.shape
        B, Hh = B2 // 2, w_hh.size(1)
        h = ig.new_zeros(2, B, Hh)
        c = ig.new_zeros(2, B, Hh)
        hs, cs, h_prev, c_prev, c_new_l, work = [], [], [], [], [], []
        fused = ig.is_cuda
        ig_s, keep_s = ig.unbind(0), keep.unbind(0)
        for s in range(S):
            hg = torch.bmm(h, w_hh).view(B2, G)
            h_prev.append(h)
            c_prev.append(c)
            if fused:
                h_new, c_new, wk = torch.ops.aten._thnn_fused_lstm_cell(ig_s[s], hg, c.view(B2, Hh))
            else:
                i, f, g, o = (ig_s[s] + hg).chunk(4, 1)
                i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
                c_new = f * c.view(B2, Hh) + i * g
                h_new = o * torch.tanh(c_new)
                wk = torch.stack([i, f, g, o])
            c_new_l.append(c_new)
            work.append(wk)
            h = h_new.view(2, B, Hh) * keep_s[s]
            c = c_new.view(2, B, Hh) * keep_s[s]
            hs.append(h)
            cs.append(c)
        ctx.fused = fused
        ctx.save_for_backward(w_hh, keep, torch.stack(h_prev), torch.stack(c_prev), torch.stack(c_new_l), torch.stack(work))
        return torch.stack(hs), torch.stack(cs)

    @staticmethod
    def backward(ctx, dhs, dcs):
        w_hh, keep, h_prev, c_prev, c_new, work = ctx.saved_tensors
        S, _, B, Hh = h_prev.shape
        B2, G = 2 * B, 4 * Hh
        w_t = w_hh.transpose(1, 2)
        dh_next = dc_next = None
        dgs = [None] * S
        dhs_s, dcs_s, keep_s = dhs.unbind(0), dcs.unbind(0), keep.unbind(0)
        for s in range(S - 1, -1, -1):
            dh = dhs_s[s] if dh_next is None else dhs_s[s] + dh_next
            dc = dcs_s[s] if dc_next is None else dcs_s[s] + dc_next
            dh_new = (dh * keep_s[s]).reshape(B2, Hh)
            dc_new = (dc * keep_s[s]).reshape(B2, Hh)
            if ctx.fused:
                dg, dc_prev, _ = torch.ops.aten._thnn_fused_lstm_cell_backward_impl(dh_new, dc_new, c_prev[s].view(B2, Hh), c_new[s], work[s], False)
            else:
                i, f, g, o = work[s].unbind(0)
                tc = torch.tanh(c_new[s])
                do = dh_new * tc
                dcn = dc_new + dh_new * o * (1 - tc * tc)
                dg = torch.cat([dcn * g * i * (1 - i), dcn * c_prev[s].view(B2, Hh) * f * (1 - f), dcn * i * (1 - g * g), do * o * (1 - o)], 1)
                dc_prev = dcn * f
            dgs[s] = dg
            dh_next = torch.bmm(dg.view(2, B, G), w_t)
            dc_next = dc_prev.view(2, B, Hh)
        dig = torch.stack(dgs)                                                                 # (S, 2B, 4Hh)
        # d w_hh[d] = sum over steps and rows of h_prev^T dgates: one batched matmul with K = S B
        hp = h_prev.permute(1, 3, 0, 2).reshape(2, Hh, S * B)
        dw = torch.bmm(hp, dig.view(S, 2, B, G).permute(1, 0, 2, 3).reshape(2, S * B, G))
        return dig, dw.to(w_hh.dtype), None


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
        outs = []
        tgt_emb = self.tgt_lut(tgt[:-1])
        # Kernel-count savings of the per-token loop (it replays from a CUDA graph and is launch-bound): the word half of the
        # first layer's input projection is one GEMM over all T steps; per step the input-feed half and the recurrent
        # projection are ONE GEMM over [feed | h0]; the padding mask enters the scores as the additive term of baddbmm.
        c0 = self.cells[0]
        e_gates = F.linear(tgt_emb, c0.weight_ih[:, :self.dim], c0.bias_ih + c0.bias_hh)   # (T-1, B, 4 dim)
        w_fh = torch.cat([c0.weight_ih[:, self.dim:], c0.weight_hh], 1)                    # (4 dim, 2 dim): [feed | h0]
        if torch.is_autocast_enabled():   # a non-leaf operand is re-cast by autocast at EVERY use (30 x 8 MB per step): cast once
            w_fh = w_fh.to(torch.get_autocast_dtype("cuda"))
        neg = torch.zeros(mask.shape, dtype=keys.dtype, device=src.device).masked_fill(mask, float("-inf")).unsqueeze(2)
        eg_steps = e_gates.unbind(0)
        for t in range(tgt_emb.size(0)):                                                   # one step per token :209-262
            hg = F.linear(torch.cat([feed, h[0]], 1), w_fh)
            h[0], c[0] = _lstm_from_gates(eg_steps[t], hg.to(e_gates.dtype), c[0].to(e_gates.dtype))
            x = h[0]
            for i in range(1, self.layers):
                x = self.drop(x)
                h[i], c[i] = self.cells[i](x, (h[i], c[i]))
                x = h[i]
            score = torch.baddbmm(neg, keys, x.unsqueeze(2).to(keys.dtype)).squeeze(2)
            ctx = torch.bmm(F.softmax(score, 1).unsqueeze(1).to(memory.dtype), memory).squeeze(1)
            feed = self.drop(torch.tanh(self.attn_out(torch.cat([ctx, x.to(ctx.dtype)], 1))))
            outs.append(feed)
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
