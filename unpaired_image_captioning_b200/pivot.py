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

    def _direction(self, x, valid, layer, reverse):
        sfx = f"_l{layer}" + ("_reverse" if reverse else "")
        w_ih, w_hh = getattr(self.rnn, "weight_ih" + sfx), getattr(self.rnn, "weight_hh" + sfx)
        b_ih, b_hh = getattr(self.rnn, "bias_ih" + sfx), getattr(self.rnn, "bias_hh" + sfx)
        S, B, _ = x.shape
        h = x.new_zeros(B, self.hidden_size)
        c = x.new_zeros(B, self.hidden_size)
        outs = [None] * S
        for t in (range(S - 1, -1, -1) if reverse else range(S)):
            # torch's fused LSTM cell (two GEMMs + one pointwise kernel, forward and backward): the captured graph is bound
            # by its NUMBER of kernels, not by their arithmetic (a first version with the gate math spelled out in elementwise
            # ops replayed ~10 000 tiny kernels per step)
            h_new, c_new = torch._VF.lstm_cell(x[t], (h, c), w_ih, w_hh, b_ih, b_hh)
            m = valid[t]
            h, c = torch.where(m, h_new, h), torch.where(m, c_new, c)    # rows past their end keep their state
            outs[t] = h * m
        return torch.stack(outs), h, c

    def forward(self, x, lengths):
        """x (S, B, D); lengths (B,) on x's device.  Returns memory (S, B, 2 hidden), (h_n, c_n) each (2 layers, B, hidden)."""
        S = x.size(0)
        valid = (torch.arange(S, device=x.device)[:, None] < lengths[None, :]).unsqueeze(2)    # (S, B, 1)
        hs, cs = [], []
        for layer in range(self.num_layers):
            fw, hf, cf = self._direction(x, valid, layer, False)
            bw, hb, cb = self._direction(x, valid, layer, True)
            x = torch.cat([fw, bw], 2)
            if layer + 1 < self.num_layers and self.dropout > 0:
                x = F.dropout(x, self.dropout, self.training)
            hs += [hf, hb]
            cs += [cf, cb]
        return x, (torch.stack(hs), torch.stack(cs))


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
        for t in range(tgt_emb.size(0)):                                                   # one step per token :209-262
            x = torch.cat([tgt_emb[t], feed], 1)
            for i, cell in enumerate(self.cells):
                h[i], c[i] = cell(x, (h[i], c[i]))
                x = self.drop(h[i]) if i + 1 < self.layers else h[i]
            score = torch.bmm(keys, x.unsqueeze(2)).squeeze(2).masked_fill(mask, float("-inf"))
            ctx = torch.bmm(F.softmax(score, 1).unsqueeze(1), memory).squeeze(1)
            feed = self.drop(torch.tanh(self.attn_out(torch.cat([ctx, x], 1))))
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
