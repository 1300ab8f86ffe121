"""BASELINE.json configs[3] / SURVEY.md §8(d) cfg 4: the TopDown decoder training step (this library, B = 256 per GPU)
next to a PyTorch restatement of the pivot translator's training step, to report decoder-only and joint samples/s and
the NMT share of the step time (which decides whether NMT kernels become "next").

The onmt translator is NOT part of the hot path: north_star keeps it in PyTorch, the reference's copy is not importable
here and its joint step is broken as shipped, so there is nothing to pin it against -- "parity unpinned".
The restatement below follows the reference's structure only (plain torch modules, cuDNN LSTM + cuBLAS):
  Embeddings      models/NMT_Models.py:27-72    lookup -> Linear -> ReLU on the encoder side, plain lookup in the decoder
  Encoder         models/NMT_Models.py:75-135   bi-LSTM over packed source sentences
  Decoder         models/NMT_Models.py:137-271  input-feed stacked LSTM, one python step per target token
  GlobalAttention onmt/modules/GlobalAttention.py:84-177   Luong "general" attention + tanh(W [c; h])
  generator       Linear + log-softmax, NLL over the non-pad target tokens
Sizes: rnn 512, word vectors 512, 2 layers, src / tgt vocab 12k / 8.6k (models/nmt/readme.md), sentence length ~U{5..30}.

Run on the GPU box:  python scripts/cfg4_joint.py [--steps 20 --warmup 3]  -> one JSON line (also gpurun_out/cfg4_joint.json).
"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

PAD = 0


class PivotNMT(nn.Module):
    def __init__(self, src_vocab=12000, tgt_vocab=8600, dim=512, layers=2, dropout=0.3):
        super().__init__()
        self.dim, self.layers = dim, layers
        self.src_lut = nn.Embedding(src_vocab, dim, padding_idx=PAD)
        self.src_mlp = nn.Linear(dim, dim)
        self.encoder = nn.LSTM(dim, dim // 2, num_layers=layers, dropout=dropout, bidirectional=True)
        self.tgt_lut = nn.Embedding(tgt_vocab, dim, padding_idx=PAD)
        self.cells = nn.ModuleList([nn.LSTMCell(2 * dim if i == 0 else dim, dim) for i in range(layers)])
        self.attn_in = nn.Linear(dim, dim, bias=False)
        self.attn_out = nn.Linear(2 * dim, dim, bias=False)
        self.drop = nn.Dropout(dropout)
        self.generator = nn.Linear(dim, tgt_vocab)

    def forward(self, src, src_len, tgt):
        """src (S, B), src_len (B,) sorted descending, tgt (T, B) with BOS first; returns the summed NLL and the token count."""
        emb = F.relu(self.src_mlp(self.src_lut(src)))
        packed = nn.utils.rnn.pack_padded_sequence(emb, src_len.cpu())
        memory, (h, c) = self.encoder(packed)
        memory = nn.utils.rnn.pad_packed_sequence(memory)[0].transpose(0, 1)              # (B, S, dim)
        fix = lambda s: torch.cat([s[0::2], s[1::2]], 2)                                  # _fix_enc_hidden :284-288
        h, c = list(fix(h)), list(fix(c))
        mask = torch.arange(memory.size(1), device=src.device)[None, :] >= src_len[:, None].to(src.device)
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


def sentences(batch, vocab, gen, lo=5, hi=30, bos=None):
    n = torch.randint(lo, hi + 1, (batch,), generator=gen)
    n, _ = torch.sort(n, descending=True)
    T = int(n.max()) + (2 if bos is not None else 0)
    x = torch.full((T, batch), PAD, dtype=torch.int64)
    for b in range(batch):
        words = torch.randint(4, vocab, (int(n[b]),), generator=gen)
        if bos is not None:
            x[0, b], x[1:1 + int(n[b]), b], x[1 + int(n[b]), b] = bos, words, 3
        else:
            x[:int(n[b]), b] = words
    return x, n


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    from unpaired_image_captioning_b200 import train_bench

    # ---- decoder step: this library (TopDown, 36 regions, B = 256 per GPU; fwd + XE + BPTT + clip + Adam, CUDA graph) ----
    st = train_bench._setup(0, 0, cfg_name="cfg4")
    dec_step = train_bench.GraphedTrainStep(st)
    dec_ms = timed(dec_step, args.steps, args.warmup)

    # ---- pivot translator step: PyTorch restatement (fwd + NLL + bwd + clip + Adam) ----
    torch.manual_seed(1234)
    gen = torch.Generator().manual_seed(1234)
    nmt = PivotNMT().cuda().train()
    optim = torch.optim.Adam(nmt.parameters(), lr=1e-3, fused=True)
    src, src_len = sentences(args.batch, 12000, gen)
    tgt, _ = sentences(args.batch, 8600, gen, bos=2)
    src, tgt = src.cuda(), tgt.cuda()

    def nmt_step():
        optim.zero_grad(set_to_none=True)
        nll, n = nmt(src, src_len, tgt)
        (nll / n).backward()
        torch.nn.utils.clip_grad_norm_(nmt.parameters(), 5.0)
        optim.step()

    nmt_ms = timed(nmt_step, args.steps, args.warmup)

    def joint():
        dec_step()
        nmt_step()

    joint_ms = timed(joint, args.steps, args.warmup)
    B = args.batch
    out = {"workload": "configs[3] (cfg 4): TopDown decoder step + PyTorch pivot-translator step, batch %d on one B200" % B,
           "decoder_ms": round(dec_ms, 3), "decoder_samples_per_s": round(B / dec_ms * 1e3, 1),
           "nmt_torch_ms": round(nmt_ms, 3), "nmt_samples_per_s": round(B / nmt_ms * 1e3, 1),
           "joint_ms": round(joint_ms, 3), "joint_samples_per_s": round(B / joint_ms * 1e3, 1),
           "nmt_share_of_joint_step": round(nmt_ms / (nmt_ms + dec_ms), 3),
           "src_len_max": int(src.size(0)), "tgt_len_max": int(tgt.size(0)), "nmt_params_M": round(sum(p.numel() for p in nmt.parameters()) / 1e6, 1),
           "note": "NMT = plain PyTorch restatement (eager, cuDNN LSTM encoder, python-loop input-feed decoder); parity unpinned (SURVEY F2/F3)"}
    line = json.dumps(out)
    print(line)
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/cfg4_joint.json", "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
