"""Deterministic synthetic inputs and option namespaces for the decoder hot path.

Mirrors what the reference's data loader hands to the model:
  * att_feats  (B, L, D) fp32, post-ReLU-like   (misc/dataloader/dataloader.py:274-280)
  * fc_feats   (B, F)    fp32, mean over regions (scripts/make_bu_data.py)
  * labels     (B, seq_length+2) int64, [0, w_1..w_n, 0...] (dataloader.py:223-224,252)
  * masks      (B, seq_length+2) fp32, ones over the first n+2 columns (dataloader.py:284-287)

Shared by tests, bench.py and the golden-fixture generator so that every arm sees
byte-identical inputs.  Pure CPU torch; no CUDA, no oracle import.
"""
from __future__ import annotations

import argparse

import torch

# The five workload shapes named by BASELINE.json `configs`.
CONFIGS = {
    # cfg 1: att2in2, 14x14 regions, vocab ~10k, seq 16, B=16 (the CPU-runnable case)
    "cfg1": dict(caption_model="att2in2", rnn_size=512, input_encoding_size=512, att_hid_size=512,
                 att_size=196, vocab_size=9999, seq_length=16, batch=16, beam_size=3),
    # cfg 2: same decoder, greedy + beam-3, B=256 on one B200
    "cfg2": dict(caption_model="att2in2", rnn_size=512, input_encoding_size=512, att_hid_size=512,
                 att_size=196, vocab_size=9999, seq_length=16, batch=256, beam_size=3),
    # cfg 3: TopDown, 36 regions, XE training, global batch 512
    "cfg3": dict(caption_model="topdown", rnn_size=512, input_encoding_size=512, att_hid_size=512,
                 att_size=36, vocab_size=9999, seq_length=16, batch=512, beam_size=3),
    # cfg 4: the decoder step of the pivot path (the onmt translator stays in PyTorch): TopDown, batch 256 per GPU
    "cfg4": dict(caption_model="topdown", rnn_size=512, input_encoding_size=512, att_hid_size=512,
                 att_size=36, vocab_size=9999, seq_length=16, batch=256, beam_size=3),
    # cfg 5: large-vocab stress, rnn 1024, vocab 30k, seq 20, beam 5
    "cfg5": dict(caption_model="att2in2", rnn_size=1024, input_encoding_size=512, att_hid_size=512,
                 att_size=196, vocab_size=29999, seq_length=20, batch=500, beam_size=5),
    # tiny shapes used by the golden fixtures and the smoke test
    "tiny_att2in2": dict(caption_model="att2in2", rnn_size=32, input_encoding_size=32, att_hid_size=32,
                         att_size=7, vocab_size=51, seq_length=6, batch=5, beam_size=3,
                         fc_feat_size=64, att_feat_size=64),
    "tiny_att2all2": dict(caption_model="att2all2", rnn_size=32, input_encoding_size=32, att_hid_size=32,
                          att_size=7, vocab_size=51, seq_length=6, batch=5, beam_size=3,
                          fc_feat_size=64, att_feat_size=64),
    "tiny_topdown": dict(caption_model="topdown", rnn_size=32, input_encoding_size=32, att_hid_size=32,
                         att_size=7, vocab_size=51, seq_length=6, batch=5, beam_size=3,
                         fc_feat_size=64, att_feat_size=64),
    "tiny_stackatt": dict(caption_model="stackatt", rnn_size=32, input_encoding_size=32, att_hid_size=32,
                          att_size=7, vocab_size=51, seq_length=6, batch=5, beam_size=3,
                          fc_feat_size=64, att_feat_size=64),
    "tiny_denseatt": dict(caption_model="denseatt", rnn_size=32, input_encoding_size=32, att_hid_size=32,
                          att_size=7, vocab_size=51, seq_length=6, batch=5, beam_size=3,
                          fc_feat_size=64, att_feat_size=64),
}


def make_opt(caption_model="att2in2", vocab_size=9999, rnn_size=512, input_encoding_size=512,
             att_hid_size=512, seq_length=16, fc_feat_size=2048, att_feat_size=2048,
             drop_prob_lm=0.0, use_bn=0, logit_layers=1, **_unused):
    """The subset of the reference's flat `opt` namespace the decoder reads (opts.py:41-56,
    models/AttModel.py:58-69,86)."""
    return argparse.Namespace(
        caption_model=caption_model, vocab_size=vocab_size, rnn_size=rnn_size,
        input_encoding_size=input_encoding_size, att_hid_size=att_hid_size,
        seq_length=seq_length, fc_feat_size=fc_feat_size, att_feat_size=att_feat_size,
        drop_prob_lm=drop_prob_lm, use_bn=use_bn, logit_layers=logit_layers,
        num_layers={"topdown": 2, "stackatt": 3, "denseatt": 3}.get(caption_model, 1))


def opt_for(name, **over):
    cfg = dict(CONFIGS[name])
    cfg.update(over)
    return make_opt(**cfg), cfg


def make_features(batch, att_size, att_feat_size=2048, seed=1234, normalise=False):
    g = torch.Generator().manual_seed(seed)
    att = torch.randn(batch, att_size, att_feat_size, generator=g).clamp_(min=0)
    if normalise:  # mirrors norm_att_feat (dataloader.py:310-311)
        att = att / att.norm(dim=2, keepdim=True).clamp_(min=1e-12)
    fc = att.mean(1)
    return fc, att


def make_captions(batch, seq_length, vocab_size, seed=1234, min_len=None):
    """labels/masks exactly as the loader builds them; words are 1..vocab_size-1 so the UNK
    index (vocab_size, the last logit) never appears as a target."""
    g = torch.Generator().manual_seed(seed + 7)
    lo = min(8, seq_length) if min_len is None else min_len
    lens = torch.randint(lo, seq_length + 1, (batch,), generator=g)
    labels = torch.zeros(batch, seq_length + 2, dtype=torch.int64)
    masks = torch.zeros(batch, seq_length + 2, dtype=torch.float32)
    for i in range(batch):
        n = int(lens[i])
        labels[i, 1:n + 1] = torch.randint(1, max(2, vocab_size), (n,), generator=g)
        masks[i, :n + 2] = 1.0
    return labels, masks


def make_att_masks(batch, att_size, seed=1234):
    """Ragged region counts (bottom-up features have 10-100 boxes, dataloader.py:274-280):
    row i keeps its first n_i regions, at least one row keeps all of them."""
    g = torch.Generator().manual_seed(seed + 13)
    n = torch.randint(max(1, att_size // 2), att_size + 1, (batch,), generator=g)
    n[0] = att_size
    m = (torch.arange(att_size)[None, :] < n[:, None]).float()
    return m


def init_state_dict(opt, seed=1234, peaked=0.0, eos_bias=0.0):
    """Random-init weights with the reference's key layout and torch's default
    initialisers (Linear/LSTMCell: U(+-1/sqrt(fan_in)); Embedding: N(0,1)).

    `peaked` scales logit.weight and `eos_bias` is added to logit.bias[0]: the variants SURVEY.md
    F6 / Appendix A ask for so that margins are wide and EOS paths are exercised."""
    g = torch.Generator().manual_seed(seed + 101)
    V, E, H, A = opt.vocab_size + 1, opt.input_encoding_size, opt.rnn_size, opt.att_hid_size
    F_, D = opt.fc_feat_size, opt.att_feat_size

    def lin(out_f, in_f, bound_in=None):
        k = 1.0 / (bound_in or in_f) ** 0.5
        w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * k
        b = (torch.rand(out_f, generator=g) * 2 - 1) * k
        return w, b

    sd = {}
    sd["embed.0.weight"] = torch.randn(V, E, generator=g)
    if opt.caption_model not in ("att2in2", "att2all2"):
        sd["fc_embed.0.weight"], sd["fc_embed.0.bias"] = lin(H, F_)
    if getattr(opt, "use_bn", 0):   # BatchNorm1d(att_feat_size) first (models/AttModel.py:79-80): non-trivial affine + running stats
        sd["att_embed.0.weight"] = 0.5 + torch.rand(D, generator=g)
        sd["att_embed.0.bias"] = 0.2 * torch.randn(D, generator=g)
        sd["att_embed.0.running_mean"] = 0.4 + 0.1 * torch.randn(D, generator=g)
        sd["att_embed.0.running_var"] = 0.25 + 0.2 * torch.rand(D, generator=g)
        sd["att_embed.0.num_batches_tracked"] = torch.tensor(3, dtype=torch.int64)
        sd["att_embed.1.weight"], sd["att_embed.1.bias"] = lin(H, D)
        if opt.use_bn == 2:   # a second BatchNorm1d(rnn_size) behind Linear + ReLU + Dropout (:84)
            sd["att_embed.4.weight"] = 0.5 + torch.rand(H, generator=g)
            sd["att_embed.4.bias"] = 0.1 * torch.randn(H, generator=g)
            sd["att_embed.4.running_mean"] = 0.2 + 0.05 * torch.randn(H, generator=g)
            sd["att_embed.4.running_var"] = 0.1 + 0.1 * torch.rand(H, generator=g)
            sd["att_embed.4.num_batches_tracked"] = torch.tensor(3, dtype=torch.int64)
    else:
        sd["att_embed.0.weight"], sd["att_embed.0.bias"] = lin(H, D)
    n_logit = getattr(opt, "logit_layers", 1)
    if n_logit == 1:
        sd["logit.weight"], sd["logit.bias"] = lin(V, H)
    else:   # [Linear(H, H), ReLU, Dropout(0.5)] x (n - 1) + Linear(H, V) in one nn.Sequential (:89-91)
        for i in range(n_logit - 1):
            sd[f"logit.{3 * i}.weight"], sd[f"logit.{3 * i}.bias"] = lin(H, H)
        sd[f"logit.{3 * (n_logit - 1)}.weight"], sd[f"logit.{3 * (n_logit - 1)}.bias"] = lin(V, H)
    last = "logit" if n_logit == 1 else f"logit.{3 * (n_logit - 1)}"
    sd["ctx2att.weight"], sd["ctx2att.bias"] = lin(A, H)
    if opt.caption_model in ("att2in2", "att2all2"):
        if opt.caption_model == "att2all2":
            sd["core.a2h.weight"], sd["core.a2h.bias"] = lin(5 * H, H)
        else:
            sd["core.a2c.weight"], sd["core.a2c.bias"] = lin(2 * H, H)
        sd["core.i2h.weight"], sd["core.i2h.bias"] = lin(5 * H, E)
        sd["core.h2h.weight"], sd["core.h2h.bias"] = lin(5 * H, H)
    elif opt.caption_model == "topdown":
        for name, in_f in (("att_lstm", E + 2 * H), ("lang_lstm", 2 * H)):
            w_ih, b_ih = lin(4 * H, in_f, bound_in=H)
            w_hh, b_hh = lin(4 * H, H, bound_in=H)
            sd[f"core.{name}.weight_ih"], sd[f"core.{name}.weight_hh"] = w_ih, w_hh
            sd[f"core.{name}.bias_ih"], sd[f"core.{name}.bias_hh"] = b_ih, b_hh
    elif opt.caption_model in ("stackatt", "denseatt"):   # models/AttModel.py:458-526: att1, att2, lstm0..2 (FCModel.LSTMCore), emb2, fusions
        for att_name in ("att1", "att2"):
            sd[f"core.{att_name}.h2att.weight"], sd[f"core.{att_name}.h2att.bias"] = lin(A, H)
            sd[f"core.{att_name}.alpha_net.weight"], sd[f"core.{att_name}.alpha_net.bias"] = lin(1, A)
        for name, in_f in (("lstm0", E + H), ("lstm1", 2 * H), ("lstm2", 2 * H)):
            sd[f"core.{name}.i2h.weight"], sd[f"core.{name}.i2h.bias"] = lin(5 * H, in_f)
            sd[f"core.{name}.h2h.weight"], sd[f"core.{name}.h2h.bias"] = lin(5 * H, H)
        sd["core.emb2.weight"], sd["core.emb2.bias"] = lin(H, H)
        if opt.caption_model == "denseatt":
            sd["core.fusion1.0.weight"], sd["core.fusion1.0.bias"] = lin(H, 2 * H)
            sd["core.fusion2.0.weight"], sd["core.fusion2.0.bias"] = lin(H, 3 * H)
    else:
        raise ValueError(f"caption_model {opt.caption_model!r} is outside the hot path")
    if opt.caption_model not in ("stackatt", "denseatt"):
        sd["core.attention.h2att.weight"], sd["core.attention.h2att.bias"] = lin(A, H)
        sd["core.attention.alpha_net.weight"], sd["core.attention.alpha_net.bias"] = lin(1, A)
    if peaked:
        sd[last + ".weight"] = sd[last + ".weight"] * peaked
    if eos_bias:
        sd[last + ".bias"] = sd[last + ".bias"].clone()
        sd[last + ".bias"][0] += eos_bias
    return sd
