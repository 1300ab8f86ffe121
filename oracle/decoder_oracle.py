"""CPU oracle for the attention-LSTM caption decoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this module; it is used by
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs as the
checker and as the timed CPU port of the reference, never as the thing shipped.

What it is: a plain fp32 PyTorch-on-CPU restatement, written as pure functions over a
`state_dict`, of the reference algorithm in /root/reference/pivot_based_eccv2018:
    models/AttModel.py     (AttModel, Attention, Att2in2Core, TopDownCore)
    models/CaptionModel.py (beam_search)
    misc/criterion.py      (LanguageModelCriterion)
The arithmetic itself lives in a third-party dependency of the reference (PyTorch, un-pinned by the
reference; here torch 2.11): nn.Linear / nn.LSTMCell / nn.Embedding / log_softmax / softmax / bmm /
sort / max.  Each function cites the reference lines it follows.

Parity pin: the reference has NO tests or golden vectors for this path.  The
oracle is pinned instead against outputs of the unmodified reference run in the build container
(`tests/golden/make_golden.py` imports /root/reference and writes `tests/golden/*.npz`); the
CPU test-suite checks this file against those fixtures (`tests/test_oracle_golden.py`), and against
the live reference when /root/reference is present (`tests/test_oracle_vs_reference.py`).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

KINDS = ("att2in2", "att2all2", "topdown", "stackatt", "denseatt")


def num_layers(kind):
    # models/AttModel.py:62 (att2in2 -> opt.num_layers == 1), :689 (topdown forces 2), :696,703 (stackatt / denseatt: 3)
    return {"topdown": 2, "stackatt": 3, "denseatt": 3}.get(kind, 1)


def _linear(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def logit_layer(sd, x):
    """self.logit (models/AttModel.py:86-91): one Linear, or logit_layers - 1 blocks of Linear + ReLU + Dropout(0.5)
    (indices 0, 3, 6, ... of the nn.Sequential; dropout inactive in eval mode) in front of it."""
    if "logit.weight" in sd:
        return _linear(sd, "logit", x)
    i = 0
    while f"logit.{i + 3}.weight" in sd:
        x = torch.relu(_linear(sd, f"logit.{i}", x))
        i += 3
    return _linear(sd, f"logit.{i}", x)


# ------------------------------------------------------------------------------------------------
# feature prologue -- models/AttModel.py:99-117 (clip_att, _prepare_feature), :30-53 (pack_wrapper)
# ------------------------------------------------------------------------------------------------
DROP_XT, DROP_ATT, DROP_FC, DROP_OUT = 0, 1, 2, 3   # dropout sites (row ids: t*B+b | b*L+l | b | b*T+t), as in _lib.py


def dropout_mask(drop, site, row_ids, cols):
    """Training-mode nn.Dropout as the product path draws it (csrc/pointwise.cu dropout_kernel): keep = u >= p with u the
    counter-based uniform of (seed, 0x40000000 + site, row, col); returns the fp32 multiplier keep / (1 - p), shape
    (len(row_ids), cols).  drop = (p, seed)."""
    p, seed = float(drop[0]), int(drop[1]) & 0xFFFFFFFFFFFFFFFF
    with np.errstate(over="ignore"):
        step = 0x40000000 + site
        step_key = np.uint32((seed ^ (seed >> 32)) & 0xFFFFFFFF) ^ np.uint32((step * 0x9E3779B1) & 0xFFFFFFFF)
        r = np.asarray(row_ids, dtype=np.uint64)
        row_key = _rng_mix(((np.uint64(step_key) + r * np.uint64(0x85EBCA77)) & np.uint64(0xFFFFFFFF)).astype(np.uint32))
        c = ((np.arange(cols, dtype=np.uint64) * np.uint64(0xC2B2AE3D)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        h = _rng_mix(row_key[:, None] ^ c[None, :])
    u = ((h >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)
    return torch.from_numpy((u >= np.float32(p)).astype(np.float32) / np.float32(1.0 - p))


_BN_TRAIN = False


class bn_training:
    """`with bn_training():` -- the BatchNorm1d of att_embed (use_bn = 1) normalises with the statistics of the batch
    (module in train() mode) instead of its running statistics."""

    def __enter__(self):
        global _BN_TRAIN
        self._old, _BN_TRAIN = _BN_TRAIN, True

    def __exit__(self, *exc):
        global _BN_TRAIN
        _BN_TRAIN = self._old


def bn_batch_stats(att_feats, att_masks):
    """Mean / biased variance over the packed valid regions (pack_wrapper, models/AttModel.py:44-53): the first
    sum(mask) regions of every image, after clip_att."""
    keep = int(att_masks.long().sum(1).max())
    n_valid = att_masks[:, :keep].long().sum(1)
    valid = torch.arange(keep)[None, :] < n_valid[:, None]
    xv = att_feats[:, :keep][valid]
    return xv.mean(0), xv.var(0, unbiased=False), int(valid.sum())


def prepare_features(sd, kind, fc_feats, att_feats, att_masks=None, drop=None):
    if att_masks is not None:  # clip_att (:99-105)
        keep = int(att_masks.long().sum(1).max())
        att_feats = att_feats[:, :keep].contiguous()
        att_masks = att_masks[:, :keep].contiguous()
    if kind not in ("att2in2", "att2all2"):  # fc_embed = Linear+ReLU(+Dropout) (:76-78); identity for att2in2 / att2all2 (:674-675,682-683)
        fc = torch.relu(_linear(sd, "fc_embed.0", fc_feats))
        if drop is not None:
            fc = fc * dropout_mask(drop, DROP_FC, np.arange(fc.size(0)), fc.size(1))
    else:
        fc = fc_feats
    # att_embed = Linear+ReLU (:79-84, use_bn=0, dropout off).  With masks the reference packs the
    # valid regions, embeds them and pads the rest with ZEROS (pad_packed_sequence, :50-51).
    if "att_embed.1.weight" in sd:
        # use_bn = 1 (:79-80): BatchNorm1d(att_feat_size) in front of the Linear.  It only runs on the packed 2-D rows, i.e.
        # with att_masks (a 3-D unmasked batch makes nn.BatchNorm1d read the region axis as channels and raise).
        if att_masks is None:
            raise RuntimeError("running_mean should contain %d elements not %d" % (att_feats.size(1), att_feats.size(2)))
        if _BN_TRAIN:
            mean, var, _ = bn_batch_stats(att_feats, att_masks)
        else:
            mean, var = sd["att_embed.0.running_mean"], sd["att_embed.0.running_var"]
        xn = (att_feats - mean) / torch.sqrt(var + 1e-5) * sd["att_embed.0.weight"] + sd["att_embed.0.bias"]
        att = torch.relu(_linear(sd, "att_embed.1", xn))
    else:
        att = torch.relu(_linear(sd, "att_embed.0", att_feats))
    if "att_embed.4.weight" in sd:
        # use_bn = 2 (:84): a second BatchNorm1d(rnn_size) behind Linear + ReLU + Dropout, on the packed valid rows
        if _BN_TRAIN:
            mean2, var2, _ = bn_batch_stats(att, att_masks)
        else:
            mean2, var2 = sd["att_embed.4.running_mean"], sd["att_embed.4.running_var"]
        att = (att - mean2) / torch.sqrt(var2 + 1e-5) * sd["att_embed.4.weight"] + sd["att_embed.4.bias"]
    if att_masks is not None:
        n_valid = att_masks.long().sum(1)
        valid = (torch.arange(att.size(1))[None, :] < n_valid[:, None]).to(att.dtype)
        att = att * valid[:, :, None]
    if drop is not None:                                               # att_embed's nn.Dropout (:79-84)
        B_, L_, H_ = att.shape
        att = att * dropout_mask(drop, DROP_ATT, np.arange(B_ * L_), H_).view(B_, L_, H_)
    p_att = _linear(sd, "ctx2att", att)  # :115 (bias also lands on padded rows, like the reference)
    return fc, att, p_att, att_masks


# ------------------------------------------------------------------------------------------------
# additive attention -- models/AttModel.py:538-558
# ------------------------------------------------------------------------------------------------
def attention(sd, h, att, p_att, att_masks=None, return_weights=False, prefix="core.attention"):
    att_h = _linear(sd, prefix + ".h2att", h)                           # :543
    hidden = torch.tanh(p_att + att_h[:, None, :])                      # :544-546
    w = sd[prefix + ".alpha_net.weight"].view(-1)
    score = hidden @ w + sd[prefix + ".alpha_net.bias"]                 # :548-549
    weight = torch.softmax(score, dim=1)                                # :551
    if att_masks is not None:                                           # :552-554
        weight = weight * att_masks.to(weight.dtype)
        weight = weight / weight.sum(1, keepdim=True)
    ctx = torch.bmm(weight[:, None, :], att).squeeze(1)                 # :555-556
    return (ctx, weight) if return_weights else ctx


# ------------------------------------------------------------------------------------------------
# recurrent cores
# ------------------------------------------------------------------------------------------------
def _lstm_cell(sd, prefix, x, h, c):
    """torch.nn.LSTMCell, gate order i, f, g, o (used at models/AttModel.py:426-427,434,441)."""
    gates = (F.linear(x, sd[prefix + ".weight_ih"], sd[prefix + ".bias_ih"]) +
             F.linear(h, sd[prefix + ".weight_hh"], sd[prefix + ".bias_hh"]))
    i, f, g, o = gates.chunk(4, dim=1)
    c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h_new = torch.sigmoid(o) * torch.tanh(c_new)
    return h_new, c_new


def core_att2in2(sd, xt, fc, att, p_att, state, att_masks=None):
    """models/AttModel.py:581-601 -- maxout LSTM whose cell input is the attended context."""
    h_prev, c_prev = state[0][-1], state[1][-1]
    H = h_prev.size(1)
    ctx = attention(sd, h_prev, att, p_att, att_masks)                  # :582
    sums = _linear(sd, "core.i2h", xt) + _linear(sd, "core.h2h", h_prev)  # :584
    sig = torch.sigmoid(sums[:, :3 * H])                                # :585-589
    i_g, f_g, o_g = sig[:, :H], sig[:, H:2 * H], sig[:, 2 * H:]
    pre = sums[:, 3 * H:] + _linear(sd, "core.a2c", ctx)                # :591-592
    g = torch.maximum(pre[:, :H], pre[:, H:])                           # :593-595
    c = f_g * c_prev + i_g * g                                          # :596
    h = o_g * torch.tanh(c)                                             # :597
    return h, (h[None], c[None])                                        # :599-601 (dropout p=0)


def core_att2all2(sd, xt, fc, att, p_att, state, att_masks=None):
    """models/AttModel.py:636-654 -- maxout LSTM whose five gate sums all receive the attended context."""
    h_prev, c_prev = state[0][-1], state[1][-1]
    H = h_prev.size(1)
    ctx = attention(sd, h_prev, att, p_att, att_masks)                  # :637
    sums = _linear(sd, "core.i2h", xt) + _linear(sd, "core.h2h", h_prev) + _linear(sd, "core.a2h", ctx)   # :639
    sig = torch.sigmoid(sums[:, :3 * H])                                # :640-644
    i_g, f_g, o_g = sig[:, :H], sig[:, H:2 * H], sig[:, 2 * H:]
    g = torch.maximum(sums[:, 3 * H:4 * H], sums[:, 4 * H:])            # :646-647
    c = f_g * c_prev + i_g * g                                          # :648
    h = o_g * torch.tanh(c)                                             # :649
    return h, (h[None], c[None])                                        # :651-653 (dropout p=0)


def core_topdown(sd, xt, fc, att, p_att, state, att_masks=None):
    """models/AttModel.py:430-446 -- attention LSTM + language LSTM."""
    h_lang_prev = state[0][-1]
    x1 = torch.cat([h_lang_prev, fc, xt], 1)                            # :431-432
    h_att, c_att = _lstm_cell(sd, "core.att_lstm", x1, state[0][0], state[1][0])    # :434
    ctx = attention(sd, h_att, att, p_att, att_masks)                   # :436
    x2 = torch.cat([ctx, h_att], 1)                                     # :438
    h_lang, c_lang = _lstm_cell(sd, "core.lang_lstm", x2, state[0][1], state[1][1])  # :441
    return h_lang, (torch.stack([h_att, h_lang]), torch.stack([c_att, c_lang]))   # :443-446


def _lstm_core(sd, prefix, x, h_prev, c_prev):
    """models/FCModel.py:14-42 (LSTMCore): the 5H maxout cell without a context term (dropout p=0)."""
    H = h_prev.size(1)
    sums = _linear(sd, prefix + ".i2h", x) + _linear(sd, prefix + ".h2h", h_prev)       # FCModel.py:27
    sig = torch.sigmoid(sums[:, :3 * H])                                # :28-32
    g = torch.maximum(sums[:, 3 * H:4 * H], sums[:, 4 * H:])            # :34-36
    c = sig[:, H:2 * H] * c_prev + sig[:, :H] * g                       # :37
    h = sig[:, 2 * H:] * torch.tanh(c)                                  # :38
    return h, c


def _stack_dense(sd, xt, fc, att, p_att, state, att_masks, dense):
    """models/AttModel.py:476-486 (StackAttCore.forward) and :516-526 (DenseAttCore.forward): three maxout cells, two
    attentions; the dense variant fuses h_0, h_1 into the third cell's input and all three into the output."""
    h0, c0 = _lstm_core(sd, "core.lstm0", torch.cat([xt, fc], 1), state[0][0], state[1][0])                  # :478 / :518
    a1 = attention(sd, h0, att, p_att, att_masks, prefix="core.att1")                                         # :479 / :519
    h1, c1 = _lstm_core(sd, "core.lstm1", torch.cat([h0, a1], 1), state[0][1], state[1][1])                  # :480 / :520
    a2 = attention(sd, h1 + _linear(sd, "core.emb2", a1), att, p_att, att_masks, prefix="core.att2")          # :481 / :521
    x2 = torch.relu(_linear(sd, "core.fusion1.0", torch.cat([h0, h1], 1))) if dense else h1                  # :522
    h2, c2 = _lstm_core(sd, "core.lstm2", torch.cat([x2, a2], 1), state[0][2], state[1][2])                  # :482 / :522
    out = torch.relu(_linear(sd, "core.fusion2.0", torch.cat([h0, h1, h2], 1))) if dense else h2             # :484 / :524
    return out, (torch.stack([h0, h1, h2]), torch.stack([c0, c1, c2]))


def core_stackatt(sd, xt, fc, att, p_att, state, att_masks=None):
    return _stack_dense(sd, xt, fc, att, p_att, state, att_masks, dense=False)


def core_denseatt(sd, xt, fc, att, p_att, state, att_masks=None):
    return _stack_dense(sd, xt, fc, att, p_att, state, att_masks, dense=True)


CORES = {"att2in2": core_att2in2, "att2all2": core_att2all2, "topdown": core_topdown, "stackatt": core_stackatt,
         "denseatt": core_denseatt}


def init_hidden(sd, kind, rows):
    # models/AttModel.py:94-97
    H = sd["ctx2att.weight"].size(1)
    z = sd["ctx2att.weight"].new_zeros(num_layers(kind), rows, H)
    return (z, z.clone())


def logprobs_state(sd, kind, it, fc, att, p_att, att_masks, state, drop=None, t=0, T=0):
    """models/AttModel.py:158-165: embed -> core -> logit -> log_softmax.  drop = (p, seed): training-mode dropout on
    the embedding (:75) and on the core's output (:431,599) at teacher-forced step t of T."""
    xt = torch.relu(sd["embed.0.weight"][it])                           # :73-75,160
    B = xt.size(0)
    if drop is not None:
        xt = xt * dropout_mask(drop, DROP_XT, t * B + np.arange(B), xt.size(1))
    out, state = CORES[kind](sd, xt, fc, att, p_att, state, att_masks)  # :162
    if drop is not None:
        out = out * dropout_mask(drop, DROP_OUT, np.arange(B) * T + t, out.size(1))
    return torch.log_softmax(logit_layer(sd, out), dim=1), state         # :163


# ------------------------------------------------------------------------------------------------
# teacher-forced forward -- models/AttModel.py:119-156 (ss_prob == 0)
# ------------------------------------------------------------------------------------------------
def teacher_forced(sd, kind, fc_feats, att_feats, seq, att_masks=None, ss_prob=0.0, ss_seed=0, return_tokens=False,
                   inputs=None, drop=None):
    """models/AttModel.py:119-156.  ss_prob > 0: scheduled sampling as in training mode (:130-143) -- per row a uniform
    draw against ss_prob decides whether the input token of step i is a sample from exp(outputs[:, i-1]) (detached)
    instead of seq[:, i].  The draws use the counter-based noise shared with the product path (see gumbel_noise):
    Gumbel-max for the token (same distribution as the reference's torch.multinomial) and column 0x7fffffff of the
    same hash for the coin.  `inputs` (B, T): feed exactly these tokens instead (to compare losses and gradients for
    the draws another implementation made)."""
    B, T = fc_feats.size(0), seq.size(1) - 1
    V = sd["embed.0.weight"].size(0)
    state = init_hidden(sd, kind, B)
    fc, att, p_att, masks = prepare_features(sd, kind, fc_feats, att_feats, att_masks, drop)
    steps, used, margins = [], [], []
    for i in range(T):
        it = seq[:, i].clone()
        margin = torch.full((B,), float("inf"))
        if inputs is not None:
            it = inputs[:, i].clone()
        elif i >= 1 and ss_prob > 0.0:                                  # :130-143
            coin = uniform_noise(ss_seed, i, B)
            keys = steps[-1].detach() + gumbel_noise(ss_seed, i, B, V)
            top2 = keys.topk(2, dim=1)
            sample_mask = coin < ss_prob
            it[sample_mask] = top2.indices[:, 0][sample_mask]
            margin[sample_mask] = (top2.values[:, 0] - top2.values[:, 1])[sample_mask]
        if i >= 1 and int(seq[:, i].sum()) == 0:                        # :148-151
            break
        lp, state = logprobs_state(sd, kind, it, fc, att, p_att, masks, state, drop, i, T)
        steps.append(lp)
        used.append(it)
        margins.append(margin)
    out = torch.stack(steps, 1)
    if out.size(1) < T:                                                 # untouched steps stay 0 (:123)
        out = torch.cat([out, out.new_zeros(B, T - out.size(1), V)], 1)
    if return_tokens:
        return out, torch.stack(used, 1), torch.stack(margins, 1)
    return out


def xe_loss(logprobs, target, mask):
    """misc/criterion.py:143-150."""
    target = target[:, :logprobs.size(1)]
    mask = mask[:, :logprobs.size(1)]
    picked = logprobs.gather(2, target.unsqueeze(2)).squeeze(2)
    return -(picked * mask).sum() / mask.sum()


def train_loss(sd, kind, fc_feats, att_feats, labels, masks, att_masks=None, ss_prob=0.0, ss_seed=0, inputs=None, drop=None):
    """The call pattern of trainer.py:164-165."""
    out = teacher_forced(sd, kind, fc_feats, att_feats, labels, att_masks, ss_prob, ss_seed, inputs=inputs, drop=drop)
    return xe_loss(out, labels[:, 1:], masks[:, 1:])


def loss_and_grads(sd, kind, fc_feats, att_feats, labels, masks, att_masks=None, ss_prob=0.0, ss_seed=0, inputs=None, drop=None):
    leaf = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    loss = train_loss(leaf, kind, fc_feats, att_feats, labels, masks, att_masks, ss_prob, ss_seed, inputs, drop)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return loss.detach(), grads


# ------------------------------------------------------------------------------------------------
# greedy sampling -- models/AttModel.py:198-253 with sample_max=1
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def sample_greedy(sd, kind, fc_feats, att_feats, seq_length, att_masks=None,
                  decoding_constraint=0, return_margins=False, relative_margins=False):
    """return_margins: also the oracle's top-2 log-prob margin per (row, step) (test harness, for the north-star exemption
    of near-ties); relative_margins divides it by logprob_scale of the row."""
    B = fc_feats.size(0)
    state = init_hidden(sd, kind, B)
    fc, att, p_att, masks = prepare_features(sd, kind, fc_feats, att_feats, att_masks)
    seq = torch.zeros(B, seq_length, dtype=torch.int64)
    seq_lp = torch.zeros(B, seq_length)
    margins = torch.full((B, seq_length), float("inf"))
    it = torch.zeros(B, dtype=torch.int64)
    unfinished = None
    for t in range(seq_length + 1):
        lp, state = logprobs_state(sd, kind, it, fc, att, p_att, masks, state)
        if decoding_constraint and t > 0:                               # :220-223
            lp = lp.clone()
            lp.scatter_(1, seq[:, t - 1:t], float("-inf"))
        if t == seq_length:                                             # :226-227
            break
        best, it = lp.max(1)                                            # :229
        top2 = lp.topk(2, dim=1).values
        margins[:, t] = (top2[:, 0] - top2[:, 1]) / (logprob_scale(lp) if relative_margins else 1.0)
        unfinished = (it > 0) if t == 0 else unfinished & (it > 0)      # :242-245
        it = it * unfinished.to(it.dtype)                               # :246
        seq[:, t] = it
        seq_lp[:, t] = best                                             # :248 (not masked)
        if int(unfinished.sum()) == 0:                                  # :250
            break
    return (seq, seq_lp, margins) if return_margins else (seq, seq_lp)


# ------------------------------------------------------------------------------------------------
# multinomial sampling -- models/AttModel.py:198-253 with sample_max=0 (self-critical roll-outs)
# The reference calls torch.multinomial(exp(logprobs / T), 1): token v with probability softmax(x / T)[v].
# Its random stream cannot be shared with a device kernel, so the product path (csrc/uic_vocab.cuh) and this
# oracle both draw by Gumbel-max -- argmax_v (logprob_v / T + g_v), the same distribution -- with g a pure
# function of (seed, step, row, column) restated here in numpy uint32 arithmetic.
# ------------------------------------------------------------------------------------------------
def _rng_mix(h):
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x7feb352d)).astype(np.uint32)
    h ^= h >> np.uint32(15)
    h = (h * np.uint32(0x846ca68b)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


def uniform_noise(seed, step, rows):
    """(rows,) fp32 uniforms in (0, 1): the per-row draw of decoding step `step` (rng_uniform at RNG_ROW_DRAW_COL)."""
    with np.errstate(over="ignore"):
        seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        step_key = np.uint32((seed ^ (seed >> 32)) & 0xFFFFFFFF) ^ np.uint32((step * 0x9E3779B1) & 0xFFFFFFFF)
        r = np.arange(rows, dtype=np.uint64)
        row_key = _rng_mix(((np.uint64(step_key) + r * np.uint64(0x85EBCA77)) & np.uint64(0xFFFFFFFF)).astype(np.uint32))
        c = np.uint32((0x7FFFFFFF * 0xC2B2AE3D) & 0xFFFFFFFF)
        h = _rng_mix(row_key ^ c)
    u = ((h >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)
    return torch.from_numpy(u).float()


def gumbel_noise(seed, step, rows, cols):
    """(rows, cols) fp32 standard Gumbel noise of decoding step `step` (csrc/uic_vocab.cuh: rng_step_key, rng_row_key,
    rng_gumbel)."""
    with np.errstate(over="ignore"):
        seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        step_key = np.uint32((seed ^ (seed >> 32)) & 0xFFFFFFFF) ^ np.uint32((step * 0x9E3779B1) & 0xFFFFFFFF)
        r = np.arange(rows, dtype=np.uint64)
        row_key = _rng_mix(((np.uint64(step_key) + r * np.uint64(0x85EBCA77)) & np.uint64(0xFFFFFFFF)).astype(np.uint32))
        c = ((np.arange(cols, dtype=np.uint64) * np.uint64(0xC2B2AE3D)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        h = _rng_mix(row_key[:, None] ^ c[None, :])
    u = ((h >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)
    return torch.from_numpy(-np.log(-np.log(u))).float()


@torch.no_grad()
def sample_multinomial(sd, kind, fc_feats, att_feats, seq_length, temperature=1.0, seed=0, att_masks=None,
                       decoding_constraint=0, return_margins=False, drop=None):
    """drop = (p, seed): the roll-out of a model in train() mode (self-critical training samples with dropout active)."""
    B = fc_feats.size(0)
    state = init_hidden(sd, kind, B)
    fc, att, p_att, masks = prepare_features(sd, kind, fc_feats, att_feats, att_masks, drop)
    seq = torch.zeros(B, seq_length, dtype=torch.int64)
    seq_lp = torch.zeros(B, seq_length)
    margins = torch.full((B, seq_length), float("inf"))
    it = torch.zeros(B, dtype=torch.int64)
    unfinished = None
    for t in range(seq_length + 1):
        lp, state = logprobs_state(sd, kind, it, fc, att, p_att, masks, state, drop, t, seq_length + 1)
        if decoding_constraint and t > 0:                               # :220-223
            lp = lp.clone()
            lp.scatter_(1, seq[:, t - 1:t], float("-inf"))
        if t == seq_length:                                             # :226-227
            break
        keys = lp / temperature + gumbel_noise(seed, t, B, lp.size(1))  # :232-237 (same distribution as multinomial)
        it = keys.argmax(1)
        top2 = keys.topk(2, dim=1).values
        margins[:, t] = top2[:, 0] - top2[:, 1]
        best = lp.gather(1, it[:, None]).squeeze(1)                     # :238
        unfinished = (it > 0) if t == 0 else unfinished & (it > 0)      # :242-245
        it = it * unfinished.to(it.dtype)                               # :246
        seq[:, t] = it
        seq_lp[:, t] = best                                             # :248 (not masked)
        if int(unfinished.sum()) == 0:                                  # :250
            break
    return (seq, seq_lp, margins) if return_margins else (seq, seq_lp)


def reward_loss(sample_logprobs, seq, reward):
    """misc/criterion.py:104-124 (RewardCriterion)."""
    mask = (seq > 0).float()
    mask = torch.cat([torch.ones(mask.size(0), 1), mask[:, :-1]], 1)
    return torch.sum(-sample_logprobs * reward * mask) / torch.sum(mask)


def rl_loss_and_grads(sd, kind, fc_feats, att_feats, seq, reward, att_masks=None, drop=None):
    """Self-critical step of trainer.py:166-173 for given sampled tokens: the roll-out's log-probs are those of
    teacher forcing on [BOS, seq] (AttModel.py:205-248 feeds `it * unfinished`, i.e. 0 after the end token)."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    B, T = seq.shape
    labels = torch.cat([torch.zeros(B, 1, dtype=torch.int64), seq, torch.zeros(B, 1, dtype=torch.int64)], 1)
    fc, att, p_att, masks = prepare_features(leaf, kind, fc_feats, att_feats, att_masks, drop)
    state = init_hidden(leaf, kind, B)
    lps = []
    for t in range(T):
        lp, state = logprobs_state(leaf, kind, labels[:, t], fc, att, p_att, masks, state, drop, t, T + 1)
        lps.append(lp.gather(1, labels[:, t + 1:t + 2]).squeeze(1))
    sample_lp = torch.stack(lps, 1)
    loss = reward_loss(sample_lp, seq, reward)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return loss.detach(), grads, sample_lp.detach()


# ------------------------------------------------------------------------------------------------
# beam search -- models/AttModel.py:167-196 + models/CaptionModel.py:33-177 (group_size == 1)
# One image at a time with Python-side candidate bookkeeping, like the reference: this is also the
# cost structure the CPU baseline is meant to show.
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def logprob_scale(lp):
    """Per-row magnitude that a relative tolerance on a row of log-probs refers to: |log-prob| of the best entry, or half
    the spread of the row (~ the largest |logit|) when that is bigger -- a peaked row has a best log-prob near 0 while its
    rounding error is that of logits of magnitude spread / 2."""
    fin = torch.where(torch.isfinite(lp), lp, lp.new_full((), float("nan")))
    hi = torch.nan_to_num(fin, nan=-float("inf")).max(1).values
    lo = torch.nan_to_num(fin, nan=float("inf")).min(1).values
    return torch.maximum(hi.abs(), 0.5 * (hi - lo))


def _rel_gaps(vals):
    """Smallest gap between neighbours of a descending score list, relative to the scores' magnitude.
    The -1000 offsets (finished beams, CaptionModel.py:167; UNK, :133) are bookkeeping, not log-probability mass: the
    rounding error of a cumulative score is proportional to the sum of its per-token log-probs, i.e. to the score with
    those offsets removed."""
    v = vals.double()
    v = v[torch.isfinite(v)]
    if v.numel() < 2:
        return float("inf")
    eff = (v + 1000.0 * torch.round(-v / 1000.0)).abs()
    scale = torch.maximum(eff[:-1], eff[1:]).clamp_min(1e-6)
    return float(((v[:-1] - v[1:]) / scale).min())


def _beam_search_one(sd, kind, state, logprobs, fc, att, p_att, masks, seq_length, beam_size,
                     decoding_constraint, max_ppl, margin_out=None):
    """margin_out (test harness only, not part of the reference): a list that receives this image's decision margins --
    per step the smallest relative gap among the b + 1 best entries of the full (rows x V) score matrix (a perturbed
    evaluation whose scores move by less than half of it selects the same beams in the same order), and last the gap
    between the two best finished hypotheses.  "Relative" = divided by the magnitude of the cumulative scores compared."""
    b, T = beam_size, seq_length
    beam_seq = torch.zeros(T, b, dtype=torch.int64)                      # CaptionModel.py:109-111
    beam_lp = torch.zeros(T, b)
    beam_sum = torch.zeros(b)
    done = []
    for t in range(T):
        lpf = logprobs.float().clone()
        if decoding_constraint and t > 0:                               # :130-131
            lpf.scatter_(1, beam_seq[t - 1].unsqueeze(1), float("-inf"))
        lpf[:, -1] -= 1000.0                                            # UNK suppression :133
        ys, ix = torch.sort(lpf, 1, True)                               # :61
        cands = []
        rows = 1 if t == 0 else b                                       # :64-66
        if margin_out is not None:
            full = (beam_sum[:rows, None] + lpf[:rows]).flatten()
            top = full.topk(min(b + 1, full.numel())).values
            if float(top[0]) > -900.0:                                  # (every beam already finished: nothing left to decide)
                margin_out.append(_rel_gaps(top))
        for c in range(min(b, ys.size(1))):                             # c-major, q-minor :67-73
            for q in range(rows):
                cands.append((beam_sum[q] + ys[q, c].item(), int(ix[q, c]), q, lpf[q, ix[q, c]]))
        cands.sort(key=lambda e: -e[0])                                 # stable :74
        new_state = [s.clone() for s in state]
        if t >= 1:
            prev_seq, prev_lp = beam_seq[:t].clone(), beam_lp[:t].clone()
        for vix in range(b):                                            # :82-95
            p, tok, q, r = cands[vix]
            if t >= 1:
                beam_seq[:t, vix] = prev_seq[:, q]
                beam_lp[:t, vix] = prev_lp[:, q]
            for s_new, s_old in zip(new_state, state):
                s_new[:, vix] = s_old[:, q]
            beam_seq[t, vix] = tok
            beam_lp[t, vix] = r
            beam_sum[vix] = p
        state = new_state
        for vix in range(b):                                            # :155-167
            if int(beam_seq[t, vix]) == 0 or t == T - 1:
                p = beam_sum[vix].item()
                done.append({"seq": beam_seq[:, vix].clone(), "logps": beam_lp[:, vix].clone(),
                             "unaug_p": beam_lp[:, vix].sum().item(),
                             "p": p / (t + 1) if max_ppl else p})
                beam_sum[vix] = -1000
        logprobs, state = logprobs_state(sd, kind, beam_seq[t], fc, att, p_att, masks, state)  # :171-172
    done.sort(key=lambda d: -d["p"])                                    # :175
    if margin_out is not None and len(done) > 1:
        margin_out.append(_rel_gaps(torch.tensor([done[0]["p"], done[1]["p"]])))
    return done[:b]


@torch.no_grad()
def _diverse_beam_search_one(sd, kind, state, logprobs, fc, att, p_att, masks, seq_length, beam_size, group_size,
                             diversity_lambda, decoding_constraint, max_ppl, margin_out=None):
    """models/CaptionModel.py:33-177 with group_size > 1 (diverse beam search): `group_size` groups of
    bdash = beam_size // group_size beams advance staggered by one step each; group divm ranks its candidates with
    diversity_lambda subtracted once per occurrence of a token among the tokens the groups before it hold at the same
    local step (:36-45), and stores the unpenalised log-prob."""
    G, T = group_size, seq_length
    bdash = beam_size // G
    seq_t = [torch.zeros(T, bdash, dtype=torch.int64) for _ in range(G)]          # :109-111
    lp_t = [torch.zeros(T, bdash) for _ in range(G)]
    sum_t = [torch.zeros(bdash) for _ in range(G)]
    done_t = [[] for _ in range(G)]
    state_t = [[s[:, g * bdash:(g + 1) * bdash].clone() for s in state] for g in range(G)]   # :115
    logprobs_t = list(logprobs.chunk(G, 0))                                       # :116
    feats = [[x[g * bdash:(g + 1) * bdash] if x is not None else None for x in (fc, att, p_att, masks)] for g in range(G)]
    for t in range(T + G - 1):                                                    # :124
        for divm in range(G):
            lt = t - divm
            if not (0 <= lt <= T - 1):                                            # :126
                continue
            lpf = logprobs_t[divm].float().clone()
            if decoding_constraint and lt > 0:                                    # :130-131
                lpf.scatter_(1, seq_t[divm][lt - 1].unsqueeze(1), float("-inf"))
            lpf[:, -1] -= 1000.0                                                  # :133
            unaug = lpf.clone()                                                   # :38
            for prev in range(divm):                                              # :39-44
                for tok in seq_t[prev][lt].tolist():
                    lpf[:, tok] -= diversity_lambda
            ys, ix = torch.sort(lpf, 1, True)                                     # :61
            rows = 1 if lt == 0 else bdash
            if margin_out is not None:                                            # (test harness, see _beam_search_one)
                full = (sum_t[divm][:rows, None] + lpf[:rows]).flatten()
                top = full.topk(min(bdash + 1, full.numel())).values
                if float(top[0]) > -900.0:
                    margin_out.append(_rel_gaps(top))
            cands = []
            for c in range(min(bdash, ys.size(1))):
                for q in range(rows):
                    cands.append((sum_t[divm][q] + ys[q, c].item(), int(ix[q, c]), q, unaug[q, ix[q, c]]))
            cands.sort(key=lambda e: -e[0])
            st = state_t[divm]
            new_state = [x.clone() for x in st]
            if lt >= 1:
                prev_seq, prev_lp = seq_t[divm][:lt].clone(), lp_t[divm][:lt].clone()
            for vix in range(bdash):
                p, tok, q, r = cands[vix]
                if lt >= 1:
                    seq_t[divm][:lt, vix] = prev_seq[:, q]
                    lp_t[divm][:lt, vix] = prev_lp[:, q]
                for s_new, s_old in zip(new_state, st):
                    s_new[:, vix] = s_old[:, q]
                seq_t[divm][lt, vix] = tok
                lp_t[divm][lt, vix] = r
                sum_t[divm][vix] = p
            state_t[divm] = new_state
            for vix in range(bdash):                                              # :155-167
                if int(seq_t[divm][lt, vix]) == 0 or lt == T - 1:
                    p = sum_t[divm][vix].item()
                    done_t[divm].append({"seq": seq_t[divm][:, vix].clone(), "logps": lp_t[divm][:, vix].clone(),
                                         "unaug_p": lp_t[divm][:, vix].sum().item(),
                                         "p": p / (lt + 1) if max_ppl else p})
                    sum_t[divm][vix] = -1000
            f = feats[divm]
            logprobs_t[divm], state_t[divm] = logprobs_state(sd, kind, seq_t[divm][lt], f[0], f[1], f[2], f[3], state_t[divm])  # :171-172
    out = []
    for g in range(G):                                                            # :175-176
        ranked = sorted(done_t[g], key=lambda d: -d["p"])
        if margin_out is not None and g == 0 and len(ranked) > 1:                 # the returned caption is group 0's best
            margin_out.append(_rel_gaps(torch.tensor([ranked[0]["p"], ranked[1]["p"]])))
        out += ranked[:bdash]
    return out


@torch.no_grad()
def sample_beam(sd, kind, fc_feats, att_feats, seq_length, beam_size=10, att_masks=None,
                decoding_constraint=0, max_ppl=0, group_size=1, diversity_lambda=0.5, return_margins=False):
    """Returns (seq (B,T) int64, seqLogprobs (B,T) fp32, done_beams list-of-lists) and, with return_margins,
    a (B,) tensor: the smallest relative decision margin along each image's search (see _beam_search_one)."""
    B = fc_feats.size(0)
    margins = torch.full((B,), float("inf"), dtype=torch.float64)
    fc, att, p_att, masks = prepare_features(sd, kind, fc_feats, att_feats, att_masks)
    V = sd["embed.0.weight"].size(0)
    assert beam_size <= V                                               # AttModel.py:173
    seq = torch.zeros(seq_length, B, dtype=torch.int64)
    seq_lp = torch.zeros(seq_length, B)
    done_beams = []
    for k in range(B):                                                  # AttModel.py:179
        state = init_hidden(sd, kind, beam_size)
        fc_k = fc[k:k + 1].expand(beam_size, fc.size(1))
        att_k = att[k:k + 1].expand(beam_size, *att.shape[1:]).contiguous()
        p_att_k = p_att[k:k + 1].expand(beam_size, *p_att.shape[1:]).contiguous()
        m_k = masks[k:k + 1].expand(beam_size, masks.size(1)).contiguous() if masks is not None else None
        it = torch.zeros(beam_size, dtype=torch.int64)
        lp, state = logprobs_state(sd, kind, it, fc_k, att_k, p_att_k, m_k, state)   # :186-190
        m_k_list = [] if return_margins else None
        if group_size > 1:
            done = _diverse_beam_search_one(sd, kind, state, lp, fc_k, att_k, p_att_k, m_k, seq_length, beam_size, group_size,
                                            diversity_lambda, decoding_constraint, max_ppl, m_k_list)
        else:
            done = _beam_search_one(sd, kind, state, lp, fc_k, att_k, p_att_k, m_k, seq_length,
                                    beam_size, decoding_constraint, max_ppl, m_k_list)
        if m_k_list:
            margins[k] = min(m_k_list)
        done_beams.append(done)
        seq[:, k] = done[0]["seq"]                                      # :193-194
        seq_lp[:, k] = done[0]["logps"]
    if return_margins:
        return seq.t(), seq_lp.t(), done_beams, margins
    return seq.t(), seq_lp.t(), done_beams


def sample(sd, kind, fc_feats, att_feats, seq_length, att_masks=None, opt=None):
    """Dispatch of AttModel._sample (models/AttModel.py:198-205)."""
    opt = opt or {}
    if opt.get("beam_size", 1) > 1:
        s, lp, _ = sample_beam(sd, kind, fc_feats, att_feats, seq_length, opt["beam_size"], att_masks,
                               opt.get("decoding_constraint", 0), opt.get("max_ppl", 0))
        return s, lp
    return sample_greedy(sd, kind, fc_feats, att_feats, seq_length, att_masks,
                         opt.get("decoding_constraint", 0))
