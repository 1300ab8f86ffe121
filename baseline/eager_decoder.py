"""The incumbent on this path: the reference's decoder architecture in stock torch.nn, run EAGERLY on the same B200
(cuBLAS / ATen kernels, one launch per op, one host sync per time step like the reference) -- BASELINE.md §3.4,
SURVEY.md §8d "second, stronger baseline".  Used by bench.py's `eager_b200_baseline` leg only: it is neither the product
(no kernel of libuic_b200.so runs here) nor the parity oracle (oracle/decoder_oracle.py, CPU).

Parameter names equal the reference's (models/AttModel.py:56-92,421-446,561-601), so the same state_dict loads.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Attention(nn.Module):          # models/AttModel.py:529-558
    def __init__(self, H, A):
        super().__init__()
        self.h2att = nn.Linear(H, A)
        self.alpha_net = nn.Linear(A, 1)

    def forward(self, h, att, p_att):
        dot = torch.tanh(p_att + self.h2att(h).unsqueeze(1))
        weight = F.softmax(self.alpha_net(dot).squeeze(2), dim=1)
        return torch.bmm(weight.unsqueeze(1), att).squeeze(1)


class _Att2in2Core(nn.Module):        # models/AttModel.py:561-601
    def __init__(self, E, H, A):
        super().__init__()
        self.a2c, self.i2h, self.h2h = nn.Linear(H, 2 * H), nn.Linear(E, 5 * H), nn.Linear(H, 5 * H)
        self.attention = _Attention(H, A)

    def forward(self, xt, fc, att, p_att, state):
        H = state[0].size(-1)
        ctx = self.attention(state[0][-1], att, p_att)
        sums = self.i2h(xt) + self.h2h(state[0][-1])
        sig = torch.sigmoid(sums[:, :3 * H])
        pre = sums[:, 3 * H:] + self.a2c(ctx)
        c = sig[:, H:2 * H] * state[1][-1] + sig[:, :H] * torch.max(pre[:, :H], pre[:, H:])
        h = sig[:, 2 * H:] * torch.tanh(c)
        return h, (h.unsqueeze(0), c.unsqueeze(0))


class _TopDownCore(nn.Module):        # models/AttModel.py:421-446
    def __init__(self, E, H, A):
        super().__init__()
        self.att_lstm, self.lang_lstm = nn.LSTMCell(E + 2 * H, H), nn.LSTMCell(2 * H, H)
        self.attention = _Attention(H, A)

    def forward(self, xt, fc, att, p_att, state):
        h_att, c_att = self.att_lstm(torch.cat([state[0][-1], fc, xt], 1), (state[0][0], state[1][0]))
        ctx = self.attention(h_att, att, p_att)
        h_lang, c_lang = self.lang_lstm(torch.cat([ctx, h_att], 1), (state[0][1], state[1][1]))
        return h_lang, (torch.stack([h_att, h_lang]), torch.stack([c_att, c_lang]))


class EagerDecoder(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.kind, self.seq_length = opt.caption_model, opt.seq_length
        V, E, H, A = opt.vocab_size + 1, opt.input_encoding_size, opt.rnn_size, opt.att_hid_size
        self.H, self.V = H, V
        self.embed = nn.Sequential(nn.Embedding(V, E), nn.ReLU())
        if self.kind == "topdown":
            self.fc_embed = nn.Sequential(nn.Linear(opt.fc_feat_size, H), nn.ReLU())
        self.att_embed = nn.Sequential(nn.Linear(opt.att_feat_size, H), nn.ReLU())
        self.logit = nn.Linear(H, V)
        self.ctx2att = nn.Linear(H, A)
        self.core = _TopDownCore(E, H, A) if self.kind == "topdown" else _Att2in2Core(E, H, A)
        self.num_layers = 2 if self.kind == "topdown" else 1

    def _prepare(self, fc, att):
        fc = self.fc_embed(fc) if self.kind == "topdown" else fc
        att = self.att_embed(att)
        return fc, att, self.ctx2att(att)

    def _state(self, B, ref):
        z = ref.new_zeros(self.num_layers, B, self.H)
        return (z, z.clone())

    def _step(self, it, fc, att, p_att, state):
        out, state = self.core(self.embed(it), fc, att, p_att, state)
        return F.log_softmax(self.logit(out), dim=1), state

    def forward(self, fc, att, seq):                       # teacher forcing, models/AttModel.py:119-156
        B, T = fc.size(0), seq.size(1) - 1
        fc, att, p_att = self._prepare(fc, att)
        state = self._state(B, att)
        outputs = att.new_zeros(B, T, self.V)
        for i in range(T):
            if i >= 1 and seq[:, i].sum() == 0:            # (host sync per step, like the reference :148-151)
                break
            lp, state = self._step(seq[:, i], fc, att, p_att, state)
            outputs[:, i] = lp
        return outputs

    @torch.no_grad()
    def sample_greedy(self, fc, att):                      # models/AttModel.py:198-253, sample_max = 1
        B = fc.size(0)
        fc, att, p_att = self._prepare(fc, att)
        state = self._state(B, att)
        seq = fc.new_zeros(B, self.seq_length, dtype=torch.long)
        lps = fc.new_zeros(B, self.seq_length)
        it = fc.new_zeros(B, dtype=torch.long)
        unfinished = None
        for t in range(self.seq_length):
            lp, state = self._step(it, fc, att, p_att, state)
            best, it = lp.max(1)
            unfinished = (it > 0) if t == 0 else unfinished & (it > 0)
            it = it * unfinished.to(it.dtype)
            seq[:, t], lps[:, t] = it, best
            if unfinished.sum() == 0:                      # (host sync per step :250)
                break
        return seq, lps


def xe_loss(logprobs, target, mask):                       # misc/criterion.py:143-150
    target, mask = target[:, :logprobs.size(1)], mask[:, :logprobs.size(1)]
    return torch.sum(-logprobs.gather(2, target.unsqueeze(2)).squeeze(2) * mask) / torch.sum(mask)
