"""misc/criterion.py:138-159 -- LanguageModelCriterion (masked cross-entropy over log-probs), and
misc/criterion.py:104-124 -- RewardCriterion (the self-critical policy-gradient loss).

Drop-in signature `LanguageModelCriterion(opt)(input, target, mask)`.  On the (B, T, V) log-prob tensor the unmodified call
pattern `crit(model(...), labels[:, 1:], masks[:, 1:])` produces, the three-line gather / mask / sum below is plain torch
(it is the reference's formula on a tensor that already exists; its backward is the hand-written BPTT of
`autograd._DecoderLogprobsFn`).  The path that never materialises (B, T, V) is `mode='forward_loss'`: the model returns a
handle whose `_uic_fused(target, mask)` runs the library's fused logit-stage / LSE / NLL kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class LanguageModelCriterion(nn.Module):
    def __init__(self, opt=None):
        super().__init__()
        self.caption_model = getattr(opt, "caption_model", "")

    def xe_loss(self, input, target, mask):
        # truncate to the same size (misc/criterion.py:144-146)
        target = target[:, :input.size(1)]
        mask = mask[:, :input.size(1)]
        if getattr(input, "_uic_fused", None) is not None:
            return input._uic_fused(target, mask)
        output = -input.gather(2, target.unsqueeze(2)).squeeze(2) * mask
        return torch.sum(output) / torch.sum(mask)

    def forward(self, input, target, mask):
        if "stackcap" in self.caption_model:
            return self.xe_loss(input[0], target, mask) + self.xe_loss(input[1], target, mask) + self.xe_loss(input[2], target, mask)
        return self.xe_loss(input, target, mask)


class RewardCriterion(nn.Module):
    """misc/criterion.py:104-124: `loss = -sum(logprob * reward * mask) / sum(mask)` with
    `mask[:, t] = 1` for the first step and wherever the previous sampled token is not 0 (so the step that emits
    the end token still counts).  `input` are the sample log-probs returned by
    `model(..., opt={'sample_max': 0}, mode='sample')`; when gradients are enabled they carry the decoder's
    autograd graph (teacher forcing on the sampled tokens), so `loss.backward()` works as in `trainer.py:166-173`."""

    def forward(self, input, seq, reward):
        input = input.contiguous().view(-1)
        reward = reward.contiguous().view(-1).to(input.dtype)
        mask = (seq > 0).to(input.dtype)
        mask = torch.cat([mask.new_ones(mask.size(0), 1), mask[:, :-1]], 1).contiguous().view(-1)
        output = -input * reward * mask
        return torch.sum(output) / torch.sum(mask)
