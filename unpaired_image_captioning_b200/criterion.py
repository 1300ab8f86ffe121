"""misc/criterion.py:138-159 -- LanguageModelCriterion (masked cross-entropy over log-probs).

Drop-in signature `LanguageModelCriterion(opt)(input, target, mask)`.  The gather/mask/sum runs on
whatever device `input` lives on with the library's row kernels when it is a CUDA tensor that does
not require grad; the differentiable path goes through `unpaired_image_captioning_b200.autograd`
(the fused loss never materialises (B, T, V)).
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
