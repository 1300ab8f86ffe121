"""Data-parallel training plumbing: one process per GPU, NCCL over NVLink for the single exchange
step of the path (gradient all-reduce), replacing the reference's single-process
torch.nn.DataParallel (trainer.py:74,88-89).

The loss normaliser of LanguageModelCriterion is sum(mask) over the WHOLE batch
(misc/criterion.py:149; DataParallel gathers the outputs on GPU 0 before the criterion runs), so
ranks first all-reduce the scalar sum(mask) and each computes  sum_local(nll) / sum_global(mask);
the all-reduced (summed) gradients are then exactly the gradients of the reference's loss.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def global_mask_sum(masks_shifted):
    """sum over all ranks of masks[:, 1:] (a 0-dim tensor on the masks' device)."""
    s = masks_shifted.float().sum()
    if world_size() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return s


class GradBucket:
    """All parameter gradients live in ONE flat fp32 buffer (p.grad are views into it), so the
    exchange step is a single NCCL all-reduce with no packing copies."""

    def __init__(self, model):
        self.params = [p for p in model.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce(self, async_op=False):
        if world_size() == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)

    def clip_(self, max_norm):
        """clip_grad_norm_ on the flat buffer (misc/optimizer.py:89-93 clips at 5.0)."""
        norm = self.flat.norm()
        scale = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
        self.flat.mul_(scale)
        return norm
