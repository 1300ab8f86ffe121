"""Data-parallel training plumbing: one process per GPU, NCCL over NVLink for the single exchange
step of the path (gradient all-reduce), replacing the reference's single-process
torch.nn.DataParallel (trainer.py:74,88-89).

The loss normaliser of LanguageModelCriterion is sum(mask) over the WHOLE batch
(misc/criterion.py:149; DataParallel gathers the outputs on GPU 0 before the criterion runs).  Every rank
therefore computes the gradients of its UN-normalised masked NLL sum and appends its local sum(mask) to the
gradient buffer: the all-reduce that sums the gradients also sums the normaliser, and one scale by its
reciprocal afterwards yields exactly the gradients of the reference's loss.  No collective gates the forward.

Overlap: BPTT accumulates into every recurrent weight until its last step, but the gradient groups become
final one after the other -- logit.* right after the fused logit stage (before BPTT starts), the recurrent
core after the time-batched wgrads, the embedding table after its scatter, the feature-side layers last.
`GradBuckets` lays the parameters out in that order in ONE flat fp32 buffer (p.grad are views into it) and
`DataParallelStep` starts each bucket's all-reduce the moment its last gradient has been written, on NCCL's
own stream, while the kernels of the next group keep running; only the last, small bucket is exposed.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def global_mask_sum(masks_shifted):
    """sum over all ranks of masks[:, 1:] (a 0-dim tensor on the masks' device): the loss normaliser for callers that
    run `mode='forward_loss'` + loss.backward() themselves (DataParallelStep does not need it)."""
    s = masks_shifted.float().sum()
    if world_size() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return s


def bucket_of(name):
    """Exchange order of a parameter (see the module docstring)."""
    if name.startswith("logit."):
        return 0
    if name.startswith("core.") or name.startswith("fc_embed."):
        return 1
    if name.startswith("embed."):
        return 2
    return 3


class GradBuckets:
    """All parameter gradients in ONE flat fp32 buffer, grouped by the order in which they become final; the last
    element is the loss normaliser (local sum(mask)), reduced together with the last bucket."""

    N_BUCKETS = 4

    def __init__(self, model):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        named.sort(key=lambda np_: bucket_of(np_[0]))          # stable: registration order inside a bucket
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n + 1, dtype=torch.float32, device=dev)
        self.grads = self.flat[:n]
        self.tail = self.flat[n:]                               # [local sum(mask)]
        self.by_name, self.bounds, off = {}, [], 0
        for b in range(self.N_BUCKETS):
            start = off
            for name, p in named:
                if bucket_of(name) == b:
                    p.grad = self.flat[off:off + p.numel()].view_as(p)
                    self.by_name[name] = p
                    off += p.numel()
            self.bounds.append((start, off))
        self.bounds[-1] = (self.bounds[-1][0], n + 1)          # the normaliser travels with the last bucket
        self.members = [{nm for nm in self.names if bucket_of(nm) == b} for b in range(self.N_BUCKETS)]

    def view(self, b):
        lo, hi = self.bounds[b]
        return self.flat[lo:hi]

    def zero(self):
        self.flat.zero_()

    def clip_(self, max_norm):
        """clip_grad_norm_ on the flat buffer (misc/optimizer.py:89-93 clips at 5.0)."""
        norm = self.grads.norm()
        scale = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
        self.grads.mul_(scale)
        return norm


GradBucket = GradBuckets   # (name of the single-bucket version of round 1; same surface: flat, zero, clip_)


class DataParallelStep:
    """loss = step(fc, att, labels, masks[, att_masks]): gradients of the reference's loss for the GLOBAL batch in
    p.grad on every rank (bucketed, overlapped all-reduce), clipped to `clip`; returns this rank's share of the loss
    (local NLL sum / global mask sum -- the shares of all ranks add up to the reference's loss).

    grad_fn(model, fc, att, labels, masks, att_masks, on_ready) -> (nll_sum, mask_sum, {name: grad of nll_sum}) defaults to
    the CUDA path (autograd.xe_sum_and_grads); tests inject a CPU implementation to exercise the exchange logic on gloo."""

    def __init__(self, model, clip=5.0, grad_fn=None, overlap=True):
        self.model, self.clip, self.overlap = model, clip, overlap
        self.buckets = GradBuckets(model)
        if grad_fn is None:
            from .autograd import xe_sum_and_grads

            def grad_fn(model, fc, att, labels, masks, att_masks, on_ready):
                model._check_tokens(labels)
                return xe_sum_and_grads(model, fc, att, labels, masks, att_masks, model._scheduled_sampling(att.device),
                                        model._dropout(att.device), on_ready)
        self.grad_fn = grad_fn

    def __call__(self, fc, att, labels, masks, att_masks=None):
        bk, world = self.buckets, world_size()
        pending = [set(m) for m in bk.members]
        works = []

        def launch(b):
            if world > 1:
                works.append(dist.all_reduce(bk.view(b), op=dist.ReduceOp.SUM, async_op=True))

        def on_ready(chunk):
            for name, grad in chunk.items():
                p = bk.by_name.get(name)
                if p is None:
                    continue
                p.grad.copy_(grad.reshape(p.shape))
                b = bucket_of(name)
                pending[b].discard(name)
                if self.overlap and not pending[b] and b < bk.N_BUCKETS - 1:
                    launch(b)

        nll_sum, mask_sum, rest = self.grad_fn(self.model, fc, att, labels, masks, att_masks, on_ready)
        on_ready({n: g for n, g in rest.items() if any(n in m for m in pending)})
        missing = [n for m in pending for n in m]
        if missing:
            raise RuntimeError(f"DataParallelStep: no gradient for {missing}")
        bk.tail.copy_(mask_sum.reshape(1))
        if not self.overlap:
            for b in range(bk.N_BUCKETS - 1):
                launch(b)
        launch(bk.N_BUCKETS - 1)
        for w in works:
            w.wait()
        inv = 1.0 / bk.tail                                     # 1 / global sum(mask), identical on every rank
        if self.clip:
            # normalisation and clip_grad_norm_ (misc/optimizer.py:89-93) in ONE pass over the 81 MB buffer:
            # ||g inv|| = inv ||g||, so the clip factor of the normalised gradients is known before they are scaled
            norm = bk.grads.norm() * inv
            bk.grads.mul_(inv * torch.clamp(self.clip / (norm + 1e-6), max=1.0))
        else:
            bk.grads.mul_(inv)
        return nll_sum * inv[0]
