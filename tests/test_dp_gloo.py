"""Host-side logic of the data-parallel path on CPU with gloo, world_size 2: dp.DataParallelStep (un-normalised local
sums, the normaliser riding in the last bucket, bucketed / overlapped all-reduce in the order the gradient groups become
final) reproduces the single-process gathered-batch result the reference's nn.DataParallel computes
(trainer.py:74,164-165).  The gradients come from the CPU oracle through the step's `grad_fn` hook, announced group by
group like autograd.xe_sum_and_grads does on the device; the exchange is the product code in dp.py."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Holder(torch.nn.Module):
    """A module whose parameter names are the reference's (so that dp.bucket_of sorts them into the exchange order)."""

    def __init__(self, sd):
        super().__init__()
        self._names = list(sd.keys())
        for i, (k, v) in enumerate(sd.items()):
            self.register_parameter("p%d" % i, torch.nn.Parameter(v.detach().clone()))

    def named_parameters(self, *a, **k):
        for i, name in enumerate(self._names):
            yield name, getattr(self, "p%d" % i)


def _oracle_grad_fn(kind, order_log):
    from oracle import decoder_oracle as O

    def grad_fn(model, fc, att, labels, masks, att_masks, on_ready):
        params = dict(model.named_parameters())
        leaf = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        out = O.teacher_forced(leaf, kind, fc, att, labels)
        tgt, m = labels[:, 1:], masks[:, 1:]
        nll_sum = -(out.gather(2, tgt.unsqueeze(2)).squeeze(2) * m).sum()
        nll_sum.backward()
        g = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
        # announce the groups in the order the device path does: logit, core + fc_embed, embed; the rest is returned
        for group in (lambda n: n.startswith("logit."), lambda n: n.startswith("core.") or n.startswith("fc_embed."),
                      lambda n: n.startswith("embed.")):
            chunk = {n: g.pop(n) for n in list(g) if group(n)}
            order_log.append(sorted(chunk))
            on_ready(chunk)
        return nll_sum.detach(), m.sum(), g
    return grad_fn


def _worker(rank, world, port, ret, overlap):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import decoder_oracle as O
    from unpaired_image_captioning_b200 import dp, shard_bounds, synth
    torch.set_num_threads(1)
    opt, cfg = synth.opt_for("tiny_topdown")
    sd = synth.init_state_dict(opt, seed=5)
    B = 7                                                         # uneven shards: 4 + 3 rows, different mask sums
    fc, att = synth.make_features(B, 7, 64, seed=5)
    labels, masks = synth.make_captions(B, 6, 51, seed=5, min_len=2)
    lo, hi = shard_bounds(B, rank, world)
    holder = _Holder(sd)
    log = []
    step = dp.DataParallelStep(holder, clip=0.0, grad_fn=_oracle_grad_fn("topdown", log), overlap=overlap)
    share = step(fc[lo:hi], att[lo:hi], labels[lo:hi], masks[lo:hi])
    total = share.detach().clone()
    dist.all_reduce(total)
    if rank == 0:
        ref_loss, ref_grads = O.loss_and_grads(sd, "topdown", fc, att, labels, masks)
        ok = abs(float(total) - float(ref_loss)) < 1e-5
        ok = ok and abs(float(step.buckets.tail) - float(masks[:, 1:].sum())) < 1e-6      # the normaliser was summed over ranks
        for k, p in holder.named_parameters():
            ok = ok and torch.allclose(p.grad, ref_grads[k], rtol=1e-4, atol=1e-6)
        # bucket layout: contiguous, in exchange order, the normaliser at the very end
        bk = step.buckets
        ok = ok and bk.bounds[0][0] == 0 and all(bk.bounds[i][1] == bk.bounds[i + 1][0] for i in range(3))
        ok = ok and bk.bounds[3][1] == bk.flat.numel() and [dp.bucket_of(n) for n in bk.names] == sorted(dp.bucket_of(n) for n in bk.names)
        ok = ok and log[0] == ["logit.bias", "logit.weight"]
        # clipping acts on the global norm of the reduced gradient
        n = bk.clip_(1e-3)
        ok = ok and abs(float(bk.grads.norm()) - 1e-3) < 1e-6 and float(n) > 1e-3
        ret.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_world_size_2_matches_gathered_batch(overlap):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() + int(overlap)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
