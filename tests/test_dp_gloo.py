"""Host-side logic of the data-parallel path on CPU with gloo, world_size 2: the global loss
normaliser and the flat gradient bucket reproduce the single-process (gathered-batch) result the
reference's nn.DataParallel computes (trainer.py:74,164-165).  The model here is the CPU oracle; the
collective plumbing is the product code in unpaired_image_captioning_b200/dp.py."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import decoder_oracle as O
    from unpaired_image_captioning_b200 import dp, synth
    torch.set_num_threads(1)
    opt, cfg = synth.opt_for("tiny_topdown")
    sd = synth.init_state_dict(opt, seed=5)
    B = 6
    fc, att = synth.make_features(B, 7, 64, seed=5)
    labels, masks = synth.make_captions(B, 6, 51, seed=5, min_len=2)
    lo, hi = rank * B // world, (rank + 1) * B // world          # contiguous shard of the batch
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(v.detach().clone()) for v in sd.values()])
    holder = Holder()
    bucket = dp.GradBucket(holder)
    bucket.zero()
    params = dict(zip(sd.keys(), holder.ps))
    norm = dp.global_mask_sum(masks[lo:hi, 1:])
    out = O.teacher_forced(params, "topdown", fc[lo:hi], att[lo:hi], labels[lo:hi])
    tgt, m = labels[lo:hi, 1:], masks[lo:hi, 1:]
    loss = -(out.gather(2, tgt.unsqueeze(2)).squeeze(2) * m).sum() / norm
    loss.backward()
    bucket.allreduce()
    total = loss.detach().clone()
    dist.all_reduce(total)
    if rank == 0:
        ref_loss, ref_grads = O.loss_and_grads(sd, "topdown", fc, att, labels, masks)
        ok = abs(float(total) - float(ref_loss)) < 1e-5
        ok = ok and abs(float(norm) - float(masks[:, 1:].sum())) < 1e-6
        for k, p in params.items():
            ok = ok and torch.allclose(p.grad, ref_grads[k], rtol=1e-4, atol=1e-6)
        # clipping acts on the global norm of the reduced gradient
        n = bucket.clip_(1e-3)
        ok = ok and abs(float(bucket.flat.norm()) - 1e-3) < 1e-6 and float(n) > 1e-3
        ret.put(bool(ok))
    dist.destroy_process_group()


def test_world_size_2_matches_gathered_batch():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
