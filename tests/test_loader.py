"""Host logic of the feature path (unpaired_image_captioning_b200/loader.py) on CPU: shard bounds, the cache's dtype
handling, and the rank-ordered gather of sharded captions over gloo (world size 2)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything_once():
    from unpaired_image_captioning_b200 import shard_bounds
    for n in (0, 1, 7, 8, 40000, 40003):
        for world in (1, 2, 3, 8):
            got = [shard_bounds(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1


def test_feature_cache_rounds_once_and_counts_bytes():
    from unpaired_image_captioning_b200 import FeatureCache
    att = torch.randn(5, 7, 16)
    fc = att.mean(1)
    c16 = FeatureCache(fc, att, dtype=torch.bfloat16, pin=False)
    c32 = FeatureCache(fc, att, torch.ones(5, 7), dtype=torch.float32, pin=False)
    assert c16.att.dtype == torch.bfloat16 and torch.equal(c16.att, att.to(torch.bfloat16)) and c16.fc.dtype == torch.float32
    assert c16.nbytes(3) == 3 * (7 * 16 * 2 + 16 * 4)
    assert c32.nbytes(2) == 2 * (7 * 16 * 4 + 16 * 4 + 7 * 4)
    assert len(c16) == 5


def _gather_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unpaired_image_captioning_b200 import gather_captions, shard_bounds
    n, T = 11, 4
    lo, hi = shard_bounds(n, rank, world)
    seqs = torch.arange(lo, hi)[:, None].repeat(1, T)          # row i of the split holds i
    out = gather_captions(seqs, seqs.float() * 0.5, n, rank, world)
    if rank == 0:
        ret["ok"] = bool(torch.equal(out[0][:, 0], torch.arange(n)) and torch.equal(out[1][:, 0], torch.arange(n) * 0.5))
    else:
        ret["other"] = out is None
    dist.destroy_process_group()


def test_gather_captions_is_rank_ordered_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gather_worker, args=(2, 29533, ret), nprocs=2, join=True)
    assert ret["ok"] and ret["other"]
