"""All-reduce time of the gradient buckets' sizes on this box (torchrun, NCCL): what the exchange step of the data-parallel
training path costs when it is not hidden.  Prints ms and algorithm bandwidth per size on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
for mb in (5, 20, 33, 81):
    x = torch.ones(mb * 2 ** 20 // 4, device="cuda")
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 20
    for _ in range(n):
        dist.all_reduce(x)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    if dist.get_rank() == 0:
        print(f"world {world}: all_reduce {mb} MB fp32: {ms * 1e3:.1f} us, algbw {mb * 2 ** 20 / ms / 1e6:.0f} GB/s, busbw {mb * 2 ** 20 / ms / 1e6 * 2 * (world - 1) / world:.0f} GB/s")
dist.destroy_process_group()
