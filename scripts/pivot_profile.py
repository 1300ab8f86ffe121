"""Kernel breakdown of one pivot-translator training step (eager, bf16 autocast) with torch.profiler: count and total
device time per kernel name.  Usage: python scripts/pivot_profile.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from unpaired_image_captioning_b200 import pivot  # noqa: E402

torch.manual_seed(0)
gen = torch.Generator().manual_seed(1)
B = 256
src, n = pivot.sentences(B, 12000, gen, lo=5, hi=30, max_len=30)
tgt, _ = pivot.sentences(B, 8600, gen, lo=5, hi=30, bos=pivot.BOS, max_len=30)
m = pivot.PivotNMT().cuda().train()
step = pivot.PivotTrainStep(m, src.size(0), tgt.size(0), B, graph=False, batch=(src.cuda(), n.cuda(), tgt.cuda()))
for _ in range(3):
    step.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step.step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"kernels {sum(r[1] for r in rows)}  device time {tot / 1e3:.2f} ms")
for k, c, t in rows[:28]:
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  n={c:5d}  avg {t / c:6.1f} us  {k[:100]}")
