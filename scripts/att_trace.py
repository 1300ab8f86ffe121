"""Pipeline timeline of CTA 0 of the attention kernel at the bench shape + event-timed throughput (debug aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unpaired_image_captioning_b200 import _lib  # noqa: E402

_lib.require_device()
lib = _lib.load()
B, beams, L, A, H = (int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (256, 3, 196, 512, 512)))
R = B * beams
g = torch.Generator(device="cpu").manual_seed(0)
p_att = (torch.randn(B, L, A, generator=g) * 0.5).cuda()
att = torch.randn(B, L, H, generator=g).cuda().to(torch.bfloat16)
att_h = torch.randn(R, A, generator=g).cuda()
w = (torch.randn(A, generator=g) * 0.2).cuda()
e_tile = _lib.exp_tile(p_att)
f = (torch.exp(2.0 * att_h) * _lib.ATT_F_SCALE).contiguous()
ctx = torch.empty(R, H, device="cuda", dtype=torch.bfloat16)


def run():
    _lib.att_step(f, A, e_tile, att, w, None, ctx, H, None, 0, None, B, beams, L, A, H)


for _ in range(3):
    run()
torch.cuda.synchronize()
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(10):
    big.zero_()   # flush L2
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    run()
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e) * 1e3)
print("us per launch (L2 flushed):", [round(t, 1) for t in ts])
trace = torch.zeros(1024, dtype=torch.int64, device="cuda")
lib.uic_gemm_set_trace(trace.data_ptr())
run()
torch.cuda.synchronize()
lib.uic_gemm_set_trace(None)
t = trace.tolist()
t0 = t[0]
print("CTA 0 warp 0 timeline (us after start): batch: data ready / region A scored / previous batch finished / region B scored")
for i in range(24):
    ev = t[1 + 4 * i: 5 + 4 * i]
    if not any(ev):
        break
    print(i, [round((x - t0) / 1e3, 2) if x else None for x in ev])
print("end", round((t[100] - t0) / 1e3, 2))
print("context: after scored-wait", [round((x - t0) / 1e3, 2) for x in t[50:70] if x])
print("context: after MMA       ", [round((x - t0) / 1e3, 2) for x in t[70:90] if x])

import statistics
starts = [t[128 + 2 * c] for c in range(448) if t[128 + 2 * c]]
ends = [t[129 + 2 * c] for c in range(448) if t[129 + 2 * c]]
if starts:
    z = min(starts)
    print("CTAs", len(starts), "start spread us", round((max(starts) - z) / 1e3, 2))
    d = sorted((e - z) / 1e3 for e in ends)
    print("end times us: min", round(d[0], 1), "p25", round(d[len(d) // 4], 1), "median", round(statistics.median(d), 1), "p75", round(d[3 * len(d) // 4], 1), "max", round(d[-1], 1))
    per = [(t[129 + 2 * c] - z) / 1e3 for c in range(len(starts))]
    print("end by CTA (every 8th):", [round(x, 1) for x in per[::8]])
