"""Pipeline timeline of CTA 0 of the statistics GEMM (logit projection of the decode loops): TMA issue, operand landing and
MMA issue times per k-block of its first tiles (debug aid, uic_gemm_set_trace)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unpaired_image_captioning_b200 import _lib  # noqa: E402

_lib.require_device()
lib = _lib.load()
M, N, K, ks = 768, 10000, 512, 3
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
b = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
parts = int(lib.uic_logit_stats_parts(M, N))
stats = torch.empty(M, parts, int(lib.uic_logit_stats_entry_floats(ks)), device="cuda")
trace = torch.zeros(1024, dtype=torch.int64, device="cuda")


def run():
    _lib.check(lib.uic_logit_stats(a.data_ptr(), K, b.data_ptr(), K, bias.data_ptr(), None, 1, stats.data_ptr(), M, N, K, ks, 1, 0.0, None, 0,
                                   torch.cuda.current_stream().cuda_stream))


for _ in range(3):
    run()
torch.cuda.synchronize()
trace.zero_()
lib.uic_gemm_set_trace(trace.data_ptr())
run()
torch.cuda.synchronize()
lib.uic_gemm_set_trace(None)
t = trace.cpu().tolist()
t0 = t[50]
nkb = K // 64
print("tma issue kb:", [t[50 + i] - t0 for i in range(nkb)])
print("landed (mma thread saw full) kb:", [t[1 + i] - t0 for i in range(nkb)])
print("mma loop start", t[0] - t0, "last commit", t[40] - t0)
ent = [t[128 + 2 * i] for i in range(148) if t[128 + 2 * i]]
ext = [t[129 + 2 * i] for i in range(148) if t[129 + 2 * i]]
print("CTAs", len(ent), "kernel span ns", max(ext) - min(ent), "first CTA life ns", t[129] - t[128])
