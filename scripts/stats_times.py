"""In-graph time of the statistics GEMM (uic_logit_stats) by kslots at the decode shapes of configs[1] and configs[4]:
how much of the kernel is the top-k epilogue.  Usage: python scripts/stats_times.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unpaired_image_captioning_b200 import _lib  # noqa: E402

_lib.require_device()
lib = _lib.load()


def time_graph(fn, n_rep=20):
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for _ in range(n_rep):
                fn()
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            g.replay()
        e.record()
        torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / (5 * n_rep)


for M, N, K in [(768, 10000, 512), (2500, 30000, 1024), (256, 10000, 512)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    parts = int(lib.uic_logit_stats_parts(M, N))
    line = [f"logit_stats {M}x{N}x{K} ({parts} parts):"]
    for ks in (1, 3, 5, 8):
        stats = torch.empty(M, parts, int(lib.uic_logit_stats_entry_floats(ks)), device="cuda")

        def st():
            _lib.check(lib.uic_logit_stats(a.data_ptr(), K, b.data_ptr(), K, bias.data_ptr(), None, 1, stats.data_ptr(), M, N, K, ks, 1, 0.0, None, 0,
                                           torch.cuda.current_stream().cuda_stream))

        t = time_graph(st)
        line.append(f"kslots={ks} {t:.1f} us ({2.0 * M * N * K / t * 1e-6:.0f} TF/s)")
    ref = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    t = time_graph(lambda: torch.matmul(a, b.t(), out=ref))
    line.append(f"cuBLAS bf16 out {t:.1f} us")
    print("  ".join(line), flush=True)
