"""Steady-state duration of single kernels inside a CUDA graph (50 back-to-back launches per replay), optionally
interleaved with an L2-evicting kernel -- debug aid for separating kernel time from launch / profiler overheads."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unpaired_image_captioning_b200 import _lib  # noqa: E402

_lib.require_device()
lib = _lib.load()
N_REP = 50


def time_graph(fn):
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for _ in range(N_REP):
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
    return s.elapsed_time(e) * 1e3 / (5 * N_REP)


evict = torch.empty(200 << 20, dtype=torch.uint8, device="cuda")
t_evict = time_graph(lambda: evict.zero_())
print(f"evict (zero 200 MB): {t_evict:.1f} us")
for M, N, K in [(768, 3072, 512), (768, 3072, 1024), (768, 1024, 512), (768, 10000, 512), (256, 2560, 1024)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    t_warm = time_graph(lambda: _lib.gemm(a, b, bias, out_f32=out))

    def both():
        evict.zero_()
        _lib.gemm(a, b, bias, out_f32=out)

    t_cold = time_graph(both) - t_evict
    ref = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    t_cublas = time_graph(lambda: torch.matmul(a, b.t(), out=ref))
    # like for like: the library GEMM above writes bf16 (2 bytes per output) and no bias; the decode loop's GEMMs write
    # fp32 gate sums.  Same output format for ours, and cuBLAS with the bias (addmm):
    out16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    t_warm16 = time_graph(lambda: _lib.gemm(a, b, bias, out_bf16=out16))
    bias16 = bias.to(torch.bfloat16)
    t_cublas_bias = time_graph(lambda: torch.addmm(bias16, a, b.t(), out=ref))
    print(f"gemm {M}x{N}x{K}: in-graph warm {t_warm:.1f} us (fp32 out), {t_warm16:.1f} us (bf16 out), after eviction {t_cold:.1f} us; "
          f"cuBLAS warm, bf16 out {t_cublas:.1f} us, with bias {t_cublas_bias:.1f} us")
# the fused-statistics variant of the logit GEMM (what the decode loops launch)
M, N, K = 768, 10000, 512
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
b = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
parts = int(lib.uic_logit_stats_parts(M, N))
for ks in (1, 3):
    stats = torch.empty(M, parts, int(lib.uic_logit_stats_entry_floats(ks)), device="cuda")

    def st():
        _lib.check(lib.uic_logit_stats(a.data_ptr(), K, b.data_ptr(), K, bias.data_ptr(), None, 1, stats.data_ptr(), M, N, K, ks, 1, 0.0, None, 0,
                                       torch.cuda.current_stream().cuda_stream))

    def st_cold():
        evict.zero_()
        st()

    print(f"logit_stats {M}x{N}x{K} kslots={ks}: in-graph warm {time_graph(st):.1f} us, after eviction {time_graph(st_cold) - t_evict:.1f} us")
if len(sys.argv) > 1:
    sys.exit(0)
B, beams, L, A, H = 256, 3, 196, 512, 512
R = B * beams
p_att = (torch.randn(B, L, A) * 0.5).cuda()
att = torch.randn(B, L, H).cuda().to(torch.bfloat16)
att_h = torch.randn(R, A).cuda()
w = (torch.randn(A) * 0.2).cuda()
e_tile = _lib.exp_tile(p_att)
f = (torch.exp(2.0 * att_h) * _lib.ATT_F_SCALE).contiguous()
ctx = torch.empty(R, H, device="cuda", dtype=torch.bfloat16)
t_att = time_graph(lambda: _lib.att_step(f, A, e_tile, att, w, None, ctx, H, None, 0, None, B, beams, L, A, H))
print(f"att_step_fwd {B}x{beams}x{L}: in-graph {t_att:.1f} us")
