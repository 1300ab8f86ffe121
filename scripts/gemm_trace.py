"""Timeline of CTA (0,0) of the tcgen05 GEMM at the decoder's shapes + event-timed throughput."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unpaired_image_captioning_b200 import _lib  # noqa: E402

_lib.require_device()
lib = _lib.load()
SHAPES = [(768, 1024, 512), (768, 3072, 1024), (768, 10000, 512), (256, 10000, 512), (50176, 512, 2048), (8704, 10000, 512)]
if len(sys.argv) > 1:   # e.g. "384x3072x1024,128x3072x1024": how the k-block period scales with the number of CTAs
    SHAPES = [tuple(int(v) for v in sh.split("x")) for sh in sys.argv[1].split(",")]
trace = torch.zeros(1024, dtype=torch.int64, device="cuda")
for M, N, K in SHAPES:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    for _ in range(3):
        _lib.gemm(a, b, bias, out_f32=out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    s.record()
    for _ in range(n):
        _lib.gemm(a, b, bias, out_f32=out)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / n * 1e3
    for _ in range(3):  # cuBLAS for comparison (library baseline)
        torch.matmul(a, b.t())
    torch.cuda.synchronize()
    s.record()
    for _ in range(n):
        torch.matmul(a, b.t())
    e.record()
    torch.cuda.synchronize()
    us_cublas = s.elapsed_time(e) / n * 1e3
    trace.zero_()
    lib.uic_gemm_set_trace(trace.data_ptr())
    _lib.gemm(a, b, bias, out_f32=out)
    torch.cuda.synchronize()
    lib.uic_gemm_set_trace(None)
    t = trace.cpu().tolist()
    nkb = (K + 63) // 64
    t0 = t[50]   # first TMA issue of CTA 0's first tile

    def rel(i):
        return t[i] - t0

    print(f"--- {M}x{N}x{K}: {us:.1f} us ({2 * M * N * K / us * 1e-6:.0f} TF/s), cuBLAS bf16 {us_cublas:.1f} us; nkb={nkb}")
    print(f"    setup done @{rel(0)}  tma issue kb0..: {[rel(50 + i) for i in range(min(nkb, 8))]}")
    print(f"    operands landed kb0..: {[rel(1 + i) for i in range(min(nkb, 32))]}")
    print(f"    last mma issued @{rel(40)}  acc ready @{rel(41)}  epilogue done @{rel(42)}  (persistent: first tile of CTA 0)")
    print(f"    epilogue warp 4: chunk 0 tmem loaded @{rel(43)} staged @{rel(44)} stored @{rel(45)}; chunk 1 @{rel(46)} @{rel(47)} @{rel(48)}")
    print(f"    CTA 0: entry @{rel(120)}  set-up done (barriers, TMEM allocation) @{rel(121)}  dependency wait passed @{rel(122)}  first TMA issue @0")
    ent = [t[128 + 2 * c] for c in range(148) if t[128 + 2 * c]]
    ext = [t[129 + 2 * c] for c in range(148) if t[129 + 2 * c]]
    if ent and ext:
        print(f"    CTA lifetimes (globaltimer ns): first entry -> last exit {max(ext) - min(ent)}; entry spread {max(ent) - min(ent)}; "
              f"median lifetime {sorted(x - e_ for x, e_ in zip(ext, ent))[len(ent) // 2]}")
