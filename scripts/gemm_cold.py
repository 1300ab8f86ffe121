"""Timeline of CTA 0 of the decoder's per-step GEMMs with a COLD L2 (as inside the decode loop, where the
attention kernel streams 200 MB between two uses of the weights) -- debug aid."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unpaired_image_captioning_b200 import _lib  # noqa: E402

_lib.require_device()
lib = _lib.load()
big = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
trace = torch.zeros(1024, dtype=torch.int64, device="cuda")
for M, N, K in [(768, 3072, 1024), (768, 1024, 512), (768, 10000, 512)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    for _ in range(3):
        _lib.gemm(a, b, bias, out_f32=out)
    for cold in (False, True):
        ts = []
        for _ in range(8):
            if cold:
                big.zero_()
            else:
                _lib.gemm(a, b, bias, out_f32=out)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            _lib.gemm(a, b, bias, out_f32=out)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) * 1e3)
        if cold:
            big.zero_()
        trace.zero_()
        lib.uic_gemm_set_trace(trace.data_ptr())
        _lib.gemm(a, b, bias, out_f32=out)
        torch.cuda.synchronize()
        lib.uic_gemm_set_trace(None)
        t = trace.cpu().tolist()
        nkb = (K + 63) // 64
        t0 = t[50]
        print(f"--- {M}x{N}x{K} {'COLD' if cold else 'warm'}: single-launch event time (incl. ~launch overhead) min {min(ts):.1f} us")
        print(f"    operands landed kb0..: {[t[1 + i] - t0 for i in range(min(nkb, 16))]}")
        print(f"    last mma @{t[40] - t0} acc ready @{t[41] - t0} epilogue done @{t[42] - t0} cycles")
        st = [t[128 + 2 * c] for c in range(448) if t[128 + 2 * c]]
        en = [t[129 + 2 * c] for c in range(448) if t[129 + 2 * c]]
        z = min(st)
        print(f"    CTAs {len(st)}: entry spread {(max(st) - z) / 1e3:.2f} us, exits min {(min(en) - z) / 1e3:.2f} median {(sorted(en)[len(en) // 2] - z) / 1e3:.2f} max {(max(en) - z) / 1e3:.2f} us")
