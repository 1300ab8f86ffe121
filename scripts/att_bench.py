"""Attention-step kernel alone: in-graph duration (32 back-to-back launches per replay) at the benchmark shapes and a
check against the fp32 torch formula.  UIC_ATT_V7=0 selects the v6 kernel everywhere (A/B in two processes).
Usage: python scripts/att_bench.py [check]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unpaired_image_captioning_b200 import _lib  # noqa: E402

_lib.require_device()
N_REP = 32


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


def reference(p_att, att, att_h, w, beams, masks=None):
    B, L, A = p_att.shape
    R = att_h.shape[0]
    pa = p_att.repeat_interleave(beams, 0)
    e = (torch.tanh(pa + att_h[:, None, :]) * w).sum(-1)
    al = torch.softmax(e, -1)
    if masks is not None:
        al = al * masks.repeat_interleave(beams, 0)
        al = al / al.sum(-1, keepdim=True)
    return torch.bmm(al[:, None, :], att.float().repeat_interleave(beams, 0)).squeeze(1), al


tag = "v6" if os.environ.get("UIC_ATT_V7") == "0" else "v7"
cases = [(256, 3, 196, 512, 512), (256, 1, 196, 512, 512), (512, 1, 36, 512, 512), (500, 5, 196, 512, 1024), (148, 3, 196, 512, 512),
         (300, 2, 100, 256, 512), (37, 3, 196, 512, 512), (256, 3, 196, 1024, 1024)]
for B, beams, L, A, H in cases:
    torch.manual_seed(B + beams)
    R = B * beams
    p_att = (torch.randn(B, L, A) * 0.5).cuda()
    att = torch.randn(B, L, H).cuda().to(torch.bfloat16)
    att_h = torch.randn(R, A).cuda()
    w = (torch.randn(A) * 0.2).cuda()
    e_tile = _lib.exp_tile(p_att)
    f = (torch.exp(2.0 * att_h) * _lib.ATT_F_SCALE).contiguous()
    ctx = torch.zeros(R, H, device="cuda", dtype=torch.bfloat16)
    ctx32 = torch.zeros(R, H, device="cuda")
    alpha = torch.zeros(R, L, device="cuda")
    run = lambda: _lib.att_step(f, A, e_tile, att, w, None, ctx, H, None, 0, None, B, beams, L, A, H)  # noqa: E731
    run()
    _lib.att_step(f, A, e_tile, att, w, None, None, 0, ctx32, H, alpha, B, beams, L, A, H)
    torch.cuda.synchronize()
    ref, al = reference(p_att, att, att_h, w, beams)
    err = (ctx.float() - ref).abs().max().item()
    err32 = (ctx32 - ref).abs().max().item()
    erra = (alpha - al).abs().max().item()
    t = time_graph(run)
    t_aux = time_graph(lambda: _lib.att_step(f, A, e_tile, att, w, None, None, 0, ctx32, H, alpha, B, beams, L, A, H))   # training forward: fp32 ctx + alpha
    gb = B * L * (A + H) * 2 / 1e9
    print(f"{tag} att_step_fwd {B}x{beams}x{L} A={A} H={H}: {t:.1f} us  {gb / t * 1e6:.0f} GB/s   (with alpha output: {t_aux:.1f} us)   max|ctx-ref| bf16 {err:.4f} f32 {err32:.5f} alpha {erra:.6f}", flush=True)
