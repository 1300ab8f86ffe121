"""configs[4] of BASELINE.json at its per-GPU chunk size (500 images, rnn 1024, vocab 30k, 20 steps, beam 5): runs the
decode a few times and prints captions/s plus a consistency check against a 16-image sub-batch (scale sanity)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import unpaired_image_captioning_b200 as uic  # noqa: E402
from unpaired_image_captioning_b200 import synth  # noqa: E402

opt, cfg = synth.opt_for("cfg5")
B = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["batch"]
model = uic.setup(opt)
model.load_state_dict(synth.init_state_dict(opt, seed=3, peaked=8.0, eos_bias=0.0))
model = model.cuda().eval()
fc, att = synth.make_features(B, cfg["att_size"], opt.att_feat_size, seed=3)
fc, att = fc.cuda(), att.cuda()
o = {"beam_size": cfg["beam_size"]}
with torch.no_grad():
    seq, lp = model(fc, None, att, None, opt=o, mode="sample")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        seq, lp = model(fc, None, att, None, opt=o, mode="sample")
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    sub, _ = model(fc[:16], None, att[:16], None, opt=o, mode="sample")
print(f"cfg5 chunk of {B} images, beam {cfg['beam_size']}: {dt * 1e3:.1f} ms per chunk = {B / dt:.0f} captions/s (incl. prologue, D2H of results)")
same = (seq[:16] == sub).all(1).float().mean()
print("rows of the first 16 images identical to a 16-image run:", float(same), "| mean caption length", float((seq > 0).sum(1).float().mean()))
print("max memory allocated GB:", torch.cuda.max_memory_allocated() / 1e9)
