"""One eager (non-graph) pass of the benchmark step between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python scripts/profile_step.py [beam|greedy|train]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import unpaired_image_captioning_b200 as uic  # noqa: E402
from unpaired_image_captioning_b200 import synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "beam"
cfg_name = sys.argv[2] if len(sys.argv) > 2 else ("cfg3" if mode == "train" else "cfg2")
opt, cfg = synth.opt_for(cfg_name)
model = uic.setup(opt)
model.load_state_dict(synth.init_state_dict(opt, seed=1234))
model = model.cuda().eval()
eng = model.engine
eng.use_graphs = False
fc, att = synth.make_features(cfg["batch"], cfg["att_size"], opt.att_feat_size, seed=1234)
fc, att = fc.cuda(), att.cuda()


train_state = None
if mode == "train":
    from unpaired_image_captioning_b200 import train_bench
    train_state = train_bench.make_state(0, 0, 1, cfg_name=cfg_name)


def step():
    if mode == "train":
        train_bench.one_train_step(st=train_state)
        return
    feats = eng.prepare(fc, att, lazy=True)
    if mode == "beam":
        eng.beam(feats, opt.seq_length, cfg["beam_size"])
    else:
        eng.greedy(feats, opt.seq_length)


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
