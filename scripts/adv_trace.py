"""Phase timeline of CTA 0 of uic_beam_advance inside a real beam decode (debug aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import unpaired_image_captioning_b200 as uic  # noqa: E402
from unpaired_image_captioning_b200 import _lib, synth  # noqa: E402

_lib.require_device()
lib = _lib.load()
opt = synth.make_opt(caption_model="att2in2", vocab_size=9999, rnn_size=512, input_encoding_size=512, att_hid_size=512, seq_length=16)
model = uic.setup(opt)
model.load_state_dict(synth.init_state_dict(opt, seed=1))
model = model.cuda().eval()
model.engine.use_graphs = False
fc, att = synth.make_features(256, 196, 2048, seed=1)
fc, att = fc.cuda(), att.cuda()
trace = torch.zeros(1024, dtype=torch.int64, device="cuda")
with torch.no_grad():
    model(fc, None, att, None, opt={"beam_size": 3}, mode="sample")
    lib.uic_gemm_set_trace(trace.data_ptr())
    model(fc, None, att, None, opt={"beam_size": 3}, mode="sample")
    torch.cuda.synchronize()
    lib.uic_gemm_set_trace(None)
t = trace.tolist()
print("last beam_advance of the decode, CTA 0 (us): merge %.2f, beam step %.2f, state move %.2f" %
      ((t[111] - t[110]) / 1e3, (t[112] - t[111]) / 1e3, (t[113] - t[112]) / 1e3))
