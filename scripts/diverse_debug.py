"""Debug: diverse beam search with one beam per group -- group 0 must equal greedy decoding."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
import unpaired_image_captioning_b200 as uic
from oracle import decoder_oracle as O
from unpaired_image_captioning_b200 import synth
from parity import compare_greedy

opt = synth.make_opt(caption_model="att2in2", vocab_size=9999, rnn_size=512, input_encoding_size=512, att_hid_size=512, seq_length=16)
sd = synth.init_state_dict(opt, seed=7, peaked=40.0, eos_bias=2.0)
fc, att = synth.make_features(5, 36, 2048, seed=7)
model = uic.setup(opt); model.load_state_dict(sd); model = model.cuda().eval()
g_ref, g_lp, margins = O.sample_greedy(sd, "att2in2", fc, att, 16, return_margins=True)
o = {"beam_size": 4, "group_size": 4, "diversity_lambda": 0.5}
ref_seq, ref_lp, ref_done = O.sample_beam(sd, "att2in2", fc, att, 16, 4, group_size=4, diversity_lambda=0.5)
seq, lp = model(fc.cuda(), None, att.cuda(), None, opt=o, mode="sample")
gs, gl = model(fc.cuda(), None, att.cuda(), None, opt={"beam_size": 1}, mode="sample")
print("margins (oracle greedy):\n", margins)
for k in range(5):
    print("img", k)
    print("  oracle greedy ", g_ref[k].tolist())
    print("  oracle dbs g0 ", ref_seq[k].tolist())
    print("  device greedy ", gs[k].tolist())
    print("  device dbs g0 ", seq[k].tolist())
    for g in range(4):
        print("   grp", g, "ref", ref_done[k][g]["seq"].tolist(), round(float(ref_done[k][g]["p"]), 3))
        print("   grp", g, "dev", model.done_beams[k][g]["seq"].tolist(), round(model.done_beams[k][g]["p"], 3))
