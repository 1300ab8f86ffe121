"""Detail of one shape_sweep case: gradient norms per parameter (device vs oracle)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
import shape_cases
import unpaired_image_captioning_b200 as uic
from oracle import decoder_oracle as O
from unpaired_image_captioning_b200 import synth

seed, case = int(sys.argv[1]), int(sys.argv[2])
c = shape_cases.cases(case + 1, seed)[case]
print(c)
kind = c["kind"]
opt = synth.make_opt(caption_model=kind, vocab_size=c["V"], rnn_size=c["H"], input_encoding_size=c["E"], att_hid_size=c["A"],
                     seq_length=c["T"], fc_feat_size=c["D"], att_feat_size=c["D"])
sd = synth.init_state_dict(opt, seed=100 + case)
fc, att = synth.make_features(c["B"], c["L"], c["D"], seed=100 + case)
labels, masks = synth.make_captions(c["B"], c["T"], c["V"], seed=100 + case, min_len=1)
am = synth.make_att_masks(c["B"], c["L"], seed=100 + case) if c["use_masks"] else None
print("att_masks", None if am is None else am.tolist())
model = uic.setup(opt); model.load_state_dict(sd); model = model.cuda().train()
cu = lambda t: None if t is None else t.cuda()
ref_loss, ref_grads = O.loss_and_grads(sd, kind, fc, att, labels, masks, am)
loss = model(cu(fc), None, cu(att), cu(labels), cu(masks), cu(am), mode="forward_loss"); loss.backward()
for name, p in model.named_parameters():
    r = ref_grads[name]
    print(f"{name:40s} ref norm {float(r.norm()):.3e}  dev norm {float(p.grad.norm()):.3e}  diff {float((p.grad.cpu() - r).norm()):.3e}")
