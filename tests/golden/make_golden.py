"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py [case ...]

For each case it builds the reference model with `models.setup(opt)`, loads a seeded state_dict,
feeds seeded synthetic inputs through the reference's public API (the call patterns of
trainer.py:164-165 and eval_utils.py:263) and stores inputs, weights and outputs in one `.npz`.
The fixtures pin `oracle/decoder_oracle.py` (tests/test_oracle_golden.py) and are also compared with
the CUDA path (tests/test_gpu_golden.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_shim  # noqa: E402
from unpaired_image_captioning_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (synth config, seed, weight variant, use att_masks).  The "peaked" variants scale
# logit.weight (wide top-2 margins, so token ids are comparable exactly under bf16 operands) and
# bias EOS so beams finish at different lengths; seeds were picked so
# that caption lengths differ across images.
CASES = {
    "att2in2_plain": ("tiny_att2in2", 1234, dict(), False),
    "att2in2_peaked": ("tiny_att2in2", 1238, dict(peaked=80.0, eos_bias=1.0), False),
    "att2in2_masked": ("tiny_att2in2", 1238, dict(peaked=80.0, eos_bias=1.0), True),
    "att2all2_peaked": ("tiny_att2all2", 1238, dict(peaked=80.0, eos_bias=1.0), False),
    "topdown_plain": ("tiny_topdown", 1234, dict(), False),
    "topdown_peaked": ("tiny_topdown", 1237, dict(peaked=80.0, eos_bias=0.0), False),
    "topdown_masked": ("tiny_topdown", 1237, dict(peaked=80.0, eos_bias=0.0), True),
    "stackatt_plain": ("tiny_stackatt", 1234, dict(), False),
    "stackatt_peaked": ("tiny_stackatt", 1256, dict(peaked=80.0, eos_bias=2.0), False),
    "denseatt_peaked": ("tiny_denseatt", 1259, dict(peaked=80.0, eos_bias=0.5), False),
    "denseatt_masked": ("tiny_denseatt", 1259, dict(peaked=80.0, eos_bias=0.5), True),
}


def run_case(models, criterion, name, cfg_name, seed, variant, use_masks):
    opt, cfg = synth.opt_for(cfg_name)
    sd = synth.init_state_dict(opt, seed=seed, **variant)
    model = models.setup(opt)
    model.load_state_dict(sd, strict=True)
    model.eval()
    crit = criterion.LanguageModelCriterion(opt)

    B, L, T = cfg["batch"], cfg["att_size"], cfg["seq_length"]
    fc, att = synth.make_features(B, L, opt.att_feat_size, seed=seed)
    labels, masks = synth.make_captions(B, T, opt.vocab_size, seed=seed, min_len=2)
    att_masks = synth.make_att_masks(B, L, seed=seed) if use_masks else None

    out = {}
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()
    out["in/fc"], out["in/att"] = fc.numpy(), att.numpy()
    out["in/labels"], out["in/masks"] = labels.numpy(), masks.numpy()
    if att_masks is not None:
        out["in/att_masks"] = att_masks.numpy()

    # teacher-forced forward + XE + backward (trainer.py:164-165,173)
    model.zero_grad()
    logprobs = model(fc, None, att, labels, att_masks)
    loss = crit(logprobs, labels[:, 1:], masks[:, 1:])
    loss.backward()
    out["out/logprobs"] = logprobs.detach().numpy()
    out["out/loss"] = np.float32(loss.item())
    for k, p in model.named_parameters():
        out["grad/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()

    with torch.no_grad():
        # greedy (eval_utils.py:263 with beam_size 1)
        seq, lp = model(fc, None, att, att_masks, opt={"beam_size": 1}, mode="sample")
        out["greedy/seq"], out["greedy/lp"] = seq.numpy(), lp.numpy()
        seq, lp = model(fc, None, att, att_masks, opt={"beam_size": 1, "decoding_constraint": 1}, mode="sample")
        out["greedy_dc/seq"], out["greedy_dc/lp"] = seq.numpy(), lp.numpy()
        # beam search variants (models/CaptionModel.py:100-106 options)
        for tag, o in (("beam3", {"beam_size": 3}),
                       ("beam3_dc", {"beam_size": 3, "decoding_constraint": 1}),
                       ("beam3_ppl", {"beam_size": 3, "max_ppl": 1}),
                       ("beam5", {"beam_size": 5}),
                       ("beam2", {"beam_size": 2})):
            seq, lp = model(fc, None, att, att_masks, opt=dict(o), mode="sample")
            out[tag + "/seq"], out[tag + "/lp"] = seq.numpy().copy(), lp.numpy().copy()
            b = o["beam_size"]
            done_p = np.full((B, b), -np.inf, dtype=np.float64)
            done_seq = np.zeros((B, b, T), dtype=np.int64)
            for k in range(B):
                for j, d in enumerate(model.done_beams[k][:b]):
                    done_p[k, j] = d["p"]
                    done_seq[k, j] = d["seq"].numpy()
            out[tag + "/done_p"], out[tag + "/done_seq"] = done_p, done_seq

    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: loss={loss.item():.6f} -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    models, criterion = reference_shim.load()
    torch.set_num_threads(1)
    only = sys.argv[1:]          # optional: case names to (re)generate; default all
    for name, (cfg_name, seed, variant, use_masks) in CASES.items():
        if only and name not in only:
            continue
        run_case(models, criterion, name, cfg_name, seed, variant, use_masks)


if __name__ == "__main__":
    main()
