"""Helpers shared by the GPU parity tests (the comparison rules themselves live in oracle/compare.py)."""
import torch  # noqa: F401

from oracle.compare import compare_beam, compare_greedy  # noqa: F401


def load_model(uic, synth, sd, kind, opt_kwargs, device="cuda"):
    opt = synth.make_opt(caption_model=kind, **opt_kwargs)
    model = uic.setup(opt)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval(), opt


def opt_kwargs_from_sd(sd, kind, seq_length):
    V, E = sd["embed.0.weight"].shape
    use_bn = int("att_embed.1.weight" in sd)            # BatchNorm1d first shifts the Linear to index 1 (AttModel.py:79-84)
    H, D = sd["att_embed.%d.weight" % use_bn].shape
    A = sd["ctx2att.weight"].shape[0]
    F = sd["fc_embed.0.weight"].shape[1] if "fc_embed.0.weight" in sd else D
    return dict(vocab_size=V - 1, rnn_size=H, input_encoding_size=E, att_hid_size=A, seq_length=seq_length,
                fc_feat_size=F, att_feat_size=D, use_bn=use_bn)
