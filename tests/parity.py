"""Helpers for comparing token sequences under the north-star tolerance: ids must be identical
except where the oracle's top-2 log-prob margin at the first differing step is inside the tolerance
(bf16 operands flip near-ties).  After an exempted flip the rest of that
row is not comparable (the two decoders follow different prefixes) and is skipped."""
import torch


def compare_greedy(seq, ref_seq, ref_margins, tol):
    """Returns (n_exact_rows, n_exempt_rows, failures[list of (row, step, margin)])."""
    exact = exempt = 0
    failures = []
    for r in range(ref_seq.size(0)):
        diff = (seq[r] != ref_seq[r]).nonzero()
        if diff.numel() == 0:
            exact += 1
            continue
        t = int(diff[0])
        m = float(ref_margins[r, t])
        if m < tol:
            exempt += 1
        else:
            failures.append((r, t, m))
    return exact, exempt, failures


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
