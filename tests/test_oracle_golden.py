"""The oracle restatement against the fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only; this is what pins `oracle/decoder_oracle.py`."""
import numpy as np
import pytest
import torch

from oracle import decoder_oracle as O

TOL = dict(rtol=1e-5, atol=2e-6)


def _inputs(g):
    i = g["in"]
    return i["fc"], i["att"], i["labels"], i["masks"], i.get("att_masks")


def test_teacher_forced_logprobs_and_loss(golden):
    fc, att, labels, masks, am = _inputs(golden)
    out = O.teacher_forced(golden["sd"], golden["kind"], fc, att, labels, am)
    torch.testing.assert_close(out, golden["out"]["logprobs"], **TOL)
    loss = O.xe_loss(out, labels[:, 1:], masks[:, 1:])
    torch.testing.assert_close(loss, golden["out"]["loss"], **TOL)


def test_gradients(golden):
    fc, att, labels, masks, am = _inputs(golden)
    _, grads = O.loss_and_grads(golden["sd"], golden["kind"], fc, att, labels, masks, am)
    for k, ref in golden["grad"].items():
        torch.testing.assert_close(grads[k], ref, rtol=1e-4, atol=1e-6, msg=lambda m, k=k: f"{k}: {m}")


@pytest.mark.parametrize("tag,dc", [("greedy", 0), ("greedy_dc", 1)])
def test_greedy(golden, tag, dc):
    fc, att, _, _, am = _inputs(golden)
    T = golden[tag]["seq"].shape[1]
    seq, lp = O.sample_greedy(golden["sd"], golden["kind"], fc, att, T, am, decoding_constraint=dc)
    assert torch.equal(seq, golden[tag]["seq"])
    torch.testing.assert_close(lp, golden[tag]["lp"], **TOL)


@pytest.mark.parametrize("tag,opts", [("beam3", dict(beam_size=3)),
                                      ("beam3_dc", dict(beam_size=3, decoding_constraint=1)),
                                      ("beam3_ppl", dict(beam_size=3, max_ppl=1)),
                                      ("beam5", dict(beam_size=5)),
                                      ("beam2", dict(beam_size=2))])
def test_beam(golden, tag, opts):
    fc, att, _, _, am = _inputs(golden)
    T = golden[tag]["seq"].shape[1]
    seq, lp, done = O.sample_beam(golden["sd"], golden["kind"], fc, att, T, att_masks=am, **opts)
    assert torch.equal(seq, golden[tag]["seq"])
    torch.testing.assert_close(lp, golden[tag]["lp"], **TOL)
    for k, beams in enumerate(done):
        for j, d in enumerate(beams):
            assert np.isclose(d["p"], float(golden[tag]["done_p"][k, j]), rtol=1e-5, atol=1e-5)
            assert torch.equal(d["seq"], golden[tag]["done_seq"][k, j])


@pytest.mark.parametrize("name", ["att2in2_peaked", "att2in2_masked", "topdown_peaked", "topdown_masked"])
def test_fixture_exercises_eos(name):
    """The peaked variants must finish beams at different lengths, otherwise the EOS bookkeeping
    of CaptionModel.beam_search (:155-167) is never pinned."""
    from conftest import load_golden
    g = load_golden(name)
    lens = set()
    for tag in ("beam3", "beam3_ppl", "greedy"):
        lens |= set((g[tag]["seq"] > 0).sum(1).tolist())
    assert len(lens) >= 3, lens
