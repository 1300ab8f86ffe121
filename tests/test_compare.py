"""The comparison rules of the parity harness (oracle/compare.py) and the decision margins the oracle reports."""
import torch

from oracle import decoder_oracle as O
from oracle.compare import compare_beam, compare_greedy
from unpaired_image_captioning_b200 import synth


def test_compare_greedy_exempts_only_near_ties_at_the_first_difference():
    ref = torch.tensor([[3, 4, 0], [5, 6, 7], [1, 2, 3]])
    got = torch.tensor([[3, 4, 0], [5, 9, 9], [1, 8, 3]])
    margins = torch.tensor([[1.0, 1.0, 1.0], [1.0, 5e-4, 1.0], [1.0, 2e-3, 1e-9]])
    exact, exempt, failures = compare_greedy(got, ref, margins, tol=1e-3)
    assert (exact, exempt) == (1, 1)
    assert failures == [(2, 1, margins[2, 1].item())]


def test_compare_beam():
    ref = torch.tensor([[3, 4, 0], [5, 6, 7], [1, 2, 3]])
    got = torch.tensor([[3, 4, 0], [5, 9, 9], [1, 8, 3]])
    exact, exempt, failures = compare_beam(got, ref, torch.tensor([0.5, 5e-4, 2e-3]), tol=1e-3)
    assert (exact, exempt) == (1, 1) and [f[0] for f in failures] == [2]


def test_rel_gaps_ignores_the_bookkeeping_offsets():
    # two children of a finished beam (sum -1000): their gap is judged against |-9.2|, not against 1009
    g = O._rel_gaps(torch.tensor([-1009.20, -1009.21]))
    assert abs(g - 0.01 / 9.21) < 1e-4
    assert O._rel_gaps(torch.tensor([-3.0])) == float("inf")
    assert O._rel_gaps(torch.tensor([-3.0, float("-inf")])) == float("inf")


def test_beam_margins_do_not_change_the_search_and_bound_perturbations():
    opt, cfg = synth.opt_for("tiny_att2in2")
    sd = synth.init_state_dict(opt, seed=1238, peaked=20.0, eos_bias=1.0)
    fc, att = synth.make_features(6, cfg["att_size"], opt.att_feat_size, seed=1238)
    seq, lp, done = O.sample_beam(sd, "att2in2", fc, att, opt.seq_length, 3)
    seq_m, lp_m, done_m, margins = O.sample_beam(sd, "att2in2", fc, att, opt.seq_length, 3, return_margins=True)
    assert torch.equal(seq, seq_m) and torch.equal(lp, lp_m)
    assert margins.shape == (6,) and bool((margins > 0).all()) and bool(torch.isfinite(margins).all())
    # a perturbation of the logit bias far below every margin must leave every caption unchanged
    sd2 = {k: v.clone() for k, v in sd.items()}
    g = torch.Generator().manual_seed(0)
    sd2["logit.bias"] += (torch.rand(sd2["logit.bias"].shape, generator=g) - 0.5) * float(margins.min()) * 1e-2
    seq2, _, _ = O.sample_beam(sd2, "att2in2", fc, att, opt.seq_length, 3)
    assert torch.equal(seq, seq2)
    # diverse beam search reports margins too
    out = O.sample_beam(sd, "att2in2", fc, att, opt.seq_length, 4, group_size=2, return_margins=True)
    assert out[3].shape == (6,) and bool(torch.isfinite(out[3]).all())


def test_relative_greedy_margins():
    opt, cfg = synth.opt_for("tiny_topdown")
    sd = synth.init_state_dict(opt, seed=3)
    fc, att = synth.make_features(4, cfg["att_size"], opt.att_feat_size, seed=3)
    s1, lp1, m_abs = O.sample_greedy(sd, "topdown", fc, att, opt.seq_length, return_margins=True)
    s2, lp2, m_rel = O.sample_greedy(sd, "topdown", fc, att, opt.seq_length, return_margins=True, relative_margins=True)
    assert torch.equal(s1, s2)
    fin = torch.isfinite(m_abs)
    assert bool((m_rel[fin] <= m_abs[fin] / lp1[fin].abs().clamp_min(1e-9) + 1e-6).all())   # scale >= |best log-prob|
