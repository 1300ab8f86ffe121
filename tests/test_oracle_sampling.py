"""CPU checks of the sampling / self-critical restatement in the oracle: the counter-based Gumbel noise is what it
claims to be, Gumbel-max over it reproduces multinomial sampling (the reference's torch.multinomial,
models/AttModel.py:231-239), and the RewardCriterion mirror equals the reference formula."""
import numpy as np
import torch

import unpaired_image_captioning_b200 as uic
from oracle import decoder_oracle as O
from unpaired_image_captioning_b200 import synth


def test_gumbel_noise_is_deterministic_and_standard():
    a = O.gumbel_noise(1234, 3, 64, 4096)
    b = O.gumbel_noise(1234, 3, 64, 4096)
    assert torch.equal(a, b)
    assert not torch.equal(a, O.gumbel_noise(1234, 4, 64, 4096)) and not torch.equal(a, O.gumbel_noise(1235, 3, 64, 4096))
    assert abs(float(a.mean()) - 0.5772) < 0.02                      # Euler-Mascheroni
    assert abs(float(a.var()) - np.pi ** 2 / 6) < 0.05
    assert bool(torch.isfinite(a).all())
    # rows and columns are decorrelated
    assert abs(float(torch.corrcoef(torch.stack([a[0], a[1]]))[0, 1])) < 0.06
    assert abs(float(torch.corrcoef(torch.stack([a[:, 0], a[:, 1]]))[0, 1])) < 0.4


def test_gumbel_max_reproduces_the_multinomial_distribution():
    torch.manual_seed(0)
    V, n, temperature = 12, 60000, 0.7
    lp = torch.log_softmax(torch.randn(V) * 2.0, 0)
    keys = lp[None, :] / temperature + O.gumbel_noise(99, 0, n, V)
    counts = torch.bincount(keys.argmax(1), minlength=V).double()
    p = torch.softmax(lp / temperature, 0).double()                 # what torch.multinomial(exp(lp / T)) draws from
    chi2 = float(((counts - n * p) ** 2 / (n * p)).sum())
    assert chi2 < 40.0, chi2                                        # 11 degrees of freedom: P(chi2 > 40) ~ 4e-5


def test_sample_multinomial_oracle_semantics():
    opt, cfg = synth.opt_for("tiny_att2in2")
    sd = synth.init_state_dict(opt, seed=5)
    fc, att = synth.make_features(6, 7, opt.att_feat_size, seed=5)
    T = opt.seq_length
    seq, lp = O.sample_multinomial(sd, "att2in2", fc, att, T, temperature=1.0, seed=7)
    seq2, lp2 = O.sample_multinomial(sd, "att2in2", fc, att, T, temperature=1.0, seed=7)
    assert torch.equal(seq, seq2) and torch.equal(lp, lp2)
    seq3, _ = O.sample_multinomial(sd, "att2in2", fc, att, T, temperature=1.0, seed=8)
    assert not torch.equal(seq, seq3)
    # zeros after the first end token; log-probs are those of the sampled tokens
    for r in range(seq.size(0)):
        z = (seq[r] == 0).nonzero()
        if z.numel():
            assert int(seq[r, int(z[0]):].abs().sum()) == 0
    assert bool((lp <= 0).all())
    # a very low temperature degenerates to greedy decoding
    g_seq, _ = O.sample_greedy(sd, "att2in2", fc, att, T)
    c_seq, _ = O.sample_multinomial(sd, "att2in2", fc, att, T, temperature=1e-4, seed=3)
    assert torch.equal(g_seq, c_seq)


def test_reward_criterion_matches_the_reference_formula():
    torch.manual_seed(1)
    B, T = 7, 9
    lp = -torch.rand(B, T)
    seq = torch.randint(0, 5, (B, T))
    seq[:, 0] = torch.randint(1, 5, (B,))
    reward = torch.randn(B, 1).expand(B, T).contiguous()
    got = uic.RewardCriterion()(lp, seq, reward)
    mask = torch.cat([torch.ones(B, 1), (seq > 0).float()[:, :-1]], 1)      # misc/criterion.py:118-119
    want = (-lp * reward * mask).sum() / mask.sum()
    assert abs(float(got) - float(want)) < 1e-6
    assert abs(float(O.reward_loss(lp, seq, reward)) - float(want)) < 1e-6


def test_scheduled_sampling_oracle_semantics():
    """ss_prob = 0 is plain teacher forcing; ss_prob = 1 replaces every input from step 1 on; the coin is uniform."""
    opt, cfg = synth.opt_for("tiny_att2in2")
    sd = synth.init_state_dict(opt, seed=5)
    fc, att = synth.make_features(16, 7, opt.att_feat_size, seed=5)
    labels, masks = synth.make_captions(16, opt.seq_length, opt.vocab_size, seed=5)
    plain = O.teacher_forced(sd, "att2in2", fc, att, labels)
    out0, used0, _ = O.teacher_forced(sd, "att2in2", fc, att, labels, ss_prob=0.0, ss_seed=3, return_tokens=True)
    assert torch.equal(plain, out0) and torch.equal(used0, labels[:, :used0.size(1)])
    out1, used1, margins = O.teacher_forced(sd, "att2in2", fc, att, labels, ss_prob=1.0, ss_seed=3, return_tokens=True)
    assert torch.equal(used1[:, 0], labels[:, 0]) and bool(torch.isfinite(margins[:, 1:]).all())
    assert not torch.equal(used1[:, 1:], labels[:, 1:used1.size(1)])
    out_h, used_h, m_h = O.teacher_forced(sd, "att2in2", fc, att, labels, ss_prob=0.5, ss_seed=3, return_tokens=True)
    frac = float(torch.isfinite(m_h[:, 1:]).float().mean())
    assert 0.3 < frac < 0.7, frac
    out_h2, used_h2, _ = O.teacher_forced(sd, "att2in2", fc, att, labels, ss_prob=0.5, ss_seed=3, return_tokens=True)
    assert torch.equal(used_h, used_h2) and torch.equal(out_h, out_h2)
    u = torch.cat([O.uniform_noise(11, t, 4096) for t in range(4)])
    assert abs(float(u.mean()) - 0.5) < 0.02 and float(u.min()) > 0.0 and float(u.max()) < 1.0


def test_dropout_mask_oracle():
    m = O.dropout_mask((0.3, 5), O.DROP_ATT, range(2000), 128)
    assert abs(float((m > 0).float().mean()) - 0.7) < 0.01
    assert abs(float(m.max()) - 1.0 / 0.7) < 1e-6 and float(m.min()) == 0.0
    assert torch.equal(m, O.dropout_mask((0.3, 5), O.DROP_ATT, range(2000), 128))
    assert not torch.equal(m, O.dropout_mask((0.3, 5), O.DROP_OUT, range(2000), 128))
    assert not torch.equal(m, O.dropout_mask((0.3, 6), O.DROP_ATT, range(2000), 128))
    # rows are addressed by id: a slice of the ids gives the slice of the mask
    assert torch.equal(m[100:200], O.dropout_mask((0.3, 5), O.DROP_ATT, range(100, 200), 128))
    opt, cfg = synth.opt_for("tiny_topdown")
    sd = synth.init_state_dict(opt, seed=5)
    fc, att = synth.make_features(4, 5, opt.att_feat_size, seed=5)
    labels, masks = synth.make_captions(4, opt.seq_length, opt.vocab_size, seed=5)
    plain = O.teacher_forced(sd, "topdown", fc, att, labels)
    dropped = O.teacher_forced(sd, "topdown", fc, att, labels, drop=(0.5, 9))
    assert not torch.allclose(plain, dropped)
    assert torch.equal(dropped, O.teacher_forced(sd, "topdown", fc, att, labels, drop=(0.5, 9)))
