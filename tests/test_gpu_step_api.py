"""The single-step surface the reference exposes (SURVEY.md §8b "signatures to keep"): _prepare_feature, init_hidden,
get_logprobs_state (models/AttModel.py:158-165), core(...) (:430-446, :581-601, :636-654) and core.attention(...)
(:538-558), each against the oracle, with the tiles in the engine's own formats and in the reference's fp32 formats."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import unpaired_image_captioning_b200 as uic  # noqa: E402
from oracle import decoder_oracle as O  # noqa: E402
from unpaired_image_captioning_b200 import synth  # noqa: E402


def _setup(kind, B, L, masks, seed=51):
    opt = synth.make_opt(caption_model=kind, vocab_size=999, rnn_size=128, input_encoding_size=64, att_hid_size=96, seq_length=8,
                         fc_feat_size=256, att_feat_size=256)
    sd = synth.init_state_dict(opt, seed=seed)
    fc, att = synth.make_features(B, L, 256, seed=seed)
    am = synth.make_att_masks(B, L, seed=seed) if masks else None
    model = uic.setup(opt)
    model.load_state_dict(sd)
    return opt, sd, model.cuda().eval(), fc, att, am


def _close(a, b, tol=4e-3):
    a, b = a.float().cpu(), b.float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = float(((a - b).abs() / b.abs().clamp_min(1.0)).max())
    assert err < tol, err


@pytest.mark.parametrize("kind", ["att2in2", "att2all2", "topdown", "stackatt", "denseatt"])
@pytest.mark.parametrize("B,L,masks", [(5, 17, False), (1, 33, True), (9, 4, True)])
def test_get_logprobs_state_steps(kind, B, L, masks):
    opt, sd, model, fc, att, am = _setup(kind, B, L, masks)
    cu = lambda t: None if t is None else t.cuda()
    o_fc, o_att, o_patt, o_m = O.prepare_features(sd, kind, fc, att, am)
    p_fc, p_att, p_patt, p_m = model._prepare_feature(cu(fc), cu(att), cu(am))
    assert p_att.shape == o_att.shape and p_patt.shape == o_patt.shape
    _close(p_att, o_att, 1e-2)                                               # bf16 tile
    _close(p_patt, o_patt, 2e-2)                                             # decoded from the bf16 tile exp(2 p_att)
    state, o_state = model.init_hidden(B), O.init_hidden(sd, kind, B)
    assert state[0].shape == o_state[0].shape and state[1].shape == o_state[1].shape
    it = torch.zeros(B, dtype=torch.int64)
    for step in range(3):
        lp, state = model.get_logprobs_state(cu(it), p_fc, p_att, p_patt, p_m, state)
        o_lp, o_state = O.logprobs_state(sd, kind, it, o_fc, o_att, o_patt, o_m, o_state)
        _close(lp, o_lp)
        _close(state[0], o_state[0], 1e-2)                                   # h went through a bf16 operand slot
        _close(state[1], o_state[1])
        assert state[0].shape == o_state[0].shape
        it = o_lp.argmax(1)
        state = (o_state[0].cuda(), o_state[1].cuda())                       # restart from the oracle's state: no drift across steps


@pytest.mark.parametrize("kind", ["att2in2", "att2all2", "topdown"])
def test_core_and_attention_accept_reference_formats(kind):
    """core(xt, fc, att, p_att, state, masks) and core.attention(h, att, p_att, masks) with the reference's own fp32
    att / p_att tensors (the call pattern of models/AttModel.py:163,436,582)."""
    B, L = 6, 21
    opt, sd, model, fc, att, am = _setup(kind, B, L, True, seed=52)
    o_fc, o_att, o_patt, o_m = O.prepare_features(sd, kind, fc, att, am)
    g = torch.Generator().manual_seed(3)
    n_layers = O.num_layers(kind)
    state = (0.5 * torch.randn(n_layers, B, 128, generator=g), 0.5 * torch.randn(n_layers, B, 128, generator=g))
    xt = torch.randn(B, 64, generator=g).clamp_(min=0)
    out, new_state = model.core(xt.cuda(), o_fc.cuda(), o_att.cuda(), o_patt.cuda(), (state[0].cuda(), state[1].cuda()), o_m.cuda())
    o_out, o_new = O.CORES[kind](sd, xt, o_fc, o_att, o_patt, state, o_m)
    _close(out, o_out, 1e-2)
    _close(new_state[0], o_new[0], 1e-2)
    _close(new_state[1], o_new[1], 1e-2)
    h = 0.5 * torch.randn(B, 128, generator=g)
    ctx = model.core.attention(h.cuda(), o_att.cuda(), o_patt.cuda(), o_m.cuda())
    _close(ctx, O.attention(sd, h, o_att, o_patt, o_m), 1e-2)
    ctx = model.core.attention(h.cuda(), o_att.cuda(), o_patt.cuda())       # att_masks is optional
    _close(ctx, O.attention(sd, h, o_att, o_patt, None), 1e-2)
