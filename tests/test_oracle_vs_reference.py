"""Oracle restatement vs the live, unmodified reference at larger shapes and more seeds.
Runs only where /root/reference exists (the build container); skipped on the GPU box."""
import pytest
import torch

from oracle import decoder_oracle as O
from oracle import reference_shim
from unpaired_image_captioning_b200 import synth

pytestmark = pytest.mark.skipif(not reference_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return reference_shim.load()


@pytest.mark.parametrize("kind,over", [("att2in2", dict(use_bn=2)), ("topdown", dict(logit_layers=2)), ("denseatt", dict(use_bn=2, logit_layers=3)),
                                       ("stackatt", dict(logit_layers=2))])
def test_second_batchnorm_and_hidden_logit_layers(ref, kind, over):
    """use_bn = 2 (models/AttModel.py:84) and logit_layers > 1 (:89-91), eval mode: log-probs, greedy, beam."""
    models, _ = ref
    opt = synth.make_opt(caption_model=kind, vocab_size=299, rnn_size=64, input_encoding_size=48, att_hid_size=40, seq_length=9,
                         fc_feat_size=96, att_feat_size=96, **over)
    sd = synth.init_state_dict(opt, seed=7, peaked=20.0, eos_bias=0.5)
    model = models.setup(opt)
    model.load_state_dict(sd, strict=True)
    model.eval()
    B, L = 5, 11
    fc, att = synth.make_features(B, L, 96, seed=7)
    labels, _ = synth.make_captions(B, 9, 299, seed=7, min_len=3)
    am = synth.make_att_masks(B, L, seed=7) if opt.use_bn else None
    with torch.no_grad():
        torch.testing.assert_close(O.teacher_forced(sd, kind, fc, att, labels, am), model(fc, None, att, labels, am), rtol=1e-5, atol=2e-6)
        for o in ({"beam_size": 1}, {"beam_size": 3}):
            rs, rlp = model(fc, None, att, am, opt=dict(o), mode="sample")
            s, lp = O.sample(sd, kind, fc, att, 9, am, dict(o))
            assert torch.equal(s, rs), o
            torch.testing.assert_close(lp, rlp, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("kind", ["att2in2", "att2all2", "topdown", "stackatt", "denseatt"])
@pytest.mark.parametrize("seed,use_masks", [(11, False), (12, True)])
def test_midsize_forward_loss_sampling(ref, kind, seed, use_masks):
    models, criterion = ref
    opt = synth.make_opt(caption_model=kind, vocab_size=299, rnn_size=64, input_encoding_size=48,
                         att_hid_size=40, seq_length=9, fc_feat_size=96, att_feat_size=96)
    sd = synth.init_state_dict(opt, seed=seed, peaked=30.0, eos_bias=0.5)
    model = models.setup(opt)
    model.load_state_dict(sd)
    model.eval()
    B, L = 6, 11
    fc, att = synth.make_features(B, L, 96, seed=seed)
    labels, masks = synth.make_captions(B, 9, 299, seed=seed, min_len=3)
    am = synth.make_att_masks(B, L, seed=seed) if use_masks else None

    ref_out = model(fc, None, att, labels, am)
    out = O.teacher_forced(sd, kind, fc, att, labels, am)
    torch.testing.assert_close(out, ref_out.detach(), rtol=1e-5, atol=2e-6)
    ref_loss = criterion.LanguageModelCriterion(opt)(ref_out, labels[:, 1:], masks[:, 1:])
    torch.testing.assert_close(O.xe_loss(out, labels[:, 1:], masks[:, 1:]), ref_loss.detach(), rtol=1e-5, atol=1e-6)

    with torch.no_grad():
        for o in ({"beam_size": 1}, {"beam_size": 3}, {"beam_size": 4, "max_ppl": 1},
                  {"beam_size": 3, "decoding_constraint": 1}):
            rs, rlp = model(fc, None, att, am, opt=dict(o), mode="sample")
            s, lp = O.sample(sd, kind, fc, att, 9, am, dict(o))
            assert torch.equal(s, rs), o
            torch.testing.assert_close(lp, rlp, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("kind", ["att2in2", "topdown"])
@pytest.mark.parametrize("o", [{"beam_size": 6, "group_size": 3, "diversity_lambda": 0.5},
                               {"beam_size": 4, "group_size": 2, "diversity_lambda": 2.0, "decoding_constraint": 1},
                               {"beam_size": 6, "group_size": 2, "diversity_lambda": 0.7, "max_ppl": 1}])
def test_diverse_beam_search(ref, kind, o):
    """group_size > 1 (models/CaptionModel.py:36-45,124-172): sequences, log-probs and the per-group done lists."""
    models, _ = ref
    opt = synth.make_opt(caption_model=kind, vocab_size=299, rnn_size=64, input_encoding_size=48,
                         att_hid_size=40, seq_length=9, fc_feat_size=96, att_feat_size=96)
    sd = synth.init_state_dict(opt, seed=21, peaked=30.0, eos_bias=0.5)
    model = models.setup(opt)
    model.load_state_dict(sd)
    model.eval()
    fc, att = synth.make_features(5, 11, 96, seed=21)
    with torch.no_grad():
        rs, rlp = model(fc, None, att, None, opt=dict(o), mode="sample")
    s, lp, done = O.sample_beam(sd, kind, fc, att, 9, o["beam_size"], None, o.get("decoding_constraint", 0), o.get("max_ppl", 0),
                                o["group_size"], o["diversity_lambda"])
    assert torch.equal(s, rs), o
    torch.testing.assert_close(lp, rlp, rtol=1e-5, atol=2e-6)
    for k in range(5):
        ref_done = model.done_beams[k]
        assert len(ref_done) == len(done[k])
        for a, b in zip(ref_done, done[k]):
            assert torch.equal(a["seq"], b["seq"])
            assert abs(a["p"] - b["p"]) < 1e-4 * max(1.0, abs(a["p"]))


@pytest.mark.parametrize("kind", ["att2in2", "topdown"])
def test_use_bn_batchnorm_in_att_embed(ref, kind):
    """use_bn = 1 (opts.py:52 default; models/AttModel.py:79-84): BatchNorm1d over the packed valid regions.  eval(): running
    statistics; train(): batch statistics (+ gradients of the BN affine and the Linear behind it); without att_masks the
    reference raises."""
    models, criterion = ref
    opt = synth.make_opt(caption_model=kind, vocab_size=299, rnn_size=64, input_encoding_size=48, att_hid_size=40, seq_length=9,
                         fc_feat_size=96, att_feat_size=96, use_bn=1)
    sd = synth.init_state_dict(opt, seed=31, peaked=30.0, eos_bias=0.5)
    model = models.setup(opt)
    model.load_state_dict(sd)
    B, L = 6, 11
    fc, att = synth.make_features(B, L, 96, seed=31)
    labels, masks = synth.make_captions(B, 9, 299, seed=31, min_len=3)
    am = synth.make_att_masks(B, L, seed=31)
    crit = criterion.LanguageModelCriterion(opt)

    model.eval()
    ref_out = model(fc, None, att, labels, am)
    torch.testing.assert_close(O.teacher_forced(sd, kind, fc, att, labels, am), ref_out.detach(), rtol=1e-5, atol=3e-6)
    with torch.no_grad():
        rs, rlp = model(fc, None, att, am, opt={"beam_size": 3}, mode="sample")
    s, lp = O.sample(sd, kind, fc, att, 9, am, {"beam_size": 3})
    assert torch.equal(s, rs)
    with pytest.raises(RuntimeError):
        model(fc, None, att, labels, None)
    with pytest.raises(RuntimeError):
        O.teacher_forced(sd, kind, fc, att, labels, None)

    model.load_state_dict(sd)          # the failed call above already touched num_batches_tracked
    model.train()
    model.zero_grad()
    loss = crit(model(fc, None, att, labels, am), labels[:, 1:], masks[:, 1:])
    loss.backward()
    with O.bn_training():
        o_loss, o_grads = O.loss_and_grads(sd, kind, fc, att, labels, masks, am)
        mean, var, n = O.bn_batch_stats(att, am)
    torch.testing.assert_close(o_loss, loss.detach(), rtol=1e-5, atol=1e-6)
    for k, p in model.named_parameters():
        torch.testing.assert_close(o_grads[k], p.grad, rtol=2e-4, atol=2e-6, msg=k)
    bn = model.att_embed[0]
    torch.testing.assert_close(bn.running_mean, 0.9 * sd["att_embed.0.running_mean"] + 0.1 * mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bn.running_var, 0.9 * sd["att_embed.0.running_var"] + 0.1 * var * n / (n - 1), rtol=1e-5, atol=1e-6)
