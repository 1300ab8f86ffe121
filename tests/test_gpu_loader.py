"""FeatureCache / FeatureStream / decode_split on the device: the double-buffered stream hands over exactly the cached
batches, and a bf16 feature cache gives the captions of the fp32 tensors it was rounded from."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import unpaired_image_captioning_b200 as uic  # noqa: E402
from oracle import decoder_oracle as O  # noqa: E402
from oracle.compare import compare_beam  # noqa: E402
from unpaired_image_captioning_b200 import synth  # noqa: E402


def test_feature_stream_yields_the_cached_batches_in_order():
    fc, att = synth.make_features(23, 9, 64, seed=3)
    masks = synth.make_att_masks(23, 9, seed=3)
    cache = uic.FeatureCache(fc, att, masks, dtype=torch.bfloat16)
    seen = []
    for fc_d, att_d, m_d, (lo, hi) in uic.FeatureStream(cache, 5, "cuda", lo=2, hi=21):
        # the consumer keeps the GPU busy for a while: the next prefetch must not overwrite what it still reads
        big = torch.randn(2048, 2048, device="cuda")
        for _ in range(3):
            big = big @ big * 1e-3
        assert torch.equal(att_d.cpu(), att[lo:hi].to(torch.bfloat16)) and torch.equal(fc_d.cpu(), fc[lo:hi])
        assert torch.equal(m_d.cpu(), masks[lo:hi])
        seen.append((lo, hi))
    assert seen == [(2, 7), (7, 12), (12, 17), (17, 21)]
    it = iter(uic.FeatureStream(cache, 10, "cuda", loop=True))
    assert [next(it)[3] for _ in range(5)] == [(0, 10), (10, 20), (20, 23), (0, 10), (10, 20)]


@pytest.mark.parametrize("kind,L", [("att2in2", 49), ("topdown", 36)])
def test_bf16_cache_decodes_like_the_fp32_features(kind, L):
    opt = synth.make_opt(caption_model=kind, vocab_size=9999, rnn_size=512, input_encoding_size=512, att_hid_size=512, seq_length=16)
    sd = synth.init_state_dict(opt, seed=5)
    model = uic.setup(opt)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    n = 21
    fc, att = synth.make_features(n, L, 2048, seed=5)
    o = {"beam_size": 3}
    seq16, lp16, _ = uic.decode_split(model, uic.FeatureCache(fc, att, dtype=torch.bfloat16), 8, o)
    seq32, lp32, _ = uic.decode_split(model, uic.FeatureCache(fc, att, dtype=torch.float32), 8, o)
    assert torch.equal(seq16, seq32)                       # the engine rounds fp32 inputs to the same bf16 operand
    torch.testing.assert_close(lp16, lp32, rtol=0, atol=0)
    # and against the oracle on the fp32 features, at the north-star tolerance
    ref_seq, ref_lp, _, margins = O.sample_beam(sd, kind, fc, att, 16, 3, return_margins=True)
    exact, exempt, failures = compare_beam(seq16, ref_seq, margins, tol=1e-3)
    assert not failures and exact >= 1, (failures, exact, exempt)
    rows = (seq16 == ref_seq).all(1)
    torch.testing.assert_close(lp16[rows], ref_lp[rows], rtol=1e-3, atol=1e-2)
