"""Error behaviour at the C ABI (SURVEY.md §8b "Errors"): bad arguments come back as negative return codes with a message
in uic_last_error(), never as a crash, a launch, or a sticky CUDA error; the Python wrapper turns them into exceptions."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from unpaired_image_captioning_b200 import _lib  # noqa: E402
from unpaired_image_captioning_b200._lib import ptr, stream  # noqa: E402

DEV = "cuda"
ERR_ARG, ERR_SHAPE, ERR_ALIGN = -1, -2, -3


def _expect(rc, *codes):
    lib = _lib.load()
    assert rc in codes, (rc, lib.uic_last_error())
    assert len(lib.uic_last_error()) > 0
    with pytest.raises(_lib.UicError):
        _lib.check(rc)


def test_bad_arguments_return_codes_and_leave_the_device_usable():
    lib = _lib.load()
    bf = torch.bfloat16
    a, b = torch.randn(64, 64, device=DEV).to(bf), torch.randn(32, 64, device=DEV).to(bf)
    c = torch.empty(64, 32, device=DEV)
    n0 = _lib.launch_count()
    # GEMM: null operand, K that is not a whole number of 16-byte rows, misaligned operand pointer
    _expect(lib.uic_gemm_bf16_ex(None, 64, ptr(b), 64, ptr(c), 32, None, 0, None, 64, 32, 64, 0, 0, 0.0, stream()), ERR_ARG)
    _expect(lib.uic_gemm_bf16_ex(ptr(a), 60, ptr(b), 60, ptr(c), 32, None, 0, None, 64, 32, 60, 0, 0, 0.0, stream()), ERR_ALIGN, ERR_SHAPE)
    _expect(lib.uic_gemm_bf16_ex(ptr(a) + 2, 64, ptr(b), 64, ptr(c), 32, None, 0, None, 63, 32, 64, 0, 0, 0.0, stream()), ERR_ALIGN)
    # sampling: k / beams / step out of range, missing previous tokens for the decoding constraint
    logits = torch.randn(4, 50, device=DEV)
    tv, ti = torch.empty(4, 20, device=DEV), torch.empty(4, 20, dtype=torch.int32, device=DEV)
    _expect(lib.uic_row_topk(ptr(logits), 50, None, ptr(tv), ptr(ti), 4, 50, 17, 0, stream()), ERR_SHAPE)
    _expect(lib.uic_row_topk(ptr(logits), 50, None, ptr(tv), ptr(ti), 4, 50, 3, _lib.SAMPLE_DECODING_CONSTRAINT, stream()), ERR_ARG)
    z32 = torch.zeros(4096, dtype=torch.int32, device=DEV)
    zf, zd, z64 = torch.zeros(4096, device=DEV), torch.zeros(512, dtype=torch.float64, device=DEV), torch.zeros(512, dtype=torch.int64, device=DEV)
    beam_args = (ptr(tv), ptr(ti), None, ptr(z32), ptr(zf), ptr(zf), ptr(z32), ptr(zf), ptr(zd), ptr(zf), ptr(z32), ptr(z32), ptr(z64))
    _expect(lib.uic_beam_step(*beam_args, 0, 8, 2, 17, 0, stream()), ERR_SHAPE)
    _expect(lib.uic_beam_step(*beam_args, 8, 8, 2, 2, 0, stream()), ERR_SHAPE)
    _expect(lib.uic_diverse_select(ptr(tv), ptr(ti), 2, ptr(z32), 1, 2, 3, 8, 0, 0.5, ptr(zf), ptr(zf), ptr(z32), stream()), ERR_SHAPE)
    # attention step: no output buffer; LSTM cell backward: a2c without its gradient buffer; statistics: pitch < cols
    _expect(lib.uic_att_step_fwd(ptr(zf), 32, ptr(a), ptr(a), ptr(zf), None, None, 0, None, 0, None, None, 0, 2, 1, 4, 32, 32, stream()), ERR_ARG)
    _expect(lib.uic_lstm_maxout_bwd(ptr(zf), 160, ptr(zf), 64, None, ptr(zf), ptr(zf), 32, None, 0, None, ptr(a), 160, None, 0, ptr(zf),
                                    2, 32, stream()), ERR_ARG)
    _expect(lib.uic_col_moments(ptr(a), 1, 32, None, 1, 64, 64, ptr(zd), ptr(zd), stream()), ERR_SHAPE)
    assert _lib.launch_count() == n0                      # nothing was launched
    torch.cuda.synchronize()                              # and nothing is pending or sticky
    _lib.gemm(a, b, out_f32=c)
    torch.testing.assert_close(c, a.float() @ b.float().t(), rtol=1e-2, atol=1e-2)


def test_python_wrappers_validate_like_torch_would():
    bf = torch.bfloat16
    a, b = torch.randn(16, 64, device=DEV).to(bf), torch.randn(8, 32, device=DEV).to(bf)
    with pytest.raises(ValueError):
        _lib.gemm(a, b, out_f32=torch.empty(16, 8, device=DEV))            # K mismatch
    with pytest.raises(ValueError):
        _lib.gemm(a.float(), b, out_f32=torch.empty(16, 8, device=DEV))    # dtype
    with pytest.raises(ValueError):
        _lib.dropout(torch.zeros(4, 4, 4, device=DEV), (0.5, torch.zeros(1, dtype=torch.int64, device=DEV)), 0)


def test_out_of_range_tokens_raise_like_the_reference():
    """The kernels clamp token ids; the module surface raises IndexError first, as nn.Embedding / gather do in the reference."""
    import unpaired_image_captioning_b200 as uic
    from unpaired_image_captioning_b200 import synth
    opt, cfg = synth.opt_for("tiny_att2in2")
    model = uic.setup(opt)
    model.load_state_dict(synth.init_state_dict(opt, seed=1))
    model = model.cuda().eval()
    fc, att = synth.make_features(3, cfg["att_size"], opt.att_feat_size, seed=1)
    labels, masks = synth.make_captions(3, opt.seq_length, opt.vocab_size, seed=1)
    bad = labels.clone()
    bad[1, 2] = opt.vocab_size + 1
    with pytest.raises(IndexError):
        model(fc.cuda(), None, att.cuda(), bad.cuda())
    with pytest.raises(IndexError):
        model(fc.cuda(), None, att.cuda(), bad.cuda(), masks.cuda(), None, mode="forward_loss")
    bad[1, 2] = -1
    with pytest.raises(IndexError):
        model(fc.cuda(), None, att.cuda(), bad.cuda())


def test_row_topk_rejects_more_than_16_candidates():
    from unpaired_image_captioning_b200 import _lib
    lib = _lib.load()
    logits = torch.randn(4, 100, device="cuda")
    val, idx = torch.empty(4, 20, device="cuda"), torch.empty(4, 20, device="cuda", dtype=torch.int32)
    rc = lib.uic_row_topk(logits.data_ptr(), 100, None, val.data_ptr(), idx.data_ptr(), 4, 100, 20, 0, torch.cuda.current_stream().cuda_stream)
    assert rc != 0 and b"16" in lib.uic_last_error()
