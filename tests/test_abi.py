"""CPU-only checks of the drop-in boundary: the shared library builds/loads, exports every symbol
include/uic_b200.h declares, and the ctypes table matches the header (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "uic_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uic_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from unpaired_image_captioning_b200 import build
    return build.build()


def test_header_declares_the_path():
    names = _declared()
    for required in ("uic_gemm_bf16", "uic_att_step_fwd", "uic_lstm_maxout_fwd", "uic_lstm_cell_fwd", "uic_lse_xent_fwd",
                     "uic_greedy_step", "uic_beam_step", "uic_last_error"):
        assert required in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in uic_b200.h but not exported"


def test_ctypes_table_matches_header(lib_path):
    from unpaired_image_captioning_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    # argument counts agree with the prototypes
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))


def test_missing_device_fails_loudly():
    import torch
    from unpaired_image_captioning_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("has a device")
    with pytest.raises(_lib.UicError):
        _lib.require_device()


def test_sass_uses_blackwell_tensor_and_tma_instructions(lib_path):
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass      # tcgen05.mma
    assert "UTMALDG" in sass      # TMA tile loads
    assert "LDTM" in sass         # tcgen05.ld (TMEM -> registers)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU or PyTorch fallback: without libuic_b200.so the binding raises instead of computing something else."""
    from unpaired_image_captioning_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libuic_b200.so"))
    with pytest.raises(_lib.UicError, match="no CPU fallback"):
        _lib.load()


def test_empty_batch_returns_empty_results():
    """Edge case of the reference interface: a batch of zero images yields empty outputs (no kernel is launched)."""
    import torch
    import unpaired_image_captioning_b200 as uic
    from unpaired_image_captioning_b200 import synth
    opt, cfg = synth.opt_for("tiny_att2in2")
    model = uic.setup(opt).eval()
    fc = torch.zeros(0, opt.fc_feat_size)
    att = torch.zeros(0, 7, opt.att_feat_size)
    seq, lp = model(fc, None, att, None, opt={"beam_size": 1}, mode="sample")
    assert seq.shape == (0, opt.seq_length) and lp.shape == (0, opt.seq_length) and seq.dtype == torch.long
    seq, lp = model(fc, None, att, None, opt={"beam_size": 3}, mode="sample")
    assert seq.shape == (0, opt.seq_length)
    out = model(fc, None, att, torch.zeros(0, opt.seq_length + 2, dtype=torch.long))
    assert out.shape == (0, opt.seq_length + 1, opt.vocab_size + 1)


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: the header must compile as C99 on its own (no C++ or torch types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    r = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-pedantic", HEADER], capture_output=True, text=True)
    assert r.returncode == 0 and not r.stderr.strip(), r.stderr
