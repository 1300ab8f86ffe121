"""Import the UNMODIFIED reference hot path from /root/reference (build container only).

TEST INFRASTRUCTURE.  Used by `tests/golden/make_golden.py` to generate the committed fixtures and
by `tests/test_oracle_vs_reference.py`; it must never be reachable from `-m gpu` tests, `smoke()` or
`bench.py`, because /root/reference does not exist on the GPU box.

Three shims, no edits to the reference:
  * a stub `nltk` (misc/utils.py:5-6 imports it at module top, the hot path never uses it);
  * `builtins.reduce` (py2 builtin used at models/CaptionModel.py:176, models/AttModel.py:91);
  * on CPU, a no-op `Tensor.cuda` (models/CaptionModel.py:131,172 hard-code `.cuda()`).
"""
import builtins
import functools
import os
import sys
import types
import warnings

REFERENCE_ROOT = "/root/reference/pivot_based_eccv2018"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def load():
    """Returns (models, criterion) modules of the reference."""
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    import torch
    if "nltk" not in sys.modules:
        nltk = types.ModuleType("nltk")
        tr = types.ModuleType("nltk.translate")
        bs = types.ModuleType("nltk.translate.bleu_score")
        bs.SmoothingFunction = object
        tr.bleu_score = bs
        nltk.translate = tr
        sys.modules.update({"nltk": nltk, "nltk.translate": tr, "nltk.translate.bleu_score": bs})
    builtins.reduce = functools.reduce
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    warnings.filterwarnings("ignore", message=".*nn.functional.(tanh|sigmoid) is deprecated.*")
    import models  # noqa: E402  (the reference's package)
    import misc.criterion as criterion  # noqa: E402
    return models, criterion
