"""Randomised shapes through the public model API against the oracle: odd region counts (1, 2, 3, 17, 33 ...), ragged
masks, batch sizes down to 1, one-step sequences, every beam width, layer widths from 32 to 512 (tests/shape_cases.py)."""
import pytest

pytestmark = pytest.mark.gpu

import shape_cases  # noqa: E402

CASES = shape_cases.cases(24, 7)


@pytest.mark.parametrize("case", range(len(CASES)))
def test_random_shape(case):
    msg, _ = shape_cases.run_case(case, CASES[case], strict_sampling=False)
    assert not msg, (CASES[case], msg)


@pytest.mark.parametrize("case", range(20))
def test_random_shape_optional_paths(case):
    """dropout training, scheduled sampling, self-critical step, use_bn, diverse beam search -- one of them per drawn shape."""
    c = shape_cases.cases(20, 12)[case]
    msg, _ = shape_cases.run_variant(case, c, shape_cases.VARIANTS[case % len(shape_cases.VARIANTS)])
    assert not msg, (c, msg)
