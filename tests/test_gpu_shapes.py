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
