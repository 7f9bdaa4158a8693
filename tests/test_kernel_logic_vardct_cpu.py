"""CPU: the VarDCT host planner + the kernels' device functions compiled for the host (tests/emul), against the
oracle on streams written by the oracle's encoder. Bit-exact: the device code performs the oracle's float
operations (explicit fmaf where libjxl uses MulAdd) in an order-independent staging."""
import numpy as np
import pytest

import emul_lib
import jxlo
import vardct_cases as vc
from conftest import read_golden


@pytest.mark.parametrize("name", [c[0] for c in vc.SMALL_CASES])
def test_small_cases_match_oracle(name):
    data, shape = vc.encoded(name)
    got = emul_lib.decode([data], 3, jxlo.UINT8, [shape])[0]
    assert np.array_equal(got, jxlo.decode(data, 3, jxlo.UINT8))


@pytest.mark.parametrize("s", [1, 2, 3, 5, 9, 12, 13, 14, 17, 18, 20, 22, 24, 26])
def test_each_transform_matches_oracle_in_float(s):
    data, shape = vc.encoded("strategy_%d" % s)
    got = emul_lib.decode([data], 3, jxlo.FLOAT, [shape])[0]
    assert np.array_equal(got.view(np.uint32), jxlo.decode(data, 3, jxlo.FLOAT).view(np.uint32))


def test_mixed_batch_and_formats():
    a, sa = vc.encoded("heuristic")
    b, sb = vc.encoded("odd_size")
    m = read_golden("sample.jxl")
    got = emul_lib.decode([a, m, b], 4, jxlo.UINT16, [sa, (50, 40), sb])
    for g, d in zip(got, [a, m, b]):
        assert np.array_equal(g, jxlo.decode(d, 4, jxlo.UINT16))


def test_single_section_vardct_fails_loudly():
    img = vc.crop(200, 200)
    data = jxlo.encode_vardct(img)
    with pytest.raises(emul_lib.EmulError, match="single-section"):
        emul_lib.decode([data], 3, jxlo.UINT8, [(200, 200)])
