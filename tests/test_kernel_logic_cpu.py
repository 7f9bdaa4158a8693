"""CPU: host planner + the kernels' device functions compiled for the host, against the oracle and the goldens.
This checks the logic the GPU will execute; the GPU parity tests proper are in test_gpu_*.py."""
import numpy as np
import pytest

import emul_lib
import jxlo
from conftest import read_golden


@pytest.mark.parametrize("nch,dt", [(4, jxlo.UINT16), (4, jxlo.UINT8), (3, jxlo.UINT8), (1, jxlo.UINT16),
                                    (2, jxlo.FLOAT16), (4, jxlo.FLOAT)])
def test_sample_matches_oracle(nch, dt):
    data = read_golden("sample.jxl")
    got = emul_lib.decode([data], nch, dt, [(50, 40)])[0]
    want = jxlo.decode(data, nch, dt)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


def test_bench_matches_oracle_in_a_batch():
    a, b = read_golden("bench.jxl"), read_golden("sample.jxl")
    got = emul_lib.decode([a, b, a], 4, jxlo.UINT8, [(1433, 2122), (50, 40), (1433, 2122)])
    want = jxlo.decode(a, 4, jxlo.UINT8)
    assert np.array_equal(got[0], want)
    assert np.array_equal(got[2], want)
    assert np.array_equal(got[1], jxlo.decode(b, 4, jxlo.UINT8))


def test_unsupported_files_fail_loudly():
    for name in ["2bit.jxl"]:  # splines
        with pytest.raises(emul_lib.EmulError):
            emul_lib.decode([read_golden(name)], 3, jxlo.UINT8, [(600, 800)])


def test_wide_predictor_path_matches_oracle():
    # 0x100 in the endianness argument is the emulation's hook for the 64-bit predictor arithmetic
    a, b = read_golden("bench.jxl"), read_golden("sample.jxl")
    got = emul_lib.decode([a, b], 4, jxlo.UINT8, [(1433, 2122), (50, 40)], endianness=0x100)
    assert np.array_equal(got[0], jxlo.decode(a, 4, jxlo.UINT8))
    assert np.array_equal(got[1], jxlo.decode(b, 4, jxlo.UINT8))
