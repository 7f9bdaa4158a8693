"""Lossless Modular streams written by the oracle's plain encoder (tests/modular_cases.py): palette, delta palette,
squeeze, every predictor / property through fixed and random trees, prefix codes, LZ77.

CPU: the stream decodes back to the source samples with the oracle (lossless round trip: pins the oracle's encoder and
decoder against each other), and the product's host planner + the kernels' device functions compiled for the host give
the oracle's samples. The GPU parity test proper is test_gpu_decode.py::test_modular_cases_match_oracle."""
import numpy as np
import pytest

import emul_lib
import jxlo
import modular_cases as mc


@pytest.mark.parametrize("name", list(mc.CASES))
def test_oracle_round_trip_is_lossless(name):
    data, img = mc.encoded(name)
    _, kw, _ = mc.CASES[name]
    bits = kw.get("bits", 8)
    got = jxlo.decode(data, img.shape[2], jxlo.UINT8 if bits <= 8 else jxlo.UINT16)
    assert np.array_equal(got, img)


@pytest.mark.parametrize("name", list(mc.CASES))
def test_kernel_logic_matches_oracle(name):
    data, img = mc.encoded(name)
    _, _, (nc, dt) = mc.CASES[name]
    want = jxlo.decode(data, nc, dt)
    got = emul_lib.decode([data], nc, dt, [img.shape[:2]])[0]
    assert np.array_equal(got, want)


def test_all_cases_in_one_batch_kernel_logic():
    names = [n for n in mc.CASES if mc.CASES[n][2] == (3, jxlo.UINT8)]
    files = [mc.encoded(n)[0] for n in names]
    shapes = [mc.encoded(n)[1].shape[:2] for n in names]
    got = emul_lib.decode(files, 3, jxlo.UINT8, shapes)
    for n, f, g in zip(names, files, got):
        assert np.array_equal(g, jxlo.decode(f, 3, jxlo.UINT8)), n
