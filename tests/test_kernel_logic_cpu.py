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
    # what the path does not cover is refused with an error, never decoded wrongly: here a lossy stream cut in the middle of
    # its sections (the refusals by feature name -- chroma subsampling, noise, blending ... -- need streams no encoder
    # in this repo writes; the planner's JXLB_CHECK messages name them)
    import vardct_cases as vc
    data = vc.encoded("dct8_plain")[0]
    with pytest.raises(emul_lib.EmulError):
        emul_lib.decode([data[:len(data) // 3]], 3, jxlo.UINT8, [(300, 400)])


def test_g5_2bit_jxl_splines_through_the_kernel_logic():
    # the reference's fifth fixture (jpegxl-rs/src/tests/decode.rs:69-80): a 2-bit RGB Modular frame whose drawing is
    # all splines; every output type against the oracle
    d = read_golden("2bit.jxl")
    for nc, dt in [(3, jxlo.UINT8), (4, jxlo.UINT16), (3, jxlo.FLOAT)]:
        got = emul_lib.decode([d], nc, dt, [(600, 800)])[0]
        assert np.array_equal(got.view(np.uint8), jxlo.decode(d, nc, dt).view(np.uint8))


def test_wide_predictor_path_matches_oracle():
    # 0x100 in the endianness argument is the emulation's hook for the 64-bit predictor arithmetic
    a, b = read_golden("bench.jxl"), read_golden("sample.jxl")
    got = emul_lib.decode([a, b], 4, jxlo.UINT8, [(1433, 2122), (50, 40)], endianness=0x100)
    assert np.array_equal(got[0], jxlo.decode(a, 4, jxlo.UINT8))
    assert np.array_equal(got[1], jxlo.decode(b, 4, jxlo.UINT8))


def test_corrupted_and_truncated_inputs_fail_with_an_error():
    """Seeded bit flips and truncations of every fixture: the planner or the stream status words report an error
    (or, rarely, the damage is harmless) -- never a crash, a hang or an out-of-bounds write (the arenas of the
    emulation are plain vectors, so a wild index shows up as a crash of this process)."""
    import numpy as np
    import vardct_cases as vc
    files = {"sample.jxl": (read_golden("sample.jxl"), (50, 40)), "sample_jpg.jxl": (read_golden("sample_jpg.jxl"), (50, 40)),
             "sample_grey.jxl": (read_golden("sample_grey.jxl"), (50, 40)), "lossy": vc.encoded("heuristic"),
             "passes": vc.encoded("three_passes")}
    # round 2: the chained alpha streams (positions handed over by the AC decode), the probe round of a small alpha
    # frame, the lossless encoder's streams and the spline decoder; output with 4 channels so that alpha is read
    img = vc.crop(300, 520, 100, 200)
    rgba = np.dstack([img, img[:, :, 0] ^ img[:, :, 2]])
    files["alpha"] = (jxlo.encode_vardct(rgba, strategy_mode=2), (300, 520))
    files["alpha_small"] = (jxlo.encode_vardct(np.ascontiguousarray(rgba[:60, :70]), strategy_mode=2), (60, 70))
    files["lossless"] = (emul_lib.encode_lossless(np.ascontiguousarray(rgba[:150, :300])), (150, 300))
    files["2bit.jxl"] = (read_golden("2bit.jxl"), (600, 800))
    files["lossy_splines"] = (jxlo.encode_vardct(img, strategy_mode=2, splines=6), (300, 520))
    rng = np.random.default_rng(7)
    errors = 0
    for name, (data, shape) in files.items():
        for trial in range(24):
            b = bytearray(data)
            if trial % 4 == 0:
                b = b[:int(len(b) * (0.2 + 0.03 * trial))]
            else:
                for _ in range(int(rng.integers(1, 4))):
                    b[int(rng.integers(2, len(b)))] ^= 1 << int(rng.integers(8))
            try:
                emul_lib.decode([bytes(b)], 4 if name.startswith(("alpha", "lossless")) else 3, jxlo.UINT8, [shape])
            except emul_lib.EmulError:
                errors += 1
    assert errors > 0.75 * 24 * len(files)  # almost every damaged file is detected (ANS final state, bounds, header checks)
