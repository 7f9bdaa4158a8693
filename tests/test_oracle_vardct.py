"""CPU: the oracle's VarDCT path against what the reference pins:

* G3  samples/sample.jpg <-> samples/sample_jpg.jxl: the JXL file is the lossless transcode of the JPEG, so
      its quantised coefficients, quant tables, DC and CfL maps are all determined by the JPEG. The test
      re-derives libjxl's float pixel path from the JPEG's own coefficients and compares pixels.
* G4  samples/sample_grey.jxl: decodes (patches + Gaborish + EPF); variant and size as the reference asserts.
* G6  libjxl's self-contained known-answer methods restated: fast DCT vs an O(N^2) double oracle
      (lib/jxl/dct_test.cc:103-160, :205-282), DC <-> LLF consistency (lib/jxl/ac_strategy_test.cc:79-142).
"""
import ctypes
import os

import numpy as np
import pytest

import jpeg_coeffs
import jxlo
from conftest import read_golden

FP = ctypes.POINTER(ctypes.c_float)


def P(a):
    return a.ctypes.data_as(FP)


@pytest.fixture(scope="module")
def L():
    lib = jxlo.lib()
    lib.jxlo_library_quant_table.restype = ctypes.c_size_t
    lib.jxlo_library_quant_table.argtypes = [ctypes.c_int, FP, ctypes.c_size_t]
    lib.jxlo_llf_from_dc.argtypes = [ctypes.c_int, FP, ctypes.c_size_t, FP]
    lib.jxlo_fast_powf.restype = ctypes.c_float
    lib.jxlo_fast_powf.argtypes = [ctypes.c_float, ctypes.c_float]
    lib.jxlo_srgb_from_linear.restype = ctypes.c_float
    lib.jxlo_srgb_from_linear.argtypes = [ctypes.c_float]
    return lib


def slow_dct(x):
    """DCT-II down axis 0, scaled so that X_0 is the mean and the inverse has no scaling (lib/jxl/dct_for_test.h)."""
    n = x.shape[0]
    i = np.arange(n)
    m = np.cos(np.pi * i[:, None] * (i[None, :] + 0.5) / n) * np.where(i[:, None] > 0, np.sqrt(2), 1) / n
    return m @ x


DCT_SIZES = [(1, 1), (1, 2), (2, 1), (2, 2), (4, 2), (4, 4), (4, 8), (8, 4), (8, 8), (8, 16), (16, 8), (16, 16), (8, 32),
             (32, 8), (16, 32), (32, 16), (32, 32), (64, 32), (32, 64), (64, 64), (128, 64), (64, 128), (128, 128),
             (256, 128), (128, 256), (256, 256)]


@pytest.mark.parametrize("rows,cols", DCT_SIZES)
def test_scaled_dct_matches_slow_dct_and_inverts(L, rows, cols):
    rng = np.random.default_rng(rows * 1000 + cols)
    px = rng.standard_normal((rows, cols)).astype(np.float32)
    co = np.zeros(rows * cols, np.float32)
    L.jxlo_scaled_dct(rows, cols, P(px), P(co))
    d = slow_dct(slow_dct(px.astype(np.float64)).T).T  # [vertical freq][horizontal freq]
    want = d if rows < cols else d.T                   # coefficient layout: min(R, C) rows x max(R, C) columns
    got = co.reshape(min(rows, cols), max(rows, cols))
    assert np.abs(got - want).max() < 1e-6             # dct_test.cc: accuracy / N
    back = np.zeros((rows, cols), np.float32)
    L.jxlo_scaled_idct(rows, cols, P(co), P(back))
    assert np.abs(back - px).max() < 4e-5 * max(1, max(rows, cols) / 64)


PLAIN = {0: (1, 1), 4: (2, 2), 5: (4, 4), 6: (2, 1), 7: (1, 2), 8: (4, 1), 9: (1, 4), 10: (4, 2), 11: (2, 4), 18: (8, 8),
         19: (8, 4), 20: (4, 8), 21: (16, 16), 22: (16, 8), 23: (8, 16), 24: (32, 32), 25: (32, 16), 26: (16, 32)}


@pytest.mark.parametrize("strategy", sorted(PLAIN))
def test_llf_from_dc_reproduces_block_means(L, strategy):
    # ac_strategy_test.cc:79-142: the lowest frequencies derived from the 1:8 image, put through the inverse
    # transform with all other coefficients zero, give pixels whose 8x8 block means are the 1:8 image.
    by, bx = PLAIN[strategy]
    rng = np.random.default_rng(strategy)
    dc = rng.standard_normal((by, bx)).astype(np.float32)
    n = 64 * by * bx
    coeffs = np.zeros(n, np.float32)
    L.jxlo_llf_from_dc(strategy, P(dc), bx, P(coeffs))
    px = np.zeros((8 * by, 8 * bx), np.float32)
    L.jxlo_transform_to_pixels(strategy, P(coeffs), P(px))
    means = px.reshape(by, 8, bx, 8).mean(axis=(1, 3))
    assert np.abs(means - dc).max() < 2e-6 * max(by, bx)


@pytest.mark.parametrize("strategy", range(27))
def test_transform_to_pixels_is_linear_and_dc_preserving(L, strategy):
    cx = [1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32][strategy]
    cy = [1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16][strategy]
    n = 64 * cx * cy
    rng = np.random.default_rng(100 + strategy)
    a = rng.standard_normal(n).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    pa, pb, pab = (np.zeros((8 * cy, 8 * cx), np.float32) for _ in range(3))
    L.jxlo_transform_to_pixels(strategy, P(a), P(pa))
    L.jxlo_transform_to_pixels(strategy, P(b), P(pb))
    L.jxlo_transform_to_pixels(strategy, P(a + b), P(pab))
    assert np.abs(pab - (pa + pb)).max() < 2e-5 * np.abs(pab).max()
    # a block holding only its DC coefficient is flat at that value
    dc_only = np.zeros(n, np.float32)
    dc_only[0] = 0.75
    flat = np.zeros((8 * cy, 8 * cx), np.float32)
    L.jxlo_transform_to_pixels(strategy, P(dc_only), P(flat))
    assert np.abs(flat - 0.75).max() < 1e-5


def test_library_quant_tables_are_finite_and_sized(L):
    rx = [1, 1, 1, 1, 2, 4, 1, 1, 2, 1, 1, 8, 4, 16, 8, 32, 16]
    ry = [1, 1, 1, 1, 2, 4, 2, 4, 4, 1, 1, 8, 8, 16, 16, 32, 32]
    for t in range(17):
        n = 3 * 64 * rx[t] * ry[t]
        buf = np.zeros(n, np.float32)
        assert L.jxlo_library_quant_table(t, P(buf), n) == n
        assert np.all(np.isfinite(buf)) and np.all(buf[1:] > 0)
    # DCT8 luma: weight of the lowest AC band is the library seed 560 (quant_weights.cc:529-556)
    buf = np.zeros(192, np.float32)
    L.jxlo_library_quant_table(0, P(buf), 192)
    assert abs(1.0 / buf[64 + 1] - 560.0) / 560.0 < 0.2


def test_fast_math_accuracy(L):
    # lib/jxl/base/fast_math-inl.h: FastPowf max relative error ~3e-5; TF_SRGB error ~5e-7
    for b, e in [(2.0, 0.5), (0.3, 2.2), (10.0, -1.5), (0.02, 0.45455)]:
        assert abs(L.jxlo_fast_powf(b, e) / (b ** e) - 1) < 1e-4
    for v in [0.0, 0.001, 0.0031308, 0.01, 0.2, 0.5, 1.0]:
        want = 12.92 * v if v <= 0.0031308 else 1.055 * v ** (1 / 2.4) - 0.055
        assert abs(L.jxlo_srgb_from_linear(v) - want) < 2e-6
    assert L.jxlo_srgb_from_linear(-0.5) == -L.jxlo_srgb_from_linear(0.5)


def _float_cfl_reconstruction(frame, ytox, ytob):
    """libjxl's pixel path for a transcoded JPEG, computed from the JPEG's own coefficients: luma is
    q * Q; chroma is stored as c - cfl_int(y) (lib/jxl/dec_group.cc:381-400 is the integer inverse) and
    dequantised as stored * Q_c + (factor / 84) * y * Q_y (lib/jxl/dec_group.cc:139-160)."""
    comps, qt = frame["comps"], frame["qt"]
    n = np.arange(8)
    basis = np.cos(np.pi * n[None, :] * (n[:, None] + 0.5) / 8) * np.where(n[None, :] > 0, 1.0, np.sqrt(0.5)) * 0.5

    def recon(cf):
        by, bx, _ = cf.shape
        out = np.zeros((by * 8, bx * 8))
        for y in range(by):
            for x in range(bx):
                out[y * 8:y * 8 + 8, x * 8:x * 8 + 8] = basis @ cf[y, x].reshape(8, 8) @ basis.T
        return out

    qy = qt[comps[0]["tq"]].astype(np.int64)
    yq = comps[0]["coeffs"].astype(np.int64)
    planes = [recon(yq * qy) + 128]
    for ci, f in ((1, ytox), (2, ytob)):
        qc = qt[comps[ci]["tq"]].astype(np.int64)
        c = comps[ci]["coeffs"].astype(np.int64)
        ratio = int(f * 2048 / 84)  # ColorCorrelation::RatioJPEG
        coeff_scale = (((2048 * qy) // qc) * ratio + 1024) >> 11
        stored = c - ((yq * coeff_scale + 1024) >> 11)
        deq = stored * qc + float(np.float32(f) / np.float32(84)) * (yq * qy)
        deq[:, :, 0] = c[:, :, 0] * qc[0]
        planes.append(recon(deq))
    yy, cb, cr = planes
    r = yy + 1.402 * cr
    g = yy - (0.114 * 1.772 / 0.587) * cb - (0.299 * 1.402 / 0.587) * cr
    b = yy + 1.772 * cb
    return np.stack([r, g, b], axis=2)[:frame["height"], :frame["width"]]


def test_g3_sample_jpg_jxl_pixels_follow_from_the_jpeg_coefficients(monkeypatch):
    frame = jpeg_coeffs.parse(read_golden("sample.jpg"))
    data = read_golden("sample_jpg.jxl")
    # quantisation-bias adjustment off: the JPEG relation is exact only for unbiased dequantisation
    monkeypatch.setenv("JXLO_NO_QUANT_BIAS", "1")
    got = jxlo.decode(data, 3, jxlo.FLOAT).astype(np.float64) * 255
    # the luma plane does not depend on the chroma-from-luma factors: exact JPEG luma
    w = np.array([0.299, 0.587, 0.114])
    want_plain = jpeg_coeffs.reconstruct_rgb_float(frame)
    assert np.abs(got @ w - want_plain @ w).max() < 0.01
    # chroma: the single 64x64 tile of this 40x50 image carries the factors (-15, 47)
    want = _float_cfl_reconstruction(frame, -15, 47)
    assert np.abs(got - want).max() < 0.01
    monkeypatch.delenv("JXLO_NO_QUANT_BIAS")
    biased = jxlo.decode(data, 3, jxlo.UINT8)
    assert biased.shape == (50, 40, 3)
    assert np.abs(biased.astype(float) - np.clip(np.round(want), 0, 255)).max() <= 6  # bias moves samples slightly


def test_g4_sample_grey_jxl_decodes_as_the_reference_asserts():
    # jpegxl-rs/src/tests/decode.rs:82-93: Pixels::Uint16, len == width * height
    d = jxlo.Decoded(read_golden("sample_grey.jxl"))
    assert (d.info.xsize, d.info.ysize, d.info.bits, d.info.num_color, d.info.xyb) == (40, 50, 16, 1, 1)
    assert d.frame_info()[0].startswith("modular type=2 6x6") and "vardct" in d.frame_info()[1]
    px = d.pixels(1, jxlo.UINT16)
    assert px.dtype == np.uint16 and px.size == 40 * 50
    # the picture is the grey JPEG XL logo on a light background: bright corner, dark glyph
    assert px[0, 0, 0] > 60000 and px.min() < 45000
