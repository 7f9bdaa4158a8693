"""GPU: VarDCT decode through the C ABI against the oracle (lossy parity is pinned on the oracle,
the oracle on the reference's goldens and known-answer tests). The bar is bit-exact output samples."""
import numpy as np
import pytest

import jxlo
import vardct_cases as vc
from conftest import read_golden

pytestmark = pytest.mark.gpu


def test_all_cases_in_one_batch(pkg):
    names = [c[0] for c in vc.SMALL_CASES + vc.STRATEGY_CASES]
    files = [vc.encoded(n)[0] for n in names]
    outs = pkg.decode_batch(files, 3, np.uint8)
    bad = [n for n, f, o in zip(names, files, outs) if not np.array_equal(o, jxlo.decode(f, 3, jxlo.UINT8))]
    assert not bad, bad


@pytest.mark.parametrize("dt,npdt", [(jxlo.UINT16, np.uint16), (jxlo.FLOAT16, np.float16), (jxlo.FLOAT, np.float32)])
def test_pixel_types(pkg, dt, npdt):
    names = ["all_strategies", "three_passes", "strategy_24"]
    files = [vc.encoded(n)[0] for n in names]
    outs = pkg.decode_batch(files, 4, npdt)
    for f, o in zip(files, outs):
        want = jxlo.decode(f, 4, dt)
        assert np.array_equal(o.view(np.uint8), want.view(np.uint8))


def test_event_api_decodes_a_lossy_file(pkg):
    data, shape = vc.encoded("heuristic")
    dec = pkg.decoder_builder().build()
    meta, px = dec.decode(data)
    assert px.variant == "Uint8" and (meta.height, meta.width) == shape and not meta.has_alpha_channel
    assert np.array_equal(px.data.reshape(shape + (3,)), jxlo.decode(data, 3, jxlo.UINT8))


def test_mixed_modular_and_vardct_batch(pkg):
    a, _ = vc.encoded("heuristic")
    m = read_golden("bench.jxl")
    outs = pkg.decode_batch([a, m, a], 4, np.uint8)
    assert np.array_equal(outs[0], jxlo.decode(a, 4, jxlo.UINT8))
    assert np.array_equal(outs[1], jxlo.decode(m, 4, jxlo.UINT8))
    assert np.array_equal(outs[2], outs[0])


def test_4k_frame_full_size(pkg):
    # BASELINE.json configs[1]: a 3840x2160 lossy frame, compared sample by sample
    img = vc.frame_4k()
    data = jxlo.encode_vardct(img, distance=1.0, strategy_mode=2)
    got = pkg.decode_batch([data, data], 3, np.uint8)
    want = jxlo.decode(data, 3, jxlo.UINT8)
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)
    err = got[0].astype(np.float64) - img
    assert 10 * np.log10(255 ** 2 / (err ** 2).mean()) > 30  # and it is the picture that was encoded


def test_token_budget_retry(pkg):
    # three passes over noisy content produce more tokens than the byte-count heuristic allows: Wait() regrows
    data, shape = vc.encoded("three_passes")
    out = pkg.decode_batch([data], 3, np.uint8)[0]
    assert np.array_equal(out, jxlo.decode(data, 3, jxlo.UINT8))


def test_single_section_frames(pkg):
    # frames of one group: the sub-streams are chained inside one section, positions come from probe launches
    files, shapes = [], []
    for h, w, mode, seed in [(80, 100, 1, 3), (256, 256, 2, 4), (17, 9, 1, 5), (200, 256, 0, 6)]:
        files.append(jxlo.encode_vardct(vc.crop(h, w, 300, 500), strategy_mode=mode, random_side_info=True, epf_iters=3, seed=seed))
    big, _ = vc.encoded("heuristic")
    outs = pkg.decode_batch(files + [big], 3, np.uint8)
    for f, o in zip(files + [big], outs):
        assert np.array_equal(o, jxlo.decode(f, 3, jxlo.UINT8))


def test_g3_sample_jpg_jxl(pkg):
    """The reference's own VarDCT fixture, written by libjxl (lossless transcode of samples/sample.jpg: container,
    YCbCr, DCT8, raw quantisation tables, single section): the GPU path equals the oracle bit for bit, and the pixels
    follow from the JPEG's coefficients (jpegxl-rs/src/tests/encode.rs:54-72 asserts the coefficient identity)."""
    import jpeg_coeffs
    from test_oracle_vardct import _float_cfl_reconstruction
    jpg = read_golden("sample_jpg.jxl")
    got8 = pkg.decode_batch([jpg, jpg], 3, np.uint8)
    want8 = jxlo.decode(jpg, 3, jxlo.UINT8)
    assert np.array_equal(got8[0], want8) and np.array_equal(got8[1], want8)
    gotf = pkg.decode_batch([jpg], 3, np.float32)[0]
    assert np.array_equal(gotf.view(np.uint32), jxlo.decode(jpg, 3, jxlo.FLOAT).view(np.uint32))
    frame = jpeg_coeffs.parse(read_golden("sample.jpg"))
    want = _float_cfl_reconstruction(frame, -15, 47)  # dequantised JPEG coefficients -> IDCT -> YCbCr -> RGB, in double
    # the decoder's quantisation-bias adjustment (lib/jxl/quantizer-inl.h:34-71) moves samples by a few levels
    assert np.abs(got8[0].astype(float) - np.clip(np.round(want), 0, 255)).max() <= 6
    dec = pkg.decoder_builder().build()
    meta, px = dec.decode(jpg)
    assert (meta.width, meta.height) == (40, 50) and px.variant == "Uint8"
    assert np.array_equal(np.asarray(px.data).reshape(50, 40, 3), want8)


def test_g4_sample_grey_jxl(pkg):
    """The reference's grey fixture (libjxl-written): reference-only XYB Modular frame + single-section VarDCT frame with
    patches, Gaborish, EPF. jpegxl-rs/src/tests/decode.rs:82-93 asserts Pixels::Uint16 and len == width * height; here
    additionally bit-exact against the oracle, alone and inside a mixed batch."""
    grey = read_golden("sample_grey.jxl")
    dec = pkg.decoder_builder().build()
    meta, px = dec.decode(grey)
    assert px.variant == "Uint16" and len(px.data) == meta.width * meta.height == 2000
    assert np.array_equal(np.asarray(px.data).reshape(50, 40, 1), jxlo.decode(grey, 1, jxlo.UINT16))
    a, _ = vc.encoded("heuristic")
    outs = pkg.decode_batch([a, grey, read_golden("sample_jpg.jxl"), grey], 3, np.uint8)
    assert np.array_equal(outs[1], jxlo.decode(grey, 3, jxlo.UINT8)) and np.array_equal(outs[3], outs[1])
    assert np.array_equal(outs[0], jxlo.decode(a, 3, jxlo.UINT8))


def test_orientation(pkg):
    # the write stage's undo_orientation on the GPU (all eight orientations, lossy RGB8 / RGBA16 and lossless RGBA8 in
    # one batch), the coded image with keep_orientation, and the event API's upright size (lib/jxl/decode.cc:2083-2090)
    img = vc.crop(70, 100, 100, 200)
    rgba = np.random.default_rng(3).integers(0, 256, (37, 53, 4)).astype(np.uint16)
    files = [jxlo.encode_vardct(img, strategy_mode=2, orientation=o) for o in range(1, 9)]
    files += [jxlo.encode_vardct(img[:35, :50], strategy_mode=2, upsampling=2, orientation=o) for o in (3, 6)]
    files += [jxlo.encode_modular(rgba, bits=8, alpha=True, orientation=o) for o in (2, 5, 7, 8)]
    for nc, npdt, dt in [(3, np.uint8, jxlo.UINT8), (4, np.uint16, jxlo.UINT16), (4, np.uint8, jxlo.UINT8)]:
        outs = pkg.decode_batch(files, nc, npdt)
        kept = pkg.decode_batch(files, nc, npdt, keep_orientation=True)
        for f, o, k in zip(files, outs, kept):
            assert np.array_equal(o, jxlo.decode(f, nc, dt, undo_orientation=True))
            assert np.array_equal(k, jxlo.decode(f, nc, dt))
    dec = pkg.decoder_builder().build()
    meta, px = dec.decode(files[5])  # orientation 6: rotated, the upright image is 100 x 70 -> 70 wide, 100 high
    assert (meta.width, meta.height) == (70, 100) and meta.orientation == 1
    assert np.array_equal(np.asarray(px.data).reshape(100, 70, 3), jxlo.decode(files[5], 3, jxlo.UINT8, undo_orientation=True))
    dec = pkg.decoder_builder().skip_reorientation(True).build()
    meta, px = dec.decode(files[5])
    assert (meta.width, meta.height) == (100, 70) and meta.orientation == 6
    assert np.array_equal(np.asarray(px.data).reshape(70, 100, 3), jxlo.decode(files[5], 3, jxlo.UINT8))


def test_alpha_channel_in_lossy_frames(pkg):
    # VarDCT frames with an alpha extra channel on the GPU: the Modular streams chained behind the AC coefficients
    # (second Modular launch), the global stream of a single-section frame (probe round); batch and event API
    def rgba(h, w, y0=100, x0=200):
        img = vc.crop(h, w, y0, x0)
        a = (img[:, :, 0].astype(np.int32) + np.arange(w)[None, :] * 3) % 256
        a[h // 3:h // 2, w // 4:w // 2] = 255
        return np.dstack([img, a.astype(np.uint8)])
    cases = [rgba(300, 520), rgba(200, 256), rgba(40, 50), rgba(257, 263, 700, 100), rgba(1000, 1500, 0, 0)]
    files = [jxlo.encode_vardct(c, strategy_mode=2) for c in cases]
    files.append(jxlo.encode_vardct(cases[0], strategy_mode=1, random_side_info=True, seed=3, epf_iters=1, dc_tree=1))
    files.append(vc.encoded("odd_size")[0])
    for nc, npdt, dt in [(4, np.uint8, jxlo.UINT8), (4, np.uint16, jxlo.UINT16), (3, np.uint8, jxlo.UINT8), (4, np.float32, jxlo.FLOAT)]:
        outs = pkg.decode_batch(files, nc, npdt)
        for f, o in zip(files, outs):
            assert np.array_equal(o.view(np.uint8), jxlo.decode(f, nc, dt).view(np.uint8))
    outs = pkg.decode_batch(files[:5], 4, np.uint8)
    for o, c in zip(outs, cases):
        assert np.array_equal(o[:, :, 3], c[:, :, 3])
    dec = pkg.decoder_builder().build()
    meta, px = dec.decode(files[0])
    assert meta.has_alpha_channel and px.variant == "Uint8"
    assert np.array_equal(np.asarray(px.data).reshape(300, 520, 4), jxlo.decode(files[0], 4, jxlo.UINT8))


def test_splines_in_lossy_frames(pkg):
    # splines drawn over VarDCT frames in front of the colour transform (DevColorStore -> DevSplineAdd), in a batch
    # with spline-free frames and the reference's Modular spline fixture
    img = vc.crop(300, 420, 100, 200)
    files = [jxlo.encode_vardct(img, strategy_mode=2, splines=5), jxlo.encode_vardct(img[:60, :70], strategy_mode=2, splines=3),
             jxlo.encode_vardct(vc.crop(520, 300, 50, 60), strategy_mode=3, splines=9, epf_iters=1, seed=4),
             jxlo.encode_vardct(vc.crop(1000, 1500, 0, 0), strategy_mode=2, splines=15),
             vc.encoded("odd_size")[0], read_golden("2bit.jxl"),
             jxlo.encode_vardct(vc.crop(64, 96, 100, 200), strategy_mode=2, upsampling=2, splines=4)]  # (in front of the upsampling)
    for nc, npdt, dt in [(3, np.uint8, jxlo.UINT8), (4, np.uint8, jxlo.UINT8), (3, np.uint16, jxlo.UINT16), (4, np.float32, jxlo.FLOAT)]:
        outs = pkg.decode_batch(files, nc, npdt)
        for f, o in zip(files, outs):
            assert np.array_equal(o.view(np.uint8), jxlo.decode(f, nc, dt).view(np.uint8))
    o = jxlo.encode_vardct(img[:100, :150], strategy_mode=2, splines=4, orientation=7)
    assert np.array_equal(pkg.decode_batch([o], 3, np.uint8)[0], jxlo.decode(o, 3, jxlo.UINT8, undo_orientation=True))
