"""The encode path: the kernels' device functions on the CPU (tests/emul) and the CUDA encoder through the C ABI must
both produce the oracle encoder's codestream byte for byte (oracle/jxlo_encode.h, gradient DC tree), and the
stream must decode (oracle and GPU decoder) to the same pixels."""
import numpy as np
import pytest

import emul_lib
import jxlo
import vardct_cases as vc

CFL = True  # the chroma-from-luma fit (row E6) runs in both encoders

CASES = [
    ("dct8", lambda: vc.crop(300, 400), dict(strategy_mode=0)),
    ("heuristic", lambda: vc.crop(300, 400, 500, 700), dict(strategy_mode=2)),
    ("single_group", lambda: vc.crop(200, 200, 100, 100), dict(strategy_mode=2)),
    ("d2_5_no_filters", lambda: vc.crop(520, 700, 300, 0), dict(strategy_mode=2, distance=2.5)),
    ("synthetic", lambda: vc.synthetic(333, 517, 4), dict(strategy_mode=2, distance=0.7)),
    # adaptive quantisation (row E3): no Gaborish (InitialQuantField sees 0.62 * distance and the transform planes
    # themselves), an image smaller than one 64x64 tile, a distance past the mean / max mixer ramp
    ("no_gab", lambda: vc.crop(150, 210, 640, 900), dict(strategy_mode=2, gab=False, epf_iters=0)),
    ("tiny", lambda: vc.crop(37, 50, 700, 1000), dict(strategy_mode=2)),
    ("d4", lambda: vc.crop(200, 330, 900, 1200), dict(strategy_mode=2, distance=4.0)),
]


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_emulated_encoder_is_byte_exact(name):
    _, make, kw = next(c for c in CASES if c[0] == name)
    img = make()
    want = jxlo.encode_vardct(img, dc_tree=1, cfl=CFL, **kw)
    assert emul_lib.encode(img, **kw) == want
    # and the stream is a picture of the input
    out = jxlo.decode(want, 3, jxlo.UINT8)
    err = out.astype(np.float64) - img
    assert 10 * np.log10(255 ** 2 / (err ** 2).mean()) > 27


@pytest.mark.gpu
def test_gpu_encoder_is_byte_exact_in_a_batch(pkg):
    imgs = [make() for _, make, kw in CASES if kw.get("distance", 1.0) == 1.0 and kw["strategy_mode"] == 2]
    enc = pkg.encoder_builder().quality(1.0).build()
    outs = enc.encode_batch(imgs)
    for img, o in zip(imgs, outs):
        assert o.data == jxlo.encode_vardct(img, dc_tree=1, cfl=CFL, strategy_mode=2)


@pytest.mark.gpu
@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_gpu_encoder_each_case(pkg, name):
    _, make, kw = next(c for c in CASES if c[0] == name)
    img = make()
    enc = pkg.JxlEncoder(quality=kw.get("distance", 1.0), speed=1 if kw["strategy_mode"] == 0 else 7)
    if "gab" in kw or "epf_iters" in kw:  # loop-filter header fields: only the batch call exposes them
        got = enc.encode_batch([img], gaborish=kw.get("gab", True), epf_iters=kw.get("epf_iters", 2))[0].data
    else:
        got = enc.encode(img.reshape(-1), img.shape[1], img.shape[0]).data
    assert got == jxlo.encode_vardct(img, dc_tree=1, cfl=CFL, **kw)


@pytest.mark.gpu
def test_gpu_round_trip_4k(pkg):
    # BASELINE.json configs[2]: encode a 4K RGB8 frame at d = 1.0 and decode it again, both on the GPU
    img = vc.frame_4k()
    enc = pkg.encoder_builder().build()
    data = enc.encode(img).data
    assert data == jxlo.encode_vardct(img, dc_tree=1, cfl=CFL, strategy_mode=2)
    out = pkg.decode_batch([data], 3, np.uint8)[0]
    err = out.astype(np.float64) - img
    assert 10 * np.log10(255 ** 2 / (err ** 2).mean()) > 33
    assert np.array_equal(out, jxlo.decode(data, 3, jxlo.UINT8))


@pytest.mark.gpu
def test_event_api_container_output_round_trips(pkg):
    # JxlEncoderUseContainer: signature box + ftyp + jxlc around the same codestream (lib/jxl/encode.cc:376-560)
    img = vc.crop(300, 400, 500, 700)
    plain = pkg.encoder_builder().build().encode(img).data
    boxed = pkg.encoder_builder().use_container(True).init_buffer_size(64).build().encode(img).data
    assert pkg.check_valid_signature(boxed) and boxed[:12] == b"\0\0\0\x0cJXL \r\n\x87\n" and boxed.endswith(plain)
    assert plain == jxlo.encode_vardct(img, dc_tree=1, cfl=CFL, strategy_mode=2)
    meta, px = pkg.decoder_builder().build().decode(boxed)
    assert (meta.width, meta.height) == (400, 300)
    assert np.array_equal(np.asarray(px.data).reshape(300, 400, 3), jxlo.decode(plain, 3, jxlo.UINT8))


def _lossless_cases():
    img = vc.crop(300, 520, 100, 200)
    rng = np.random.default_rng(5)
    return [("rgb8", img, dict(bits=8, alpha=False, rct=6)),
            ("one_group", np.ascontiguousarray(img[:100, :128]), dict(bits=8, alpha=False, rct=6)),
            ("four_groups", np.ascontiguousarray(img[:200, :256]), dict(bits=8, alpha=False, rct=6)),
            ("rgba8", np.dstack([img, img[:, :, 0] ^ img[:, :, 1]]), dict(bits=8, alpha=True, rct=6)),
            ("grey8", np.ascontiguousarray(img[:, :, 1:2]), dict(bits=8, alpha=False, rct=-1)),
            ("grey_alpha8", np.ascontiguousarray(img[:257, :263, :2]), dict(bits=8, alpha=True, rct=-1)),
            ("rgb16", img.astype(np.uint16) * 257 ^ 0x35, dict(bits=16, alpha=False, rct=6)),
            ("noise16", rng.integers(0, 65536, (70, 300, 4)).astype(np.uint16), dict(bits=16, alpha=True, rct=6))]


@pytest.mark.parametrize("case", range(8))
def test_lossless_encoder_logic_matches_the_oracle(case):
    # the lossless (Modular) encoder's device functions on the CPU: the stream equals the oracle's EncodeModular under the
    # same fixed choices (YCoCg-R, libjxl's fixed gradient tree, 128 x 128 groups) byte for byte, and decodes -- by the
    # oracle's decoder and by the decode kernels' logic -- to the input
    name, a, kw = _lossless_cases()[case]
    got = emul_lib.encode_lossless(a)
    assert got == jxlo.encode_modular(a.astype(np.uint16), tree=1, predictor=5, group_size_shift=0, **kw), name
    dt = jxlo.UINT8 if a.dtype == np.uint8 else jxlo.UINT16
    assert np.array_equal(jxlo.decode(got, a.shape[2], dt), a)
    assert np.array_equal(emul_lib.decode([got], a.shape[2], dt, [a.shape[:2]])[0], a)


@pytest.mark.gpu
def test_gpu_lossless_encoder(pkg):
    # the CUDA lossless encoder: byte-exact against the oracle, GPU decode of its output == input (round trip), through the
    # batch call and through the libjxl-compatible JxlEncoder* calls the way jpegxl-rs drives them (lossless(true),
    # uses_original_profile(true), has_alpha)
    cases = _lossless_cases()
    for kind in (np.uint8, np.uint16):
        for nch in (1, 2, 3, 4):
            batch = [(n, a, kw) for n, a, kw in cases if a.dtype == kind and a.shape[2] == nch]
            if not batch:
                continue
            enc = pkg.JxlEncoder(lossless=True, uses_original_profile=True, has_alpha=nch in (2, 4))
            outs = enc.encode_batch([a for _, a, _ in batch])
            for (n, a, kw), o in zip(batch, outs):
                assert o.data == jxlo.encode_modular(a.astype(np.uint16), tree=1, predictor=5, group_size_shift=0, **kw), n
            dec = pkg.decode_batch([o.data for o in outs], nch, kind)
            for (n, a, kw), d in zip(batch, dec):
                assert np.array_equal(d, a), n
    name, a, kw = cases[3]  # RGBA8 through the event-style API
    enc = pkg.JxlEncoder(lossless=True, uses_original_profile=True, has_alpha=True)
    res = enc.encode(a)
    assert res.data == jxlo.encode_modular(a.astype(np.uint16), tree=1, predictor=5, group_size_shift=0, **kw)
    big = vc.frame_4k()
    out = pkg.JxlEncoder(lossless=True, uses_original_profile=True).encode_batch([big])[0].data
    assert np.array_equal(pkg.decode_batch([out], 3, np.uint8)[0], big)


def _rgba(h, w, y0=100, x0=200):
    img = vc.crop(h, w, y0, x0)
    a = (img[:, :, 0].astype(np.int32) + np.arange(w)[None, :] * 3) % 256
    a[h // 3:h // 2, w // 4:w // 2] = 255
    return np.dstack([img, a.astype(np.uint8)])


ALPHA_SIZES = [(300, 520), (200, 256), (40, 50), (257, 263)]


@pytest.mark.parametrize("h,w", ALPHA_SIZES)
def test_lossy_encoder_with_alpha_logic_matches_the_oracle(h, w):
    # RGBA8 -> lossy colour + a lossless 8-bit alpha extra channel in the frame's Modular sub-streams (the global stream
    # when the image fits one group, else behind the coefficients of every AC group): byte-exact against the oracle's
    # encoder, alpha decodes to the input, colour to what the stream without alpha decodes to
    im = _rgba(h, w)
    got = emul_lib.encode(im)
    assert got == jxlo.encode_vardct(im, dc_tree=1, cfl=CFL, strategy_mode=2)
    dec = jxlo.decode(got, 4, jxlo.UINT8)
    assert np.array_equal(dec[:, :, 3], im[:, :, 3])
    assert np.array_equal(dec[:, :, :3], jxlo.decode(emul_lib.encode(np.ascontiguousarray(im[:, :, :3])), 3, jxlo.UINT8))
    assert np.array_equal(emul_lib.decode([got], 4, jxlo.UINT8, [(h, w)])[0], dec)


@pytest.mark.gpu
def test_gpu_lossy_encoder_with_alpha(pkg):
    ims = [_rgba(h, w) for h, w in ALPHA_SIZES] + [_rgba(1000, 1500, 0, 0)]
    enc = pkg.JxlEncoder(has_alpha=True)
    outs = enc.encode_batch(ims)
    for im, o in zip(ims, outs):
        assert o.data == jxlo.encode_vardct(im, dc_tree=1, cfl=CFL, strategy_mode=2)
    dec = pkg.decode_batch([o.data for o in outs], 4, np.uint8)
    for im, d, o in zip(ims, dec, outs):
        assert np.array_equal(d[:, :, 3], im[:, :, 3])
        assert np.array_equal(d, jxlo.decode(o.data, 4, jxlo.UINT8))
    res = pkg.JxlEncoder(has_alpha=True).encode(ims[0])  # the libjxl-compatible calls, as jpegxl-rs drives them
    assert res.data == outs[0].data


def test_register_dct8_is_bit_identical_to_the_staged_forward_dct():
    # the DCT8X8 fast path of k_enc_coeffs (DevFwdDct8: a line in registers) states the generic staged forward DCT
    # (CoopDCT, n = 8) operation by operation: bit-identical coefficients on random lines of every magnitude
    import ctypes
    L = emul_lib.lib()
    L.jxlb_emul_fwd_dct8_mismatches.restype = ctypes.c_long
    assert L.jxlb_emul_fwd_dct8_mismatches(ctypes.c_uint32(7), ctypes.c_uint32(200000)) == 0


def test_reciprocal_division_of_the_rans_writer_is_exact():
    # DevRansPushWarp divides the coder state by a symbol frequency as umulhi(state, 0xFFFFFFFF // f) plus one
    # correction step: exact for every frequency of a 12-bit ANS table and every 32-bit state (the estimate is never more
    # than one short: 0xFFFFFFFF // f >= 2^32 / f - 1)
    rng = np.random.default_rng(11)
    for f in range(1, 4097):
        edge = np.array([0, 1, f - 1, f, f + 1, 2**32 - 1, 2**32 - f, (2**32 // f) * f - 1, ((2**32 - 1) // f) * f], dtype=np.uint64)
        st = np.concatenate([rng.integers(0, 2**32, 64, dtype=np.uint64), edge[edge < 2**32]])
        q = (st * np.uint64(0xFFFFFFFF // f)) >> np.uint64(32)
        r = st - q * np.uint64(f)
        fix = r >= f
        q, r = q + fix, r - fix * np.uint64(f)
        assert np.array_equal(q, st // np.uint64(f)) and np.all(r < f), f
