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
