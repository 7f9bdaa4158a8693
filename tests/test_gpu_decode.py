"""GPU: parity of the CUDA path (through the C ABI) with the oracle and the reference's goldens."""
import hashlib
import json
import os

import numpy as np
import pytest

import jxlo
from conftest import GOLDEN, read_golden

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(GOLDEN, "golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_simple_sample_uint16(pkg):
    # jpegxl-rs/src/tests/decode.rs:44-67 + image.rs:158-174
    dec = pkg.decoder_builder().build()
    meta, px = dec.decode(read_golden("sample.jxl"))
    assert px.variant == "Uint16"
    assert len(px) == meta.width * meta.height * 4
    assert (meta.width, meta.height, meta.num_color_channels, meta.has_alpha_channel) == (40, 50, 3, True)
    assert sha(px.data) == G["sample.jxl"]["sha256"]


def test_bench_jxl_rgba8_golden(pkg):
    # the reference's criterion bench input (jpegxl-rs/benches/decode.rs:10-40): decode_with::<u8>
    dec = pkg.decoder_builder().build()
    meta, px = dec.decode_with(read_golden("bench.jxl"), np.uint8)
    assert px.size == 2122 * 1433 * 4
    assert sha(px) == G["bench.jxl"]["sha256"]


@pytest.mark.parametrize("nch", [1, 2, 3, 4])
@pytest.mark.parametrize("dt,npdt", [(jxlo.UINT8, np.uint8), (jxlo.UINT16, np.uint16), (jxlo.FLOAT16, np.float16),
                                     (jxlo.FLOAT, np.float32)])
def test_pixel_types_match_oracle(pkg, nch, dt, npdt):
    # jpegxl-rs/src/tests/decode.rs:95-120
    data = read_golden("sample.jxl")
    dec = pkg.decoder_builder().pixel_format(pkg.PixelFormat(num_channels=nch)).build()
    if nch < 3:
        # lib/jxl/decode.cc:2334-2337: grey output of a colour image is an API error
        with pytest.raises(pkg.DecodeError):
            dec.decode_with(data, npdt)
        return
    meta, px = dec.decode_with(data, npdt)
    want = jxlo.decode(data, nch, dt)
    assert px.size == 40 * 50 * nch
    assert np.array_equal(px.view(np.uint8), want.reshape(-1).view(np.uint8))


def test_big_endian_and_alignment(pkg):
    data = read_golden("sample.jxl")
    dec = pkg.decoder_builder().pixel_format(pkg.PixelFormat(num_channels=3, endianness=pkg.JXL_BIG_ENDIAN)).build()
    _, be = dec.decode_with(data, np.uint16)
    want = jxlo.decode(data, 3, jxlo.UINT16)
    assert np.array_equal(be.reshape(want.shape), want)  # wrapper converts to native values like jpegxl-rs
    bd = pkg.BatchDecoder(0)
    bd.set_input([data], 3, pkg.JXL_TYPE_UINT8, align=64)
    bd.run()
    bd.wait()
    raw = bd.read_output(0)
    ref = jxlo.Decoded(data).pixels(3, jxlo.UINT8, align=64)
    stride = 128
    assert raw.size == stride * 50
    assert np.array_equal(raw.reshape(50, stride)[:, :120], ref.reshape(50, stride)[:, :120])


def test_batch_of_mixed_frames_matches_oracle(pkg):
    a, b = read_golden("bench.jxl"), read_golden("sample.jxl")
    files = [a, b, a, b, a]
    outs = pkg.decode_batch(files, 4, np.uint8)
    wa, wb = jxlo.decode(a, 4, jxlo.UINT8), jxlo.decode(b, 4, jxlo.UINT8)
    for f, o in zip(files, outs):
        assert np.array_equal(o, wa if f is a else wb)


def test_full_size_batch_property(pkg):
    # size-independent property at bench size: every replica of the same codestream decodes to the same
    # checksum, equal to the golden one
    a = read_golden("bench.jxl")
    outs = pkg.decode_batch([a] * 16, 4, np.uint8)
    assert {sha(o) for o in outs} == {G["bench.jxl"]["sha256"]}


def test_errors(pkg):
    dec = pkg.decoder_builder().build()
    with pytest.raises(pkg.InvalidInput):
        dec.decode(b"")
    with pytest.raises(pkg.InvalidInput):
        dec.decode(b"\0" * 64)
    with pytest.raises(pkg.DecodeError):
        dec.decode(read_golden("sample.jxl")[:300])
    # corrupted payload: flipped bytes inside a group section must be reported, not silently decoded
    bad = bytearray(read_golden("bench.jxl"))
    for i in range(600000, 600064):
        bad[i] ^= 0x5A
    bd = pkg.BatchDecoder(0)
    bd.set_input([bytes(bad)], 4, pkg.JXL_TYPE_UINT8)
    bd.run()
    with pytest.raises(pkg.GenericError):
        bd.wait()


def test_modular_cases_match_oracle(pkg):
    """The Modular branches the reference's fixtures do not reach (tests/modular_cases.py: palette, delta palette with
    explicit and implicit entries, squeeze, all predictors / properties through random trees, prefix codes, LZ77 with
    and without special distances), one batch per output format, bit for bit against the oracle."""
    import modular_cases as mc
    by_format = {}
    for name, (_, _, fmt) in mc.CASES.items():
        by_format.setdefault(fmt, []).append(name)
    npdt = {jxlo.UINT8: np.uint8, jxlo.UINT16: np.uint16}
    for (nc, dt), names in by_format.items():
        files = [mc.encoded(n)[0] for n in names]
        outs = pkg.decode_batch(files, nc, npdt[dt])
        for n, f, o in zip(names, files, outs):
            assert np.array_equal(o, jxlo.decode(f, nc, dt)), n
            img = mc.encoded(n)[1]
            if nc == img.shape[2]:
                assert np.array_equal(o, img), n  # and lossless against the source samples


def test_sample_2bit_as_the_reference_test(pkg):
    # jpegxl-rs/src/tests/decode.rs:69-80: decoder.decode(SAMPLE_JXL_2BIT) -> Pixels::Uint8 of width * height * 3 samples;
    # here also equal to the oracle (splines evaluated per pixel in the output kernel), in a batch with other files
    d = read_golden("2bit.jxl")
    dec = pkg.decoder_builder().build()
    meta, px = dec.decode(d)
    assert px.variant == "Uint8" and np.asarray(px.data).size == meta.width * meta.height * 3 == 800 * 600 * 3
    assert np.array_equal(np.asarray(px.data).reshape(600, 800, 3), jxlo.decode(d, 3, jxlo.UINT8))
    outs = pkg.decode_batch([read_golden("sample.jxl"), d, d], 4, np.uint16)
    assert np.array_equal(outs[1], jxlo.decode(d, 4, jxlo.UINT16)) and np.array_equal(outs[2], outs[1])
    assert np.array_equal(outs[0], jxlo.decode(read_golden("sample.jxl"), 4, jxlo.UINT16))


def test_streaming_calls_plan_ahead_and_run_to_host(pkg):
    # PlanBatch / CommitPlan / RunToHost (include/jxl_b200.h): the next batch is parsed while the current one decodes,
    # and the frames land in the caller's buffers wave by wave -- two different batches alternating on one handle
    # (lossy multi-group, lossy with alpha, single-section probe round, Modular, splines), pinned and pageable buffers
    import torch
    import vardct_cases as vc
    img = vc.crop(300, 520, 100, 200)
    rgba = np.dstack([img, img[:, :, 0] ^ img[:, :, 2]])
    a = [vc.encoded("heuristic")[0], read_golden("sample.jxl"), jxlo.encode_vardct(rgba, strategy_mode=2), read_golden("2bit.jxl")]
    b = [read_golden("bench.jxl"), jxlo.encode_vardct(img[:60, :70], strategy_mode=2, splines=3), vc.encoded("three_passes")[0]]
    want = {0: [jxlo.decode(f, 4, jxlo.UINT8) for f in a], 1: [jxlo.decode(f, 4, jxlo.UINT8) for f in b]}
    d = pkg.BatchDecoder(0)
    stream = torch.cuda.Stream()
    d.plan(a, 4, pkg.JXL_TYPE_UINT8)
    for step in range(5):
        cur = step % 2
        d.commit()
        sizes = [d.out_size(i) for i in range(len(want[cur]))]
        if step < 3:
            outs = [torch.empty(n, dtype=torch.uint8).pin_memory().numpy() for n in sizes]
        else:
            outs = [np.empty(n, dtype=np.uint8) for n in sizes]
        d.run_to_host(outs, stream.cuda_stream)
        d.plan(b if cur == 0 else a, 4, pkg.JXL_TYPE_UINT8)  # while the kernels run
        d.wait(stream.cuda_stream)
        for o, w in zip(outs, want[cur]):
            assert np.array_equal(o.reshape(w.shape), w)
        assert np.array_equal(d.read_output(0).reshape(want[cur][0].shape), want[cur][0])  # the device copy is still there
    with pytest.raises(pkg.GenericError):
        d.run_to_host([np.empty(4, dtype=np.uint8)] * len(want[0]), stream.cuda_stream)  # buffers too small
    d2 = pkg.BatchDecoder(0)
    with pytest.raises(pkg.GenericError):
        d2.commit()  # nothing planned
