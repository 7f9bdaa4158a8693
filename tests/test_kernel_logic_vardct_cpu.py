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


@pytest.mark.parametrize("name", ["heuristic", "odd_size"])
def test_per_pixel_render_kernels_still_match(name, monkeypatch):
    # frames without patches take the fused render tile (DevRenderTile) by default; the per-pixel kernels they
    # replaced stay in use for frames with patches and must give the same samples
    monkeypatch.setenv("JXLB_EMUL_UNFUSED", "1")
    data, shape = vc.encoded(name)
    got = emul_lib.decode([data], 3, jxlo.UINT8, [shape])[0]
    assert np.array_equal(got, jxlo.decode(data, 3, jxlo.UINT8))


@pytest.mark.parametrize("h,w,epf,gab", [(33, 70, 3, True), (5, 3, 3, True), (64, 128, 2, False), (100, 65, 1, True),
                                         (96, 192, 0, True), (71, 64, 3, False)])
def test_fused_render_tile_borders(h, w, epf, gab):
    # tile edges against frame edges: frames narrower than the halo (mirroring bounces twice), frames that end exactly
    # on a tile edge, one row / column past it, every combination of stages
    img = vc.crop(h, w, 640, 960)
    data = jxlo.encode_vardct(img, strategy_mode=2, gab=gab, epf_iters=epf, random_side_info=True, seed=h + w)
    for nc, dt in [(3, jxlo.UINT8), (4, jxlo.UINT16), (3, jxlo.FLOAT)]:
        got = emul_lib.decode([data], nc, dt, [(h, w)])[0]
        assert np.array_equal(got.view(np.uint8), jxlo.decode(data, nc, dt).view(np.uint8))


def test_small_div_is_exact_over_its_domain():
    # DevSmallDiv (jxlb_vardct_dev.h): i / d as (i * ceil(2^20 / d)) >> 20 for the tile index splits of k_render_fused
    # (i < 80 * 46 cells, 64 <= d <= 80 columns)
    for d in range(64, 81):
        m = ((1 << 20) + d - 1) // d
        i = np.arange(0, 4096, dtype=np.uint64)
        assert np.array_equal((i * m) >> 20, i // d)
        assert int(i.max()) * m < 2 ** 32


SINGLE_SECTION = [(80, 100, 1, 3), (256, 256, 2, 4), (17, 9, 1, 5), (200, 256, 0, 6)]


@pytest.mark.parametrize("h,w,mode,seed", SINGLE_SECTION)
def test_single_section_frames_match_oracle(h, w, mode, seed):
    # one group, one pass: DC global, DC group, AC global and AC group share a section; the planner learns where
    # the DC / AC-metadata chain ends from a probe launch of the Modular decode kernel (ProbeCtx)
    img = vc.crop(h, w, 300, 500)
    data = jxlo.encode_vardct(img, strategy_mode=mode, random_side_info=True, epf_iters=3, seed=seed)
    got = emul_lib.decode([data], 3, jxlo.UINT8, [(h, w)])[0]
    assert np.array_equal(got, jxlo.decode(data, 3, jxlo.UINT8))


def test_g3_sample_jpg_jxl_matches_oracle():
    # the reference's own VarDCT fixture (libjxl-written: container, YCbCr, raw quantisation tables decoded in a
    # second probe round, single section), in a batch with other files
    jpg = read_golden("sample_jpg.jxl")
    a, sa = vc.encoded("odd_size")
    got = emul_lib.decode([jpg, a, jpg], 3, jxlo.UINT8, [(50, 40), sa, (50, 40)])
    want = jxlo.decode(jpg, 3, jxlo.UINT8)
    assert np.array_equal(got[0], want) and np.array_equal(got[2], want)
    assert np.array_equal(got[1], jxlo.decode(a, 3, jxlo.UINT8))
    gotf = emul_lib.decode([jpg], 3, jxlo.FLOAT, [(50, 40)])[0]
    assert np.array_equal(gotf.view(np.uint32), jxlo.decode(jpg, 3, jxlo.FLOAT).view(np.uint32))


def test_g4_sample_grey_jxl_matches_oracle():
    # the reference's grey fixture (libjxl-written): a reference-only XYB Modular frame, then a single-section VarDCT
    # frame with patches from it, Gaborish and one EPF iteration; 16-bit grey output as jpegxl-rs asks for
    # (jpegxl-rs/src/tests/decode.rs:82-93), RGB8 and float
    grey = read_golden("sample_grey.jxl")
    for nc, dt in [(1, jxlo.UINT16), (3, jxlo.UINT8), (1, jxlo.FLOAT), (2, jxlo.UINT16)]:
        got = emul_lib.decode([grey], nc, dt, [(50, 40)])[0]
        assert np.array_equal(got.view(np.uint8), jxlo.decode(grey, nc, dt).view(np.uint8))
    a, sa = vc.encoded("odd_size")
    got = emul_lib.decode([a, grey, read_golden("sample_jpg.jxl")], 3, jxlo.UINT8, [sa, (50, 40), (50, 40)])
    assert np.array_equal(got[1], jxlo.decode(grey, 3, jxlo.UINT8))
    assert np.array_equal(got[0], jxlo.decode(a, 3, jxlo.UINT8))


def test_ac_metadata_channels_take_the_ynw_table_path(monkeypatch):
    # libjxl's fixed AC-metadata tree (tests y, N, W only) is decoded through the (y, N, W) bucket table
    # (DevChannel::nw_lut); with JXLB200_NO_NW_LUT=1 the same channels take the generic tree walk. Same samples.
    img = vc.crop(200, 300, 100, 200)
    # (dc_tree=1: libjxl's fixed gradient DC tree, one property -> the 1-D table variant of the same path)
    cases = [jxlo.encode_vardct(img, strategy_mode=3, distance=1.0, epf_iters=1),
             jxlo.encode_vardct(img, strategy_mode=1, random_side_info=True, seed=11, epf_iters=3),
             jxlo.encode_vardct(img, strategy_mode=2, dc_tree=1)]
    for data in cases:
        want = jxlo.decode(data, 3, jxlo.UINT8)
        got = emul_lib.decode([data], 3, jxlo.UINT8, [(200, 300)])[0]
        chans, wp, nw = emul_lib.last_plan_stats()
        assert np.array_equal(got, want)
        assert chans == 7 and wp + nw >= 4 and nw >= 1, (chans, wp, nw)
        monkeypatch.setenv("JXLB200_NO_NW_LUT", "1")
        got2 = emul_lib.decode([data], 3, jxlo.UINT8, [(200, 300)])[0]
        assert emul_lib.last_plan_stats()[2] == 0
        assert np.array_equal(got2, want)
        monkeypatch.delenv("JXLB200_NO_NW_LUT")


def test_dc_chains_take_the_warp_cooperative_path(monkeypatch):
    # DC-group chains under libjxl's fixed trees are decoded one per warp (k_modular_decode_coop, DevDecodeModularStreamCoop:
    # bit window, branch-free hybrid integers, per-row predictor paths); JXLB200_NO_COOP=1 sends them back to the
    # lock-step kernels. Weighted DC tree (34 clusters: two of them outside the 32 lanes -> the unspeculated decode),
    # gradient DC tree (DevCoopFastRow), multi-group frames, odd sizes. Same samples either way.
    imgs = [vc.crop(200, 300, 100, 200), vc.crop(257, 263, 700, 100), vc.crop(520, 700, 300, 500)]
    cases = [(jxlo.encode_vardct(imgs[0], strategy_mode=3, distance=1.0, epf_iters=1), imgs[0].shape[:2]),
             (jxlo.encode_vardct(imgs[1], strategy_mode=1, random_side_info=True, seed=11, epf_iters=3, dc_tree=1), imgs[1].shape[:2]),
             (jxlo.encode_vardct(imgs[2], strategy_mode=2, dc_tree=1, distance=0.5), imgs[2].shape[:2]),
             (jxlo.encode_vardct(imgs[2], strategy_mode=2, dc_tree=0, distance=4.0), imgs[2].shape[:2])]
    for data, shape in cases:
        want = jxlo.decode(data, 3, jxlo.UINT8)
        got = emul_lib.decode([data], 3, jxlo.UINT8, [shape])[0]
        coop, total = emul_lib.last_coop_streams()
        assert coop == total >= 1, (coop, total)
        assert np.array_equal(got, want)
        monkeypatch.setenv("JXLB200_NO_COOP", "1")
        got2 = emul_lib.decode([data], 3, jxlo.UINT8, [shape])[0]
        assert emul_lib.last_coop_streams()[0] == 0
        assert np.array_equal(got2, want)
        monkeypatch.delenv("JXLB200_NO_COOP")
    # the 64-bit predictor path of the same kernel
    data, shape = cases[1]
    got = emul_lib.decode([data], 3, jxlo.UINT8, [shape], endianness=0x100)[0]
    assert np.array_equal(got, jxlo.decode(data, 3, jxlo.UINT8))


@pytest.mark.parametrize("o", range(1, 9))
def test_orientation_is_undone_like_the_write_stage(o):
    # libjxl's default output turns the image upright (stage_write.cc:131-135, :163-172, :345-366: flips, dither pattern at
    # the flipped position, transposed store). The kernels' store (DevOrient) against the oracle's restatement and, for
    # 16-bit samples (no dither), against the plain numpy flips of the coded image. Lossy (fused tile and, upsampled, the
    # per-pixel colour kernel) and lossless Modular (RGBA).
    import modular_cases as mc
    img = vc.crop(70, 100, 100, 200)
    flip_x, flip_y, transpose = o in (2, 3, 8, 7), o in (4, 3, 6, 7), o >= 5
    def upright(a):
        a = a[:, ::-1] if flip_x else a
        a = a[::-1] if flip_y else a
        return a.transpose(1, 0, 2) if transpose else a
    rng = np.random.default_rng(o)
    rgba = rng.integers(0, 256, (37, 53, 4)).astype(np.uint16)
    files = [jxlo.encode_vardct(img, strategy_mode=2, orientation=o),
             jxlo.encode_vardct(img[:35, :50], strategy_mode=2, upsampling=2, orientation=o),
             jxlo.encode_modular(rgba, bits=8, alpha=True, orientation=o)]
    shapes = [(70, 100), (70, 100), (37, 53)]
    out_shapes = [(w, h) if transpose else (h, w) for h, w in shapes]
    for nc, dt in [(3, jxlo.UINT8), (4, jxlo.UINT16), (4, jxlo.UINT8)]:
        got = emul_lib.decode(files, nc, dt, out_shapes, endianness=0x400)
        kept = emul_lib.decode(files, nc, dt, shapes)
        for g, k, f in zip(got, kept, files):
            want = jxlo.decode(f, nc, dt, undo_orientation=True)
            assert np.array_equal(g, want)
            assert np.array_equal(k, jxlo.decode(f, nc, dt))
            if dt == jxlo.UINT16:
                assert np.array_equal(g, upright(k))


def _rgba(h, w, y0=100, x0=200):
    img = vc.crop(h, w, y0, x0)
    a = (img[:, :, 0].astype(np.int32) + np.arange(w)[None, :] * 3) % 256
    a[h // 3:h // 2, w // 4:w // 2] = 255
    return np.dstack([img, a.astype(np.uint8)])


def test_vardct_frames_with_an_alpha_channel(monkeypatch):
    # Lossy frames with an alpha extra channel (lib/jxl/dec_frame.cc:266-365, :478-560): the alpha samples travel in the
    # frame's Modular sub-streams -- the global stream when the image fits one group (its end is where the DC group
    # starts in a single-section frame: a probe round), else one stream per AC group that starts where the group's
    # coefficients end (DevStream::chain_slot: position from the AC decode, GroupHeader parsed by the device, second
    # Modular launch). One per warp and, with JXLB200_NO_COOP=1, in lock-step bundles; every output type; alpha dropped
    # for RGB output; in a batch with files without alpha; with an orientation.
    cases = [_rgba(300, 520), _rgba(200, 256), _rgba(40, 50), _rgba(257, 263, 700, 100)]
    files = [jxlo.encode_vardct(c, strategy_mode=2) for c in cases]
    files.append(jxlo.encode_vardct(cases[0], strategy_mode=1, random_side_info=True, seed=3, epf_iters=1, dc_tree=1))
    files.append(jxlo.encode_vardct(cases[3], strategy_mode=2, orientation=6))
    shapes = [c.shape[:2] for c in cases] + [cases[0].shape[:2], cases[3].shape[:2]]
    for env in (None, "JXLB200_NO_COOP"):
        if env:
            monkeypatch.setenv(env, "1")
        for nc, dt in [(4, jxlo.UINT8), (4, jxlo.UINT16), (3, jxlo.UINT8), (4, jxlo.FLOAT)]:
            got = emul_lib.decode(files, nc, dt, shapes)
            for g, f in zip(got, files):
                assert np.array_equal(g.view(np.uint8), jxlo.decode(f, nc, dt).view(np.uint8))
        if env:
            monkeypatch.delenv(env)
    got = emul_lib.decode(files[:4], 4, jxlo.UINT8, shapes[:4])
    for g, c in zip(got, cases):
        assert np.array_equal(g[:, :, 3], c[:, :, 3])  # alpha is lossless
    plain, sp = vc.encoded("odd_size")
    mixed = emul_lib.decode([files[0], plain, read_golden("sample.jxl"), files[1]], 4, jxlo.UINT8, [shapes[0], sp, (50, 40), shapes[1]])
    for g, f in zip(mixed, [files[0], plain, read_golden("sample.jxl"), files[1]]):
        assert np.array_equal(g, jxlo.decode(f, 4, jxlo.UINT8))
    got = emul_lib.decode([files[5]], 4, jxlo.UINT8, [(263, 257)], endianness=0x400)[0]
    assert np.array_equal(got, jxlo.decode(files[5], 4, jxlo.UINT8, undo_orientation=True))


def test_splines_in_lossy_frames():
    # Splines over a VarDCT frame (lib/jxl/dec_frame.cc:286-305, render_pipeline/stage_splines.cc): read from DC global
    # in front of the DC quantisation, drawn in XYB with the frame's base colour correlation behind the loop filters;
    # multi-group and single-section (probe round) frames, with alpha and an orientation, beside spline-free frames and
    # the reference's spline fixture in one batch
    img = vc.crop(300, 420, 100, 200)
    files = [jxlo.encode_vardct(img, strategy_mode=2, splines=5), jxlo.encode_vardct(img[:60, :70], strategy_mode=2, splines=3),
             jxlo.encode_vardct(vc.crop(520, 300, 50, 60), strategy_mode=3, splines=9, epf_iters=1, seed=4),
             jxlo.encode_vardct(_rgba(257, 263, 700, 100), strategy_mode=2, splines=6),
             vc.encoded("odd_size")[0], read_golden("2bit.jxl")]
    shapes = [(300, 420), (60, 70), (520, 300), (257, 263), vc.encoded("odd_size")[1], (600, 800)]
    plain = jxlo.decode(jxlo.encode_vardct(img, strategy_mode=2), 3, jxlo.UINT8)
    assert (plain != jxlo.decode(files[0], 3, jxlo.UINT8)).any(axis=2).mean() > 0.05  # the splines are visible
    for nc, dt in [(3, jxlo.UINT8), (4, jxlo.UINT8), (3, jxlo.UINT16), (4, jxlo.FLOAT)]:
        got = emul_lib.decode(files, nc, dt, shapes)
        for g, f in zip(got, files):
            assert np.array_equal(g.view(np.uint8), jxlo.decode(f, nc, dt).view(np.uint8))
    o = jxlo.encode_vardct(img[:100, :150], strategy_mode=2, splines=4, orientation=7)
    got = emul_lib.decode([o], 3, jxlo.UINT8, [(150, 100)], endianness=0x400)[0]
    assert np.array_equal(got, jxlo.decode(o, 3, jxlo.UINT8, undo_orientation=True))
    # Modular frames with splines written by the oracle's encoder (2bit.jxl is the reference's own)
    m = np.random.default_rng(1).integers(0, 256, (120, 200, 3)).astype(np.uint16)
    e = jxlo.encode_modular(m, bits=8, tree=1, splines=4)
    assert np.array_equal(emul_lib.decode([e], 3, jxlo.UINT8, [(120, 200)])[0], jxlo.decode(e, 3, jxlo.UINT8))


def test_splines_in_an_upsampled_frame():
    # libjxl draws splines on the coded planes in front of the upsampling stage (lib/jxl/dec_cache.cc:178-196): the
    # upsampling kernel adds them to every tap (DevUpsamplePixel); 2x and 4x, every output type, with an orientation
    img = vc.crop(64, 96, 100, 200)
    for up, kw in [(2, {}), (4, {}), (2, dict(orientation=6))]:
        d = jxlo.encode_vardct(img, strategy_mode=2, upsampling=up, splines=4, **kw)
        plain = jxlo.encode_vardct(img, strategy_mode=2, upsampling=up, **kw)
        assert (jxlo.decode(d, 3, jxlo.UINT8) != jxlo.decode(plain, 3, jxlo.UINT8)).any()
        for nc, dt in [(3, jxlo.UINT8), (4, jxlo.UINT16), (3, jxlo.FLOAT)]:
            got = emul_lib.decode([d], nc, dt, [(64 * up, 96 * up)])[0]
            assert np.array_equal(got.view(np.uint8), jxlo.decode(d, nc, dt).view(np.uint8)), (up, kw, nc, dt)
