"""CPU: the C-ABI library loads and exports every symbol include/jxl_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "jxl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(Jxl(?:B200)?[A-Za-z]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/jxl_b200.h but not exported"
    assert sorted(pkg.EXPORTED_SYMBOLS) == syms


def test_version_and_signature(pkg):
    lib = pkg.load_library()
    assert lib.JxlDecoderVersion() == 11002  # jpegxl-sys/src/lib.rs:77-83
    assert lib.JxlSignatureCheck(b"\xff\x0a", 2) == 2
    assert lib.JxlSignatureCheck(b"\0\0\0\x0cJXL \r\n\x87\n", 12) == 3
    assert lib.JxlSignatureCheck(b"\0\0\0\0", 4) == 1
    assert lib.JxlSignatureCheck(b"", 0) == 0
    assert pkg.check_valid_signature(b"") is None
    assert pkg.check_valid_signature(b"\0" * 64) is False


def test_basic_info_layout(pkg):
    assert ctypes.sizeof(pkg.JxlBasicInfo) == 204
    assert ctypes.sizeof(pkg.JxlPixelFormat) == 24


def test_no_silent_cpu_fallback_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.CannotCreateDecoder):
        pkg.BatchDecoder(0)
    dec = pkg.decoder_builder().build()
    data = open(os.path.join(ROOT, "tests", "golden", "sample.jxl"), "rb").read()
    with pytest.raises(pkg.DecodeError):
        dec.decode(data)


def test_invalid_input_errors(pkg):
    dec = pkg.decoder_builder().build()
    for bad in [b"", b"\0" * 64]:
        with pytest.raises(pkg.InvalidInput):
            dec.decode(bad)


# ---- the libjxl-compatible encoder entry points (no pixels are processed without a GPU)
def test_encoder_api_protocol_and_error_codes(pkg):
    import numpy as np
    lib = pkg.load_library()
    assert lib.JxlEncoderVersion() == 11002
    assert ctypes.sizeof(pkg.JxlColorEncoding) == 104
    # lib/jxl/encode.cc:1547-1552
    assert lib.JxlEncoderDistanceFromQuality(100.0) == 0.0
    assert abs(lib.JxlEncoderDistanceFromQuality(90.0) - 1.0) < 1e-6
    assert abs(lib.JxlEncoderDistanceFromQuality(10.0) - (53.0 / 30 - 11.5 + 25.0)) < 1e-5
    info = pkg.JxlBasicInfo()
    lib.JxlEncoderInitBasicInfo(ctypes.byref(info))
    assert (info.bits_per_sample, info.num_color_channels, info.orientation, info.tps_numerator) == (8, 3, 1, 10)
    col = pkg.JxlColorEncoding()
    lib.JxlColorEncodingSetToLinearSRGB(ctypes.byref(col), 0)
    assert (col.color_space, col.white_point, col.primaries, col.transfer_function) == (0, 1, 1, 8)
    enc = lib.JxlEncoderCreate(None)
    fs = lib.JxlEncoderFrameSettingsCreate(enc, None)
    fmt = pkg.JxlPixelFormat(3, 2, 0, 0)
    px = np.zeros((4, 4, 3), np.uint8)
    # no basic info yet -> API usage error (lib/jxl/encode.cc:2291-2295)
    assert lib.JxlEncoderAddImageFrame(fs, ctypes.byref(fmt), px.ctypes.data, px.nbytes) == 1
    assert lib.JxlEncoderGetError(enc) == 0x81
    assert lib.JxlEncoderSetFrameDistance(fs, 30.0) == 1 and lib.JxlEncoderGetError(enc) == 0x81
    assert lib.JxlEncoderFrameSettingsSetOption(fs, 0, 11) == 1  # effort 11 needs expert options
    assert lib.JxlEncoderFrameSettingsSetOption(fs, 0, 7) == 0 and lib.JxlEncoderFrameSettingsSetOption(fs, 1, 4) == 0
    info.xsize = info.ysize = 4
    assert lib.JxlEncoderSetBasicInfo(enc, ctypes.byref(info)) == 0
    # 16-bit input, alpha, lossless, JPEG, boxes: outside the CUDA encoder -> NotSupported
    fmt16 = pkg.JxlPixelFormat(3, 3, 0, 0)
    assert lib.JxlEncoderAddImageFrame(fs, ctypes.byref(fmt16), px.ctypes.data, px.nbytes * 2) == 1
    assert lib.JxlEncoderGetError(enc) == 0x80
    assert lib.JxlEncoderAddJPEGFrame(fs, b"\xff\xd8", 2) == 1 and lib.JxlEncoderGetError(enc) == 0x80
    assert lib.JxlEncoderAddBox(enc, b"Exif", b"abcd", 4, 0) == 1 and lib.JxlEncoderGetError(enc) == 0x81
    assert lib.JxlEncoderUseBoxes(enc) == 0
    assert lib.JxlEncoderAddBox(enc, b"Exif", b"abcd", 4, 0) == 1 and lib.JxlEncoderGetError(enc) == 0x80
    # too small a buffer
    assert lib.JxlEncoderAddImageFrame(fs, ctypes.byref(fmt), px.ctypes.data, 10) == 1
    assert lib.JxlEncoderGetError(enc) == 0x81
    assert lib.JxlEncoderAddImageFrame(fs, ctypes.byref(fmt), px.ctypes.data, px.nbytes) == 0
    lib.JxlEncoderReset(enc)
    lib.JxlEncoderDestroy(enc)


def test_encoder_mirror_reports_the_reference_error_variants(pkg):
    import numpy as np
    img = np.zeros((8, 8, 3), np.uint8)
    with pytest.raises(pkg.EncodeError, match="ApiUsage"):  # lossless needs uses_original_profile (encode.cc:1476-1484)
        pkg.encoder_builder().lossless(True).build().encode(img)
    with pytest.raises(pkg.EncodeError, match="NotSupported"):  # lossless takes integer samples
        pkg.encoder_builder().lossless(True).uses_original_profile(True).build().encode(np.zeros((8, 8, 3), np.float32))
    with pytest.raises(pkg.EncodeError, match="NotSupported"):
        pkg.encoder_builder().build().encode(np.zeros((8, 8, 3), np.uint16))
    with pytest.raises(pkg.EncodeError, match="NotSupported"):  # lossy alpha is 8-bit
        pkg.encoder_builder().has_alpha(True).build().encode(np.zeros((8, 8, 4), np.uint16))
    with pytest.raises(pkg.EncodeError, match="NotSupported"):
        pkg.encoder_builder().build().encode_jpeg(b"\xff\xd8\xff\xd9")
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(pkg.EncodeError, match="GenericError.*no usable CUDA device"):
            pkg.encoder_builder().build().encode(img)
        with pytest.raises(pkg.EncodeError, match="GenericError.*no usable CUDA device"):  # the lossless path: no CPU fallback either
            pkg.encoder_builder().lossless(True).uses_original_profile(True).build().encode(img)


def test_thread_runner_symbols_behave(pkg):
    lib = pkg.load_library()
    init_t = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t)
    func_t = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_size_t)
    seen, inits = [], []
    init = init_t(lambda opaque, n: inits.append(n) or 0)
    func = func_t(lambda opaque, value, thread: seen.append((value, thread)))
    for create, run, destroy in [(lambda: lib.JxlThreadParallelRunnerCreate(None, 4), lib.JxlThreadParallelRunner,
                                  lib.JxlThreadParallelRunnerDestroy),
                                 (lambda: lib.JxlResizableParallelRunnerCreate(None), lib.JxlResizableParallelRunner,
                                  lib.JxlResizableParallelRunnerDestroy)]:
        run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, init_t, func_t, ctypes.c_uint32, ctypes.c_uint32]
        r = create()
        seen.clear()
        assert run(r, None, init, func, 3, 8) == 0
        assert [v for v, _ in seen] == [3, 4, 5, 6, 7]
        assert run(r, None, init, func, 5, 5) == 0 and run(r, None, init, func, 6, 5) != 0
        destroy(r)
    assert lib.JxlThreadParallelRunnerDefaultNumWorkerThreads() >= 1
    assert lib.JxlResizableParallelRunnerSuggestThreads(256, 256) == 1


def test_color_encoding_event_and_memory_manager(pkg):
    # jpegxl-rs with `icc_profile` subscribes to COLOR_ENCODING and then asks for the ICC profile
    # (jpegxl-rs/src/decode.rs:334-347): the event arrives after BASIC_INFO, the ICC calls fail loudly. Header parsing
    # only: runs without a GPU. A memory manager is accepted (jpegxl-rs/src/memory.rs:24-40).
    import ctypes
    lib = pkg.load_library()
    data = open(os.path.join(ROOT, "tests", "golden", "sample.jxl"), "rb").read()
    lib.JxlDecoderCreate.restype = ctypes.c_void_p
    lib.JxlDecoderCreate.argtypes = [ctypes.c_void_p]
    mm = (ctypes.c_void_p * 3)()
    dec = ctypes.c_void_p(lib.JxlDecoderCreate(ctypes.addressof(mm)))
    assert dec.value
    try:
        assert lib.JxlDecoderSubscribeEvents(dec, pkg.JXL_DEC_BASIC_INFO | 0x100 | pkg.JXL_DEC_FULL_IMAGE) == 0
        assert lib.JxlDecoderSetInput(dec, data, len(data)) == 0
        lib.JxlDecoderCloseInput(dec)
        assert lib.JxlDecoderProcessInput(dec) == pkg.JXL_DEC_BASIC_INFO
        assert lib.JxlDecoderProcessInput(dec) == 0x100  # JXL_DEC_COLOR_ENCODING
        size = ctypes.c_size_t(7)
        assert lib.JxlDecoderGetICCProfileSize(dec, 1, ctypes.byref(size)) == pkg.JXL_DEC_ERROR and size.value == 0
        assert lib.JxlDecoderProcessInput(dec) == pkg.JXL_DEC_NEED_IMAGE_OUT_BUFFER
    finally:
        lib.JxlDecoderDestroy(dec)
