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
