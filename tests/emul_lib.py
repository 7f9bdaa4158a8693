"""ctypes wrapper of tests/emul/libjxlb_emul.so: the kernels' __host__ __device__ bodies run on the CPU.
Test infrastructure only (logic check without a GPU)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_NP = {0: np.float32, 2: np.uint8, 3: np.uint16, 5: np.float16}


def lib():
    global _lib
    if _lib is None:
        d = os.path.join(HERE, "emul")
        subprocess.check_call(["make", "-s"], cwd=d)
        _lib = ctypes.CDLL(os.path.join(d, "libjxlb_emul.so"))
    return _lib


class EmulError(Exception):
    pass


def decode(files, num_channels, data_type, shapes, endianness=0, align=0):
    """shapes: list of (h, w). Returns one array per file."""
    n = len(files)
    arr = (ctypes.c_char_p * n)(*files)
    sizes = (ctypes.c_size_t * n)(*[len(f) for f in files])
    bps = np.dtype(_NP[data_type]).itemsize
    cap = sum(((h * w * num_channels * bps + 255) // 256 + 1) * 256 for h, w in shapes) + 4096
    out = np.zeros(cap, np.uint8)
    offs = (ctypes.c_uint64 * n)()
    err = ctypes.create_string_buffer(512)
    rc = lib().jxlb_emul_decode(arr, sizes, ctypes.c_size_t(n), num_channels, data_type, endianness,
                                ctypes.c_size_t(align), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(cap),
                                offs, err, ctypes.c_size_t(512))
    if rc != 0:
        raise EmulError(err.value.decode())
    res = []
    for i, (h, w) in enumerate(shapes):
        nb = h * w * num_channels * bps
        res.append(out[offs[i]:offs[i] + nb].view(_NP[data_type]).reshape(h, w, num_channels))
    return res


def last_plan_stats():
    """(channels, channels on the weighted-predictor LUT path, channels on the (y, N, W) table path) of the last decode."""
    v = (ctypes.c_uint64 * 3)()
    lib().jxlb_emul_last_plan_stats(v)
    return tuple(int(x) for x in v)


def last_coop_streams():
    """(streams decoded one per warp by k_modular_decode_coop, all Modular streams) of the last decode."""
    v = (ctypes.c_uint64 * 2)()
    lib().jxlb_emul_last_coop_streams(v)
    return tuple(int(x) for x in v)


def encode(rgb, distance=1.0, strategy_mode=2, gab=True, epf_iters=2, dc_smoothing=True) -> bytes:
    """RGB8 (H, W, 3) or RGBA8 (H, W, 4: alpha as a lossless extra channel) -> codestream, the encoder kernels' device
    functions run on the CPU."""
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, nch = rgb.shape
    dc_smoothing = int(bool(dc_smoothing)) | (2 if nch == 4 else 0)
    cap = h * w * 6 + (1 << 20)
    out = np.zeros(cap, np.uint8)
    err = ctypes.create_string_buffer(512)
    L = lib()
    L.jxlb_emul_encode.restype = ctypes.c_long
    n = L.jxlb_emul_encode(rgb.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(w), ctypes.c_uint32(h),
                           ctypes.c_float(distance), ctypes.c_int(strategy_mode), ctypes.c_int(int(gab)),
                           ctypes.c_uint32(epf_iters), ctypes.c_int(int(dc_smoothing)),
                           out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(cap), err, ctypes.c_size_t(512))
    if n < 0:
        raise EmulError(err.value.decode())
    return out[:n].tobytes()


def encode_lossless(img) -> bytes:
    """(H, W, C) uint8 / uint16 samples (C = 1 grey, 2 grey + alpha, 3 RGB, 4 RGBA) -> lossless codestream, the lossless
    encoder's device functions run on the CPU."""
    img = np.ascontiguousarray(img)
    assert img.dtype in (np.uint8, np.uint16) and img.ndim == 3
    h, w, c = img.shape
    cap = img.nbytes * 2 + (1 << 20)
    out = np.zeros(cap, np.uint8)
    err = ctypes.create_string_buffer(512)
    L = lib()
    L.jxlb_emul_encode_lossless.restype = ctypes.c_long
    n = L.jxlb_emul_encode_lossless(img.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(w), ctypes.c_uint32(h), ctypes.c_uint32(c),
                                    ctypes.c_uint32(8 * img.dtype.itemsize), out.ctypes.data_as(ctypes.c_void_p),
                                    ctypes.c_size_t(cap), err, ctypes.c_size_t(512))
    if n < 0:
        raise EmulError(err.value.decode())
    return out[:n].tobytes()
