"""ctypes wrapper of the oracle (oracle/libjxlo.so). Checker only: tests, smoke(), bench cpu_baseline."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLOAT, UINT8, UINT16, FLOAT16 = 0, 2, 3, 5
_NP = {FLOAT: np.float32, UINT8: np.uint8, UINT16: np.uint16, FLOAT16: np.float16}


class Info(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in
                "xsize ysize bits exp num_color num_extra alpha_bits xyb orientation num_frames".split()]


_lib = None
_fast = False


def use_fast_build(on=True):
    """Timing arms of bench.py only: load oracle/libjxlo_fast.so (-O3 -march=x86-64-v3, bit-identical output) instead of
    the -O2 checker build. Must be called before the first use."""
    global _fast, _lib
    _fast = bool(on)
    _lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "libjxlo_fast.so" if _fast else "libjxlo.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-s"], cwd=os.path.join(ROOT, "oracle"))
        L = ctypes.CDLL(path)
        L.jxlo_decode.restype = ctypes.c_void_p
        L.jxlo_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
        L.jxlo_output_size.restype = ctypes.c_size_t
        L.jxlo_output_size.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_size_t]
        L.jxlo_write_pixels.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                        ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        L.jxlo_frame_info.restype = ctypes.c_char_p
        L.jxlo_frame_info.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.jxlo_get_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(Info)]
        L.jxlo_free.argtypes = [ctypes.c_void_p]
        L.jxlo_encode_vardct.restype = ctypes.c_size_t
        L.jxlo_encode_vardct.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, ctypes.c_int,
                                         ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_uint32, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
        L.jxlo_encoded_copy.argtypes = [ctypes.c_void_p]
        L.jxlo_encode_modular.restype = ctypes.c_size_t
        L.jxlo_encode_modular.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_char_p,
                                          ctypes.c_size_t]
        _lib = L
    return _lib


class OracleError(Exception):
    pass


class Decoded:
    def __init__(self, data: bytes):
        L = lib()
        err = ctypes.create_string_buffer(512)
        self.h = L.jxlo_decode(data, len(data), err, 512)
        if not self.h:
            raise OracleError(err.value.decode())
        self.info = Info()
        L.jxlo_get_info(self.h, ctypes.byref(self.info))

    def frame_info(self):
        return [lib().jxlo_frame_info(self.h, i).decode() for i in range(self.info.num_frames)]

    def pixels(self, num_channels, data_type, endianness=0, align=0, raw=False, undo_orientation=False):
        """undo_orientation: libjxl's default output (JxlDecoderSetKeepOrientation false): the image turned upright."""
        L = lib()
        n = L.jxlo_output_size(self.h, num_channels, data_type, align)
        h, w = self.info.ysize, self.info.xsize
        if undo_orientation and self.info.orientation >= 5:
            assert align <= 1
            h, w = w, h
        buf = np.zeros(n, np.uint8)
        rc = L.jxlo_write_pixels(self.h, num_channels, data_type, endianness | (0x400 if undo_orientation else 0), align,
                                 buf.ctypes.data, n)
        assert rc == 0
        if raw or align > 1:
            return buf
        return buf.view(_NP[data_type]).reshape(h, w, num_channels)

    def __del__(self):
        if getattr(self, "h", None):
            lib().jxlo_free(self.h)
            self.h = None


def decode(data: bytes, num_channels: int, data_type: int, endianness: int = 0, undo_orientation: bool = False) -> np.ndarray:
    return Decoded(data).pixels(num_channels, data_type, endianness, undo_orientation=undo_orientation)


def encode_vardct(rgb: np.ndarray, distance=1.0, strategy_mode=2, seed=1, gab=True, epf_iters=2, dc_smoothing=True,
                  random_side_info=False, num_passes=1, dc_tree=0, inverse_gaborish=True, coeff_orders=True, cfl=True,
                  adaptive_quant=True, prefix_codes=False, upsampling=1, orientation=1, splines=0) -> bytes:
    """RGB8 (H, W, 3) -> a VarDCT codestream written by the oracle's plain encoder (stream generator). (H, W, 4): the
    fourth channel travels as a lossless 8-bit alpha extra channel in the frame's Modular sub-streams."""
    L = lib()
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, c = rgb.shape
    alpha = None
    if c == 4:
        alpha = np.ascontiguousarray(rgb[:, :, 3])
        rgb = np.ascontiguousarray(rgb[:, :, :3])
        c = 3
        L.jxlo_set_next_alpha(alpha.ctypes.data_as(ctypes.c_void_p))
    assert c == 3
    err = ctypes.create_string_buffer(512)
    gab_arg = int(bool(gab)) | (0 if inverse_gaborish else 2) | (0 if coeff_orders else 4) | (0 if cfl else 8) | (0 if adaptive_quant else 16) | (32 if prefix_codes else 0) | ({1: 0, 2: 1, 4: 2, 8: 3}[upsampling] << 6) | ((orientation - 1) << 8) | (splines << 11)
    n = L.jxlo_encode_vardct(rgb.ctypes.data, w, h, distance, strategy_mode, seed, gab_arg, epf_iters,
                             int(dc_smoothing), int(random_side_info), num_passes, dc_tree, err, 512)
    if n == 0:
        raise OracleError(err.value.decode())
    out = np.zeros(n, np.uint8)
    L.jxlo_encoded_copy(out.ctypes.data)
    return out.tobytes()


def encode_modular(img: np.ndarray, bits=8, alpha=False, group_size_shift=1, tree=0, predictor=5, seed=1, rct=-1,
                   palette_colors=0, palette_deltas=0, palette_predictor=0, squeeze=False, prefix=False, lz77=False,
                   lz77_min_symbol=224, orientation=1, splines=0) -> bytes:
    """(H, W, C) integer samples -> a lossless Modular codestream written by the oracle's plain encoder
    (oracle/jxlo_enc_modular.h). C = 1 / 3 colour channels (+ 1 when alpha)."""
    L = lib()
    img = np.ascontiguousarray(img, np.uint16)
    h, w, c = img.shape
    num_color = c - (1 if alpha else 0)
    assert num_color in (1, 3)
    params = np.array([bits, num_color, int(alpha), group_size_shift, tree, predictor, seed, rct + 1, palette_colors,
                       palette_deltas, palette_predictor, int(squeeze), int(prefix) | (int(lz77) << 1), lz77_min_symbol, orientation, splines],
                      np.uint32)
    err = ctypes.create_string_buffer(512)
    n = L.jxlo_encode_modular(img.ctypes.data, w, h, params.ctypes.data, err, 512)
    if n == 0:
        raise OracleError(err.value.decode())
    out = np.zeros(n, np.uint8)
    L.jxlo_encoded_copy(out.ctypes.data)
    return out.tobytes()
