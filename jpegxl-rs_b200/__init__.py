"""jxl_b200: host-side mirror of the jpegxl-rs decode API over the B200 C ABI.

The reference's safe wrapper is Rust (jpegxl-rs/src/decode.rs); there is no Rust
toolchain in this image, so the same surface is mirrored here in Python on top of
``libjxl_b200.so`` (include/jxl_b200.h) through ctypes:

* ``decoder_builder()`` / ``JxlDecoder``   -> jpegxl-rs/src/decode.rs:85-204 (builder fields)
* ``JxlDecoder.decode`` / ``decode_with``  -> jpegxl-rs/src/decode.rs:440-491
* ``Metadata`` / ``Pixels``                -> jpegxl-rs/src/decode/result.rs:25-75
* ``DecodeError`` variants                 -> jpegxl-rs/src/errors.rs:27-60
* ``decode_batch``                         -> new: the batch extension (JxlB200Decoder*)

All pixel work happens in the CUDA kernels; importing this module without the built
library or without a GPU raises (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjxl_b200.so")

# ---- enums (jpegxl-sys/src/common/types.rs:45-103, jpegxl-sys/src/decode.rs:84-247) ----
JXL_TYPE_FLOAT, JXL_TYPE_UINT8, JXL_TYPE_UINT16, JXL_TYPE_FLOAT16 = 0, 2, 3, 5
JXL_NATIVE_ENDIAN, JXL_LITTLE_ENDIAN, JXL_BIG_ENDIAN = 0, 1, 2
JXL_DEC_SUCCESS, JXL_DEC_ERROR, JXL_DEC_NEED_MORE_INPUT = 0, 1, 2
JXL_DEC_NEED_IMAGE_OUT_BUFFER = 5
JXL_DEC_BASIC_INFO, JXL_DEC_FULL_IMAGE = 0x40, 0x1000

_NP_DTYPE = {JXL_TYPE_FLOAT: np.float32, JXL_TYPE_UINT8: np.uint8, JXL_TYPE_UINT16: np.uint16,
             JXL_TYPE_FLOAT16: np.float16}


class JxlPixelFormat(ctypes.Structure):
    _fields_ = [("num_channels", ctypes.c_uint32), ("data_type", ctypes.c_int), ("endianness", ctypes.c_int),
                ("align", ctypes.c_size_t)]


class JxlBasicInfo(ctypes.Structure):
    _fields_ = [("have_container", ctypes.c_int), ("xsize", ctypes.c_uint32), ("ysize", ctypes.c_uint32),
                ("bits_per_sample", ctypes.c_uint32), ("exponent_bits_per_sample", ctypes.c_uint32),
                ("intensity_target", ctypes.c_float), ("min_nits", ctypes.c_float),
                ("relative_to_max_display", ctypes.c_int), ("linear_below", ctypes.c_float),
                ("uses_original_profile", ctypes.c_int), ("have_preview", ctypes.c_int),
                ("have_animation", ctypes.c_int), ("orientation", ctypes.c_uint32),
                ("num_color_channels", ctypes.c_uint32), ("num_extra_channels", ctypes.c_uint32),
                ("alpha_bits", ctypes.c_uint32), ("alpha_exponent_bits", ctypes.c_uint32),
                ("alpha_premultiplied", ctypes.c_int), ("preview_xsize", ctypes.c_uint32),
                ("preview_ysize", ctypes.c_uint32), ("tps_numerator", ctypes.c_uint32),
                ("tps_denominator", ctypes.c_uint32), ("num_loops", ctypes.c_uint32),
                ("have_timecodes", ctypes.c_int), ("intrinsic_xsize", ctypes.c_uint32),
                ("intrinsic_ysize", ctypes.c_uint32), ("padding", ctypes.c_uint8 * 100)]


assert ctypes.sizeof(JxlBasicInfo) == 204  # lib/jxl/decode.cc:2061


class JxlB200Stats(ctypes.Structure):
    _fields_ = [("compressed_bytes", ctypes.c_uint64), ("output_bytes", ctypes.c_uint64), ("pixels", ctypes.c_uint64),
                ("num_streams", ctypes.c_uint64), ("arena_bytes", ctypes.c_uint64),
                ("kernel_launches", ctypes.c_uint32), ("num_ac_streams", ctypes.c_uint32),
                ("vardct_frames", ctypes.c_uint32), ("wave_frames", ctypes.c_uint32)]


class JxlB200EncodeOptions(ctypes.Structure):
    _fields_ = [("distance", ctypes.c_float), ("strategy_mode", ctypes.c_int), ("gaborish", ctypes.c_int),
                ("epf_iters", ctypes.c_uint32), ("dc_smoothing", ctypes.c_int), ("has_alpha", ctypes.c_int)]


class JxlColorEncoding(ctypes.Structure):
    """jpegxl-sys/src/color/color_encoding.rs:125-159."""
    _fields_ = [("color_space", ctypes.c_int), ("white_point", ctypes.c_int), ("white_point_xy", ctypes.c_double * 2),
                ("primaries", ctypes.c_int), ("primaries_red_xy", ctypes.c_double * 2),
                ("primaries_green_xy", ctypes.c_double * 2), ("primaries_blue_xy", ctypes.c_double * 2),
                ("transfer_function", ctypes.c_int), ("gamma", ctypes.c_double), ("rendering_intent", ctypes.c_int)]


class EncodeError(Exception):
    """Mirrors jpegxl-rs/src/errors.rs:62-98."""


# JxlEncoderError -> the EncodeError variant names of jpegxl-rs/src/errors.rs:62-98 (check_enc_status, encode.rs:205-221)
_ENC_ERROR_NAMES = {1: "GenericError", 2: "OutOfMemory", 3: "Jbrd", 4: "BadInput", 0x80: "NotSupported", 0x81: "ApiUsage"}
JXL_ENC_SUCCESS, JXL_ENC_ERROR, JXL_ENC_NEED_MORE_OUTPUT = 0, 1, 2
JXL_ENC_FRAME_SETTING_EFFORT, JXL_ENC_FRAME_SETTING_DECODING_SPEED = 0, 1


class DecodeError(Exception):
    """Mirrors jpegxl-rs/src/errors.rs:27-60."""


class CannotCreateDecoder(DecodeError):
    pass


class GenericError(DecodeError):
    pass


class InvalidInput(DecodeError):
    pass


class UnsupportedBitWidth(DecodeError):
    pass


class NotImplementedFeature(DecodeError):
    pass


_lib = None


def load_library() -> ctypes.CDLL:
    """Loads libjxl_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a) first")
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    lib.JxlB200DecoderCreate.restype = vp
    lib.JxlB200DecoderCreate.argtypes = [ctypes.c_int]
    lib.JxlB200DecoderDestroy.argtypes = [vp]
    lib.JxlB200DecoderGetError.restype = ctypes.c_char_p
    lib.JxlB200DecoderGetError.argtypes = [vp]
    lib.JxlB200DecoderSetInputBatch.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(sz), sz,
                                                ctypes.POINTER(JxlPixelFormat), ctypes.c_int]
    lib.JxlB200DecoderSetKeepOrientation.argtypes = [vp, ctypes.c_int]
    lib.JxlB200DecoderNumFrames.restype = sz
    lib.JxlB200DecoderNumFrames.argtypes = [vp]
    lib.JxlB200DecoderGetBasicInfo.argtypes = [vp, sz, ctypes.POINTER(JxlBasicInfo)]
    lib.JxlB200DecoderImageOutBufferSize.restype = sz
    lib.JxlB200DecoderImageOutBufferSize.argtypes = [vp, sz]
    lib.JxlB200DecoderRun.argtypes = [vp, vp]
    lib.JxlB200DecoderWait.argtypes = [vp, vp]
    lib.JxlB200DecoderDeviceOutput.restype = vp
    lib.JxlB200DecoderDeviceOutput.argtypes = [vp, sz]
    lib.JxlB200DecoderReadOutput.argtypes = [vp, sz, vp, sz]
    lib.JxlB200DecoderReadOutputs.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(sz), sz]
    lib.JxlB200DecoderPlanBatch.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(sz), sz,
                                            ctypes.POINTER(JxlPixelFormat), ctypes.c_int]
    lib.JxlB200DecoderCommitPlan.argtypes = [vp]
    lib.JxlB200DecoderRunToHost.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(sz), sz]
    lib.JxlB200DecoderGetStats.argtypes = [vp, ctypes.POINTER(JxlB200Stats)]
    lib.JxlB200DecoderSetProfiling.argtypes = [vp, ctypes.c_int]
    lib.JxlB200DecoderSetPhaseMask.argtypes = [vp, ctypes.c_uint32]
    lib.JxlB200DecoderDeviceOutputBytes.restype = ctypes.c_size_t
    lib.JxlB200DecoderDeviceOutputBytes.argtypes = [vp]
    lib.JxlB200DecoderGetKernelTimes.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint32)]
    lib.JxlB200DecoderGetKernelTimesEx.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.c_uint32,
                                                   ctypes.POINTER(ctypes.c_uint32)]
    lib.JxlB200EncoderCreate.restype = vp
    lib.JxlB200EncoderCreate.argtypes = [ctypes.c_int]
    lib.JxlB200EncoderDestroy.argtypes = [vp]
    lib.JxlB200EncoderGetError.restype = ctypes.c_char_p
    lib.JxlB200EncoderGetError.argtypes = [vp]
    lib.JxlB200EncoderEncodeBatch.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint32),
                                              ctypes.POINTER(ctypes.c_uint32), sz, ctypes.POINTER(JxlB200EncodeOptions)]
    lib.JxlB200EncoderEncodeLosslessBatch.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint32),
                                                      ctypes.POINTER(ctypes.c_uint32), sz, ctypes.c_uint32, ctypes.c_uint32]
    lib.JxlB200EncoderOutputSize.restype = sz
    lib.JxlB200EncoderOutputSize.argtypes = [vp, sz]
    lib.JxlB200EncoderReadOutput.argtypes = [vp, sz, vp, sz]
    lib.JxlB200EncoderGetPhaseTimes.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    lib.JxlEncoderVersion.restype = ctypes.c_uint32
    lib.JxlEncoderCreate.restype = vp
    lib.JxlEncoderCreate.argtypes = [vp]
    lib.JxlEncoderReset.argtypes = [vp]
    lib.JxlEncoderDestroy.argtypes = [vp]
    lib.JxlEncoderGetError.argtypes = [vp]
    lib.JxlB200EncoderApiMessage.restype = ctypes.c_char_p
    lib.JxlB200EncoderApiMessage.argtypes = [vp]
    lib.JxlEncoderSetParallelRunner.argtypes = [vp, vp, vp]
    lib.JxlEncoderFrameSettingsCreate.restype = vp
    lib.JxlEncoderFrameSettingsCreate.argtypes = [vp, vp]
    lib.JxlEncoderUseContainer.argtypes = [vp, ctypes.c_int]
    lib.JxlEncoderUseBoxes.argtypes = [vp]
    lib.JxlEncoderAddBox.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, sz, ctypes.c_int]
    lib.JxlEncoderStoreJPEGMetadata.argtypes = [vp, ctypes.c_int]
    lib.JxlEncoderSetFrameLossless.argtypes = [vp, ctypes.c_int]
    lib.JxlEncoderSetFrameDistance.argtypes = [vp, ctypes.c_float]
    lib.JxlEncoderDistanceFromQuality.restype = ctypes.c_float
    lib.JxlEncoderDistanceFromQuality.argtypes = [ctypes.c_float]
    lib.JxlEncoderFrameSettingsSetOption.argtypes = [vp, ctypes.c_int, ctypes.c_int64]
    lib.JxlEncoderInitBasicInfo.argtypes = [ctypes.POINTER(JxlBasicInfo)]
    lib.JxlEncoderSetBasicInfo.argtypes = [vp, ctypes.POINTER(JxlBasicInfo)]
    lib.JxlEncoderSetColorEncoding.argtypes = [vp, ctypes.POINTER(JxlColorEncoding)]
    lib.JxlColorEncodingSetToSRGB.argtypes = [ctypes.POINTER(JxlColorEncoding), ctypes.c_int]
    lib.JxlColorEncodingSetToLinearSRGB.argtypes = [ctypes.POINTER(JxlColorEncoding), ctypes.c_int]
    lib.JxlEncoderAddImageFrame.argtypes = [vp, ctypes.POINTER(JxlPixelFormat), vp, sz]
    lib.JxlEncoderAddJPEGFrame.argtypes = [vp, ctypes.c_char_p, sz]
    lib.JxlEncoderCloseInput.argtypes = [vp]
    lib.JxlEncoderProcessOutput.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(sz)]
    lib.JxlThreadParallelRunnerCreate.restype = vp
    lib.JxlThreadParallelRunnerCreate.argtypes = [vp, sz]
    lib.JxlThreadParallelRunnerDestroy.argtypes = [vp]
    lib.JxlThreadParallelRunnerDefaultNumWorkerThreads.restype = sz
    lib.JxlResizableParallelRunnerCreate.restype = vp
    lib.JxlResizableParallelRunnerCreate.argtypes = [vp]
    lib.JxlResizableParallelRunnerSetThreads.argtypes = [vp, sz]
    lib.JxlResizableParallelRunnerSuggestThreads.restype = ctypes.c_uint32
    lib.JxlResizableParallelRunnerSuggestThreads.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
    lib.JxlResizableParallelRunnerDestroy.argtypes = [vp]
    lib.JxlDecoderVersion.restype = ctypes.c_uint32
    lib.JxlSignatureCheck.argtypes = [ctypes.c_char_p, sz]
    lib.JxlDecoderCreate.restype = vp
    lib.JxlDecoderCreate.argtypes = [vp]
    lib.JxlDecoderReset.argtypes = [vp]
    lib.JxlDecoderDestroy.argtypes = [vp]
    lib.JxlDecoderSetParallelRunner.argtypes = [vp, vp, vp]
    lib.JxlDecoderSubscribeEvents.argtypes = [vp, ctypes.c_int]
    lib.JxlDecoderSetKeepOrientation.argtypes = [vp, ctypes.c_int]
    lib.JxlDecoderSetUnpremultiplyAlpha.argtypes = [vp, ctypes.c_int]
    lib.JxlDecoderSetRenderSpotcolors.argtypes = [vp, ctypes.c_int]
    lib.JxlDecoderSetCoalescing.argtypes = [vp, ctypes.c_int]
    lib.JxlDecoderSetDesiredIntensityTarget.argtypes = [vp, ctypes.c_float]
    lib.JxlDecoderSetInput.argtypes = [vp, ctypes.c_char_p, sz]
    lib.JxlDecoderCloseInput.argtypes = [vp]
    lib.JxlDecoderProcessInput.argtypes = [vp]
    lib.JxlDecoderGetBasicInfo.argtypes = [vp, ctypes.POINTER(JxlBasicInfo)]
    lib.JxlDecoderImageOutBufferSize.argtypes = [vp, ctypes.POINTER(JxlPixelFormat), ctypes.POINTER(sz)]
    lib.JxlDecoderSetImageOutBuffer.argtypes = [vp, ctypes.POINTER(JxlPixelFormat), vp, sz]
    _lib = lib
    return lib


# Every symbol include/jxl_b200.h declares (checked by the CPU test-suite).
EXPORTED_SYMBOLS = [
    "JxlB200DecoderCreate", "JxlB200DecoderDestroy", "JxlB200DecoderGetError", "JxlB200DecoderSetInputBatch",
    "JxlB200DecoderSetKeepOrientation", "JxlB200DecoderNumFrames", "JxlB200DecoderGetBasicInfo", "JxlB200DecoderImageOutBufferSize", "JxlB200DecoderRun",
    "JxlB200DecoderWait", "JxlB200DecoderDeviceOutput", "JxlB200DecoderReadOutput", "JxlB200DecoderReadOutputs",
    "JxlB200DecoderPlanBatch", "JxlB200DecoderCommitPlan", "JxlB200DecoderRunToHost",
    "JxlB200DecoderGetStats", "JxlB200DecoderSetProfiling", "JxlB200DecoderSetPhaseMask", "JxlB200DecoderDeviceOutputBytes", "JxlB200DecoderGetKernelTimes",
    "JxlB200DecoderGetKernelTimesEx", "JxlB200EncoderCreate", "JxlB200EncoderDestroy", "JxlB200EncoderGetError",
    "JxlB200EncoderEncodeBatch", "JxlB200EncoderEncodeLosslessBatch", "JxlB200EncoderOutputSize", "JxlB200EncoderReadOutput", "JxlB200EncoderGetPhaseTimes",
    "JxlDecoderVersion", "JxlSignatureCheck", "JxlDecoderCreate", "JxlDecoderReset",
    "JxlDecoderDestroy", "JxlDecoderSetParallelRunner", "JxlDecoderSubscribeEvents", "JxlDecoderSetKeepOrientation",
    "JxlDecoderSetUnpremultiplyAlpha", "JxlDecoderSetRenderSpotcolors", "JxlDecoderSetCoalescing",
    "JxlDecoderSetDesiredIntensityTarget", "JxlDecoderSetInput", "JxlDecoderCloseInput", "JxlDecoderProcessInput",
    "JxlDecoderGetBasicInfo", "JxlDecoderImageOutBufferSize", "JxlDecoderSetImageOutBuffer",
    "JxlDecoderGetICCProfileSize", "JxlDecoderGetColorAsICCProfile", "JxlDecoderSetJPEGBuffer", "JxlDecoderReleaseJPEGBuffer",
    # the encoder symbols jpegxl-rs calls (jpegxl-rs/src/encode.rs:156-467)
    "JxlEncoderVersion", "JxlEncoderCreate", "JxlEncoderReset", "JxlEncoderDestroy", "JxlEncoderGetError",
    "JxlEncoderSetParallelRunner", "JxlEncoderFrameSettingsCreate", "JxlEncoderUseContainer", "JxlEncoderUseBoxes",
    "JxlEncoderAddBox", "JxlEncoderStoreJPEGMetadata", "JxlEncoderSetFrameLossless", "JxlEncoderSetFrameDistance",
    "JxlEncoderDistanceFromQuality", "JxlEncoderFrameSettingsSetOption", "JxlEncoderInitBasicInfo",
    "JxlEncoderSetBasicInfo", "JxlEncoderSetColorEncoding", "JxlColorEncodingSetToSRGB", "JxlColorEncodingSetToLinearSRGB",
    "JxlEncoderAddImageFrame", "JxlEncoderAddJPEGFrame", "JxlEncoderCloseInput", "JxlEncoderProcessOutput",
    "JxlB200EncoderApiMessage",
    # libjxl_threads (jpegxl-sys/src/threads)
    "JxlThreadParallelRunner", "JxlThreadParallelRunnerCreate", "JxlThreadParallelRunnerDestroy",
    "JxlThreadParallelRunnerDefaultNumWorkerThreads", "JxlResizableParallelRunner", "JxlResizableParallelRunnerCreate",
    "JxlResizableParallelRunnerSetThreads", "JxlResizableParallelRunnerSuggestThreads", "JxlResizableParallelRunnerDestroy",
]


@dataclass
class PixelFormat:
    """jpegxl-rs/src/decode.rs:52-83. num_channels == 0 means colour (+ alpha if present)."""
    num_channels: int = 0
    endianness: int = JXL_NATIVE_ENDIAN
    align: int = 0


@dataclass
class Metadata:
    """jpegxl-rs/src/decode/result.rs:25-49."""
    width: int
    height: int
    intensity_target: float
    min_nits: float
    orientation: int
    num_color_channels: int
    has_alpha_channel: bool
    intrinsic_width: int
    intrinsic_height: int
    icc_profile: Optional[bytes] = None


@dataclass
class Pixels:
    """jpegxl-rs/src/decode/result.rs:51-75: the variant is the numpy dtype."""
    data: np.ndarray
    data_type: int

    @property
    def variant(self) -> str:
        return {JXL_TYPE_FLOAT: "Float", JXL_TYPE_UINT8: "Uint8", JXL_TYPE_UINT16: "Uint16",
                JXL_TYPE_FLOAT16: "Float16"}[self.data_type]

    def __len__(self) -> int:
        return int(self.data.size)


def _metadata(info: JxlBasicInfo) -> Metadata:
    return Metadata(info.xsize, info.ysize, info.intensity_target, info.min_nits, info.orientation,
                    info.num_color_channels, info.alpha_bits > 0, info.intrinsic_xsize, info.intrinsic_ysize)


def check_valid_signature(data: bytes) -> Optional[bool]:
    """jpegxl-rs/src/utils.rs:24-33."""
    sig = load_library().JxlSignatureCheck(bytes(data), len(data))
    if sig == 0:
        return None
    return sig in (2, 3)


def _default_data_type(info: JxlBasicInfo) -> int:
    """jpegxl-rs/src/decode.rs:394-404."""
    bits, exp = info.bits_per_sample, info.exponent_bits_per_sample
    if exp == 0 and bits <= 8:
        return JXL_TYPE_UINT8
    if exp == 0 and bits <= 16:
        return JXL_TYPE_UINT16
    if bits == 16:
        return JXL_TYPE_FLOAT16
    if bits == 32:
        return JXL_TYPE_FLOAT
    raise UnsupportedBitWidth(bits)


class JxlDecoder:
    """Event-loop decoder over the libjxl-compatible subset (one image per call)."""

    def __init__(self, pixel_format: Optional[PixelFormat] = None, skip_reorientation: Optional[bool] = None,
                 unpremul_alpha: Optional[bool] = None, render_spotcolors: Optional[bool] = None,
                 coalescing: Optional[bool] = None, desired_intensity_target: Optional[float] = None,
                 decompress: Optional[bool] = None, icc_profile: bool = False, parallel_runner=None):
        self._lib = load_library()
        self._dec = self._lib.JxlDecoderCreate(None)
        if not self._dec:
            raise CannotCreateDecoder()
        self.pixel_format = pixel_format
        self.skip_reorientation = skip_reorientation
        self.unpremul_alpha = unpremul_alpha
        self.render_spotcolors = render_spotcolors
        self.coalescing = coalescing
        self.desired_intensity_target = desired_intensity_target
        self.decompress = decompress
        self.icc_profile = icc_profile
        self.parallel_runner = parallel_runner

    def __del__(self):
        if getattr(self, "_dec", None):
            self._lib.JxlDecoderDestroy(self._dec)
            self._dec = None

    def _check(self, status: int) -> None:
        if status != JXL_DEC_SUCCESS:
            raise GenericError(f"decoder status {status}")

    def _setup(self) -> None:  # jpegxl-rs/src/decode.rs:327-366
        lib, dec = self._lib, self._dec
        self._check(lib.JxlDecoderSetParallelRunner(dec, None, None))
        self._check(lib.JxlDecoderSubscribeEvents(dec, JXL_DEC_BASIC_INFO | JXL_DEC_FULL_IMAGE))
        if self.skip_reorientation is not None:
            self._check(lib.JxlDecoderSetKeepOrientation(dec, int(self.skip_reorientation)))
        if self.unpremul_alpha is not None:
            self._check(lib.JxlDecoderSetUnpremultiplyAlpha(dec, int(self.unpremul_alpha)))
        if self.render_spotcolors is not None:
            self._check(lib.JxlDecoderSetRenderSpotcolors(dec, int(self.render_spotcolors)))
        if self.coalescing is not None:
            self._check(lib.JxlDecoderSetCoalescing(dec, int(self.coalescing)))
        if self.desired_intensity_target is not None:
            self._check(lib.JxlDecoderSetDesiredIntensityTarget(dec, float(self.desired_intensity_target)))

    def _decode_internal(self, data: bytes, data_type: Optional[int]) -> Tuple[Metadata, Pixels]:
        valid = check_valid_signature(data)
        if not valid:
            raise InvalidInput()
        lib, dec = self._lib, self._dec
        self._setup()
        data = bytes(data)
        self._check(lib.JxlDecoderSetInput(dec, data, len(data)))
        lib.JxlDecoderCloseInput(dec)
        info = JxlBasicInfo()
        fmt = JxlPixelFormat()
        buf = None
        try:
            while True:
                status = lib.JxlDecoderProcessInput(dec)
                if status in (JXL_DEC_NEED_MORE_INPUT, JXL_DEC_ERROR):
                    raise GenericError()
                if status == JXL_DEC_BASIC_INFO:
                    self._check(lib.JxlDecoderGetBasicInfo(dec, ctypes.byref(info)))
                elif status == JXL_DEC_NEED_IMAGE_OUT_BUFFER:
                    dt = data_type if data_type is not None else _default_data_type(info)
                    f = self.pixel_format or PixelFormat()
                    fmt.num_channels = f.num_channels or (info.num_color_channels + (1 if info.alpha_bits > 0 else 0))
                    fmt.data_type, fmt.endianness, fmt.align = dt, f.endianness, f.align
                    size = ctypes.c_size_t(0)
                    self._check(lib.JxlDecoderImageOutBufferSize(dec, ctypes.byref(fmt), ctypes.byref(size)))
                    buf = np.zeros(size.value, dtype=np.uint8)
                    self._check(lib.JxlDecoderSetImageOutBuffer(dec, ctypes.byref(fmt), buf.ctypes.data, size.value))
                elif status == JXL_DEC_FULL_IMAGE:
                    pass
                elif status == JXL_DEC_SUCCESS:
                    break
                else:
                    raise NotImplementedFeature(f"event {status:#x}")
        finally:
            lib.JxlDecoderReset(dec)
        pixels = buf.view(_NP_DTYPE[fmt.data_type])
        if fmt.endianness == JXL_BIG_ENDIAN and pixels.dtype.itemsize > 1:
            pixels = pixels.byteswap()  # jpegxl-rs/src/common.rs:37-125 converts to native values
        return _metadata(info), Pixels(pixels, fmt.data_type)

    def decode(self, data: bytes) -> Tuple[Metadata, Pixels]:
        """jpegxl-rs/src/decode.rs:440-455."""
        return self._decode_internal(data, None)

    def decode_with(self, data: bytes, dtype) -> Tuple[Metadata, np.ndarray]:
        """jpegxl-rs/src/decode.rs:457-491; dtype in {u8, u16, f16, f32} as a numpy dtype."""
        dt = {np.dtype(np.uint8): JXL_TYPE_UINT8, np.dtype(np.uint16): JXL_TYPE_UINT16,
              np.dtype(np.float16): JXL_TYPE_FLOAT16, np.dtype(np.float32): JXL_TYPE_FLOAT}[np.dtype(dtype)]
        meta, px = self._decode_internal(data, dt)
        return meta, px.data


def decoder_builder(**kwargs):
    """jpegxl-rs/src/lib.rs:36-42: `decoder_builder().…build()`."""

    class _Builder:
        def __init__(self, kw):
            self._kw = dict(kw)

        def __getattr__(self, name):
            def setter(value):
                self._kw[name] = value
                return self
            return setter

        def build(self) -> JxlDecoder:
            return JxlDecoder(**self._kw)

    return _Builder(kwargs)


class BatchDecoder:
    """The batch extension: n independent codestreams -> one set of kernel launches."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        self._dec = self._lib.JxlB200DecoderCreate(device)
        if not self._dec:
            raise CannotCreateDecoder("no usable CUDA device (jxl_b200 has no CPU fallback)")
        self._keep = None

    def __del__(self):
        if getattr(self, "_dec", None):
            self._lib.JxlB200DecoderDestroy(self._dec)
            self._dec = None

    def _err(self) -> str:
        return self._lib.JxlB200DecoderGetError(self._dec).decode()

    def set_input(self, files: Sequence[bytes], num_channels: int = 4, data_type: int = JXL_TYPE_UINT8,
                  endianness: int = JXL_NATIVE_ENDIAN, align: int = 0, threads: int = 0, keep_orientation: bool = False) -> None:
        """keep_orientation: leave the images as coded (jpegxl-rs `skip_reorientation`); default: turned upright."""
        self._lib.JxlB200DecoderSetKeepOrientation(self._dec, int(keep_orientation))
        n = len(files)
        bufs = [np.frombuffer(f, dtype=np.uint8) for f in files]
        ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (ctypes.c_size_t * n)(*[b.size for b in bufs])
        fmt = JxlPixelFormat(num_channels, data_type, endianness, align)
        if threads <= 0:
            threads = min(os.cpu_count() or 1, 32)
        rc = self._lib.JxlB200DecoderSetInputBatch(self._dec, ptrs, sizes, n, ctypes.byref(fmt), threads)
        if rc != 0:
            raise GenericError(self._err())
        self.format = fmt
        self.num_frames = n

    def plan(self, files: Sequence[bytes], num_channels: int = 4, data_type: int = JXL_TYPE_UINT8,
             endianness: int = JXL_NATIVE_ENDIAN, align: int = 0, threads: int = 0, keep_orientation: bool = False) -> None:
        """The host half of set_input (parse into a pending plan): may run while this handle's kernels are in flight.
        commit() uploads it (after wait() of the batch before)."""
        self._lib.JxlB200DecoderSetKeepOrientation(self._dec, int(keep_orientation))
        n = len(files)
        bufs = [np.frombuffer(f, dtype=np.uint8) for f in files]
        ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (ctypes.c_size_t * n)(*[b.size for b in bufs])
        fmt = JxlPixelFormat(num_channels, data_type, endianness, align)
        if threads <= 0:
            threads = min(os.cpu_count() or 1, 32)
        if self._lib.JxlB200DecoderPlanBatch(self._dec, ptrs, sizes, n, ctypes.byref(fmt), threads) != 0:
            raise GenericError(self._err())
        self._pending = (fmt, n)

    def commit(self) -> None:
        if self._lib.JxlB200DecoderCommitPlan(self._dec) != 0:
            raise GenericError(self._err())
        self.format, self.num_frames = self._pending

    def run(self, cuda_stream: int = 0) -> None:
        if self._lib.JxlB200DecoderRun(self._dec, ctypes.c_void_p(cuda_stream)) != 0:
            raise GenericError(self._err())

    def run_to_host(self, outs: Sequence[np.ndarray], cuda_stream: int = 0) -> None:
        """run(), with frame i copied into outs[i] (pinned memory for the copies to overlap the kernels) as soon as it
        is rendered; complete when wait() returns. The arrays must stay alive until then."""
        n = len(outs)
        key = (id(outs), n)
        if getattr(self, "_outs_key", None) != key:  # (the same list step after step: the pointer arrays are kept)
            self._outs_ptrs = (ctypes.c_void_p * n)(*[o.ctypes.data for o in outs])
            self._outs_sizes = (ctypes.c_size_t * n)(*[o.nbytes for o in outs])
            self._outs_key, self._outs_ref = key, outs
        ptrs, sizes = self._outs_ptrs, self._outs_sizes
        if self._lib.JxlB200DecoderRunToHost(self._dec, ctypes.c_void_p(cuda_stream), ptrs, sizes, n) != 0:
            raise GenericError(self._err())

    def wait(self, cuda_stream: int = 0) -> None:
        if self._lib.JxlB200DecoderWait(self._dec, ctypes.c_void_p(cuda_stream)) != 0:
            raise GenericError(self._err())

    def basic_info(self, i: int) -> JxlBasicInfo:
        info = JxlBasicInfo()
        if self._lib.JxlB200DecoderGetBasicInfo(self._dec, i, ctypes.byref(info)) != 0:
            raise GenericError("bad frame index")
        return info

    def out_size(self, i: int) -> int:
        return self._lib.JxlB200DecoderImageOutBufferSize(self._dec, i)

    def device_output(self, i: int) -> int:
        return self._lib.JxlB200DecoderDeviceOutput(self._dec, i)

    def device_output_bytes(self) -> int:
        return self._lib.JxlB200DecoderDeviceOutputBytes(self._dec)

    def device_output_tensor(self):
        """The batch's whole output buffer in HBM as a torch uint8 CUDA tensor (no copy): frames back to back, each
        256-byte aligned. What the multi-GPU gather sends."""
        import torch

        class _Buf:
            pass

        n = self._lib.JxlB200DecoderDeviceOutputBytes(self._dec)
        b = _Buf()
        b.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (self.device_output(0), False), "version": 2}
        return torch.as_tensor(b, device="cuda")

    def read_output(self, i: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        n = self.out_size(i)
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        if self._lib.JxlB200DecoderReadOutput(self._dec, i, out.ctypes.data, out.nbytes) != 0:
            raise GenericError(self._err())
        return out

    def read_outputs(self, outs: Sequence[np.ndarray]) -> None:
        n = len(outs)
        ptrs = (ctypes.c_void_p * n)(*[o.ctypes.data for o in outs])
        sizes = (ctypes.c_size_t * n)(*[o.nbytes for o in outs])
        if self._lib.JxlB200DecoderReadOutputs(self._dec, ptrs, sizes, n) != 0:
            raise GenericError(self._err())

    def set_phase_mask(self, classes=None) -> None:
        """Profiling hook: run() launches only the named kernel classes (None = all)."""
        mask = 0xFFFFFFFF if classes is None else sum(1 << self.KERNEL_CLASSES.index(c) for c in classes)
        self._lib.JxlB200DecoderSetPhaseMask(self._dec, mask)

    def set_profiling(self, enabled: bool) -> None:
        self._lib.JxlB200DecoderSetProfiling(self._dec, int(enabled))

    def kernel_times(self):
        """(ms per kernel accumulated [decode, group transforms, global transforms, output], runs)."""
        ms = (ctypes.c_double * 4)()
        runs = ctypes.c_uint32(0)
        self._lib.JxlB200DecoderGetKernelTimes(self._dec, ms, ctypes.byref(runs))
        return list(ms), runs.value

    KERNEL_CLASSES = ["modular_decode", "group_programs", "frame_levels", "write_output", "dc_finish", "ac_decode",
                      "dequant_idct", "filters", "color_write"]

    def kernel_times_ex(self):
        """({kernel class: accumulated ms}, runs) over all kernel classes (include/jxl_b200.h)."""
        n = len(self.KERNEL_CLASSES)
        ms = (ctypes.c_double * n)()
        runs = ctypes.c_uint32(0)
        self._lib.JxlB200DecoderGetKernelTimesEx(self._dec, ms, n, ctypes.byref(runs))
        return dict(zip(self.KERNEL_CLASSES, list(ms))), runs.value

    def stats(self) -> JxlB200Stats:
        st = JxlB200Stats()
        self._lib.JxlB200DecoderGetStats(self._dec, ctypes.byref(st))
        return st


def decode_batch(files: Sequence[bytes], num_channels: int = 4, dtype=np.uint8, device: int = 0,
                 keep_orientation: bool = False) -> List[np.ndarray]:
    """Decodes n files on the GPU and returns one (H, W, C) array per file."""
    dt = {np.dtype(np.uint8): JXL_TYPE_UINT8, np.dtype(np.uint16): JXL_TYPE_UINT16,
          np.dtype(np.float16): JXL_TYPE_FLOAT16, np.dtype(np.float32): JXL_TYPE_FLOAT}[np.dtype(dtype)]
    dec = BatchDecoder(device)
    dec.set_input(files, num_channels, dt, keep_orientation=keep_orientation)
    dec.run()
    dec.wait()
    out = []
    for i in range(len(files)):
        info = dec.basic_info(i)
        raw = dec.read_output(i)
        out.append(raw.view(dtype).reshape(info.ysize, info.xsize, num_channels))
    return out


# ---- multi-GPU: frames are independent, frame i -> rank i mod R ----
def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """Indices of the frames rank `rank` of `world` decodes: round-robin, so every rank gets a similar mix."""
    return list(range(rank, n, world))


def gather_frames(local: Sequence[np.ndarray], n_total: int, rank: int, world: int, dst: int = 0, device=None):
    """The one collective of the path: gathers the decoded frames of every rank on rank `dst` (returns the list in
    the original order there, None elsewhere). Works on any torch.distributed backend (NCCL with `device` a CUDA
    device, gloo on the CPU): sizes are exchanged first, then one padded gather of bytes."""
    import torch
    import torch.distributed as dist
    dev = device if device is not None else torch.device("cpu")
    meta = [(a.shape, a.dtype.str) for a in local]
    metas = [None] * world
    dist.all_gather_object(metas, meta)
    sizes = [sum(int(np.prod(sh)) * np.dtype(dt).itemsize for sh, dt in m) for m in metas]
    cap = max(max(sizes), 1)
    flat = np.zeros(cap, np.uint8)
    off = 0
    for a in local:
        b = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        flat[off:off + b.size] = b
        off += b.size
    mine = torch.from_numpy(flat).to(dev)
    bufs = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(mine, bufs, dst=dst)
    if rank != dst:
        return None
    out = [None] * n_total
    for r in range(world):
        raw = bufs[r].cpu().numpy()
        off = 0
        for i, (sh, dt) in zip(shard_indices(n_total, r, world), metas[r]):
            nbytes = int(np.prod(sh)) * np.dtype(dt).itemsize
            out[i] = raw[off:off + nbytes].view(np.dtype(dt)).reshape(sh).copy()
            off += nbytes
    return out


def decode_batch_distributed(files: Sequence[bytes], num_channels: int = 4, dtype=np.uint8, dst: int = 0, decode_fn=None,
                             device=None):
    """Every rank decodes its shard of `files` (all ranks pass the same list); rank `dst` gets all frames.
    `decode_fn(files, num_channels, dtype)` defaults to decode_batch on this rank's GPU."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    idx = shard_indices(len(files), rank, world)
    fn = decode_fn or decode_batch
    local = fn([files[i] for i in idx], num_channels, dtype) if idx else []
    return gather_frames(local, len(files), rank, world, dst, device)


# ---- encoder: mirror of jpegxl-rs's JxlEncoder builder for the lossy RGB8 path ----
@dataclass
class EncoderResult:
    """jpegxl-rs/src/encode.rs:505-509."""
    data: bytes


class JxlEncoder:
    """jpegxl-rs/src/encode.rs:60-187 (builder fields) and :477-486 (`encode::<u8, u8>`): lossy VarDCT of RGB8 input on
    the GPU. `quality` is the Butteraugli distance as in jpegxl-rs (default 1.0); `speed` selects the AcStrategy
    search: the fastest tiers use 8x8 DCTs only, the others the variance heuristic over 8x8 ... 64x64.

    `encode()` drives the libjxl-compatible JxlEncoder* entry points in the order jpegxl-rs does
    (setup_encoder -> add_frame -> CloseInput -> ProcessOutput with a doubling buffer, encode.rs:225-378);
    `encode_batch()` is the batch extension (one call, many images). What the CUDA encoder does not cover
    (lossless, alpha, non-8-bit samples, JPEG transcoding, boxes) fails with the reference's error variants."""

    def __init__(self, has_alpha: bool = False, lossless: bool = False, speed: int = 7, quality: float = 1.0,
                 use_container: bool = False, uses_original_profile: bool = False, decoding_speed: int = 0,
                 init_buffer_size: int = 512 * 1024, device: int = 0):
        self._lib = load_library()
        self._handle = self._lib.JxlEncoderCreate(None)
        if not self._handle:
            raise EncodeError("CannotCreateEncoder")
        self._options = self._lib.JxlEncoderFrameSettingsCreate(self._handle, None)
        self._enc = None  # batch encoder, created on first use
        self.device = device
        self.has_alpha, self.lossless, self.speed, self.quality = has_alpha, lossless, speed, quality
        self.use_container, self.uses_original_profile = use_container, uses_original_profile
        self.decoding_speed, self.init_buffer_size = decoding_speed, max(32, init_buffer_size)

    def __del__(self):
        if getattr(self, "_enc", None):
            self._lib.JxlB200EncoderDestroy(self._enc)
            self._enc = None
        if getattr(self, "_handle", None):
            self._lib.JxlEncoderDestroy(self._handle)
            self._handle = None

    # -- jpegxl-rs/src/encode.rs:205-221
    def _check(self, status: int):
        if status == JXL_ENC_SUCCESS:
            return
        if status == JXL_ENC_NEED_MORE_OUTPUT:
            raise EncodeError("NeedMoreOutput")
        code = self._lib.JxlEncoderGetError(self._handle)
        raise EncodeError("%s: %s" % (_ENC_ERROR_NAMES.get(code, "GenericError"),
                                      self._lib.JxlB200EncoderApiMessage(self._handle).decode()))

    # -- jpegxl-rs/src/encode.rs:225-254
    def _set_options(self):
        L = self._lib
        self._check(L.JxlEncoderUseContainer(self._handle, int(self.use_container)))
        if self.lossless is not None:
            self._check(L.JxlEncoderSetFrameLossless(self._options, int(self.lossless)))
        self._check(L.JxlEncoderFrameSettingsSetOption(self._options, JXL_ENC_FRAME_SETTING_EFFORT, int(self.speed)))
        self._check(L.JxlEncoderSetFrameDistance(self._options, float(self.quality)))
        self._check(L.JxlEncoderFrameSettingsSetOption(self._options, JXL_ENC_FRAME_SETTING_DECODING_SPEED,
                                                       int(self.decoding_speed)))

    # -- jpegxl-rs/src/encode.rs:257-319
    def _setup_encoder(self, width: int, height: int, bits: int, exp: int, has_alpha: bool):
        L = self._lib
        self._set_options()
        info = JxlBasicInfo()
        L.JxlEncoderInitBasicInfo(ctypes.byref(info))
        info.xsize, info.ysize = width, height
        info.have_container = int(self.use_container)
        info.uses_original_profile = int(self.uses_original_profile)
        info.bits_per_sample, info.exponent_bits_per_sample = bits, exp
        if has_alpha:
            info.num_extra_channels, info.alpha_bits, info.alpha_exponent_bits = 1, bits, exp
        self._check(L.JxlEncoderSetBasicInfo(self._handle, ctypes.byref(info)))

    # -- jpegxl-rs/src/encode.rs:345-378
    def _internal(self) -> bytes:
        L = self._lib
        L.JxlEncoderCloseInput(self._handle)
        buf = (ctypes.c_uint8 * self.init_buffer_size)()
        written = 0
        while True:
            next_out = ctypes.c_void_p(ctypes.addressof(buf) + written)
            avail = ctypes.c_size_t(len(buf) - written)
            status = L.JxlEncoderProcessOutput(self._handle, ctypes.byref(next_out), ctypes.byref(avail))
            written = next_out.value - ctypes.addressof(buf)
            if status != JXL_ENC_NEED_MORE_OUTPUT:
                break
            bigger = (ctypes.c_uint8 * (len(buf) * 2))()
            ctypes.memmove(bigger, buf, written)
            buf = bigger
        try:
            self._check(status)
        finally:
            L.JxlEncoderReset(self._handle)
            self._options = L.JxlEncoderFrameSettingsCreate(self._handle, None)
        return bytes(buf[:written])

    def _batch_options(self) -> JxlB200EncodeOptions:
        return JxlB200EncodeOptions(float(self.quality), 0 if self.speed <= 2 else 2, 1, 2, 1)

    def encode_batch(self, images: Sequence[np.ndarray], gaborish: bool = True, epf_iters: int = 2) -> List[EncoderResult]:
        """One JxlB200EncoderEncodeBatch call over independent RGB8 images. `gaborish` / `epf_iters`: the loop-filter
        fields of the frame header (libjxl's encoder derives them from the distance; the defaults are its d = 1 values)."""
        if self.lossless:
            return self.encode_lossless_batch(images)
        if self.use_container:
            raise EncodeError("NotSupported: container output in batch mode")
        if self._enc is None:
            self._enc = self._lib.JxlB200EncoderCreate(self.device)
            if not self._enc:
                raise EncodeError("CannotCreateEncoder: no usable CUDA device (jxl_b200 has no CPU fallback)")
        imgs = [np.ascontiguousarray(a, np.uint8) for a in images]
        nch = 4 if self.has_alpha else 3
        for a in imgs:
            if a.ndim != 3 or a.shape[2] != nch:
                raise EncodeError("ApiUsage: expected (height, width, %d) uint8 arrays" % nch)
        n = len(imgs)
        ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in imgs])
        xs = (ctypes.c_uint32 * n)(*[a.shape[1] for a in imgs])
        ys = (ctypes.c_uint32 * n)(*[a.shape[0] for a in imgs])
        opt = self._batch_options()
        opt.gaborish = 1 if gaborish else 0
        opt.epf_iters = int(epf_iters)
        opt.has_alpha = 1 if self.has_alpha else 0
        if self._lib.JxlB200EncoderEncodeBatch(self._enc, ptrs, xs, ys, n, ctypes.byref(opt)) != 0:
            raise EncodeError(self._lib.JxlB200EncoderGetError(self._enc).decode())
        out = []
        for i in range(n):
            size = self._lib.JxlB200EncoderOutputSize(self._enc, i)
            buf = np.empty(size, np.uint8)
            if self._lib.JxlB200EncoderReadOutput(self._enc, i, buf.ctypes.data, size) != 0:
                raise EncodeError("internal: output read failed")
            out.append(EncoderResult(buf.tobytes()))
        return out

    def encode_lossless_batch(self, images: Sequence[np.ndarray]) -> List[EncoderResult]:
        """One JxlB200EncoderEncodeLosslessBatch call: (H, W, C) uint8 or uint16 arrays of one dtype and channel count
        (C = 1 grey, 2 grey + alpha, 3 RGB, 4 RGBA). The decoded images equal the inputs bit for bit."""
        if self._enc is None:
            self._enc = self._lib.JxlB200EncoderCreate(self.device)
            if not self._enc:
                raise EncodeError("CannotCreateEncoder: no usable CUDA device (jxl_b200 has no CPU fallback)")
        imgs = [np.ascontiguousarray(a) for a in images]
        dt, nch = imgs[0].dtype, imgs[0].shape[2] if imgs[0].ndim == 3 else 0
        for a in imgs:
            if a.ndim != 3 or a.dtype != dt or a.shape[2] != nch or dt not in (np.uint8, np.uint16) or not 1 <= nch <= 4:
                raise EncodeError("ApiUsage: expected (height, width, channels) uint8 / uint16 arrays of one kind")
        n = len(imgs)
        ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in imgs])
        xs = (ctypes.c_uint32 * n)(*[a.shape[1] for a in imgs])
        ys = (ctypes.c_uint32 * n)(*[a.shape[0] for a in imgs])
        if self._lib.JxlB200EncoderEncodeLosslessBatch(self._enc, ptrs, xs, ys, n, nch, 8 * dt.itemsize) != 0:
            raise EncodeError(self._lib.JxlB200EncoderGetError(self._enc).decode())
        out = []
        for i in range(n):
            size = self._lib.JxlB200EncoderOutputSize(self._enc, i)
            buf = np.empty(size, np.uint8)
            if self._lib.JxlB200EncoderReadOutput(self._enc, i, buf.ctypes.data, size) != 0:
                raise EncodeError("internal: output read failed")
            out.append(EncoderResult(buf.tobytes()))
        return out

    def encode(self, data: np.ndarray, width: Optional[int] = None, height: Optional[int] = None) -> EncoderResult:
        """jpegxl-rs/src/encode.rs:477-486: `data` is height * width * channels samples (or an (H, W, C) array) of
        u8 / u16 / f32; channels = 3, + 1 with `has_alpha`."""
        a = np.asarray(data)
        if a.dtype not in (np.uint8, np.uint16, np.float32, np.float16):
            a = a.astype(np.uint8)
        channels = 4 if self.has_alpha else 3
        if a.ndim == 1:
            if not width or not height or a.size != width * height * channels:
                raise EncodeError("ApiUsage: buffer size does not match width * height * channels")
            a = a.reshape(height, width, channels)
        a = np.ascontiguousarray(a)
        height, width = a.shape[:2]
        bits, exp, dt = {np.dtype(np.uint8): (8, 0, JXL_TYPE_UINT8), np.dtype(np.uint16): (16, 0, JXL_TYPE_UINT16),
                         np.dtype(np.float16): (16, 5, JXL_TYPE_FLOAT16), np.dtype(np.float32): (32, 8, JXL_TYPE_FLOAT)}[a.dtype]
        self._setup_encoder(width, height, bits, exp, self.has_alpha)
        fmt = JxlPixelFormat(a.shape[2], dt, JXL_NATIVE_ENDIAN, 0)
        try:
            self._check(self._lib.JxlEncoderAddImageFrame(self._options, ctypes.byref(fmt), a.ctypes.data, a.nbytes))
        except EncodeError:
            self._lib.JxlEncoderReset(self._handle)
            self._options = self._lib.JxlEncoderFrameSettingsCreate(self._handle, None)
            raise
        return EncoderResult(self._internal())

    def encode_jpeg(self, data: bytes) -> EncoderResult:
        """jpegxl-rs/src/encode.rs:444-466. JPEG transcoding is not built: NotSupported, as the C API reports."""
        self._set_options()
        self._check(self._lib.JxlEncoderStoreJPEGMetadata(self._handle, 1))
        self._check(self._lib.JxlEncoderAddJPEGFrame(self._options, data, len(data)))
        return EncoderResult(self._internal())

    def phase_times(self):
        ms = (ctypes.c_double * 3)()
        if self._enc is None:
            return {"tokens_ms": 0.0, "host_tables_ms": 0.0, "emit_ms": 0.0}
        self._lib.JxlB200EncoderGetPhaseTimes(self._enc, ms)
        return {"tokens_ms": ms[0], "host_tables_ms": ms[1], "emit_ms": ms[2]}


def encoder_builder(**kwargs):
    """jpegxl-rs/src/lib.rs:36-42: `encoder_builder().quality(1.0).build()`."""

    class _Builder:
        def __init__(self, kw):
            self._kw = dict(kw)

        def __getattr__(self, name):
            def setter(value):
                self._kw[name] = value
                return self
            return setter

        def build(self) -> JxlEncoder:
            return JxlEncoder(**self._kw)

    return _Builder(kwargs)
