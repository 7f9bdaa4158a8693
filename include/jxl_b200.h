/* jxl_b200 -- C ABI of the B200-native JPEG XL hot path.
 *
 * Drop-in boundary: the symbols below are what jpegxl-sys would bind for the
 * decode/encode path (jpegxl-sys/src/decode.rs:363-1532 declares libjxl's
 * JxlDecoder* functions as `extern "C-unwind"`; jpegxl-rs drives them from
 * jpegxl-rs/src/decode.rs:207-325). Two groups of entry points:
 *
 *  1. JxlB200* batch API (new): one call decodes a batch of independent
 *     codestreams on the GPU. A single 256x256 group is a serial entropy-coded
 *     chain, so a B200 is filled by #groups x #frames, not by one image.
 *  2. Jxl* libjxl-compatible subset (same names, argument meaning, status codes and
 *     event order as libjxl 0.11.2, lib/include/jxl/decode.h) so that jpegxl-rs's
 *     event loop runs unchanged; internally a batch of one.
 *
 * Plain pointers and sizes only. All functions return 0 (JXL_DEC_SUCCESS) on
 * success unless stated otherwise. The library needs a CUDA device: without one
 * every compute entry point fails (there is no CPU fallback).
 */
#ifndef JXL_B200_H_
#define JXL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- types shared with libjxl (jpegxl-sys/src/common/types.rs:25-148) ---- */
typedef int JXL_BOOL;
typedef enum { JXL_TYPE_FLOAT = 0, JXL_TYPE_UINT8 = 2, JXL_TYPE_UINT16 = 3, JXL_TYPE_FLOAT16 = 5 } JxlDataType;
typedef enum { JXL_NATIVE_ENDIAN = 0, JXL_LITTLE_ENDIAN = 1, JXL_BIG_ENDIAN = 2 } JxlEndianness;
typedef struct {
  uint32_t num_channels;
  JxlDataType data_type;
  JxlEndianness endianness;
  size_t align;
} JxlPixelFormat;

/* JxlBasicInfo, byte-identical to libjxl 0.11.2 (lib/include/jxl/codestream_header.h;
 * jpegxl-sys/src/metadata/codestream_header.rs:108-238; 204 bytes, lib/jxl/decode.cc:2061). */
typedef struct { uint32_t xsize, ysize; } JxlPreviewHeader;
typedef struct { uint32_t tps_numerator, tps_denominator, num_loops; JXL_BOOL have_timecodes; } JxlAnimationHeader;
typedef struct {
  JXL_BOOL have_container;
  uint32_t xsize;
  uint32_t ysize;
  uint32_t bits_per_sample;
  uint32_t exponent_bits_per_sample;
  float intensity_target;
  float min_nits;
  JXL_BOOL relative_to_max_display;
  float linear_below;
  JXL_BOOL uses_original_profile;
  JXL_BOOL have_preview;
  JXL_BOOL have_animation;
  uint32_t orientation; /* JxlOrientation */
  uint32_t num_color_channels;
  uint32_t num_extra_channels;
  uint32_t alpha_bits;
  uint32_t alpha_exponent_bits;
  JXL_BOOL alpha_premultiplied;
  JxlPreviewHeader preview;
  JxlAnimationHeader animation;
  uint32_t intrinsic_xsize;
  uint32_t intrinsic_ysize;
  uint8_t padding[100];
} JxlBasicInfo;

/* ---- 1. batch API ---- */
typedef struct JxlB200Decoder JxlB200Decoder;

/* Creates a batch decoder bound to CUDA device `device`. NULL if no usable GPU. */
JxlB200Decoder* JxlB200DecoderCreate(int device);
void JxlB200DecoderDestroy(JxlB200Decoder* dec);
/* Message of the last failure on this handle (never NULL). */
const char* JxlB200DecoderGetError(const JxlB200Decoder* dec);

/* Host parse of `n` files (headers, TOC, histograms, MA trees, group headers) on
 * `num_threads` host threads, then upload of bitstreams and tables to HBM.
 * The files are only read during the call. Replaces libjxl's
 * JxlDecoderSetInput + the header part of JxlDecoderProcessInput
 * (jpegxl-rs/src/decode.rs:231-252). */
int JxlB200DecoderSetInputBatch(JxlB200Decoder* dec, const uint8_t* const* files, const size_t* sizes, size_t n,
                                const JxlPixelFormat* format, int num_threads);
/* For the batches set afterwards: 1 = leave the images as coded (JxlDecoderSetKeepOrientation); default 0 = libjxl's
 * default, the output is turned upright (lib/jxl/render_pipeline/stage_write.cc:271-288) and GetBasicInfo reports the
 * upright size with orientation 1 (lib/jxl/decode.cc:2083-2090). */
int JxlB200DecoderSetKeepOrientation(JxlB200Decoder* dec, int keep);
size_t JxlB200DecoderNumFrames(const JxlB200Decoder* dec);
int JxlB200DecoderGetBasicInfo(const JxlB200Decoder* dec, size_t i, JxlBasicInfo* info);
/* Bytes of frame i in the requested pixel format (JxlDecoderImageOutBufferSize). */
size_t JxlB200DecoderImageOutBufferSize(const JxlB200Decoder* dec, size_t i);

/* Runs the decode kernels for the whole batch on `cuda_stream` (a cudaStream_t
 * passed as void*, NULL = the decoder's own stream). Asynchronous: returns after
 * the launches. The pixels stay in HBM. */
int JxlB200DecoderRun(JxlB200Decoder* dec, void* cuda_stream);
/* Waits for the last Run and reports per-stream decode errors. */
int JxlB200DecoderWait(JxlB200Decoder* dec, void* cuda_stream);
/* Device pointer (HBM) of frame i's pixels after Run. */
void* JxlB200DecoderDeviceOutput(const JxlB200Decoder* dec, size_t i);
/* Size in bytes of the batch's whole output buffer in HBM: frames back to back from JxlB200DecoderDeviceOutput(dec, 0),
 * each 256-byte aligned (what a caller hands to a collective, e.g. ncclGather, without a host round trip). */
size_t JxlB200DecoderDeviceOutputBytes(const JxlB200Decoder* dec);
/* Device -> host copy of frame i into `dst` (JxlDecoderSetImageOutBuffer semantics:
 * rows of align_up(xsize * channels * bytes, align) bytes). Synchronous. */
int JxlB200DecoderReadOutput(JxlB200Decoder* dec, size_t i, void* dst, size_t size);
/* Device -> host copy of all frames; dsts[i] receives frame i. */
int JxlB200DecoderReadOutputs(JxlB200Decoder* dec, void* const* dsts, const size_t* sizes, size_t n);

/* Streaming use of one handle (the decode loop of a server: batch k + 1 is parsed while batch k decodes, and batch k's
 * pixels leave for the host while its later frames are still being rendered):
 *   PlanBatch   = the host half of SetInputBatch (parse on `num_threads` threads into a pending plan; no device work
 *                 except the probe rounds of single-section frames) -- may run while the handle's kernels are in flight;
 *   CommitPlan  = the device half (upload of the pending plan's bitstreams and tables; call after Wait of the batch
 *                 before). SetInputBatch is PlanBatch + CommitPlan;
 *   RunToHost   = Run, with frame i copied into dsts[i] (pinned host memory for the copies to overlap) on a copy stream
 *                 as soon as the kernels that write it are through; the buffers are complete when Wait returns.
 * The same role as feeding libjxl's decoder from one thread while another consumes JXL_DEC_FULL_IMAGE
 * (jpegxl-rs/src/decode.rs:231-327 runs them back to back on one thread). */
int JxlB200DecoderPlanBatch(JxlB200Decoder* dec, const uint8_t* const* files, const size_t* sizes, size_t n,
                            const JxlPixelFormat* format, int num_threads);
int JxlB200DecoderCommitPlan(JxlB200Decoder* dec);
int JxlB200DecoderRunToHost(JxlB200Decoder* dec, void* cuda_stream, void* const* dsts, const size_t* sizes, size_t n);

/* Workload figures for benchmarks. */
typedef struct {
  uint64_t compressed_bytes; /* bitstream bytes resident in HBM */
  uint64_t output_bytes;     /* pixels written per Run */
  uint64_t pixels;           /* sum of xsize * ysize */
  uint64_t num_streams;      /* entropy-coded streams = decode-kernel threads */
  uint64_t arena_bytes;      /* intermediate int32 planes */
  uint32_t kernel_launches;  /* launches per Run */
  uint32_t num_ac_streams;   /* VarDCT (frame, group, pass) coefficient streams among num_streams */
  uint32_t vardct_frames;    /* lossy frames in the batch */
  uint32_t wave_frames;      /* lossy frames whose pixel planes are live at once */
} JxlB200Stats;
int JxlB200DecoderGetStats(const JxlB200Decoder* dec, JxlB200Stats* stats);
/* Per-kernel device time: when enabled, Run brackets each kernel with CUDA events on the
 * launching stream (the timed region of a benchmark). ms4 = accumulated milliseconds of
 * {entropy decode, group transforms, global transforms, output write}; runs = number of Runs. */
int JxlB200DecoderSetProfiling(JxlB200Decoder* dec, int enabled);
/* Profiling hook: restricts the following Run calls to the kernel classes whose bit is set (bit k = class k of
 * JxlB200DecoderGetKernelTimesEx); the others keep the results of the last full Run. 0xFFFFFFFF (default) = all.
 * Used by tools/interference.py to time one class next to another; never by the decode path proper. */
int JxlB200DecoderSetPhaseMask(JxlB200Decoder* dec, uint32_t mask);
int JxlB200DecoderGetKernelTimes(JxlB200Decoder* dec, double* ms4, uint32_t* runs);
/* All kernel classes: ms[0..n) = {Modular entropy decode, group transforms, global transforms,
 * Modular output write, VarDCT DC finish, AC entropy decode, dequant + inverse transforms,
 * Gaborish + EPF, colour + output write}; n <= 9. */
#define JXL_B200_NUM_KERNEL_CLASSES 9
int JxlB200DecoderGetKernelTimesEx(JxlB200Decoder* dec, double* ms, uint32_t n, uint32_t* runs);

/* ---- 2. libjxl-compatible subset (decode) ---- */
typedef struct JxlDecoderStruct JxlDecoder;
typedef enum {
  JXL_DEC_SUCCESS = 0, JXL_DEC_ERROR = 1, JXL_DEC_NEED_MORE_INPUT = 2, JXL_DEC_NEED_PREVIEW_OUT_BUFFER = 3,
  JXL_DEC_NEED_IMAGE_OUT_BUFFER = 5, JXL_DEC_JPEG_NEED_MORE_OUTPUT = 6, JXL_DEC_BOX_NEED_MORE_OUTPUT = 7,
  JXL_DEC_BASIC_INFO = 0x40, JXL_DEC_COLOR_ENCODING = 0x100, JXL_DEC_PREVIEW_IMAGE = 0x200, JXL_DEC_FRAME = 0x400,
  JXL_DEC_FULL_IMAGE = 0x1000, JXL_DEC_JPEG_RECONSTRUCTION = 0x2000, JXL_DEC_BOX = 0x4000,
  JXL_DEC_FRAME_PROGRESSION = 0x8000, JXL_DEC_BOX_COMPLETE = 0x10000
} JxlDecoderStatus;
typedef enum { JXL_SIG_NOT_ENOUGH_BYTES = 0, JXL_SIG_INVALID = 1, JXL_SIG_CODESTREAM = 2, JXL_SIG_CONTAINER = 3 } JxlSignature;

uint32_t JxlDecoderVersion(void);
JxlSignature JxlSignatureCheck(const uint8_t* buf, size_t len);
/* memory_manager: accepted, its callbacks are never invoked (jpegxl-rs may pass one, jpegxl-rs/src/memory.rs:24-40; the
 * GPU path owns device memory and pinned staging). Events: BASIC_INFO, then COLOR_ENCODING when subscribed (the ICC
 * calls fail: ICC synthesis is not built), NEED_IMAGE_OUT_BUFFER, FULL_IMAGE, SUCCESS. */
JxlDecoder* JxlDecoderCreate(const void* memory_manager);
void JxlDecoderReset(JxlDecoder* dec);
void JxlDecoderDestroy(JxlDecoder* dec);
/* Accepted and ignored: the GPU path does not use host thread pools for pixels. */
JxlDecoderStatus JxlDecoderSetParallelRunner(JxlDecoder* dec, void* parallel_runner, void* parallel_runner_opaque);
JxlDecoderStatus JxlDecoderSubscribeEvents(JxlDecoder* dec, int events_wanted);
JxlDecoderStatus JxlDecoderSetKeepOrientation(JxlDecoder* dec, JXL_BOOL skip_reorientation);
JxlDecoderStatus JxlDecoderSetUnpremultiplyAlpha(JxlDecoder* dec, JXL_BOOL unpremul_alpha);
JxlDecoderStatus JxlDecoderSetRenderSpotcolors(JxlDecoder* dec, JXL_BOOL render_spotcolors);
JxlDecoderStatus JxlDecoderSetCoalescing(JxlDecoder* dec, JXL_BOOL coalescing);
JxlDecoderStatus JxlDecoderSetDesiredIntensityTarget(JxlDecoder* dec, float desired_intensity_target);
JxlDecoderStatus JxlDecoderSetInput(JxlDecoder* dec, const uint8_t* data, size_t size);
void JxlDecoderCloseInput(JxlDecoder* dec);
JxlDecoderStatus JxlDecoderProcessInput(JxlDecoder* dec);
JxlDecoderStatus JxlDecoderGetBasicInfo(const JxlDecoder* dec, JxlBasicInfo* info);
JxlDecoderStatus JxlDecoderImageOutBufferSize(const JxlDecoder* dec, const JxlPixelFormat* format, size_t* size);
JxlDecoderStatus JxlDecoderSetImageOutBuffer(JxlDecoder* dec, const JxlPixelFormat* format, void* buffer, size_t size);
/* Declared because jpegxl-rs binds them (jpegxl-rs/src/decode.rs:260-283, :329-357); ICC synthesis and JPEG
 * reconstruction are not built: these return JXL_DEC_ERROR / 0. `target` is JxlColorProfileTarget. */
JxlDecoderStatus JxlDecoderGetICCProfileSize(const JxlDecoder* dec, int target, size_t* size);
JxlDecoderStatus JxlDecoderGetColorAsICCProfile(const JxlDecoder* dec, int target, uint8_t* icc_profile, size_t size);
JxlDecoderStatus JxlDecoderSetJPEGBuffer(JxlDecoder* dec, uint8_t* data, size_t size);
size_t JxlDecoderReleaseJPEGBuffer(JxlDecoder* dec);

/* ---- 3. batch encoder (lossy VarDCT) ----
 * Replaces what jpegxl-rs reaches through JxlEncoderAddImageFrame + JxlEncoderProcessOutput
 * (jpegxl-rs/src/encode.rs:323-378) for RGB8 input: one call encodes a batch of independent images. */
typedef struct JxlB200Encoder JxlB200Encoder;
typedef struct {
  float distance;      /* JxlEncoderSetFrameDistance: 1.0 = visually lossless target */
  int strategy_mode;   /* 0: 8x8 DCT only; 2: variance heuristic over 8x8 ... 64x64 */
  int gaborish;        /* loop-filter flags written to the frame header (decoder side filters) */
  uint32_t epf_iters;
  int dc_smoothing;
  int has_alpha;       /* 1: the input is interleaved RGBA8; alpha becomes an 8-bit extra channel, coded losslessly */
} JxlB200EncodeOptions;
JxlB200Encoder* JxlB200EncoderCreate(int device);
void JxlB200EncoderDestroy(JxlB200Encoder* enc);
const char* JxlB200EncoderGetError(const JxlB200Encoder* enc);
/* rgb[i]: xsizes[i] * ysizes[i] interleaved RGB8 (options->has_alpha: RGBA8) samples (sRGB). Synchronous. */
int JxlB200EncoderEncodeBatch(JxlB200Encoder* enc, const uint8_t* const* rgb, const uint32_t* xsizes, const uint32_t* ysizes,
                              size_t n, const JxlB200EncodeOptions* options);
/* Lossless (Modular) batch -- jpegxl-rs `lossless(true)` (jpegxl-rs/src/encode.rs:143, :230-234; libjxl:
 * JxlEncoderSetFrameLossless, lib/jxl/enc_modular.cc): pixels[i] = xsizes[i] * ysizes[i] interleaved samples of
 * num_channels each (1 grey, 2 grey + alpha, 3 RGB, 4 RGBA; sRGB), bits_per_sample 8 (uint8) or 16 (uint16, native
 * endian). YCoCg-R + libjxl's fixed gradient tree, groups of 128 x 128 (group_size_shift 0); the decoded image equals the input bit for bit. */
int JxlB200EncoderEncodeLosslessBatch(JxlB200Encoder* enc, const void* const* pixels, const uint32_t* xsizes, const uint32_t* ysizes,
                                      size_t n, uint32_t num_channels, uint32_t bits_per_sample);
size_t JxlB200EncoderOutputSize(const JxlB200Encoder* enc, size_t i);
int JxlB200EncoderReadOutput(const JxlB200Encoder* enc, size_t i, uint8_t* dst, size_t size);
/* Device time of the last EncodeBatch: {pixels -> tokens + histograms, host tables (incl. D2H), rANS emission} in ms. */
int JxlB200EncoderGetPhaseTimes(const JxlB200Encoder* enc, double* ms3);

/* ---- 4. libjxl-compatible subset (encode) ----
 * The 23 encoder symbols jpegxl-rs calls (jpegxl-rs/src/encode.rs:156-467, jpegxl-rs/src/encode/options.rs:46-70;
 * declared in jpegxl-sys/src/encoder/encode.rs), same names, argument meaning, status and error codes as libjxl
 * 0.11.2 (lib/include/jxl/encode.h, lib/jxl/encode.cc); internally a batch of one on the CUDA encoder. What the CUDA
 * encoder does not cover returns JXL_ENC_ERROR with JxlEncoderGetError() == JXL_ENC_ERR_NOT_SUPPORTED: lossless
 * (Modular) encoding, alpha / extra channels, samples other than 8-bit sRGB, JPEG transcoding, metadata boxes. */
typedef struct JxlEncoderStruct JxlEncoder;
typedef struct JxlEncoderFrameSettingsStruct JxlEncoderFrameSettings;
typedef enum { JXL_ENC_SUCCESS = 0, JXL_ENC_ERROR = 1, JXL_ENC_NEED_MORE_OUTPUT = 2 } JxlEncoderStatus;
typedef enum {
  JXL_ENC_ERR_OK = 0, JXL_ENC_ERR_GENERIC = 1, JXL_ENC_ERR_OOM = 2, JXL_ENC_ERR_JBRD = 3, JXL_ENC_ERR_BAD_INPUT = 4,
  JXL_ENC_ERR_NOT_SUPPORTED = 0x80, JXL_ENC_ERR_API_USAGE = 0x81
} JxlEncoderError;
/* jpegxl-sys/src/encoder/encode.rs:110-341; the ids between DECODING_SPEED and LAST are accepted without effect. */
typedef enum {
  JXL_ENC_FRAME_SETTING_EFFORT = 0, JXL_ENC_FRAME_SETTING_DECODING_SPEED = 1, JXL_ENC_FRAME_SETTING_LAST = 39,
  JXL_ENC_FRAME_SETTING_FILL_ENUM = 65535
} JxlEncoderFrameSettingId;
/* JxlColorEncoding, byte-identical to libjxl (jpegxl-sys/src/color/color_encoding.rs:30-159). */
typedef enum { JXL_COLOR_SPACE_RGB = 0, JXL_COLOR_SPACE_GRAY, JXL_COLOR_SPACE_XYB, JXL_COLOR_SPACE_UNKNOWN } JxlColorSpace;
typedef enum { JXL_WHITE_POINT_D65 = 1, JXL_WHITE_POINT_CUSTOM = 2, JXL_WHITE_POINT_E = 10, JXL_WHITE_POINT_DCI = 11 } JxlWhitePoint;
typedef enum { JXL_PRIMARIES_SRGB = 1, JXL_PRIMARIES_CUSTOM = 2, JXL_PRIMARIES_2100 = 9, JXL_PRIMARIES_P3 = 11 } JxlPrimaries;
typedef enum {
  JXL_TRANSFER_FUNCTION_709 = 1, JXL_TRANSFER_FUNCTION_UNKNOWN = 2, JXL_TRANSFER_FUNCTION_LINEAR = 8,
  JXL_TRANSFER_FUNCTION_SRGB = 13, JXL_TRANSFER_FUNCTION_PQ = 16, JXL_TRANSFER_FUNCTION_DCI = 17,
  JXL_TRANSFER_FUNCTION_HLG = 18, JXL_TRANSFER_FUNCTION_GAMMA = 65535
} JxlTransferFunction;
typedef enum {
  JXL_RENDERING_INTENT_PERCEPTUAL = 0, JXL_RENDERING_INTENT_RELATIVE, JXL_RENDERING_INTENT_SATURATION,
  JXL_RENDERING_INTENT_ABSOLUTE
} JxlRenderingIntent;
typedef struct {
  JxlColorSpace color_space;
  JxlWhitePoint white_point;
  double white_point_xy[2];
  JxlPrimaries primaries;
  double primaries_red_xy[2];
  double primaries_green_xy[2];
  double primaries_blue_xy[2];
  JxlTransferFunction transfer_function;
  double gamma;
  JxlRenderingIntent rendering_intent;
} JxlColorEncoding;

uint32_t JxlEncoderVersion(void);
/* memory_manager: accepted, not used (the codestream is the only host allocation; pixels live in device memory). */
JxlEncoder* JxlEncoderCreate(const void* memory_manager);
void JxlEncoderReset(JxlEncoder* enc);
void JxlEncoderDestroy(JxlEncoder* enc);
JxlEncoderError JxlEncoderGetError(JxlEncoder* enc);
/* Accepted and ignored, as on the decoder. */
JxlEncoderStatus JxlEncoderSetParallelRunner(JxlEncoder* enc, void* parallel_runner, void* parallel_runner_opaque);
JxlEncoderFrameSettings* JxlEncoderFrameSettingsCreate(JxlEncoder* enc, const JxlEncoderFrameSettings* source);
JxlEncoderStatus JxlEncoderUseContainer(JxlEncoder* enc, JXL_BOOL use_container);
JxlEncoderStatus JxlEncoderUseBoxes(JxlEncoder* enc);
JxlEncoderStatus JxlEncoderAddBox(JxlEncoder* enc, const char* type, const uint8_t* contents, size_t size, JXL_BOOL compress_box);
JxlEncoderStatus JxlEncoderStoreJPEGMetadata(JxlEncoder* enc, JXL_BOOL store_jpeg_metadata);
JxlEncoderStatus JxlEncoderSetFrameLossless(JxlEncoderFrameSettings* frame_settings, JXL_BOOL lossless);
JxlEncoderStatus JxlEncoderSetFrameDistance(JxlEncoderFrameSettings* frame_settings, float distance);
float JxlEncoderDistanceFromQuality(float quality);
JxlEncoderStatus JxlEncoderFrameSettingsSetOption(JxlEncoderFrameSettings* frame_settings, JxlEncoderFrameSettingId option, int64_t value);
void JxlEncoderInitBasicInfo(JxlBasicInfo* info);
JxlEncoderStatus JxlEncoderSetBasicInfo(JxlEncoder* enc, const JxlBasicInfo* info);
JxlEncoderStatus JxlEncoderSetColorEncoding(JxlEncoder* enc, const JxlColorEncoding* color);
void JxlColorEncodingSetToSRGB(JxlColorEncoding* color_encoding, JXL_BOOL is_gray);
void JxlColorEncodingSetToLinearSRGB(JxlColorEncoding* color_encoding, JXL_BOOL is_gray);
JxlEncoderStatus JxlEncoderAddImageFrame(const JxlEncoderFrameSettings* frame_settings, const JxlPixelFormat* pixel_format,
                                         const void* buffer, size_t size);
JxlEncoderStatus JxlEncoderAddJPEGFrame(const JxlEncoderFrameSettings* frame_settings, const uint8_t* buffer, size_t size);
void JxlEncoderCloseInput(JxlEncoder* enc);
/* Runs the CUDA encoder on the first call after a frame was added; NEED_MORE_OUTPUT until everything is written. */
JxlEncoderStatus JxlEncoderProcessOutput(JxlEncoder* enc, uint8_t** next_out, size_t* avail_out);
/* Text of the last error (extension; libjxl only prints it in debug builds). */
const char* JxlB200EncoderApiMessage(const JxlEncoder* enc);

/* ---- 5. libjxl_threads symbols (jpegxl-sys/src/threads/{thread,resizable}_parallel_runner.rs) ----
 * jpegxl-rs hands the two runner functions to SetParallelRunner by value (jpegxl-rs/src/parallel/). The GPU path
 * never calls a runner; the symbols exist so that the crate links, and they behave like libjxl's runner with zero
 * worker threads (the range runs on the calling thread). */
typedef int (*JxlParallelRunInit)(void* jpegxl_opaque, size_t num_threads);
typedef void (*JxlParallelRunFunction)(void* jpegxl_opaque, uint32_t value, size_t thread_id);
int JxlThreadParallelRunner(void* runner_opaque, void* jpegxl_opaque, JxlParallelRunInit init, JxlParallelRunFunction func,
                            uint32_t start_range, uint32_t end_range);
void* JxlThreadParallelRunnerCreate(const void* memory_manager, size_t num_worker_threads);
void JxlThreadParallelRunnerDestroy(void* runner_opaque);
size_t JxlThreadParallelRunnerDefaultNumWorkerThreads(void);
int JxlResizableParallelRunner(void* runner_opaque, void* jpegxl_opaque, JxlParallelRunInit init, JxlParallelRunFunction func,
                               uint32_t start_range, uint32_t end_range);
void* JxlResizableParallelRunnerCreate(const void* memory_manager);
void JxlResizableParallelRunnerSetThreads(void* runner_opaque, size_t num_threads);
uint32_t JxlResizableParallelRunnerSuggestThreads(uint64_t xsize, uint64_t ysize);
void JxlResizableParallelRunnerDestroy(void* runner_opaque);

#ifdef __cplusplus
}
#endif

#endif /* JXL_B200_H_ */
