// jxl_b200: CUDA kernels (sm_100a) and the C ABI declared in include/jxl_b200.h.
//
// Kernel map:
//   k_modular_decode  one thread = one Modular entropy-coded stream (group x pass x frame)
//   k_group_programs  one CTA = one group's inverse transforms + scatter into the frame planes
//   k_frame_level     grid-wide: the k-th global inverse transform of every frame
//   k_write_output    int32 planes -> interleaved u8/u16/f16/f32 pixels, coalesced stores
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/jxl_b200.h"
#include "host/jxlb_batch.h"
#include "kernels/jxlb_finish_dev.h"

namespace jxlb {

// ------------------------------------------------------------------ kernels
// One warp per CTA: 32 streams in lock step. Properties live in shared memory as
// [property][lane] (bank = lane, conflict free for any per-lane property index).
template <typename WT>
__global__ void __launch_bounds__(32) k_modular_decode(DevPools P) {
  __shared__ int32_t props_s[kDevMaxProps * 32];
  __shared__ uint32_t div_s[64];
  const uint32_t lane = threadIdx.x;
  for (uint32_t i = lane; i < 64; i += 32) div_s[i] = (1u << 24) / (i + 1);
  __syncwarp();
  const uint32_t s = blockIdx.x * 32 + lane;
  DevLaneMem m;
  m.props = props_s + lane;
  m.props_stride = 32;
  m.divlut = div_s;
  m.ring_w = P.wp_width;
  m.lane_stride = 32;
  m.ring = P.ring + static_cast<size_t>(blockIdx.x) * 3 * P.wp_width * 32 + lane;
  m.wp = P.wp_scratch + static_cast<size_t>(blockIdx.x) * 10 * (P.wp_width + 2) * 32 + lane;
  const bool valid = s < P.num_streams;
  const uint32_t status = DevDecodeModularStream<WT>(P, s, m, P.warp_dims + P.warp_dims_off[blockIdx.x],
                                                     P.warp_chans[blockIdx.x], valid);
  if (valid) P.status[s] = status;
}

__global__ void __launch_bounds__(256) k_group_programs(DevPools P, const DevOp* ops, const DevProgram* programs) {
  const DevProgram pr = programs[blockIdx.x];
  for (uint32_t o = pr.op_begin; o < pr.op_end; o++) {
    DevRunOp(P, ops[o], threadIdx.x, blockDim.x);
    __syncthreads();
    __threadfence_block();
  }
}

// blockIdx.y selects the frame, blockIdx.x / gridDim.x partitions the elements.
__global__ void __launch_bounds__(256) k_frame_level(DevPools P, const DevOp* ops, const DevProgram* level) {
  const DevProgram pr = level[blockIdx.y];
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  for (uint32_t o = pr.op_begin; o < pr.op_end; o++) DevRunOp(P, ops[o], tid, nthreads);
}

// One thread per pixel; blockIdx.y = frame.
__global__ void __launch_bounds__(256) k_write_output(DevPools P, const DevFrameOut* frames, uint8_t* out) {
  const DevFrameOut& fo = frames[blockIdx.y];
  const uint64_t n = static_cast<uint64_t>(fo.xsize) * fo.ysize;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t y = static_cast<uint32_t>(i / fo.xsize), x = static_cast<uint32_t>(i % fo.xsize);
    DevWritePixel(P, fo, out, x, y);
  }
}

// RGBA8 fast path: 4 integer planes -> one 32-bit store per pixel.
__global__ void __launch_bounds__(256) k_write_output_rgba8(DevPools P, const DevFrameOut* frames, uint8_t* out) {
  const DevFrameOut& fo = frames[blockIdx.y];
  const uint64_t n = static_cast<uint64_t>(fo.xsize) * fo.ysize;
  for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t y = static_cast<uint32_t>(i / fo.xsize), x = static_cast<uint32_t>(i % fo.xsize);
    const float d = DevDither(x, y);
    uint32_t px = 0;
#pragma unroll
    for (uint32_t c = 0; c < 4; c++) {
      float v = DevSampleFloat(P, fo, c, x, y);
      v = __fadd_rn(__fmul_rn(v, 255.0f), d);
      if (!(v >= 0.0f)) v = 0.0f;
      if (v > 255.0f) v = 255.0f;
      px |= static_cast<uint32_t>(__float2int_rn(v)) << (8 * c);
    }
    *reinterpret_cast<uint32_t*>(out + fo.out_off + fo.stride * y + 4ull * x) = px;
  }
}

// ------------------------------------------------------------------ runtime
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  ~DevBuf() { Free(); }
  void Free() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t Alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    Free();
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t Upload(const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = Alloc(v.size());
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

}  // namespace jxlb

using namespace jxlb;

#define CUDA_OK(expr)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      dec->error = std::string(#expr) + ": " + cudaGetErrorString(e_);                       \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

struct JxlB200Decoder {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string error;
  std::unique_ptr<BatchPlan> plan;
  DevBuf<uint8_t> d_bytes, d_out;
  DevBuf<DevAlias> d_alias;
  DevBuf<uint32_t> d_prefix, d_cfg, d_refs, d_lz77, d_status, d_warp_chans, d_warp_dims_off, d_warp_dims;
  DevBuf<DevTreeNode> d_tree;
  DevBuf<DevCode> d_codes;
  DevBuf<DevChannel> d_chans;
  DevBuf<DevStream> d_streams;
  DevBuf<DevPlane> d_planes;
  DevBuf<DevOp> d_ops;
  DevBuf<DevProgram> d_group_programs, d_levels;
  DevBuf<DevFrameOut> d_frames;
  DevBuf<int32_t> d_arena, d_wp, d_ring;
  std::vector<size_t> level_off;  // offset of each level inside d_levels
  std::vector<uint32_t> h_status;
  bool uniform_rgba8 = false;
  uint32_t launches = 0;
  DevPools pools{};
  // optional per-kernel timing (CUDA events on the launching stream)
  bool profiling = false;
  static constexpr int kEvRuns = 64;          // event sets kept before folding
  cudaEvent_t ev[kEvRuns][5] = {};
  bool ev_created = false;
  int ev_used = 0;                            // recorded, not yet folded
  double kernel_ms[4] = {0, 0, 0, 0};         // decode, group programs, frame levels, output (accumulated)
  uint32_t profiled_runs = 0;
};

static void FoldEvents(JxlB200Decoder* dec) {
  for (int r = 0; r < dec->ev_used; r++) {
    if (cudaEventSynchronize(dec->ev[r][4]) != cudaSuccess) continue;
    for (int k = 0; k < 4; k++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, dec->ev[r][k], dec->ev[r][k + 1]);
      dec->kernel_ms[k] += ms;
    }
    dec->profiled_runs++;
  }
  dec->ev_used = 0;
}

extern "C" {

JxlB200Decoder* JxlB200DecoderCreate(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= device || device < 0) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  JxlB200Decoder* dec = new JxlB200Decoder();
  dec->device = device;
  if (cudaStreamCreateWithFlags(&dec->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete dec;
    return nullptr;
  }
  return dec;
}

void JxlB200DecoderDestroy(JxlB200Decoder* dec) {
  if (!dec) return;
  cudaSetDevice(dec->device);
  if (dec->stream) cudaStreamDestroy(dec->stream);
  delete dec;
}

const char* JxlB200DecoderGetError(const JxlB200Decoder* dec) { return dec ? dec->error.c_str() : "null decoder"; }

int JxlB200DecoderSetInputBatch(JxlB200Decoder* dec, const uint8_t* const* files, const size_t* sizes, size_t n,
                                const JxlPixelFormat* format, int num_threads) {
  if (!dec || !files || !sizes || !format || n == 0) return 1;
  dec->error.clear();
  PixelFormat fmt;
  fmt.num_channels = format->num_channels;
  fmt.data_type = format->data_type;
  fmt.endianness = format->endianness;
  fmt.align = format->align;
  if (fmt.num_channels < 1 || fmt.num_channels > 4 ||
      !(fmt.data_type == 0 || fmt.data_type == 2 || fmt.data_type == 3 || fmt.data_type == 5)) {
    dec->error = "invalid pixel format";
    return 1;
  }
  std::unique_ptr<BatchPlan> plan(new BatchPlan());
  try {
    PlanBatch(files, sizes, n, fmt, num_threads, plan.get());
  } catch (const std::exception& e) {
    dec->error = e.what();
    return 1;
  }
  CUDA_OK(cudaSetDevice(dec->device));
  cudaStream_t s = dec->stream;
  BatchPlan& b = *plan;
  CUDA_OK(dec->d_bytes.Upload(b.bytes, s));
  CUDA_OK(dec->d_alias.Upload(b.alias, s));
  CUDA_OK(dec->d_prefix.Upload(b.prefix, s));
  CUDA_OK(dec->d_cfg.Upload(b.cfg, s));
  CUDA_OK(dec->d_refs.Upload(b.refs, s));
  CUDA_OK(dec->d_tree.Upload(b.tree, s));
  CUDA_OK(dec->d_codes.Upload(b.codes, s));
  CUDA_OK(dec->d_chans.Upload(b.chans, s));
  CUDA_OK(dec->d_streams.Upload(b.streams, s));
  CUDA_OK(dec->d_planes.Upload(b.planes, s));
  CUDA_OK(dec->d_ops.Upload(b.ops, s));
  CUDA_OK(dec->d_group_programs.Upload(b.group_programs, s));
  std::vector<DevProgram> all_levels;
  dec->level_off.clear();
  for (const auto& lvl : b.levels) {
    dec->level_off.push_back(all_levels.size());
    all_levels.insert(all_levels.end(), lvl.begin(), lvl.end());
  }
  CUDA_OK(dec->d_levels.Upload(all_levels, s));
  CUDA_OK(dec->d_frames.Upload(b.frames, s));
  CUDA_OK(dec->d_warp_chans.Upload(b.warp_chans, s));
  CUDA_OK(dec->d_warp_dims_off.Upload(b.warp_dims_off, s));
  CUDA_OK(dec->d_warp_dims.Upload(b.warp_dims, s));
  CUDA_OK(dec->d_arena.Alloc(b.arena_size + 16));
  const size_t num_warps = (b.streams.size() + 31) / 32;
  CUDA_OK(dec->d_wp.Alloc(num_warps * 10 * (b.wp_width + 2) * 32 + 16));
  CUDA_OK(dec->d_ring.Alloc(num_warps * 3 * b.wp_width * 32 + 16));
  CUDA_OK(dec->d_lz77.Alloc(static_cast<size_t>(b.lz77_slots) << 20));
  CUDA_OK(dec->d_status.Alloc(b.streams.size()));
  CUDA_OK(dec->d_out.Alloc(b.out_size));
  CUDA_OK(cudaStreamSynchronize(s));
  DevPools& P = dec->pools;
  P.words = reinterpret_cast<const uint32_t*>(dec->d_bytes.p);
  P.alias = dec->d_alias.p;
  P.prefix = dec->d_prefix.p;
  P.cfg = dec->d_cfg.p;
  P.tree = dec->d_tree.p;
  P.chans = dec->d_chans.p;
  P.streams = dec->d_streams.p;
  P.planes = dec->d_planes.p;
  P.refs = dec->d_refs.p;
  P.codes = dec->d_codes.p;
  P.arena = dec->d_arena.p;
  P.wp_scratch = dec->d_wp.p;
  P.ring = dec->d_ring.p;
  P.wp_width = b.wp_width;
  P.lz77 = dec->d_lz77.p;
  P.status = dec->d_status.p;
  P.num_streams = b.streams.size();
  P.warp_chans = dec->d_warp_chans.p;
  P.warp_dims_off = dec->d_warp_dims_off.p;
  P.warp_dims = dec->d_warp_dims.p;
  dec->uniform_rgba8 = fmt.num_channels == 4 && fmt.data_type == 2;
  for (const DevFrameOut& fo : b.frames)
    for (int c = 0; c < 4; c++)
      if (fo.is_float[c] || fo.stride % 4) dec->uniform_rgba8 = false;
  dec->plan = std::move(plan);
  return 0;
}

size_t JxlB200DecoderNumFrames(const JxlB200Decoder* dec) { return dec && dec->plan ? dec->plan->frames.size() : 0; }

static void FillBasicInfo(const BasicInfo& bi, bool have_container, JxlBasicInfo* info) {
  std::memset(info, 0, sizeof(*info));
  const ImageMetadata& m = bi.meta;
  info->have_container = have_container;
  info->xsize = bi.xsize;
  info->ysize = bi.ysize;
  info->bits_per_sample = m.bit_depth.bits;
  info->exponent_bits_per_sample = m.bit_depth.exp_bits;
  info->intensity_target = m.intensity_target;
  info->min_nits = m.min_nits;
  info->relative_to_max_display = m.relative_to_max_display;
  info->linear_below = m.linear_below;
  info->uses_original_profile = !m.xyb_encoded;
  info->have_preview = m.have_preview;
  info->have_animation = m.have_animation;
  info->orientation = m.orientation;
  info->num_color_channels = m.color.IsGray() ? 1 : 3;
  info->num_extra_channels = m.extra.size();
  int a = m.AlphaIndex();
  if (a >= 0) {
    info->alpha_bits = m.extra[a].bit_depth.bits;
    info->alpha_exponent_bits = m.extra[a].bit_depth.exp_bits;
    info->alpha_premultiplied = m.extra[a].alpha_associated;
  }
  info->preview.xsize = m.preview_size.xsize;
  info->preview.ysize = m.preview_size.ysize;
  info->animation.tps_numerator = m.tps_num;
  info->animation.tps_denominator = m.tps_den;
  info->animation.num_loops = m.num_loops;
  info->animation.have_timecodes = m.have_timecodes;
  info->intrinsic_xsize = m.have_intrinsic_size ? m.intrinsic_size.xsize : bi.xsize;
  info->intrinsic_ysize = m.have_intrinsic_size ? m.intrinsic_size.ysize : bi.ysize;
}

int JxlB200DecoderGetBasicInfo(const JxlB200Decoder* dec, size_t i, JxlBasicInfo* info) {
  if (!dec || !dec->plan || i >= dec->plan->info.size() || !info) return 1;
  FillBasicInfo(dec->plan->info[i], false, info);
  return 0;
}

size_t JxlB200DecoderImageOutBufferSize(const JxlB200Decoder* dec, size_t i) {
  if (!dec || !dec->plan || i >= dec->plan->frames.size()) return 0;
  return dec->plan->frame_out_size[i];
}

int JxlB200DecoderRun(JxlB200Decoder* dec, void* cuda_stream) {
  if (!dec || !dec->plan) return 1;
  CUDA_OK(cudaSetDevice(dec->device));
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : dec->stream;
  const BatchPlan& b = *dec->plan;
  const DevPools& P = dec->pools;
  uint32_t launches = 0;
  const bool prof = dec->profiling;
  if (prof && !dec->ev_created) {
    for (auto& set : dec->ev)
      for (auto& e : set) CUDA_OK(cudaEventCreate(&e));
    dec->ev_created = true;
  }
  if (prof && dec->ev_used == JxlB200Decoder::kEvRuns) FoldEvents(dec);  // blocks only every 64 runs
  cudaEvent_t* ev = prof ? dec->ev[dec->ev_used] : nullptr;
  if (prof) cudaEventRecord(ev[0], s);
  if (!b.streams.empty()) {
    const uint32_t block = 32;
    if (b.narrow) {
      k_modular_decode<int32_t><<<(b.streams.size() + block - 1) / block, block, 0, s>>>(P);
    } else {
      k_modular_decode<int64_t><<<(b.streams.size() + block - 1) / block, block, 0, s>>>(P);
    }
    launches++;
  }
  if (prof) cudaEventRecord(ev[1], s);
  if (!b.group_programs.empty()) {
    k_group_programs<<<b.group_programs.size(), 256, 0, s>>>(P, dec->d_ops.p, dec->d_group_programs.p);
    launches++;
  }
  if (prof) cudaEventRecord(ev[2], s);
  for (size_t k = 0; k < b.levels.size(); k++) {
    const uint32_t tiles = std::max<uint32_t>(1, std::min<uint32_t>(1024, (b.max_frame_pixels + 1023) / 1024));
    dim3 grid(tiles, b.levels[k].size());
    k_frame_level<<<grid, 256, 0, s>>>(P, dec->d_ops.p, dec->d_levels.p + dec->level_off[k]);
    launches++;
  }
  if (prof) cudaEventRecord(ev[3], s);
  {
    const uint32_t tiles = std::max<uint32_t>(1, std::min<uint32_t>(4096, (b.max_frame_pixels + 255) / 256));
    dim3 grid(tiles, b.frames.size());
    if (dec->uniform_rgba8) {
      k_write_output_rgba8<<<grid, 256, 0, s>>>(P, dec->d_frames.p, dec->d_out.p);
    } else {
      k_write_output<<<grid, 256, 0, s>>>(P, dec->d_frames.p, dec->d_out.p);
    }
    launches++;
  }
  if (prof) {
    cudaEventRecord(ev[4], s);
    dec->ev_used++;
  }
  dec->launches = launches;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int JxlB200DecoderSetProfiling(JxlB200Decoder* dec, int enabled) {
  if (!dec) return 1;
  dec->profiling = enabled != 0;
  for (double& m : dec->kernel_ms) m = 0;
  dec->profiled_runs = 0;
  dec->ev_used = 0;
  return 0;
}

int JxlB200DecoderGetKernelTimes(JxlB200Decoder* dec, double* ms4, uint32_t* runs) {
  if (!dec || !ms4 || !runs) return 1;
  FoldEvents(dec);
  for (int k = 0; k < 4; k++) ms4[k] = dec->kernel_ms[k];
  *runs = dec->profiled_runs;
  return 0;
}

int JxlB200DecoderWait(JxlB200Decoder* dec, void* cuda_stream) {
  if (!dec || !dec->plan) return 1;
  CUDA_OK(cudaSetDevice(dec->device));
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : dec->stream;
  dec->h_status.resize(dec->plan->streams.size());
  if (!dec->h_status.empty())
    CUDA_OK(cudaMemcpyAsync(dec->h_status.data(), dec->d_status.p, dec->h_status.size() * 4, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  for (size_t i = 0; i < dec->h_status.size(); i++) {
    if (dec->h_status[i] != 0) {
      dec->error = "entropy-coded stream " + std::to_string(i) + " failed (status " + std::to_string(dec->h_status[i]) +
                   ": 1 = read past section end, 2 = bad ANS final state)";
      return 1;
    }
  }
  return 0;
}

void* JxlB200DecoderDeviceOutput(const JxlB200Decoder* dec, size_t i) {
  if (!dec || !dec->plan || i >= dec->plan->frames.size()) return nullptr;
  return dec->d_out.p + dec->plan->frames[i].out_off;
}

int JxlB200DecoderReadOutput(JxlB200Decoder* dec, size_t i, void* dst, size_t size) {
  if (!dec || !dec->plan || i >= dec->plan->frames.size() || !dst) return 1;
  if (size < dec->plan->frame_out_size[i]) {
    dec->error = "output buffer too small";
    return 1;
  }
  CUDA_OK(cudaSetDevice(dec->device));
  CUDA_OK(cudaMemcpyAsync(dst, dec->d_out.p + dec->plan->frames[i].out_off, dec->plan->frame_out_size[i],
                          cudaMemcpyDeviceToHost, dec->stream));
  CUDA_OK(cudaStreamSynchronize(dec->stream));
  return 0;
}

int JxlB200DecoderReadOutputs(JxlB200Decoder* dec, void* const* dsts, const size_t* sizes, size_t n) {
  if (!dec || !dec->plan || n != dec->plan->frames.size() || !dsts || !sizes) return 1;
  CUDA_OK(cudaSetDevice(dec->device));
  for (size_t i = 0; i < n; i++) {
    if (sizes[i] < dec->plan->frame_out_size[i]) {
      dec->error = "output buffer too small";
      return 1;
    }
    CUDA_OK(cudaMemcpyAsync(dsts[i], dec->d_out.p + dec->plan->frames[i].out_off, dec->plan->frame_out_size[i],
                            cudaMemcpyDeviceToHost, dec->stream));
  }
  CUDA_OK(cudaStreamSynchronize(dec->stream));
  return 0;
}

int JxlB200DecoderGetStats(const JxlB200Decoder* dec, JxlB200Stats* st) {
  if (!dec || !dec->plan || !st) return 1;
  const BatchPlan& b = *dec->plan;
  st->compressed_bytes = b.compressed_bytes;
  st->output_bytes = 0;
  for (uint64_t s : b.frame_out_size) st->output_bytes += s;
  st->pixels = b.total_pixels;
  st->num_streams = b.streams.size();
  st->arena_bytes = b.arena_size * 4;
  st->kernel_launches = (b.streams.empty() ? 0 : 1) + (b.group_programs.empty() ? 0 : 1) + b.levels.size() + 1;
  return 0;
}

// ------------------------------------------------------------------ libjxl-compatible subset
struct JxlDecoderStruct {
  JxlB200Decoder* gpu = nullptr;
  const uint8_t* input = nullptr;
  size_t input_size = 0;
  bool input_closed = false;
  int events_wanted = 0;
  int stage = 0;  // 0 start, 1 basic info sent, 2 waiting for buffer, 3 full image sent, 4 done
  bool have_info = false;
  BasicInfo info;
  bool have_container = false;
  JxlPixelFormat format{};
  void* out_buffer = nullptr;
  size_t out_size = 0;
  bool keep_orientation = false;
};

uint32_t JxlDecoderVersion(void) { return 11002; }  // lib/jxl/version.h: 0.11.2

JxlSignature JxlSignatureCheck(const uint8_t* buf, size_t len) {  // lib/jxl/decode.cc:137-175
  if (len == 0) return JXL_SIG_NOT_ENOUGH_BYTES;
  static const uint8_t kBox[12] = {0, 0, 0, 0xC, 'J', 'X', 'L', ' ', 0xD, 0xA, 0x87, 0xA};
  if (buf[0] == 0xFF) {
    if (len < 2) return JXL_SIG_NOT_ENOUGH_BYTES;
    return buf[1] == 0x0A ? JXL_SIG_CODESTREAM : JXL_SIG_INVALID;
  }
  size_t n = len < 12 ? len : 12;
  if (std::memcmp(buf, kBox, n) != 0) return JXL_SIG_INVALID;
  return len < 12 ? JXL_SIG_NOT_ENOUGH_BYTES : JXL_SIG_CONTAINER;
}

JxlDecoder* JxlDecoderCreate(const void* memory_manager) {
  if (memory_manager != nullptr) return nullptr;
  JxlDecoder* d = new JxlDecoderStruct();
  return d;
}

void JxlDecoderReset(JxlDecoder* dec) {
  if (!dec) return;
  JxlB200Decoder* gpu = dec->gpu;
  *dec = JxlDecoderStruct();
  dec->gpu = gpu;
}

void JxlDecoderDestroy(JxlDecoder* dec) {
  if (!dec) return;
  JxlB200DecoderDestroy(dec->gpu);
  delete dec;
}

JxlDecoderStatus JxlDecoderSetParallelRunner(JxlDecoder* dec, void*, void*) { return dec ? JXL_DEC_SUCCESS : JXL_DEC_ERROR; }

JxlDecoderStatus JxlDecoderSubscribeEvents(JxlDecoder* dec, int events_wanted) {
  if (!dec || dec->stage != 0) return JXL_DEC_ERROR;
  if (events_wanted & 63) return JXL_DEC_ERROR;  // lib/jxl/decode.cc:862-870
  dec->events_wanted = events_wanted;
  return JXL_DEC_SUCCESS;
}
JxlDecoderStatus JxlDecoderSetKeepOrientation(JxlDecoder* dec, JXL_BOOL v) {
  if (!dec || dec->stage != 0) return JXL_DEC_ERROR;
  dec->keep_orientation = v;
  return JXL_DEC_SUCCESS;
}
JxlDecoderStatus JxlDecoderSetUnpremultiplyAlpha(JxlDecoder* dec, JXL_BOOL v) {
  if (!dec || dec->stage != 0) return JXL_DEC_ERROR;
  return v ? JXL_DEC_ERROR : JXL_DEC_SUCCESS;  // un-premultiply is not implemented on the GPU path
}
JxlDecoderStatus JxlDecoderSetRenderSpotcolors(JxlDecoder* dec, JXL_BOOL) { return dec && dec->stage == 0 ? JXL_DEC_SUCCESS : JXL_DEC_ERROR; }
JxlDecoderStatus JxlDecoderSetCoalescing(JxlDecoder* dec, JXL_BOOL) { return dec && dec->stage == 0 ? JXL_DEC_SUCCESS : JXL_DEC_ERROR; }
JxlDecoderStatus JxlDecoderSetDesiredIntensityTarget(JxlDecoder* dec, float v) {
  if (!dec || dec->stage != 0 || v < 0) return JXL_DEC_ERROR;
  return JXL_DEC_SUCCESS;
}

JxlDecoderStatus JxlDecoderSetInput(JxlDecoder* dec, const uint8_t* data, size_t size) {
  if (!dec || dec->input) return JXL_DEC_ERROR;
  dec->input = data;
  dec->input_size = size;
  return JXL_DEC_SUCCESS;
}
void JxlDecoderCloseInput(JxlDecoder* dec) {
  if (dec) dec->input_closed = true;
}

// Event order as jpegxl-rs expects it (jpegxl-rs/src/decode.rs:234-324):
// BASIC_INFO -> NEED_IMAGE_OUT_BUFFER -> FULL_IMAGE -> SUCCESS.
JxlDecoderStatus JxlDecoderProcessInput(JxlDecoder* dec) {
  if (!dec) return JXL_DEC_ERROR;
  if (!dec->input || dec->input_size == 0) return dec->input_closed ? JXL_DEC_ERROR : JXL_DEC_NEED_MORE_INPUT;
  JxlSignature sig = JxlSignatureCheck(dec->input, dec->input_size);
  if (sig == JXL_SIG_INVALID) return JXL_DEC_ERROR;
  if (sig == JXL_SIG_NOT_ENOUGH_BYTES) return dec->input_closed ? JXL_DEC_ERROR : JXL_DEC_NEED_MORE_INPUT;
  if (!dec->have_info) {
    try {
      CodestreamView v = FindCodestream(dec->input, dec->input_size);
      dec->info = ReadBasicInfo(v.data, v.size);
      dec->have_container = sig == JXL_SIG_CONTAINER;
      dec->have_info = true;
    } catch (const std::exception&) {
      return JXL_DEC_ERROR;
    }
  }
  if (dec->stage == 0) {
    dec->stage = 1;
    if (dec->events_wanted & JXL_DEC_BASIC_INFO) return JXL_DEC_BASIC_INFO;
  }
  if (dec->stage == 1) {
    if (!(dec->events_wanted & JXL_DEC_FULL_IMAGE)) {
      dec->stage = 4;
      return JXL_DEC_SUCCESS;
    }
    if (!dec->out_buffer) return JXL_DEC_NEED_IMAGE_OUT_BUFFER;
    if (!dec->gpu) dec->gpu = JxlB200DecoderCreate(0);
    if (!dec->gpu) return JXL_DEC_ERROR;  // no CUDA device: fail, never fall back
    const uint8_t* files[1] = {dec->input};
    size_t sizes[1] = {dec->input_size};
    if (JxlB200DecoderSetInputBatch(dec->gpu, files, sizes, 1, &dec->format, 1) != 0) return JXL_DEC_ERROR;
    if (JxlB200DecoderRun(dec->gpu, nullptr) != 0) return JXL_DEC_ERROR;
    if (JxlB200DecoderWait(dec->gpu, nullptr) != 0) return JXL_DEC_ERROR;
    if (JxlB200DecoderReadOutput(dec->gpu, 0, dec->out_buffer, dec->out_size) != 0) return JXL_DEC_ERROR;
    dec->stage = 3;
    return JXL_DEC_FULL_IMAGE;
  }
  dec->stage = 4;
  return JXL_DEC_SUCCESS;
}

JxlDecoderStatus JxlDecoderGetBasicInfo(const JxlDecoder* dec, JxlBasicInfo* info) {
  if (!dec || !dec->have_info) return JXL_DEC_NEED_MORE_INPUT;
  if (info) FillBasicInfo(dec->info, dec->have_container, info);
  return JXL_DEC_SUCCESS;
}

JxlDecoderStatus JxlDecoderImageOutBufferSize(const JxlDecoder* dec, const JxlPixelFormat* format, size_t* size) {
  if (!dec || !dec->have_info || !format || !size) return JXL_DEC_ERROR;
  if (format->num_channels < 1 || format->num_channels > 4) return JXL_DEC_ERROR;
  if (format->num_channels < 3 && !dec->info.meta.color.IsGray()) return JXL_DEC_ERROR;  // lib/jxl/decode.cc:2334-2337
  PixelFormat f;
  f.num_channels = format->num_channels;
  f.data_type = format->data_type;
  f.endianness = format->endianness;
  f.align = format->align;
  *size = OutputStride(dec->info.xsize, f) * dec->info.ysize;
  return JXL_DEC_SUCCESS;
}

JxlDecoderStatus JxlDecoderSetImageOutBuffer(JxlDecoder* dec, const JxlPixelFormat* format, void* buffer, size_t size) {
  size_t need = 0;
  if (JxlDecoderImageOutBufferSize(dec, format, &need) != JXL_DEC_SUCCESS) return JXL_DEC_ERROR;
  if (!buffer || size < need) return JXL_DEC_ERROR;
  dec->format = *format;
  dec->out_buffer = buffer;
  dec->out_size = size;
  return JXL_DEC_SUCCESS;
}

}  // extern "C"
