// jxl_b200: CUDA kernels (sm_100a) and the C ABI declared in include/jxl_b200.h.
//
// Kernel map:
//   k_modular_decode  one thread = one Modular entropy-coded stream (group x pass x frame)
//   k_group_programs  one CTA = one group's inverse transforms + scatter into the frame planes
//   k_frame_level     grid-wide: the k-th global inverse transform of every frame
//   k_write_output    int32 planes -> interleaved u8/u16/f16/f32 pixels, coalesced stores
// VarDCT frames (kernels/jxlb_vardct_dev.h):
//   k_dc_finish       one CTA = one DC group: DC dequantisation, block side information, EPF sigma
//   k_dc_smooth       adaptive DC smoothing
//   k_ac_decode       one thread = one (frame, group, pass) AC stream -> sparse coefficient tokens
//   k_dequant_idct    one CTA = one 256x256 group: token scatter, dequant + CfL, inverse transforms;
//                     warp per varblock up to 32x32 (register-resident IDCTs)
//   k_idct_mid        64x32 / 32x64 / 64x64 varblocks, one CTA per varblock, staged transforms in shared memory
//   k_idct_big        varblocks of 128x128 and larger, scratch in global memory
//   k_gaborish / k_epf / k_color_write   render stages, one thread per pixel
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/jxl_b200.h"
#include "host/jxlb_batch.h"
#include "kernels/jxlb_modular_coop_dev.h"
#include "kernels/jxlb_finish_dev.h"
#include "kernels/jxlb_vardct_dev.h"
#include "kernels/jxlb_enc_dev.h"
#include "kernels/jxlb_encl_dev.h"
#include "host/jxlb_enc_host.h"
#include "host/jxlb_encl_host.h"

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
  const uint32_t s = P.stream0 + blockIdx.x * 32 + lane;
  DevLaneMem m;
  m.props = props_s + lane;
  m.props_stride = 32;
  m.divlut = div_s;
  m.ring_w = P.wp_width;
  m.lane_stride = 32;
  m.ring = P.ring + static_cast<size_t>(blockIdx.x) * 3 * P.wp_width * 32 + lane;
  m.wp = P.wp_scratch + static_cast<size_t>(blockIdx.x) * 10 * (P.wp_width + 2) * 32 + lane;
  const bool valid = s < P.num_streams;
  uint64_t end_pos = 0;
  const uint32_t status = DevDecodeModularStream<WT, 32>(P, s, m, P.warp_dims + P.warp_dims_off[blockIdx.x],
                                                         P.warp_chans[blockIdx.x], valid, &end_pos);
  if (valid) {
    P.status[s] = status;
    if (P.end_bits) P.end_bits[s] = end_pos;
  }
}

// Few streams (a batch of lossy frames has only four DC-group chains per 4K frame): the chip is empty and each
// stream is a latency-bound serial chain, so spread them out -- kSparseLanes streams per warp -- and keep each
// stream's row ring and weighted-predictor rows in shared memory instead of global memory.
constexpr uint32_t kSparseLanes = 8;
template <typename WT>
__global__ void __launch_bounds__(32) k_modular_decode_sparse(DevPools P) {
  extern __shared__ int32_t sparse_smem[];
  __shared__ int32_t props_s[kDevMaxProps * 32];
  __shared__ uint32_t div_s[64];
  const uint32_t lane = threadIdx.x;
  for (uint32_t i = lane; i < 64; i += 32) div_s[i] = (1u << 24) / (i + 1);
  __syncwarp();
  const uint32_t s = P.stream0 + blockIdx.x * kSparseLanes + lane;
  DevLaneMem m;
  m.props = props_s + lane;
  m.props_stride = 32;
  m.divlut = div_s;
  m.ring_w = P.wp_width;
  m.lane_stride = kSparseLanes;
  m.ring = sparse_smem + (lane < kSparseLanes ? lane : 0);
  m.wp = sparse_smem + 2 * P.wp_width * kSparseLanes + (lane < kSparseLanes ? lane : 0);
  const bool valid = lane < kSparseLanes && s < P.num_streams;
  const uint32_t b0 = (blockIdx.x * kSparseLanes) / 32;  // loop bounds of the 32-stream bundle these streams belong to (a superset)
  uint64_t end_pos = 0;
  const uint32_t status = DevDecodeModularStream<WT, kSparseLanes>(P, s, m, P.warp_dims + P.warp_dims_off[b0], P.warp_chans[b0], valid, &end_pos);
  if (valid) {
    P.status[s] = status;
    if (P.end_bits) P.end_bits[s] = end_pos;
  }
}

// One warp per stream (jxlb_modular_coop_dev.h): the chains of the lossy path -- DC + AC-metadata of a DC group under
// libjxl's fixed trees -- whose latency per sample sets the time of the whole kernel. Shared memory: the warp's two
// sample rows and the weighted predictor's five error rows.
constexpr uint32_t kCoopWarps = 4;  // per CTA: few CTA slots per SM stay taken while the chains run
template <typename WT>
__global__ void __launch_bounds__(32 * kCoopWarps) k_modular_decode_coop(DevPools P) {
  extern __shared__ int32_t coop_smem[];
  __shared__ uint32_t div_s[64];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 64) div_s[threadIdx.x] = (1u << 24) / (threadIdx.x + 1);
  __syncthreads();
  const uint32_t s = P.coop0 + blockIdx.x * kCoopWarps + warp;
  if (s >= P.stream0) return;
  int32_t* mine = coop_smem + warp * (7 * P.wp_width + 10);
  uint64_t end_pos = 0;
  const uint32_t status = DevDecodeModularStreamCoop<WT>(P, s, mine, mine + 2 * P.wp_width, P.wp_width, div_s, &end_pos);
  if (lane == 0) {
    P.status[s] = status;
    if (P.end_bits) P.end_bits[s] = end_pos;
  }
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
  if (fo.vardct) return;
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
  if (fo.vardct) return;
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

// ------------------------------------------------------------------ VarDCT kernels
// blockIdx.x indexes a flat list of (frame, DC group) pairs.
__global__ void __launch_bounds__(256) k_dc_finish(DevPools P, DevVPools V, const uint2* dcg_list) {
  __shared__ uint32_t occ_s[kDcOccWords];  // one bit per block of the DC group (256 x 256 blocks): covered or not
  __shared__ uint32_t stage_s[kDcStageEntries];
  const uint2 e = dcg_list[blockIdx.x];
  DevDcGroupFinish<2>(P, V, e.x, e.y, threadIdx.x, blockDim.x, blockIdx.x, occ_s, stage_s);
}

__global__ void __launch_bounds__(256) k_dc_smooth(DevVPools V) {
  const DevVFrame& vf = V.frames[blockIdx.y];
  if (vf.skip_dc_smoothing) return;
  const uint32_t n = vf.xblocks * vf.yblocks;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    DevDcSmoothBlock(V, vf, i % vf.xblocks, i / vf.xblocks);
}

// One thread per (frame, group): the group's varblock list for the AC streams (DevBuildBlockList).
__global__ void __launch_bounds__(128) k_block_lists(DevVPools V, uint32_t max_groups, uint32_t num_frames) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max_groups * num_frames) return;
  const DevVFrame& vf = V.frames[i / max_groups];
  const uint32_t g = i % max_groups;
  if (g < vf.xgroups * vf.ygroups) DevBuildBlockList(V, vf, g);
}

// kAcWarps independent warps per CTA, 32 AC streams in lock step each; the streams are latency-bound single warps.
// One warp per CTA is fastest alone, but 1080 one-warp CTAs per 256 frames take the SMs' CTA slots (32 per SM) away
// from the per-pixel kernels of other batches (tools/interference.py): several warps per CTA by default.
template <uint32_t kAcWarps>
__global__ void __launch_bounds__(32 * kAcWarps) k_ac_decode(DevPools P, DevVPools V) {
  __shared__ uint8_t colnz_s[kAcWarps][96 * 32];
  __shared__ uint16_t ctxtab_s[128];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) ctxtab_s[i] = static_cast<uint16_t>(V.upool[V.ctxtab_off + i]);
  for (uint32_t i = lane; i < 96 * 32; i += 32) colnz_s[warp][i] = 0;
  __syncthreads();
  const uint32_t s = (blockIdx.x * kAcWarps + warp) * 32 + lane;
  DevAcLaneMem m;
  m.colnz = colnz_s[warp] + lane;
  m.stride = 32;
  m.freq_ctx = ctxtab_s;
  m.nnz_ctx = ctxtab_s + 64;
  const bool valid = s < V.num_streams;
  const uint32_t status = V.ac_plain_ans ? DevDecodeAcStream<true>(P, V, s, m, valid) : DevDecodeAcStream<false>(P, V, s, m, valid);
  if (valid) V.ac_status[s] = status;
}

// ---- AC decode, one CTA per (frame, pass): the pass's alias tables (cp.async.bulk into shared memory, completion on
// an mbarrier), uint configs and context map are staged once per CTA and every symbol's context -> cluster -> alias
// entry chain then runs on shared-memory loads (the one-warp kernel above takes them from L1 / L2: with 32 lanes from
// 32 different frames its hit rate is 78 %, ncu long_scoreboard 5 cycles per issue). Lane = one group's stream, the
// unit's streams longest first so that the lanes of a warp end together. Plain ANS codes only.
__device__ __forceinline__ uint32_t SmemAddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

struct AcFrameSmemLayout {
  uint32_t alias_bytes, cfg_off, ctx_off, colnz_off, tab_off, bar_off, total;
};
__host__ __device__ inline AcFrameSmemLayout AcFrameLayout(uint32_t alias_entries, uint32_t clusters, uint32_t ctx_bytes, uint32_t warps) {
  AcFrameSmemLayout L;
  L.alias_bytes = alias_entries * 8;
  L.cfg_off = L.alias_bytes;
  L.ctx_off = L.cfg_off + ((clusters * 4 + 15) & ~15u);
  L.colnz_off = L.ctx_off + ((ctx_bytes + 15) & ~15u);
  L.tab_off = L.colnz_off + warps * 96 * 32;
  L.bar_off = L.tab_off + 256;
  L.total = L.bar_off + 16;
  return L;
}

__global__ void __launch_bounds__(256) k_ac_decode_frame(DevPools P, DevVPools V, const DevAcUnit* units, const uint32_t* unit_streams) {
  extern __shared__ __align__(128) uint8_t ac_smem[];
  const DevAcUnit u = units[blockIdx.x];
  const DevVFrame& vf = V.frames[u.frame];
  const DevCode code = P.codes[vf.ac_code[u.pass]];
  const uint32_t warps = blockDim.x >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ctx_bytes = vf.num_histograms * vf.num_ctxs * 495;
  const AcFrameSmemLayout L = AcFrameLayout(code.num_clusters << code.log_alpha_size, code.num_clusters, ctx_bytes, warps);
  DevAlias* alias_s = reinterpret_cast<DevAlias*>(ac_smem);
  uint32_t* cfg_s = reinterpret_cast<uint32_t*>(ac_smem + L.cfg_off);
  uint8_t* ctx_s = ac_smem + L.ctx_off;
  uint8_t* colnz_s = ac_smem + L.colnz_off;
  uint16_t* tab_s = reinterpret_cast<uint16_t*>(ac_smem + L.tab_off);
  const uint32_t bar = SmemAddr(ac_smem + L.bar_off);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // the alias tables of this code: code.alias_off is a multiple of 32 entries (every code's table is
    // clusters << log_alpha entries, log_alpha >= 5), so source, destination and size are 16-byte multiples
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(L.alias_bytes) : "memory");
    const uint8_t* src = reinterpret_cast<const uint8_t*>(P.alias + code.alias_off);
    for (uint32_t off = 0; off < L.alias_bytes; off += 32768) {
      const uint32_t n = min(32768u, L.alias_bytes - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       SmemAddr(ac_smem + off)),
                   "l"(src + off), "r"(n), "r"(bar)
                   : "memory");
    }
  }
  // meanwhile: uint configs, context map, the two context tables, the non-zero columns
  for (uint32_t i = threadIdx.x; i < code.num_clusters; i += blockDim.x) cfg_s[i] = P.cfg[code.cfg_off + i];
  const uint8_t* ctx_g = V.cpool + vf.ctx_map_off[u.pass];
  for (uint32_t i = threadIdx.x; i < ctx_bytes; i += blockDim.x) ctx_s[i] = ctx_g[i];
  for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) tab_s[i] = static_cast<uint16_t>(V.upool[V.ctxtab_off + i]);
  for (uint32_t i = threadIdx.x; i < warps * 96 * 32; i += blockDim.x) colnz_s[i] = 0;
  {  // wait for the bulk copies (phase 0)
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(bar)
          : "memory");
    }
  }
  __syncthreads();
  DevAcLaneMem m;
  m.colnz = colnz_s + warp * 96 * 32 + lane;
  m.stride = 32;
  m.freq_ctx = tab_s;
  m.nnz_ctx = tab_s + 64;
  m.alias_s = alias_s;
  m.cfg_s = cfg_s;
  m.ctx_map_s = ctx_s;
  // (a unit with more streams than threads: every warp takes another 32 after it finished -- 8K x 8K frames and beyond)
  for (uint32_t base = warp * 32; base < u.count; base += warps * 32) {
    const bool valid = base + lane < u.count;
    const uint32_t s = valid ? unit_streams[u.first + base + lane] : 0;
    const uint32_t status = DevDecodeAcStream<true, true>(P, V, s, m, valid);
    if (valid) V.ac_status[s] = status;
    __syncwarp();
    if (base + warps * 32 < u.count)
      for (uint32_t i = lane; i < 96 * 32; i += 32) colnz_s[warp * 96 * 32 + i] = 0;
    __syncwarp();
  }
}

constexpr uint32_t kSortKeys = 64;  // k_dequant_idct's list order: special 8x8 transforms by strategy (keys 0 .. 31), then the rest (32 .. 63)
constexpr uint32_t kIdctThreads = 128;                                  // 4 warps, one small varblock each at a time
constexpr uint32_t kIdctSmemFloats = (kIdctThreads / 32) * kFastBufFloats;  // 51.7 KB: four CTAs per SM
constexpr uint32_t kMidThreads = 256;
constexpr uint32_t kMidSmemFloats = kFastBufFloats64;                   // 49 KB: three padded 64x64 channels

// blockIdx.x = group, blockIdx.y = frame - frame0. Varblocks of up to 32x32 pixels: one warp each, handed out
// dynamically from a list of the group's varblocks that the CTA compacts first (cell | strategy << 10). A warp asks
// for its next varblock and requests that block's token ranges and raw quant (seven words, one per lane) before it
// starts on the current one, so the chain list -> metadata -> tokens -> tables of a varblock is not waited for link by
// link. Plain DCTs take the register-IDCT fast path, the 8x8 special transforms the generic one.
__global__ void __launch_bounds__(kIdctThreads) k_dequant_idct(DevVPools V, uint32_t frame0, uint32_t* has_mid, uint32_t* mid_count,
                                                                uint2* mid_list) {
  extern __shared__ float idct_smem[];
  __shared__ uint32_t next_s, count_s, next8_s, count8_s, left8_s, nextS_s, countS_s;
  __shared__ uint32_t bucket_s[2 * kSortKeys];  // entries per sort key, then the key's fill position
  // Varblocks of the group. list8_s, from the front: 8x8 DCTs of single-pass frames (four per warp at a time). list_s:
  // everything else as it is found -- special 8x8 transforms of single-pass frames from the back, the rest from the
  // front -- then both sorted by strategy into the free tail of list8_s (every first block is in exactly one list:
  // 1024 entries hold them all): [.. | the rest | special], read backwards from 1023. Sorted, the warps of a CTA run
  // the same transform code at the same time (the kernel was waiting for instruction fetches: ncu no_instruction
  // stall 13.9 cycles per issue with the lists in cell order, profiles/r2_ncu_k_dequant_idct_before_sort.txt).
  __shared__ uint16_t list_s[1024], list8_s[1024];
  const DevVFrame& vf = V.frames[frame0 + blockIdx.y];
  const uint32_t g = blockIdx.x;
  if (g >= vf.xgroups * vf.ygroups) return;
  const uint32_t x0 = (g % vf.xgroups) * 32, y0 = (g / vf.xgroups) * 32;
  const uint32_t xs = min(32u, vf.xblocks - x0), ys = min(32u, vf.yblocks - y0);
  const uint8_t* acs = V.barena + vf.acs;
  if (threadIdx.x == 0) next_s = count_s = next8_s = count8_s = left8_s = nextS_s = countS_s = 0;
  if (threadIdx.x < 2 * kSortKeys) bucket_s[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool single_pass = vf.num_passes == 1;
  bool mid = false;
  for (uint32_t cell = threadIdx.x; cell < 1024; cell += kIdctThreads) {
    const uint32_t bx = cell & 31, by = cell >> 5;
    if (bx >= xs || by >= ys) continue;
    const uint8_t a = acs[static_cast<size_t>(y0 + by) * vf.xblocks + x0 + bx];
    if (!(a & 1) || a == 0xFF) continue;
    const StrategyInfo si = UnpackStrategyInfo(V.upool[V.sinfo_off + (a >> 1)]);
    if (static_cast<uint32_t>(si.cx) * si.cy > 16) {
      mid = true;
      // 64x32 / 32x64 / 64x64: onto the wave's list for k_idct_mid (frame, block position)
      if (static_cast<uint32_t>(si.cx) * si.cy <= 64)
        mid_list[atomicAdd(mid_count, 1u)] = make_uint2(frame0 + blockIdx.y, (y0 + by) * vf.xblocks + x0 + bx);
      continue;
    }
    if ((a >> 1) == 0 && single_pass) {
      list8_s[atomicAdd(&count8_s, 1u)] = static_cast<uint16_t>(cell);
    } else if (single_pass && !si.plain_dct && static_cast<uint32_t>(si.cx) * si.cy == 1) {
      list_s[1023 - atomicAdd(&countS_s, 1u)] = static_cast<uint16_t>(cell | (static_cast<uint32_t>(a >> 1) << 10));
      atomicAdd(&bucket_s[(a >> 1)], 1u);
    } else {
      list_s[atomicAdd(&count_s, 1u)] = static_cast<uint16_t>(cell | (static_cast<uint32_t>(a >> 1) << 10));
      atomicAdd(&bucket_s[32 + (a >> 1)], 1u);
    }
  }
  if (mid) *has_mid = 1;  // some frame of the batch needs k_idct_mid / k_idct_big
  __syncthreads();
  const uint32_t totalS = countS_s, total = count_s;
  if (totalS + total) {  // counting sort by (special first, strategy): position p of the order lives at list8_s[1023 - p]
    if (threadIdx.x == 0) {
      uint32_t off = 0;
      for (uint32_t b = 0; b < kSortKeys; b++) {
        bucket_s[kSortKeys + b] = off;
        off += bucket_s[b];
      }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < totalS + total; i += kIdctThreads) {
      const bool special = i < totalS;
      const uint16_t e = special ? list_s[1023 - i] : list_s[i - totalS];
      list8_s[1023 - atomicAdd(&bucket_s[kSortKeys + (special ? 0 : 32) + (e >> 10)], 1u)] = e;
    }
    __syncthreads();
  }
  float* wbuf = idct_smem + warp * kFastBufFloats;
  const size_t nb = static_cast<size_t>(vf.xblocks) * vf.yblocks;
  // 8x8 DCT blocks, four per warp at a time: the eight lanes of a quarter warp run the varblock function on their own
  // block (tid = lane & 7 of 8 threads, own slice of the warp's buffer). All four follow the same control flow up to
  // loop trip counts, so the function's warp-wide synchronisation points are reached by every lane. Lane 8q + l
  // (l < 7) fetches word l of block q's metadata one quad ahead. What does not fill a quad goes one block at a time.
  {
    const uint32_t total8 = count8_s, quads = total8 / 4, sub = lane >> 3, t8 = lane & 7;
    float* qbuf = wbuf + sub * 296;  // 3 * 8 * 9 floats of coefficients (+ the LLF scratch the 8x8 DCT never touches)
    auto fetch8 = [&](uint32_t* cell_out) -> uint32_t {
      uint32_t q = 0;
      if (lane == 0) q = atomicAdd(&next8_s, 1u);
      q = __shfl_sync(0xFFFFFFFFu, q, 0);
      if (q >= quads) {
        *cell_out = 0xFFFFFFFFu;
        return 0;
      }
      const uint32_t cell = list8_s[q * 4 + sub];
      *cell_out = cell;
      const size_t pos = static_cast<size_t>(y0 + (cell >> 5)) * vf.xblocks + x0 + (cell & 31);
      uint32_t word = 0;
      if (t8 < 3) word = V.uarena[vf.tok_start + t8 * nb + pos];
      else if (t8 < 6) word = V.uarena[vf.tok_count + (t8 - 3) * nb + pos];
      else if (t8 == 6) word = reinterpret_cast<const uint16_t*>(V.barena + vf.rawq)[pos];
      return word;
    };
    uint32_t cell, word = fetch8(&cell);
    while (cell != 0xFFFFFFFFu) {
      uint32_t next_cell;
      const uint32_t next_word = fetch8(&next_cell);
      DevBlockMeta meta;
      for (uint32_t c = 0; c < 3; c++) {
        meta.start[c] = __shfl_sync(0xFFFFFFFFu, word, (lane & 24) + c);
        meta.count[c] = __shfl_sync(0xFFFFFFFFu, word, (lane & 24) + 3 + c);
      }
      meta.rawq = __shfl_sync(0xFFFFFFFFu, word, (lane & 24) + 6);
      DevVarblockFast<1, 32>(V, vf, x0 + (cell & 31), y0 + (cell >> 5), 0, qbuf, t8, 8, &meta);
      cell = next_cell;
      word = next_word;
    }
    // the up to three blocks that do not fill a quad, one warp each
    for (;;) {
      uint32_t k = 0;
      if (lane == 0) k = atomicAdd(&left8_s, 1u);
      k = __shfl_sync(0xFFFFFFFFu, k, 0);
      if (quads * 4 + k >= total8) break;
      const uint32_t c8 = list8_s[quads * 4 + k];
      DevVarblockFast<1, 32>(V, vf, x0 + (c8 & 31), y0 + (c8 >> 5), 0, wbuf, lane, 32);
    }
  }
  // special 8x8 transforms, eight per warp at a time: a group of four lanes per varblock (dequantisation from the
  // tokens by all four, then one lane per channel runs the transform)
  if (totalS) {
    const uint32_t sub = lane >> 2, t4 = lane & 3;
    float* sbuf = wbuf + sub * 3 * kSpecialChStride;
    for (;;) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&nextS_s, 8u);
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (base >= totalS) break;
      const bool active = base + sub < totalS;
      const uint32_t e = active ? list8_s[1023 - (base + sub)] : 0;
      const uint32_t cell = e & 1023;
      DevBlockMeta meta{};
      if (active) meta = DevLoadBlockMeta(V, vf, static_cast<size_t>(y0 + (cell >> 5)) * vf.xblocks + x0 + (cell & 31));
      DevVarblockSpecial<1>(V, vf, x0 + (cell & 31), y0 + (cell >> 5), active ? e >> 10 : 1, sbuf, t4, 4, meta, active);
    }
  }
  // lane l < 6 holds tok_start / tok_count of channel l % 3 (l < 3: start), lane 6 the raw quant of the varblock
  auto fetch = [&](uint32_t* entry) -> uint32_t {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(&next_s, 1u);
    i = __shfl_sync(0xFFFFFFFFu, i, 0);
    if (i >= total) {
      *entry = 0xFFFFFFFFu;
      return 0;
    }
    const uint32_t e = list8_s[1023 - (totalS + i)];
    *entry = e;
    const uint32_t cell = e & 1023;
    const size_t pos = static_cast<size_t>(y0 + (cell >> 5)) * vf.xblocks + x0 + (cell & 31);
    uint32_t word = 0;
    if (single_pass) {
      if (lane < 3) word = V.uarena[vf.tok_start + lane * nb + pos];
      else if (lane < 6) word = V.uarena[vf.tok_count + (lane - 3) * nb + pos];
      else if (lane == 6) word = reinterpret_cast<const uint16_t*>(V.barena + vf.rawq)[pos];
    }
    return word;
  };
  uint32_t entry, word = fetch(&entry);
  while (entry != 0xFFFFFFFFu) {
    uint32_t next_entry;
    const uint32_t next_word = fetch(&next_entry);
    const uint32_t cell = entry & 1023, strategy = entry >> 10;
    const uint32_t bx = x0 + (cell & 31), by = y0 + (cell >> 5);
    const StrategyInfo si = UnpackStrategyInfo(V.upool[V.sinfo_off + strategy]);
    if (si.plain_dct) {
      if (single_pass) {
        DevBlockMeta meta;
        for (uint32_t c = 0; c < 3; c++) {
          meta.start[c] = __shfl_sync(0xFFFFFFFFu, word, c);
          meta.count[c] = __shfl_sync(0xFFFFFFFFu, word, 3 + c);
        }
        meta.rawq = __shfl_sync(0xFFFFFFFFu, word, 6);
        DevVarblockFast<1, 32>(V, vf, bx, by, strategy, wbuf, lane, 32, &meta);
      } else {
        DevVarblockFast<1, 32>(V, vf, bx, by, strategy, wbuf, lane, 32);
      }
    } else {
      DevVarblock<1>(V, vf, bx, by, strategy, wbuf, lane, 32);
    }
    entry = next_entry;
    word = next_word;
  }
}

// 64x32, 32x64 and 64x64 varblocks: the whole CTA per varblock, one 64-point register IDCT per thread and line.
// The CTAs stride over the list of such varblocks that k_dequant_idct made for the wave, so that a smooth group with
// many of them does not become one CTA's serial tail.
__global__ void __launch_bounds__(kMidThreads) k_idct_mid(DevVPools V, const uint32_t* mid_count, const uint2* mid_list) {
  extern __shared__ float idct_smem[];
  const uint32_t n = *mid_count;
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const uint2 e = mid_list[i];
    const DevVFrame& vf = V.frames[e.x];
    const uint8_t a = V.barena[vf.acs + e.y];
    DevVarblockFast<2, 64>(V, vf, e.y % vf.xblocks, e.y / vf.xblocks, a >> 1, idct_smem, threadIdx.x, kMidThreads);
  }
}

// Persistent CTAs with 4 * 65536 floats of global scratch each: varblocks above 64x64 pixels.
__global__ void __launch_bounds__(256) k_idct_big(DevVPools V, uint32_t frame0, uint32_t num_frames, float* scratch,
                                                  const uint32_t* has_mid) {
  __shared__ uint32_t list_s[64], count_s;
  if (*has_mid == 0) return;
  float* buf = scratch + static_cast<size_t>(blockIdx.x) * 4 * 65536;
  for (uint32_t f = 0; f < num_frames; f++) {
    const DevVFrame& vf = V.frames[frame0 + f];
    const uint8_t* acs = V.barena + vf.acs;
    for (uint32_t g = blockIdx.x; g < vf.xgroups * vf.ygroups; g += gridDim.x) {
      const uint32_t x0 = (g % vf.xgroups) * 32, y0 = (g / vf.xgroups) * 32;
      const uint32_t xs = min(32u, vf.xblocks - x0), ys = min(32u, vf.yblocks - y0);
      __syncthreads();
      if (threadIdx.x == 0) count_s = 0;
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < xs * ys; i += blockDim.x) {
        const uint8_t a = acs[static_cast<size_t>(y0 + i / xs) * vf.xblocks + x0 + i % xs];
        if (!(a & 1) || a == 0xFF) continue;
        const StrategyInfo si = UnpackStrategyInfo(V.upool[V.sinfo_off + (a >> 1)]);
        if (static_cast<uint32_t>(si.cx) * si.cy <= 64) continue;
        const uint32_t k = atomicAdd(&count_s, 1u);
        if (k < 64) list_s[k] = i | (static_cast<uint32_t>(a >> 1) << 16);  // at most 4 such varblocks fit a group
      }
      __syncthreads();
      const uint32_t n = min(count_s, 64u);
      for (uint32_t k = 0; k < n; k++) {
        const uint32_t i = list_s[k] & 0xFFFF, strategy = list_s[k] >> 16;
        DevVarblock<2>(V, vf, x0 + i % xs, y0 + i / xs, strategy, buf, threadIdx.x, blockDim.x);
        __threadfence_block();
      }
    }
  }
}

// Render stages: blockDim (32, 8), grid (x tiles, y tiles, frames of the wave).
// `skip_fused`: frames that k_render_fused handles (DevRenderFused) are left alone.
__global__ void __launch_bounds__(256) k_gaborish(DevVPools V, uint32_t frame0, uint32_t in_set, uint32_t out_set, uint32_t skip_fused) {
  const DevVFrame& vf = V.frames[frame0 + blockIdx.z];
  const uint32_t x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (!vf.gab || x >= vf.xsize || y >= vf.ysize || (skip_fused && DevRenderFused(vf))) return;
  const bool interior = x >= 1 && y >= 1 && x + 1 < vf.xsize && y + 1 < vf.ysize;
  for (uint32_t c = 0; c < 3; c++) {
    if (interior) {
      DevGaborishPixel<true>(V, vf, in_set, out_set, c, static_cast<int>(x), static_cast<int>(y));
    } else {
      DevGaborishPixel<false>(V, vf, in_set, out_set, c, static_cast<int>(x), static_cast<int>(y));
    }
  }
}

// `sets` packs, per frame class, which plane set holds the input: frames with / without
// Gaborish (and different EPF iteration counts) take different numbers of ping-pong steps, so the
// current set is derived per frame from its own flags.
__device__ __forceinline__ uint32_t SetBeforeStage(const DevVFrame& vf, uint32_t stage) {
  uint32_t set = vf.gab ? 1u : 0u;
  if (stage > 0 && vf.epf_iters >= 3) set ^= 1;  // stage 0 ran
  if (stage > 1 && vf.epf_iters >= 1) set ^= 1;  // stage 1 ran
  if (stage > 2 && vf.epf_iters >= 2) set ^= 1;  // stage 2 ran
  return set;
}

__global__ void __launch_bounds__(256) k_epf(DevVPools V, uint32_t frame0, uint32_t stage, uint32_t skip_fused) {
  const DevVFrame& vf = V.frames[frame0 + blockIdx.z];
  const bool runs = vf.epf_iters > 0 && !(stage == 0 && vf.epf_iters < 3) && !(stage == 2 && vf.epf_iters < 2) &&
                    !(skip_fused && DevRenderFused(vf));
  const uint32_t x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (!runs || x >= vf.xsize || y >= vf.ysize) return;
  const uint32_t set = SetBeforeStage(vf, stage);
  if (x >= 3 && y >= 3 && x + 3 < vf.xsize && y + 3 < vf.ysize) {
    DevEpfPixel<true>(V, vf, stage, set, set ^ 1, static_cast<int>(x), static_cast<int>(y));
  } else {
    DevEpfPixel<false>(V, vf, stage, set, set ^ 1, static_cast<int>(x), static_cast<int>(y));
  }
}

// Gaborish + EPF + colour + output write of one 64x32 tile with the intermediate planes in shared memory
// (DevRenderTile): grid (x tiles, y tiles, frames of the wave), 6 * cap floats of dynamic shared memory.
__global__ void __launch_bounds__(256) k_render_fused(DevVPools V, uint32_t frame0, uint32_t cap, uint32_t stride) {
  extern __shared__ float rt_sm[];
  const DevVFrame& vf = V.frames[frame0 + blockIdx.z];
  const int tx0 = blockIdx.x * kRtW, ty0 = blockIdx.y * kRtH;
  const int xsize = static_cast<int>(vf.xsize), ysize = static_cast<int>(vf.ysize);
  if (!DevRenderFused(vf) || tx0 >= xsize || ty0 >= ysize) return;
  const int H = static_cast<int>(DevRenderHalo(vf.gab, vf.epf_iters));
  const bool interior = tx0 - H >= 0 && ty0 - H >= 0 && tx0 + kRtW + H <= xsize && ty0 + kRtH + H <= ysize;
  if (stride == kRtStrideSmall) {  // (the stride of the batch: its largest halo decides)
    if (interior) {
      DevRenderTile<2, true, kRtStrideSmall>(V, vf, tx0, ty0, threadIdx.x, blockDim.x, rt_sm, cap);
    } else {
      DevRenderTile<2, false, kRtStrideSmall>(V, vf, tx0, ty0, threadIdx.x, blockDim.x, rt_sm, cap);
    }
  } else if (interior) {
    DevRenderTile<2, true, kRtStrideLarge>(V, vf, tx0, ty0, threadIdx.x, blockDim.x, rt_sm, cap);
  } else {
    DevRenderTile<2, false, kRtStrideLarge>(V, vf, tx0, ty0, threadIdx.x, blockDim.x, rt_sm, cap);
  }
}

// Reference-only frames: Modular samples -> float XYB planes. blockIdx.y = reference frame.
__global__ void __launch_bounds__(256) k_ref_frames(DevPools P, DevVPools V) {
  const DevRefFrame& rf = V.ref_frames[blockIdx.y];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < rf.w * rf.h; i += gridDim.x * blockDim.x)
    DevRefFrameSample(P, V, rf, i);
}

// Patches of one frame per CTA, one after the other (a later patch may cover an earlier one).
__global__ void __launch_bounds__(256) k_patches(DevVPools V, uint32_t frame0) {
  const DevVFrame& vf = V.frames[frame0 + blockIdx.x];
  const uint32_t set = SetBeforeStage(vf, 3);
  for (uint32_t k = 0; k < vf.patch_count; k++) {
    const DevPatch& p = V.patches[vf.patch_begin + k];
    for (uint32_t i = threadIdx.x; i < p.xsize * p.ysize; i += blockDim.x) DevPatchPixel(V, vf, p, set, i % p.xsize, i / p.xsize);
    __syncthreads();
  }
}

// Upsampled frames: one thread per output pixel of the upsampled planes (DevUpsamplePixel).
__global__ void __launch_bounds__(256) k_upsample(DevVPools V, uint32_t frame0) {
  const DevVFrame& vf = V.frames[frame0 + blockIdx.z];
  const uint32_t x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (vf.upsampling <= 1 || x >= vf.xsize * vf.upsampling || y >= vf.ysize * vf.upsampling) return;
  DevUpsamplePixel(V, vf, SetBeforeStage(vf, 3), x, y);
}

__global__ void __launch_bounds__(256) k_color_write(DevVPools V, uint32_t frame0, uint32_t skip_fused) {
  const DevVFrame& vf = V.frames[frame0 + blockIdx.z];
  const uint32_t y = blockIdx.y * 8 + threadIdx.y;
  if (skip_fused && DevRenderFused(vf)) return;
  if (vf.upsampling > 1) {  // the image's own size, from the upsampled planes
    if (y >= vf.up_ysize) return;
    for (uint32_t k = 0; k < 4; k++) {
      const uint32_t x = (blockIdx.x * 4 + k) * 32 + threadIdx.x;
      if (x < vf.up_xsize) DevColorPixel(V, vf, 0, x, y);
    }
    return;
  }
  if (y >= vf.ysize) return;
  const uint32_t set = SetBeforeStage(vf, 3);
  if (vf.out_type == 2 && vf.out_channels == 3 && vf.out_stride % 4 == 0 && vf.orient == 0 && !vf.has_splines) {
    // RGB8: each thread converts 4 consecutive pixels and writes 12 bytes as three words
    const uint32_t x = (blockIdx.x * 32 + threadIdx.x) * 4;
    if (x + 4 <= vf.xsize) {
      DevColorPixelsRgb8x4(V, vf, set, x, y);
    } else {
      for (uint32_t i = x; i < vf.xsize; i++) DevColorPixel(V, vf, set, i, y);
    }
    return;
  }
  for (uint32_t k = 0; k < 4; k++) {  // grid.x covers 128 pixels per CTA row
    const uint32_t x = (blockIdx.x * 4 + k) * 32 + threadIdx.x;
    if (x < vf.xsize) DevColorPixel(V, vf, set, x, y);
  }
}

// ------------------------------------------------------------------ runtime
// Pinned staging for the tables of a batch: a pageable source makes cudaMemcpyAsync stage and wait copy by copy (40
// pools: 83 ms per 256-frame batch, measured); copied here first, the uploads are enqueued back to back. Blocks are
// kept across batches; Reset() after the stream that read them was synchronised.
struct PinnedStage {
  struct Block { uint8_t* p; size_t cap; };
  std::vector<Block> blocks;
  size_t cur = 0, used = 0;
  ~PinnedStage() {
    for (Block& b : blocks) cudaFreeHost(b.p);
  }
  void Reset() { cur = 0; used = 0; }
  uint8_t* Take(size_t bytes) {
    bytes = (bytes + 255) & ~size_t{255};
    while (cur < blocks.size() && used + bytes > blocks[cur].cap) {
      cur++;
      used = 0;
    }
    if (cur == blocks.size()) {
      Block b{nullptr, std::max<size_t>(bytes, size_t{8} << 20)};
      if (cudaHostAlloc(reinterpret_cast<void**>(&b.p), b.cap, cudaHostAllocDefault) != cudaSuccess) return nullptr;
      blocks.push_back(b);
      used = 0;
    }
    uint8_t* p = blocks[cur].p + used;
    used += bytes;
    return p;
  }
};

// Pinned host words for the status read-back of a run: into pageable memory the asynchronous copy turns into a blocking
// one that waits inside the driver for the run's kernels (other threads' launches and uploads queue up behind it).
struct PinnedWords {
  uint32_t* p = nullptr;
  size_t n = 0, cap = 0;
  ~PinnedWords() {
    if (p) cudaFreeHost(p);
  }
  void resize(size_t count) {
    if (count > cap) {
      if (p) cudaFreeHost(p);
      p = nullptr;
      cap = 0;
      const size_t want = count + count / 4 + 64;
      if (cudaHostAlloc(reinterpret_cast<void**>(&p), want * 4, cudaHostAllocDefault) == cudaSuccess) cap = want;
    }
    n = cap >= count ? count : 0;
  }
  uint32_t* data() { return p; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  uint32_t operator[](size_t i) const { return p[i]; }
};

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
  cudaError_t Upload(const std::vector<T>& v, cudaStream_t s, PinnedStage* stage = nullptr) {
    cudaError_t e = Alloc(v.size());
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    const void* src = v.data();
    if (stage && v.size() * sizeof(T) >= 16384) {  // (smaller ones travel inside the command)
      if (uint8_t* st = stage->Take(v.size() * sizeof(T))) {
        std::memcpy(st, v.data(), v.size() * sizeof(T));
        src = st;
      }
    }
    return cudaMemcpyAsync(p, src, v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
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

enum KernelClass {
  kKModular = 0, kKGroupPrograms, kKFrameLevels, kKWriteOutput, kKDcFinish, kKAcDecode, kKDequantIdct, kKFilters,
  kKColorWrite, kNumKernelClasses
};

// ------------------------------------------------------------------ SM partition (CUDA green contexts)
// The entropy kernels (Modular chains, DC finish, AC decode) are latency-bound single warps that sit on their SMs for
// 50 - 150 ms and park registers / shared memory there; the per-pixel kernels of the other handles in flight then find
// less room on every SM (tools/interference.py). With JXLB200_ENTROPY_SMS=n the SMs are split once per device into an
// entropy partition of n SMs and a pixel partition of the rest (cuDevSmResourceSplitByCount -> cuGreenCtxCreate), and
// every handle launches its entropy kernels into a stream of the first and its per-pixel kernels into a stream of the
// second. The driver entry points are fetched at run time (cudaGetDriverEntryPoint): the library does not link libcuda.
struct SmPartition {
  bool ok = false;
  CUgreenCtx ctx[2] = {nullptr, nullptr};  // 0: entropy, 1: pixel
  uint32_t sms[2] = {0, 0};
  CUresult (*stream_create)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
};

static SmPartition* GetSmPartition(int device) {
  static std::mutex mu;
  static std::map<int, SmPartition> parts;
  std::lock_guard<std::mutex> lock(mu);
  auto it = parts.find(device);
  if (it != parts.end()) return it->second.ok ? &it->second : nullptr;
  SmPartition& p = parts[device];
  const char* e = std::getenv("JXLB200_ENTROPY_SMS");
  const unsigned want = e ? std::atoi(e) : 0;
  if (want == 0) return nullptr;
  auto entry = [](const char* name) -> void* {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) return nullptr;
    return fn;
  };
  auto dev_get = reinterpret_cast<CUresult (*)(CUdevice*, int)>(entry("cuDeviceGet"));
  auto get_res = reinterpret_cast<CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType)>(entry("cuDeviceGetDevResource"));
  auto split = reinterpret_cast<CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                                             unsigned int)>(entry("cuDevSmResourceSplitByCount"));
  auto gen_desc = reinterpret_cast<CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int)>(entry("cuDevResourceGenerateDesc"));
  auto ctx_create = reinterpret_cast<CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int)>(entry("cuGreenCtxCreate"));
  p.stream_create = reinterpret_cast<CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int)>(entry("cuGreenCtxStreamCreate"));
  if (!dev_get || !get_res || !split || !gen_desc || !ctx_create || !p.stream_create) return nullptr;
  cudaFree(nullptr);  // the primary context exists
  CUdevice dev;
  CUdevResource all, part[2];
  unsigned int groups = 1;
  if (dev_get(&dev, device) != CUDA_SUCCESS || get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return nullptr;
  if (split(&part[0], &groups, &all, &part[1], 0, want) != CUDA_SUCCESS || groups != 1 || part[1].sm.smCount == 0) return nullptr;
  for (int i = 0; i < 2; i++) {
    CUdevResourceDesc desc;
    if (gen_desc(&desc, &part[i], 1) != CUDA_SUCCESS) return nullptr;
    if (ctx_create(&p.ctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return nullptr;
    p.sms[i] = part[i].sm.smCount;
  }
  std::fprintf(stderr, "jxl_b200: SM partition on device %d: %u SMs entropy kernels, %u SMs per-pixel kernels\n", device, p.sms[0],
               p.sms[1]);
  p.ok = true;
  return &p;
}

struct JxlB200Decoder {
  int device = 0;
  cudaStream_t stream = nullptr;
  // SM partition (JXLB200_ENTROPY_SMS): this handle's streams in the entropy / pixel green contexts and the events that
  // order caller stream -> entropy kernels -> per-pixel kernels -> caller stream
  cudaStream_t part_stream[2] = {nullptr, nullptr};
  cudaEvent_t part_ev[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t pix_done = nullptr;  // end of this handle's last per-pixel phase (PixelTurn)
  bool keep_orientation = false;   // JxlB200DecoderSetKeepOrientation: for the batches that follow
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
  DevBuf<DevProgram> d_group_programs, d_levels, d_late_group_programs, d_late_levels;
  DevBuf<uint64_t> d_chain_pos;   // DevStream::chain_slot / DevAcStream::chain_slot
  DevBuf<float> d_spl_seg;        // splines of Modular frames
  DevBuf<uint32_t> d_spl_idx;
  std::vector<size_t> late_level_off;
  DevBuf<DevFrameOut> d_frames;
  DevBuf<int32_t> d_arena, d_wp, d_ring;
  // pinned staging for the codestream bytes of a batch (grows, kept across batches). Two of them, used in turn: the
  // batch planned ahead (JxlB200DecoderPlanBatch) fills one while the upload of the current batch may still read the other
  uint8_t* h_bytes[2] = {nullptr, nullptr};
  size_t h_bytes_cap[2] = {0, 0};
  int h_turn = 0;
  PinnedStage stage;                     // UploadPlan's pinned staging of the tables
  std::unique_ptr<BatchPlan> next_plan;  // JxlB200DecoderPlanBatch -> JxlB200DecoderCommitPlan
  PixelFormat next_fmt;
  // JxlB200DecoderRunToHost: the frames of a wave are copied to the caller's host buffers on `copy_stream` as soon as
  // the wave's last kernel is through (event), while the next waves compute
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_ev;
  cudaEvent_t copy_done = nullptr;
  cudaEvent_t sync_ev = nullptr;  // BlockingSync
  bool copy_pending = false;
  std::vector<void*> host_dsts;
  DevBuf<uint64_t> d_end_bits;    // probe launches only
  uint32_t probe_launches = 0;    // Modular decode launches made while planning (probe rounds)
  // VarDCT
  DevBuf<DevVFrame> d_vframes;
  DevBuf<DevAcStream> d_ac_streams;
  DevBuf<DevAcUnit> d_ac_units;
  DevBuf<uint32_t> d_ac_unit_streams;
  uint32_t ac_frame_smem = 0, ac_frame_threads = 0;  // k_ac_decode_frame's launch shape; 0: the one-warp kernel
  DevBuf<float> d_fpool, d_farena, d_big_scratch;
  DevBuf<uint16_t> d_opool, d_lut;
  DevBuf<uint8_t> d_cpool, d_barena;
  DevBuf<uint32_t> d_upool, d_uarena, d_tokens, d_ac_status, d_ac_used, d_dc_status;
  DevBuf<uint2> d_mid_list;  // varblocks of 64x32 ... 64x64 pixels of the wave in flight (k_dequant_idct -> k_idct_mid)
  DevBuf<uint2> d_dcg_list;
  DevBuf<DevPatch> d_patches;
  DevBuf<DevRefFrame> d_ref_frames;
  std::vector<uint2> dcg_list;
  PinnedWords h_ac_status, h_ac_used, h_dc_status;
  DevVPools vpools{};
  uint32_t max_groups = 0, max_xsize = 0, max_ysize = 0, max_blocks = 0;
  uint32_t max_up_xsize = 0, max_up_ysize = 0;  // largest frame after upsampling
  bool any_gab = false;
  uint32_t max_epf = 0;
  bool fused_render = std::getenv("JXLB200_UNFUSED_RENDER") == nullptr;
  std::vector<size_t> level_off;  // offset of each level inside d_levels
  PinnedWords h_status;
  bool uniform_rgba8 = false;
  bool any_modular_frame = false;
  uint32_t launches = 0;
  DevPools pools{};
  // optional per-kernel timing (CUDA events on the launching stream)
  uint32_t phase_mask = 0xFFFFFFFFu;  // profiling hook (JxlB200DecoderSetPhaseMask): kernel classes that Run launches
  bool profiling = false;
  struct Timed { cudaEvent_t a, b; int cls; };
  std::vector<Timed> timed;            // recorded, not yet folded
  std::vector<cudaEvent_t> free_events;
  double kernel_ms[kNumKernelClasses] = {};
  uint32_t profiled_runs = 0;
  static constexpr uint32_t kBigCtas = 148;
};

static cudaEvent_t GetEvent(JxlB200Decoder* dec) {
  if (!dec->free_events.empty()) {
    cudaEvent_t e = dec->free_events.back();
    dec->free_events.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

static void FoldEvents(JxlB200Decoder* dec) {
  for (auto& t : dec->timed) {
    float ms = 0;
    if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess)
      dec->kernel_ms[t.cls] += ms;
    dec->free_events.push_back(t.a);
    dec->free_events.push_back(t.b);
  }
  dec->timed.clear();
}

// Brackets the launches of one kernel class with events when profiling is on.
struct ScopedTimer {
  JxlB200Decoder* dec;
  cudaStream_t s;
  cudaEvent_t b = nullptr;
  int cls;
  ScopedTimer(JxlB200Decoder* d, cudaStream_t st, int c) : dec(d), s(st), cls(c) {
    if (!dec->profiling) return;
    cudaEvent_t a = GetEvent(dec);
    b = GetEvent(dec);
    cudaEventRecord(a, s);
    dec->timed.push_back({a, b, cls});
  }
  ~ScopedTimer() {
    if (b) cudaEventRecord(b, s);
  }
};

// Per-pixel phases of the handles of one device take turns, in the order in which their runs were enqueued: each waits
// for the end of the previously enqueued one. The entropy kernels of a run are chains of dependent instructions on a
// few warps -- they need time, not the machine -- while the per-pixel kernels fill every SM; when several handles are
// driven in lock step (a loop that enqueues one run after the other) their phases line up, all entropy phases overlap
// each other and then all per-pixel phases fight for the SMs. With the turns the per-pixel phases run back to back and
// every other handle's entropy phase runs underneath them. Measured (profiles/r2_pixel_turns.txt): the per-pixel phase then
// takes 2.2 x its time alone, because the entropy kernels underneath hold registers that its CTAs need -- the step is
// bound by register-file capacity either way and does not get shorter. Opt-in: JXLB200_PIXEL_TURNS=1.
struct PixelTurn {
  std::mutex mu;
  cudaEvent_t last = nullptr;
  JxlB200Decoder* owner = nullptr;
};
static PixelTurn g_pixel_turn[64];
static bool PixelTurnsOn() {
  static const bool on = std::getenv("JXLB200_PIXEL_TURNS") && std::atoi(std::getenv("JXLB200_PIXEL_TURNS")) != 0;
  return on;
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
  if (device < 64 && cudaEventCreateWithFlags(&dec->pix_done, cudaEventDisableTiming) != cudaSuccess) dec->pix_done = nullptr;
  if (SmPartition* sp = GetSmPartition(device)) {
    bool ok = true;
    for (int i = 0; i < 2; i++) {
      CUstream st = nullptr;
      ok = ok && sp->stream_create(&st, sp->ctx[i], CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS;
      dec->part_stream[i] = reinterpret_cast<cudaStream_t>(st);
    }
    for (int i = 0; i < 3; i++) ok = ok && cudaEventCreateWithFlags(&dec->part_ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) dec->part_stream[0] = dec->part_stream[1] = nullptr;  // (falls back to one stream)
  }
  cudaFuncSetAttribute(k_dequant_idct, cudaFuncAttributeMaxDynamicSharedMemorySize, kIdctSmemFloats * sizeof(float));
  cudaFuncSetAttribute(k_idct_mid, cudaFuncAttributeMaxDynamicSharedMemorySize, kMidSmemFloats * sizeof(float));
  cudaFuncSetAttribute(k_render_fused, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       6 * DevRenderTileFloats(kRtMaxHalo) * sizeof(float));
  cudaFuncSetAttribute(k_modular_decode_coop<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_modular_decode_coop<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_modular_decode_sparse<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(k_modular_decode_sparse<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(k_ac_decode_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  // (Measured and not kept, profiles/r2_ab_carveout.txt: cudaFuncAttributePreferredSharedMemoryCarveout = max shared on
  // every kernel, on the theory that a long-running entropy CTA pins its SM's L1 / shared-memory split and keeps render
  // CTAs out. The slowdown of the per-pixel kernels next to AC decode stayed 4 x, and the entropy kernels lost 8 - 20 %
  // to the smaller L1: the interference is issue slots, not the carve-out.)
  return dec;
}

void JxlB200DecoderDestroy(JxlB200Decoder* dec) {
  if (!dec) return;
  cudaSetDevice(dec->device);
  FoldEvents(dec);
  for (cudaEvent_t e : dec->free_events) cudaEventDestroy(e);
  if (dec->pix_done) {
    PixelTurn& turn = g_pixel_turn[dec->device];
    std::lock_guard<std::mutex> lock(turn.mu);
    if (turn.owner == dec) {
      turn.owner = nullptr;
      turn.last = nullptr;
    }
    cudaEventDestroy(dec->pix_done);
  }
  for (uint8_t* hb : dec->h_bytes)
    if (hb) cudaFreeHost(hb);
  if (dec->copy_stream) {
    cudaStreamSynchronize(dec->copy_stream);
    cudaStreamDestroy(dec->copy_stream);
  }
  for (cudaEvent_t ev : dec->copy_ev) cudaEventDestroy(ev);
  if (dec->copy_done) cudaEventDestroy(dec->copy_done);
  if (dec->sync_ev) cudaEventDestroy(dec->sync_ev);
  if (dec->stream) cudaStreamDestroy(dec->stream);
  for (cudaStream_t st : dec->part_stream)
    if (st) cudaStreamDestroy(st);
  for (cudaEvent_t ev : dec->part_ev)
    if (ev) cudaEventDestroy(ev);
  delete dec;
}

const char* JxlB200DecoderGetError(const JxlB200Decoder* dec) { return dec ? dec->error.c_str() : "null decoder"; }

static int UploadTokensLayout(JxlB200Decoder* dec) {
  BatchPlan& b = *dec->plan;
  CUDA_OK(dec->d_ac_streams.Upload(b.ac_streams, dec->stream));
  CUDA_OK(dec->d_tokens.Alloc(b.tok_size + 16));
  dec->vpools.streams = dec->d_ac_streams.p;
  dec->vpools.tokens = dec->d_tokens.p;
  return 0;
}

static int LaunchModular(JxlB200Decoder* dec, const BatchPlan& b, cudaStream_t s, int phase, uint32_t* launches);

// Waits for stream `s` on an event that sleeps (cudaEventBlockingSync) instead of cudaStreamSynchronize's spinning: a
// server keeps several handles in flight from as many threads, and eight threads spinning for a second each take half
// of a 16-core host away from the planning threads of the next batches.
static cudaError_t BlockingSync(JxlB200Decoder* dec, cudaStream_t s) {
  static const bool spin = std::getenv("JXLB200_SPIN_WAIT") != nullptr;
  if (spin) return cudaStreamSynchronize(s);
  if (!dec->sync_ev) {
    cudaError_t e = cudaEventCreateWithFlags(&dec->sync_ev, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaEventRecord(dec->sync_ev, s);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(dec->sync_ev);
}

// Uploads the pools of `b` and allocates the arenas; `dec->pools` / `dec->vpools` describe them afterwards.
static int UploadPlan(JxlB200Decoder* dec, const BatchPlan& b, const PixelFormat& fmt, bool want_end_bits) {
  CUDA_OK(cudaSetDevice(dec->device));
  cudaStream_t s = dec->stream;
  static const bool use_stage = std::getenv("JXLB200_NO_PINNED_STAGE") == nullptr;
  PinnedStage* stage = use_stage ? &dec->stage : nullptr;
  if (stage) stage->Reset();  // (the stream was synchronised at the end of the last upload)
  if (b.ext_bytes) {  // pinned staging filled by the planning threads: one asynchronous copy
    CUDA_OK(dec->d_bytes.Alloc(b.ext_bytes_size));
    CUDA_OK(cudaMemcpyAsync(dec->d_bytes.p, b.ext_bytes, b.ext_bytes_size, cudaMemcpyHostToDevice, s));
  } else {
    CUDA_OK(dec->d_bytes.Upload(b.bytes, s, stage));
  }
  static const bool up_timing = std::getenv("JXLB200_RUN_TIMING") != nullptr;
  const auto ut0 = std::chrono::steady_clock::now();
  auto ut_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ut0).count(); };
  CUDA_OK(dec->d_alias.Upload(b.alias, s, stage));
  CUDA_OK(dec->d_prefix.Upload(b.prefix, s, stage));
  CUDA_OK(dec->d_cfg.Upload(b.cfg, s, stage));
  CUDA_OK(dec->d_refs.Upload(b.refs, s, stage));
  CUDA_OK(dec->d_lut.Upload(b.lut, s, stage));
  CUDA_OK(dec->d_tree.Upload(b.tree, s, stage));
  CUDA_OK(dec->d_codes.Upload(b.codes, s, stage));
  CUDA_OK(dec->d_chans.Upload(b.chans, s, stage));
  CUDA_OK(dec->d_streams.Upload(b.streams, s, stage));
  CUDA_OK(dec->d_planes.Upload(b.planes, s, stage));
  CUDA_OK(dec->d_ops.Upload(b.ops, s, stage));
  CUDA_OK(dec->d_group_programs.Upload(b.group_programs, s, stage));
  std::vector<DevProgram> all_levels;
  dec->level_off.clear();
  for (const auto& lvl : b.levels) {
    dec->level_off.push_back(all_levels.size());
    all_levels.insert(all_levels.end(), lvl.begin(), lvl.end());
  }
  CUDA_OK(dec->d_levels.Upload(all_levels, s, stage));
  CUDA_OK(dec->d_late_group_programs.Upload(b.late_group_programs, s, stage));
  std::vector<DevProgram> all_late_levels;
  dec->late_level_off.clear();
  for (const auto& lvl : b.late_levels) {
    dec->late_level_off.push_back(all_late_levels.size());
    all_late_levels.insert(all_late_levels.end(), lvl.begin(), lvl.end());
  }
  CUDA_OK(dec->d_late_levels.Upload(all_late_levels, s, stage));
  CUDA_OK(dec->d_chain_pos.Alloc(b.chain_slots + 1));
  CUDA_OK(dec->d_spl_seg.Upload(b.spl_seg, s, stage));
  CUDA_OK(dec->d_spl_idx.Upload(b.spl_idx, s, stage));
  CUDA_OK(dec->d_frames.Upload(b.frames, s, stage));
  CUDA_OK(dec->d_warp_chans.Upload(b.warp_chans, s, stage));
  CUDA_OK(dec->d_warp_dims_off.Upload(b.warp_dims_off, s, stage));
  CUDA_OK(dec->d_warp_dims.Upload(b.warp_dims, s, stage));
  CUDA_OK(dec->d_arena.Alloc(b.arena_size + 16));
  const size_t num_warps = b.warp_chans.size() + 1;  // (lock-step bundles of both launches)
  CUDA_OK(dec->d_wp.Alloc(num_warps * 10 * (b.wp_width + 2) * 32 + 16));
  CUDA_OK(dec->d_ring.Alloc(num_warps * 3 * b.wp_width * 32 + 16));
  CUDA_OK(dec->d_lz77.Alloc(static_cast<size_t>(b.lz77_slots) << 20));
  CUDA_OK(dec->d_status.Alloc(b.streams.size()));
  if (want_end_bits) CUDA_OK(dec->d_end_bits.Alloc(b.streams.size()));
  CUDA_OK(dec->d_out.Alloc(b.out_size));
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
  P.lut = dec->d_lut.p;
  P.codes = dec->d_codes.p;
  P.arena = dec->d_arena.p;
  P.wp_scratch = dec->d_wp.p;
  P.ring = dec->d_ring.p;
  P.wp_width = b.wp_width;
  P.lz77 = dec->d_lz77.p;
  P.status = dec->d_status.p;
  P.end_bits = want_end_bits ? dec->d_end_bits.p : nullptr;
  P.num_streams = b.streams.size();
  P.stream0 = b.num_coop;
  P.coop0 = 0;
  P.chain_pos = dec->d_chain_pos.p;
  P.spl_seg = dec->d_spl_seg.p;
  P.spl_idx = dec->d_spl_idx.p;
  P.warp_chans = dec->d_warp_chans.p;
  P.warp_dims_off = dec->d_warp_dims_off.p;
  P.warp_dims = dec->d_warp_dims.p;
  dec->uniform_rgba8 = fmt.num_channels == 4 && fmt.data_type == 2;
  dec->any_modular_frame = false;
  for (const DevFrameOut& fo : b.frames) {
    if (fo.vardct) continue;
    dec->any_modular_frame = true;
    for (int c = 0; c < 4; c++)
      if (fo.is_float[c] || fo.stride % 4 || fo.orient != 0 || fo.has_splines) dec->uniform_rgba8 = false;
  }
  // ---- VarDCT
  dec->dcg_list.clear();
  dec->max_groups = dec->max_xsize = dec->max_ysize = dec->max_blocks = dec->max_epf = 0;
  dec->max_up_xsize = dec->max_up_ysize = 0;
  dec->any_gab = false;
  if (!b.vframes.empty()) {
    const SharedVarDCTTables& sh = SharedVarDCTTables::Get();
    for (uint32_t f = 0; f < b.vframes.size(); f++) {
      const DevVFrame& vf = b.vframes[f];
      for (uint32_t g = 0; g < vf.xdcgroups * vf.ydcgroups; g++) dec->dcg_list.push_back(make_uint2(f, g));
      dec->max_groups = std::max(dec->max_groups, vf.xgroups * vf.ygroups);
      dec->max_xsize = std::max(dec->max_xsize, vf.xsize);
      dec->max_ysize = std::max(dec->max_ysize, vf.ysize);
      dec->max_up_xsize = std::max(dec->max_up_xsize, vf.upsampling > 1 ? vf.xsize * vf.upsampling : vf.xsize);
      dec->max_up_ysize = std::max(dec->max_up_ysize, vf.upsampling > 1 ? vf.ysize * vf.upsampling : vf.ysize);
      dec->max_blocks = std::max(dec->max_blocks, vf.xblocks * vf.yblocks);
      dec->max_epf = std::max(dec->max_epf, vf.epf_iters);
      dec->any_gab |= vf.gab != 0;
    }
    CUDA_OK(dec->d_vframes.Upload(b.vframes, s, stage));
    CUDA_OK(dec->d_fpool.Upload(b.fpool, s, stage));
    CUDA_OK(dec->d_opool.Upload(b.opool, s, stage));
    CUDA_OK(dec->d_cpool.Upload(b.cpool, s, stage));
    CUDA_OK(dec->d_upool.Upload(b.upool, s, stage));
    CUDA_OK(dec->d_dcg_list.Upload(dec->dcg_list, s, stage));
    CUDA_OK(dec->d_patches.Upload(b.patches, s, stage));
    CUDA_OK(dec->d_ref_frames.Upload(b.ref_frames, s, stage));
    CUDA_OK(dec->d_farena.Alloc(b.farena_size + 16));
    CUDA_OK(dec->d_barena.Alloc(b.barena_size + 16));
    CUDA_OK(dec->d_uarena.Alloc(b.uarena_size + 16));
    CUDA_OK(dec->d_ac_status.Alloc(b.ac_streams.size() + 1));
    CUDA_OK(dec->d_ac_used.Alloc(b.ac_streams.size() + 1));
    CUDA_OK(dec->d_dc_status.Alloc(dec->dcg_list.size() + 2));
    CUDA_OK(dec->d_big_scratch.Alloc(static_cast<size_t>(JxlB200Decoder::kBigCtas) * 4 * 65536));
    // (a 32x32-block group holds at most 32 varblocks of 64x32 pixels)
    CUDA_OK(dec->d_mid_list.Alloc(static_cast<size_t>(std::max<uint32_t>(1, b.wave_frames)) * dec->max_groups * 32 + 1));
    DevVPools& V = dec->vpools;
    V = DevVPools{};
    V.frames = dec->d_vframes.p;
    V.num_streams = b.ac_streams.size();
    V.fpool = dec->d_fpool.p;
    V.opool = dec->d_opool.p;
    V.cpool = dec->d_cpool.p;
    V.upool = dec->d_upool.p;
    V.farena = dec->d_farena.p;
    V.barena = dec->d_barena.p;
    V.uarena = dec->d_uarena.p;
    V.ac_status = dec->d_ac_status.p;
    V.ac_used = dec->d_ac_used.p;
    V.dc_status = dec->d_dc_status.p;
    V.wc_off = sh.wc_off;
    V.llf_off = sh.llf_off;
    V.afv_off = sh.afv_off;
    V.sinfo_off = sh.sinfo_off;
    V.ctxtab_off = sh.ctxtab_off;
    V.out = dec->d_out.p;
    V.ref_frames = dec->d_ref_frames.p;
    V.num_ref_frames = b.ref_frames.size();
    V.patches = dec->d_patches.p;
    V.chain_pos = dec->d_chain_pos.p;
    V.arena = dec->d_arena.p;
    V.planes = dec->d_planes.p;
    V.spl_seg = dec->d_spl_seg.p;
    V.spl_idx = dec->d_spl_idx.p;
    V.ac_plain_ans = 1;
    for (const DevVFrame& vf : b.vframes)
      for (uint32_t p = 0; p < vf.num_passes; p++)
        if (b.codes[vf.ac_code[p]].use_prefix || b.codes[vf.ac_code[p]].lz77_enabled) V.ac_plain_ans = 0;
    CUDA_OK(dec->d_ac_streams.Upload(b.ac_streams, s, stage));
    CUDA_OK(dec->d_tokens.Alloc(b.tok_size + 16));
    V.streams = dec->d_ac_streams.p;
    V.tokens = dec->d_tokens.p;
    // The CTA-per-(frame, pass) AC kernel (tables TMA-staged in shared memory) is opt-in, JXLB200_AC_FRAME=1; it needs
    // plain ANS codes and tables that fit the shared memory of an SM next to a second CTA. Measured on the bench batch
    // (profiles/r2_ab_ac_frame_kernel.txt): 74.6 ms alone against 52.9 ms for the one-warp kernel, whose warps hold 32
    // streams of equal length from 32 frames and therefore run in lock step, while a frame's own 135 streams differ in
    // length and diverge; its 93 KB of shared memory per CTA also starve the render CTAs of the other handles.
    dec->ac_frame_smem = dec->ac_frame_threads = 0;
    if (V.ac_plain_ans && !b.ac_units.empty() && std::getenv("JXLB200_AC_FRAME")) {
      uint32_t max_count = 0, max_smem = 0;
      for (const DevAcUnit& u : b.ac_units) max_count = std::max(max_count, u.count);
      const uint32_t warps = std::min<uint32_t>(8, (max_count + 31) / 32);
      for (const DevAcUnit& u : b.ac_units) {
        const DevVFrame& vf = b.vframes[u.frame];
        const DevCode& c = b.codes[vf.ac_code[u.pass]];
        max_smem = std::max(max_smem, AcFrameLayout(c.num_clusters << c.log_alpha_size, c.num_clusters,
                                                    vf.num_histograms * vf.num_ctxs * 495, warps).total);
      }
      if (max_smem <= 110 * 1024) {
        dec->ac_frame_smem = max_smem;
        dec->ac_frame_threads = warps * 32;
        CUDA_OK(dec->d_ac_units.Upload(b.ac_units, s, stage));
        CUDA_OK(dec->d_ac_unit_streams.Upload(b.ac_unit_streams, s, stage));
      }
    }
  }
  const double ut_enq = ut_ms();
  CUDA_OK(BlockingSync(dec, s));
  if (up_timing)
    std::fprintf(stderr, "UploadPlan: enqueue %.2f ms, + sync %.2f ms (bitstreams %.1f MB pinned, alias %.1f MB, orders %.1f MB, fpool %.1f MB, "
                 "ac streams %.1f MB)\n", ut_enq, ut_ms(), b.ext_bytes_size / 1e6, b.alias.size() * sizeof(DevAlias) / 1e6,
                 b.opool.size() * 2 / 1e6, b.fpool.size() * 4 / 1e6, b.ac_streams.size() * sizeof(DevAcStream) / 1e6);
  return 0;
}

int JxlB200DecoderPlanBatch(JxlB200Decoder* dec, const uint8_t* const* files, const size_t* sizes, size_t n,
                            const JxlPixelFormat* format, int num_threads) {
  if (!dec || !files || !sizes || !format || n == 0) return 1;
  dec->error.clear();
  dec->next_plan.reset();
  PixelFormat fmt;
  fmt.num_channels = format->num_channels;
  fmt.data_type = format->data_type;
  fmt.endianness = format->endianness;
  fmt.align = format->align;
  fmt.keep_orientation = dec->keep_orientation;
  if (fmt.num_channels < 1 || fmt.num_channels > 4 ||
      !(fmt.data_type == 0 || fmt.data_type == 2 || fmt.data_type == 3 || fmt.data_type == 5)) {
    dec->error = "invalid pixel format";
    return 1;
  }
  // Probe rounds (single-section frames, raw quantisation tables): the Modular decode kernel runs over the probe
  // streams alone and the host reads back where each one ended and the samples it asked for.
  const ProbeFn probe = [dec](const BatchPlan& pb, std::vector<uint64_t>* end_bits, std::vector<int32_t>* arena) {
    JxlB200Decoder tmp;
    tmp.device = dec->device;
    tmp.stream = dec->stream;  // borrowed
    struct EventGuard {
      JxlB200Decoder* d;
      ~EventGuard() {
        if (d->sync_ev) cudaEventDestroy(d->sync_ev);
      }
    } guard{&tmp};
    PixelFormat pf;
    std::vector<uint32_t> status(pb.streams.size());
    end_bits->assign(pb.streams.size(), 0);
    arena->assign(pb.arena_size, 0);
    auto ok = [&](cudaError_t e) {
      if (e != cudaSuccess) throw Error(std::string("probe launch: ") + cudaGetErrorString(e));
    };
    if (UploadPlan(&tmp, pb, pf, true) != 0) throw Error("probe launch: " + tmp.error);
    ok(cudaMemsetAsync(tmp.d_arena.p, 0, pb.arena_size * sizeof(int32_t), tmp.stream));
    uint32_t probe_launches = 0;
    if (LaunchModular(&tmp, pb, tmp.stream, 0, &probe_launches) != 0) throw Error("probe launch: " + tmp.error);
    ok(cudaGetLastError());
    ok(cudaMemcpyAsync(status.data(), tmp.d_status.p, status.size() * 4, cudaMemcpyDeviceToHost, tmp.stream));
    ok(cudaMemcpyAsync(end_bits->data(), tmp.d_end_bits.p, end_bits->size() * 8, cudaMemcpyDeviceToHost, tmp.stream));
    if (pb.arena_size)
      ok(cudaMemcpyAsync(arena->data(), tmp.d_arena.p, pb.arena_size * sizeof(int32_t), cudaMemcpyDeviceToHost, tmp.stream));
    ok(cudaStreamSynchronize(tmp.stream));
    dec->probe_launches++;
    for (size_t i = 0; i < status.size(); i++)
      if (status[i] != 0) throw Error("chained sub-stream " + std::to_string(i) + " failed (status " + std::to_string(status[i]) + ")");
  };
  std::unique_ptr<BatchPlan> plan(new BatchPlan());
  try {
    if (cudaSetDevice(dec->device) != cudaSuccess) throw Error("cudaSetDevice failed");
    const int turn = dec->h_turn ^= 1;
    const BytesAlloc bytes_alloc = [dec, turn](size_t size) -> uint8_t* {
      if (size > dec->h_bytes_cap[turn]) {
        if (dec->h_bytes[turn]) cudaFreeHost(dec->h_bytes[turn]);
        dec->h_bytes[turn] = nullptr;
        dec->h_bytes_cap[turn] = 0;
        const size_t cap = size + size / 4;
        if (cudaHostAlloc(reinterpret_cast<void**>(&dec->h_bytes[turn]), cap, cudaHostAllocDefault) != cudaSuccess) return nullptr;
        dec->h_bytes_cap[turn] = cap;
      }
      return dec->h_bytes[turn];
    };
    PlanBatch(files, sizes, n, fmt, num_threads, plan.get(), probe, bytes_alloc);
  } catch (const std::exception& e) {
    dec->error = e.what();
    return 1;
  }
  dec->next_plan = std::move(plan);
  dec->next_fmt = fmt;
  return 0;
}

static void FinishCopies(JxlB200Decoder* dec) {
  if (dec->copy_pending) BlockingSync(dec, dec->copy_stream);
  dec->copy_pending = false;
  dec->host_dsts.clear();
}

int JxlB200DecoderCommitPlan(JxlB200Decoder* dec) {
  if (!dec) return 1;
  if (!dec->next_plan) {
    dec->error = "no planned batch to commit";
    return 1;
  }
  const auto t0 = std::chrono::steady_clock::now();
  auto ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  CUDA_OK(cudaSetDevice(dec->device));
  FinishCopies(dec);
  // The upload reallocates / overwrites the device pools of the previous batch: the handle has no plan until it
  // succeeded (a failed upload must not leave Run / ReadOutput with the old grid sizes over freed pools).
  std::unique_ptr<BatchPlan> plan = std::move(dec->next_plan);
  dec->plan.reset();
  const double t_free = ms();
  dec->pools = DevPools{};
  dec->vpools = DevVPools{};
  if (UploadPlan(dec, *plan, dec->next_fmt, false) != 0) return 1;
  dec->plan = std::move(plan);
  if (std::getenv("JXLB200_RUN_TIMING")) std::fprintf(stderr, "CommitPlan: old plan freed after %.2f ms, uploaded after %.2f ms\n", t_free, ms());
  return 0;
}

int JxlB200DecoderSetInputBatch(JxlB200Decoder* dec, const uint8_t* const* files, const size_t* sizes, size_t n,
                                const JxlPixelFormat* format, int num_threads) {
  if (JxlB200DecoderPlanBatch(dec, files, sizes, n, format, num_threads) != 0) return 1;
  return JxlB200DecoderCommitPlan(dec);
}

int JxlB200DecoderSetKeepOrientation(JxlB200Decoder* dec, int keep) {
  if (!dec) return 1;
  dec->keep_orientation = keep != 0;
  return 0;
}

size_t JxlB200DecoderNumFrames(const JxlB200Decoder* dec) { return dec && dec->plan ? dec->plan->frames.size() : 0; }

// keep_orientation false (libjxl's default): the sizes are those of the upright image and the orientation reads
// identity (lib/jxl/decode.cc:2083-2090).
static void FillBasicInfo(const BasicInfo& bi, bool have_container, JxlBasicInfo* info, bool keep_orientation = true) {
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
  if (!keep_orientation) {
    if (m.orientation >= 5) std::swap(info->xsize, info->ysize);
    info->orientation = 1;
  }
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
  FillBasicInfo(dec->plan->info[i], false, info, dec->keep_orientation);
  return 0;
}

size_t JxlB200DecoderImageOutBufferSize(const JxlB200Decoder* dec, size_t i) {
  if (!dec || !dec->plan || i >= dec->plan->frames.size()) return 0;
  return dec->plan->frame_out_size[i];
}

// phase 0: the streams whose position the host knows; phase 1 (after the AC decode kernel): the ones chained behind AC
// coefficient streams (BatchPlan::num_early). Each phase: its one-per-warp streams, then its lock-step bundles.
static int LaunchModular(JxlB200Decoder* dec, const BatchPlan& b, cudaStream_t s, int phase, uint32_t* launches) {
  DevPools P = dec->pools;
  const uint32_t block = 32;
  const uint32_t first = phase == 0 ? 0 : b.num_early, last = phase == 0 ? b.num_early : static_cast<uint32_t>(b.streams.size());
  const uint32_t ncoop = phase == 0 ? b.num_coop : b.late_coop;
  if (first == last) return 0;
  P.coop0 = first;
  P.stream0 = first + ncoop;
  P.num_streams = last;
  if (phase == 1) {  // the bundles of the late streams follow the early ones in the warp tables
    P.warp_chans += b.early_warps;
    P.warp_dims_off += b.early_warps;
  }
  if (ncoop != 0) {  // one warp per stream: the chains under libjxl's fixed trees
    const size_t coop_smem = static_cast<size_t>(7 * b.wp_width + 10) * sizeof(int32_t) * kCoopWarps;
    if (coop_smem > 200 * 1024) return 1;  // (channel widths are bounded by the group size: never)
    const uint32_t coop_grid = (ncoop + kCoopWarps - 1) / kCoopWarps;
    if (b.narrow) {
      k_modular_decode_coop<int32_t><<<coop_grid, 32 * kCoopWarps, coop_smem, s>>>(P);
    } else {
      k_modular_decode_coop<int64_t><<<coop_grid, 32 * kCoopWarps, coop_smem, s>>>(P);
    }
    CUDA_OK(cudaGetLastError());
    (*launches)++;
  }
  const size_t rest = last - first - ncoop;
  if (rest == 0) return 0;
  const size_t sparse_smem = static_cast<size_t>(7 * b.wp_width + 10) * kSparseLanes * sizeof(int32_t);
  // (JXLB200_MODULAR_DENSE=1: always the dense-lane kernel -- 32 streams per warp, rows in HBM, 6 KB of shared memory
  // per CTA instead of 58 KB: slower alone, but it leaves the SMs' shared memory to the kernels of other batches)
  static const bool force_dense = std::getenv("JXLB200_MODULAR_DENSE") != nullptr;
  if (!force_dense && rest <= 2 * 148 * kSparseLanes && sparse_smem <= 160 * 1024) {
    const uint32_t grid = (rest + kSparseLanes - 1) / kSparseLanes;
    if (b.narrow) {
      k_modular_decode_sparse<int32_t><<<grid, block, sparse_smem, s>>>(P);
    } else {
      k_modular_decode_sparse<int64_t><<<grid, block, sparse_smem, s>>>(P);
    }
  } else if (b.narrow) {
    k_modular_decode<int32_t><<<(rest + block - 1) / block, block, 0, s>>>(P);
  } else {
    k_modular_decode<int64_t><<<(rest + block - 1) / block, block, 0, s>>>(P);
  }
  CUDA_OK(cudaGetLastError());
  (*launches)++;
  return 0;
}

int JxlB200DecoderRun(JxlB200Decoder* dec, void* cuda_stream) {
  if (!dec || !dec->plan) return 1;
  CUDA_OK(cudaSetDevice(dec->device));
  cudaStream_t caller = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : dec->stream;
  // Without an SM partition everything runs on the caller's stream. With one: entropy kernels on the handle's stream in
  // the entropy partition (`s`), per-pixel kernels on its stream in the pixel partition (`sp`), ordered by events.
  const bool parted = dec->part_stream[0] != nullptr && !dec->plan->vframes.empty();
  cudaStream_t s = parted ? dec->part_stream[0] : caller;
  cudaStream_t sp = parted ? dec->part_stream[1] : caller;
  const bool to_host = !dec->host_dsts.empty();
  std::vector<uint8_t> copied;
  uint32_t waves_done = 0;
  if (dec->copy_pending) CUDA_OK(cudaStreamWaitEvent(caller, dec->copy_done, 0));  // (a read-back of the old output)
  if (to_host) {
    if (!dec->copy_stream) {
      CUDA_OK(cudaStreamCreateWithFlags(&dec->copy_stream, cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&dec->copy_done, cudaEventDisableTiming));
    }
    copied.assign(dec->plan->frames.size(), 0);
  }
  // frames [..] of the list -> the caller's buffers, on the copy stream, once `after` has reached this point
  auto copy_out = [&](cudaStream_t after, const uint32_t* frames, size_t count) -> cudaError_t {
    if (count == 0) return cudaSuccess;
    if (waves_done >= dec->copy_ev.size()) {
      cudaEvent_t ev = nullptr;
      cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
      dec->copy_ev.push_back(ev);
    }
    cudaEvent_t ev = dec->copy_ev[waves_done++];
    cudaError_t e = cudaEventRecord(ev, after);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(dec->copy_stream, ev, 0);
    // frames that follow each other in the caller's memory as they do in the output buffer (one pinned arena laid out
    // like JxlB200DecoderDeviceOutput) leave in one copy
    static const bool coalesce = std::getenv("JXLB200_NO_COALESCE") == nullptr;
    for (size_t k = 0; k < count && e == cudaSuccess;) {
      const uint32_t f0 = frames[k];
      const uint64_t off0 = dec->plan->frames[f0].out_off;
      uint8_t* const dst0 = static_cast<uint8_t*>(dec->host_dsts[f0]);
      uint64_t len = dec->plan->frame_out_size[f0];
      copied[f0] = 1;
      size_t j = k + 1;
      for (; coalesce && j < count; j++) {
        const uint32_t fj = frames[j];
        const uint64_t offj = dec->plan->frames[fj].out_off;
        if (offj < off0 + len || static_cast<uint8_t*>(dec->host_dsts[fj]) != dst0 + (offj - off0)) break;
        len = offj - off0 + dec->plan->frame_out_size[fj];
        copied[fj] = 1;
      }
      e = cudaMemcpyAsync(dst0, dec->d_out.p + off0, len, cudaMemcpyDeviceToHost, dec->copy_stream);
      k = j;
    }
    return e;
  };
  static const bool run_timing = std::getenv("JXLB200_RUN_TIMING") != nullptr;  // stderr: host time of the enqueue
  const auto rt0 = std::chrono::steady_clock::now();
  auto rt_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - rt0).count(); };
  double rt_entropy = 0, rt_waves = 0;
  if (parted) {
    CUDA_OK(cudaEventRecord(dec->part_ev[0], caller));
    CUDA_OK(cudaStreamWaitEvent(s, dec->part_ev[0], 0));
    CUDA_OK(cudaStreamWaitEvent(sp, dec->part_ev[0], 0));  // (the previous run's read-back of the output is over)
  }
  const BatchPlan& b = *dec->plan;
  const DevPools& P = dec->pools;
  uint32_t launches = 0;
  const uint32_t pm = dec->phase_mask;
  if (!b.streams.empty() && (pm & (1u << kKModular))) {
    ScopedTimer t(dec, s, kKModular);
    if (LaunchModular(dec, b, s, 0, &launches) != 0) return 1;
  }
  if (!b.group_programs.empty()) {
    ScopedTimer t(dec, s, kKGroupPrograms);
    k_group_programs<<<b.group_programs.size(), 256, 0, s>>>(P, dec->d_ops.p, dec->d_group_programs.p);
    launches++;
  }
  {
    ScopedTimer t(dec, s, kKFrameLevels);
    for (size_t k = 0; k < b.levels.size(); k++) {
      const uint32_t tiles = std::max<uint32_t>(1, std::min<uint32_t>(1024, (b.max_frame_pixels + 1023) / 1024));
      dim3 grid(tiles, b.levels[k].size());
      k_frame_level<<<grid, 256, 0, s>>>(P, dec->d_ops.p, dec->d_levels.p + dec->level_off[k]);
      launches++;
    }
  }
  if (dec->any_modular_frame) {
    ScopedTimer t(dec, s, kKWriteOutput);
    const uint32_t tiles = std::max<uint32_t>(1, std::min<uint32_t>(4096, (b.max_frame_pixels + 255) / 256));
    dim3 grid(tiles, b.frames.size());
    if (dec->uniform_rgba8) {
      k_write_output_rgba8<<<grid, 256, 0, s>>>(P, dec->d_frames.p, dec->d_out.p);
    } else {
      k_write_output<<<grid, 256, 0, s>>>(P, dec->d_frames.p, dec->d_out.p);
    }
    launches++;
  }
  if (!b.vframes.empty()) {
    const DevVPools& V = dec->vpools;
    const uint32_t nvf = b.vframes.size();
    if (!b.ref_frames.empty()) {
      ScopedTimer t(dec, s, kKFrameLevels);
      k_ref_frames<<<dim3(16, b.ref_frames.size()), 256, 0, s>>>(P, V);
      launches++;
    }
    if (pm & (1u << kKDcFinish)) {
      ScopedTimer t(dec, s, kKDcFinish);
      CUDA_OK(cudaMemsetAsync(dec->d_dc_status.p, 0, (dec->dcg_list.size() + 2) * 4, s));
      // (64 threads: the serial scan of a DC group is one thread's work; a wider CTA only parks warps on the SMs that
      // the per-pixel kernels of other batches need -- tools/interference.py)
      static const uint32_t dcf_threads = std::getenv("JXLB200_DCF_THREADS") ? std::atoi(std::getenv("JXLB200_DCF_THREADS")) : 64;
      k_dc_finish<<<dec->dcg_list.size(), dcf_threads, 0, s>>>(P, V, dec->d_dcg_list.p);
      dim3 grid(std::max<uint32_t>(1, std::min<uint32_t>(256, (dec->max_blocks + 255) / 256)), nvf);
      k_dc_smooth<<<grid, 256, 0, s>>>(V);
      launches += 2;
    }
    if (pm & (1u << kKAcDecode)) {
      ScopedTimer t(dec, s, kKAcDecode);
      k_block_lists<<<(dec->max_groups * nvf + 127) / 128, 128, 0, s>>>(V, dec->max_groups, nvf);
      launches++;
      static const uint32_t ac_warps = std::getenv("JXLB200_AC_WARPS") ? std::atoi(std::getenv("JXLB200_AC_WARPS")) : 1;
      const uint32_t nst = b.ac_streams.size();
      if (dec->ac_frame_smem != 0) {
        k_ac_decode_frame<<<b.ac_units.size(), dec->ac_frame_threads, dec->ac_frame_smem, s>>>(P, V, dec->d_ac_units.p,
                                                                                                dec->d_ac_unit_streams.p);
      } else if (ac_warps >= 8) {
        k_ac_decode<8><<<(nst + 255) / 256, 256, 0, s>>>(P, V);
      } else if (ac_warps >= 4) {
        k_ac_decode<4><<<(nst + 127) / 128, 128, 0, s>>>(P, V);
      } else {
        k_ac_decode<1><<<(nst + 31) / 32, 32, 0, s>>>(P, V);
      }
      launches++;
    }
    if (b.num_early != b.streams.size() || !b.late_group_programs.empty() || !b.late_levels.empty()) {
      // the extra channels of VarDCT frames: the streams chained behind the AC coefficients, then their copies into the
      // frame's planes and the global inverse transforms
      {
        ScopedTimer t(dec, s, kKModular);
        if (LaunchModular(dec, b, s, 1, &launches) != 0) return 1;
      }
      if (!b.late_group_programs.empty()) {
        ScopedTimer t(dec, s, kKGroupPrograms);
        k_group_programs<<<b.late_group_programs.size(), 256, 0, s>>>(P, dec->d_ops.p, dec->d_late_group_programs.p);
        launches++;
      }
      ScopedTimer t(dec, s, kKFrameLevels);
      for (size_t k = 0; k < b.late_levels.size(); k++) {
        const uint32_t tiles = std::max<uint32_t>(1, std::min<uint32_t>(1024, (b.max_frame_pixels + 1023) / 1024));
        dim3 grid(tiles, b.late_levels[k].size());
        k_frame_level<<<grid, 256, 0, s>>>(P, dec->d_ops.p, dec->d_late_levels.p + dec->late_level_off[k]);
        launches++;
      }
    }
    rt_entropy = rt_ms();
    if (parted) {  // the per-pixel waves start when this batch's entropy kernels are through
      CUDA_OK(cudaEventRecord(dec->part_ev[1], s));
      CUDA_OK(cudaStreamWaitEvent(sp, dec->part_ev[1], 0));
    }
    const dim3 px_block(32, 8);
    const bool turns = dec->pix_done != nullptr && PixelTurnsOn();
    std::unique_lock<std::mutex> turn_lock;
    if (turns) {  // (held until this phase is enqueued: the order of the turns is the order of the enqueues)
      PixelTurn& turn = g_pixel_turn[dec->device];
      turn_lock = std::unique_lock<std::mutex>(turn.mu);
      if (turn.last != nullptr && turn.owner != dec) CUDA_OK(cudaStreamWaitEvent(sp, turn.last, 0));
    }
    for (uint32_t f0 = 0; f0 < nvf; f0 += b.wave_frames) {
      const uint32_t nf = std::min<uint32_t>(b.wave_frames, nvf - f0);
      if (pm & (1u << kKDequantIdct)) {
        ScopedTimer t(dec, sp, kKDequantIdct);
        uint32_t* has_mid = dec->d_dc_status.p + dec->dcg_list.size();  // spare words after the DC status words
        uint32_t* mid_count = has_mid + 1;
        if (f0 != 0 || !(pm & (1u << kKDcFinish))) CUDA_OK(cudaMemsetAsync(mid_count, 0, 4, sp));  // (the first wave's was cleared with the DC status)
        k_dequant_idct<<<dim3(dec->max_groups, nf), kIdctThreads, kIdctSmemFloats * sizeof(float), sp>>>(V, f0, has_mid, mid_count,
                                                                                                       dec->d_mid_list.p);
        k_idct_mid<<<4 * 148, kMidThreads, kMidSmemFloats * sizeof(float), sp>>>(V, mid_count, dec->d_mid_list.p);
        k_idct_big<<<JxlB200Decoder::kBigCtas, 256, 0, sp>>>(V, f0, nf, dec->d_big_scratch.p, has_mid);
        launches += 3;
      }
      const dim3 px_grid((dec->max_xsize + 31) / 32, (dec->max_ysize + 7) / 8, nf);
      // Frames without patches: one fused kernel (tiles in shared memory). The per-pixel kernels follow only when
      // the batch has patches (or the fused path is switched off: JXLB200_UNFUSED_RENDER=1, for comparison).
      const uint32_t skip_fused = dec->fused_render ? 1 : 0;
      if (dec->fused_render && (pm & (1u << kKFilters))) {
        ScopedTimer t(dec, sp, kKFilters);
        const uint32_t halo = DevRenderHalo(dec->any_gab ? 1 : 0, dec->max_epf);
        const uint32_t cap = DevRenderTileFloats(halo);
        const dim3 rt_grid((dec->max_xsize + kRtW - 1) / kRtW, (dec->max_ysize + kRtH - 1) / kRtH, nf);
        k_render_fused<<<rt_grid, 256, 6 * cap * sizeof(float), sp>>>(V, f0, cap, DevRenderStride(halo));
        launches++;
      }
      if (!dec->fused_render || !b.patches.empty() || b.any_upsampling) {
        {
          ScopedTimer t(dec, sp, kKFilters);
          if (dec->any_gab) {
            k_gaborish<<<px_grid, px_block, 0, sp>>>(V, f0, 0, 1, skip_fused);
            launches++;
          }
          for (uint32_t stage = 0; stage < 3; stage++) {
            if (dec->max_epf == 0 || (stage == 0 && dec->max_epf < 3) || (stage == 2 && dec->max_epf < 2)) continue;
            k_epf<<<px_grid, px_block, 0, sp>>>(V, f0, stage, skip_fused);
            launches++;
          }
        }
        if (!b.patches.empty()) {
          ScopedTimer t(dec, sp, kKFilters);
          k_patches<<<nf, 256, 0, sp>>>(V, f0);
          launches++;
        }
        if (b.any_upsampling) {
          ScopedTimer t(dec, sp, kKFilters);
          k_upsample<<<dim3((dec->max_up_xsize + 31) / 32, (dec->max_up_ysize + 7) / 8, nf), px_block, 0, sp>>>(V, f0);
          launches++;
        }
        {
          ScopedTimer t(dec, sp, kKColorWrite);
          const dim3 cw_grid((dec->max_up_xsize + 127) / 128, (dec->max_up_ysize + 7) / 8, nf);
          k_color_write<<<cw_grid, px_block, 0, sp>>>(V, f0, skip_fused);
          launches++;
        }
      }
      if (to_host) CUDA_OK(copy_out(sp, b.vframe_frame.data() + f0, nf));  // this wave's frames leave while the next computes
    }
    if (turns) {
      PixelTurn& turn = g_pixel_turn[dec->device];
      CUDA_OK(cudaEventRecord(dec->pix_done, sp));
      turn.last = dec->pix_done;
      turn.owner = dec;
    }
  }
  if (parted) {
    CUDA_OK(cudaEventRecord(dec->part_ev[2], sp));
    CUDA_OK(cudaStreamWaitEvent(caller, dec->part_ev[2], 0));
    CUDA_OK(cudaEventRecord(dec->part_ev[2], s));  // (entropy-side kernels of Modular frames, if any, end here)
    CUDA_OK(cudaStreamWaitEvent(caller, dec->part_ev[2], 0));
  }
  rt_waves = rt_ms();
  if (to_host) {  // the Modular frames (written by the output kernel above)
    std::vector<uint32_t> rest;
    for (uint32_t i = 0; i < copied.size(); i++)
      if (!copied[i]) rest.push_back(i);
    CUDA_OK(copy_out(caller, rest.data(), rest.size()));
    CUDA_OK(cudaEventRecord(dec->copy_done, dec->copy_stream));
    dec->copy_pending = true;
  }
  if (dec->profiling) dec->profiled_runs++;
  dec->launches = launches;
  CUDA_OK(cudaGetLastError());
  if (run_timing)
    std::fprintf(stderr, "Run enqueue: entropy kernels %.2f ms, + waves %.2f ms, + tail %.2f ms (%u launches, to_host %d)\n", rt_entropy,
                 rt_waves, rt_ms(), launches, to_host ? 1 : 0);
  return 0;
}

int JxlB200DecoderRunToHost(JxlB200Decoder* dec, void* cuda_stream, void* const* dsts, const size_t* sizes, size_t n) {
  if (!dec || !dec->plan || n != dec->plan->frames.size() || !dsts || !sizes) return 1;
  for (size_t i = 0; i < n; i++) {
    if (!dsts[i] || sizes[i] < dec->plan->frame_out_size[i]) {
      dec->error = "output buffer too small";
      return 1;
    }
  }
  const auto t0 = std::chrono::steady_clock::now();
  CUDA_OK(cudaSetDevice(dec->device));
  FinishCopies(dec);
  dec->host_dsts.assign(dsts, dsts + n);
  const int rc = JxlB200DecoderRun(dec, cuda_stream);
  if (rc != 0) FinishCopies(dec);
  if (std::getenv("JXLB200_RUN_TIMING"))
    std::fprintf(stderr, "RunToHost: %.2f ms in the call\n",
                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  return rc;
}

int JxlB200DecoderSetPhaseMask(JxlB200Decoder* dec, uint32_t mask) {
  if (!dec) return 1;
  dec->phase_mask = mask;
  return 0;
}

int JxlB200DecoderSetProfiling(JxlB200Decoder* dec, int enabled) {
  if (!dec) return 1;
  FoldEvents(dec);
  dec->profiling = enabled != 0;
  for (double& m : dec->kernel_ms) m = 0;
  dec->profiled_runs = 0;
  return 0;
}

int JxlB200DecoderGetKernelTimes(JxlB200Decoder* dec, double* ms4, uint32_t* runs) {
  if (!dec || !ms4 || !runs) return 1;
  FoldEvents(dec);
  for (int k = 0; k < 4; k++) ms4[k] = dec->kernel_ms[k];
  *runs = dec->profiled_runs;
  return 0;
}

int JxlB200DecoderGetKernelTimesEx(JxlB200Decoder* dec, double* ms, uint32_t n, uint32_t* runs) {
  if (!dec || !ms || !runs) return 1;
  FoldEvents(dec);
  for (uint32_t k = 0; k < n; k++) ms[k] = k < kNumKernelClasses ? dec->kernel_ms[k] : 0.0;
  *runs = dec->profiled_runs;
  return 0;
}

static int CheckOnce(JxlB200Decoder* dec, cudaStream_t s, bool* overflow) {
  BatchPlan& b = *dec->plan;
  *overflow = false;
  dec->h_status.resize(b.streams.size());
  if (dec->h_status.size() != b.streams.size()) {
    dec->error = "out of memory (pinned status words)";
    return 1;
  }
  if (!dec->h_status.empty())
    CUDA_OK(cudaMemcpyAsync(dec->h_status.data(), dec->d_status.p, dec->h_status.size() * 4, cudaMemcpyDeviceToHost, s));
  if (!b.vframes.empty()) {
    dec->h_ac_status.resize(b.ac_streams.size());
    dec->h_ac_used.resize(b.ac_streams.size());
    dec->h_dc_status.resize(dec->dcg_list.size());
    if (dec->h_ac_status.size() != b.ac_streams.size() || dec->h_ac_used.size() != b.ac_streams.size() ||
        dec->h_dc_status.size() != dec->dcg_list.size()) {
      dec->error = "out of memory (pinned status words)";
      return 1;
    }
    CUDA_OK(cudaMemcpyAsync(dec->h_ac_status.data(), dec->d_ac_status.p, dec->h_ac_status.size() * 4, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaMemcpyAsync(dec->h_ac_used.data(), dec->d_ac_used.p, dec->h_ac_used.size() * 4, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaMemcpyAsync(dec->h_dc_status.data(), dec->d_dc_status.p, dec->h_dc_status.size() * 4, cudaMemcpyDeviceToHost, s));
  }
  CUDA_OK(BlockingSync(dec, s));
  for (size_t i = 0; i < dec->h_status.size(); i++) {
    if (dec->h_status[i] != 0) {
      dec->error = "entropy-coded stream " + std::to_string(i) + " failed (status " + std::to_string(dec->h_status[i]) +
                   ": 1 = read past section end, 2 = bad ANS final state, 4 = unsupported chained stream header)";
      return 1;
    }
  }
  if (!b.vframes.empty()) {
    for (size_t i = 0; i < dec->h_dc_status.size(); i++) {
      if (dec->h_dc_status[i] != 0) {
        dec->error = "DC group " + std::to_string(i) + ": corrupted AC metadata (status " + std::to_string(dec->h_dc_status[i]) + ")";
        return 1;
      }
    }
    for (size_t i = 0; i < dec->h_ac_status.size(); i++) {
      const uint32_t st = dec->h_ac_status[i];
      if (st == kVTokenOverflow) {
        *overflow = true;
      } else if (st != 0) {
        dec->error = "AC stream " + std::to_string(i) + " failed (status " + std::to_string(st) +
                     ": 1 = read past section end, 2 = bad ANS final state, 8 = corrupted, 16 = coefficient outside 16 bits)";
        return 1;
      }
    }
  }
  return 0;
}

static int WaitForRun(JxlB200Decoder* dec, void* cuda_stream);

int JxlB200DecoderWait(JxlB200Decoder* dec, void* cuda_stream) {
  if (!dec || !dec->plan) return 1;
  const int rc = WaitForRun(dec, cuda_stream);
  FinishCopies(dec);  // (JxlB200DecoderRunToHost: the frames are in the caller's buffers when Wait returns)
  return rc;
}

static int WaitForRun(JxlB200Decoder* dec, void* cuda_stream) {
  CUDA_OK(cudaSetDevice(dec->device));
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : dec->stream;
  bool overflow = false;
  if (CheckOnce(dec, s, &overflow) != 0) return 1;
  if (!overflow) return 0;
  // Some AC stream held more non-zero coefficients than its token budget (sized from the
  // section's byte count): lay the token arena out again with what was needed and decode again.
  if (!GrowTokenCapacity(dec->plan.get(), dec->h_ac_used.data())) {
    dec->error = "internal: token overflow without growth";
    return 1;
  }
  if (UploadTokensLayout(dec) != 0) return 1;
  if (JxlB200DecoderRun(dec, cuda_stream) != 0) return 1;
  if (CheckOnce(dec, s, &overflow) != 0) return 1;
  if (overflow) {
    dec->error = "internal: token overflow after growth";
    return 1;
  }
  return 0;
}

size_t JxlB200DecoderDeviceOutputBytes(const JxlB200Decoder* dec) { return dec && dec->plan ? dec->plan->out_size : 0; }

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
  CUDA_OK(BlockingSync(dec, dec->stream));
  return 0;
}

int JxlB200DecoderGetStats(const JxlB200Decoder* dec, JxlB200Stats* st) {
  if (!dec || !dec->plan || !st) return 1;
  const BatchPlan& b = *dec->plan;
  st->compressed_bytes = b.compressed_bytes;
  st->output_bytes = 0;
  for (uint64_t s : b.frame_out_size) st->output_bytes += s;
  st->pixels = b.total_pixels;
  st->num_streams = b.streams.size() + b.ac_streams.size();
  st->arena_bytes = b.arena_size * 4 + b.farena_size * 4 + b.barena_size + b.uarena_size * 4 + b.tok_size * 4;
  st->kernel_launches = dec->launches ? dec->launches
                                      : (b.streams.empty() ? 0 : 1) + (b.group_programs.empty() ? 0 : 1) + b.levels.size() + 1;
  st->num_ac_streams = b.ac_streams.size();
  st->vardct_frames = b.vframes.size();
  st->wave_frames = b.vframes.empty() ? 0 : b.wave_frames;
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
  // libjxl copies the JxlMemoryManager struct and allocates through it (lib/jxl/decode.cc:772-787); jpegxl-rs may pass
  // one (jpegxl-rs/src/memory.rs:24-40). The GPU path owns device memory and pinned staging, which a host allocator
  // callback cannot provide: the manager is accepted and its callbacks are never invoked.
  (void)memory_manager;
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
    dec->stage = 5;
    if (dec->events_wanted & JXL_DEC_BASIC_INFO) return JXL_DEC_BASIC_INFO;
  }
  if (dec->stage == 5) {
    // libjxl emits COLOR_ENCODING after the headers and before the first frame (lib/jxl/decode.cc:1149-1500); jpegxl-rs
    // subscribes to it when `icc_profile` is set and then asks for the ICC profile (jpegxl-rs/src/decode.rs:334-347):
    // the ICC calls below fail loudly (ICC synthesis is SURVEY.md 8f N2, not built) instead of an empty profile.
    dec->stage = 1;
    if (dec->events_wanted & JXL_DEC_COLOR_ENCODING) return JXL_DEC_COLOR_ENCODING;
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
    JxlB200DecoderSetKeepOrientation(dec->gpu, dec->keep_orientation ? 1 : 0);
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
  if (info) FillBasicInfo(dec->info, dec->have_container, info, dec->keep_orientation);
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
  const bool transpose = !dec->keep_orientation && dec->info.meta.orientation >= 5;
  *size = transpose ? OutputStride(dec->info.ysize, f) * dec->info.xsize : OutputStride(dec->info.xsize, f) * dec->info.ysize;
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

// jpegxl-rs asks for the ICC profile only when the `icc_profile` option is set (jpegxl-rs/src/decode.rs:329-357).
// libjxl synthesises one from the enumerated colour encoding (lib/jxl/cms/jxl_cms_internal.h); that generator and the
// ICC codec are container / metadata work, not built: the calls fail, as libjxl's do when it cannot
// produce a profile.
JxlDecoderStatus JxlDecoderGetICCProfileSize(const JxlDecoder* dec, int target, size_t* size) {
  (void)dec;
  (void)target;
  if (size) *size = 0;
  return JXL_DEC_ERROR;
}

JxlDecoderStatus JxlDecoderGetColorAsICCProfile(const JxlDecoder* dec, int target, uint8_t* icc_profile, size_t size) {
  (void)dec;
  (void)target;
  (void)icc_profile;
  (void)size;
  return JXL_DEC_ERROR;
}

// JPEG reconstruction is not built: JXL_DEC_JPEG_RECONSTRUCTION is never emitted, so jpegxl-rs
// never reaches these (jpegxl-rs/src/decode.rs:260-283); they refuse, like libjxl outside that event.
JxlDecoderStatus JxlDecoderSetJPEGBuffer(JxlDecoder* dec, uint8_t* data, size_t size) {
  (void)dec;
  (void)data;
  (void)size;
  return JXL_DEC_ERROR;
}

size_t JxlDecoderReleaseJPEGBuffer(JxlDecoder* dec) {
  (void)dec;
  return 0;
}

}  // extern "C"

// ================================================================== encoder
namespace jxlb {

// Every encoder kernel covers the whole batch: the last grid dimension is the frame.
__device__ __forceinline__ DevEPools FramePools(const DevEPools& E, const DevEFrame& ef) {
  DevEPools El = E;
  El.tree = E.tree + ef.tree_off;
  return El;
}

__global__ void __launch_bounds__(256) k_enc_xyb(DevEPools E, const DevEFrame* frames) {
  const DevEFrame& ef = frames[blockIdx.z];
  const uint32_t x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x < ef.xblocks * 8 && y < ef.yblocks * 8) DevEncXybPixel(E, ef, x, y);
}

// blockIdx.z = frame * 3 + channel
__global__ void __launch_bounds__(256) k_enc_gaborish_inv(DevEPools E, const DevEFrame* frames) {
  const DevEFrame& ef = frames[blockIdx.z / 3];
  const uint32_t c = blockIdx.z % 3;
  const int32_t x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  const int32_t PW = ef.xblocks * 8, PH = ef.yblocks * 8;
  if (!ef.gab || x >= PW || y >= PH) return;
  if (x >= 2 && y >= 2 && x + 2 < PW && y + 2 < PH) {
    DevEncGaborishInvPixel<true>(E, ef, c, x, y);
  } else {
    DevEncGaborishInvPixel<false>(E, ef, c, x, y);
  }
}

// blockIdx.x = tile (64x64 pixels), blockIdx.y = frame: libjxl's initial quantisation field (E3).
__global__ void __launch_bounds__(256) k_enc_aq(DevEPools E, const DevEFrame* frames) {
  __shared__ float aq_sm[kAqSmemFloats];
  const DevEFrame& ef = frames[blockIdx.y];
  const uint32_t tiles_x = (ef.xblocks + 7) / 8, tiles_y = (ef.yblocks + 7) / 8;
  if (!ef.adaptive || blockIdx.x >= tiles_x * tiles_y) return;
  DevEncAqTile<2>(E, ef, blockIdx.x % tiles_x, blockIdx.x / tiles_x, threadIdx.x, blockDim.x, aq_sm);
}

// one thread per 64x64-pixel tile (the greedy choice is serial inside a tile; tiles are independent: DevEncStrategyTile)
__global__ void __launch_bounds__(64) k_enc_strategy(DevEPools E, const DevEFrame* frames) {
  const DevEFrame& ef = frames[blockIdx.y];
  const uint32_t tiles_x = (ef.xblocks + 7) / 8, tiles_y = (ef.yblocks + 7) / 8;
  const uint32_t t = blockIdx.x * 64 + threadIdx.x;
  if (t < tiles_x * tiles_y) DevEncStrategyTile(E, ef, (t % tiles_x) * 8, (t / tiles_x) * 8);
}

__global__ void __launch_bounds__(32) k_enc_number(DevEPools E, const DevEFrame* frames) {
  const DevEFrame& ef = frames[blockIdx.y];
  const uint32_t g = blockIdx.x * 32 + threadIdx.x;
  if (g < ef.xdcgroups * ef.ydcgroups) DevEncNumberBlocks(E, ef, g);
}

__global__ void __launch_bounds__(256) k_enc_dc(DevEPools E, const DevEFrame* frames) {
  const DevEFrame& ef = frames[blockIdx.y];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ef.xblocks * ef.yblocks) DevEncDcBlock(E, ef, i % ef.xblocks, i / ef.xblocks);
}

constexpr uint32_t kEncThreads = 128;
constexpr uint32_t kEncSmemFloats = 4 * 4096;  // four buffers of a 64x64 varblock, or one 32x32 set per warp

// blockIdx.x = tile (64x64 pixels), blockIdx.y = frame: chroma-from-luma fit over the float coefficients.
__global__ void __launch_bounds__(128) k_enc_cfl(DevEPools E, const DevEFrame* frames) {
  extern __shared__ float cfl_vals[];  // 4 * 4096 floats
  __shared__ float red[64];
  const DevEFrame& ef = frames[blockIdx.y];
  if (blockIdx.x >= ef.cmw * ef.cmh) return;
  DevEncCflTile<2>(E, ef, blockIdx.x % ef.cmw, blockIdx.x / ef.cmw, threadIdx.x, blockDim.x, cfl_vals, red);
}

// blockIdx.x = group: MODE 0 forward transforms, MODE 1 quantisation; warp per varblock up to 32x32, CTA per 64x64 class.
template <int MODE>
__global__ void __launch_bounds__(kEncThreads) k_enc_coeffs(DevEPools E, const DevEFrame* frames) {
  extern __shared__ float enc_smem[];
  __shared__ uint32_t next_s, has_big_s;
  const DevEFrame& ef = frames[blockIdx.y];
  const uint32_t g = blockIdx.x;
  if (g >= ef.xgroups * ef.ygroups) return;
  const uint32_t x0 = (g % ef.xgroups) * 32, y0 = (g / ef.xgroups) * 32;
  const uint32_t xs = min(32u, ef.xblocks - x0), ys = min(32u, ef.yblocks - y0);
  const uint8_t* acs = E.barena + ef.acs;
  if (threadIdx.x == 0) {
    next_s = 0;
    has_big_s = 0;
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* wbuf = enc_smem + warp * 4096;
  const uint32_t total = xs * ys;
  for (;;) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(&next_s, 1u);
    i = __shfl_sync(0xFFFFFFFFu, i, 0);
    if (i >= total) break;
    const uint32_t bx = i % xs, by = i / xs;
    const uint8_t a = acs[static_cast<size_t>(y0 + by) * ef.xblocks + x0 + bx];
    if (!(a & 1)) continue;
    const StrategyInfo si = UnpackStrategyInfo(E.upool[E.sinfo_off + (a >> 1)]);
    if (static_cast<uint32_t>(si.cx) * si.cy > 16) {
      if (lane == 0) has_big_s = 1;
      continue;
    }
    if (MODE == 0 && (a >> 1) == 0) {  // DCT8X8, most of a photograph: in registers
      DevEncDct8Warp(E, ef, x0 + bx, y0 + by, wbuf, lane);
      continue;
    }
    DevEncVarblock<1, MODE>(E, ef, x0 + bx, y0 + by, a >> 1, wbuf, lane, 32);
  }
  __syncthreads();
  if (!has_big_s) return;
  for (uint32_t i = 0; i < total; i++) {
    const uint32_t bx = i % xs, by = i / xs;
    const uint8_t a = acs[static_cast<size_t>(y0 + by) * ef.xblocks + x0 + bx];
    if (!(a & 1)) continue;
    const StrategyInfo si = UnpackStrategyInfo(E.upool[E.sinfo_off + (a >> 1)]);
    if (static_cast<uint32_t>(si.cx) * si.cy <= 16) continue;
    DevEncVarblock<2, MODE>(E, ef, x0 + bx, y0 + by, a >> 1, enc_smem, threadIdx.x, kEncThreads);
  }
}

// AdjustQuantBlockAC: thread per (block, channel); blockIdx.y = channel, blockIdx.z = frame.
__global__ void __launch_bounds__(128) k_enc_adjust(DevEPools E, const DevEFrame* frames) {
  const DevEFrame& ef = frames[blockIdx.z];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ef.xblocks * ef.yblocks) DevEncAdjustVarblockChannel(E, ef, i % ef.xblocks, i / ef.xblocks, blockIdx.y);
}

// Coefficient-order statistics: thread per group, then CTA per group (see DevEncOrderStatsGroup).
__global__ void __launch_bounds__(32) k_enc_group_orders(DevEPools E, const DevEFrame* frames) {
  const DevEFrame& ef = frames[blockIdx.y];
  const uint32_t g = blockIdx.x * 32 + threadIdx.x;
  if (g < ef.xgroups * ef.ygroups) DevEncGroupOrders(E, ef, g);
}

__global__ void __launch_bounds__(256) k_enc_order_stats(DevEPools E, const DevEFrame* frames) {
  __shared__ uint32_t local[kCustomOrderCounters + 1024];
  const DevEFrame& ef = frames[blockIdx.y];
  if (blockIdx.x >= ef.xgroups * ef.ygroups) return;
  DevEncOrderStatsGroup<2>(E, ef, blockIdx.x, threadIdx.x, blockDim.x, local);
}

// Tokenisation, data-parallel: thread per (block, channel) for the statistics and for the tokens, thread per group
// for the offsets in between. blockIdx.y = channel, blockIdx.z = frame.
__global__ void __launch_bounds__(128) k_enc_block_stats(DevEPools E, const DevEFrame* frames) {
  const DevEFrame& ef = frames[blockIdx.z];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ef.xblocks * ef.yblocks) DevEncBlockStats(E, ef, i % ef.xblocks, i / ef.xblocks, blockIdx.y);
}

__global__ void __launch_bounds__(32) k_enc_token_offsets(DevEPools E, const DevEFrame* frames) {
  const DevEFrame& ef = frames[blockIdx.y];
  const uint32_t g = blockIdx.x * 32 + threadIdx.x;
  if (g < ef.xgroups * ef.ygroups) DevEncTokenOffsets(E, ef, g);
}

__global__ void __launch_bounds__(128) k_enc_block_tokens(DevEPools E, const DevEFrame* frames) {
  __shared__ uint16_t ctxtab_s[128];
  for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) ctxtab_s[i] = static_cast<uint16_t>(E.upool[E.ctxtab_off + i]);
  __syncthreads();
  const DevEFrame& ef = frames[blockIdx.z];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ef.xblocks * ef.yblocks) DevEncBlockTokens(E, ef, i % ef.xblocks, i / ef.xblocks, blockIdx.y, ctxtab_s, ctxtab_s + 64);
}

// blockIdx.y = DC group, blockIdx.z = frame; one thread per sample of the DC + AC-metadata streams
__global__ void __launch_bounds__(256) k_enc_modular(DevEPools E, const DevEFrame* frames) {
  const DevEFrame& ef = frames[blockIdx.z];
  const uint32_t g = blockIdx.y;
  if (g >= ef.xdcgroups * ef.ydcgroups) return;
  const DevEPools El = FramePools(E, ef);
  const DevDcGroupLayout L = DevDcGroupGeometry(El, ef, g);
  const uint32_t total = L.dc_tokens + L.meta_tokens;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    DevEncModularSample(El, ef, g, L, i);
}

// Alpha extra channel: one thread per pixel (DevEncAlphaSample); blockIdx.y = frame.
__global__ void __launch_bounds__(256) k_enc_alpha(DevEPools E, const DevEFrame* frames) {
  const DevEFrame& ef = frames[blockIdx.y];
  if (!ef.has_alpha) return;
  const DevEPools El = FramePools(E, ef);
  const uint64_t n = static_cast<uint64_t>(ef.xsize) * ef.ysize;
  for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) DevEncAlphaSample(El, ef, i);
}

// The halves of the DC-group sections: the longest chains of a frame (up to 196 K tokens), one CTA of one warp per
// SM, the reverse tables of the half's most used contexts staged in shared memory (`slots`: per (frame, half) 64 bytes,
// cluster -> slot or 0xFF; kEmitDcSlots tables of 8 KB): the load on the coder's serial chain is a shared-memory load
// instead of an L2 hit. blockIdx.x = half * dc_blocks + DC group.
constexpr uint32_t kEmitDcSlots = 27;
__global__ void __launch_bounds__(32) k_enc_emit_dc(DevEPools E, const DevEFrame* frames, const uint32_t* fs_tables,
                                                    const uint16_t* rev_tables, uint32_t* words, const uint64_t* off, uint64_t* first,
                                                    uint32_t dc_blocks, const uint8_t* slots) {
  extern __shared__ uint4 emit_smem[];
  __shared__ uint8_t sslot[64];
  __shared__ uint4 stage[32];
  const DevEFrame& ef = frames[blockIdx.y];
  const uint32_t lane = threadIdx.x;
  const uint32_t half = blockIdx.x >= dc_blocks ? 1 : 0, g = blockIdx.x - half * dc_blocks;
  const uint32_t ndc = ef.xdcgroups * ef.ydcgroups;
  if (g >= ndc) return;
  const DevEncCode code{fs_tables + ef.code_off[0], rev_tables + ef.code_off[1]};
  const uint8_t* my_slots = slots + (static_cast<size_t>(blockIdx.y) * 2 + half) * 64;
  sslot[lane] = my_slots[lane];
  sslot[lane + 32] = my_slots[lane + 32];
  __syncwarp();
  for (uint32_t c = 0; c < 64; c++) {  // (8 KB per staged cluster, 16 bytes per lane and step)
    const uint32_t slot = sslot[c];
    if (slot == 0xFF) continue;
    const uint4* src = reinterpret_cast<const uint4*>(code.reverse + c * 4096);
    for (uint32_t i = lane; i < 512; i += 32) emit_smem[slot * 512 + i] = src[i];
  }
  __syncwarp();
  const uint32_t sec = ef.sec_base + g + half * (ndc + ef.xgroups * ef.ygroups);
  const uint64_t pos = DevEncEmitDcGroupWarp(E, ef, g, code, words, off[sec + 1] * 32, lane, half, stage,
                                             reinterpret_cast<const uint16_t*>(emit_smem), sslot);
  if (lane == 0) first[sec] = pos;
}

// rANS emission: one warp per section (DevRansPushWarp: the lanes fetch and split the tokens, lane 0 codes), written
// back to front so that it ends at the end of its region (`off[sec]` .. `off[sec + 1]`, in words); `first[sec]`
// receives the bit position of its first bit. This kernel: the AC-group sections (sections [ndc, ndc + ngroups) of a
// frame; [0, ndc) = DC halves, [ndc + ngroups, 2 ndc + ngroups) = metadata halves of the DC-group sections:
// k_enc_emit_dc, on a second stream at the same time).
__global__ void __launch_bounds__(32) k_enc_emit(DevEPools E, const DevEFrame* frames, const uint32_t* fs_tables,
                                                 const uint16_t* rev_tables, uint32_t* words, const uint64_t* off, uint64_t* first) {
  __shared__ uint4 stage[32];
  const DevEFrame& ef = frames[blockIdx.y];
  const uint32_t lane = threadIdx.x;
  const uint32_t g = blockIdx.x;
  if (g >= ef.xgroups * ef.ygroups) return;
  const DevEncCode code{fs_tables + ef.code_off[2], rev_tables + ef.code_off[3]};
  const uint32_t n = static_cast<uint32_t>(E.iarena[ef.group_tokens + g]);
  const uint32_t sec = ef.sec_base + ef.xdcgroups * ef.ydcgroups + g;
  if (ef.has_alpha && !(ef.xsize <= 256 && ef.ysize <= 256)) {  // the group's alpha stream follows its coefficients
    const DevEncCode mod{fs_tables + ef.code_off[0], rev_tables + ef.code_off[1]};
    const uint32_t gx = g % ef.xgroups, gy = g / ef.xgroups;
    const uint32_t gw = ef.xsize - (gx << 8) < 256 ? ef.xsize - (gx << 8) : 256, gh = ef.ysize - (gy << 8) < 256 ? ef.ysize - (gy << 8) : 256;
    const uint64_t pos = DevEncEmitAcGroupWarp(E.tokens + ef.ac_tokens + static_cast<size_t>(g) * 3 * 65536, n, code, words,
                                               off[sec + 1] * 32, lane, stage, E.tokens + ef.alpha_tokens + static_cast<size_t>(g) * 65536,
                                               gw * gh, &mod);
    if (lane == 0) first[sec] = pos;
    return;
  }
  const uint64_t pos = DevEncEmitAcGroupWarp(E.tokens + ef.ac_tokens + static_cast<size_t>(g) * 3 * 65536, n, code, words,
                                             off[sec + 1] * 32, lane, stage);
  if (lane == 0) first[sec] = pos;
}

// Section k's used words: ranges[3k] = first source word, [3k + 1] = destination word, [3k + 2] = count.
__global__ void __launch_bounds__(256) k_enc_compact(const uint32_t* words, uint32_t* out, const uint64_t* ranges) {
  const uint64_t src = ranges[3 * blockIdx.x], dst = ranges[3 * blockIdx.x + 1], n = ranges[3 * blockIdx.x + 2];
  for (uint64_t i = threadIdx.x; i < n; i += 256) out[dst + i] = words[src + i];
}

// ---- lossless (Modular) encoder: kernels/jxlb_encl_dev.h
__global__ void __launch_bounds__(256) k_encl_planes(DevLPools L, const DevLFrame* frames) {
  const DevLFrame& f = frames[blockIdx.y];
  const uint64_t n = static_cast<uint64_t>(f.xsize) * f.ysize;
  for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) DevEnclPlanes(L, f, i);
}

__global__ void __launch_bounds__(256) k_encl_tokens(DevLPools L, const DevLFrame* frames) {
  const DevLFrame& f = frames[blockIdx.y];
  const uint64_t n = static_cast<uint64_t>(f.xsize) * f.ysize * f.nch;
  for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) DevEnclToken(L, f, i);
}

// One warp per group section (a serial rANS chain each, DevRansPushWarp); `off[sec]` .. `off[sec + 1]` is the section's region in words.
__global__ void __launch_bounds__(32) k_encl_emit(DevLPools L, const DevLFrame* frames, const uint32_t* fs_tables,
                                                  const uint16_t* rev_tables, uint32_t* words, const uint64_t* off, uint64_t* first) {
  __shared__ uint4 stage[32];
  const DevLFrame& f = frames[blockIdx.y];
  const uint32_t g = blockIdx.x;
  if (g >= f.xgroups * f.ygroups) return;
  const DevEncCode code{fs_tables + f.code_off[0], rev_tables + f.code_off[1]};
  const uint32_t sec = f.sec_base + g;
  const bool global_only = f.xsize <= kEnclGroupDim && f.ysize <= kEnclGroupDim;  // the tokens follow the host's global header
  const uint64_t pos = DevEnclEmitGroupWarp(L, f, g, code, words, off[sec + 1] * 32, !global_only, threadIdx.x, stage);
  if (threadIdx.x == 0) first[sec] = pos;
}

}  // namespace jxlb

struct JxlB200Encoder {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string error;
  std::vector<std::vector<uint8_t>> outputs;
  double phase_ms[3] = {0, 0, 0};  // kernels before the histogram sync, host table building, emit kernels
  // device buffers, kept across batches (a batch of 32 4K frames uses ~20 GB: allocating and freeing that per call
  // costs more than the kernels)
  DevBuf<uint8_t> d_in, d_barena, d_cluster, d_sample;
  DevBuf<float> d_farena, d_fpool, d_lut;
  DevBuf<int32_t> d_iarena;
  DevBuf<uint2> d_tokens;
  DevBuf<uint16_t> d_opool, d_custom, d_rev;
  DevBuf<uint32_t> d_upool, d_words, d_fs;
  DevBuf<uint8_t> d_slots;        // k_enc_emit_dc: cluster -> shared-memory slot per (frame, half)
  cudaStream_t stream2 = nullptr;  // the DC halves are emitted next to the AC groups
  DevBuf<uint64_t> d_off, d_bits, d_ranges;
  DevBuf<uint32_t> d_compact;  // FetchSections
  DevBuf<DevEncTreeNode> d_trees;
  DevBuf<DevEFrame> d_efs;
  DevBuf<DevLFrame> d_lfs;      // lossless encoder
  DevBuf<int32_t> d_lconst;     // its cutoffs and leaf table
  uint8_t* h_in = nullptr;      // pinned staging of the input frames (UploadFrames); read by the transfers of the
  size_t h_in_cap = 0;          // current call only: every call ends with a stream synchronisation
  uint32_t* h_words = nullptr;  // pinned: the emitted sections of a batch
  size_t h_words_cap = 0;
};

#undef CUDA_OK
#define CUDA_OK(expr)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      enc->error = std::string(#expr) + ": " + cudaGetErrorString(e_);                       \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

extern "C" {

JxlB200Encoder* JxlB200EncoderCreate(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= device || device < 0) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  JxlB200Encoder* enc = new JxlB200Encoder();
  enc->device = device;
  if (cudaStreamCreateWithFlags(&enc->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete enc;
    return nullptr;
  }
  cudaFuncSetAttribute(k_enc_coeffs<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncSmemFloats * sizeof(float));
  cudaFuncSetAttribute(k_enc_coeffs<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncSmemFloats * sizeof(float));
  cudaFuncSetAttribute(k_enc_cfl, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4096 * sizeof(float));
  cudaFuncSetAttribute(k_enc_emit_dc, cudaFuncAttributeMaxDynamicSharedMemorySize, kEmitDcSlots * 8192);
  return enc;
}

void JxlB200EncoderDestroy(JxlB200Encoder* enc) {
  if (!enc) return;
  cudaSetDevice(enc->device);
  if (enc->h_words) cudaFreeHost(enc->h_words);
  if (enc->h_in) cudaFreeHost(enc->h_in);
  if (enc->stream2) cudaStreamDestroy(enc->stream2);
  if (enc->stream) cudaStreamDestroy(enc->stream);
  delete enc;
}

const char* JxlB200EncoderGetError(const JxlB200Encoder* enc) { return enc ? enc->error.c_str() : "null encoder"; }

// The caller's frames (pageable memory as a rule: a direct cudaMemcpyAsync stages them on one thread at a few GB/s) ->
// device: host threads copy them into pinned staging, each enqueuing its frame's transfer as soon as its copy is done.
static int UploadFrames(JxlB200Encoder* enc, cudaStream_t s, const void* const* src, const std::vector<uint64_t>& dev_off,
                        const std::vector<size_t>& sizes) {
  JxlB200Encoder* dec = enc;  // (CUDA_OK reports through ->error)
  const size_t n = sizes.size();
  std::vector<size_t> host_off(n);
  size_t total = 0;
  for (size_t i = 0; i < n; i++) {
    host_off[i] = total;
    total += (sizes[i] + 255) & ~size_t{255};
  }
  if (total > enc->h_in_cap) {
    if (enc->h_in) cudaFreeHost(enc->h_in);
    enc->h_in = nullptr;
    enc->h_in_cap = 0;
    if (cudaHostAlloc(reinterpret_cast<void**>(&enc->h_in), total, cudaHostAllocDefault) == cudaSuccess) enc->h_in_cap = total;
  }
  if (enc->h_in_cap < total) {  // no pinned memory to be had: the plain way
    for (size_t i = 0; i < n; i++) CUDA_OK(cudaMemcpyAsync(enc->d_in.p + dev_off[i], src[i], sizes[i], cudaMemcpyHostToDevice, s));
    return 0;
  }
  std::atomic<size_t> next{0};
  std::atomic<int> failed{0};
  auto work = [&]() {
    cudaSetDevice(enc->device);
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= n) break;
      std::memcpy(enc->h_in + host_off[i], src[i], sizes[i]);
      if (cudaMemcpyAsync(enc->d_in.p + dev_off[i], enc->h_in + host_off[i], sizes[i], cudaMemcpyHostToDevice, s) != cudaSuccess) failed = 1;
    }
  };
  const size_t nthreads = std::max<size_t>(1, std::min<size_t>(n, std::min<unsigned>(16, std::thread::hardware_concurrency())));
  std::vector<std::thread> pool;
  for (size_t t = 0; t < nthreads; t++) pool.emplace_back(work);
  for (auto& t : pool) t.join();
  if (failed) {
    enc->error = "upload of the input frames failed";
    return 1;
  }
  return 0;
}

// The sections were written back to front into regions sized for the worst case (6 bytes per token): what a section
// really uses is the tail of its region. Reads where every section starts, packs the used words of all sections into
// one buffer on the device (k_enc_compact, a CTA per section) and copies that -- 0.13 GB instead of 2.5 GB per 16
// lossless 4K frames. On return (stream synchronised) section k is bits [first[k], first[k] + nbits[k]) of enc->h_words.
static int FetchSections(JxlB200Encoder* enc, cudaStream_t s, const std::vector<uint64_t>& h_off, size_t nsec,
                         std::vector<uint64_t>* first, std::vector<uint64_t>* nbits) {
  JxlB200Encoder* dec = enc;  // (CUDA_OK reports through ->error)
  std::vector<uint64_t> h_bits(nsec);
  CUDA_OK(cudaMemcpyAsync(h_bits.data(), enc->d_bits.p, nsec * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  std::vector<uint64_t> ranges(3 * nsec);  // source word, destination word, count
  first->resize(nsec);
  nbits->resize(nsec);
  uint64_t total = 0;
  for (size_t k = 0; k < nsec; k++) {
    const uint64_t w0 = h_bits[k] >> 5, w1 = h_off[k + 1];
    if (w0 > w1 || w0 < h_off[k]) {
      enc->error = "internal: section outside its region";
      return 1;
    }
    ranges[3 * k] = w0;
    ranges[3 * k + 1] = total;
    ranges[3 * k + 2] = w1 - w0;
    (*first)[k] = total * 32 + (h_bits[k] & 31);
    (*nbits)[k] = w1 * 32 - h_bits[k];
    total += w1 - w0;
  }
  CUDA_OK(enc->d_ranges.Upload(ranges, s));
  CUDA_OK(enc->d_compact.Alloc(total + 16));
  k_enc_compact<<<static_cast<uint32_t>(nsec), 256, 0, s>>>(enc->d_words.p, enc->d_compact.p, enc->d_ranges.p);
  if (total + 16 > enc->h_words_cap) {  // pinned, kept across batches
    if (enc->h_words) cudaFreeHost(enc->h_words);
    enc->h_words = nullptr;
    enc->h_words_cap = 0;
    const size_t cap = total + 16 + total / 4;
    CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&enc->h_words), cap * sizeof(uint32_t), cudaHostAllocDefault));
    enc->h_words_cap = cap;
  }
  CUDA_OK(cudaMemcpyAsync(enc->h_words, enc->d_compact.p, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  CUDA_OK(cudaGetLastError());
  return 0;
}

int JxlB200EncoderEncodeBatch(JxlB200Encoder* enc, const uint8_t* const* rgb, const uint32_t* xsizes, const uint32_t* ysizes,
                              size_t n, const JxlB200EncodeOptions* opt) {
  if (!enc || !rgb || !xsizes || !ysizes || !opt || n == 0) return 1;
  enc->error.clear();
  enc->outputs.clear();
  EncParams p;
  p.distance = opt->distance;
  p.strategy_mode = opt->strategy_mode;
  p.gab = opt->gaborish != 0;
  p.epf_iters = opt->epf_iters;
  p.dc_smoothing = opt->dc_smoothing != 0;
  p.alpha = opt->has_alpha != 0;
  const size_t in_ch = p.alpha ? 4 : 3;
  if (!(p.strategy_mode == 0 || p.strategy_mode == 2) || p.epf_iters > 3 || !(p.distance > 0.0f)) {
    enc->error = "invalid encode options";
    return 1;
  }
  CUDA_OK(cudaSetDevice(enc->device));
  cudaStream_t s = enc->stream;
  try {
    const SharedVarDCTTables& sh = SharedVarDCTTables::Get();
    uint32_t num_ac_clusters = 0;
    const std::vector<uint8_t> ac_cluster_of = AcContextClusters(&num_ac_clusters);
    // ---- layout of every frame in the arenas
    struct Frame {
      DevEFrame ef;
      EncLayout L;
      EncTree tree;
      uint32_t global_scale, quant_dc;
      uint64_t tree_off;
      EncGlobals G;
      std::vector<uint64_t> ac_off, dc_off, meta_off;  // word offsets of the sections (DC halves, AC groups, metadata halves)
      uint64_t bits_off;                     // index of this frame's first entry in d_bits
    };
    std::vector<Frame> fr(n);
    uint64_t fbase = 0, ibase = 0, bbase = 0, tbase = 0, inbase = 0, treebase = 0;
    std::vector<DevEncTreeNode> all_trees;
    for (size_t i = 0; i < n; i++) {
      Frame& f = fr[i];
      JXLB_CHECK(xsizes[i] > 0 && ysizes[i] > 0 && xsizes[i] <= (1u << 16) && ysizes[i] <= (1u << 16), "bad image size");
      f.ef = DevEFrame{};
      FillQuantizer(p, &f.ef, &f.global_scale, &f.quant_dc);
      FrameHeader fh0;
      fh0.xsize = xsizes[i];
      fh0.ysize = ysizes[i];
      f.tree = BuildEncTree(ToFrameDimensions(fh0).num_dc_groups);
      f.L = LayoutEncFrame(xsizes[i], ysizes[i], num_ac_clusters, f.tree.num_leaves, &f.ef, p.alpha);
      DevEFrame& e = f.ef;
      for (int c = 0; c < 3; c++) {
        e.xyb[c] += fbase;
        e.xyb_raw[c] += fbase;
        e.coef[c] += ibase;
        e.dcq[c] += ibase;
        e.blk_nz[c] += ibase;
        e.blk_ntok[c] += ibase;
        e.blk_bucket[c] += ibase;
      }
      e.quant_field += fbase;
      e.adj_thres += fbase;
      e.adj_quant += ibase;
      e.raw_quant += bbase;
      e.first_index += ibase;
      e.block_of_num += ibase;
      e.order_mask += ibase;
      e.group_first += ibase;
      e.zero_counts += ibase;
      e.dcg_count += ibase;
      e.group_tokens += ibase;
      e.ac_hist += ibase;
      e.mod_hist += ibase;
      e.acs += bbase;
      e.ytox += bbase;
      e.ytob += bbase;
      e.ac_tokens += tbase;
      e.mod_tokens += tbase;
      e.alpha_tokens += tbase;
      e.rgb = inbase;
      f.tree_off = treebase;
      all_trees.insert(all_trees.end(), f.tree.nodes.begin(), f.tree.nodes.end());
      treebase += f.tree.nodes.size();
      fbase += f.L.fsize;
      ibase += f.L.isize;
      bbase += f.L.bsize;
      tbase += f.L.tsize;
      inbase += (static_cast<uint64_t>(xsizes[i]) * ysizes[i] * in_ch + 15) & ~uint64_t{15};
    }
    DevBuf<uint8_t>&d_in = enc->d_in, &d_barena = enc->d_barena, &d_cluster = enc->d_cluster;
    DevBuf<float>&d_farena = enc->d_farena, &d_fpool = enc->d_fpool, &d_lut = enc->d_lut;
    DevBuf<int32_t>& d_iarena = enc->d_iarena;
    DevBuf<uint2>& d_tokens = enc->d_tokens;
    DevBuf<uint16_t>& d_opool = enc->d_opool;
    DevBuf<uint32_t>& d_upool = enc->d_upool;
    DevBuf<DevEncTreeNode>& d_trees = enc->d_trees;
    CUDA_OK(d_in.Alloc(inbase + 16));
    CUDA_OK(d_farena.Alloc(fbase + 16));
    CUDA_OK(d_iarena.Alloc(ibase + 16));
    CUDA_OK(d_barena.Alloc(bbase + 16));
    CUDA_OK(d_tokens.Alloc(tbase + 16));
    CUDA_OK(d_fpool.Upload(sh.fpool, s));
    CUDA_OK(d_opool.Upload(sh.opool, s));
    CUDA_OK(d_upool.Upload(sh.upool, s));
    CUDA_OK(d_cluster.Upload(ac_cluster_of, s));
    CUDA_OK(d_trees.Upload(all_trees, s));
    std::vector<float> lut(256);
    for (int i = 0; i < 256; i++) lut[i] = SrgbToLinearHost(i / 255.0f);
    CUDA_OK(d_lut.Upload(lut, s));
    {
      std::vector<uint64_t> dev_off(n);
      std::vector<size_t> in_sizes(n);
      for (size_t i = 0; i < n; i++) {
        dev_off[i] = fr[i].ef.rgb;
        in_sizes[i] = static_cast<size_t>(xsizes[i]) * ysizes[i] * in_ch;
      }
      if (UploadFrames(enc, s, reinterpret_cast<const void* const*>(rgb), dev_off, in_sizes) != 0) return 1;
    }
    CUDA_OK(cudaMemsetAsync(d_iarena.p, 0, (ibase + 16) * sizeof(int32_t), s));
    CUDA_OK(cudaMemsetAsync(d_barena.p, 0xFF, bbase + 16, s));
    DevEPools E{};
    E.bytes_in = d_in.p;
    E.farena = d_farena.p;
    E.iarena = d_iarena.p;
    E.barena = d_barena.p;
    E.tokens = d_tokens.p;
    E.srgb_lut = d_lut.p;
    E.fpool = d_fpool.p;
    E.opool = d_opool.p;
    E.upool = d_upool.p;
    for (int i = 0; i < 17; i++) E.table_off[i] = sh.table_off[i];
    for (int i = 0; i < 13; i++) E.order_off[i] = sh.order_off[i];
    E.wc_off = sh.wc_off;
    E.sinfo_off = sh.sinfo_off;
    E.ctxtab_off = sh.ctxtab_off;
    E.ac_cluster_of = d_cluster.p;
    cudaEvent_t ev[4];
    for (auto& e : ev) CUDA_OK(cudaEventCreate(&e));
    CUDA_OK(cudaEventRecord(ev[0], s));
    // ---- phase 1: pixels -> tokens + histograms, all frames of the batch in every launch
    uint32_t maxW = 0, maxH = 0, max_groups = 0, max_dcg = 0;
    std::vector<DevEFrame> efs(n);
    for (size_t i = 0; i < n; i++) {
      fr[i].ef.tree_off = static_cast<uint32_t>(fr[i].tree_off);
      efs[i] = fr[i].ef;
      const FrameDimensions& d = fr[i].L.dim;
      maxW = std::max<uint32_t>(maxW, d.xsize_blocks);
      maxH = std::max<uint32_t>(maxH, d.ysize_blocks);
      max_groups = std::max<uint32_t>(max_groups, d.num_groups);
      max_dcg = std::max<uint32_t>(max_dcg, d.num_dc_groups);
    }
    DevBuf<DevEFrame>& d_efs = enc->d_efs;
    CUDA_OK(d_efs.Upload(efs, s));
    E.tree = d_trees.p;
    const uint32_t nf = static_cast<uint32_t>(n);
    k_enc_xyb<<<dim3((maxW * 8 + 31) / 32, maxH, nf), dim3(32, 8), 0, s>>>(E, d_efs.p);
    if (p.gab) k_enc_gaborish_inv<<<dim3((maxW * 8 + 31) / 32, maxH, nf * 3), dim3(32, 8), 0, s>>>(E, d_efs.p);
    if (p.adaptive_quant) k_enc_aq<<<dim3(((maxW + 7) / 8) * ((maxH + 7) / 8), nf), 256, 0, s>>>(E, d_efs.p);
    k_enc_strategy<<<dim3((((maxW + 7) / 8) * ((maxH + 7) / 8) + 63) / 64, nf), 64, 0, s>>>(E, d_efs.p);
    k_enc_number<<<dim3((max_dcg + 31) / 32, nf), 32, 0, s>>>(E, d_efs.p);
    k_enc_dc<<<dim3((maxW * maxH + 255) / 256, nf), 256, 0, s>>>(E, d_efs.p);
    k_enc_coeffs<0><<<dim3(max_groups, nf), kEncThreads, kEncSmemFloats * sizeof(float), s>>>(E, d_efs.p);
    k_enc_cfl<<<dim3(((maxW + 7) / 8) * ((maxH + 7) / 8), nf), 128, 4 * 4096 * sizeof(float), s>>>(E, d_efs.p);
    if (p.adaptive_quant) k_enc_adjust<<<dim3((maxW * maxH + 127) / 128, 3, nf), 128, 0, s>>>(E, d_efs.p);
    k_enc_coeffs<1><<<dim3(max_groups, nf), kEncThreads, kEncSmemFloats * sizeof(float), s>>>(E, d_efs.p);
    // ---- coefficient orders: zero counts on the device, sort + permutation coding on the host, orders back
    std::vector<CustomOrders> orders(n);
    DevBuf<uint16_t>& d_custom = enc->d_custom;
    DevBuf<uint8_t>& d_sample = enc->d_sample;
    std::vector<uint16_t> custom_pool;
    std::vector<uint8_t> sample_bits;
    if (p.coeff_orders) {
      sample_bits = MakeOrderSampleBits(static_cast<size_t>(maxW) * maxH);
      CUDA_OK(d_sample.Upload(sample_bits, s));
      E.sample_bits = d_sample.p;
      k_enc_group_orders<<<dim3((max_groups + 31) / 32, nf), 32, 0, s>>>(E, d_efs.p);
      k_enc_order_stats<<<dim3(max_groups, nf), 256, 0, s>>>(E, d_efs.p);
      std::vector<std::vector<int32_t>> h_stats(n);
      for (size_t i = 0; i < n; i++) {
        const uint64_t first = fr[i].ef.order_mask, count = fr[i].ef.zero_counts + kCustomOrderCounters - first;
        h_stats[i].resize(count);
        CUDA_OK(cudaMemcpyAsync(h_stats[i].data(), d_iarena.p + first, count * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
      }
      CUDA_OK(cudaStreamSynchronize(s));
      {
        std::atomic<size_t> next{0};
        auto work = [&]() {
          for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n) break;
            const DevEFrame& e = fr[i].ef;
            orders[i] = ComputeCustomOrders(static_cast<uint32_t>(h_stats[i][0]), h_stats[i].data() + (e.zero_counts - e.order_mask),
                                            e.xblocks, e.yblocks);
          }
        };
        const size_t nthreads = std::max<size_t>(1, std::min<size_t>(n, std::min<unsigned>(16, std::thread::hardware_concurrency())));
        std::vector<std::thread> pool;
        for (size_t t = 0; t < nthreads; t++) pool.emplace_back(work);
        for (auto& t : pool) t.join();
      }
      for (size_t i = 0; i < n; i++)
        for (uint32_t ord = 0; ord < kNumCustomOrders; ord++) {
          if (!(orders[i].used & (1u << ord))) continue;
          for (uint32_t c = 0; c < 3; c++) {
            JXLB_CHECK(custom_pool.size() < 0xFFFFFFFFu, "custom order pool too large");
            efs[i].custom_order[3 * ord + c] = fr[i].ef.custom_order[3 * ord + c] = static_cast<uint32_t>(custom_pool.size());
            custom_pool.insert(custom_pool.end(), orders[i].order[ord][c].begin(), orders[i].order[ord][c].end());
          }
        }
      CUDA_OK(d_custom.Upload(custom_pool, s));
      E.opool_custom = d_custom.p;
      CUDA_OK(d_efs.Upload(efs, s));
    }
    k_enc_block_stats<<<dim3((maxW * maxH + 127) / 128, 3, nf), 128, 0, s>>>(E, d_efs.p);
    k_enc_token_offsets<<<dim3((max_groups + 31) / 32, nf), 32, 0, s>>>(E, d_efs.p);
    k_enc_block_tokens<<<dim3((maxW * maxH + 127) / 128, 3, nf), 128, 0, s>>>(E, d_efs.p);
    k_enc_modular<<<dim3(256, max_dcg, nf), 256, 0, s>>>(E, d_efs.p);
    if (p.alpha) k_enc_alpha<<<dim3(148 * 8, nf), 256, 0, s>>>(E, d_efs.p);
    CUDA_OK(cudaEventRecord(ev[1], s));
    // ---- host: histograms -> codes, global sections, section layout
    uint64_t words_total = 0, nsec = 0;
    std::vector<uint16_t> h_rev;
    std::vector<uint32_t> h_fs;
    struct CodeOff { uint64_t mod_fs, mod_r, ac_fs, ac_r; };
    std::vector<CodeOff> code_off(n);
    std::vector<std::vector<int32_t>> h_small(n);
    for (size_t i = 0; i < n; i++) {  // counts + histograms of every frame: one copy each, one synchronisation
      const Frame& f = fr[i];
      const uint64_t first = f.ef.dcg_count, count = f.ef.mod_hist + static_cast<uint64_t>(f.L.num_leaves) * 256 - first;
      h_small[i].resize(count);
      CUDA_OK(cudaMemcpyAsync(h_small[i].data(), d_iarena.p + first, count * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    // an image that fits one group carries its alpha samples in the global section, which the host writes: their tokens
    std::vector<std::vector<uint2>> h_alpha(n);
    for (size_t i = 0; i < n && p.alpha; i++) {
      if (xsizes[i] > 256 || ysizes[i] > 256) continue;
      h_alpha[i].resize(static_cast<size_t>(xsizes[i]) * ysizes[i]);
      CUDA_OK(cudaMemcpyAsync(h_alpha[i].data(), d_tokens.p + fr[i].ef.alpha_tokens, h_alpha[i].size() * sizeof(uint2),
                              cudaMemcpyDeviceToHost, s));
    }
    CUDA_OK(cudaStreamSynchronize(s));
    {  // histogram normalisation, header coding and table building per frame, on host threads
      std::atomic<size_t> next{0};
      std::vector<std::string> errors(n);
      auto work = [&]() {
        for (;;) {
          const size_t i = next.fetch_add(1);
          if (i >= n) break;
          try {
            Frame& f = fr[i];
            const uint64_t first = f.ef.dcg_count;
            const uint32_t* ac_hist = reinterpret_cast<const uint32_t*>(h_small[i].data() + (f.ef.ac_hist - first));
            const uint32_t* mod_hist = reinterpret_cast<const uint32_t*>(h_small[i].data() + (f.ef.mod_hist - first));
            std::vector<std::pair<uint32_t, uint32_t>> global_alpha;
            for (const uint2& t : h_alpha[i]) global_alpha.push_back({t.x, t.y});
            BuildEncGlobals(p, f.L, f.tree, ac_cluster_of, f.global_scale, f.quant_dc, mod_hist, ac_hist, orders[i], &f.G,
                            h_alpha[i].empty() ? nullptr : &global_alpha);
          } catch (const std::exception& e) {
            errors[i] = e.what();
          }
        }
      };
      const size_t nthreads = std::max<size_t>(1, std::min<size_t>(n, std::min<unsigned>(16, std::thread::hardware_concurrency())));
      std::vector<std::thread> pool;
      for (size_t t = 0; t < nthreads; t++) pool.emplace_back(work);
      for (auto& t : pool) t.join();
      for (const std::string& e : errors)
        if (!e.empty()) throw Error(e);
    }
    for (size_t i = 0; i < n; i++) {
      Frame& f = fr[i];
      const FrameDimensions& d = f.L.dim;
      const uint64_t first = f.ef.dcg_count;
      const int32_t* dcg_count = h_small[i].data();
      const int32_t* group_tokens = h_small[i].data() + (f.ef.group_tokens - first);
      f.bits_off = nsec;
      for (uint32_t g = 0; g < d.num_dc_groups; g++) {
        const uint32_t gx = g % d.xsize_dc_groups, gy = g / d.xsize_dc_groups;
        const uint64_t xs = std::min<uint64_t>(256, d.xsize_blocks - gx * 256), ys = std::min<uint64_t>(256, d.ysize_blocks - gy * 256);
        f.dc_off.push_back(words_total);  // the DC half of the section (the metadata half: below)
        words_total += (3 * xs * ys * 6 + 64) / 4 + 4;
      }
      for (uint32_t g = 0; g < d.num_groups; g++) {
        f.ac_off.push_back(words_total);
        words_total += (static_cast<uint64_t>(group_tokens[g]) * 6 + 64) / 4 + 4;
        if (p.alpha) words_total += (65536 * 6 + 64) / 4 + 4;  // the group's alpha stream
      }
      for (uint32_t g = 0; g < d.num_dc_groups; g++) {
        const uint32_t gx = g % d.xsize_dc_groups, gy = g / d.xsize_dc_groups;
        const uint64_t xs = std::min<uint64_t>(256, d.xsize_blocks - gx * 256), ys = std::min<uint64_t>(256, d.ysize_blocks - gy * 256);
        const uint64_t toks = 2 * ((xs + 7) / 8) * ((ys + 7) / 8) + 2 * static_cast<uint64_t>(dcg_count[g]) + xs * ys;
        f.meta_off.push_back(words_total);
        words_total += (toks * 6 + 64 + 32) / 4 + 4;
      }
      nsec += 2 * d.num_dc_groups + d.num_groups;
      auto push_rev = [&](const std::vector<uint16_t>& v) {
        const uint64_t off = h_rev.size();
        h_rev.insert(h_rev.end(), v.begin(), v.end());
        return off;
      };
      auto push_fs = [&](const std::vector<uint32_t>& v) {
        const uint64_t off = h_fs.size();
        h_fs.insert(h_fs.end(), v.begin(), v.end());
        return off;
      };
      code_off[i] = CodeOff{push_fs(f.G.mod_code.Fs()), push_rev(f.G.mod_code.reverse), push_fs(f.G.ac_code.Fs()),
                            push_rev(f.G.ac_code.reverse)};
    }
    DevBuf<uint16_t>& d_rev = enc->d_rev;
    DevBuf<uint32_t>&d_words = enc->d_words, &d_fs = enc->d_fs;
    DevBuf<uint64_t>&d_off = enc->d_off, &d_bits = enc->d_bits;
    std::vector<uint64_t> h_off;
    for (size_t i = 0; i < n; i++) {
      h_off.insert(h_off.end(), fr[i].dc_off.begin(), fr[i].dc_off.end());
      h_off.insert(h_off.end(), fr[i].ac_off.begin(), fr[i].ac_off.end());
      h_off.insert(h_off.end(), fr[i].meta_off.begin(), fr[i].meta_off.end());
    }
    h_off.push_back(words_total);  // region of section k: words [h_off[k], h_off[k + 1])
    CUDA_OK(d_rev.Upload(h_rev, s));
    CUDA_OK(d_fs.Upload(h_fs, s));
    CUDA_OK(d_off.Upload(h_off, s));
    CUDA_OK(d_bits.Alloc(nsec + 1));
    CUDA_OK(d_words.Alloc(words_total + 16));
    CUDA_OK(cudaMemsetAsync(d_words.p, 0, (words_total + 16) * sizeof(uint32_t), s));
    CUDA_OK(cudaEventRecord(ev[2], s));
    // ---- phase 2: rANS emission of every section of every frame
    for (size_t i = 0; i < n; i++) {
      const CodeOff& co = code_off[i];
      const uint64_t offs[4] = {co.mod_fs, co.mod_r, co.ac_fs, co.ac_r};
      for (int k = 0; k < 4; k++) efs[i].code_off[k] = offs[k];
      efs[i].sec_base = static_cast<uint32_t>(fr[i].bits_off);
    }
    CUDA_OK(d_efs.Upload(efs, s));
    // which reverse tables the DC halves keep in shared memory: per frame and half the contexts of that half's subtree
    // (the root of the global tree splits on the stream id: left = AC metadata, right = DC), most used first
    std::vector<uint8_t> h_slots(n * 2 * 64, 0xFF);
    for (size_t i = 0; i < n; i++) {
      const Frame& f = fr[i];
      const std::vector<DevEncTreeNode>& nodes = f.tree.nodes;
      if (nodes.empty() || nodes[0].prop != 1) continue;
      const uint32_t* mod_hist = reinterpret_cast<const uint32_t*>(h_small[i].data() + (f.ef.mod_hist - f.ef.dcg_count));
      for (uint32_t half = 0; half < 2; half++) {
        std::vector<std::pair<uint64_t, uint32_t>> leaves;  // (tokens, leaf)
        std::vector<uint32_t> stack = {half == 0 ? nodes[0].r : nodes[0].l};
        while (!stack.empty()) {
          const uint32_t k = stack.back();
          stack.pop_back();
          if (k >= nodes.size()) continue;
          if (nodes[k].prop < 0) {
            uint64_t count = 0;
            for (uint32_t t = 0; t < 256; t++) count += mod_hist[static_cast<size_t>(nodes[k].l) * 256 + t];
            if (count != 0 && nodes[k].l < 64) leaves.push_back({count, nodes[k].l});
          } else {
            stack.push_back(nodes[k].l);
            stack.push_back(nodes[k].r);
          }
        }
        std::sort(leaves.begin(), leaves.end(), [](const auto& a, const auto& b) { return a.first > b.first || (a.first == b.first && a.second < b.second); });
        for (uint32_t k = 0; k < leaves.size() && k < kEmitDcSlots; k++) h_slots[(i * 2 + half) * 64 + leaves[k].second] = static_cast<uint8_t>(k);
      }
    }
    CUDA_OK(enc->d_slots.Upload(h_slots, s));
    const uint32_t dc_blocks = max_dcg;
    if (!enc->stream2) CUDA_OK(cudaStreamCreateWithFlags(&enc->stream2, cudaStreamNonBlocking));
    cudaEvent_t fork = nullptr, join = nullptr;
    CUDA_OK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
    CUDA_OK(cudaEventRecord(fork, s));
    CUDA_OK(cudaStreamWaitEvent(enc->stream2, fork, 0));
    k_enc_emit_dc<<<dim3(2 * dc_blocks, nf), 32, kEmitDcSlots * 8192, enc->stream2>>>(E, d_efs.p, d_fs.p, d_rev.p, d_words.p, d_off.p,
                                                                                     d_bits.p, dc_blocks, enc->d_slots.p);
    CUDA_OK(cudaEventRecord(join, enc->stream2));
    k_enc_emit<<<dim3(max_groups, nf), 32, 0, s>>>(E, d_efs.p, d_fs.p, d_rev.p, d_words.p, d_off.p, d_bits.p);
    CUDA_OK(cudaStreamWaitEvent(s, join, 0));
    CUDA_OK(cudaEventRecord(ev[3], s));
    CUDA_OK(cudaEventDestroy(fork));
    CUDA_OK(cudaEventDestroy(join));
    std::vector<uint64_t> h_bits, h_nbits;
    if (FetchSections(enc, s, h_off, nsec, &h_bits, &h_nbits) != 0) return 1;
    uint32_t* const h_words = enc->h_words;
    float ms = 0;
    cudaEventElapsedTime(&ms, ev[0], ev[1]);
    enc->phase_ms[0] = ms;
    cudaEventElapsedTime(&ms, ev[1], ev[2]);
    enc->phase_ms[1] = ms;
    cudaEventElapsedTime(&ms, ev[2], ev[3]);
    enc->phase_ms[2] = ms;
    for (auto& e : ev) cudaEventDestroy(e);
    // ---- assemble (host threads)
    enc->outputs.assign(n, std::vector<uint8_t>());
    {
      std::atomic<size_t> next{0};
      auto work = [&]() {
        for (;;) {
          const size_t i = next.fetch_add(1);
          if (i >= n) break;
          const Frame& f = fr[i];
          const FrameDimensions& d = f.L.dim;
          std::vector<EncSection> dcg, acg;  // h_bits[sec] = first bit of the section, its end = end of its region
          for (uint32_t g = 0; g < d.num_dc_groups; g++) {
            const uint64_t sec = f.bits_off + g;
            const uint64_t sec2 = sec + d.num_dc_groups + d.num_groups;  // the metadata half
            dcg.push_back({h_words, h_bits[sec], h_nbits[sec], h_bits[sec2], h_nbits[sec2]});
          }
          for (uint32_t g = 0; g < d.num_groups; g++) {
            const uint64_t sec = f.bits_off + d.num_dc_groups + g;
            acg.push_back({h_words, h_bits[sec], h_nbits[sec]});
          }
          enc->outputs[i] = AssembleCodestream(p, f.L, f.G, dcg, acg);
        }
      };
      const size_t nthreads = std::max<size_t>(1, std::min<size_t>(n, std::min<unsigned>(16, std::thread::hardware_concurrency())));
      std::vector<std::thread> pool;
      for (size_t t = 0; t < nthreads; t++) pool.emplace_back(work);
      for (auto& t : pool) t.join();
    }
  } catch (const std::exception& e) {
    enc->error = e.what();
    return 1;
  }
  return 0;
}

// Lossless batch: pixels[i] = xsizes[i] * ysizes[i] interleaved samples, num_channels each (1 grey, 2 grey + alpha, 3 RGB,
// 4 RGBA), bits_per_sample 8 (uint8) or 16 (uint16, native endian). kernels/jxlb_encl_dev.h describes the stream.
int JxlB200EncoderEncodeLosslessBatch(JxlB200Encoder* enc, const void* const* pixels, const uint32_t* xsizes, const uint32_t* ysizes,
                                      size_t n, uint32_t num_channels, uint32_t bits_per_sample) {
  if (!enc || !pixels || !xsizes || !ysizes || n == 0) return 1;
  enc->error.clear();
  enc->outputs.clear();
  if (num_channels < 1 || num_channels > 4 || !(bits_per_sample == 8 || bits_per_sample == 16)) {
    enc->error = "invalid lossless encode options";
    return 1;
  }
  CUDA_OK(cudaSetDevice(enc->device));
  cudaStream_t s = enc->stream;
  try {
    const EnclTree tree = BuildEnclTree();
    std::vector<EnclParams> ps(n);
    std::vector<DevLFrame> lf(n);
    std::vector<uint64_t> h_off;  // word offset of every group section, then the end
    uint64_t in_bytes = 0, planes = 0, toks = 0, words_total = 0;
    uint32_t max_groups = 0;
    uint64_t max_samples = 0;
    const uint32_t bytes = bits_per_sample / 8;
    for (size_t i = 0; i < n; i++) {
      JXLB_CHECK(xsizes[i] > 0 && ysizes[i] > 0 && xsizes[i] <= (1u << 16) && ysizes[i] <= (1u << 16) && pixels[i], "bad image");
      EnclParams& p = ps[i];
      p.xsize = xsizes[i];
      p.ysize = ysizes[i];
      p.nch = num_channels;
      p.bits = bits_per_sample;
      DevLFrame& f = lf[i];
      f = DevLFrame{};
      f.xsize = p.xsize;
      f.ysize = p.ysize;
      f.xgroups = p.XGroups();
      f.ygroups = p.YGroups();
      f.nch = p.nch;
      f.bytes = bytes;
      const uint64_t px = static_cast<uint64_t>(p.xsize) * p.ysize;
      f.in_off = in_bytes;
      in_bytes += (px * p.nch * bytes + 15) & ~uint64_t{15};
      f.plane_off = planes;
      planes += px * p.nch;
      f.tok_off = toks;
      const uint32_t groups = f.xgroups * f.ygroups;
      toks += static_cast<uint64_t>(groups) * p.nch * kEnclGroupSamples;
      f.hist_off = i * 34 * 256;
      f.sec_base = static_cast<uint32_t>(h_off.size());
      for (uint32_t g = 0; g < groups; g++) {
        const uint32_t gx = g % f.xgroups, gy = g / f.xgroups;
        const uint64_t gw = std::min<uint32_t>(kEnclGroupDim, p.xsize - gx * kEnclGroupDim);
        const uint64_t gh = std::min<uint32_t>(kEnclGroupDim, p.ysize - gy * kEnclGroupDim);
        h_off.push_back(words_total);
        words_total += EnclSectionWords(gw * gh * p.nch);
      }
      max_groups = std::max(max_groups, groups);
      max_samples = std::max(max_samples, px * p.nch);
    }
    h_off.push_back(words_total);
    const size_t nsec = h_off.size() - 1;
    JXLB_CHECK(nsec < (size_t{1} << 31), "too many sections");
    CUDA_OK(enc->d_in.Alloc(in_bytes + 16));
    CUDA_OK(enc->d_iarena.Alloc(planes + n * 34 * 256 + 16));
    CUDA_OK(enc->d_tokens.Alloc(toks + 16));
    std::vector<int32_t> consts(kEnclCutoffValues, kEnclCutoffValues + 33);
    for (int k = 0; k < 34; k++) consts.push_back(static_cast<int32_t>(tree.leaf_of[k]));
    CUDA_OK(enc->d_lconst.Upload(consts, s));
    {
      std::vector<uint64_t> dev_off(n);
      std::vector<size_t> in_sizes(n);
      for (size_t i = 0; i < n; i++) {
        dev_off[i] = lf[i].in_off;
        in_sizes[i] = static_cast<size_t>(xsizes[i]) * ysizes[i] * num_channels * bytes;
      }
      if (UploadFrames(enc, s, pixels, dev_off, in_sizes) != 0) return 1;
    }
    uint32_t* d_hist = reinterpret_cast<uint32_t*>(enc->d_iarena.p + planes);
    CUDA_OK(cudaMemsetAsync(d_hist, 0, n * 34 * 256 * sizeof(uint32_t), s));
    CUDA_OK(enc->d_lfs.Upload(lf, s));
    DevLPools L{};
    L.in = enc->d_in.p;
    L.planes = enc->d_iarena.p;
    L.tokens = enc->d_tokens.p;
    L.hist = d_hist;
    L.cutoffs = enc->d_lconst.p;
    L.leaf_of = reinterpret_cast<const uint32_t*>(enc->d_lconst.p + 33);
    cudaEvent_t ev[4];
    for (auto& e : ev) CUDA_OK(cudaEventCreate(&e));
    CUDA_OK(cudaEventRecord(ev[0], s));
    const uint32_t nf = static_cast<uint32_t>(n);
    const uint32_t px_blocks = static_cast<uint32_t>(std::min<uint64_t>(148 * 16, (max_samples + 255) / 256));
    k_encl_planes<<<dim3(px_blocks, nf), 256, 0, s>>>(L, enc->d_lfs.p);
    k_encl_tokens<<<dim3(px_blocks, nf), 256, 0, s>>>(L, enc->d_lfs.p);
    CUDA_OK(cudaEventRecord(ev[1], s));
    std::vector<uint32_t> hist(n * 34 * 256);
    CUDA_OK(cudaMemcpyAsync(hist.data(), d_hist, hist.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    // ---- host: the global section of every frame and the encoder tables of its code
    std::vector<BitWriter> globals(n);
    std::vector<uint32_t> all_fs;
    std::vector<uint16_t> all_rev;
    for (size_t i = 0; i < n; i++) {
      EncCode code;
      WriteEnclGlobal(globals[i], ps[i], tree, hist.data() + i * 34 * 256, &code);
      lf[i].code_off[0] = all_fs.size();
      lf[i].code_off[1] = all_rev.size();
      const std::vector<uint32_t> fsv = code.Fs();
      all_fs.insert(all_fs.end(), fsv.begin(), fsv.end());
      all_rev.insert(all_rev.end(), code.reverse.begin(), code.reverse.end());
    }
    CUDA_OK(enc->d_fs.Upload(all_fs, s));
    CUDA_OK(enc->d_rev.Upload(all_rev, s));
    CUDA_OK(enc->d_lfs.Upload(lf, s));
    CUDA_OK(enc->d_off.Upload(h_off, s));
    CUDA_OK(enc->d_bits.Alloc(nsec + 1));
    CUDA_OK(enc->d_words.Alloc(words_total + 16));
    CUDA_OK(cudaMemsetAsync(enc->d_words.p, 0, (words_total + 16) * sizeof(uint32_t), s));
    CUDA_OK(cudaEventRecord(ev[2], s));
    k_encl_emit<<<dim3(max_groups, nf), 32, 0, s>>>(L, enc->d_lfs.p, enc->d_fs.p, enc->d_rev.p, enc->d_words.p,
                                                                enc->d_off.p, enc->d_bits.p);
    CUDA_OK(cudaEventRecord(ev[3], s));
    std::vector<uint64_t> h_bits, h_nbits;
    if (FetchSections(enc, s, h_off, nsec, &h_bits, &h_nbits) != 0) return 1;
    float ms = 0;
    for (int k = 0; k < 3; k++) {
      cudaEventElapsedTime(&ms, ev[k], ev[k + 1]);
      enc->phase_ms[k] = ms;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    enc->outputs.assign(n, std::vector<uint8_t>());
    {  // assemble (host threads: a bit-wise copy of every section into its codestream)
      std::atomic<size_t> next{0};
      std::vector<std::string> errors(n);
      auto work = [&]() {
        for (;;) {
          const size_t i = next.fetch_add(1);
          if (i >= n) break;
          try {
            std::vector<EncSection> groups;
            for (uint32_t g = 0; g < lf[i].xgroups * lf[i].ygroups; g++) {
              const uint64_t sec = lf[i].sec_base + g;
              groups.push_back({enc->h_words, h_bits[sec], h_nbits[sec]});
            }
            enc->outputs[i] = AssembleEncl(ps[i], globals[i], groups);
          } catch (const std::exception& e) {
            errors[i] = e.what();
          }
        }
      };
      const size_t nthreads = std::max<size_t>(1, std::min<size_t>(n, std::min<unsigned>(16, std::thread::hardware_concurrency())));
      std::vector<std::thread> pool;
      for (size_t t = 0; t < nthreads; t++) pool.emplace_back(work);
      for (auto& t : pool) t.join();
      for (const std::string& e : errors)
        if (!e.empty()) throw Error(e);
    }
  } catch (const std::exception& e) {
    enc->error = e.what();
    return 1;
  }
  return 0;
}

size_t JxlB200EncoderOutputSize(const JxlB200Encoder* enc, size_t i) {
  return enc && i < enc->outputs.size() ? enc->outputs[i].size() : 0;
}

int JxlB200EncoderReadOutput(const JxlB200Encoder* enc, size_t i, uint8_t* dst, size_t size) {
  if (!enc || i >= enc->outputs.size() || !dst || size < enc->outputs[i].size()) return 1;
  std::memcpy(dst, enc->outputs[i].data(), enc->outputs[i].size());
  return 0;
}

int JxlB200EncoderGetPhaseTimes(const JxlB200Encoder* enc, double* ms3) {
  if (!enc || !ms3) return 1;
  for (int k = 0; k < 3; k++) ms3[k] = enc->phase_ms[k];
  return 0;
}

}  // extern "C"

#include "jxl_encode_api.inc"
