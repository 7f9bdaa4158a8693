// jxl_b200 device code of the VarDCT path (lossy frames).
//
// Kernel map (wrappers in jxl_b200.cu; the bodies below are __host__ __device__ so that
// tests/ can run the same statements on the CPU, the product never does):
//   DevDcGroupFinish   one CTA per DC group: quantised DC -> float DC (+ chroma-from-luma on DC),
//                      quant-DC bucket, AC-metadata rows -> strategy / quant / sharpness / sigma maps
//                        lib/jxl/compressed_dc.cc:197-290, lib/jxl/dec_modular.cc:437-532, lib/jxl/epf.cc:39-147
//   DevDcSmoothPixel   adaptive DC smoothing, lib/jxl/compressed_dc.cc:124-195
//   DevDecodeAcStream  one thread per (frame, group, pass): block contexts, non-zero counts, the
//                      coefficient symbol loop; emits sparse tokens
//                        lib/jxl/dec_group.cc:454-527, :534-645, lib/jxl/ac_context.h, lib/jxl/dec_ans.h:168-255
//   DevVarblock        one warp / CTA per varblock: token scatter, dequantisation + quant bias + chroma
//                      from luma, lowest frequencies from DC, the 27 inverse transforms
//                        lib/jxl/dec_group.cc:98-166, lib/jxl/quantizer-inl.h:34-71,
//                        lib/jxl/dec_transforms-inl.h:30-813, lib/jxl/dct-inl.h:50-335
//   DevGaborishPixel / DevEpfPixel / DevColorPixel   per-pixel render stages
//                        lib/jxl/render_pipeline/stage_{gaborish,epf,xyb,ycbcr,from_linear,write}.cc
//
// Floating point: every multiply-add that libjxl writes as MulAdd / NegMulAdd is an explicit fmaf;
// everything else is a separately rounded IEEE operation (the library is built with --fmad=false),
// so the results do not depend on how the work is spread over threads.
#ifndef JXLB_VARDCT_DEV_H_
#define JXLB_VARDCT_DEV_H_

#include <math.h>

#include "jxlb_finish_dev.h"
#include "jxlb_vardct_desc.h"
#include "jxlb_wc.h"

namespace jxlb {

struct DevVPools {
  const DevVFrame* frames;
  const DevAcStream* streams;
  uint32_t num_streams;
  const float* fpool;
  const uint16_t* opool;
  const uint8_t* cpool;
  const uint32_t* upool;
  float* farena;
  uint8_t* barena;
  uint32_t* uarena;
  uint32_t* tokens;
  uint32_t* ac_status;  // per AC stream
  uint32_t* ac_used;    // per AC stream: tokens produced (may exceed the capacity -> retry with more)
  uint32_t* dc_status;  // per (frame, DC group)
  uint32_t wc_off, llf_off, afv_off;  // fpool: WcMultipliers, DC -> LLF resample scales, AFV basis
  uint32_t sinfo_off;                 // upool: packed StrategyInfo x 27
  uint32_t ctxtab_off;                // upool: kCoeffFreqContext[64] then kCoeffNumNonzeroContext[64]
  uint8_t* out;
  const DevRefFrame* ref_frames;  // reference-only frames of the batch
  uint32_t num_ref_frames;
  const DevPatch* patches;
  uint32_t ac_plain_ans;  // every AC coefficient code of the batch is ANS without LZ77
  uint64_t* chain_pos;    // DevAcStream::chain_slot: the bit position behind the stream's last coefficient
  const int32_t* arena;   // the Modular sample arena and plane table (alpha of VarDCT frames with extra channels)
  const DevPlane* planes;
  const float* spl_seg;   // spline draw cache: segments, row offsets + index lists (DevSplineAdd)
  const uint32_t* spl_idx;
};

#if defined(__CUDACC__)
#define JXLB_UNROLL _Pragma("unroll")
#else
#define JXLB_UNROLL
#endif

template <int SCOPE>
JXLB_HD void CoopSync() {
#if defined(__CUDA_ARCH__)
  if (SCOPE == 1) __syncwarp();
  if (SCOPE == 2) __syncthreads();
#endif
}

#if defined(__CUDA_ARCH__)
#define JXLB_WARP_ALL(p) __all_sync(0xFFFFFFFFu, (p))
#else
#define JXLB_WARP_ALL(p) (p)
#endif

constexpr float kDevSqrt2 = 1.41421356237f;

// ---------------------------------------------------------------- DC groups
// `occ`: kDcOccWords words of scratch (shared memory on the device): one bit per block of the group, set when a varblock
// covers it. The serial scan below finds the next free block with a find-first-set per 32 blocks and marks / tests a
// varblock with one word per block row (a varblock never crosses a 32-block boundary); the strategy bytes go straight
// to the frame's map.
constexpr uint32_t kDcOccWords = 256 * 8;
JXLB_HD uint32_t DevFfs(uint32_t v) {  // 1-based index of the lowest set bit (v != 0)
#if defined(__CUDA_ARCH__)
  return static_cast<uint32_t>(__ffs(static_cast<int>(v)));
#else
  return static_cast<uint32_t>(__builtin_ffs(static_cast<int>(v)));
#endif
}
constexpr uint32_t kDcStageEntries = 2048;  // list entries staged per chunk (uint16 each) by DevDcGroupFinish

// `stage`: kDcStageEntries words of scratch (shared memory on the device) for the serial scan.
template <int SCOPE>
JXLB_HD void DevDcGroupFinish(const DevPools& P, const DevVPools& V, uint32_t frame, uint32_t g, uint32_t tid, uint32_t nt,
                              uint32_t status_index, uint32_t* occ, uint32_t* stage) {
  const DevVFrame& vf = V.frames[frame];
  const uint32_t W = vf.xblocks, H = vf.yblocks;
  const uint32_t gx = g % vf.xdcgroups, gy = g / vf.xdcgroups;
  const uint32_t x0 = gx * 256, y0 = gy * 256;
  const uint32_t xs = W - x0 < 256 ? W - x0 : 256, ys = H - y0 < 256 ? H - y0 : 256;
  const uint32_t* dcg = V.upool + vf.dcg_index + 8 * g;
  const int32_t* qy = P.arena + P.planes[dcg[0]].off;
  const int32_t* qx = P.arena + P.planes[dcg[1]].off;
  const int32_t* qb = P.arena + P.planes[dcg[2]].off;
  const int32_t* m_ytox = P.arena + P.planes[dcg[3]].off;
  const int32_t* m_ytob = P.arena + P.planes[dcg[4]].off;
  const DevPlane pl_rows = P.planes[dcg[5]];
  const int32_t* m_rows = P.arena + pl_rows.off;
  const int32_t* m_sharp = P.arena + P.planes[dcg[6]].off;
  const uint32_t extra_precision = dcg[7];
  const float mul = 1.0f / static_cast<float>(1u << extra_precision);
  const float fac_x = vf.mul_dc[0] * mul, fac_y = vf.mul_dc[1] * mul, fac_b = vf.mul_dc[2] * mul;
  float* dcx = V.farena + vf.dc[0];
  float* dcy = V.farena + vf.dc[1];
  float* dcb = V.farena + vf.dc[2];
  uint8_t* acs_frame = V.barena + vf.acs;
  // the map the scan works on: element (x, y) of the group at acs[y * astride + x]
  uint8_t* acs = acs_frame + static_cast<size_t>(y0) * W + x0;
  const uint32_t astride = W;
  uint8_t* qdc = V.barena + vf.qdc;
  uint8_t* sharp = V.barena + vf.sharp;
  uint16_t* rawq = reinterpret_cast<uint16_t*>(V.barena + vf.rawq);
  const uint32_t* thr = V.upool + vf.bctx_off;
  uint32_t status = 0;
  // (a) per block: dequantised DC, quant-DC bucket, sharpness; strategy map cleared
  for (uint32_t i = tid; i < xs * ys; i += nt) {
    const uint32_t x = i % xs, y = i / xs;
    const size_t pos = static_cast<size_t>(y0 + y) * W + x0 + x;
    const int32_t vx = qx[i], vy = qy[i], vb = qb[i];
    const float in_x = static_cast<float>(vx) * fac_x;
    const float in_y = static_cast<float>(vy) * fac_y;
    const float in_b = static_cast<float>(vb) * fac_b;
    dcy[pos] = in_y;
    dcx[pos] = fmaf(in_y, vf.cfl_dc_x, in_x);
    dcb[pos] = fmaf(in_y, vf.cfl_dc_b, in_b);
    uint32_t bucket = 0;
    if (vf.num_dc_ctxs > 1) {
      const uint32_t n0 = vf.num_dc_thr[0], n1 = vf.num_dc_thr[1], n2 = vf.num_dc_thr[2];
      uint32_t bx = 0, by = 0, bb = 0;
      for (uint32_t t = 0; t < n0; t++) bx += vx > static_cast<int32_t>(thr[t]);
      for (uint32_t t = 0; t < n1; t++) by += vy > static_cast<int32_t>(thr[n0 + t]);
      for (uint32_t t = 0; t < n2; t++) bb += vb > static_cast<int32_t>(thr[n0 + n1 + t]);
      bucket = (bx * (n2 + 1) + bb) * (n1 + 1) + by;
    }
    qdc[pos] = static_cast<uint8_t>(bucket);
    const int32_t sh = m_sharp[i];
    if (sh < 0 || sh >= 8) status |= kVBadStream;
    sharp[pos] = static_cast<uint8_t>(sh & 7);
    acs[static_cast<size_t>(y) * astride + x] = 0xFF;
  }
  for (uint32_t i = tid; i < ys * 8; i += nt) {  // blocks past the group's width count as covered
    const uint32_t first = (i & 7) * 32;
    occ[i] = first >= xs ? 0xFFFFFFFFu : (xs - first >= 32 ? 0u : 0xFFFFFFFFu << (xs - first));
  }
  // colour correlation maps (one entry per 64x64 tile)
  const uint32_t cw = (xs + 7) >> 3, chh = (ys + 7) >> 3, cx0 = x0 >> 3, cy0 = y0 >> 3;
  int8_t* ytox = reinterpret_cast<int8_t*>(V.barena + vf.ytox);
  int8_t* ytob = reinterpret_cast<int8_t*>(V.barena + vf.ytob);
  for (uint32_t i = tid; i < cw * chh; i += nt) {
    const uint32_t x = i % cw, y = i / cw;
    int32_t a = m_ytox[i], b = m_ytob[i];
    a = a < -128 ? -128 : (a > 127 ? 127 : a);
    b = b < -128 ? -128 : (b > 127 ? 127 : b);
    ytox[(cy0 + y) * vf.cmw + cx0 + x] = static_cast<int8_t>(a);
    ytob[(cy0 + y) * vf.cmw + cx0 + x] = static_cast<int8_t>(b);
  }
  CoopSync<SCOPE>();
  // (b) the strategy / quant rows list one entry per varblock in raster order of their top-left
  // blocks; where the next varblock starts depends on the extents of all earlier ones: serial. The list is
  // staged chunk by chunk by all threads as packed words -- strategy | quant << 8 | cx << 16 | cy << 24 -- so that
  // the scanning thread waits for one shared-memory word per varblock, requested one varblock ahead; it keeps the
  // occupancy word it is filling in a register, marks the rows below in shared memory and stores only the varblock's
  // first cell (strategy byte, raw quant). The cells a varblock covers are filled in parallel in (c).
  {
    const uint32_t cap = pl_rows.w;
    const uint32_t count = static_cast<uint32_t>(m_rows[static_cast<size_t>(cap) * 2]);
    const int32_t* row_strategy = m_rows;
    const int32_t* row_quant = m_rows + cap;
    uint32_t num = 0, iy = 0, wi = 0;  // (thread 0's scan position, kept across chunks)
    for (uint32_t chunk = 0; chunk == 0 || chunk < count; chunk += kDcStageEntries) {
      const uint32_t chunk_end = chunk + kDcStageEntries < count ? chunk + kDcStageEntries : count;
      for (uint32_t i = chunk + tid; i < chunk_end; i += nt) {
        const int32_t raw = row_strategy[i];
        int32_t q = row_quant[i];
        q = q < 0 ? 0 : (q > 255 ? 255 : q);
        uint32_t e = 0xFFu;  // invalid strategy
        if (static_cast<uint32_t>(raw) < kNumStrategies) {
          const StrategyInfo si = UnpackStrategyInfo(V.upool[V.sinfo_off + raw]);
          e = static_cast<uint32_t>(raw) | (static_cast<uint32_t>(q) << 8) | (static_cast<uint32_t>(si.cx) << 16) |
              (static_cast<uint32_t>(si.cy) << 24);
        }
        stage[i - chunk] = e;
      }
      CoopSync<SCOPE>();
      if (tid == 0 && !(status & kVBadStream)) {
        const bool last_chunk = chunk_end == count;
        bool paused = false;
        uint32_t e_next = num < chunk_end ? stage[num - chunk] : 0xFFu;
        for (; iy < ys && !paused && !(status & kVBadStream); iy++, wi = 0) {
          uint32_t* orow = occ + iy * 8;
          const uint32_t rows_left_in_group = 32 - (iy & 31), rows_left = ys - iy;
          for (; wi < 8; wi++) {
            uint32_t w = orow[wi];
            while (~w != 0) {
              if (num >= chunk_end) {
                if (last_chunk) status |= kVBadStream;  // more varblocks than list entries
                paused = true;
                break;
              }
              const uint32_t e = e_next;
              e_next = num + 1 < chunk_end ? stage[num + 1 - chunk] : 0xFFu;  // (requested one varblock ahead)
              const uint32_t b = DevFfs(~w) - 1, ix = (wi << 5) + b;
              const uint32_t raw = e & 0xFF, cx = (e >> 16) & 0xFF, cy = e >> 24;
              // a varblock stays inside its 256x256 group (32 blocks) and inside the DC group
              if (raw == 0xFF || b + cx > 32 || ix + cx > xs || cy > rows_left_in_group || cy > rows_left) {
                status |= kVBadStream;
                break;
              }
              const uint32_t mask = (cx >= 32 ? 0xFFFFFFFFu : ((1u << cx) - 1u)) << b;
              uint32_t overlap = w & mask;
              w |= mask;
              for (uint32_t jy = 1; jy < cy; jy++) {
                uint32_t& o = orow[jy * 8 + wi];
                overlap |= o & mask;
                o |= mask;
              }
              if (overlap) {
                status |= kVBadStream;
                break;
              }
              acs[static_cast<size_t>(iy) * astride + ix] = static_cast<uint8_t>((raw << 1) | 1);
              rawq[static_cast<size_t>(y0 + iy) * W + x0 + ix] = static_cast<uint16_t>(1 + ((e >> 8) & 0xFF));
              num++;
            }
            orow[wi] = w;
            if (paused || (status & kVBadStream)) break;  // (before wi moves on: the scan resumes inside this word)
          }
          if (paused || (status & kVBadStream)) break;  // (resume at the same row and word with the next chunk)
        }
      }
      CoopSync<SCOPE>();
    }
  }
  // (c) every varblock fills the cells it covers: strategy bytes (first-block flag clear) and, with EPF, the inverse
  // sigma per block (ComputeSigma)
  {
    float* inv_sigma = V.farena + vf.inv_sigma;
    const float kInvSigmaNum = -1.1715728752538099024f;
    const bool epf = vf.epf_iters > 0;
    for (uint32_t i = tid; i < xs * ys; i += nt) {
      const uint32_t x = i % xs, y = i / xs;
      const size_t pos = static_cast<size_t>(y0 + y) * W + x0 + x;
      const uint8_t a = acs[static_cast<size_t>(y) * astride + x];
      if (a == 0xFF || !(a & 1)) continue;  // (cells of other varblocks: 0xFF until their owner gets here, or filled)
      const StrategyInfo si = UnpackStrategyInfo(V.upool[V.sinfo_off + (a >> 1)]);
      const float sigma_quant = epf ? vf.epf_quant_mul / (vf.global_scale_f * static_cast<float>(rawq[pos]) * kInvSigmaNum) : 0.0f;
      for (uint32_t jy = 0; jy < si.cy; jy++)
        for (uint32_t jx = 0; jx < si.cx; jx++) {
          if (jy | jx) acs[static_cast<size_t>(y + jy) * astride + x + jx] = static_cast<uint8_t>(a & 0xFE);
          if (epf) {
            const size_t q = pos + static_cast<size_t>(jy) * W + jx;
            float sigma = sigma_quant * vf.epf_sharp_lut[sharp[q]];
            sigma = sigma < -1e-4f ? sigma : -1e-4f;
            inv_sigma[q] = 1.0f / sigma;
          }
        }
    }
  }
  if (status && V.dc_status) V.dc_status[status_index] = status;  // pre-zeroed by the host
}

// One block of AdaptiveDCSmoothing: dc -> dc_final (only called when smoothing is on).
JXLB_HD void DevDcSmoothBlock(const DevVPools& V, const DevVFrame& vf, uint32_t x, uint32_t y) {
  const uint32_t W = vf.xblocks, H = vf.yblocks;
  const size_t pos = static_cast<size_t>(y) * W + x;
  if (x == 0 || y == 0 || x + 1 >= W || y + 1 >= H) {
    for (int c = 0; c < 3; c++) V.farena[vf.dc_final[c] + pos] = V.farena[vf.dc[c] + pos];
    return;
  }
  const float w1 = 0.20345139757231578f, w2 = 0.0334829185968739f;
  const float w0 = 1.0f - 4.0f * (w1 + w2);
  float mc[3], sm[3];
  float gap = 0.5f;
  for (int c = 0; c < 3; c++) {
    const float* rm = V.farena + vf.dc[c] + pos;
    const float* rt = rm - W;
    const float* rb = rm + W;
    mc[c] = rm[0];
    const float corner = (rt[-1] + rt[1]) + (rb[-1] + rb[1]);
    const float side = (rm[-1] + rm[1]) + (rt[0] + rb[0]);
    sm[c] = fmaf(corner, w2, fmaf(side, w1, mc[c] * w0));
    const float d = fabsf((mc[c] - sm[c]) / vf.mul_dc[c]);
    gap = gap > d ? gap : d;
  }
  float factor = fmaf(-4.0f, gap, 3.0f);
  if (factor < 0.0f) factor = 0.0f;
  for (int c = 0; c < 3; c++) V.farena[vf.dc_final[c] + pos] = fmaf(sm[c] - mc[c], factor, mc[c]);
}

// ---------------------------------------------------------------- AC coefficient streams
// The varblocks of group `g` in decode order (raster order of their first blocks, lib/jxl/dec_group.cc:168-442) with
// everything the AC stream needs to start one: word 0 = cell (by * 32 + bx inside the group) | cx << 10 |
// log2(covered blocks) << 16 | coefficient-order id << 20 | strategy << 24; word 1 = the block context
// (lib/jxl/ac_context.h:99-109) of channel c in byte c. One thread per group: a serial pass over <= 1024 strategy
// bytes, run once per batch between the DC kernels and AC decode, so that the lock-step AC lanes -- which diverge on
// every varblock start -- fetch one prefetchable record instead of walking the strategy map, the raw quant field,
// the quant-DC buckets and the block-context thresholds link by link.
JXLB_HD void DevBuildBlockList(const DevVPools& V, const DevVFrame& vf, uint32_t g) {
  const uint32_t W = vf.xblocks;
  const uint32_t x0 = (g % vf.xgroups) * 32, y0 = (g / vf.xgroups) * 32;
  const uint32_t xs = W - x0 < 32 ? W - x0 : 32, ys = vf.yblocks - y0 < 32 ? vf.yblocks - y0 : 32;
  const uint8_t* acs = V.barena + vf.acs;
  const uint8_t* qdc = V.barena + vf.qdc;
  const uint16_t* rawq = reinterpret_cast<const uint16_t*>(V.barena + vf.rawq);
  const uint32_t* bthr = V.upool + vf.bctx_off;
  const uint32_t nqf = vf.num_qf_thr;
  const uint32_t bmap_off = vf.num_dc_thr[0] + vf.num_dc_thr[1] + vf.num_dc_thr[2] + nqf;
  uint32_t* list = V.uarena + vf.blist + 2 * (static_cast<size_t>(y0) * W + static_cast<size_t>(x0) * ys);
  uint32_t n = 0;
  bool hole = false;
  for (uint32_t by = 0; by < ys && !hole; by++) {
    for (uint32_t bx = 0; bx < xs;) {
      const size_t pos = static_cast<size_t>(y0 + by) * W + x0 + bx;
      const uint8_t a = acs[pos];
      if (a == 0xFF) {
        hole = true;
        break;
      }
      const StrategyInfo si = UnpackStrategyInfo(JXLB_LDG(V.upool + V.sinfo_off + (a >> 1)));
      if (a & 1) {
        const uint32_t qf = rawq[pos];
        uint32_t qf_idx = 0;
        for (uint32_t t = 0; t < nqf; t++) qf_idx += qf > bthr[bmap_off - nqf + t];
        uint32_t ctxs = 0;
        for (uint32_t c = 0; c < 3; c++) {
          uint32_t idx = (c < 2 ? c ^ 1 : 2) * kNumOrders + si.order;
          idx = idx * (nqf + 1) + qf_idx;
          idx = idx * vf.num_dc_ctxs + qdc[pos];
          ctxs |= ((bthr[bmap_off + (idx >> 2)] >> (8 * (idx & 3))) & 0xFF) << (8 * c);
        }
        list[2 * n] = (by * 32 + bx) | (static_cast<uint32_t>(si.cx) << 10) | (static_cast<uint32_t>(si.log2_covered) << 16) |
                      (static_cast<uint32_t>(si.order) << 20) | (static_cast<uint32_t>(a >> 1) << 24);
        list[2 * n + 1] = ctxs;
        n++;
      }
      bx += si.cx;
    }
  }
  V.uarena[vf.blist_count + g] = hole ? 0xFFFFFFFFu : n;
}

struct DevAcLaneMem {
  uint8_t* colnz;          // [3 * 32] entries, element stride `stride`: last non-zero bucket per column
  uint32_t stride;
  const uint16_t* freq_ctx;  // kCoeffFreqContext[64]
  const uint16_t* nnz_ctx;   // kCoeffNumNonzeroContext[64]
  // kShared instances: the pass's tables in shared memory
  const DevAlias* alias_s = nullptr;
  const uint32_t* cfg_s = nullptr;
  const uint8_t* ctx_map_s = nullptr;
};

// kPlainAns: every AC code of the batch is ANS without LZ77 (the host checked): no mode tests per symbol.
// kShared (implies kPlainAns): alias tables, uint configs and the context map are read from m.alias_s / cfg_s / ctx_map_s.
template <bool kPlainAns = false, bool kShared = false>
JXLB_HD uint32_t DevDecodeAcStream(const DevPools& P, const DevVPools& V, uint32_t s, const DevAcLaneMem& m, bool valid) {
#define JXLB_AC_CTX(p) (kShared ? *(p) : JXLB_LDG(p))
  enum { kNeedBlock = 0, kReadNz = 1, kCoeff = 2, kDone = 3 };
  uint32_t mode = kDone, status = 0;
  DevAcStream st{};
  DevBits br{};
  DevSymbolReader reader{};
  DevCode code{};
  const DevVFrame* vf = nullptr;
  uint32_t W = 0, x0 = 0, y0 = 0, xs = 0, ys = 0, shift_unused = 0;
  uint32_t ctx_offset = 0, num_ctxs = 0;
  const uint8_t* ctx_map = nullptr;
  uint32_t* tok = nullptr;
  uint32_t* ts = nullptr;
  uint32_t* tc = nullptr;
  size_t nb = 0;
  if (valid) {
    st = V.streams[s];
    vf = V.frames + st.frame;
    W = vf->xblocks;
    nb = static_cast<size_t>(W) * vf->yblocks;
    const uint32_t gx = st.group % vf->xgroups, gy = st.group / vf->xgroups;
    x0 = gx * 32;
    y0 = gy * 32;
    xs = W - x0 < 32 ? W - x0 : 32;
    ys = vf->yblocks - y0 < 32 ? vf->yblocks - y0 : 32;
    br.Init(P.words, st.bit_pos, st.bit_end);
    num_ctxs = vf->num_ctxs;
    uint32_t selector_bits = 0;
    while ((1u << selector_bits) < vf->num_histograms) selector_bits++;
    const uint32_t cur = selector_bits ? br.Read(selector_bits) : 0;
    if (cur >= vf->num_histograms) status |= kVBadStream;
    ctx_offset = cur * num_ctxs * 495;
    code = P.codes[vf->ac_code[st.pass]];
    reader.Init(P, code, br, 0, nullptr);
    ctx_map = V.cpool + vf->ctx_map_off[st.pass];
    if (kShared) {
      reader.alias = m.alias_s;
      reader.cfg = m.cfg_s;
      ctx_map = m.ctx_map_s;
    }
    tok = V.tokens + st.tok_off;
    ts = V.uarena + vf->tok_start + static_cast<size_t>(st.pass) * 3 * nb;
    tc = V.uarena + vf->tok_count + static_cast<size_t>(st.pass) * 3 * nb;
    mode = (status == 0) ? kNeedBlock : kDone;
  }
  (void)shift_unused;
  // the group's varblock list (DevBuildBlockList): `rec` is the record of the varblock being decoded, `next` the one
  // after it, requested when `rec` is taken
  const uint32_t* blist = nullptr;
  uint32_t nblocks = 0, bi = 0;
  uint32_t next_x = 0, next_y = 0;
  if (valid && mode != kDone) {
    blist = V.uarena + vf->blist + 2 * (static_cast<size_t>(y0) * W + static_cast<size_t>(x0) * ys);
    nblocks = V.uarena[vf->blist_count + st.group];
    if (nblocks == 0xFFFFFFFFu) {
      status |= kVBadStream;
      mode = kDone;
      nblocks = 0;
    } else if (nblocks > 0) {
      next_x = blist[0];
      next_y = blist[1];
    }
  }
  const uint32_t LS = m.stride;
  uint32_t bx = 0, by = 0, ci = 3;                 // position inside the group, channel step (Y, X, B)
  uint32_t cx = 1, log2c = 0, covered = 1, size = 64, ord = 0, bctxs = 0;
  uint32_t c = 0, k = 0, nz = 0, prev = 0, histo_offset = 0, ctx = 0, ntok = 0, chan_start = 0;
  uint32_t pre_zero = 0, pre_nonzero = 0;
  bool pre_valid = false;
  size_t pos = 0;
  const uint16_t* order = nullptr;
  for (;;) {
    if (mode == kNeedBlock) {
      if (ci >= 3) {  // next varblock in raster order of the top-left blocks
        if (bi >= nblocks) {
          mode = kDone;
        } else {
          const uint32_t rec = next_x;
          bctxs = next_y;
          bi++;
          if (bi < nblocks) {
            next_x = blist[2 * bi];
            next_y = blist[2 * bi + 1];
          }
          bx = rec & 31;
          by = (rec >> 5) & 31;
          cx = (rec >> 10) & 63;
          log2c = (rec >> 16) & 15;
          ord = (rec >> 20) & 15;
          covered = 1u << log2c;
          size = covered * 64;
          pos = static_cast<size_t>(y0 + by) * W + x0 + bx;
          ci = 0;
        }
      }
      if (mode != kDone) {
        c = ci == 0 ? 1 : (ci == 1 ? 0 : 2);
        const uint32_t block_ctx = (bctxs >> (8 * c)) & 0xFF;
        // predicted number of non-zeros from the top and left neighbours
        uint32_t predicted;
        const uint8_t* col = m.colnz + static_cast<size_t>(c * 32 + bx) * LS;
        if (bx == 0) {
          predicted = by == 0 ? 32 : col[0];
        } else if (by == 0) {
          predicted = col[-static_cast<ptrdiff_t>(LS)];
        } else {
          predicted = (static_cast<uint32_t>(col[0]) + col[-static_cast<ptrdiff_t>(LS)] + 1) / 2;
        }
        uint32_t bucket = predicted >= 64 ? 64 : predicted;
        bucket = bucket < 8 ? bucket : 4 + bucket / 2;
        ctx = ctx_offset + bucket * num_ctxs + block_ctx;
        histo_offset = ctx_offset + num_ctxs * 37 + 458 * block_ctx;
        order = V.opool + JXLB_LDG(V.upool + vf->order_index + st.pass * 39 + 3 * ord + c);
        mode = kReadNz;
      }
    }
    if (JXLB_WARP_ALL(mode == kDone)) break;
    if (mode == kDone) continue;
    uint32_t cluster;
    if (mode == kCoeff && pre_valid) {
      cluster = prev ? pre_nonzero : pre_zero;  // looked up while the previous symbol was being decoded
    } else {
      if (mode == kCoeff) {
        const uint32_t nzl = (nz + covered - 1) >> log2c;
        ctx = histo_offset + (m.nnz_ctx[nzl] + m.freq_ctx[k >> log2c]) * 2 + prev;
      }
      cluster = JXLB_AC_CTX(ctx_map + ctx);
    }
    pre_valid = false;
    if (mode == kCoeff && k + 1 < size) {
      // The context of the next coefficient depends on this one only through (is it zero?): fetch the cluster
      // of both outcomes now, off the critical path of the serial rANS chain.
      const uint32_t f = m.freq_ctx[(k + 1) >> log2c];
      pre_zero = JXLB_AC_CTX(ctx_map + histo_offset + (m.nnz_ctx[(nz + covered - 1) >> log2c] + f) * 2);
      pre_nonzero = nz > 1 ? JXLB_AC_CTX(ctx_map + histo_offset + (m.nnz_ctx[(nz - 1 + covered - 1) >> log2c] + f) * 2 + 1) : 0;
      pre_valid = true;
    }
    const uint32_t u = kShared ? reader.template ReadUintPlainAns<true>(cluster, br)
                               : (kPlainAns ? reader.ReadUintPlainAns(cluster, br) : reader.ReadUint(cluster, br));
    bool chan_done = false;
    if (mode == kReadNz) {
      nz = u;
      if (nz > size - covered) {
        status |= kVBadStream;
        mode = kDone;
        continue;
      }
      const uint8_t v = static_cast<uint8_t>((nz + covered - 1) >> log2c);
      uint8_t* col = m.colnz + static_cast<size_t>(c * 32 + bx) * LS;
      for (uint32_t i = 0; i < cx; i++) col[static_cast<size_t>(i) * LS] = v;
      chan_start = ntok;
      ts[c * nb + pos] = static_cast<uint32_t>(st.tok_off) + (ntok < st.tok_cap ? ntok : st.tok_cap);
      k = covered;
      prev = nz > size / 16 ? 0 : 1;
      if (nz == 0) {
        chan_done = true;
      } else {
        mode = kCoeff;
      }
    } else {
      if (u != 0) {
        const uint32_t magnitude = u >> 1, neg_sign = (~u) & 1;
        const int32_t coeff = static_cast<int32_t>(magnitude ^ (neg_sign - 1));
        if (coeff > 32767 || coeff < -32767) status |= kVUnsupported;  // the token format holds 16-bit values
        if (ntok < st.tok_cap) tok[ntok] = static_cast<uint32_t>(JXLB_LDG(order + k)) | (static_cast<uint32_t>(coeff) << 16);
        ntok++;
        nz--;
        prev = 1;
      } else {
        prev = 0;
      }
      k++;
      if (nz == 0) {
        chan_done = true;
      } else if (k >= size) {
        status |= kVBadStream;  // non-zeros left at the end of the block
        chan_done = true;
      }
    }
    if (chan_done) {
      pre_valid = false;
      tc[c * nb + pos] = (ntok < st.tok_cap ? ntok : st.tok_cap) - (chan_start < st.tok_cap ? chan_start : st.tok_cap);
      ci++;
      mode = kNeedBlock;
    }
  }
  if (valid) {
    if (!code.use_prefix && reader.state != (0x13u << 16)) status |= kVBadFinalState;
    if (br.Pos() > st.bit_end) status |= kVOverread;
    if (ntok > st.tok_cap) status |= kVTokenOverflow;
    V.ac_used[s] = ntok;
    if (st.chain_slot != 0) V.chain_pos[st.chain_slot - 1] = br.Pos();
  }
  return status;
#undef JXLB_AC_CTX
}

// ---------------------------------------------------------------- 1-D transforms, staged
// An N-point transform of libjxl's recursive DCT (lib/jxl/dct-inl.h:167-222) unrolled into
// 2 * log2(N) - 1 stages in which every element is computed independently: any number of
// threads can share a stage, and the arithmetic per element is exactly the recursion's.
// A "line" is the 1-D signal; element e of line l lives at buf[l * ls + e * es].

// Inverse. Returns the buffer that holds the result (src or dst).
template <int SCOPE>
JXLB_HD float* CoopIDCT(uint32_t n, uint32_t lines, uint32_t ls, uint32_t es, float* src, float* dst, const float* wc,
                        uint32_t tid, uint32_t nt) {
  if (n < 2) return src;
  uint32_t log2n = 0;
  while ((1u << log2n) < n) log2n++;
  const uint32_t total = lines * n;
  // downward: even/odd split of every sub-problem of size m, B-transpose on the odd half
  for (uint32_t m = n; m > 2; m >>= 1) {
    const uint32_t h = m >> 1;
    for (uint32_t i = tid; i < total; i += nt) {
      const uint32_t l = i >> log2n, e = i & (n - 1);
      const uint32_t sub = e & ~(m - 1), r = e & (m - 1);
      const float* line = src + static_cast<size_t>(l) * ls;
      float v;
      if (r < h) {
        v = line[static_cast<size_t>(sub + 2 * r) * es];
      } else {
        const uint32_t j = r - h;
        const float t = line[static_cast<size_t>(sub + 2 * j + 1) * es];
        v = j == 0 ? t * kDevSqrt2 : t + line[static_cast<size_t>(sub + 2 * j - 1) * es];
      }
      dst[static_cast<size_t>(l) * ls + static_cast<size_t>(e) * es] = v;
    }
    CoopSync<SCOPE>();
    float* t = src;
    src = dst;
    dst = t;
  }
  // 2-point butterflies
  for (uint32_t i = tid; i < total; i += nt) {
    const uint32_t l = i >> log2n, e = i & (n - 1);
    const float* line = src + static_cast<size_t>(l) * ls;
    const float a = line[static_cast<size_t>(e & ~1u) * es], b = line[static_cast<size_t>(e | 1u) * es];
    dst[static_cast<size_t>(l) * ls + static_cast<size_t>(e) * es] = (e & 1) ? a - b : a + b;
  }
  CoopSync<SCOPE>();
  {
    float* t = src;
    src = dst;
    dst = t;
  }
  // upward: out[i] = in1 + mul[i] * in2, out[m - 1 - i] = in1 - mul[i] * in2
  for (uint32_t m = 4; m <= n; m <<= 1) {
    const uint32_t h = m >> 1;
    const float* mul = wc + (m / 2 - 2);
    for (uint32_t i = tid; i < total; i += nt) {
      const uint32_t l = i >> log2n, e = i & (n - 1);
      const uint32_t sub = e & ~(m - 1), r = e & (m - 1);
      const float* line = src + static_cast<size_t>(l) * ls;
      const uint32_t j = r < h ? r : m - 1 - r;
      const float in1 = line[static_cast<size_t>(sub + j) * es], in2 = line[static_cast<size_t>(sub + h + j) * es];
      const float w = mul[j];
      dst[static_cast<size_t>(l) * ls + static_cast<size_t>(e) * es] = fmaf(r < h ? w : -w, in2, in1);
    }
    CoopSync<SCOPE>();
    float* t = src;
    src = dst;
    dst = t;
  }
  return src;
}

// Forward (DCT1DImpl), scaled by 1 / n. Returns the buffer that holds the result.
template <int SCOPE>
JXLB_HD float* CoopDCT(uint32_t n, uint32_t lines, uint32_t ls, uint32_t es, float* src, float* dst, const float* wc,
                       uint32_t tid, uint32_t nt) {
  uint32_t log2n = 0;
  while ((1u << log2n) < n) log2n++;
  const uint32_t total = lines * n;
  if (n >= 2) {
    for (uint32_t m = n; m > 2; m >>= 1) {  // AddReverse / SubReverse + Multiply
      const uint32_t h = m >> 1;
      const float* mul = wc + (m / 2 - 2);
      for (uint32_t i = tid; i < total; i += nt) {
        const uint32_t l = i >> log2n, e = i & (n - 1);
        const uint32_t sub = e & ~(m - 1), r = e & (m - 1);
        const float* line = src + static_cast<size_t>(l) * ls;
        const uint32_t j = r < h ? r : r - h;
        const float a = line[static_cast<size_t>(sub + j) * es], b = line[static_cast<size_t>(sub + m - 1 - j) * es];
        dst[static_cast<size_t>(l) * ls + static_cast<size_t>(e) * es] = r < h ? a + b : (a - b) * mul[j];
      }
      CoopSync<SCOPE>();
      float* t = src;
      src = dst;
      dst = t;
    }
    for (uint32_t i = tid; i < total; i += nt) {
      const uint32_t l = i >> log2n, e = i & (n - 1);
      const float* line = src + static_cast<size_t>(l) * ls;
      const float a = line[static_cast<size_t>(e & ~1u) * es], b = line[static_cast<size_t>(e | 1u) * es];
      dst[static_cast<size_t>(l) * ls + static_cast<size_t>(e) * es] = (e & 1) ? a - b : a + b;
    }
    CoopSync<SCOPE>();
    {
      float* t = src;
      src = dst;
      dst = t;
    }
    for (uint32_t m = 4; m <= n; m <<= 1) {  // B + InverseEvenOdd
      const uint32_t h = m >> 1;
      for (uint32_t i = tid; i < total; i += nt) {
        const uint32_t l = i >> log2n, e = i & (n - 1);
        const uint32_t sub = e & ~(m - 1), r = e & (m - 1);
        const float* line = src + static_cast<size_t>(l) * ls;
        float v;
        if ((r & 1) == 0) {
          v = line[static_cast<size_t>(sub + (r >> 1)) * es];
        } else {
          const uint32_t j = r >> 1;
          const float t0 = line[static_cast<size_t>(sub + h + j) * es];
          if (j == 0) {
            v = fmaf(t0, kDevSqrt2, line[static_cast<size_t>(sub + h + 1) * es]);
          } else if (j + 1 < h) {
            v = t0 + line[static_cast<size_t>(sub + h + j + 1) * es];
          } else {
            v = t0;
          }
        }
        dst[static_cast<size_t>(l) * ls + static_cast<size_t>(e) * es] = v;
      }
      CoopSync<SCOPE>();
      float* t = src;
      src = dst;
      dst = t;
    }
  }
  const float scale = 1.0f / static_cast<float>(n);
  for (uint32_t i = tid; i < total; i += nt) {
    const uint32_t l = i >> log2n, e = i & (n - 1);
    const size_t at = static_cast<size_t>(l) * ls + static_cast<size_t>(e) * es;
    src[at] = scale * src[at];
  }
  CoopSync<SCOPE>();
  return src;
}

// ---------------------------------------------------------------- scalar 8x8 special transforms
// IDENTITY, DCT2X2, DCT4X4, DCT4X8, DCT8X4 and AFV0-3 cover one 8x8 block; one thread per channel
// runs libjxl's statements (lib/jxl/dec_transforms-inl.h:61-88, :380-449, :458-576).
template <int N>
JXLB_HD void ScalarIDCT1D(const float* from, int fs, float* to, int ts, const float* wc) {
  if (N == 1) {
    to[0] = from[0];
    return;
  }
  if (N == 2) {
    const float a = from[0], b = from[fs];
    to[0] = a + b;
    to[ts] = a - b;
    return;
  }
  constexpr int H = N / 2 > 0 ? N / 2 : 1;
  float tmp[N];
  for (int i = 0; i < H; i++) tmp[i] = from[2 * i * fs];
  for (int i = H; i < N; i++) tmp[i] = from[(2 * (i - H) + 1) * fs];
  ScalarIDCT1D<H>(tmp, 1, tmp, 1, wc);
  for (int i = H - 1; i > 0; i--) tmp[H + i] = tmp[H + i] + tmp[H + i - 1];
  tmp[H] = tmp[H] * kDevSqrt2;
  ScalarIDCT1D<H>(tmp + H, 1, tmp + H, 1, wc);
  const float* mul = wc + (N / 2 - 2);
  for (int i = 0; i < H; i++) {
    const float in1 = tmp[i], in2 = tmp[H + i];
    to[i * ts] = fmaf(mul[i], in2, in1);
    to[(N - i - 1) * ts] = fmaf(-mul[i], in2, in1);
  }
}
template <>
JXLB_HD void ScalarIDCT1D<0>(const float*, int, float*, int, const float*) {}

// ComputeScaledIDCT<R, C> for R, C in {4, 8}; `from` is in the min x max layout.
template <int R, int C>
JXLB_HD void ScalarScaledIDCT(const float* from, float* to, int to_stride, const float* wc) {
  float a[R * C], b[R * C];
  // x direction (C-point) first, then y (R-point); the layout rules follow ComputeScaledIDCT
  if (R < C) {
    // from: R rows x C cols, [yfreq][xfreq]
    for (int i = 0; i < R; i++) ScalarIDCT1D<C>(from + i * C, 1, a + i * C, 1, wc);      // a[yfreq][x]
    for (int x = 0; x < C; x++) ScalarIDCT1D<R>(a + x, C, b + x, C, wc);                  // b[y][x]
    for (int y = 0; y < R; y++)
      for (int x = 0; x < C; x++) to[y * to_stride + x] = b[y * C + x];
  } else {
    // from: C rows x R cols, [xfreq][yfreq]
    for (int j = 0; j < R; j++) ScalarIDCT1D<C>(from + j, R, a + j, R, wc);               // a[x][yfreq]
    for (int x = 0; x < C; x++) ScalarIDCT1D<R>(a + x * R, 1, b + x * R, 1, wc);          // b[x][y]
    for (int y = 0; y < R; y++)
      for (int x = 0; x < C; x++) to[y * to_stride + x] = b[x * R + y];
  }
}

JXLB_HD void ScalarIDCT2TopBlock(int S, float* block) {  // in place on an 8x8 row-major block
  float temp[64];
  const int num = S / 2;
  for (int y = 0; y < num; y++)
    for (int x = 0; x < num; x++) {
      const float c00 = block[y * 8 + x], c01 = block[y * 8 + num + x];
      const float c10 = block[(y + num) * 8 + x], c11 = block[(y + num) * 8 + num + x];
      temp[y * 2 * 8 + x * 2] = c00 + c01 + c10 + c11;
      temp[y * 2 * 8 + x * 2 + 1] = c00 + c01 - c10 - c11;
      temp[(y * 2 + 1) * 8 + x * 2] = c00 - c01 + c10 - c11;
      temp[(y * 2 + 1) * 8 + x * 2 + 1] = c00 - c01 - c10 + c11;
    }
  for (int y = 0; y < S; y++)
    for (int x = 0; x < S; x++) block[y * 8 + x] = temp[y * 8 + x];
}

// coefficients: 64 floats (8x8 row-major); pixels: 8 rows of `stride`.
JXLB_HD void DevSpecialToPixels(uint32_t strategy, const float* coefficients, float* pixels, int stride, const float* wc,
                                const float* afv_basis) {
  switch (strategy) {
    case 1: {  // IDENTITY
      const float b00 = coefficients[0], b01 = coefficients[1], b10 = coefficients[8], b11 = coefficients[9];
      const float dcs[4] = {b00 + b01 + b10 + b11, b00 + b01 - b10 - b11, b00 - b01 + b10 - b11, b00 - b01 - b10 + b11};
      for (int y = 0; y < 2; y++)
        for (int x = 0; x < 2; x++) {
          const float block_dc = dcs[y * 2 + x];
          float residual_sum = 0;
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) {
              if (ix == 0 && iy == 0) continue;
              residual_sum += coefficients[(y + iy * 2) * 8 + x + ix * 2];
            }
          const float centre = block_dc - residual_sum * (1.0f / 16);
          pixels[(4 * y + 1) * stride + 4 * x + 1] = centre;
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) {
              if (ix == 1 && iy == 1) continue;
              pixels[(y * 4 + iy) * stride + x * 4 + ix] = coefficients[(y + iy * 2) * 8 + x + ix * 2] + centre;
            }
          pixels[y * 4 * stride + x * 4] = coefficients[(y + 2) * 8 + x + 2] + centre;
        }
      return;
    }
    case 13:    // DCT8X4
    case 12: {  // DCT4X8
      const float block0 = coefficients[0], block1 = coefficients[8];
      const float dcs[2] = {block0 + block1, block0 - block1};
      for (int h = 0; h < 2; h++) {
        float block[32];
        block[0] = dcs[h];
        for (int iy = 0; iy < 4; iy++)
          for (int ix = 0; ix < 8; ix++) {
            if (ix == 0 && iy == 0) continue;
            block[iy * 8 + ix] = coefficients[(h + iy * 2) * 8 + ix];
          }
        if (strategy == 13) {
          ScalarScaledIDCT<8, 4>(block, pixels + h * 4, stride, wc);
        } else {
          ScalarScaledIDCT<4, 8>(block, pixels + h * 4 * stride, stride, wc);
        }
      }
      return;
    }
    case 3: {  // DCT4X4
      const float b00 = coefficients[0], b01 = coefficients[1], b10 = coefficients[8], b11 = coefficients[9];
      const float dcs[4] = {b00 + b01 + b10 + b11, b00 + b01 - b10 - b11, b00 - b01 + b10 - b11, b00 - b01 - b10 + b11};
      for (int y = 0; y < 2; y++)
        for (int x = 0; x < 2; x++) {
          float block[16];
          block[0] = dcs[y * 2 + x];
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) {
              if (ix == 0 && iy == 0) continue;
              block[iy * 4 + ix] = coefficients[(y + iy * 2) * 8 + x + ix * 2];
            }
          ScalarScaledIDCT<4, 4>(block, pixels + y * 4 * stride + x * 4, stride, wc);
        }
      return;
    }
    case 2: {  // DCT2X2
      float coeffs[64];
      for (int i = 0; i < 64; i++) coeffs[i] = coefficients[i];
      ScalarIDCT2TopBlock(2, coeffs);
      ScalarIDCT2TopBlock(4, coeffs);
      ScalarIDCT2TopBlock(8, coeffs);
      for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) pixels[y * stride + x] = coeffs[y * 8 + x];
      return;
    }
    default: {  // AFV0..3
      const int afv_kind = static_cast<int>(strategy) - 14;
      const int afv_x = afv_kind & 1, afv_y = afv_kind / 2;
      const float block00 = coefficients[0], block01 = coefficients[1], block10 = coefficients[8];
      const float dcs[3] = {(block00 + block10 + block01) * 4.0f, (block00 + block10 - block01), block00 - block10};
      float coeff[16];
      coeff[0] = dcs[0];
      for (int iy = 0; iy < 4; iy++)
        for (int ix = 0; ix < 4; ix++) {
          if (ix == 0 && iy == 0) continue;
          coeff[iy * 4 + ix] = coefficients[iy * 2 * 8 + ix * 2];
        }
      float block[32];
      for (int i = 0; i < 16; i++) {
        float pixel = 0.0f;
        for (int j = 0; j < 16; j++) pixel = fmaf(coeff[j], afv_basis[j * 16 + i], pixel);
        block[i] = pixel;
      }
      for (int iy = 0; iy < 4; iy++)
        for (int ix = 0; ix < 4; ix++)
          pixels[(iy + afv_y * 4) * stride + afv_x * 4 + ix] = block[(afv_y == 1 ? 3 - iy : iy) * 4 + (afv_x == 1 ? 3 - ix : ix)];
      block[0] = dcs[1];
      for (int iy = 0; iy < 4; iy++)
        for (int ix = 0; ix < 4; ix++) {
          if (ix == 0 && iy == 0) continue;
          block[iy * 4 + ix] = coefficients[iy * 2 * 8 + ix * 2 + 1];
        }
      ScalarScaledIDCT<4, 4>(block, pixels + afv_y * 4 * stride + (afv_x == 1 ? 0 : 4), stride, wc);
      block[0] = dcs[2];
      for (int iy = 0; iy < 4; iy++)
        for (int ix = 0; ix < 8; ix++) {
          if (ix == 0 && iy == 0) continue;
          block[iy * 8 + ix] = coefficients[(1 + iy * 2) * 8 + ix];
        }
      ScalarScaledIDCT<4, 8>(block, pixels + (afv_y == 1 ? 0 : 4) * stride, stride, wc);
      return;
    }
  }
}

// ---------------------------------------------------------------- one varblock
// AdjustQuantBias with an exact reciprocal (lib/jxl/quantizer-inl.h:34-71; the reference's x86
// builds use the 12-bit rcpps approximation there, see DESIGN.md "numerics").
JXLB_HD float DevAdjustQuantBias(int c, int32_t quant_i, const float* biases) {
  const float quant = static_cast<float>(quant_i);
  const float abs_quant = fabsf(quant);
  if (abs_quant < 1.125f) {
    if (!(abs_quant > 0.0f)) return 0.0f;
    return quant_i < 0 ? -biases[c] : biases[c];
  }
  return fmaf(-biases[3], 1.0f / quant, quant);
}

// Decodes varblock (bx, by) [absolute block coordinates, strategy `s`] of frame `vf` into the
// pixel planes pix[0]. `buf` provides 4 * 64 * covered floats of scratch visible to the `nt`
// cooperating threads (shared memory for <= 64x64 pixels, global memory above).
template <int SCOPE>
JXLB_HD void DevVarblock(const DevVPools& V, const DevVFrame& vf, uint32_t bx, uint32_t by, uint32_t s, float* buf,
                         uint32_t tid, uint32_t nt) {
  const StrategyInfo si = UnpackStrategyInfo(V.upool[V.sinfo_off + s]);
  const uint32_t covered = static_cast<uint32_t>(si.cx) * si.cy;
  const uint32_t N = covered * 64;
  const uint32_t W = vf.xblocks;
  const size_t nb = static_cast<size_t>(W) * vf.yblocks;
  const size_t pos = static_cast<size_t>(by) * W + bx;
  float* ch[3] = {buf, buf + N, buf + 2 * static_cast<size_t>(N)};
  float* scratch = buf + 3 * static_cast<size_t>(N);
  int32_t* qi = reinterpret_cast<int32_t*>(buf);
  const float* wc = V.fpool + V.wc_off;
  // 1. quantised coefficients: zero, then add every pass's tokens (positions are unique within a pass)
  for (uint32_t i = tid; i < 3 * N; i += nt) qi[i] = 0;
  CoopSync<SCOPE>();
  for (uint32_t p = 0; p < vf.num_passes; p++) {
    const uint32_t shift = vf.pass_shift[p];
    for (uint32_t c = 0; c < 3; c++) {
      const size_t e = (static_cast<size_t>(p) * 3 + c) * nb + pos;
      const uint32_t start = V.uarena[vf.tok_start + e], count = V.uarena[vf.tok_count + e];
      const uint32_t* tok = V.tokens + start;
      int32_t* q = qi + static_cast<size_t>(c) * N;
      for (uint32_t i = tid; i < count; i += nt) {
        const uint32_t t = JXLB_LDG(tok + i);
        const int32_t val = static_cast<int16_t>(t >> 16);
        const uint32_t k = t & 0xFFFF;
        if (k < N) q[k] = static_cast<int32_t>(static_cast<uint32_t>(q[k]) + (static_cast<uint32_t>(val) << shift));
      }
    }
    CoopSync<SCOPE>();
  }
  // 2. dequantisation, quantisation-bias adjustment, chroma from luma (DequantBlock)
  {
    const uint16_t* rawq = reinterpret_cast<const uint16_t*>(V.barena + vf.rawq);
    const float scaled = vf.inv_global_scale / static_cast<float>(rawq[pos]);
    const float sd0 = scaled * vf.x_dm, sd1 = scaled, sd2 = scaled * vf.b_dm;
    const size_t tile = static_cast<size_t>(by / 8) * vf.cmw + bx / 8;
    const int8_t fx = reinterpret_cast<const int8_t*>(V.barena + vf.ytox)[tile];
    const int8_t fb = reinterpret_cast<const int8_t*>(V.barena + vf.ytob)[tile];
    const float x_cc = vf.base_x + static_cast<float>(fx) * vf.color_scale;
    const float b_cc = vf.base_b + static_cast<float>(fb) * vf.color_scale;
    const float* dm = V.fpool + vf.table_off[si.table];
    for (uint32_t k = tid; k < N; k += nt) {
      const float x_mul = JXLB_LDG(dm + k) * sd0, y_mul = JXLB_LDG(dm + N + k) * sd1, b_mul = JXLB_LDG(dm + 2 * N + k) * sd2;
      const int32_t qx = qi[k], qy = qi[N + k], qb = qi[2 * N + k];
      const float dq_x = DevAdjustQuantBias(0, qx, vf.biases) * x_mul;
      const float dq_y = DevAdjustQuantBias(1, qy, vf.biases) * y_mul;
      const float dq_b = DevAdjustQuantBias(2, qb, vf.biases) * b_mul;
      ch[0][k] = fmaf(x_cc, dq_y, dq_x);
      ch[1][k] = dq_y;
      ch[2][k] = fmaf(b_cc, dq_y, dq_b);
    }
  }
  CoopSync<SCOPE>();
  // 3. lowest frequencies from the DC image (LowestFrequenciesFromDC)
  const uint32_t Rb = si.cy, Cb = si.cx;
  for (uint32_t c = 0; c < 3; c++) {
    const float* dcp = V.farena + vf.dc_final[c] + pos;
    if (!si.plain_dct || s == 0) {
      if (tid == 0) ch[c][0] = dcp[0];
      continue;
    }
    float* a = scratch;
    float* b = scratch + covered;
    for (uint32_t i = tid; i < covered; i += nt) a[i] = dcp[static_cast<size_t>(i / Cb) * W + i % Cb];
    CoopSync<SCOPE>();
    float* r1 = CoopDCT<SCOPE>(Rb, Cb, 1, Cb, a, b, wc, tid, nt);            // along y for every x
    float* r2 = CoopDCT<SCOPE>(Cb, Rb, Cb, 1, r1, r1 == a ? b : a, wc, tid, nt);  // along x for every y
    const float* ks = V.fpool + V.llf_off;
    if (Rb < Cb) {
      const uint32_t out_stride = 8 * Cb;
      for (uint32_t i = tid; i < covered; i += nt) {
        const uint32_t y = i / Cb, x = i % Cb;
        ch[c][y * out_stride + x] = r2[y * Cb + x] * ks[Rb - 1 + y] * ks[Cb - 1 + x];
      }
    } else {
      const uint32_t out_stride = 8 * Rb;
      for (uint32_t i = tid; i < covered; i += nt) {
        const uint32_t y = i / Rb, x = i % Rb;  // y < Cb (x frequency), x < Rb (y frequency)
        ch[c][y * out_stride + x] = r2[x * Cb + y] * ks[Cb - 1 + y] * ks[Rb - 1 + x];
      }
    }
    CoopSync<SCOPE>();
  }
  CoopSync<SCOPE>();
  // 4. inverse transform into the pixel planes
  const uint32_t PW = W * 8;
  if (!si.plain_dct) {
    for (uint32_t c = tid; c < 3; c += nt) {
      float* out = V.farena + vf.pix[0][c] + static_cast<size_t>(by) * 8 * PW + static_cast<size_t>(bx) * 8;
      DevSpecialToPixels(s, ch[c], out, static_cast<int>(PW), wc, V.fpool + V.afv_off);
    }
    CoopSync<SCOPE>();
    return;
  }
  const uint32_t R = 8 * Rb, C = 8 * Cb;
  for (uint32_t c = 0; c < 3; c++) {
    float* out = V.farena + vf.pix[0][c] + static_cast<size_t>(by) * 8 * PW + static_cast<size_t>(bx) * 8;
    const float* res;
    if (R >= C) {
      // coef[i * R + j]: i = horizontal frequency, j = vertical frequency
      float* r1 = CoopIDCT<SCOPE>(C, R, 1, R, ch[c], scratch, wc, tid, nt);
      res = CoopIDCT<SCOPE>(R, C, R, 1, r1, r1 == ch[c] ? scratch : ch[c], wc, tid, nt);  // res[x * R + y]
      for (uint32_t i = tid; i < R * C; i += nt) {
        const uint32_t y = i / C, x = i % C;
        out[static_cast<size_t>(y) * PW + x] = res[x * R + y];
      }
    } else {
      // coef[i * C + j]: i = vertical frequency, j = horizontal frequency
      float* r1 = CoopIDCT<SCOPE>(C, R, C, 1, ch[c], scratch, wc, tid, nt);
      res = CoopIDCT<SCOPE>(R, C, 1, C, r1, r1 == ch[c] ? scratch : ch[c], wc, tid, nt);  // res[y * C + x]
      for (uint32_t i = tid; i < R * C; i += nt) {
        const uint32_t y = i / C, x = i % C;
        out[static_cast<size_t>(y) * PW + x] = res[y * C + x];
      }
    }
    CoopSync<SCOPE>();
  }
}

// ---------------------------------------------------------------- fast path: varblocks up to 32x32
// Same arithmetic as DevVarblock, organised for the common block sizes:
//  * coefficients live in a padded layout (row stride max(R, C) + 1 floats) so that both the row-wise and the
//    column-wise pass touch 32 different shared-memory banks;
//  * single-pass frames dequantise straight from the tokens (cost ~ non-zeros, not block size): Y first, then
//    X / B with their chroma-from-luma term read back from Y;
//  * every 1-D IDCT is one thread's register-resident, fully unrolled recursion (RegIDCT), in place; the second
//    pass stores its result directly into the pixel planes.
template <int N>
struct RegIDCT {
  static JXLB_HD void Run(float* v) {
    constexpr int H = N / 2;
    float t[N];
JXLB_UNROLL
    for (int i = 0; i < H; i++) {
      t[i] = v[2 * i];
      t[H + i] = v[2 * i + 1];
    }
    RegIDCT<H>::Run(t);
JXLB_UNROLL
    for (int i = H - 1; i > 0; i--) t[H + i] = t[H + i] + t[H + i - 1];
    t[H] = t[H] * kDevSqrt2;
    RegIDCT<H>::Run(t + H);
JXLB_UNROLL
    for (int i = 0; i < H; i++) {
      const float w = WcMul(N, i);
      v[i] = fmaf(w, t[H + i], t[i]);
      v[N - 1 - i] = fmaf(-w, t[H + i], t[i]);
    }
  }
};
template <>
struct RegIDCT<2> {
  static JXLB_HD void Run(float* v) {
    const float a = v[0], b = v[1];
    v[0] = a + b;
    v[1] = a - b;
  }
};

// In-place N-point IDCT of `lines` lines per channel (3 channels of P floats each): element e of line l at
// base[c * P + l * lstride + e * estride].
template <int N>
JXLB_HD void IdctLinesInPlace(float* buf, uint32_t P, uint32_t lines, uint32_t lstride, uint32_t estride, uint32_t tid,
                              uint32_t nt) {
  for (uint32_t l = tid; l < 3 * lines; l += nt) {
    const uint32_t c = l / lines, li = l - c * lines;
    float* p = buf + c * P + li * lstride;
    float v[N];
JXLB_UNROLL
    for (int e = 0; e < N; e++) v[e] = p[e * estride];
    RegIDCT<N>::Run(v);
JXLB_UNROLL
    for (int e = 0; e < N; e++) p[e * estride] = v[e];
  }
}

// Second pass: line x of channel c holds the column x of the block; results go to out[c][y * PW + x].
template <int N>
JXLB_HD void IdctLinesToPixels(const float* buf, uint32_t P, uint32_t lines, uint32_t lstride, uint32_t estride,
                               float* const out[3], uint32_t PW, uint32_t tid, uint32_t nt) {
  for (uint32_t l = tid; l < 3 * lines; l += nt) {
    const uint32_t c = l / lines, x = l - c * lines;
    const float* p = buf + c * P + x * lstride;
    float v[N];
JXLB_UNROLL
    for (int e = 0; e < N; e++) v[e] = p[e * estride];
    RegIDCT<N>::Run(v);
    float* o = out[c] + x;
JXLB_UNROLL
    for (int e = 0; e < N; e++) o[static_cast<size_t>(e) * PW] = v[e];
  }
}

constexpr uint32_t kFastBufFloats = 3 * 32 * 33 + 64;    // three padded 32x32 channels + LLF scratch
constexpr uint32_t kFastBufFloats64 = 3 * 64 * 65 + 128;  // the same for blocks up to 64x64

// Varblock with a plain DCT of at most MAXN x MAXN pixels (MAXN = 32: strategies 0, 4..11; MAXN = 64: also
// 18..20). `buf`: kFastBufFloats (kFastBufFloats64) floats shared by the cooperating threads.
// Where the tokens of a varblock are (single-pass frames) and its raw quant value: what a kernel can fetch one
// varblock ahead of the one it is working on.
struct DevBlockMeta {
  uint32_t start[3], count[3];  // per channel X, Y, B: first token, token count
  uint32_t rawq;
};

JXLB_HD DevBlockMeta DevLoadBlockMeta(const DevVPools& V, const DevVFrame& vf, size_t pos) {
  const size_t nb = static_cast<size_t>(vf.xblocks) * vf.yblocks;
  DevBlockMeta m;
  for (uint32_t c = 0; c < 3; c++) {
    m.start[c] = V.uarena[vf.tok_start + c * nb + pos];
    m.count[c] = V.uarena[vf.tok_count + c * nb + pos];
  }
  m.rawq = reinterpret_cast<const uint16_t*>(V.barena + vf.rawq)[pos];
  return m;
}

template <int SCOPE, int MAXN>
JXLB_HD void DevVarblockFast(const DevVPools& V, const DevVFrame& vf, uint32_t bx, uint32_t by, uint32_t s, float* buf,
                             uint32_t tid, uint32_t nt, const DevBlockMeta* prefetched = nullptr) {
  const StrategyInfo si = UnpackStrategyInfo(V.upool[V.sinfo_off + s]);
  const uint32_t Rb = si.cy, Cb = si.cx, covered = Rb * Cb;
  const uint32_t R = 8 * Rb, C = 8 * Cb, N = R * C;
  const uint32_t mx = R > C ? R : C, mn = R > C ? C : R;
  uint32_t log2mx = 3;
  while ((1u << log2mx) < mx) log2mx++;
  const uint32_t S = mx + 1, P = mn * S;
  const uint32_t W = vf.xblocks;
  const size_t nb = static_cast<size_t>(W) * vf.yblocks;
  const size_t pos = static_cast<size_t>(by) * W + bx;
  float* ch[3] = {buf, buf + P, buf + 2 * P};
  float* scratch = buf + 3 * P;
  const float* wc = V.fpool + V.wc_off;
  // dequantisation constants (DequantBlock)
  DevBlockMeta meta;
  if (prefetched) {
    meta = *prefetched;
  } else if (vf.num_passes == 1) {
    meta = DevLoadBlockMeta(V, vf, pos);
  } else {
    meta.rawq = reinterpret_cast<const uint16_t*>(V.barena + vf.rawq)[pos];
  }
  // single pass: the first `nt` tokens of all three channels are requested before anything waits on them
  uint32_t first_tok[3] = {0, 0, 0};
  if (vf.num_passes == 1) {
    for (uint32_t c = 0; c < 3; c++)
      if (tid < meta.count[c]) first_tok[c] = JXLB_LDG(V.tokens + meta.start[c] + tid);
  }
  const float scaled = vf.inv_global_scale / static_cast<float>(meta.rawq);
  const float sd[3] = {scaled * vf.x_dm, scaled, scaled * vf.b_dm};
  const size_t tile = static_cast<size_t>(by / 8) * vf.cmw + bx / 8;
  const float x_cc = vf.base_x + static_cast<float>(reinterpret_cast<const int8_t*>(V.barena + vf.ytox)[tile]) * vf.color_scale;
  const float b_cc = vf.base_b + static_cast<float>(reinterpret_cast<const int8_t*>(V.barena + vf.ytob)[tile]) * vf.color_scale;
  const float* dm = V.fpool + vf.table_off[si.table];
  for (uint32_t i = tid; i < 3 * P; i += nt) buf[i] = 0.0f;
  CoopSync<SCOPE>();
  if (vf.num_passes == 1) {
    // Y tokens: Y = dq_y; X = x_cc * dq_y (+ 0), B = b_cc * dq_y (+ 0)
    {
      const uint32_t count = meta.count[1];
      const uint32_t* tok = V.tokens + meta.start[1];
      for (uint32_t i = tid; i < count; i += nt) {
        const uint32_t t = i == tid ? first_tok[1] : JXLB_LDG(tok + i);
        const uint32_t k = t & 0xFFFF;
        if (k >= N) continue;
        const int32_t q = static_cast<int16_t>(t >> 16);
        const float dq_y = DevAdjustQuantBias(1, q, vf.biases) * (JXLB_LDG(dm + N + k) * sd[1]);
        const uint32_t kp = k + (k >> log2mx);
        ch[1][kp] = dq_y;
        ch[0][kp] = fmaf(x_cc, dq_y, 0.0f);
        ch[2][kp] = fmaf(b_cc, dq_y, 0.0f);
      }
    }
    CoopSync<SCOPE>();
    for (uint32_t c = 0; c < 3; c += 2) {
      const uint32_t count = meta.count[c];
      const uint32_t* tok = V.tokens + meta.start[c];
      const float cc = c == 0 ? x_cc : b_cc;
      for (uint32_t i = tid; i < count; i += nt) {
        const uint32_t t = i == tid ? first_tok[c] : JXLB_LDG(tok + i);
        const uint32_t k = t & 0xFFFF;
        if (k >= N) continue;
        const int32_t q = static_cast<int16_t>(t >> 16);
        const float dq = DevAdjustQuantBias(static_cast<int>(c), q, vf.biases) * (JXLB_LDG(dm + c * N + k) * sd[c]);
        const uint32_t kp = k + (k >> log2mx);
        ch[c][kp] = fmaf(cc, ch[1][kp], dq);
      }
    }
  } else {
    // several passes: sum the integer contributions first, then dequantise every position
    int32_t* qi = reinterpret_cast<int32_t*>(buf);
    for (uint32_t p = 0; p < vf.num_passes; p++) {
      const uint32_t shift = vf.pass_shift[p];
      for (uint32_t c = 0; c < 3; c++) {
        const size_t e = (static_cast<size_t>(p) * 3 + c) * nb + pos;
        const uint32_t start = V.uarena[vf.tok_start + e], count = V.uarena[vf.tok_count + e];
        const uint32_t* tok = V.tokens + start;
        int32_t* q = qi + c * P;
        for (uint32_t i = tid; i < count; i += nt) {
          const uint32_t t = JXLB_LDG(tok + i);
          const int32_t val = static_cast<int16_t>(t >> 16);
          const uint32_t k = t & 0xFFFF;
          if (k >= N) continue;
          const uint32_t kp = k + (k >> log2mx);
          q[kp] = static_cast<int32_t>(static_cast<uint32_t>(q[kp]) + (static_cast<uint32_t>(val) << shift));
        }
      }
      CoopSync<SCOPE>();
    }
    for (uint32_t k = tid; k < N; k += nt) {
      const uint32_t kp = k + (k >> log2mx);
      const int32_t qx = qi[kp], qy = qi[P + kp], qb = qi[2 * P + kp];
      if ((qx | qy | qb) == 0) continue;  // stays +0.0f: the bit pattern of integer 0
      const float dq_x = DevAdjustQuantBias(0, qx, vf.biases) * (JXLB_LDG(dm + k) * sd[0]);
      const float dq_y = DevAdjustQuantBias(1, qy, vf.biases) * (JXLB_LDG(dm + N + k) * sd[1]);
      const float dq_b = DevAdjustQuantBias(2, qb, vf.biases) * (JXLB_LDG(dm + 2 * N + k) * sd[2]);
      ch[0][kp] = fmaf(x_cc, dq_y, dq_x);
      ch[1][kp] = dq_y;
      ch[2][kp] = fmaf(b_cc, dq_y, dq_b);
    }
  }
  CoopSync<SCOPE>();
  // lowest frequencies from the DC image
  if (s == 0) {
    for (uint32_t c = tid; c < 3; c += nt) ch[c][0] = V.farena[vf.dc_final[c] + pos];
  } else {
    const float* ks = V.fpool + V.llf_off;
    for (uint32_t c = 0; c < 3; c++) {
      const float* dcp = V.farena + vf.dc_final[c] + pos;
      float* a = scratch;
      float* b = scratch + covered;
      for (uint32_t i = tid; i < covered; i += nt) a[i] = dcp[static_cast<size_t>(i / Cb) * W + i % Cb];
      CoopSync<SCOPE>();
      float* r1 = CoopDCT<SCOPE>(Rb, Cb, 1, Cb, a, b, wc, tid, nt);
      float* r2 = CoopDCT<SCOPE>(Cb, Rb, Cb, 1, r1, r1 == a ? b : a, wc, tid, nt);
      if (Rb < Cb) {
        for (uint32_t i = tid; i < covered; i += nt) {
          const uint32_t y = i / Cb, x = i % Cb;
          ch[c][y * S + x] = r2[y * Cb + x] * ks[Rb - 1 + y] * ks[Cb - 1 + x];
        }
      } else {
        for (uint32_t i = tid; i < covered; i += nt) {
          const uint32_t y = i / Rb, x = i % Rb;
          ch[c][y * S + x] = r2[x * Cb + y] * ks[Cb - 1 + y] * ks[Rb - 1 + x];
        }
      }
      CoopSync<SCOPE>();
    }
  }
  CoopSync<SCOPE>();
  // pass 1: C-point IDCT along x for each of the R lines
  const uint32_t l1 = R >= C ? 1 : S, e1 = R >= C ? S : 1;
  switch (C) {
    case 8: IdctLinesInPlace<8>(buf, P, R, l1, e1, tid, nt); break;
    case 16: IdctLinesInPlace<16>(buf, P, R, l1, e1, tid, nt); break;
    case 32: IdctLinesInPlace<32>(buf, P, R, l1, e1, tid, nt); break;
    default:
      if (MAXN >= 64) IdctLinesInPlace<(MAXN >= 64 ? 64 : 32)>(buf, P, R, l1, e1, tid, nt);
      break;
  }
  CoopSync<SCOPE>();
  // pass 2: R-point IDCT along y for each of the C columns, straight into the pixel planes
  const uint32_t PW = W * 8;
  const size_t origin = static_cast<size_t>(by) * 8 * PW + static_cast<size_t>(bx) * 8;
  float* const out[3] = {V.farena + vf.pix[0][0] + origin, V.farena + vf.pix[0][1] + origin, V.farena + vf.pix[0][2] + origin};
  const uint32_t l2 = R >= C ? S : 1, e2 = R >= C ? 1 : S;
  switch (R) {
    case 8: IdctLinesToPixels<8>(buf, P, C, l2, e2, out, PW, tid, nt); break;
    case 16: IdctLinesToPixels<16>(buf, P, C, l2, e2, out, PW, tid, nt); break;
    case 32: IdctLinesToPixels<32>(buf, P, C, l2, e2, out, PW, tid, nt); break;
    default:
      if (MAXN >= 64) IdctLinesToPixels<(MAXN >= 64 ? 64 : 32)>(buf, P, C, l2, e2, out, PW, tid, nt);
      break;
  }
  CoopSync<SCOPE>();
}

// One 8x8 varblock with a special transform (IDENTITY, DCT2X2, DCT4X4, DCT4X8, DCT8X4, AFV0-3) of a single-pass frame by
// `nt` cooperating threads (a group of four lanes in k_dequant_idct, eight such varblocks per warp at a time; SCOPE 1:
// every lane of the warp calls this in lock step, `active` = the lane's group has a varblock). Same arithmetic as
// DevVarblock: dequantisation straight from the tokens as in DevVarblockFast, DC into coefficient 0, then one thread per
// channel runs libjxl's statements (DevSpecialToPixels). `buf`: 3 channels of kSpecialChStride floats.
constexpr uint32_t kSpecialChStride = 65;  // (odd: the channel slots of the warp's 24 transform lanes start in different banks)
template <int SCOPE>
JXLB_HD void DevVarblockSpecial(const DevVPools& V, const DevVFrame& vf, uint32_t bx, uint32_t by, uint32_t s, float* buf,
                                uint32_t tid, uint32_t nt, const DevBlockMeta& meta, bool active) {
  const uint32_t W = vf.xblocks;
  const size_t pos = static_cast<size_t>(by) * W + bx;
  float* ch[3] = {buf, buf + kSpecialChStride, buf + 2 * kSpecialChStride};
  const StrategyInfo si = UnpackStrategyInfo(V.upool[V.sinfo_off + s]);
  const float* dm = V.fpool + vf.table_off[si.table];
  float sd[3] = {0.0f, 0.0f, 0.0f}, x_cc = 0.0f, b_cc = 0.0f;
  if (active) {
    const float scaled = vf.inv_global_scale / static_cast<float>(meta.rawq);
    sd[0] = scaled * vf.x_dm;
    sd[1] = scaled;
    sd[2] = scaled * vf.b_dm;
    const size_t tile = static_cast<size_t>(by / 8) * vf.cmw + bx / 8;
    x_cc = vf.base_x + static_cast<float>(reinterpret_cast<const int8_t*>(V.barena + vf.ytox)[tile]) * vf.color_scale;
    b_cc = vf.base_b + static_cast<float>(reinterpret_cast<const int8_t*>(V.barena + vf.ytob)[tile]) * vf.color_scale;
    for (uint32_t i = tid; i < 3 * kSpecialChStride; i += nt) buf[i] = 0.0f;
  }
  CoopSync<SCOPE>();
  if (active) {
    const uint32_t* tok = V.tokens + meta.start[1];
    for (uint32_t i = tid; i < meta.count[1]; i += nt) {
      const uint32_t t = JXLB_LDG(tok + i);
      const uint32_t k = t & 0xFFFF;
      if (k >= 64) continue;
      const int32_t q = static_cast<int16_t>(t >> 16);
      const float dq_y = DevAdjustQuantBias(1, q, vf.biases) * (JXLB_LDG(dm + 64 + k) * sd[1]);
      ch[1][k] = dq_y;
      ch[0][k] = fmaf(x_cc, dq_y, 0.0f);
      ch[2][k] = fmaf(b_cc, dq_y, 0.0f);
    }
  }
  CoopSync<SCOPE>();
  if (active) {
    for (uint32_t c = 0; c < 3; c += 2) {
      const uint32_t* tok = V.tokens + meta.start[c];
      const float cc = c == 0 ? x_cc : b_cc;
      for (uint32_t i = tid; i < meta.count[c]; i += nt) {
        const uint32_t t = JXLB_LDG(tok + i);
        const uint32_t k = t & 0xFFFF;
        if (k >= 64) continue;
        const int32_t q = static_cast<int16_t>(t >> 16);
        const float dq = DevAdjustQuantBias(static_cast<int>(c), q, vf.biases) * (JXLB_LDG(dm + c * 64 + k) * sd[c]);
        ch[c][k] = fmaf(cc, ch[1][k], dq);
      }
    }
  }
  CoopSync<SCOPE>();
  if (active) {
    const uint32_t PW = W * 8;
    for (uint32_t c = tid; c < 3; c += nt) {
      ch[c][0] = V.farena[vf.dc_final[c] + pos];
      float* out = V.farena + vf.pix[0][c] + static_cast<size_t>(by) * 8 * PW + static_cast<size_t>(bx) * 8;
      DevSpecialToPixels(s, ch[c], out, static_cast<int>(PW), V.fpool + V.wc_off, V.fpool + V.afv_off);
    }
  }
  CoopSync<SCOPE>();
}

// ---------------------------------------------------------------- render stages (per pixel)
JXLB_HD int DevMirror(int x, int size) {  // lib/jxl/image_ops.h:184-195
  while (x < 0 || x >= size) x = x < 0 ? -x - 1 : 2 * size - 1 - x;
  return x;
}

// INTERIOR: the caller guarantees that every access stays inside the frame (no mirroring needed).
template <bool INTERIOR>
struct DevPlaneView {
  const float* p;
  uint32_t stride;
  int xsize, ysize;
  JXLB_HD float At(int x, int y) const {
    if (INTERIOR) return p[static_cast<uint32_t>(y) * stride + static_cast<uint32_t>(x)];
    return p[static_cast<size_t>(DevMirror(y, ysize)) * stride + DevMirror(x, xsize)];
  }
};

template <bool INTERIOR>
JXLB_HD DevPlaneView<INTERIOR> DevView(const DevVPools& V, const DevVFrame& vf, uint32_t set, uint32_t c) {
  DevPlaneView<INTERIOR> v;
  v.p = V.farena + vf.pix[set][c];
  v.stride = vf.xblocks * 8;
  v.xsize = static_cast<int>(vf.xsize);
  v.ysize = static_cast<int>(vf.ysize);
  return v;
}

// Gaborish, one sample of channel c (lib/jxl/render_pipeline/stage_gaborish.cc:22-100). VIEW: anything with At(x, y)
// in frame coordinates (a plane in HBM, or a tile of it in shared memory).
template <class VIEW>
JXLB_HD float DevGaborishValue(const VIEW& m, const float* w, int x, int y) {
  const float sum0 = m.At(x, y);
  const float sum1 = (m.At(x - 1, y) + m.At(x + 1, y)) + (m.At(x, y - 1) + m.At(x, y + 1));
  const float sum2 = (m.At(x - 1, y - 1) + m.At(x + 1, y - 1)) + (m.At(x - 1, y + 1) + m.At(x + 1, y + 1));
  return fmaf(sum2, w[2], fmaf(sum1, w[1], sum0 * w[0]));
}

template <bool INTERIOR>
JXLB_HD void DevGaborishPixel(const DevVPools& V, const DevVFrame& vf, uint32_t in_set, uint32_t out_set, uint32_t c, int x,
                              int y) {
  const DevPlaneView<INTERIOR> m = DevView<INTERIOR>(V, vf, in_set, c);
  V.farena[vf.pix[out_set][c] + static_cast<size_t>(y) * m.stride + x] = DevGaborishValue(m, vf.gab_w[c], x, y);
}

// One pixel of EPF stage 0 / 1 / 2 (lib/jxl/render_pipeline/stage_epf.cc:43-500).
JXLB_HD float DevEpfSigma(const DevVPools& V, const DevVFrame& vf, int x, int y) {
  const uint32_t sbx = static_cast<uint32_t>(x) / 8 < vf.xblocks - 1 ? static_cast<uint32_t>(x) / 8 : vf.xblocks - 1;
  const uint32_t sby = static_cast<uint32_t>(y) / 8 < vf.yblocks - 1 ? static_cast<uint32_t>(y) / 8 : vf.yblocks - 1;
  return V.farena[vf.inv_sigma + static_cast<size_t>(sby) * vf.xblocks + sbx];
}

// The three channels of pixel (x, y) after EPF stage `stage`; m[c].At reads the stage's input in frame coordinates.
template <class VIEW>
JXLB_HD void DevEpfValue(const VIEW* m, const DevVFrame& vf, uint32_t stage, float row_sigma, int x, int y, float* out_x,
                         float* out_y, float* out_b) {
  float X = m[0].At(x, y), Y = m[1].At(x, y), B = m[2].At(x, y);
  const float kMinSigma = -3.90524291751269967465540850526868f;
  if (!(row_sigma < kMinSigma)) {
    const float sm = vf.epf_sigma_scale[stage];
    const float bsm = sm * vf.epf_border_sad_mul;
    const int iy = y & 7, ix = x & 7;  // (x, y are in-frame: non-negative)
    const float sad_mul = (iy == 0 || iy == 7 || ix == 0 || ix == 7) ? bsm : sm;
    const float inv_sigma = row_sigma * sad_mul;
    float w = 1.0f;
#define JXLB_EPF_ADD(dx, dy, sad)                                \
  {                                                              \
    float weight = fmaf((sad), inv_sigma, 1.0f);                 \
    if (weight < 0.0f) weight = 0.0f;                            \
    w = w + weight;                                              \
    X = fmaf(weight, m[0].At(x + (dx), y + (dy)), X);            \
    Y = fmaf(weight, m[1].At(x + (dx), y + (dy)), Y);            \
    B = fmaf(weight, m[2].At(x + (dx), y + (dy)), B);            \
  }
    if (stage == 0) {
      const int sads_off[12][2] = {{-2, 0}, {-1, -1}, {-1, 0}, {-1, 1}, {0, -2}, {0, -1},
                                   {0, 1},  {0, 2},   {1, -1}, {1, 0},  {1, 1},  {2, 0}};  // {row, col}
      const int plus_off[5][2] = {{0, 0}, {-1, 0}, {0, -1}, {1, 0}, {0, 1}};
      float sads[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < 3; c++) {
        const float scale = vf.epf_channel_scale[c];
        for (int i = 0; i < 12; i++) {
          float sad = 0.0f;
          for (int k = 0; k < 5; k++) {
            const float r11 = m[c].At(x + plus_off[k][1], y + plus_off[k][0]);
            const float c11 = m[c].At(x + sads_off[i][1] + plus_off[k][1], y + sads_off[i][0] + plus_off[k][0]);
            sad = sad + fabsf(r11 - c11);
          }
          sads[i] = fmaf(sad, scale, sads[i]);
        }
      }
      for (int i = 0; i < 12; i++) JXLB_EPF_ADD(sads_off[i][1], sads_off[i][0], sads[i]);
    } else if (stage == 1) {
      float sad0 = 0, sad1 = 0, sad2 = 0, sad3 = 0;
      for (int c = 0; c < 3; c++) {
#define JXLB_P(col, row) m[c].At(x + (col) - 2, y + (row) - 2)
        const float p20 = JXLB_P(2, 0), p21 = JXLB_P(2, 1);
        float sad0c = fabsf(p20 - p21);
        const float p11 = JXLB_P(1, 1);
        float sad1c = fabsf(p11 - p21);
        const float p31 = JXLB_P(3, 1);
        float sad2c = fabsf(p31 - p21);
        const float p02 = JXLB_P(0, 2), p12 = JXLB_P(1, 2);
        sad1c = sad1c + fabsf(p02 - p12);
        sad0c = sad0c + fabsf(p11 - p12);
        const float p22 = JXLB_P(2, 2);
        float t = fabsf(p12 - p22);
        sad1c = sad1c + t;
        sad2c = sad2c + t;
        t = fabsf(p22 - p21);
        float sad3c = t;
        sad0c = sad0c + t;
        const float p32 = JXLB_P(3, 2);
        sad0c = sad0c + fabsf(p31 - p32);
        t = fabsf(p22 - p32);
        sad1c = sad1c + t;
        sad2c = sad2c + t;
        const float p42 = JXLB_P(4, 2);
        sad2c = sad2c + fabsf(p42 - p32);
        const float p13 = JXLB_P(1, 3);
        sad3c = sad3c + fabsf(p13 - p12);
        const float p23 = JXLB_P(2, 3);
        t = fabsf(p22 - p23);
        sad0c = sad0c + t;
        sad3c = sad3c + t;
        sad1c = sad1c + fabsf(p13 - p23);
        const float p33 = JXLB_P(3, 3);
        sad2c = sad2c + fabsf(p33 - p23);
        sad3c = sad3c + fabsf(p33 - p32);
        const float p24 = JXLB_P(2, 4);
        sad3c = sad3c + fabsf(p24 - p23);
#undef JXLB_P
        const float scale = vf.epf_channel_scale[c];
        sad0 = fmaf(sad0c, scale, sad0);
        sad1 = fmaf(sad1c, scale, sad1);
        sad2 = fmaf(sad2c, scale, sad2);
        sad3 = fmaf(sad3c, scale, sad3);
      }
      JXLB_EPF_ADD(0, -1, sad0);
      JXLB_EPF_ADD(-1, 0, sad1);
      JXLB_EPF_ADD(1, 0, sad2);
      JXLB_EPF_ADD(0, 1, sad3);
    } else {
      const float rx = X, ry = Y, rb = B;
      const int offs[4][2] = {{0, -1}, {-1, 0}, {1, 0}, {0, 1}};
      for (int i = 0; i < 4; i++) {
        const int dx = offs[i][0], dy = offs[i][1];
        const float cx = m[0].At(x + dx, y + dy), cy = m[1].At(x + dx, y + dy), cb = m[2].At(x + dx, y + dy);
        float sad = fabsf(cx - rx) * vf.epf_channel_scale[0];
        sad = fmaf(fabsf(cy - ry), vf.epf_channel_scale[1], sad);
        sad = fmaf(fabsf(cb - rb), vf.epf_channel_scale[2], sad);
        JXLB_EPF_ADD(dx, dy, sad);
      }
    }
#undef JXLB_EPF_ADD
    const float inv_w = 1.0f / w;
    X = X * inv_w;
    Y = Y * inv_w;
    B = B * inv_w;
  }
  *out_x = X;
  *out_y = Y;
  *out_b = B;
}

template <bool INTERIOR>
JXLB_HD void DevEpfPixel(const DevVPools& V, const DevVFrame& vf, uint32_t stage, uint32_t in_set, uint32_t out_set, int x,
                         int y) {
  const DevPlaneView<INTERIOR> m[3] = {DevView<INTERIOR>(V, vf, in_set, 0), DevView<INTERIOR>(V, vf, in_set, 1),
                                       DevView<INTERIOR>(V, vf, in_set, 2)};
  const size_t at = static_cast<size_t>(y) * m[0].stride + x;
  float X, Y, B;
  DevEpfValue(m, vf, stage, DevEpfSigma(V, vf, x, y), x, y, &X, &Y, &B);
  V.farena[vf.pix[out_set][0] + at] = X;
  V.farena[vf.pix[out_set][1] + at] = Y;
  V.farena[vf.pix[out_set][2] + at] = B;
}

// ---- colour: XYB -> linear RGB -> transfer function (lib/jxl/dec_xyb-inl.h:37-83,
// lib/jxl/cms/transfer_functions-inl.h:219-242, lib/jxl/base/fast_math-inl.h:46-90)
JXLB_HD float DevRational2(float x, const float p[3], const float q[3]) {
  float yp = p[2], yq = q[2];
  yp = fmaf(yp, x, p[1]);
  yq = fmaf(yq, x, q[1]);
  yp = fmaf(yp, x, p[0]);
  yq = fmaf(yq, x, q[0]);
  return yp / yq;
}

JXLB_HD float DevBitsToFloat(int32_t b) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}
JXLB_HD int32_t DevFloatToBits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  int32_t b;
  memcpy(&b, &f, 4);
  return b;
#endif
}

JXLB_HD float DevFastPowf(float base, float exponent) {
  const float lp[3] = {-1.8503833400518310E-06f, 1.4287160470083755E+00f, 7.4245873327820566E-01f};
  const float lq[3] = {9.9032814277590719E-01f, 1.0096718572241148E+00f, 1.7409343003366853E-01f};
  const int32_t x_bits = DevFloatToBits(base);
  const int32_t exp_bits = x_bits - 0x3f2aaaab;
  const int32_t exp_shifted = exp_bits >> 23;
  const float mantissa = DevBitsToFloat(x_bits - static_cast<int32_t>(static_cast<uint32_t>(exp_shifted) << 23));
  const float log2 = DevRational2(mantissa - 1.0f, lp, lq) + static_cast<float>(exp_shifted);
  const float x = log2 * exponent;
  const float floorx = floorf(x);
  const float exp = DevBitsToFloat(static_cast<int32_t>(static_cast<uint32_t>(static_cast<int32_t>(floorx) + 127) << 23));
  const float frac = x - floorx;
  float num = frac + 1.01749063e+01f;
  num = fmaf(num, frac, 4.88687798e+01f);
  num = fmaf(num, frac, 9.85506591e+01f);
  num = num * exp;
  float den = fmaf(frac, 2.10242958e-01f, -2.22328856e-02f);
  den = fmaf(den, frac, -1.94414990e+01f);
  den = fmaf(den, frac, 9.85506633e+01f);
  return num / den;
}

JXLB_HD float DevSrgbFromLinear(float v) {
  const float p[5] = {-5.135152395e-04f, 5.287254571e-03f, 3.903842876e-01f, 1.474205315e+00f, 7.352629620e-01f};
  const float q[5] = {1.004519624e-02f, 3.036675394e-01f, 1.340816930e+00f, 9.258482155e-01f, 2.424867759e-02f};
  const float x = fabsf(v);
  const float linear = x * 12.92f;
  const float s = sqrtf(x);
  float yp = p[4], yq = q[4];
  for (int i = 3; i >= 0; i--) {
    yp = fmaf(yp, s, p[i]);
    yq = fmaf(yq, s, q[i]);
  }
  const float poly = yp / yq;
  const float magnitude = x > 0.0031308f ? poly : linear;
  return copysignf(magnitude, v);
}

JXLB_HD float DevFromLinear(const DevVFrame& vf, float v) {
  if (vf.tf == 2) return v <= 1e-5f ? 0.0f : DevFastPowf(v, vf.inv_gamma);
  if (vf.tf == 1) return DevSrgbFromLinear(v);
  return v;
}

// float sample -> output sample at (x, y), channel slot `idx` of the row (stage_write.cc:98-113)
JXLB_HD void DevStoreSample(uint8_t* row, size_t idx, float v, uint32_t data_type, uint32_t big_endian, uint32_t x,
                            uint32_t y) {
  if (data_type == 2 || data_type == 3) {
    const float mul = data_type == 2 ? 255.0f : 65535.0f;
    v = v * mul;
    if (data_type == 2) v = v + DevDither(x, y);
    if (!(v >= 0.0f)) v = 0.0f;
    if (v > mul) v = mul;
#if defined(__CUDA_ARCH__)
    const int r = __float2int_rn(v);
#else
    const int r = static_cast<int>(lrintf(v));
#endif
    if (data_type == 2) {
      row[idx] = static_cast<uint8_t>(r);
    } else {
      uint16_t u = static_cast<uint16_t>(r);
      if (big_endian) u = static_cast<uint16_t>((u >> 8) | (u << 8));
      reinterpret_cast<uint16_t*>(row)[idx] = u;
    }
  } else if (data_type == 5) {
    uint16_t u = DevFloatToHalf(v);
    if (big_endian) u = static_cast<uint16_t>((u >> 8) | (u << 8));
    reinterpret_cast<uint16_t*>(row)[idx] = u;
  } else {
    uint32_t u = static_cast<uint32_t>(DevFloatToBits(v));
    if (big_endian) u = ((u & 0xFF) << 24) | ((u & 0xFF00) << 8) | ((u >> 8) & 0xFF00) | (u >> 24);
    reinterpret_cast<uint32_t*>(row)[idx] = u;
  }
}

// Colour transform of one pixel of the filtered planes -> non-linear output samples in [0, 1].
JXLB_HD void DevColorTransform(const DevVFrame& vf, float p0, float p1, float p2, float* out_r, float* out_g, float* out_b) {
  float r, g, b;
  if (vf.color_transform == 0) {
    float gamma_r = p1 + p0, gamma_g = p1 - p0, gamma_b = p2;
    gamma_r = gamma_r - vf.opsin_bias_cbrt[0];
    gamma_g = gamma_g - vf.opsin_bias_cbrt[1];
    gamma_b = gamma_b - vf.opsin_bias_cbrt[2];
    const float r2 = gamma_r * gamma_r, g2 = gamma_g * gamma_g, b2 = gamma_b * gamma_b;
    const float mixed_r = fmaf(r2, gamma_r, vf.opsin_bias[0]);
    const float mixed_g = fmaf(g2, gamma_g, vf.opsin_bias[1]);
    const float mixed_b = fmaf(b2, gamma_b, vf.opsin_bias[2]);
    const float* mt = vf.inv_mat;
    float lr = mt[0] * mixed_r, lg = mt[3] * mixed_r, lb = mt[6] * mixed_r;
    lr = fmaf(mt[1], mixed_g, lr);
    lg = fmaf(mt[4], mixed_g, lg);
    lb = fmaf(mt[7], mixed_g, lb);
    lr = fmaf(mt[2], mixed_b, lr);
    lg = fmaf(mt[5], mixed_b, lg);
    lb = fmaf(mt[8], mixed_b, lb);
    r = DevFromLinear(vf, lr);
    g = DevFromLinear(vf, lg);
    b = DevFromLinear(vf, lb);
  } else if (vf.color_transform == 2) {
    const float c128 = 128.0f / 255, crcr = 1.402f, cgcb = -0.114f * 1.772f / 0.587f, cgcr = -0.299f * 1.402f / 0.587f,
                cbcb = 1.772f;
    const float yv = p1 + c128, cb = p0, cr = p2;
    r = fmaf(crcr, cr, yv);
    g = fmaf(cgcr, cr, fmaf(cgcb, cb, yv));
    b = fmaf(cbcb, cb, yv);
  } else {
    r = p0;
    g = p1;
    b = p2;
  }
  *out_r = r;
  *out_g = g;
  *out_b = b;
}

// float -> 8-bit sample with the ordered dither of stage_write.cc:98-113
JXLB_HD uint32_t DevToU8(float v, uint32_t x, uint32_t y) {
  v = v * 255.0f;
  v = v + DevDither(x, y);
  if (!(v >= 0.0f)) v = 0.0f;
  if (v > 255.0f) v = 255.0f;
#if defined(__CUDA_ARCH__)
  return static_cast<uint32_t>(__float2int_rn(v));
#else
  return static_cast<uint32_t>(lrintf(v));
#endif
}

// One output pixel: colour transform + sample conversion + interleaved store.
// Sample i of a reference-only frame: int -> float XYB (see DevRefFrame).
JXLB_HD void DevRefFrameSample(const DevPools& P, const DevVPools& V, const DevRefFrame& rf, uint32_t i) {
  const int32_t vy = P.arena[P.planes[rf.plane_y].off + i];
  const int32_t vx = P.arena[P.planes[rf.plane_x].off + i];
  const int32_t vb = P.arena[P.planes[rf.plane_b].off + i];
  V.farena[rf.dst[0] + i] = static_cast<float>(vx) * rf.factor[0];
  V.farena[rf.dst[1] + i] = static_cast<float>(vy) * rf.factor[1];
  V.farena[rf.dst[2] + i] = static_cast<float>(vb + vy) * rf.factor[2];
}

// Pixel (ix, iy) of one patch, all three channels, onto plane set `set` of the frame.
JXLB_HD void DevPatchPixel(const DevVPools& V, const DevVFrame& vf, const DevPatch& p, uint32_t set, uint32_t ix, uint32_t iy) {
  const uint32_t x = p.x + ix, y = p.y + iy;
  if (x >= vf.xsize || y >= vf.ysize || p.mode == 0) return;
  const size_t at = static_cast<size_t>(y) * (vf.xblocks * 8) + x;
  const size_t from = static_cast<size_t>(p.y0 + iy) * p.src_w + p.x0 + ix;
  for (uint32_t c = 0; c < 3; c++) {
    const float v = V.farena[p.src[c] + from];
    float& o = V.farena[vf.pix[set][c] + at];
    if (p.mode == 1) {
      o = v;
    } else if (p.mode == 2) {
      o = o + v;
    } else {
      o = o * (p.clamp ? fminf(1.0f, fmaxf(0.0f, v)) : v);
    }
  }
}

// One output pixel from its three filtered samples: colour transform + sample conversion + interleaved store.
JXLB_HD void DevColorStore(const DevVPools& V, const DevVFrame& vf, float p0, float p1, float p2, uint32_t x, uint32_t y) {
  if (vf.has_splines && vf.upsampling == 1) {  // (rare) SplineStage sits behind the loop filters and patches, in XYB
    // (in an upsampled frame it sits in front of the upsampling: DevUpsamplePixel)
    float v[3] = {p0, p1, p2};
    DevSplineAdd(V.spl_idx + vf.spl_rows, V.spl_idx + vf.spl_idx, V.spl_seg + vf.spl_seg, vf.xsize, x, y, v);
    p0 = v[0];
    p1 = v[1];
    p2 = v[2];
  }
  float r, g, b;
  DevColorTransform(vf, p0, p1, p2, &r, &g, &b);
  // alpha: the frame's Modular extra channel (int -> float like DevSampleFloat), else opaque
  float alpha = 1.0f;
  if (vf.alpha_plane != kNoPlane && (vf.out_channels == 2 || vf.out_channels == 4)) {
    const DevPlane pl = V.planes[vf.alpha_plane];
    alpha = static_cast<float>(V.arena[pl.off + static_cast<size_t>(y) * pl.w + x]) * vf.alpha_factor;
  }
  if (vf.orient != 0) {  // (rare) the store position and the dither position follow the orientation
    uint32_t fx, fy, orow, ocol;
    DevOrient(vf.orient, vf.up_xsize, vf.up_ysize, x, y, &fx, &fy, &orow, &ocol);
    uint8_t* row = V.out + vf.out_off + vf.out_stride * orow;
    const uint32_t nc = vf.out_channels, num_color = nc < 3 ? 1 : 3;
    const float col[3] = {r, g, b};
    for (uint32_t c = 0; c < nc; c++)
      DevStoreSample(row, static_cast<size_t>(ocol) * nc + c, c < num_color ? col[c] : alpha, vf.out_type, vf.out_big_endian, fx, fy);
    return;
  }
  uint8_t* row = V.out + vf.out_off + vf.out_stride * y;
  const uint32_t nc = vf.out_channels;
  if (nc == 4 && vf.out_type == 2) {  // RGBA8: one aligned 32-bit store
    const uint32_t a8 = vf.alpha_plane != kNoPlane ? DevToU8(alpha, x, y) : 255u;
    reinterpret_cast<uint32_t*>(row)[x] = DevToU8(r, x, y) | (DevToU8(g, x, y) << 8) | (DevToU8(b, x, y) << 16) | (a8 << 24);
    return;
  }
  const uint32_t num_color = nc < 3 ? 1 : 3;
  const float col[3] = {r, g, b};
  for (uint32_t c = 0; c < nc; c++) {
    const float v = c < num_color ? col[c] : alpha;
    DevStoreSample(row, static_cast<size_t>(x) * nc + c, v, vf.out_type, vf.out_big_endian, x, y);
  }
}

JXLB_HD void DevColorPixel(const DevVPools& V, const DevVFrame& vf, uint32_t set, uint32_t x, uint32_t y) {
  if (vf.upsampling > 1) {  // the colour transform runs on the upsampled planes
    const size_t at = static_cast<size_t>(y) * vf.up_stride + x;
    DevColorStore(V, vf, V.farena[vf.up_pix[0] + at], V.farena[vf.up_pix[1] + at], V.farena[vf.up_pix[2] + at], x, y);
    return;
  }
  const size_t at = static_cast<size_t>(y) * (vf.xblocks * 8) + x;
  DevColorStore(V, vf, V.farena[vf.pix[set][0] + at], V.farena[vf.pix[set][1] + at], V.farena[vf.pix[set][2] + at], x, y);
}

// Upsampling (lib/jxl/render_pipeline/stage_upsampling.cc:28-170): output pixel (ox, oy) of the three channels from
// the 5 x 5 neighbourhood (mirrored at the frame edges) of input pixel (ox / N, oy / N) in plane set `set`: MulAdd in
// the order iy = -2 .. 2, ix = -2 .. 2, clamped to the neighbourhood's minimum and maximum; kernel selection as
// Kernel<N> (:93-112).
JXLB_HD void DevUpsamplePixel(const DevVPools& V, const DevVFrame& vf, uint32_t set, uint32_t ox, uint32_t oy) {
  const uint32_t N = vf.upsampling;
  const int x = static_cast<int>(ox / N), y = static_cast<int>(oy / N);
  const uint32_t kx = ox % N, ky = oy % N, half = N / 2;
  // kernel[a][b][c][d]: a, c from y; b, d from x
  const bool fy = ky >= half && N > 1, fx = kx >= half && N > 1;  // second half of the period: mirrored kernel
  const uint32_t a = N == 2 ? 0 : (fy ? half - 1 - (ky % half) : ky % half);
  const uint32_t b = N == 2 ? 0 : (fx ? half - 1 - (kx % half) : kx % half);
  const float* kernel = V.fpool + vf.up_kernel + (a * 4 + b) * 25;
  const uint32_t PW = vf.xblocks * 8;
  const int xsize = static_cast<int>(vf.xsize), ysize = static_cast<int>(vf.ysize);
  if (vf.has_splines) {
    // (rare) libjxl draws the splines on the coded planes in front of the upsampling (lib/jxl/dec_cache.cc:178-196):
    // every tap is the plane sample with the splines of its own position added; taps and channels in the order of the
    // loop below, so that each channel's sum is the same sequence of operations
    float result[3] = {0.0f, 0.0f, 0.0f}, mn[3], mx[3];
    for (int iy = -2; iy <= 2; iy++) {
      const uint32_t sy = DevMirror(y + iy, ysize);
      const uint32_t kr = static_cast<uint32_t>(fy ? 2 - iy : iy + 2);
      for (int ix = -2; ix <= 2; ix++) {
        const uint32_t sx = DevMirror(x + ix, xsize);
        const size_t at = static_cast<size_t>(sy) * PW + sx;
        float v[3] = {V.farena[vf.pix[set][0] + at], V.farena[vf.pix[set][1] + at], V.farena[vf.pix[set][2] + at]};
        DevSplineAdd(V.spl_idx + vf.spl_rows, V.spl_idx + vf.spl_idx, V.spl_seg + vf.spl_seg, vf.xsize, sx, sy, v);
        const uint32_t kc = static_cast<uint32_t>(fx ? 2 - ix : ix + 2);
        const float w = kernel[kr * 5 + kc];
        for (uint32_t c = 0; c < 3; c++) {
          if (iy == -2 && ix == -2) mn[c] = mx[c] = v[c];
          result[c] = fmaf(w, v[c], result[c]);
          mn[c] = mn[c] < v[c] ? mn[c] : v[c];
          mx[c] = v[c] < mx[c] ? mx[c] : v[c];
        }
      }
    }
    for (uint32_t c = 0; c < 3; c++) {
      const float lo = result[c] < mn[c] ? mn[c] : result[c];
      V.farena[vf.up_pix[c] + static_cast<size_t>(oy) * vf.up_stride + ox] = mx[c] < lo ? mx[c] : lo;
    }
    return;
  }
  for (uint32_t c = 0; c < 3; c++) {
    const float* p = V.farena + vf.pix[set][c];
    const float centre = p[static_cast<size_t>(y) * PW + x];
    float result = 0.0f, mn = centre, mx = centre;
    for (int iy = -2; iy <= 2; iy++) {
      const size_t row = static_cast<size_t>(DevMirror(y + iy, ysize)) * PW;
      const uint32_t kr = static_cast<uint32_t>(fy ? 2 - iy : iy + 2);
      for (int ix = -2; ix <= 2; ix++) {
        const float v = p[row + DevMirror(x + ix, xsize)];
        const uint32_t kc = static_cast<uint32_t>(fx ? 2 - ix : ix + 2);
        result = fmaf(kernel[kr * 5 + kc], v, result);
        mn = mn < v ? mn : v;
        mx = v < mx ? mx : v;
      }
    }
    const float lo = result < mn ? mn : result;
    V.farena[vf.up_pix[c] + static_cast<size_t>(oy) * vf.up_stride + ox] = mx < lo ? mx : lo;
  }
}

// Four consecutive RGB8 pixels (x multiple of 4, row 4-byte aligned): 12 bytes as three 32-bit stores.
JXLB_HD void DevColorStoreRgb8x4(const DevVPools& V, const DevVFrame& vf, const float* p0, const float* p1, const float* p2,
                                 uint32_t x, uint32_t y) {
  uint32_t b[12];
JXLB_UNROLL
  for (uint32_t i = 0; i < 4; i++) {
    float r, g, bl;
    DevColorTransform(vf, p0[i], p1[i], p2[i], &r, &g, &bl);
    b[3 * i] = DevToU8(r, x + i, y);
    b[3 * i + 1] = DevToU8(g, x + i, y);
    b[3 * i + 2] = DevToU8(bl, x + i, y);
  }
  uint32_t* o = reinterpret_cast<uint32_t*>(V.out + vf.out_off + vf.out_stride * y + 3 * static_cast<size_t>(x));
  o[0] = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
  o[1] = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
  o[2] = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
}

JXLB_HD void DevColorPixelsRgb8x4(const DevVPools& V, const DevVFrame& vf, uint32_t set, uint32_t x, uint32_t y) {
  const size_t at = static_cast<size_t>(y) * (vf.xblocks * 8) + x;
  float p0[4], p1[4], p2[4];
#if defined(__CUDA_ARCH__)
  {  // planes and rows are 16-byte aligned and x is a multiple of 4
    const float4 a0 = *reinterpret_cast<const float4*>(V.farena + vf.pix[set][0] + at);
    const float4 a1 = *reinterpret_cast<const float4*>(V.farena + vf.pix[set][1] + at);
    const float4 a2 = *reinterpret_cast<const float4*>(V.farena + vf.pix[set][2] + at);
    p0[0] = a0.x; p0[1] = a0.y; p0[2] = a0.z; p0[3] = a0.w;
    p1[0] = a1.x; p1[1] = a1.y; p1[2] = a1.z; p1[3] = a1.w;
    p2[0] = a2.x; p2[1] = a2.y; p2[2] = a2.z; p2[3] = a2.w;
  }
#else
  for (uint32_t i = 0; i < 4; i++) {
    p0[i] = V.farena[vf.pix[set][0] + at + i];
    p1[i] = V.farena[vf.pix[set][1] + at + i];
    p2[i] = V.farena[vf.pix[set][2] + at + i];
  }
#endif
  DevColorStoreRgb8x4(V, vf, p0, p1, p2, x, y);
}

// ---------------------------------------------------------------- fused render tile
// Gaborish -> EPF 0 / 1 / 2 -> colour transform -> output samples for one kRtW x kRtH tile of a frame, with the
// intermediate planes in shared memory instead of HBM: the IDCT output of the tile plus a halo of 1 (Gaborish) + 3 + 2 + 1
// (the EPF stages that run) pixels is read once, every stage shrinks the valid region by its own radius, and the only
// store is the interleaved output. Stages are the same functions as the per-pixel kernels (DevGaborishValue,
// DevEpfValue, DevColorStore) on a view of the tile, so the samples are bit-identical. Mirroring happens when a stage
// reads (frame coordinates -> mirrored frame coordinates -> tile), as in the reference, so only in-frame pixels are
// ever computed. Frames with patches keep the per-pixel kernels (patches go between EPF and the colour transform).
constexpr int kRtW = 64, kRtH = 32, kRtMaxHalo = 7;
// Row stride of a tile in shared memory, a compile-time constant so that a stage's taps are the centre address plus
// immediates: 72 floats when the halo is at most 4 (Gaborish + two EPF stages), 80 otherwise.
constexpr int kRtStrideSmall = kRtW + 2 * 4, kRtStrideLarge = 80;

JXLB_HD uint32_t DevRenderHalo(uint32_t gab, uint32_t epf_iters) {
  return (gab ? 1u : 0u) + (epf_iters >= 3 ? 3u : 0u) + (epf_iters >= 1 ? 2u : 0u) + (epf_iters >= 2 ? 1u : 0u);
}
JXLB_HD bool DevRenderFused(const DevVFrame& vf) { return vf.patch_count == 0 && vf.upsampling <= 1; }
JXLB_HD uint32_t DevRenderStride(uint32_t halo) { return halo <= 4 ? kRtStrideSmall : kRtStrideLarge; }
// floats of one channel of one tile buffer for a batch whose largest halo is `halo`
JXLB_HD uint32_t DevRenderTileFloats(uint32_t halo) { return DevRenderStride(halo) * (kRtH + 2 * halo); }

template <bool INTERIOR, int STRIDE>
struct DevTileView {
  const float* p;  // tile sample (x0, y0) of the frame
  int x0, y0, xsize, ysize;
  JXLB_HD float At(int x, int y) const {
    if (!INTERIOR) {
      x = DevMirror(x, xsize);
      y = DevMirror(y, ysize);
    }
    return p[(y - y0) * STRIDE + (x - x0)];
  }
};

// i / d and i % d for i * ceil(2^20 / d) < 2^32 and i < 2^20 / d * ... (i < 4096, 64 <= d <= 80 here): exact, because the
// error of the rounded-up reciprocal, i / 2^20 < 1 / 256, stays below 1 / d.
struct DevSmallDiv {
  uint32_t d, m;
  JXLB_HD explicit DevSmallDiv(uint32_t div) : d(div), m(((1u << 20) + div - 1) / div) {}
  JXLB_HD uint32_t Div(uint32_t i) const { return (i * m) >> 20; }
};

// `sm`: 6 * cap floats (two sets of three channel tiles), rows STRIDE floats apart. (tx0, ty0): frame coordinates of
// the tile's first pixel.
template <int SCOPE, bool INTERIOR, int STRIDE>
JXLB_HD void DevRenderTile(const DevVPools& V, const DevVFrame& vf, int tx0, int ty0, uint32_t tid, uint32_t nt, float* sm,
                           uint32_t cap) {
  const int H = static_cast<int>(DevRenderHalo(vf.gab, vf.epf_iters));
  const int SW = kRtW + 2 * H, SH = kRtH + 2 * H, ox = tx0 - H, oy = ty0 - H;
  const int xsize = static_cast<int>(vf.xsize), ysize = static_cast<int>(vf.ysize);
  const uint32_t PW = vf.xblocks * 8;
  float* cur = sm;
  float* nxt = sm + 3 * static_cast<size_t>(cap);
  {
    const DevSmallDiv dv(static_cast<uint32_t>(SW));
    for (uint32_t i = tid; i < static_cast<uint32_t>(SW * SH); i += nt) {
      const uint32_t sy = dv.Div(i), sx = i - sy * dv.d;
      const int fx = ox + static_cast<int>(sx), fy = oy + static_cast<int>(sy);
      if (INTERIOR || (fx >= 0 && fx < xsize && fy >= 0 && fy < ysize)) {
        const size_t from = static_cast<size_t>(fy) * PW + fx;
        const uint32_t at = sy * STRIDE + sx;
        cur[at] = V.farena[vf.pix[0][0] + from];
        cur[cap + at] = V.farena[vf.pix[0][1] + from];
        cur[2 * cap + at] = V.farena[vf.pix[0][2] + from];
      }
    }
  }
  CoopSync<SCOPE>();
  int r = H;
  for (uint32_t step = 0; step < 4; step++) {  // step 0: Gaborish, 1..3: EPF stage step - 1
    const bool runs = step == 0 ? vf.gab != 0
                                : (vf.epf_iters > 0 && !(step == 1 && vf.epf_iters < 3) && !(step == 3 && vf.epf_iters < 2));
    if (!runs) continue;
    r -= step == 0 ? 1 : (step == 1 ? 3 : (step == 2 ? 2 : 1));
    const int rw = kRtW + 2 * r, rh = kRtH + 2 * r;
    DevTileView<INTERIOR, STRIDE> m[3];
    for (int c = 0; c < 3; c++) {
      m[c].p = cur + c * static_cast<size_t>(cap);
      m[c].x0 = ox;
      m[c].y0 = oy;
      m[c].xsize = xsize;
      m[c].ysize = ysize;
    }
    const DevSmallDiv dv(static_cast<uint32_t>(rw));
    for (uint32_t i = tid; i < static_cast<uint32_t>(rw * rh); i += nt) {
      const uint32_t ry = dv.Div(i), rx = i - ry * dv.d;
      const int fx = tx0 - r + static_cast<int>(rx), fy = ty0 - r + static_cast<int>(ry);
      if (!INTERIOR && !(fx >= 0 && fx < xsize && fy >= 0 && fy < ysize)) continue;
      const uint32_t at = static_cast<uint32_t>((fy - oy) * STRIDE + (fx - ox));
      if (step == 0) {
        nxt[at] = DevGaborishValue(m[0], vf.gab_w[0], fx, fy);
        nxt[cap + at] = DevGaborishValue(m[1], vf.gab_w[1], fx, fy);
        nxt[2 * cap + at] = DevGaborishValue(m[2], vf.gab_w[2], fx, fy);
      } else {
        float X, Y, B;
        DevEpfValue(m, vf, step - 1, DevEpfSigma(V, vf, fx, fy), fx, fy, &X, &Y, &B);
        nxt[at] = X;
        nxt[cap + at] = Y;
        nxt[2 * cap + at] = B;
      }
    }
    CoopSync<SCOPE>();
    float* t = cur;
    cur = nxt;
    nxt = t;
  }
  // colour transform + output samples of the kRtW x kRtH tile
  const bool x4 = vf.out_type == 2 && vf.out_channels == 3 && vf.out_stride % 4 == 0 && vf.orient == 0 && !vf.has_splines;
  for (uint32_t i = tid; i < static_cast<uint32_t>(kRtW / 4 * kRtH); i += nt) {
    const int lx = static_cast<int>(i % (kRtW / 4)) * 4, ly = static_cast<int>(i / (kRtW / 4));
    const int fx = tx0 + lx, fy = ty0 + ly;
    if (fy >= ysize || fx >= xsize) continue;
    const float* p = cur + (ly + H) * STRIDE + lx + H;
    if (x4 && fx + 4 <= xsize) {
      DevColorStoreRgb8x4(V, vf, p, p + cap, p + 2 * static_cast<size_t>(cap), static_cast<uint32_t>(fx), static_cast<uint32_t>(fy));
    } else {
      for (int k = 0; k < 4 && fx + k < xsize; k++)
        DevColorStore(V, vf, p[k], p[cap + k], p[2 * static_cast<size_t>(cap) + k], static_cast<uint32_t>(fx + k), static_cast<uint32_t>(fy));
    }
  }
}

}  // namespace jxlb

#endif  // JXLB_VARDCT_DEV_H_
