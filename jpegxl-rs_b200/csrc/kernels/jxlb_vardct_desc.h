// jxl_b200: plain-old-data descriptors of the VarDCT path, shared by the host planner
// and the CUDA kernels. One DevVFrame per VarDCT frame; every pointer-like field is an
// element offset into one of the batch arenas:
//   farena  float     DC planes, pixel planes (XYB), inverse sigma
//   barena  uint8_t   per-block side information (strategy, quant-DC bucket, EPF sharpness,
//                     raw quant (uint16), chroma-from-luma maps)
//   uarena  uint32_t  per-(pass, channel, block) token ranges
//   tokens  uint32_t  non-zero quantised coefficients as (position | value << 16)
// and into the read-only pools (fpool: dequantisation tables; opool: coefficient orders
// (uint16); cpool: context maps (uint8); upool: small integer tables).
#ifndef JXLB_VARDCT_DESC_H_
#define JXLB_VARDCT_DESC_H_

#include "jxlb_dev.h"

namespace jxlb {

constexpr uint32_t kMaxPasses = 11;
constexpr uint32_t kNumStrategies = 27;
constexpr uint32_t kNumOrders = 13;

struct DevVFrame {
  uint32_t xsize, ysize;      // pixels
  uint32_t xblocks, yblocks;  // 8x8 blocks (pixel planes are xblocks * 8 wide)
  uint32_t xgroups, ygroups;  // 256x256 groups
  uint32_t xdcgroups, ydcgroups;
  uint32_t cmw, cmh;          // 64x64 colour-correlation tiles
  uint32_t num_passes;
  uint32_t pass_shift[kMaxPasses];
  uint32_t skip_dc_smoothing, gab, epf_iters;
  // farena (float index)
  uint64_t dc[3];       // dequantised DC, xblocks * yblocks
  uint64_t dc_final[3]; // after adaptive smoothing (== dc when smoothing is skipped)
  uint64_t pix[2][3];   // two sets of pixel planes (filters ping-pong), (xblocks*8) * (yblocks*8)
  uint64_t inv_sigma;   // xblocks * yblocks
  // barena (byte index)
  uint64_t acs, qdc, sharp, rawq /* uint16, 2-aligned */, ytox, ytob;
  // uarena (uint32 index): [pass][channel][block] first token / token count
  uint64_t tok_start, tok_count;
  // uarena: per 256x256 group the list of its varblocks in decode order, two words each (DevBuildBlockList); the
  // list of the group with first block (x0, y0) and ys block rows starts at blist + 2 * (y0 * xblocks + x0 * ys).
  // blist_count: entries per group (0xFFFFFFFF: the strategy map of the group has a hole).
  uint64_t blist, blist_count;
  // quantiser + colour correlation
  float mul_dc[3], inv_mul_dc[3];
  float cfl_dc_x, cfl_dc_b;
  float inv_global_scale, global_scale_f;
  float x_dm, b_dm;
  float color_scale, base_x, base_b;
  float biases[4];
  // block context map (upool index): see DevBlockCtx layout in jxlb_vardct_dev.h
  uint32_t bctx_off;
  uint32_t num_ctxs, num_dc_ctxs, num_qf_thr, num_dc_thr[3];
  // dequantisation tables (fpool index), per quant table kind
  uint32_t table_off[17];
  // coefficient orders: upool index of num_passes * 39 opool offsets ([pass][3 * ord + c])
  uint32_t order_index;
  // AC entropy codes
  uint32_t ac_code[kMaxPasses];      // DevCode index per pass
  uint32_t ctx_map_off[kMaxPasses];  // cpool index per pass
  uint32_t num_histograms;
  // DC groups: upool index of num_dc_groups * 8 words: 7 plane ids (Y, X, B quantised DC;
  // ytox, ytob, strategy/quant rows, sharpness), extra_precision
  uint32_t dcg_index;
  // loop filter
  float gab_w[3][3];  // per channel: centre, side, corner weights (normalised)
  float epf_sigma_scale[3];   // per stage 0, 1, 2 (already times 1.65)
  float epf_border_sad_mul;
  float epf_channel_scale[3];
  float epf_quant_mul;
  float epf_sharp_lut[8];
  // colour
  uint32_t color_transform;  // 0 XYB, 1 none, 2 YCbCr
  float inv_mat[9], opsin_bias[3], opsin_bias_cbrt[3];
  uint32_t tf;          // 0 linear, 1 sRGB, 2 gamma (FastPowf with inv_gamma)
  float inv_gamma;
  // output
  uint32_t out_channels, out_type, out_big_endian;
  uint64_t out_off, out_stride;
  // patches (lib/jxl/dec_patch_dictionary.cc): DevPatch range, applied in order after the loop filters
  uint32_t patch_begin, patch_count;
  // frame upsampling (lib/jxl/render_pipeline/stage_upsampling.cc): 1, 2, 4 or 8; the filtered planes are upsampled
  // into `up_pix` (row stride up_stride) before the colour transform, which then runs at up_xsize x up_ysize
  uint32_t upsampling, up_xsize, up_ysize, up_stride;
  uint32_t up_kernel;  // fpool index of kernel[4][4][5][5]
  uint32_t orient;     // undo_orientation of the output store: 1 flip x, 2 flip y, 4 transpose (stage_write.cc:271-288)
  // alpha of a VarDCT frame with extra channels: the Modular plane (DevPlane index, kNoPlane: opaque) and its
  // int -> float factor (lib/jxl/dec_modular.cc:534-708)
  uint32_t alpha_plane;
  float alpha_factor;
  uint64_t up_pix[3];  // farena index
  // splines (lib/jxl/splines.cc; stage_splines.cc): added to the X, Y, B samples in front of the colour transform;
  // offsets into the batch's spline pools as DevFrameOut::spl_* (jxlb_dev.h)
  uint32_t has_splines, spl_pad_;
  uint64_t spl_rows, spl_idx, spl_seg;
};

// A reference-only frame (lib/jxl/frame_header.h kReferenceOnly, saved before the colour transform): its Modular
// sample planes become three float XYB planes in farena (ModularImageToDecodedRect, lib/jxl/dec_modular.cc:534-708:
// modular channel 0 is Y, 1 is X, 2 is B - Y; factor = DC quantisation step per channel).
struct DevRefFrame {
  uint32_t plane_y, plane_x, plane_b;  // DevPlane ids
  uint32_t w, h;
  float factor[3];                     // for X, Y, B
  uint64_t dst[3];                     // farena index of the X, Y, B planes (w * h each)
};

// One patch position (lib/jxl/dec_patch_dictionary.cc:28-200, blending of lib/jxl/blending.cc:42-160 for the three
// colour channels): the rectangle (x0, y0, xsize, ysize) of a reference frame goes to (x, y) of the frame.
struct DevPatch {
  uint64_t src[3];   // farena index of the reference planes
  uint32_t src_w;    // their row stride
  uint32_t x0, y0, xsize, ysize;
  uint32_t x, y;
  uint32_t mode;     // 0 none, 1 replace, 2 add, 3 multiply
  uint32_t clamp;
  uint32_t pad_;
};

// One AC entropy-coded stream = (frame, group, pass) = one thread of the AC decode kernel.
struct DevAcStream {
  uint64_t bit_pos, bit_end;  // absolute bit offsets in the byte pool
  uint32_t frame;             // DevVFrame index
  uint32_t group;
  uint32_t pass;
  uint32_t tok_cap;           // capacity in tokens
  uint64_t tok_off;           // first token (tokens index)
  uint32_t chain_slot;        // != 0: the position behind the last coefficient goes to chain_pos[chain_slot - 1] (DevStream::chain_slot)
  uint32_t pad_;
};

// The AC streams of one (frame, pass): what one CTA of k_ac_decode_frame decodes with the pass's alias tables, uint
// configs and context map staged in shared memory. `first` indexes the batch's list of stream ids (longest first).
struct DevAcUnit {
  uint32_t first, count;
  uint32_t frame, pass;
};

// AcStrategy geometry (lib/jxl/ac_strategy.h:32-80, :148-174), kStrategyOrder and
// kCoeffOrderOffset / 3 (lib/jxl/coeff_order.h:28-47), quant table kind (lib/jxl/quant_weights.h:343-353).
struct StrategyInfo {
  uint8_t cx, cy, log2_covered, order, table, plain_dct;
};

JXLB_HD StrategyInfo GetStrategyInfo(uint32_t s) {
  const uint8_t kCX[27] = {1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32};
  const uint8_t kCY[27] = {1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16};
  const uint8_t kLog2[27] = {0, 0, 0, 0, 2, 4, 1, 1, 2, 2, 3, 3, 0, 0, 0, 0, 0, 0, 6, 5, 5, 8, 7, 7, 10, 9, 9};
  const uint8_t kOrder[27] = {0, 1, 1, 1, 2, 3, 4, 4, 5, 5, 6, 6, 1, 1, 1, 1, 1, 1, 7, 8, 8, 9, 10, 10, 11, 12, 12};
  const uint8_t kTable[27] = {0, 1, 2, 3, 4, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10, 10, 11, 12, 12, 13, 14, 14, 15, 16, 16};
  StrategyInfo i;
  i.cx = kCX[s];
  i.cy = kCY[s];
  i.log2_covered = kLog2[s];
  i.order = kOrder[s];
  i.table = kTable[s];
  i.plain_dct = (s == 0 || (s >= 4 && s <= 11) || s >= 18) ? 1 : 0;
  return i;
}

JXLB_HD uint32_t PackStrategyInfo(const StrategyInfo& i) {
  return i.cx | (i.cy << 6) | (i.log2_covered << 12) | (i.order << 16) | (i.table << 20) | (i.plain_dct << 25);
}
JXLB_HD StrategyInfo UnpackStrategyInfo(uint32_t w) {
  StrategyInfo i;
  i.cx = w & 63;
  i.cy = (w >> 6) & 63;
  i.log2_covered = (w >> 12) & 15;
  i.order = (w >> 16) & 15;
  i.table = (w >> 20) & 31;
  i.plain_dct = (w >> 25) & 1;
  return i;
}

// Token encoding: position (index into the varblock's coefficient array) in the low 16
// bits, value in the high 16. Values outside int16 use the escape kTokEscape in the value
// field and carry the int32 value in the following word.
constexpr uint32_t kTokEscape = 0x8000u;

enum VStatus : uint32_t {
  kVOk = 0, kVOverread = 1, kVBadFinalState = 2, kVTokenOverflow = 4, kVBadStream = 8, kVUnsupported = 16
};

}  // namespace jxlb

#endif  // JXLB_VARDCT_DESC_H_
