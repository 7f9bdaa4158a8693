// jxl_b200 encoder: descriptors shared by the host side and the kernels.
#ifndef JXLB_ENC_DESC_H_
#define JXLB_ENC_DESC_H_

#include "jxlb_vardct_desc.h"

#if !defined(__CUDACC__)
// host build of the device functions (tests only): the CUDA vector type used for tokens
struct uint2 {
  unsigned int x, y;
};
inline uint2 make_uint2(unsigned int x, unsigned int y) {
  uint2 r;
  r.x = x;
  r.y = y;
  return r;
}
#endif

namespace jxlb {

// Node of the global Modular tree in the encoder's form: inner node: prop >= 0, a = split value,
// l / r = child indices (property > split ? l : r); leaf: prop = -1, a = predictor, l = context (leaf id).
struct DevEncTreeNode {
  int32_t prop;
  int32_t a;
  uint32_t l, r;
};

// Coefficient orders that may be customised (lib/jxl/enc_coeff_order.cc:47-76: orders 0..6, blocks up to 32x32):
// coefficients per order and where the zero counters of (order, channel) start: kCustomOrderBase[ord] + c * size.
constexpr uint32_t kNumCustomOrders = 7;
constexpr uint32_t kCustomOrderCounters = 6912;
JXLB_HD uint32_t CustomOrderSize(uint32_t ord) {
  const uint16_t k[7] = {64, 64, 256, 1024, 128, 256, 512};
  return k[ord];
}
JXLB_HD uint32_t CustomOrderBase(uint32_t ord) {
  const uint16_t k[7] = {0, 192, 384, 1152, 4224, 4608, 5376};
  return k[ord];
}

// One frame being encoded. Offsets index the encoder's arenas (element units of the arena's type).
struct DevEFrame {
  uint32_t xsize, ysize;
  uint32_t xblocks, yblocks;  // 8x8 blocks; planes are xblocks * 8 wide (edge-replicated padding)
  uint32_t xgroups, ygroups, xdcgroups, ydcgroups;
  uint32_t strategy_mode;
  float distance;
  float inv_global_scale, mul_dc[3], x_dm, b_dm;
  float x_qm_mul, b_qm_mul;  // encoder-side multipliers: pow(1.25, qm_scale - 2) (lib/jxl/enc_cache.cc:66-67)
  float biases[4];
  // input
  uint64_t rgb;        // byte arena: interleaved RGB8
  // float arena
  uint64_t xyb[3];      // the planes the transforms read
  uint64_t xyb_raw[3];  // gab != 0: XYB before the inverse Gaborish (k_enc_xyb writes here, k_enc_gaborish_inv reads)
  uint32_t gab;
  float gabinv_w[6];    // c, r, R, d, D, L of lib/jxl/convolve.h WeightsSymmetric5
  // chroma from luma: the forward transforms leave the float coefficients in the xyb_raw planes (free by then, each
  // varblock's coefficients in layout order inside its pixel footprint), the fit reads them, the quantisation reads both
  uint32_t cfl;         // 1: fit the factors per 64x64 tile (lib/jxl/enc_chroma_from_luma.cc), 0: all zero
  uint32_t cmw, cmh;    // tiles
  uint64_t ytox, ytob;  // byte arena: int8 per tile
  // adaptive quantisation (lib/jxl/enc_adaptive_quantization.cc; constants derived on the host, FillQuantizer)
  uint32_t adaptive;      // 1: quant field from k_enc_aq, raw quant per varblock; 0: raw quant 16 everywhere
  float aq_target;        // butteraugli target InitialQuantField is called with
  float aq_mul, aq_add;   // PerBlockModulations: scale * dampen, (1 - dampen) * base_level
  float aq_erosion[4];    // FuzzyErosion weights of the 4 smallest values
  float aq_mixer;         // AdjustQuantField: mean_max_mixer
  float cfl_scale128;     // Quantizer::Scale() * 128 (times the raw quant: the weighting of the CfL fit)
  uint64_t quant_field;   // float arena: xblocks * yblocks
  uint64_t raw_quant;     // byte arena: raw quant - 1 at the first block of every varblock
  uint64_t adj_thres;     // float arena: 4 planes of xblocks * yblocks: the Y dead-zone thresholds of the varblock (k_enc_adjust)
  uint64_t adj_quant;     // int arena (zeroed): the varblock's adjusted quant, max over the channels
  // int arena
  uint64_t coef[3];    // quantised coefficients, stored in each varblock's pixel footprint (row-major)
  uint64_t dcq[3];     // quantised DC: [0] = Y, [1] = X, [2] = B, xblocks * yblocks
  uint64_t first_index;  // per block: index of the varblock in its DC group's list (first blocks only)
  uint64_t block_of_num; // per DC group region: list position -> block position (y * xblocks + x)
  uint64_t dcg_count;    // per DC group: number of varblocks
  uint64_t group_tokens; // per AC group: number of tokens written
  uint64_t blk_nz[3], blk_ntok[3], blk_bucket[3];  // per block and channel: see DevEncBlockStats
  // coefficient-order statistics (read back by the host before the tokenisation kernels)
  uint64_t order_mask;   // one word: bit ord set when a varblock with that coefficient order exists
  uint64_t group_first;  // per AC group: number of varblocks
  uint64_t zero_counts;  // kCustomOrderCounters words: zero coefficients per (order, channel, position) over the sampled varblocks
  uint32_t custom_order[3 * kNumCustomOrders];  // [3 * ord + c]: offset into the custom order pool, 0xFFFFFFFF = natural
  uint64_t ac_hist;      // [num_ac_clusters][256]
  uint64_t mod_hist;     // [num_leaves][256]
  // byte arena
  uint64_t acs;        // strategy << 1 | is_first per block
  // token arena (uint2: context, value)
  uint64_t ac_tokens;  // group g at ac_tokens + 3 * 65536 * g
  uint64_t mod_tokens; // DC group g at mod_tokens + mod_tokens_stride * g
  uint64_t mod_tokens_stride;
  // batch bookkeeping
  uint32_t tree_off;   // first node of this frame's tree in the tree pool
  uint32_t sec_base;   // index of this frame's first section (DC groups, then AC groups) in the offset / length arrays
  uint64_t code_off[6];  // uint16 table pool: Modular code freq, start, reverse; AC code freq, start, reverse
  // An 8-bit alpha extra channel (input: interleaved RGBA8), coded losslessly in the frame's Modular sub-streams under the
  // global tree (lib/jxl/enc_modular.cc:1258-1500): group g's tokens at alpha_tokens + g * 65536, in the global stream
  // (group 0, stream id 0) when the image fits one group, else behind the coefficients of each AC group.
  uint32_t has_alpha, pad_;
  uint64_t alpha_tokens;
};

}  // namespace jxlb

#endif  // JXLB_ENC_DESC_H_
