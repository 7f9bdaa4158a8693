// jxl_b200: plain-old-data descriptors shared by the host planner and the CUDA
// kernels. Everything the device needs is expressed as offsets into a handful
// of flat pools that are uploaded once per batch:
//   bitstream bytes | alias entries | prefix tables | uint configs | tree nodes |
//   channel descriptors | stream descriptors | plane descriptors | transform ops.
#ifndef JXLB_DEV_H_
#define JXLB_DEV_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define JXLB_HD __host__ __device__ __forceinline__
#define JXLB_ALIGN(n) __align__(n)
#else
#define JXLB_HD inline
#define JXLB_ALIGN(n) alignas(n)
#endif

namespace jxlb {

// 8-byte alias-table entry, same field meaning as libjxl's AliasTable::Entry
// (lib/jxl/ans_common.h:67-76) so one 64-bit load fetches everything.
struct JXLB_ALIGN(8) DevAlias {
  uint8_t cutoff;
  uint8_t right_value;
  uint16_t freq0;
  uint16_t offsets1;
  uint16_t freq1_xor_freq0;
};

// One entropy code (histogram set). Offsets index the pools.
struct DevCode {
  uint32_t alias_off;    // DevAlias index of cluster 0
  uint32_t cfg_off;      // uint32 index: per-cluster packed HybridUintConfig
  uint32_t prefix_off;   // uint32 index: per-cluster table offsets (relative to prefix pool), then tables
  uint32_t ctx_map_off;  // byte index into ctx pool (context -> cluster); unused when leaves are pre-mapped
  uint32_t num_clusters;
  uint32_t log_alpha_size;
  uint32_t use_prefix;
  uint32_t lz77_enabled;
  uint32_t lz77_min_symbol;
  uint32_t lz77_min_length;
  uint32_t lz77_length_cfg;    // packed config
  uint32_t lz77_dist_cluster;  // cluster of the distance context
};

// Packed HybridUintConfig: split_exponent | msb << 8 | lsb << 16.
JXLB_HD uint32_t PackCfg(uint32_t split_exp, uint32_t msb, uint32_t lsb) {
  return split_exp | (msb << 8) | (lsb << 16);
}

// MA-tree node, 16 bytes. Inner node: prop >= 0, a = splitval, b = left child
// (property > splitval), c = right child. Leaf: prop = -1,
// a = cluster | predictor << 16, b = offset (int32), c = multiplier.
struct JXLB_ALIGN(16) DevTreeNode {
  int32_t prop;
  int32_t a;
  uint32_t b;
  uint32_t c;
};

struct DevPlane {
  uint64_t off;  // int32 index into the sample arena
  uint32_t w, h;
};

// One channel of one Modular stream, in decode order.
struct DevChannel {
  uint32_t plane;     // DevPlane index receiving the samples
  uint32_t prop0;     // value of property 0 (channel index inside the stream's image)
  uint32_t ref_off;   // index into the ref pool: plane ids of reference channels
  uint32_t ref_count;
  uint32_t tree_off;  // DevTreeNode index of this channel's tree: the stream's MA tree with the
                      // static properties 0 (channel) and 1 (stream id) already resolved
  uint32_t uses_wp;   // that pruned tree needs the weighted predictor
  // Chained sub-streams (a VarDCT DC group section holds the DC stream and the AC-metadata
  // stream back to back, lib/jxl/dec_frame.cc:315-339): when `preamble` is set the previous
  // entropy-coded stream ends before this channel and a new one begins: `count_bits` bits
  // (a value whose + 1 is the width of the `dyn` channel), a GroupHeader that must select the
  // global tree without transforms, and a fresh ANS state.
  uint32_t preamble;
  uint32_t count_bits;
  uint32_t dyn;       // 1: the width of this channel is the count read in the preamble; the plane is
                      // allocated at its upper bound and the count is stored after the plane's samples
  // Fast path (libjxl's fixed weighted-predictor trees, lib/jxl/modular/encoding/enc_encoding.cc:266-273, and
  // its own LUT fast path, lib/jxl/modular/encoding/encoding.h:70-131): when every inner node of the pruned
  // tree tests the weighted predictor's max-error property and every leaf is (Weighted, offset 0, multiplier 1),
  // lut[clamp(property, lut_lo, lut_lo + lut_size - 1) - lut_lo] is the leaf's cluster.
  uint32_t wp_lut;    // 1: use the LUT
  uint32_t lut_off;   // index into the lut pool (uint16)
  int32_t lut_lo;
  uint32_t lut_size;
  // Second fast path (libjxl's fixed AC-metadata tree, lib/jxl/modular/encoding/enc_encoding.cc:218-264): the pruned
  // tree only tests y, N and W with few thresholds, its leaves have offset 0 and multiplier 1 and no weighted
  // predictor: lut_off then points at a (y, N, W) bucket table (layout: jxlb_modular_dev.h, kNwOff*).
  // nw_lut == 2 (libjxl's fixed gradient DC tree, enc_encoding.cc:274-282): only property 9 (W + N - NW) is tested,
  // lut[clamp(property, lut_lo, lut_lo + lut_size - 1) - lut_lo] = cluster | predictor << 8.
  uint32_t nw_lut;
  // Warp-cooperative decode (kernels/jxlb_modular_coop_dev.h): `coop` != 0 when the channel has one of the table paths
  // above (1: every cluster of the table is on the list, 2: some are not); coop_list_off -> lut pool: count (<= 32), then the clusters the lanes speculate on (lane l takes entry l;
  // the ones nearest property value 0 when the channel uses more than 32); coop_lut_off -> a copy of the channel's
  // table (same layout) whose leaves are  lane | predictor << 8  when the cluster is on the list and
  // 0x8000 | cluster | predictor << 8  when it is not (decoded without speculation).
  uint32_t coop, coop_lut_off, coop_list_off;
};

// Layout of a (y, N, W) bucket table in the lut pool (uint16 units, DevChannel::lut_off): number of y thresholds;
// the thresholds (int32 as two uint16, ascending, padded with INT32_MAX) of y, N and W; then
// (ny + 1) x (kNwThresholds + 1) x (kNwThresholds + 1) leaves as cluster | predictor << 8.
constexpr uint32_t kNwThresholds = 4, kNwMaxY = 7;
constexpr uint32_t kNwOffY = 2, kNwOffN = kNwOffY + 2 * kNwMaxY, kNwOffW = kNwOffN + 2 * kNwThresholds,
                   kNwOffTable = kNwOffW + 2 * kNwThresholds;

// One Modular entropy-coded stream = one thread of the decode kernel.
struct DevStream {
  uint64_t bit_pos;      // absolute bit offset of the first symbol's ANS state in the byte pool
  uint64_t bit_end;      // absolute bit offset of the end of the section
  uint32_t code;         // DevCode index
  uint32_t tree_off;     // DevTreeNode index of the root
  uint32_t stream_id;    // property 1
  uint32_t chan_begin, chan_end;  // DevChannel range
  uint32_t dist_multiplier;       // LZ77 special distances
  uint32_t wp_params[3];          // p1C | p2C<<8 | p3Ca<<16 | p3Cb<<24 ; p3Cc | p3Cd<<8 | p3Ce<<16 ; w0|w1<<8|w2<<16|w3<<24
  uint32_t uses_wp;               // tree needs the weighted predictor
  uint32_t num_props;             // number of properties the tree reads (>= 16)
  uint32_t max_w;                 // widest channel (sizes the WP scratch)
  uint32_t scratch_slot;          // WP scratch slot
  uint32_t lz77_slot;             // LZ77 window slot or 0xFFFFFFFF
  // Streams that begin where an AC coefficient stream ends (the extra channels of a VarDCT frame: the Modular stream of
  // an AC group follows the group's coefficients bit by bit, lib/jxl/dec_frame.cc:478-560): chain_slot - 1 indexes the
  // positions the AC decode kernel leaves behind; the stream then starts with its GroupHeader (the first channel carries
  // `preamble`) and is decoded in the second Modular launch. 0: bit_pos is final.
  uint32_t chain_slot;
  uint32_t pad_;
};

enum DevOpKind : uint32_t {
  kOpRCT = 0,        // a,b,c = planes; p0 = rct_type
  kOpPalette = 1,    // a = index plane (= first output), b = palette plane, c = first extra output plane (nb-1 consecutive);
                     // p0 = nb, p1 = bit_depth, p2 = nb_deltas, p3 = predictor
  kOpCopy = 2,       // a = src plane, b = dst plane; p0 = dst x, p1 = dst y
  kOpHSqueeze = 3,   // a = avg plane, b = residual plane, c = out plane
  kOpVSqueeze = 4,
};

struct DevOp {
  uint32_t kind;
  uint32_t a, b, c;
  uint32_t p0, p1, p2, p3;
  uint32_t wp_params[3];
  uint32_t pad;
};

// One group's (or one frame's global) inverse-transform program.
struct DevProgram {
  uint32_t op_begin, op_end;
};

// Output conversion of one frame (lib/jxl/dec_modular.cc:534-708 int->float, then
// lib/jxl/render_pipeline/stage_write.cc:98-113 float->sample).
struct DevFrameOut {
  uint32_t xsize, ysize;
  uint32_t num_channels;   // of the caller's pixel format (1..4)
  uint32_t data_type;      // JxlDataType: 0 float, 2 u8, 3 u16, 5 f16
  uint32_t big_endian;
  uint32_t plane[4];       // source planes per output channel; 0xFFFFFFFF = constant 1.0 (opaque alpha)
  float factor[4];         // int -> float multiplier per output channel
  uint32_t is_float[4];    // source samples are custom floats: bits | exp_bits << 8 (0 = integer)
  uint64_t out_off;        // byte offset of this frame in the output buffer
  uint64_t stride;         // bytes per output row
  uint32_t vardct;         // 1: the frame is written by the VarDCT colour kernel, not by k_write_output
  uint32_t orient;         // undo_orientation of the output store: 1 flip x, 2 flip y, 4 transpose (stage_write.cc:271-288)
  // Splines (lib/jxl/splines.cc, render_pipeline/stage_splines.cc) of a Modular frame: added to the three colour samples
  // before the conversion. spl_rows -> uint32 pool: ysize + 1 offsets into the frame's index list, spl_idx -> the list
  // (segment numbers, per row in draw order), spl_seg -> float pool: kSplineSegmentWords words per segment.
  uint32_t has_splines, spl_pad_;
  uint64_t spl_rows, spl_idx, spl_seg;
};

// One spline segment in the float pool: centre, inverse sigma, sigma / 4 * intensity, the three colour multipliers, and
// (as int32 bits) the first and one-past-last column it touches (llround(centre -/+ maximum distance), splines.cc:99-103).
constexpr uint32_t kSplineSegmentWords = 10;

constexpr uint32_t kNoPlane = 0xFFFFFFFFu;

// Where the write stage puts source pixel (x, y) of a w x h image (stage_write.cc:163-172, :345-366): flipped position
// (fx, fy) -- which also indexes the dither pattern -- then row / column of the store.
JXLB_HD void DevOrient(uint32_t orient, uint32_t w, uint32_t h, uint32_t x, uint32_t y, uint32_t* fx, uint32_t* fy,
                       uint32_t* row, uint32_t* col) {
  *fx = (orient & 1) ? w - 1 - x : x;
  *fy = (orient & 2) ? h - 1 - y : y;
  *row = (orient & 4) ? *fx : *fy;
  *col = (orient & 4) ? *fy : *fx;
}

}  // namespace jxlb

#endif  // JXLB_DEV_H_
