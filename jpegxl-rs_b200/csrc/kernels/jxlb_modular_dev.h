// jxl_b200 device code: one Modular entropy-coded stream decoded by one thread.
//
// The per-sample loop of libjxl's DecodeModularChannelMAANS
// (lib/jxl/modular/encoding/encoding.cc:143-483, generic WP path) re-expressed for
// SIMT: every lane owns one (frame, group, pass) stream, walks the flat MA tree,
// pulls one rANS / prefix symbol (lib/jxl/dec_ans.h:168-195, :223-255, :286-343)
// and applies the predictor (lib/jxl/modular/encoding/context_predict.h:355-488,
// weighted predictor :134-214). A 256x256 group is a strictly serial chain, so the
// parallelism is #groups x #frames of the batch.
//
// The functions are __host__ __device__ so that tests/ can run the very same
// statements on the CPU when no GPU is present (logic check only; the product
// never executes them on the host).
#ifndef JXLB_MODULAR_DEV_H_
#define JXLB_MODULAR_DEV_H_

#include "jxlb_dev.h"

namespace jxlb {

struct DevPools {
  const uint32_t* words;  // bitstream pool viewed as little-endian 32-bit words
  const DevAlias* alias;
  const uint32_t* prefix;
  const uint32_t* cfg;
  const DevTreeNode* tree;
  const DevChannel* chans;
  const DevStream* streams;
  const DevPlane* planes;
  const uint32_t* refs;
  const uint16_t* lut;
  const DevCode* codes;
  int32_t* arena;
  int32_t* wp_scratch;   // per warp: 5 arrays x 2 rows x (wp_width + 2) x 32 lanes, [pos][lane]
  int32_t* ring;         // per warp: 3 rows x wp_width x 32 lanes, [x][lane]
  uint32_t wp_width;     // widest channel of the batch
  uint32_t* lz77;        // per slot: 1 << 20 entries
  uint32_t* status;      // per stream: 0 ok, else error bits
  uint64_t* end_bits;    // per stream: first bit after the stream (probe launches only, else null)
  const uint64_t* chain_pos;  // DevStream::chain_slot: where the AC decode kernel stopped reading (null without such streams)
  const float* spl_seg;       // splines of Modular frames: segments (kSplineSegmentWords each), row offsets + index lists
  const uint32_t* spl_idx;
  uint32_t num_streams;
  // streams [0, stream0) are decoded one per warp by k_modular_decode_coop (jxlb_modular_coop_dev.h), the lock-step
  // kernels take [stream0, num_streams): their warp 0 starts at stream0
  uint32_t stream0;
  uint32_t coop0;  // first stream of this launch's one-per-warp part: [coop0, stream0)
  // per warp of 32 consecutive streams: number of channel slots and, per slot, the
  // largest (width, height) among its lanes -- the warp-uniform loop bounds
  const uint32_t* warp_chans;
  const uint32_t* warp_dims_off;
  const uint32_t* warp_dims;
};

enum DevStatus : uint32_t { kStatusOk = 0, kStatusOverread = 1, kStatusBadFinalState = 2, kStatusUnsupported = 4,
                            kStatusNotRun = 0x80000000u };

#if defined(__CUDA_ARCH__)
#define JXLB_LDG(p) __ldg(p)
#define JXLB_WARP_ALL_M(p) __all_sync(0xFFFFFFFFu, (p))
#else
#define JXLB_LDG(p) (*(p))
#define JXLB_WARP_ALL_M(p) (p)
#endif

struct DevBits {
  const uint32_t* words;
  uint64_t next_word;
  uint64_t limit_word;  // first word that lies entirely behind the section: reads as zero from there on
  uint64_t buf;
  uint32_t n;

  // Words at or behind `bit_end` are never loaded (lib/jxl/dec_bit_reader.h:84-103, 213-225: libjxl's reader also
  // returns zeros past the end and reports the over-read at the end): a corrupt stream cannot walk out of the
  // byte pool, the callers' Pos() > bit_end test then flags it.
  JXLB_HD uint32_t Load(uint64_t w) const { return w < limit_word ? JXLB_LDG(words + w) : 0u; }
  JXLB_HD void Init(const uint32_t* w, uint64_t bit_pos, uint64_t bit_end) {
    words = w;
    next_word = bit_pos >> 5;
    limit_word = (bit_end + 31) >> 5;
    uint32_t sh = static_cast<uint32_t>(bit_pos & 31);
    buf = static_cast<uint64_t>(Load(next_word)) >> sh;
    next_word++;
    n = 32 - sh;
  }
  JXLB_HD void Fill() {
    if (n < 32) {
      buf |= static_cast<uint64_t>(Load(next_word)) << n;
      next_word++;
      n += 32;
    }
  }
  // k <= 32
  JXLB_HD uint32_t Read(uint32_t k) {
    Fill();
    uint32_t v = static_cast<uint32_t>(buf & ((uint64_t{1} << k) - 1));
    buf >>= k;
    n -= k;
    return v;
  }
  JXLB_HD uint32_t Peek(uint32_t k) {
    Fill();
    return static_cast<uint32_t>(buf & ((uint64_t{1} << k) - 1));
  }
  JXLB_HD void Skip(uint32_t k) {
    buf >>= k;
    n -= k;
  }
  JXLB_HD uint64_t Pos() const { return next_word * 32 - n; }
};

JXLB_HD uint32_t DevReadHybrid(uint32_t cfg, uint32_t token, DevBits& br) {
  const uint32_t split_exp = cfg & 0xFF, msb = (cfg >> 8) & 0xFF, lsb = (cfg >> 16) & 0xFF;
  const uint32_t split_token = 1u << split_exp;
  if (token < split_token) return token;
  const uint32_t in_token = msb + lsb;
  const uint32_t nbits = (split_exp - in_token + ((token - split_token) >> in_token)) & 31;
  const uint32_t low = token & ((1u << lsb) - 1);
  token >>= lsb;
  const uint32_t bits = br.Read(nbits);
  const uint32_t hi = (1u << msb) | (token & ((1u << msb) - 1));
  return (((hi << nbits) | bits) << lsb) | low;
}

constexpr uint32_t kDevLZ77Mask = (1u << 20) - 1;

struct DevSymbolReader {
  const DevAlias* alias;
  const uint32_t* cfg;
  const uint32_t* prefix;  // per-cluster offsets followed by tables
  uint32_t state;
  uint32_t log_alpha;
  uint32_t use_prefix;
  // LZ77
  uint32_t lz77_enabled, lz77_min_symbol, lz77_min_length, lz77_length_cfg, lz77_dist_cluster;
  uint32_t num_special, dist_mult;
  uint32_t num_to_copy;
  uint64_t copy_pos, num_decoded;
  uint32_t* window;

  // read_state = false: the stream begins with a GroupHeader (a chained stream's preamble reads the state afterwards)
  JXLB_HD void Init(const DevPools& P, const DevCode& c, DevBits& br, uint32_t dist_multiplier, uint32_t* win, bool read_state = true) {
    alias = P.alias + c.alias_off;
    cfg = P.cfg + c.cfg_off;
    prefix = P.prefix + c.prefix_off;
    log_alpha = c.log_alpha_size;
    use_prefix = c.use_prefix;
    state = (use_prefix || !read_state) ? (0x13u << 16) : br.Read(32);
    lz77_enabled = c.lz77_enabled;
    lz77_min_symbol = c.lz77_min_symbol;
    lz77_min_length = c.lz77_min_length;
    lz77_length_cfg = c.lz77_length_cfg;
    lz77_dist_cluster = c.lz77_dist_cluster;
    num_special = dist_multiplier == 0 ? 0 : 120;
    dist_mult = dist_multiplier;
    num_to_copy = 0;
    copy_pos = 0;
    num_decoded = 0;
    window = win;
  }

  JXLB_HD uint32_t ReadSymbol(uint32_t cluster, DevBits& br) {
    if (use_prefix) {
      const uint32_t* t = prefix + JXLB_LDG(prefix + cluster);
      br.Fill();
      uint32_t w = static_cast<uint32_t>(br.buf);
      uint32_t e = JXLB_LDG(t + (w & 0xFF));
      if (e & 0x80000000u) {
        uint32_t bits = (e >> 16) & 0xFF;
        e = JXLB_LDG(t + (e & 0xFFFF) + ((w >> 8) & ((1u << bits) - 1)));
      }
      br.Skip((e >> 16) & 0xFF);
      return e & 0xFFFF;
    }
    const uint32_t log_entry = 12 - log_alpha;
    const uint32_t res = state & 0xFFF;
    const uint32_t i = res >> log_entry;
    const uint32_t pos = res & ((1u << log_entry) - 1);
#if defined(__CUDA_ARCH__)
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(alias + (cluster << log_alpha) + i));
    const uint32_t cutoff = raw.x & 0xFF, right_value = (raw.x >> 8) & 0xFF, freq0 = raw.x >> 16;
    const uint32_t offsets1 = raw.y & 0xFFFF, fx = raw.y >> 16;
#else
    const DevAlias& e = alias[(cluster << log_alpha) + i];
    const uint32_t cutoff = e.cutoff, right_value = e.right_value, freq0 = e.freq0;
    const uint32_t offsets1 = e.offsets1, fx = e.freq1_xor_freq0;
#endif
    const bool right = pos >= cutoff;
    const uint32_t sym = right ? right_value : i;
    const uint32_t offset = (right ? offsets1 : 0u) + pos;
    const uint32_t freq = right ? (freq0 ^ fx) : freq0;
    state = freq * (state >> 12) + offset;
    if (state < (1u << 16)) state = (state << 16) | br.Read(16);
    return sym;
  }

  JXLB_HD uint32_t SpecialDistance(uint32_t i) const {
    // lib/jxl/dec_ans.h:121-143 kSpecialDistances, packed as (first + 8) | second << 5.
    const uint8_t k[120] = {
        0x28, 0x09, 0x29, 0x27, 0x48, 0x0A, 0x49, 0x47, 0x2A, 0x26, 0x4A, 0x46, 0x68, 0x0B, 0x69, 0x67, 0x2B, 0x25, 0x6A, 0x66,
        0x4B, 0x45, 0x88, 0x0C, 0x89, 0x87, 0x2C, 0x24, 0x6B, 0x65, 0x8A, 0x86, 0x4C, 0x44, 0xA8, 0x8B, 0x85, 0x6C, 0x64, 0x0D,
        0xA9, 0xA7, 0x2D, 0x23, 0xAA, 0xA6, 0x4D, 0x43, 0x8C, 0x84, 0xAB, 0xA5, 0x6D, 0x63, 0xC8, 0x0E, 0xC9, 0xC7, 0x2E, 0x22,
        0xCA, 0xC6, 0x4E, 0x42, 0xAC, 0xA4, 0x8D, 0x83, 0xCB, 0xC5, 0x6E, 0x62, 0xE8, 0x0F, 0xE9, 0xE7, 0xAD, 0xA3, 0x2F, 0x21,
        0xCC, 0xC4, 0x8E, 0x82, 0xEA, 0xE6, 0x4F, 0x41, 0xEB, 0xE5, 0x6F, 0x61, 0xCD, 0xC3, 0xAE, 0xA2, 0x10, 0xEC, 0xE4, 0x8F,
        0x81, 0x30, 0x50, 0xCE, 0xC2, 0x70, 0xED, 0xE3, 0xAF, 0xA1, 0x90, 0xEE, 0xE2, 0xCF, 0xC1, 0xB0, 0xEF, 0xE1, 0xD0, 0xF0};
    const int first = static_cast<int>(k[i] & 0x1F) - 8;
    const int second = static_cast<int>(k[i] >> 5);
    const int d = first + static_cast<int>(dist_mult) * second;
    return d > 1 ? static_cast<uint32_t>(d) : 1u;
  }

  JXLB_HD uint32_t CopyOne() {
    uint32_t v = window[(copy_pos++) & kDevLZ77Mask];
    num_to_copy--;
    window[(num_decoded++) & kDevLZ77Mask] = v;
    return v;
  }

  // The code is ANS without LZ77 (checked for the whole warp by the caller): no mode tests per symbol.
  // kShared: `alias` and `cfg` point into shared memory (k_ac_decode_frame stages the tables there).
  template <bool kShared = false>
  JXLB_HD uint32_t ReadUintPlainAns(uint32_t cluster, DevBits& br) {
    const uint32_t log_entry = 12 - log_alpha;
    const uint32_t res = state & 0xFFF;
    const uint32_t i = res >> log_entry;
    const uint32_t pos = res & ((1u << log_entry) - 1);
    const uint32_t cfg_word = kShared ? cfg[cluster] : JXLB_LDG(cfg + cluster);
#if defined(__CUDA_ARCH__)
    const uint2 raw = kShared ? *reinterpret_cast<const uint2*>(alias + (cluster << log_alpha) + i)
                              : __ldg(reinterpret_cast<const uint2*>(alias + (cluster << log_alpha) + i));
    const uint32_t cutoff = raw.x & 0xFF, right_value = (raw.x >> 8) & 0xFF, freq0 = raw.x >> 16;
    const uint32_t offsets1 = raw.y & 0xFFFF, fx = raw.y >> 16;
#else
    const DevAlias& e = alias[(cluster << log_alpha) + i];
    const uint32_t cutoff = e.cutoff, right_value = e.right_value, freq0 = e.freq0;
    const uint32_t offsets1 = e.offsets1, fx = e.freq1_xor_freq0;
#endif
    const bool right = pos >= cutoff;
    const uint32_t sym = right ? right_value : i;
    const uint32_t offset = (right ? offsets1 : 0u) + pos;
    const uint32_t freq = right ? (freq0 ^ fx) : freq0;
    state = freq * (state >> 12) + offset;
    if (state < (1u << 16)) state = (state << 16) | br.Read(16);
    return DevReadHybrid(cfg_word, sym, br);
  }

  JXLB_HD uint32_t ReadUint(uint32_t cluster, DevBits& br) {
    if (!lz77_enabled) return DevReadHybrid(JXLB_LDG(cfg + cluster), ReadSymbol(cluster, br), br);
    if (num_to_copy > 0) return CopyOne();
    uint32_t token = ReadSymbol(cluster, br);
    if (token >= lz77_min_symbol) {
      num_to_copy = DevReadHybrid(lz77_length_cfg, token - lz77_min_symbol, br) + lz77_min_length;
      uint32_t dtok = ReadSymbol(lz77_dist_cluster, br);
      uint64_t distance = DevReadHybrid(JXLB_LDG(cfg + lz77_dist_cluster), dtok, br);
      if (distance < num_special) {
        distance = SpecialDistance(static_cast<uint32_t>(distance));
      } else {
        distance = distance + 1 - num_special;
      }
      if (distance > num_decoded) distance = num_decoded;
      if (distance > (1u << 20)) distance = 1u << 20;
      copy_pos = num_decoded - distance;
      if (distance == 0) {
        uint32_t fill = num_to_copy < (1u << 20) ? num_to_copy : (1u << 20);
        for (uint32_t k = 0; k < fill; k++) window[k] = 0;
      }
      if (num_to_copy < lz77_min_length) return 0;
      return CopyOne();
    }
    uint32_t v = DevReadHybrid(JXLB_LDG(cfg + cluster), token, br);
    window[(num_decoded++) & kDevLZ77Mask] = v;
    return v;
  }
};

JXLB_HD DevTreeNode DevLoadNode(const DevTreeNode* p) {
#if defined(__CUDA_ARCH__)
  const int4 raw = __ldg(reinterpret_cast<const int4*>(p));
  DevTreeNode n;
  n.prop = raw.x;
  n.a = raw.y;
  n.b = static_cast<uint32_t>(raw.z);
  n.c = static_cast<uint32_t>(raw.w);
  return n;
#else
  return *p;
#endif
}

JXLB_HD int32_t DevUnpackSigned(uint32_t u) { return static_cast<int32_t>((u >> 1) ^ (~(u & 1) + 1)); }

JXLB_HD int32_t DevClampedGradient(int32_t n, int32_t w, int32_t l) {
  const int32_t m = n < w ? n : w, M = n < w ? w : n;
  const int32_t grad = static_cast<int32_t>(static_cast<uint32_t>(n) + static_cast<uint32_t>(w) - static_cast<uint32_t>(l));
  const int32_t g = l < m ? M : grad;
  return l > M ? m : g;
}

JXLB_HD int64_t DevAbs64(int64_t v) { return v < 0 ? -v : v; }

struct DevNeighbors {
  int64_t left, top, topleft, topright, leftleft, toptop, toprightright;
};

JXLB_HD DevNeighbors DevLoadNeighbors(const int32_t* row, const int32_t* prev, const int32_t* prevprev, int x, int y, int w) {
  DevNeighbors n;
  n.left = x ? row[x - 1] : (y ? prev[x] : 0);
  n.top = y ? prev[x] : n.left;
  n.topleft = (x && y) ? prev[x - 1] : n.left;
  n.topright = (x + 1 < w && y) ? prev[x + 1] : n.top;
  n.leftleft = x > 1 ? row[x - 2] : n.left;
  n.toptop = y > 1 ? prevprev[x] : n.top;
  n.toprightright = (x + 2 < w && y) ? prev[x + 2] : n.topright;
  return n;
}

JXLB_HD int64_t DevPredictOne(uint32_t p, const DevNeighbors& n, int64_t wp_pred) {
  switch (p) {
    case 0: return 0;
    case 1: return n.left;
    case 2: return n.top;
    case 3: return (n.left + n.top) / 2;
    case 4: {
      const int64_t pp = n.left + n.top - n.topleft;
      return DevAbs64(pp - n.left) < DevAbs64(pp - n.top) ? n.left : n.top;
    }
    case 5: return DevClampedGradient(static_cast<int32_t>(n.left), static_cast<int32_t>(n.top), static_cast<int32_t>(n.topleft));
    case 6: return wp_pred;
    case 7: return n.topright;
    case 8: return n.topleft;
    case 9: return n.leftleft;
    case 10: return (n.left + n.topleft) / 2;
    case 11: return (n.topleft + n.top) / 2;
    case 12: return (n.top + n.topright) / 2;
    case 13: return (6 * n.top - 2 * n.toptop + 7 * n.left + n.leftleft + n.toprightright + 3 * n.topright + 8) / 16;
    default: return 0;
  }
}

JXLB_HD uint32_t DevFloorLog2(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return 63 - __clzll(static_cast<long long>(v));
#else
  return 63 - __builtin_clzll(v);
#endif
}

JXLB_HD uint32_t DevFloorLog2_32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return 31 - __clz(static_cast<int>(v));
#else
  return 31 - __builtin_clz(v);
#endif
}

// Per-lane working memory of the decode kernel. One warp decodes 32 streams in lock
// step, so every array that is indexed by the position inside the row is laid out
// [position][lane]: the 32 lanes of a load/store instruction then touch 32 consecutive
// words (one 128-byte line) instead of 32 different lines.
struct DevLaneMem {
  int32_t* props;          // properties: props[p * props_stride]
  uint32_t props_stride;   // 32 (shared memory, bank = lane) on the device
  const uint32_t* divlut;  // 64-entry (1 << 24) / (i + 1)
  int32_t* ring;           // 3 rows x ring_w samples: ring[(r * ring_w + x) * lane_stride]
  int32_t* wp;             // 5 arrays x 2 rows x (ring_w + 2): wp[((a * 2 + r) * (ring_w + 2) + pos) * lane_stride]
  uint32_t ring_w;
  uint32_t lane_stride;
};

// Plain weighted predictor (lib/jxl/modular/encoding/context_predict.h:134-214) with
// the error arrays in memory; only the rare delta-palette inverse uses it. The decode
// kernel keeps the same state in registers (see DevDecodeModularStream).
struct DevWPPlain {
  int32_t p1C, p2C, p3Ca, p3Cb, p3Cc, p3Cd, p3Ce;
  uint32_t w[4];
  int64_t prediction[4];
  int64_t pred;
  int32_t* base;
  uint32_t row_len;
  const uint32_t* divlut;

  JXLB_HD void InitPlain(const uint32_t params[3], int32_t* scratch, uint32_t xsize, const uint32_t* lut) {
    p1C = params[0] & 0xFF; p2C = (params[0] >> 8) & 0xFF; p3Ca = (params[0] >> 16) & 0xFF; p3Cb = params[0] >> 24;
    p3Cc = params[1] & 0xFF; p3Cd = (params[1] >> 8) & 0xFF; p3Ce = (params[1] >> 16) & 0xFF;
    for (int i = 0; i < 4; i++) w[i] = (params[2] >> (8 * i)) & 0xFF;
    base = scratch;
    row_len = xsize + 2;
    divlut = lut;
    pred = 0;
    for (int i = 0; i < 4; i++) prediction[i] = 0;
  }
  JXLB_HD int32_t& At(uint32_t a, uint32_t r, uint32_t pos) const { return base[(a * 2 + r) * row_len + pos]; }
  JXLB_HD void Reset(uint32_t xsize) {
    for (uint32_t a = 0; a < 5; a++)
      for (uint32_t r = 0; r < 2; r++)
        for (uint32_t k = 0; k < xsize + 2; k++) At(a, r, k) = 0;
  }
  JXLB_HD uint32_t ErrorWeight(uint32_t x, uint32_t maxweight) const {
    int shift = static_cast<int>(DevFloorLog2(static_cast<uint64_t>(x) + 1)) - 5;
    if (shift < 0) shift = 0;
    return 4 + ((maxweight * divlut[x >> shift]) >> shift);
  }
  JXLB_HD int64_t Predict(uint32_t x, uint32_t y, uint32_t xsize, int64_t N, int64_t W, int64_t NE, int64_t NW, int64_t NN,
                          int32_t* max_error) {
    const uint32_t cur = (y & 1) ? 0 : 1, prev = cur ^ 1;
    const uint32_t pos_N = x;
    const uint32_t pos_NE = x < xsize - 1 ? x + 1 : x;
    const uint32_t pos_NW = x > 0 ? x - 1 : x;
    uint32_t weights[4];
    for (uint32_t i = 0; i < 4; i++) {
      const uint32_t e = static_cast<uint32_t>(At(i, prev, pos_N)) + static_cast<uint32_t>(At(i, prev, pos_NE)) +
                         static_cast<uint32_t>(At(i, prev, pos_NW));
      weights[i] = ErrorWeight(e, w[i]);
    }
    N *= 8; W *= 8; NE *= 8; NW *= 8; NN *= 8;
    const int64_t teW = x == 0 ? 0 : At(4, cur, x - 1);
    const int64_t teN = At(4, prev, pos_N), teNW = At(4, prev, pos_NW), teNE = At(4, prev, pos_NE);
    const int64_t sumWN = teN + teW;
    if (max_error) {
      int64_t p = teW;
      if (DevAbs64(teN) > DevAbs64(p)) p = teN;
      if (DevAbs64(teNW) > DevAbs64(p)) p = teNW;
      if (DevAbs64(teNE) > DevAbs64(p)) p = teNE;
      *max_error = static_cast<int32_t>(p);
    }
    prediction[0] = W + NE - N;
    prediction[1] = N - (((sumWN + teNE) * p1C) >> 5);
    prediction[2] = W - (((sumWN + teNW) * p2C) >> 5);
    prediction[3] = N - ((teNW * p3Ca + teN * p3Cb + teNE * p3Cc + (NN - N) * p3Cd + (NW - W) * p3Ce) >> 5);
    uint32_t wsum = weights[0] + weights[1] + weights[2] + weights[3];
    const uint32_t log_weight = DevFloorLog2(wsum);
    wsum = 0;
    for (int i = 0; i < 4; i++) {
      weights[i] >>= log_weight - 4;
      wsum += weights[i];
    }
    int64_t sum = (wsum >> 1) - 1;
    for (int i = 0; i < 4; i++) sum += prediction[i] * weights[i];
    pred = (sum * divlut[wsum - 1]) >> 24;
    if (((teN ^ teW) | (teN ^ teNW)) > 0) return (pred + 3) >> 3;
    int64_t mx = W > NE ? W : NE;
    if (N > mx) mx = N;
    int64_t mn = W < NE ? W : NE;
    if (N < mn) mn = N;
    if (pred > mx) pred = mx;
    if (pred < mn) pred = mn;
    return (pred + 3) >> 3;
  }
  JXLB_HD void Update(int64_t val, uint32_t x, uint32_t y) {
    const uint32_t cur = (y & 1) ? 0 : 1, prev = cur ^ 1;
    val *= 8;
    At(4, cur, x) = static_cast<int32_t>(pred - val);
    for (uint32_t i = 0; i < 4; i++) {
      const uint32_t err = static_cast<uint32_t>((DevAbs64(prediction[i] - val) + 3) >> 3);
      At(i, cur, x) = static_cast<int32_t>(err);
      At(i, prev, x + 1) = static_cast<int32_t>(static_cast<uint32_t>(At(i, prev, x + 1)) + err);
    }
  }
};

constexpr int kDevMaxProps = 16 + 4 * 8;  // static 2 + 13 + WP + up to 8 reference channels

template <typename WT>
JXLB_HD WT DevAbsW(WT v) { return v < 0 ? -v : v; }

template <typename WT>
JXLB_HD WT DevPredictW(uint32_t p, WT left, WT top, WT topleft, WT topright, WT leftleft, WT toptop, WT toprightright,
                       WT wp_pred) {
  switch (p) {
    case 0: return 0;
    case 1: return left;
    case 2: return top;
    case 3: return (left + top) / 2;
    case 4: {
      const WT pp = left + top - topleft;
      return DevAbsW<WT>(pp - left) < DevAbsW<WT>(pp - top) ? left : top;
    }
    case 5: return DevClampedGradient(static_cast<int32_t>(left), static_cast<int32_t>(top), static_cast<int32_t>(topleft));
    case 6: return wp_pred;
    case 7: return topright;
    case 8: return topleft;
    case 9: return leftleft;
    case 10: return (left + topleft) / 2;
    case 11: return (topleft + top) / 2;
    case 12: return (top + topright) / 2;
    case 13: return (6 * top - 2 * toptop + 7 * left + leftleft + toprightright + 3 * topright + 8) / 16;
    default: return 0;
  }
}

JXLB_HD int32_t DevNwThreshold(const uint16_t* base, uint32_t i) {
  return static_cast<int32_t>(static_cast<uint32_t>(JXLB_LDG(base + 2 * i)) | (static_cast<uint32_t>(JXLB_LDG(base + 2 * i + 1)) << 16));
}

struct DevWpParams {
  int32_t p1C, p2C, p3Ca, p3Cb, p3Cc, p3Cd, p3Ce;
  uint32_t w[4];
};

// The sample loops of one channel slot of the warp (see DevDecodeModularStream). kFast: every lane that has this
// slot decodes it through the weighted-predictor LUT -- no properties, no tree walk, no reference channels; the
// generic code is compiled out of that instance.
//
// Row storage (per lane, `[position][lane]`): the lane's row ring is two rows, A and B. A holds the previous row and
// is overwritten in place by the current one (the previous row's samples around x live in the sliding registers
// tl / t / tr / trr, loaded three positions ahead of the write), B receives the previous row's sample just before A
// loses it and so holds the row above the previous one when the next row reads it. The weighted predictor's five
// error rows are updated in place the same way (entries x - 1 .. x + 1 of the previous row are in registers).
//
// kNw: every lane that has this slot decodes it through the (y, N, W) bucket table (DevChannel::nw_lut): the pruned
// tree only tests the row, the sample above and the sample to the left -- libjxl's fixed AC-metadata tree
// (lib/jxl/modular/encoding/enc_encoding.cc:218-264) -- so the leaf is table[bucket(y)][bucket(N)][bucket(W)], a bucket
// being the number of the property's thresholds below the value: a handful of independent compares and one load
// instead of the property stores and the dependent node loads of the tree walk; no weighted predictor, no references.
template <typename WT, bool kFast, uint32_t kLS, bool kPlainAns, bool kNw = false>
JXLB_HD void DevDecodeChannelRows(const DevPools& P, const DevLaneMem& m, const DevChannel& ch, const DevWpParams& wpp,
                                  const int w, const int h, const int max_w, const int max_h, const uint32_t stride,
                                  int32_t* out, const bool lane_direct, const bool lane_uses_wp, const DevTreeNode* tree,
                                  DevSymbolReader& reader, DevBits& br) {
  const bool direct = kFast ? false : lane_direct;  // (the fast path never reads neighbours from the output plane)
  // kLS: the lane stride when the kernel fixes it at compile time (index arithmetic becomes shifts), 0 = m.lane_stride
  const uint32_t PS = m.props_stride, LS = kLS ? kLS : m.lane_stride, RW = m.ring_w, WL = m.ring_w + 2;
  int32_t* props = m.props;
  const uint32_t* divlut = m.divlut;
  const bool uses_wp = kFast ? true : (kNw ? false : lane_uses_wp);
  // The first two levels of the tree are walked for every sample: keep them in registers.
  DevTreeNode root{}, root_l{}, root_r{};
  root.prop = -1;
  if (!kFast && !kNw && h > 0) {
    root = DevLoadNode(tree);
    if (root.prop >= 0) {
      root_l = DevLoadNode(tree + root.b);
      root_r = DevLoadNode(tree + root.c);
    }
  }
  const uint32_t RS = direct ? 1 : LS;
  const uint16_t* lut = P.lut + ch.lut_off;
  const int32_t lut_lo = ch.lut_lo, lut_hi = ch.lut_lo + static_cast<int32_t>(ch.lut_size) - 1;
  // kNw: thresholds of N and W in registers (padded with INT32_MAX: never exceeded), the row's table slice per row
  int32_t nw_tn[kNwThresholds] = {}, nw_tw[kNwThresholds] = {};
  uint32_t nw_ny = 0;
  // nw_lut == 2: the tree only tests property 9 (W + N - NW, libjxl's fixed gradient DC tree): lut[clamp(p9) - lut_lo]
  const bool grad_lut = kNw && ch.nw_lut == 2;
  if (kNw && h > 0 && !grad_lut) {
    nw_ny = JXLB_LDG(lut);
    for (uint32_t i = 0; i < kNwThresholds; i++) {
      nw_tn[i] = DevNwThreshold(lut + kNwOffN, i);
      nw_tw[i] = DevNwThreshold(lut + kNwOffW, i);
    }
  }
  const uint16_t* nw_row = lut;
  int32_t* pe[4];
  for (uint32_t i = 0; i < 4; i++) pe[i] = m.wp + static_cast<uint32_t>(i * WL) * LS;
  int32_t* er = m.wp + static_cast<uint32_t>(4 * WL) * LS;
  int32_t* rowA = m.ring;
  int32_t* rowB = m.ring + static_cast<uint32_t>(RW) * LS;
  if (uses_wp) {  // the rows above y = 0 read as zero
    for (uint32_t a = 0; a < 5; a++)
      for (int q = 0; q < w + 2; q++) m.wp[static_cast<uint32_t>(a * WL + q) * LS] = 0;
  }
  for (int y = 0; y < max_h; y++) {
    const bool row_on = y < h;
    int32_t* out_row = out + static_cast<size_t>(y) * stride;
    int32_t* row = direct ? out_row : rowA;
    const int32_t* prev = direct ? out_row - stride : rowA;
    const int32_t* prevprev = direct ? out_row - 2 * static_cast<size_t>(stride) : rowB;
    if (!kFast && !kNw) props[2 * PS] = y;
    if (kNw && row_on && !grad_lut) {
      uint32_t by = 0;
      for (uint32_t i = 0; i < nw_ny; i++) by += y > DevNwThreshold(lut + kNwOffY, i) ? 1u : 0u;
      nw_row = lut + kNwOffTable + by * ((kNwThresholds + 1) * (kNwThresholds + 1));
    }
    int32_t prev_grad = 0;  // property 9 of the previous pixel
    // sliding neighbourhood (valid when y > 0): t = prev[x], tl = prev[x-1], tr = prev[x+1], trr = prev[x+2]
    int32_t left = 0, leftleft = 0, t = 0, tl = 0, tr = 0, trr = 0;
    uint32_t eNW[4] = {0, 0, 0, 0}, eN[4] = {0, 0, 0, 0}, eNE[4] = {0, 0, 0, 0};
    int32_t teW = 0, teNW = 0, teN = 0, teNE = 0;
    if (row_on && w > 0) {
      if (y > 0) {
        t = prev[0];
        tr = w > 1 ? prev[static_cast<uint32_t>(1) * RS] : t;
        trr = w > 2 ? prev[static_cast<uint32_t>(2) * RS] : tr;
      }
      if (uses_wp) {
        for (uint32_t i = 0; i < 4; i++) {
          eN[i] = static_cast<uint32_t>(pe[i][0]);
          eNW[i] = eN[i];
          eNE[i] = w > 1 ? static_cast<uint32_t>(pe[i][static_cast<uint32_t>(1) * LS]) : eN[i];
        }
        teN = er[0];
        teNW = teN;
        teNE = w > 1 ? er[static_cast<uint32_t>(1) * LS] : teN;
      }
    }
    for (int x = 0; x < max_w; x++) {
      if (row_on && x < w) {
        // neighbours with the edge rules of context_predict.h:496-504
        const WT n_left = x ? left : (y ? t : 0);
        const WT n_top = y ? t : n_left;
        const WT n_topleft = (x && y) ? tl : n_left;
        const WT n_topright = (x + 1 < w && y) ? tr : n_top;
        const WT n_leftleft = x > 1 ? leftleft : n_left;
        const WT n_toptop = y > 1 ? prevprev[static_cast<uint32_t>(x) * RS] : n_top;
        const WT n_toprightright = (x + 2 < w && y) ? trr : n_topright;
        if (!kFast && !kNw) {
          props[3 * PS] = x;
          props[4 * PS] = static_cast<int32_t>(n_top > 0 ? n_top : -n_top);
          props[5 * PS] = static_cast<int32_t>(n_left > 0 ? n_left : -n_left);
          props[6 * PS] = static_cast<int32_t>(n_top);
          props[7 * PS] = static_cast<int32_t>(n_left);
          props[8 * PS] = static_cast<int32_t>(n_left - prev_grad);
          prev_grad = static_cast<int32_t>(n_left + n_top - n_topleft);
          props[9 * PS] = prev_grad;
          props[10 * PS] = static_cast<int32_t>(n_left - n_topleft);
          props[11 * PS] = static_cast<int32_t>(n_topleft - n_top);
          props[12 * PS] = static_cast<int32_t>(n_top - n_topright);
          props[13 * PS] = static_cast<int32_t>(n_top - n_toptop);
          props[14 * PS] = static_cast<int32_t>(n_left - n_leftleft);
        }
        int32_t wp_max_error = 0;
        WT wp_pred = 0, wp_raw = 0;
        WT prediction[4] = {0, 0, 0, 0};
        if (uses_wp) {
          uint32_t weights[4];
          for (uint32_t i = 0; i < 4; i++) {
            const uint32_t e = eN[i] + eNE[i] + eNW[i];
            // (32-bit sums cannot wrap when 16-bit buffers suffice: every error term is below 2^20)
            int shift = static_cast<int>(sizeof(WT) == 4 ? DevFloorLog2_32(e + 1) : DevFloorLog2(static_cast<uint64_t>(e) + 1)) - 5;
            if (shift < 0) shift = 0;
            weights[i] = 4 + ((wpp.w[i] * divlut[e >> shift]) >> shift);
          }
          const WT N8 = n_top * 8, W8 = n_left * 8, NE8 = n_topright * 8, NW8 = n_topleft * 8, NN8 = n_toptop * 8;
          const WT eW = x == 0 ? 0 : teW;
          const WT sumWN = static_cast<WT>(teN) + eW;
          {
            WT pm = eW;
            if (DevAbsW<WT>(teN) > DevAbsW<WT>(pm)) pm = teN;
            if (DevAbsW<WT>(teNW) > DevAbsW<WT>(pm)) pm = teNW;
            if (DevAbsW<WT>(teNE) > DevAbsW<WT>(pm)) pm = teNE;
            wp_max_error = static_cast<int32_t>(pm);
            if (!kFast) props[15 * PS] = wp_max_error;
          }
          prediction[0] = W8 + NE8 - N8;
          prediction[1] = N8 - (((sumWN + teNE) * wpp.p1C) >> 5);
          prediction[2] = W8 - (((sumWN + teNW) * wpp.p2C) >> 5);
          prediction[3] = N8 - ((static_cast<WT>(teNW) * wpp.p3Ca + static_cast<WT>(teN) * wpp.p3Cb +
                                 static_cast<WT>(teNE) * wpp.p3Cc + (NN8 - N8) * wpp.p3Cd + (NW8 - W8) * wpp.p3Ce) >> 5);
          uint32_t wsum = weights[0] + weights[1] + weights[2] + weights[3];
          const uint32_t log_weight = DevFloorLog2_32(wsum);
          wsum = 0;
          for (int i = 0; i < 4; i++) {
            weights[i] >>= log_weight - 4;
            wsum += weights[i];
          }
          WT sum = static_cast<WT>((wsum >> 1) - 1);
          for (int i = 0; i < 4; i++) sum += prediction[i] * static_cast<WT>(weights[i]);
          wp_raw = static_cast<WT>((static_cast<int64_t>(sum) * static_cast<int64_t>(divlut[wsum - 1])) >> 24);
          if (!(((static_cast<WT>(teN) ^ eW) | (static_cast<WT>(teN) ^ static_cast<WT>(teNW))) > 0)) {
            WT mx = W8 > NE8 ? W8 : NE8;
            if (N8 > mx) mx = N8;
            WT mn = W8 < NE8 ? W8 : NE8;
            if (N8 < mn) mn = N8;
            if (wp_raw > mx) wp_raw = mx;
            if (wp_raw < mn) wp_raw = mn;
          }
          wp_pred = (wp_raw + 3) >> 3;
        }
        for (uint32_t r = 0; !kFast && !kNw && r < ch.ref_count; r++) {
          const DevPlane rp = P.planes[P.refs[ch.ref_off + r]];
          const int32_t* rrow = P.arena + rp.off + static_cast<size_t>(y) * w;
          const int32_t* rprev = y ? rrow - w : rrow;
          const int64_t v = rrow[x];
          const int64_t vleft = x ? rrow[x - 1] : 0;
          const int64_t vtop = y ? rprev[x] : vleft;
          const int64_t vtopleft = (x && y) ? rprev[x - 1] : vleft;
          const int64_t vpred = DevClampedGradient(static_cast<int32_t>(vleft), static_cast<int32_t>(vtop), static_cast<int32_t>(vtopleft));
          props[(16 + 4 * r + 0) * PS] = static_cast<int32_t>(DevAbs64(v));
          props[(16 + 4 * r + 1) * PS] = static_cast<int32_t>(v);
          props[(16 + 4 * r + 2) * PS] = static_cast<int32_t>(DevAbs64(v - vpred));
          props[(16 + 4 * r + 3) * PS] = static_cast<int32_t>(v - vpred);
        }
        int32_t val;
        if (kFast) {
          // single-property tree on the max-error property, leaves (Weighted, 0, 1): one table lookup
          const int32_t pv = wp_max_error < lut_lo ? lut_lo : (wp_max_error > lut_hi ? lut_hi : wp_max_error);
          const uint32_t cluster = JXLB_LDG(lut + (pv - lut_lo));
          const uint32_t u = kPlainAns ? reader.ReadUintPlainAns(cluster, br) : reader.ReadUint(cluster, br);
          val = static_cast<int32_t>(static_cast<uint32_t>(DevUnpackSigned(u)) + static_cast<uint32_t>(wp_pred));
        } else if (kNw) {
          uint32_t e;  // cluster | predictor << 8
          if (grad_lut) {
            const WT p9 = n_left + n_top - n_topleft;
            const WT pv = p9 < static_cast<WT>(lut_lo) ? static_cast<WT>(lut_lo) : (p9 > static_cast<WT>(lut_hi) ? static_cast<WT>(lut_hi) : p9);
            e = JXLB_LDG(lut + static_cast<uint32_t>(static_cast<int32_t>(pv) - lut_lo));
          } else {
            uint32_t bn = 0, bw = 0;
            for (uint32_t i = 0; i < kNwThresholds; i++) {
              bn += n_top > static_cast<WT>(nw_tn[i]) ? 1u : 0u;
              bw += n_left > static_cast<WT>(nw_tw[i]) ? 1u : 0u;
            }
            e = JXLB_LDG(nw_row + bn * (kNwThresholds + 1) + bw);
          }
          const uint32_t u = kPlainAns ? reader.ReadUintPlainAns(e & 0xFF, br) : reader.ReadUint(e & 0xFF, br);
          const WT guess = DevPredictW<WT>(e >> 8, n_left, n_top, n_topleft, n_topright, n_leftleft, n_toptop,
                                           n_toprightright, 0);
          val = static_cast<int32_t>(static_cast<uint32_t>(DevUnpackSigned(u)) + static_cast<uint32_t>(guess));
        } else {
          DevTreeNode node = root;
          if (node.prop >= 0) node = props[node.prop * PS] > node.a ? root_l : root_r;
          while (node.prop >= 0) {
            const uint32_t pos = props[node.prop * PS] > node.a ? node.b : node.c;
            node = DevLoadNode(tree + pos);
          }
          const uint32_t cluster = static_cast<uint32_t>(node.a) & 0xFFFF;
          const uint32_t predictor = static_cast<uint32_t>(node.a) >> 16;
          const uint32_t u = kPlainAns ? reader.ReadUintPlainAns(cluster, br) : reader.ReadUint(cluster, br);
          const WT guess = static_cast<WT>(static_cast<int32_t>(node.b)) +
                           DevPredictW<WT>(predictor, n_left, n_top, n_topleft, n_topright, n_leftleft, n_toptop,
                                           n_toprightright, wp_pred);
          // low 32 bits of (unpacked * multiplier + guess), as in make_pixel (encoding.cc:168-173)
          val = static_cast<int32_t>(static_cast<uint32_t>(DevUnpackSigned(u)) * node.c + static_cast<uint32_t>(guess));
        }
        if (!direct) rowB[static_cast<uint32_t>(x) * LS] = t;  // the sample A loses below: next row's N-N
        row[static_cast<uint32_t>(x) * RS] = val;
        out_row[x] = val;
        leftleft = left;
        left = val;
        // slide the previous-row window
        tl = t;
        t = tr;
        tr = trr;
        if (y > 0 && x + 3 < w) trr = prev[static_cast<uint32_t>(x + 3) * RS];
        if (uses_wp) {
          const WT val8 = static_cast<WT>(val) * 8;
          const int32_t te = static_cast<int32_t>(wp_raw - val8);
          const bool more = x + 2 < w;  // position x + 2 exists in the previous row
          const int32_t er_ahead = more ? er[static_cast<uint32_t>(x + 2) * LS] : 0;
          er[static_cast<uint32_t>(x) * LS] = te;
          for (uint32_t i = 0; i < 4; i++) {
            const uint32_t err = static_cast<uint32_t>((DevAbsW<WT>(prediction[i] - val8) + 3) >> 3);
            // next pixel: NW <- N, N <- (entry x + 1) + err, NE <- entry x + 2 (or N at the row end)
            const uint32_t n_next = eNE[i] + err;
            eNW[i] = eN[i];
            eN[i] = n_next;
            eNE[i] = more ? static_cast<uint32_t>(pe[i][static_cast<uint32_t>(x + 2) * LS]) : n_next;
            pe[i][static_cast<uint32_t>(x) * LS] = static_cast<int32_t>(err);
          }
          teW = te;
          teNW = teN;
          teN = teNE;
          teNE = more ? er_ahead : teN;
        }
      }
    }
  }
}

// Decodes every channel of stream `s` (lane `s % 32` of warp `s / 32`).
//
// WT is the arithmetic width of predictor math: int32_t when the codestream promises
// that 16-bit buffers suffice (ImageMetadata::modular_16_bit_buffer_sufficient,
// lib/jxl/image_metadata.cc:322) -- every intermediate of the reference's int64
// formulas then fits 32 bits except the final 28 x 24 bit product of the weighted
// average, which stays 64-bit -- and int64_t otherwise.
//
// All 32 lanes run the same (channel, y, x) loop nest with warp-uniform bounds
// (warp_dims: per channel slot the largest width / height of the warp) and mask
// themselves off outside their own plane: the warp never splits.
//
// The weighted predictor's error rows are kept as sliding NW/N/NE registers; the
// reference's `pred_errors[prev_row + x + 1] += err` (context_predict.h:209) only
// ever influences the next two pixels of the same row, so it is applied to the
// registers and never stored.
template <typename WT, uint32_t kLS = 0>
JXLB_HD uint32_t DevDecodeModularStream(const DevPools& P, uint32_t s, const DevLaneMem& m, const uint32_t* warp_dims,
                                        uint32_t warp_chans, bool lane_valid, uint64_t* end_pos = nullptr) {
  DevStream st{};
  DevCode code{};
  DevBits br{};
  DevSymbolReader reader{};
  uint32_t my_chans = 0;
  if (lane_valid) {
    st = P.streams[s];
    code = P.codes[st.code];
    const bool chained = st.chain_slot != 0;  // starts (with its GroupHeader) where an AC coefficient stream ended
    br.Init(P.words, chained ? P.chain_pos[st.chain_slot - 1] : st.bit_pos, st.bit_end);
    uint32_t* window = (st.lz77_slot != 0xFFFFFFFFu) ? P.lz77 + (static_cast<size_t>(st.lz77_slot) << 20) : nullptr;
    reader.Init(P, code, br, st.dist_multiplier, window, /*read_state=*/!chained);
    my_chans = st.chan_end - st.chan_begin;
  }
  // every lane of the warp reads plain ANS (no prefix codes, no LZ77): the per-symbol mode tests are compiled out
  const bool plain_ans = JXLB_WARP_ALL_M(!lane_valid || (!code.use_prefix && !code.lz77_enabled));
  const uint32_t PS = m.props_stride;
  int32_t* props = m.props;
  for (int i = 0; i < kDevMaxProps; i++) props[i * PS] = 0;
  props[1 * PS] = static_cast<int32_t>(st.stream_id);
  // weighted predictor parameters
  DevWpParams wpp;
  wpp.p1C = st.wp_params[0] & 0xFF; wpp.p2C = (st.wp_params[0] >> 8) & 0xFF; wpp.p3Ca = (st.wp_params[0] >> 16) & 0xFF;
  wpp.p3Cb = st.wp_params[0] >> 24; wpp.p3Cc = st.wp_params[1] & 0xFF; wpp.p3Cd = (st.wp_params[1] >> 8) & 0xFF;
  wpp.p3Ce = (st.wp_params[1] >> 16) & 0xFF;
  uint32_t status = kStatusOk;
  uint32_t dyn_count = 0;
  for (int i = 0; i < 4; i++) wpp.w[i] = (st.wp_params[2] >> (8 * i)) & 0xFF;

  for (uint32_t k = 0; k < warp_chans; k++) {
    const int max_w = static_cast<int>(warp_dims[2 * k]), max_h = static_cast<int>(warp_dims[2 * k + 1]);
    DevChannel ch{};
    int w = 0, h = 0;
    uint32_t stride = 0;
    bool direct = false;  // neighbours are read from the output plane instead of the lane's row ring
    int32_t* out = nullptr;
    if (k < my_chans) {
      ch = P.chans[st.chan_begin + k];
      if (ch.preamble) {
        // the previous entropy-coded stream ends here; the next one follows bit by bit
        if (!code.use_prefix && reader.state != (0x13u << 16)) status |= kStatusBadFinalState;
        dyn_count = br.Read(ch.count_bits) + 1;
        const uint32_t use_global_tree = br.Read(1);
        if (br.Read(1)) {  // default weighted-predictor header (context_predict.h:37-61)
          wpp.p1C = 16; wpp.p2C = 10; wpp.p3Ca = 7; wpp.p3Cb = 7; wpp.p3Cc = 7; wpp.p3Cd = 0; wpp.p3Ce = 0;
          wpp.w[0] = 0xd; wpp.w[1] = 0xc; wpp.w[2] = 0xc; wpp.w[3] = 0xc;
        } else {
          wpp.p1C = br.Read(5); wpp.p2C = br.Read(5); wpp.p3Ca = br.Read(5); wpp.p3Cb = br.Read(5); wpp.p3Cc = br.Read(5);
          wpp.p3Cd = br.Read(5); wpp.p3Ce = br.Read(5);
          for (int i = 0; i < 4; i++) wpp.w[i] = br.Read(4);
        }
        const uint32_t transforms_selector = br.Read(2);
        if (!use_global_tree || transforms_selector != 0) status |= kStatusUnsupported;
        reader.state = code.use_prefix ? (0x13u << 16) : br.Read(32);
      }
      const DevPlane pl = P.planes[ch.plane];
      w = static_cast<int>(pl.w);
      h = static_cast<int>(pl.h);
      stride = pl.w;
      out = P.arena + pl.off;
      if (ch.dyn) {
        if (dyn_count > pl.w) {
          status |= kStatusUnsupported;
          dyn_count = pl.w;
        }
        w = static_cast<int>(dyn_count);
        out[static_cast<size_t>(pl.w) * pl.h] = static_cast<int32_t>(dyn_count);
        direct = true;
        if (ch.uses_wp) status |= kStatusUnsupported;
      }
    }
    const DevTreeNode* tree = P.tree + ch.tree_off;
    const bool uses_wp = ch.uses_wp != 0 && !direct;
    // every lane of the warp that has this channel slot qualifies for the weighted-predictor LUT path
    const bool fast = JXLB_WARP_ALL_M(k >= my_chans || (ch.wp_lut != 0 && uses_wp && ch.ref_count == 0));
    const bool nw = JXLB_WARP_ALL_M(k >= my_chans || ch.nw_lut != 0);
    props[0] = static_cast<int32_t>(ch.prop0);
    props[15 * PS] = 0;
    if (nw && plain_ans) {
      DevDecodeChannelRows<WT, false, kLS, true, true>(P, m, ch, wpp, w, h, max_w, max_h, stride, out, direct, uses_wp, tree, reader, br);
    } else if (nw) {
      DevDecodeChannelRows<WT, false, kLS, false, true>(P, m, ch, wpp, w, h, max_w, max_h, stride, out, direct, uses_wp, tree, reader, br);
    } else if (fast && plain_ans) {
      DevDecodeChannelRows<WT, true, kLS, true>(P, m, ch, wpp, w, h, max_w, max_h, stride, out, direct, uses_wp, tree, reader, br);
    } else if (fast) {
      DevDecodeChannelRows<WT, true, kLS, false>(P, m, ch, wpp, w, h, max_w, max_h, stride, out, direct, uses_wp, tree, reader, br);
    } else if (plain_ans) {
      DevDecodeChannelRows<WT, false, kLS, true>(P, m, ch, wpp, w, h, max_w, max_h, stride, out, direct, uses_wp, tree, reader, br);
    } else {
      DevDecodeChannelRows<WT, false, kLS, false>(P, m, ch, wpp, w, h, max_w, max_h, stride, out, direct, uses_wp, tree, reader, br);
    }
  }
  if (lane_valid) {
    if (!code.use_prefix && reader.state != (0x13u << 16)) status |= kStatusBadFinalState;
    if (br.Pos() > st.bit_end) status |= kStatusOverread;
    if (end_pos) *end_pos = br.Pos();
  }
  return status;
}

}  // namespace jxlb

#endif  // JXLB_MODULAR_DEV_H_
