// jxl_b200 device code: one Modular entropy-coded stream decoded by one WARP (k_modular_decode_coop).
//
// A VarDCT DC group is one strictly serial chain of up to ~330 000 samples (three DC channels, then the AC-metadata
// stream, lib/jxl/dec_frame.cc:315-339), and a 4K frame has only four of them: the batch offers ~1000 chains, far
// fewer than the chip has lanes, so the time of the kernel is the time of its longest chain and what matters is the
// latency of one sample. The serial dependence of a sample is
//     previous value -> context (cluster) -> alias entry of (cluster, ANS state) -> symbol, new state, refill
//                    -> extra bits of the hybrid integer -> value,
// (lib/jxl/dec_ans.h:168-195, :286-343). The warp cuts it in two by speculation: lane l decodes the next integer AS IF
// its cluster were l -- alias entry, state update, the 16-bit refill and the extra bits, all from the same (warp-uniform)
// ANS state and bit window -- while all lanes compute the real context from the previous value; three shuffles then
// pick the integer, the state and the number of consumed bits of the real cluster's lane. The per-sample latency
// becomes max(speculative decode, context) + one shuffle instead of their sum plus two dependent global loads.
//
// Eligible streams (host: BundleStreams): plain ANS without LZ77, every channel decodable through
// one of the table paths of DevDecodeChannelRows -- the gradient-property table or the (y, N, W) bucket table
// (libjxl's fixed DC / AC-metadata trees, lib/jxl/modular/encoding/enc_encoding.cc:218-282) or the weighted-predictor
// table (its default DC tree, :266-273). Everything else stays with the lock-step kernels.
//
// All lanes execute the same instructions on the same values (the previous-row window, the neighbours, the context);
// only the speculative decode differs by lane. Rows live in shared memory ([x], one copy per warp), written by lane 0.
//
// __host__ __device__ like the rest of the decode path: on the host (tests/emul) there is one "lane", which decodes
// for the real cluster directly; the window, refill and hybrid-integer code is the same text.
#ifndef JXLB_MODULAR_COOP_DEV_H_
#define JXLB_MODULAR_COOP_DEV_H_

#include "jxlb_modular_dev.h"

namespace jxlb {

#if defined(__CUDA_ARCH__)
#define JXLB_SHFL(v, src) __shfl_sync(0xFFFFFFFFu, (v), (src))
#define JXLB_SYNCWARP() __syncwarp()
#define JXLB_LANE() (threadIdx.x & 31u)
#else
#define JXLB_SHFL(v, src) (v)
#define JXLB_SYNCWARP() ((void)0)
#define JXLB_LANE() 0u
#endif

// Keeps a pointer that is computed once (table of the lane's cluster, table biased by its lowest property value) as
// ONE 64-bit register pair, so that the per-sample address is a single widening multiply-add instead of the compiler's
// re-association into base + offset + index.
template <typename T>
JXLB_HD const T* DevOpaque(const T* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+l"(p));
#endif
  return p;
}

JXLB_HD uint32_t DevFunnelR(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, s);
#else
  return static_cast<uint32_t>(((static_cast<uint64_t>(hi) << 32) | lo) >> (s & 31));
#endif
}

// The warp's bit window: four consecutive words of the stream and the bit offset inside the first (o < 32), so that
// 64 bits are always available at the position -- a refill (16) plus the extra bits of one integer (<= 31) need 47.
// Words at or behind the section end read as zero (DevBits::Load); w3 is loaded one word before it can be needed.
struct DevCoopBits {
  const uint32_t* words;  // word 0 = the word holding the stream's first bit
  uint64_t word0;         // its index in the pool (for Pos)
  uint32_t next, limit;   // relative to word 0; sections are far below 2^32 words
  uint32_t w0, w1, w2, w3, o;

  JXLB_HD uint32_t Load(uint32_t w) const { return w < limit ? JXLB_LDG(words + w) : 0u; }
  JXLB_HD void Init(const uint32_t* w, uint64_t bit_pos, uint64_t bit_end) {
    word0 = bit_pos >> 5;
    words = w + word0;
    const uint64_t lim = ((bit_end + 31) >> 5);
    limit = lim > word0 ? static_cast<uint32_t>(lim - word0) : 0u;
    o = static_cast<uint32_t>(bit_pos & 31);
    w0 = Load(0);
    w1 = Load(1);
    w2 = Load(2);
    w3 = Load(3);
    next = 4;
  }
  JXLB_HD uint64_t Pos() const { return (word0 + next - 4) * 32 + o; }
  JXLB_HD void Advance(uint32_t k) {  // k <= 63
    o += k;
    if (o >= 32) {
      w0 = w1; w1 = w2; w2 = w3;
      w3 = Load(next++);
      o -= 32;
      if (o >= 32) {
        w0 = w1; w1 = w2; w2 = w3;
        w3 = Load(next++);
        o -= 32;
      }
    }
  }
};

// What a lane needs to decode an integer of "its" cluster: the cluster's alias table and hybrid-integer configuration.
struct DevCoopLane {
  const DevAlias* alias;  // table of the lane's cluster
  uint32_t split_token, split_exp, in_token, lsb, msb;
  JXLB_HD void Set(const DevAlias* alias_base, const uint32_t* cfg, uint32_t cluster, uint32_t log_alpha) {
    alias = DevOpaque(alias_base + (static_cast<size_t>(cluster) << log_alpha));
    const uint32_t c = JXLB_LDG(cfg + cluster);
    split_exp = c & 0xFF;
    msb = (c >> 8) & 0xFF;
    lsb = (c >> 16) & 0xFF;
    split_token = 1u << split_exp;
    in_token = msb + lsb;
  }
};

struct DevCoopReader {
  const DevAlias* alias_base;
  const uint32_t* cfg;
  uint32_t log_alpha, log_entry, pos_mask;
  uint32_t state;  // warp-uniform
  DevCoopLane mine;

  JXLB_HD void Init(const DevPools& P, const DevCode& c) {
    alias_base = P.alias + c.alias_off;
    cfg = P.cfg + c.cfg_off;
    log_alpha = c.log_alpha_size;
    log_entry = 12 - log_alpha;
    pos_mask = (1u << log_entry) - 1;
  }

  // The lanes of the warp take the clusters of the channel's list (DevChannel::coop_list_off): lane l the l-th.
  JXLB_HD void SetChannel(const uint16_t* list) {
    const uint32_t count = JXLB_LDG(list), lane = JXLB_LANE();
    mine.Set(alias_base, cfg, JXLB_LDG(list + 1 + (lane < count ? lane : 0u)), log_alpha);
#if !defined(__CUDA_ARCH__)
    host_list = list;
#endif
  }

  // One integer of the cluster of lane `L`, from the warp's state and the 64 bits at the bit position:
  // lib/jxl/dec_ans.h:168-195 (alias lookup, state update, refill) and :286-343 (hybrid integer, DevReadHybrid without
  // branches: below the split the token is the value and no bits follow). `st` / `k`: the state afterwards and the
  // number of bits consumed.
  JXLB_HD uint32_t Decode(const DevCoopLane& L, uint32_t vlo, uint32_t vhi, uint32_t* st_out, uint32_t* k_out) const {
    const uint32_t res = state & 0xFFF;
    const uint32_t i = res >> log_entry, pos = res & pos_mask;
#if defined(__CUDA_ARCH__)
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(L.alias + i));
    const uint32_t cutoff = raw.x & 0xFF, right_value = (raw.x >> 8) & 0xFF, freq0 = raw.x >> 16;
    const uint32_t offsets1 = raw.y & 0xFFFF, fx = raw.y >> 16;
#else
    const DevAlias& e = L.alias[i];
    const uint32_t cutoff = e.cutoff, right_value = e.right_value, freq0 = e.freq0;
    const uint32_t offsets1 = e.offsets1, fx = e.freq1_xor_freq0;
#endif
    const bool right = pos >= cutoff;
    const uint32_t sym = right ? right_value : i;
    const uint32_t offset = (right ? offsets1 : 0u) + pos;
    const uint32_t freq = right ? (freq0 ^ fx) : freq0;
    uint32_t st = freq * (state >> 12) + offset;
    const bool refill = st < (1u << 16);
    st = refill ? ((st << 16) | (vlo & 0xFFFF)) : st;
    const uint32_t v2 = refill ? DevFunnelR(vlo, vhi, 16) : vlo;  // the 32 bits behind the refill
    const bool direct = sym < L.split_token;
    const uint32_t nbits = direct ? 0u : ((L.split_exp - L.in_token + ((sym - L.split_token) >> L.in_token)) & 31);
    const uint32_t low = sym & ((1u << L.lsb) - 1);
    const uint32_t tok = sym >> L.lsb;
    const uint32_t bits = v2 & ((1u << nbits) - 1);
    const uint32_t hi = (1u << L.msb) | (tok & ((1u << L.msb) - 1));
    const uint32_t big = (((hi << nbits) | bits) << L.lsb) | low;
    *st_out = st;
    *k_out = (refill ? 16u : 0u) + nbits;
    return direct ? sym : big;
  }

  // `leaf` (warp-uniform, from the channel's coop table): the lane that speculates on the sample's cluster, or
  // 0x8000 | cluster when no lane does. Every lane decodes for its own cluster; three shuffles pick the real one.
  // kMiss = false: the channel's table has no leaf outside the list (DevChannel::coop == 1), the test is compiled out.
  template <bool kMiss = true>
  JXLB_HD uint32_t ReadUint(uint32_t leaf, DevCoopBits& br) {
    const uint32_t vlo = DevFunnelR(br.w0, br.w1, br.o), vhi = DevFunnelR(br.w1, br.w2, br.o);
    uint32_t u, st, k;
#if defined(__CUDA_ARCH__)
    // (speculate first, repair afterwards: the common path of a sample stays one basic block, in which the scheduler
    // interleaves the context chain with this one)
    u = Decode(mine, vlo, vhi, &st, &k);
    u = JXLB_SHFL(u, leaf);
    st = JXLB_SHFL(st, leaf);
    k = JXLB_SHFL(k, leaf);
    if (kMiss && (leaf & 0x8000u)) {  // (uniform, rare) a cluster outside the list: all lanes decode it
      DevCoopLane other;
      other.Set(alias_base, cfg, leaf & 0xFF, log_alpha);
      u = Decode(other, vlo, vhi, &st, &k);
    }
#else
    // the host's single lane decodes the real cluster: of the list entry, or the one the leaf carries
    DevCoopLane other;
    other.Set(alias_base, cfg, (leaf & 0x8000u) ? (leaf & 0xFF) : host_list[1 + leaf], log_alpha);
    u = Decode(other, vlo, vhi, &st, &k);
#endif
    state = st;
    br.Advance(k);
    return u;
  }
#if !defined(__CUDA_ARCH__)
  const uint16_t* host_list = nullptr;
#endif
};

// The x loop of one row for the table paths whose row has ONE predictor out of Zero / Left / Gradient (kPred = 0, 1, 5)
// -- every row libjxl's fixed gradient DC tree and fixed AC-metadata tree produce (enc_encoding.cc:218-282; the host
// records the row's predictor behind the coop table): only W, N and NW are needed, so the previous row is a three-deep
// prefetch queue and nothing else slides. The edge rules of context_predict.h:496-504 are folded into the initial values
// (W = N at x = 0, NW = W at x = 0) and a clamped prefetch index; on the first row N = NW = W.
// kMode: 1 = (y, N, W) bucket table (`tab` = the row's slice), 2 = gradient-property table.
template <typename WT, int kMode, int kPred, bool kDirect, bool kMiss>
JXLB_HD void DevCoopFastRow(const uint16_t* tab, const int32_t lut_lo, const int32_t lut_hi, const int32_t* nw_tn,
                            const int32_t* nw_tw, const int w, const bool first_row, const int32_t* prev, int32_t* row,
                            int32_t* out_row, const bool lane0, DevCoopReader& reader, DevCoopBits& br) {
  const int last = w - 1;
  const uint16_t* tab0 = DevOpaque(kMode == 2 ? tab - lut_lo : tab);  // gradient table indexed by the clamped property itself
  int32_t t = 0, q1 = 0, q2 = 0;  // N of x, x + 1, x + 2
  if (!first_row) {
    t = prev[0];
    q1 = prev[1 < last ? 1 : last];
    q2 = prev[2 < last ? 2 : last];
  }
  int32_t left = t, tl = t;  // x = 0: W = N (0 on the first row), NW = W
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
  for (int x = 0; x < w; x++) {
    const WT n_left = left;
    const WT n_top = first_row ? left : t;
    const WT n_topleft = first_row ? left : tl;
    uint32_t e;
    if (kMode == 2) {
      const WT p9 = n_left + n_top - n_topleft;
      const WT pv = p9 < static_cast<WT>(lut_lo) ? static_cast<WT>(lut_lo) : (p9 > static_cast<WT>(lut_hi) ? static_cast<WT>(lut_hi) : p9);
      e = JXLB_LDG(tab0 + static_cast<int32_t>(pv));
    } else {
      uint32_t bn = 0, bw = 0;
      for (uint32_t i = 0; i < kNwThresholds; i++) {
        bn += n_top > static_cast<WT>(nw_tn[i]) ? 1u : 0u;
        bw += n_left > static_cast<WT>(nw_tw[i]) ? 1u : 0u;
      }
      e = JXLB_LDG(tab0 + (bn * (kNwThresholds + 1) + bw));
    }
    const uint32_t u = reader.template ReadUint<kMiss>(e & 0x80FF, br);
    const WT guess = kPred == 0 ? static_cast<WT>(0)
                                : (kPred == 1 ? n_left
                                              : static_cast<WT>(DevClampedGradient(static_cast<int32_t>(n_left), static_cast<int32_t>(n_top),
                                                                                   static_cast<int32_t>(n_topleft))));
    const int32_t val = static_cast<int32_t>(static_cast<uint32_t>(DevUnpackSigned(u)) + static_cast<uint32_t>(guess));
    JXLB_SYNCWARP();  // the other lanes read this entry (as the previous row's) three samples ago: order that before the write
    if (lane0) {
      row[x] = val;
      if (!kDirect) out_row[x] = val;
    }
    left = val;
    tl = t;
    t = q1;
    q1 = q2;
    if (!first_row && x < last) {  // (at the last sample the entry has just been overwritten and nothing reads q2 again)
      const int ahead = x + 3 < last ? x + 3 : last;
      q2 = prev[ahead];
    }
  }
}

// Row loops of one channel for the warp. kMode: 1 = (y, N, W) bucket table, 2 = gradient-property table, 3 =
// weighted-predictor table (DevChannel::nw_lut / wp_lut; same tables and arithmetic as DevDecodeChannelRows, which
// stays the reference for these paths in the lock-step kernels). Row storage as there: A holds the previous row and is
// overwritten in place by the current one, B receives the sample A loses and is the row above the previous one for the
// next row; the weighted predictor's five error rows are updated in place the same way. `rows` = 2 x ring_w samples,
// `wp` = 5 x (ring_w + 2), both private to the warp, written by lane 0 and read by all lanes (one __syncwarp per row:
// entries are written at x and read at x + 2 / x + 3 of the NEXT row at the earliest).
template <typename WT, int kMode>
JXLB_HD void DevCoopChannelRows(const DevPools& P, const DevChannel& ch, const DevWpParams& wpp, const int w, const int h,
                                const uint32_t stride, int32_t* out, const bool direct, int32_t* rows, int32_t* wp,
                                const uint32_t ring_w, const uint32_t* divlut, DevCoopReader& reader, DevCoopBits& br) {
  const bool lane0 = JXLB_LANE() == 0;
  const uint16_t* lut = P.lut + ch.coop_lut_off;  // leaves in lanes
  reader.SetChannel(P.lut + ch.coop_list_off);
  const int32_t lut_lo = ch.lut_lo, lut_hi = ch.lut_lo + static_cast<int32_t>(ch.lut_size) - 1;
  int32_t nw_tn[kNwThresholds] = {}, nw_tw[kNwThresholds] = {};
  uint32_t nw_ny = 0;
  if (kMode == 1 && h > 0) {
    nw_ny = JXLB_LDG(lut);
    for (uint32_t i = 0; i < kNwThresholds; i++) {
      nw_tn[i] = DevNwThreshold(lut + kNwOffN, i);
      nw_tw[i] = DevNwThreshold(lut + kNwOffW, i);
    }
  }
  const uint16_t* nw_row = lut;
  if (kMode != 3 && !(kMode == 2 && direct)) {
    // the predictor of every row bucket, recorded by the host behind the table (0xFF = mixed): when each is Zero, Left
    // or Gradient -- libjxl's fixed trees -- the rows go through DevCoopFastRow
    const uint16_t* rp = kMode == 1 ? lut + kNwOffTable + (nw_ny + 1) * ((kNwThresholds + 1) * (kNwThresholds + 1)) : lut + ch.lut_size;
    const uint32_t buckets = kMode == 1 ? nw_ny + 1 : 1;
    bool all_fast = h > 0;
    for (uint32_t b = 0; b < buckets; b++) {
      const uint32_t pr = JXLB_LDG(rp + b);
      all_fast = all_fast && (pr == 0 || pr == 1 || pr == 5);
    }
    if (all_fast) {
      for (int y = 0; y < h; y++) {
        int32_t* out_row = out + static_cast<size_t>(y) * stride;
        int32_t* row = direct ? out_row : rows;
        const int32_t* prev = direct ? out_row - stride : rows;
        uint32_t by = 0;
        if (kMode == 1) {
          for (uint32_t i = 0; i < nw_ny; i++) by += y > DevNwThreshold(lut + kNwOffY, i) ? 1u : 0u;
          nw_row = lut + kNwOffTable + by * ((kNwThresholds + 1) * (kNwThresholds + 1));
        }
        const uint32_t pr = JXLB_LDG(rp + by);
        const bool first = y == 0;
#define JXLB_FAST_ROW(P_, D_, M_) \
  DevCoopFastRow<WT, kMode, P_, D_, M_>(nw_row, lut_lo, lut_hi, nw_tn, nw_tw, w, first, prev, row, out_row, lane0, reader, br)
        if (ch.coop == 1) {
          if (kMode == 1 && direct) {
            if (pr == 0) JXLB_FAST_ROW(0, true, false); else if (pr == 1) JXLB_FAST_ROW(1, true, false); else JXLB_FAST_ROW(5, true, false);
          } else {
            if (pr == 0) JXLB_FAST_ROW(0, false, false); else if (pr == 1) JXLB_FAST_ROW(1, false, false); else JXLB_FAST_ROW(5, false, false);
          }
        } else {
          if (kMode == 1 && direct) {
            if (pr == 0) JXLB_FAST_ROW(0, true, true); else if (pr == 1) JXLB_FAST_ROW(1, true, true); else JXLB_FAST_ROW(5, true, true);
          } else {
            if (pr == 0) JXLB_FAST_ROW(0, false, true); else if (pr == 1) JXLB_FAST_ROW(1, false, true); else JXLB_FAST_ROW(5, false, true);
          }
        }
#undef JXLB_FAST_ROW
        JXLB_SYNCWARP();  // lane 0's row before the next row reads it
      }
      return;
    }
  }
  const uint32_t WL = ring_w + 2;
  int32_t* pe[4];
  for (uint32_t i = 0; i < 4; i++) pe[i] = wp + i * WL;
  int32_t* er = wp + 4 * WL;
  int32_t* rowA = rows;
  int32_t* rowB = rows + ring_w;
  if (kMode == 3) {  // the rows above y = 0 read as zero
    if (lane0)
      for (uint32_t q = 0; q < 5 * WL; q++) wp[q] = 0;
    JXLB_SYNCWARP();
  }
  for (int y = 0; y < h; y++) {
    int32_t* out_row = out + static_cast<size_t>(y) * stride;
    int32_t* row = direct ? out_row : rowA;
    const int32_t* prev = direct ? out_row - stride : rowA;
    const int32_t* prevprev = direct ? out_row - 2 * static_cast<size_t>(stride) : rowB;
    if (kMode == 1) {
      uint32_t by = 0;
      for (uint32_t i = 0; i < nw_ny; i++) by += y > DevNwThreshold(lut + kNwOffY, i) ? 1u : 0u;
      nw_row = lut + kNwOffTable + by * ((kNwThresholds + 1) * (kNwThresholds + 1));
    }
    int32_t left = 0, leftleft = 0, t = 0, tl = 0, tr = 0, trr = 0;
    uint32_t eNW[4] = {0, 0, 0, 0}, eN[4] = {0, 0, 0, 0}, eNE[4] = {0, 0, 0, 0};
    int32_t teW = 0, teNW = 0, teN = 0, teNE = 0;
    if (w > 0) {
      if (y > 0) {
        t = prev[0];
        tr = w > 1 ? prev[1] : t;
        trr = w > 2 ? prev[2] : tr;
      }
      if (kMode == 3) {
        for (uint32_t i = 0; i < 4; i++) {
          eN[i] = static_cast<uint32_t>(pe[i][0]);
          eNW[i] = eN[i];
          eNE[i] = w > 1 ? static_cast<uint32_t>(pe[i][1]) : eN[i];
        }
        teN = er[0];
        teNW = teN;
        teNE = w > 1 ? er[1] : teN;
      }
    }
    for (int x = 0; x < w; x++) {
      // neighbours with the edge rules of context_predict.h:496-504
      const WT n_left = x ? left : (y ? t : 0);
      const WT n_top = y ? t : n_left;
      const WT n_topleft = (x && y) ? tl : n_left;
      const WT n_topright = (x + 1 < w && y) ? tr : n_top;
      const WT n_leftleft = x > 1 ? leftleft : n_left;
      const WT n_toptop = y > 1 ? prevprev[x] : n_top;
      const WT n_toprightright = (x + 2 < w && y) ? trr : n_topright;
      int32_t val;
      WT wp_raw = 0;
      WT prediction[4] = {0, 0, 0, 0};
      if (kMode == 3) {
        // weighted predictor (context_predict.h:134-214), as in DevDecodeChannelRows
        uint32_t weights[4];
        for (uint32_t i = 0; i < 4; i++) {
          const uint32_t e = eN[i] + eNE[i] + eNW[i];
          int shift = static_cast<int>(sizeof(WT) == 4 ? DevFloorLog2_32(e + 1) : DevFloorLog2(static_cast<uint64_t>(e) + 1)) - 5;
          if (shift < 0) shift = 0;
          weights[i] = 4 + ((wpp.w[i] * divlut[e >> shift]) >> shift);
        }
        const WT N8 = n_top * 8, W8 = n_left * 8, NE8 = n_topright * 8, NW8 = n_topleft * 8, NN8 = n_toptop * 8;
        const WT eW = x == 0 ? 0 : teW;
        const WT sumWN = static_cast<WT>(teN) + eW;
        WT pm = eW;
        if (DevAbsW<WT>(teN) > DevAbsW<WT>(pm)) pm = teN;
        if (DevAbsW<WT>(teNW) > DevAbsW<WT>(pm)) pm = teNW;
        if (DevAbsW<WT>(teNE) > DevAbsW<WT>(pm)) pm = teNE;
        const int32_t wp_max_error = static_cast<int32_t>(pm);
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
        const WT wp_pred = (wp_raw + 3) >> 3;
        const int32_t pv = wp_max_error < lut_lo ? lut_lo : (wp_max_error > lut_hi ? lut_hi : wp_max_error);
        const uint32_t leaf = JXLB_LDG(lut + (pv - lut_lo));
        const uint32_t u = reader.ReadUint(leaf, br);
        val = static_cast<int32_t>(static_cast<uint32_t>(DevUnpackSigned(u)) + static_cast<uint32_t>(wp_pred));
      } else {
        uint32_t e;  // lane | predictor << 8, or 0x8000 | cluster | predictor << 8
        if (kMode == 2) {
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
        const WT guess = DevPredictW<WT>((e >> 8) & 0xF, n_left, n_top, n_topleft, n_topright, n_leftleft, n_toptop, n_toprightright, 0);
        const uint32_t u = reader.ReadUint(e & 0x80FF, br);
        val = static_cast<int32_t>(static_cast<uint32_t>(DevUnpackSigned(u)) + static_cast<uint32_t>(guess));
      }
      JXLB_SYNCWARP();  // the other lanes' reads of these entries (previous row, up to three samples ago) before lane 0's writes
      if (lane0) {
        if (!direct) rowB[x] = t;  // the sample A loses below: next row's N-N
        row[x] = val;
        if (!direct) out_row[x] = val;
      }
      leftleft = left;
      left = val;
      tl = t;
      t = tr;
      tr = trr;
      if (y > 0 && x + 3 < w) trr = prev[x + 3];
      if (kMode == 3) {
        const WT val8 = static_cast<WT>(val) * 8;
        const int32_t te = static_cast<int32_t>(wp_raw - val8);
        const bool more = x + 2 < w;  // position x + 2 exists in the previous row
        const int32_t er_ahead = more ? er[x + 2] : 0;
        if (lane0) er[x] = te;
        for (uint32_t i = 0; i < 4; i++) {
          const uint32_t err = static_cast<uint32_t>((DevAbsW<WT>(prediction[i] - val8) + 3) >> 3);
          const uint32_t n_next = eNE[i] + err;
          eNW[i] = eN[i];
          eN[i] = n_next;
          eNE[i] = more ? static_cast<uint32_t>(pe[i][x + 2]) : n_next;
          if (lane0) pe[i][x] = static_cast<int32_t>(err);
        }
        teW = te;
        teNW = teN;
        teN = teNE;
        teNE = more ? er_ahead : teN;
      }
    }
    JXLB_SYNCWARP();  // lane 0's row (shared memory or, for a `direct` channel, the output plane) before the next row reads it
  }
}

// Host side of the eligibility test lives in BundleStreams (jxlb_batch.h); this is the device-side walk of one
// stream's channels, the warp-uniform twin of DevDecodeModularStream (same preamble handling and status bits).
// `rows`: 2 x ring_w samples, `wp`: 5 x (ring_w + 2) samples of the warp's shared memory.
template <typename WT>
JXLB_HD uint32_t DevDecodeModularStreamCoop(const DevPools& P, uint32_t s, int32_t* rows, int32_t* wp, uint32_t ring_w,
                                            const uint32_t* divlut, uint64_t* end_pos) {
  const DevStream st = P.streams[s];
  const DevCode code = P.codes[st.code];
  const bool chained = st.chain_slot != 0;  // starts (with its GroupHeader) where an AC coefficient stream ended
  DevCoopBits br;
  br.Init(P.words, chained ? P.chain_pos[st.chain_slot - 1] : st.bit_pos, st.bit_end);
  DevCoopReader reader;
  reader.Init(P, code);
  if (chained) {
    reader.state = 0x13u << 16;  // (what the preamble of the first channel expects of a finished stream)
  } else {  // the initial ANS state (32 bits)
    reader.state = DevFunnelR(br.w0, br.w1, br.o);
    br.Advance(32);
  }
  DevWpParams wpp;
  wpp.p1C = st.wp_params[0] & 0xFF; wpp.p2C = (st.wp_params[0] >> 8) & 0xFF; wpp.p3Ca = (st.wp_params[0] >> 16) & 0xFF;
  wpp.p3Cb = st.wp_params[0] >> 24; wpp.p3Cc = st.wp_params[1] & 0xFF; wpp.p3Cd = (st.wp_params[1] >> 8) & 0xFF;
  wpp.p3Ce = (st.wp_params[1] >> 16) & 0xFF;
  for (int i = 0; i < 4; i++) wpp.w[i] = (st.wp_params[2] >> (8 * i)) & 0xFF;
  uint32_t status = kStatusOk;
  uint32_t dyn_count = 0;
  auto read_bits = [&](uint32_t k) {  // k <= 32, warp-uniform
    const uint32_t v = DevFunnelR(br.w0, br.w1, br.o);
    br.Advance(k);
    return k >= 32 ? v : (v & ((1u << k) - 1));
  };
  for (uint32_t c = st.chan_begin; c < st.chan_end; c++) {
    const DevChannel ch = P.chans[c];
    if (ch.preamble) {
      // the previous entropy-coded stream ends here; the next one follows bit by bit (DevDecodeModularStream)
      if (reader.state != (0x13u << 16)) status |= kStatusBadFinalState;
      dyn_count = read_bits(ch.count_bits) + 1;
      const uint32_t use_global_tree = read_bits(1);
      if (read_bits(1)) {  // default weighted-predictor header (context_predict.h:37-61)
        wpp.p1C = 16; wpp.p2C = 10; wpp.p3Ca = 7; wpp.p3Cb = 7; wpp.p3Cc = 7; wpp.p3Cd = 0; wpp.p3Ce = 0;
        wpp.w[0] = 0xd; wpp.w[1] = 0xc; wpp.w[2] = 0xc; wpp.w[3] = 0xc;
      } else {
        wpp.p1C = read_bits(5); wpp.p2C = read_bits(5); wpp.p3Ca = read_bits(5); wpp.p3Cb = read_bits(5);
        wpp.p3Cc = read_bits(5); wpp.p3Cd = read_bits(5); wpp.p3Ce = read_bits(5);
        for (int i = 0; i < 4; i++) wpp.w[i] = read_bits(4);
      }
      const uint32_t transforms_selector = read_bits(2);
      if (!use_global_tree || transforms_selector != 0) status |= kStatusUnsupported;
      reader.state = read_bits(32);
    }
    const DevPlane pl = P.planes[ch.plane];
    int w = static_cast<int>(pl.w);
    const int h = static_cast<int>(pl.h);
    int32_t* out = P.arena + pl.off;
    bool direct = false;
    if (ch.dyn) {
      if (dyn_count > pl.w) {
        status |= kStatusUnsupported;
        dyn_count = pl.w;
      }
      w = static_cast<int>(dyn_count);
      if (JXLB_LANE() == 0) out[static_cast<size_t>(pl.w) * pl.h] = static_cast<int32_t>(dyn_count);
      direct = true;
      if (ch.uses_wp) status |= kStatusUnsupported;
    }
    if (ch.nw_lut == 2) {
      DevCoopChannelRows<WT, 2>(P, ch, wpp, w, h, pl.w, out, direct, rows, wp, ring_w, divlut, reader, br);
    } else if (ch.nw_lut == 1) {
      DevCoopChannelRows<WT, 1>(P, ch, wpp, w, h, pl.w, out, direct, rows, wp, ring_w, divlut, reader, br);
    } else if (ch.wp_lut && !direct) {
      DevCoopChannelRows<WT, 3>(P, ch, wpp, w, h, pl.w, out, direct, rows, wp, ring_w, divlut, reader, br);
    } else {
      status |= kStatusUnsupported;  // (the host never sends such a stream here)
    }
  }
  if (reader.state != (0x13u << 16)) status |= kStatusBadFinalState;
  if (br.Pos() > st.bit_end) status |= kStatusOverread;
  if (end_pos) *end_pos = br.Pos();
  return status;
}

}  // namespace jxlb

#endif  // JXLB_MODULAR_COOP_DEV_H_
