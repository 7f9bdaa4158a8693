// jxl_b200 encoder, host side: everything that is O(KB) per frame -- quantiser parameters, the global
// Modular tree, histogram normalisation and the ANS tables, the codestream / frame headers, the TOC and the
// assembly of the sections the kernels produce. Mirrors the host part of jpegxl-rs's encode path
// (jpegxl-rs/src/encode.rs:345-378 -> lib/jxl/enc_frame.cc:302-478 MakeFrameHeader, :1253-1424 EncodeGroups,
// lib/jxl/enc_ans.cc:119-364 histogram normalisation and coding, :1731-1816 WriteTokens, lib/jxl/enc_toc.cc,
// lib/jxl/enc_modular.cc fixed trees). The per-sample work (XYB, strategy choice, forward transforms, quantisation,
// tokenisation, histogram counting, rANS emission) runs in the kernels of kernels/jxlb_enc_dev.h.
#ifndef JXLB_ENC_HOST_H_
#define JXLB_ENC_HOST_H_

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <vector>

#include "../kernels/jxlb_enc_desc.h"
#include "jxlb_vardct_plan.h"

namespace jxlb {

class BitWriter {
 public:
  void Write(unsigned n, uint64_t v) {  // n <= 56, LSB first
    for (unsigned i = 0; i < n; i++) {
      if ((bits_ & 7) == 0) bytes_.push_back(0);
      if ((v >> i) & 1) bytes_.back() |= static_cast<uint8_t>(1u << (bits_ & 7));
      bits_++;
    }
  }
  void ZeroPadToByte() { bits_ = (bits_ + 7) & ~size_t{7}; }
  size_t BitsWritten() const { return bits_; }
  const std::vector<uint8_t>& Bytes() const { return bytes_; }
  void AppendBits(const uint8_t* data, size_t nbits) {
    for (size_t i = 0; i < nbits; i++) Write(1, (data[i >> 3] >> (i & 7)) & 1);
  }
  // Appends `nbits` bits that start at bit `first` of the 32-bit word array `words` (word-wise, any alignment).
  void AppendWordBits(const uint32_t* words, uint64_t first, uint64_t nbits) {
    while (nbits > 0) {
      const uint32_t sh = static_cast<uint32_t>(first & 31);
      uint64_t v = words[first >> 5] >> sh;
      uint32_t have = 32 - sh;
      if (have < 32 && nbits > have) {
        v |= static_cast<uint64_t>(words[(first >> 5) + 1]) << have;
        have = 32;  // (only `take` of them are used)
      }
      const uint32_t take = static_cast<uint32_t>(std::min<uint64_t>(nbits, std::min<uint32_t>(have, 32)));
      if ((bits_ & 7) == 0 && take == 32) {  // fast path: byte-aligned output
        for (int k = 0; k < 4; k++) bytes_.push_back(static_cast<uint8_t>(v >> (8 * k)));
        bits_ += 32;
      } else {
        Write(take, v & ((uint64_t{1} << take) - 1));
      }
      first += take;
      nbits -= take;
    }
  }
  void AppendBytes(const uint8_t* data, size_t n) {
    ZeroPadToByte();
    bytes_.insert(bytes_.end(), data, data + n);
    bits_ = bytes_.size() * 8;
  }

 private:
  std::vector<uint8_t> bytes_;
  size_t bits_ = 0;
};

inline void WriteU32(BitWriter& w, uint32_t v, U32Dist d0, U32Dist d1, U32Dist d2, U32Dist d3) {
  const U32Dist d[4] = {d0, d1, d2, d3};
  for (uint32_t s = 0; s < 4; s++) {
    if (d[s].bits == 0xFF) {
      if (d[s].offset == v) {
        w.Write(2, s);
        return;
      }
    } else if (v >= d[s].offset && (d[s].bits >= 32 || v - d[s].offset < (1ull << d[s].bits))) {
      w.Write(2, s);
      w.Write(d[s].bits, v - d[s].offset);
      return;
    }
  }
  throw Error("value not representable in a U32 field");
}

inline void WriteU64(BitWriter& w, uint64_t v) {
  if (v == 0) {
    w.Write(2, 0);
  } else if (v <= 16) {
    w.Write(2, 1);
    w.Write(4, v - 1);
  } else if (v <= 272) {
    w.Write(2, 2);
    w.Write(8, v - 17);
  } else {
    throw Error("internal: large U64 not needed by the encoder");
  }
}

inline void WriteVarLenUint8(BitWriter& w, uint32_t v) {
  if (v == 0) {
    w.Write(1, 0);
    return;
  }
  w.Write(1, 1);
  const unsigned n = FloorLog2(v);
  w.Write(3, n);
  w.Write(n, v - (1u << n));
}

inline uint32_t PackSignedH(int32_t v) { return (static_cast<uint32_t>(v) << 1) ^ (v < 0 ? 0xFFFFFFFFu : 0u); }

// ---------------------------------------------------------------- histograms -> ANS code
// Normalises counts to a sum of 4096 keeping every used symbol.
inline std::vector<int32_t> NormalizeHistogram(const uint32_t* counts, size_t n) {
  uint64_t total = 0;
  size_t used = 0, alphabet = 0;
  for (size_t i = 0; i < n; i++) {
    total += counts[i];
    used += counts[i] != 0;
    if (counts[i]) alphabet = i + 1;
  }
  std::vector<int32_t> out(alphabet, 0);
  if (total == 0) {
    out.assign(1, kAnsTabSize);
    return out;
  }
  JXLB_CHECK(used <= kAnsTabSize, "too many symbols");
  int64_t sum = 0;
  for (size_t i = 0; i < alphabet; i++) {
    if (!counts[i]) continue;
    int32_t v = static_cast<int32_t>((static_cast<uint64_t>(counts[i]) * kAnsTabSize + total / 2) / total);
    if (v < 1) v = 1;
    out[i] = v;
    sum += v;
  }
  int64_t diff = static_cast<int64_t>(kAnsTabSize) - sum;
  while (diff != 0) {  // spread the correction over the largest entries
    size_t best = 0;
    for (size_t i = 0; i < out.size(); i++)
      if (out[i] > out[best]) best = i;
    const int64_t step = diff > 0 ? diff : std::max<int64_t>(diff, -(out[best] - 1));
    JXLB_CHECK(step != 0, "cannot normalise histogram");
    out[best] += static_cast<int32_t>(step);
    diff -= step;
  }
  return out;
}

// The histogram in the format ReadAnsHistogram parses (lib/jxl/dec_ans.cc:51-187), full precision.
inline void WriteAnsHistogram(BitWriter& w, const std::vector<int32_t>& counts) {
  size_t used = 0, last = 0, first = 0;
  for (size_t i = 0; i < counts.size(); i++)
    if (counts[i]) {
      if (!used) first = i;
      used++;
      last = i;
    }
  if (used == 1) {
    w.Write(1, 1);  // simple
    w.Write(1, 0);  // one symbol
    WriteVarLenUint8(w, first);
    return;
  }
  w.Write(1, 0);  // not simple
  w.Write(1, 0);  // not flat
  w.Write(3, 7);  // shift = 13: unary 3 ...
  w.Write(3, 14 - 8);  // ... then (shift + 1) - 8
  const size_t length = std::max<size_t>(3, last + 1);
  WriteVarLenUint8(w, length - 3);
  static const uint8_t kLen[14] = {5, 4, 4, 4, 4, 4, 3, 3, 3, 3, 3, 6, 7, 7};
  static const uint8_t kCode[14] = {17, 11, 15, 3, 9, 7, 4, 2, 5, 6, 0, 33, 1, 65};
  std::vector<int> logcounts(length, 0);
  int omit_log = -1;
  size_t omit_pos = 0;
  for (size_t i = 0; i < length; i++) {
    const int32_t c = i < counts.size() ? counts[i] : 0;
    logcounts[i] = c == 0 ? 0 : static_cast<int>(FloorLog2(c)) + 1;
    if (logcounts[i] > omit_log) {
      omit_log = logcounts[i];
      omit_pos = i;
    }
    w.Write(kLen[logcounts[i]], kCode[logcounts[i]]);
  }
  for (size_t i = 0; i < length; i++) {
    const int code = logcounts[i];
    if (i == omit_pos || code <= 1) continue;
    const int bitcount = PopulationCountPrecision(code - 1, 13);
    w.Write(bitcount, (counts[i] - (1 << (code - 1))) >> (code - 1 - bitcount));
  }
}

// One entropy code over `num_ctx` contexts with a fixed context -> cluster map, HybridUintConfig(4, 2, 0), no
// LZ77: writes the header, and fills the device tables for the rANS writer (kernels/jxlb_enc_dev.h).
struct EncCode {
  uint32_t num_clusters = 1, log_alpha = 5;
  std::vector<uint16_t> freq;   // [cluster][256]
  std::vector<uint16_t> start;  // [cluster][256]: first slot of the symbol in `reverse`
  std::vector<uint16_t> reverse;  // [cluster][4096]: (symbol, offset) -> rANS residue
  // device form of freq / start: [cluster][256] = freq | start << 16
  std::vector<uint32_t> Fs() const {
    std::vector<uint32_t> v(freq.size());
    for (size_t i = 0; i < freq.size(); i++) v[i] = freq[i] | (static_cast<uint32_t>(start[i]) << 16);
    return v;
  }
};

inline void WriteSimpleCode(BitWriter& w, const std::vector<uint32_t>& values, uint32_t num_values_alphabet);

inline void WriteContextMap(BitWriter& w, const std::vector<uint8_t>& cluster_of, uint32_t num_clusters);

// `hist`: [num_clusters][256] counts of hybrid tokens.
inline void WriteCodeHeader(BitWriter& w, const std::vector<uint8_t>& cluster_of, uint32_t num_clusters, const uint32_t* hist,
                            EncCode* code) {
  w.Write(1, 0);  // no LZ77
  if (cluster_of.size() > 1) WriteContextMap(w, cluster_of, num_clusters);
  w.Write(1, 0);  // ANS
  size_t max_alphabet = 1;
  for (uint32_t c = 0; c < num_clusters; c++)
    for (size_t i = 0; i < 256; i++)
      if (hist[c * 256 + i]) max_alphabet = std::max(max_alphabet, i + 1);
  uint32_t log_alpha = 5;
  while ((size_t{1} << log_alpha) < max_alphabet) log_alpha++;
  JXLB_CHECK(log_alpha <= 8, "alphabet too large");
  w.Write(2, log_alpha - 5);
  for (uint32_t c = 0; c < num_clusters; c++) {  // HybridUintConfig(4, 2, 0)
    w.Write(CeilLog2(log_alpha + 1), 4);
    w.Write(CeilLog2(4 + 1), 2);
    w.Write(CeilLog2(4 - 2 + 1), 0);
  }
  code->num_clusters = num_clusters;
  code->log_alpha = log_alpha;
  code->freq.assign(static_cast<size_t>(num_clusters) * 256, 0);
  code->start.assign(static_cast<size_t>(num_clusters) * 256, 0);
  code->reverse.assign(static_cast<size_t>(num_clusters) * 4096, 0);
  const uint32_t ts = 1u << log_alpha;
  std::vector<AliasEntry> alias(ts);
  for (uint32_t c = 0; c < num_clusters; c++) {
    const std::vector<int32_t> counts = NormalizeHistogram(hist + c * 256, 256);
    WriteAnsHistogram(w, counts);
    uint32_t acc = 0;
    for (size_t s = 0; s < counts.size(); s++) {
      code->freq[c * 256 + s] = counts[s];
      code->start[c * 256 + s] = acc;
      acc += counts[s];
    }
    BuildAliasTable(counts, log_alpha, alias.data());
    // the writer's reverse map comes from the decoder's own lookup (lib/jxl/enc_ans.cc:44-68)
    const uint32_t log_entry = kAnsLogTabSize - log_alpha;
    for (uint32_t res = 0; res < kAnsTabSize; res++) {
      const AliasEntry& e = alias[res >> log_entry];
      const uint32_t pos = res & ((1u << log_entry) - 1);
      const bool right = pos >= e.cutoff;
      const uint32_t sym = right ? e.right_value : (res >> log_entry);
      const uint32_t offset = (right ? e.offsets1 : 0) + pos;
      if (sym < counts.size() && offset < static_cast<uint32_t>(counts[sym]))
        code->reverse[c * 4096 + code->start[c * 256 + sym] + offset] = res;
    }
  }
}

// HybridUintConfig(4, 2, 0)::Encode (lib/jxl/dec_ans.h:73-90)
inline void HybridEncode420(uint32_t v, uint32_t* token, uint32_t* nbits, uint32_t* bits) {
  if (v < 16) {
    *token = v;
    *nbits = 0;
    *bits = 0;
    return;
  }
  const uint32_t n = FloorLog2(v), m = v - (1u << n);
  *token = 16 + ((n - 4) << 2) + (m >> (n - 2));
  *nbits = n - 2;
  *bits = m & ((1u << (n - 2)) - 1);
}

// Host-side rANS writer for the few small streams the host writes itself (tree, context maps).
inline void WriteHostTokens(BitWriter& w, const EncCode& code, const std::vector<uint8_t>& cluster_of,
                            const std::vector<std::pair<uint32_t, uint32_t>>& tokens /* ctx, value */) {
  struct Out { uint32_t nbits, bits; };
  std::vector<Out> out;
  uint32_t state = kAnsSignature << 16;
  for (size_t i = tokens.size(); i-- > 0;) {
    const uint32_t c = cluster_of[tokens[i].first];
    uint32_t token, nbits, bits;
    HybridEncode420(tokens[i].second, &token, &nbits, &bits);
    if (nbits) out.push_back({nbits, bits});
    const uint32_t f = code.freq[c * 256 + token];
    JXLB_CHECK(f > 0, "token outside the histogram");
    if ((state >> (32 - kAnsLogTabSize)) >= f) {
      out.push_back({16, state & 0xFFFF});
      state >>= 16;
    }
    state = ((state / f) << kAnsLogTabSize) | code.reverse[c * 4096 + code.start[c * 256 + token] + state % f];
  }
  w.Write(32, state);
  for (size_t i = out.size(); i-- > 0;) w.Write(out[i].nbits, out[i].bits);
}

inline void WriteHostStream(BitWriter& w, size_t num_ctx, const std::vector<uint8_t>& cluster_of, uint32_t num_clusters,
                            const std::vector<std::pair<uint32_t, uint32_t>>& tokens) {
  std::vector<uint32_t> hist(static_cast<size_t>(num_clusters) * 256, 0);
  for (const auto& t : tokens) {
    uint32_t token, nbits, bits;
    HybridEncode420(t.second, &token, &nbits, &bits);
    JXLB_CHECK(token < 256, "token too large");
    hist[cluster_of[t.first] * 256 + token]++;
  }
  (void)num_ctx;
  EncCode code;
  WriteCodeHeader(w, cluster_of, num_clusters, hist.data(), &code);
  WriteHostTokens(w, code, cluster_of, tokens);
}

inline void WriteContextMap(BitWriter& w, const std::vector<uint8_t>& cluster_of, uint32_t num_clusters) {
  if (num_clusters == 1) {
    w.Write(1, 1);  // simple
    w.Write(2, 0);  // zero bits per entry
    return;
  }
  if (num_clusters <= 8 && cluster_of.size() < 64) {
    const unsigned bits = CeilLog2(num_clusters);
    w.Write(1, 1);
    w.Write(2, bits);
    for (uint8_t c : cluster_of) w.Write(bits, c);
    return;
  }
  w.Write(1, 0);  // not simple
  w.Write(1, 0);  // no move-to-front
  std::vector<std::pair<uint32_t, uint32_t>> toks;
  toks.reserve(cluster_of.size());
  for (uint8_t c : cluster_of) toks.push_back({0, c});
  WriteHostStream(w, 1, std::vector<uint8_t>(1, 0), 1, toks);
}

// ---------------------------------------------------------------- the global Modular tree
// Root: stream id (static property 1) separates the DC streams (fixed gradient tree on property 9,
// lib/jxl/modular/encoding/enc_encoding.cc:274-282 kGradientFixedDC) from the AC-metadata streams (kACMeta, :218-265).
struct EncTree {
  std::vector<DevEncTreeNode> nodes;  // breadth-first, leaves numbered in that order
  std::vector<std::pair<uint32_t, uint32_t>> tokens;
  uint32_t num_leaves = 0;
};

inline EncTree BuildEncTree(uint32_t num_dc_groups) {
  struct N { int prop = -1; int32_t split = 0; int l = -1, r = -1; uint32_t pred = 0; };
  std::vector<N> n;
  auto leaf = [&](uint32_t pred) {
    n.push_back(N());
    n.back().pred = pred;
    return static_cast<int>(n.size() - 1);
  };
  auto split = [&](int prop, int32_t val, int l, int r) {
    N e;
    e.prop = prop;
    e.split = val;
    e.l = l;
    e.r = r;
    n.push_back(e);
    return static_cast<int>(n.size() - 1);
  };
  static const int32_t kCutoffs[33] = {-500, -392, -255, -191, -127, -95, -63, -47, -31, -23, -15, -11, -7, -4, -3, -1, 0,
                                       1, 3, 5, 7, 11, 15, 23, 31, 47, 63, 95, 127, 191, 255, 392, 500};
  std::function<int(size_t, size_t)> fixed = [&](size_t begin, size_t end) -> int {
    const int id = static_cast<int>(n.size());
    n.push_back(N());
    if (begin >= end) {
      n[id].pred = 5;  // Gradient
      return id;
    }
    const size_t mid = (begin + end) / 2;
    n[id].prop = 9;
    n[id].split = kCutoffs[mid];
    const int l = fixed(mid + 1, end);
    const int r = fixed(begin, mid);
    n[id].l = l;
    n[id].r = r;
    return id;
  };
  const int dc = fixed(0, 33);
  auto four = [&](int prop, uint32_t pred) {
    const int hi = split(prop, 11, leaf(pred), leaf(pred));
    const int lo = split(prop, 3, leaf(pred), leaf(pred));
    return split(prop, 5, hi, lo);
  };
  const int qf = four(7, 1 /* Left */), acs = four(7, 0 /* Zero */);
  const int acs_qf = split(2, 0, qf, acs);
  const int epf_hi = split(7, 3, leaf(0), leaf(0)), epf_lo = split(7, 3, leaf(0), leaf(0));
  const int epf = split(6, 3, epf_hi, epf_lo);
  const int c23 = split(0, 2, epf, acs_qf);
  const int c01 = split(0, 0, leaf(5), leaf(5));
  const int meta = split(0, 1, c23, c01);
  const int root = split(1, static_cast<int32_t>(num_dc_groups), meta, dc);
  EncTree t;
  std::vector<int> queue = {root};
  for (size_t k = 0; k < queue.size(); k++) {
    const N& e = n[queue[k]];
    DevEncTreeNode d{};
    if (e.prop < 0) {
      d.prop = -1;
      d.a = static_cast<int32_t>(e.pred);
      d.l = t.num_leaves++;
      t.tokens.push_back({1, 0});
      t.tokens.push_back({2, e.pred});
      t.tokens.push_back({3, 0});
      t.tokens.push_back({4, 0});
      t.tokens.push_back({5, 0});
    } else {
      d.prop = e.prop;
      d.a = e.split;
      d.l = queue.size();
      d.r = queue.size() + 1;
      queue.push_back(e.l);
      queue.push_back(e.r);
      t.tokens.push_back({1, static_cast<uint32_t>(e.prop + 1)});
      t.tokens.push_back({0, PackSignedH(e.split)});
    }
    t.nodes.push_back(d);
  }
  return t;
}

// Static clustering of the 495 * 15 AC contexts of the default block context map.
inline std::vector<uint8_t> AcContextClusters(uint32_t* num_clusters) {
  const uint32_t num_ctxs = 15, n = num_ctxs * 495;
  std::vector<uint8_t> cl(n, 0);
  const uint32_t nz_end = num_ctxs * 37;
  for (uint32_t ctx = 0; ctx < n; ctx++) {
    if (ctx < nz_end) {
      const uint32_t bucket = ctx / num_ctxs, block_ctx = ctx % num_ctxs;
      cl[ctx] = static_cast<uint8_t>((block_ctx < 7 ? 0 : 10) + std::min<uint32_t>(9, bucket / 4));
    } else {
      const uint32_t rel = ctx - nz_end;
      const uint32_t block_ctx = rel / 458, zdc = rel % 458;
      const uint32_t idx = zdc >> 1, prev = zdc & 1;
      cl[ctx] = static_cast<uint8_t>(20 + (block_ctx < 7 ? 0 : 60) + std::min<uint32_t>(29, idx / 8) * 2 + prev);
    }
  }
  std::vector<int> remap(256, -1);
  int next = 0;
  for (uint8_t& c : cl) {
    if (remap[c] < 0) remap[c] = next++;
    c = static_cast<uint8_t>(remap[c]);
  }
  *num_clusters = next;
  return cl;
}

// ---------------------------------------------------------------- headers
struct EncParams {
  float distance = 1.0f;
  int strategy_mode = 2;  // 0: DCT8 only, 2: variance heuristic over {64x64, 32x32, 16x16, 16x8, 8x16, 8x8}
  bool gab = true;
  uint32_t epf_iters = 2;
  bool dc_smoothing = true;
  uint32_t x_qm_scale = 3, b_qm_scale = 2;
  bool coeff_orders = true;  // coefficient orders from zero counts (lib/jxl/enc_coeff_order.cc); false: natural orders
  bool cfl = true;           // chroma-from-luma factors per tile (lib/jxl/enc_chroma_from_luma.cc); false: zero
  // libjxl's effort-7 quantiser: InitialQuantField (lib/jxl/enc_adaptive_quantization.cc), global scale and DC quantiser
  // from ComputeGlobalScaleAndQuant (lib/jxl/quantizer.cc:39-69), AdjustQuantField; false: raw quant 16 everywhere
  bool adaptive_quant = true;
  // the input has a fourth, alpha channel: an 8-bit extra channel coded losslessly next to the lossy colour
  bool alpha = false;
};

inline void WriteImageHeaders(BitWriter& w, uint32_t xsize, uint32_t ysize, bool alpha = false) {
  w.Write(16, 0x0AFF);
  w.Write(1, 0);  // not "small"
  WriteU32(w, ysize, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  w.Write(3, 0);  // no fixed aspect ratio
  WriteU32(w, xsize, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  if (!alpha) {
    w.Write(1, 1);  // ImageMetadata all_default: 8-bit sRGB, XYB encoded
  } else {          // the same with one extra channel (lib/jxl/image_metadata.cc:258-330)
    w.Write(1, 0);  // not all_default
    w.Write(1, 0);  // no extra_fields
    w.Write(1, 0);  // integer samples
    WriteU32(w, 8, Val(8), Val(10), Val(12), BitsOffset(6, 1));
    w.Write(1, 1);  // modular_16_bit_buffer_sufficient
    WriteU32(w, 1, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(12, 1));
    w.Write(1, 1);  // ExtraChannelInfo all_default: 8-bit alpha
    w.Write(1, 1);  // xyb_encoded
    w.Write(1, 1);  // ColorEncoding all_default (sRGB)
    WriteU64(w, 0);  // extensions
  }
  w.Write(1, 1);  // CustomTransformData all_default
  w.ZeroPadToByte();
}

inline void WriteFrameHeader(BitWriter& w, const EncParams& p) {
  w.Write(1, 0);  // not all_default
  w.Write(2, kRegularFrame);
  w.Write(1, 0);  // VarDCT
  WriteU64(w, p.dc_smoothing ? uint64_t{0} : uint64_t{kFlagSkipAdaptiveDCSmoothing});
  WriteU32(w, 1, Val(1), Val(2), Val(4), Val(8));  // upsampling
  if (p.alpha) WriteU32(w, 1, Val(1), Val(2), Val(4), Val(8));  // of the extra channel
  w.Write(3, p.x_qm_scale);
  w.Write(3, p.b_qm_scale);
  WriteU32(w, 1, Val(1), Val(2), Val(3), BitsOffset(3, 4));  // one pass
  w.Write(1, 0);  // no custom size or origin
  WriteU32(w, 0, Val(0), Val(1), Val(2), BitsOffset(2, 3));  // blend mode kReplace
  if (p.alpha) WriteU32(w, 0, Val(0), Val(1), Val(2), BitsOffset(2, 3));  // of the extra channel
  w.Write(1, 1);  // is_last
  WriteU32(w, 0, Val(0), Bits(4), BitsOffset(5, 16), BitsOffset(10, 48));  // name
  if (p.gab && p.epf_iters == 2) {
    w.Write(1, 1);  // loop filter all_default
  } else {
    w.Write(1, 0);
    w.Write(1, p.gab ? 1 : 0);
    if (p.gab) w.Write(1, 0);  // default weights
    w.Write(2, p.epf_iters);
    if (p.epf_iters > 0) {
      w.Write(1, 0);  // epf_sharp_custom
      w.Write(1, 0);  // epf_weight_custom
      w.Write(1, 0);  // epf_sigma_custom
    }
    WriteU64(w, 0);  // loop-filter extensions
  }
  WriteU64(w, 0);  // frame-header extensions
}

inline void WriteToc(BitWriter& w, const std::vector<size_t>& sizes) {
  w.Write(1, 0);  // not permuted
  w.ZeroPadToByte();
  for (size_t s : sizes) WriteU32(w, s, Bits(10), BitsOffset(14, 1024), BitsOffset(22, 17408), BitsOffset(30, 4211712));
  w.ZeroPadToByte();
}

// InitialQuantDC, lib/jxl/enc_adaptive_quantization.cc:1255-1267
inline float InitialQuantDC(float butteraugli_target) {
  const float kDcMul = 0.3f, kDcQuantPow = 0.83f, kDcQuant = 1.095924047623553f;
  const float butteraugli_target_dc = std::max<float>(
      0.5f * butteraugli_target, std::min<float>(butteraugli_target, kDcMul * std::pow((1.0f / kDcMul) * butteraugli_target, kDcQuantPow)));
  return std::min(kDcQuant / butteraugli_target_dc, 50.f);
}

// Quantiser parameters of a frame. Adaptive: libjxl's effort-7 choice, lib/jxl/enc_heuristics.cc:1055, :1104-1117
// (ComputeGlobalScaleAndQuant(InitialQuantDC(d), 0.39 / d, 0), lib/jxl/quantizer.cc:39-69); else the derivation of the
// oracle's plain stream generator (quant ~ 0.79 / distance around a raw quant of 16).
inline void FillQuantizer(const EncParams& p, DevEFrame* ef, uint32_t* global_scale_out, uint32_t* quant_dc_out) {
  const float quant_ac = 0.79f / std::max(0.1f, p.distance);
  const int base_raw = 16;
  int global_scale = std::max(1, std::min(65535 + 8192, static_cast<int>(quant_ac * 65536 / base_raw + 0.5f)));
  int quant_dc = std::max(1, std::min(65536, static_cast<int>(0.9f / std::max(0.1f, p.distance) * 65536 / global_scale + 0.5f)));
  ef->adaptive = p.adaptive_quant ? 1 : 0;
  if (p.adaptive_quant) {
    const float quant_dc_f = InitialQuantDC(p.distance);
    const float q = 0.39 / p.distance;
    float scale = 65536 * (q - 0.0f) / 5.0f;
    if (scale < 1) scale = 1;
    if (scale > (1 << 15)) scale = 1 << 15;
    int new_global_scale = static_cast<int>(scale);
    const int scaled_quant_dc = static_cast<int>(quant_dc_f * 4096 * 1.6);
    if (new_global_scale > scaled_quant_dc) {
      new_global_scale = scaled_quant_dc;
      if (new_global_scale <= 0) new_global_scale = 1;
    }
    global_scale = new_global_scale;
    const float inv_gs = static_cast<float>(1.0 * 65536 / new_global_scale);
    float fval = quant_dc_f * inv_gs + 0.5f;
    fval = std::min<float>(1 << 16, fval);
    quant_dc = static_cast<int>(fval);
    // InitialQuantField / PerBlockModulations / FuzzyErosion constants (lib/jxl/enc_adaptive_quantization.cc:306-321,
    // :380-410, :1269-1276), AdjustQuantField's mixer (:1211-1223)
    const float target = p.gab ? p.distance : p.distance * 0.62f;
    ef->aq_target = target;
    const float aq_scale = 0.725f / target * 1.0f;
    const float base_level = 0.48f * aq_scale;
    float dampen = 1.0f;
    if (target >= 2.0f) {
      dampen = 1.0f - ((target - 2.0f) / (14.0f - 2.0f));
      if (dampen < 0) dampen = 0;
    }
    ef->aq_mul = aq_scale * dampen;
    ef->aq_add = (1.0f - dampen) * base_level;
    float fmul = 0.0f;
    if (target < 2.0f) fmul = (2.0f - target) * (1.0f / 2.0f);
    float k0 = 0.125f + fmul * 0.0f, k1 = 0.10f + fmul * -0.10f, k2 = 0.09f + fmul * -0.09f, k3 = 0.06f + fmul * -0.06f;
    const float kTotal = 0.29959705784054957f;
    const float norm = kTotal / (k0 + k1 + k2 + k3);
    ef->aq_erosion[0] = k0 * norm;
    ef->aq_erosion[1] = k1 * norm;
    ef->aq_erosion[2] = k2 * norm;
    ef->aq_erosion[3] = k3 * norm;
    float mixer = 1.0f;
    if (p.distance > 1.54138f) mixer = std::max(0.0f, mixer - (p.distance - 1.54138f) * 0.56391f);
    ef->aq_mixer = mixer;
  }
  const float inv_global_scale = 1.0 * 65536 / global_scale;
  const float dc_quant[3] = {1.0f / 4096, 1.0f / 512, 1.0f / 256};
  for (int c = 0; c < 3; c++) ef->mul_dc[c] = (inv_global_scale / quant_dc) * dc_quant[c];
  ef->inv_global_scale = inv_global_scale;
  ef->x_dm = std::pow(1 / (1.25f), p.x_qm_scale - 2.0f);
  ef->b_dm = std::pow(1 / (1.25f), p.b_qm_scale - 2.0f);
  ef->x_qm_mul = std::pow(1.25f, p.x_qm_scale - 2.0f);
  ef->b_qm_mul = std::pow(1.25f, p.b_qm_scale - 2.0f);
  ef->distance = p.distance;
  ef->strategy_mode = p.strategy_mode;
  for (int i = 0; i < 4; i++) ef->biases[i] = kDefaultQuantBias[i];
  // chroma-from-luma fit: Quantizer::Scale() * kStrangeMultiplier * raw quant (lib/jxl/enc_chroma_from_luma.cc:311-316)
  ef->cfl = p.cfl ? 1 : 0;
  ef->cfl_scale128 = (global_scale * (1.0f / 65536)) * 128.0f;
  // inverse Gaborish weights (lib/jxl/enc_gaborish.cc:21-48 with mul = 1, as lib/jxl/enc_heuristics.cc:1121-1131 passes)
  ef->gab = p.gab ? 1 : 0;
  {
    static const float kGaborish[5] = {-0.09495815671340026, -0.041031725066768575, 0.013710004822696948,
                                       0.006510206083837737, -0.0014789063378272242};
    const float mul = 1.0f;
    double sum = 1.0 + mul * 4 * (kGaborish[0] + kGaborish[1] + kGaborish[2] + kGaborish[4] + 2 * kGaborish[3]);
    if (sum < 1e-5) sum = 1e-5;
    const float normalize = static_cast<float>(1.0 / sum);
    const float nm = mul * normalize;
    const float w[6] = {normalize, nm * kGaborish[0], nm * kGaborish[2], nm * kGaborish[1], nm * kGaborish[4], nm * kGaborish[3]};
    for (int i = 0; i < 6; i++) ef->gabinv_w[i] = w[i];
  }
  *global_scale_out = global_scale;
  *quant_dc_out = quant_dc;
}

// Arena layout of one frame (offsets relative to the frame's base in each arena).
struct EncLayout {
  FrameDimensions dim;
  uint32_t num_ac_clusters = 0, num_leaves = 0;
  uint64_t fsize = 0, isize = 0, bsize = 0, tsize = 0;  // floats, int32s, bytes, tokens
};

inline EncLayout LayoutEncFrame(uint32_t xsize, uint32_t ysize, uint32_t num_ac_clusters, uint32_t num_leaves, DevEFrame* ef,
                                bool alpha = false) {
  EncLayout L;
  FrameHeader fh;
  fh.xsize = xsize;
  fh.ysize = ysize;
  L.dim = ToFrameDimensions(fh);
  const FrameDimensions& d = L.dim;
  const uint64_t W = d.xsize_blocks, H = d.ysize_blocks, px = W * 8 * H * 8, nb = W * H;
  ef->xsize = xsize;
  ef->ysize = ysize;
  ef->xblocks = W;
  ef->yblocks = H;
  ef->xgroups = d.xsize_groups;
  ef->ygroups = d.ysize_groups;
  ef->xdcgroups = d.xsize_dc_groups;
  ef->ydcgroups = d.ysize_dc_groups;
  uint64_t f = 0, i = 0;
  for (int c = 0; c < 3; c++) {
    ef->xyb[c] = f;
    f += px;
  }
  for (int c = 0; c < 3; c++) {  // (only used when the frame signals Gaborish)
    ef->xyb_raw[c] = f;
    f += px;
  }
  ef->quant_field = f;
  f += (nb + 15) & ~uint64_t{15};
  ef->adj_thres = f;
  f += 4 * ((nb + 15) & ~uint64_t{15});
  for (int c = 0; c < 3; c++) {
    ef->coef[c] = i;
    i += px;
  }
  for (int c = 0; c < 3; c++) {
    ef->dcq[c] = i;
    i += nb;
  }
  ef->first_index = i; i += nb;
  ef->adj_quant = i; i += nb;
  ef->block_of_num = i; i += d.num_dc_groups * 65536;
  for (int c = 0; c < 3; c++) {
    ef->blk_nz[c] = i; i += nb;
    ef->blk_ntok[c] = i; i += nb;
    ef->blk_bucket[c] = i; i += nb;
  }
  // (the host reads everything from here to the end of the frame's region back after the tokenisation kernels)
  ef->order_mask = i; i += 1;
  ef->group_first = i; i += d.num_groups;
  ef->zero_counts = i; i += kCustomOrderCounters;
  for (uint32_t& o : ef->custom_order) o = 0xFFFFFFFFu;
  ef->dcg_count = i; i += d.num_dc_groups;
  ef->group_tokens = i; i += d.num_groups;
  ef->ac_hist = i; i += static_cast<uint64_t>(num_ac_clusters) * 256;
  ef->mod_hist = i; i += static_cast<uint64_t>(num_leaves) * 256;
  ef->acs = 0;
  ef->ac_tokens = 0;
  ef->mod_tokens = d.num_groups * 3 * 65536;
  ef->mod_tokens_stride = 6 * 65536 + 2048;
  L.fsize = f;
  L.isize = (i + 3) & ~uint64_t{3};
  ef->cmw = static_cast<uint32_t>((W + 7) / 8);
  ef->cmh = static_cast<uint32_t>((H + 7) / 8);
  ef->ytox = (nb + 15) & ~uint64_t{15};
  ef->ytob = ef->ytox + ((static_cast<uint64_t>(ef->cmw) * ef->cmh + 15) & ~uint64_t{15});
  ef->raw_quant = ef->ytob + ((static_cast<uint64_t>(ef->cmw) * ef->cmh + 15) & ~uint64_t{15});
  L.bsize = ef->raw_quant + ((nb + 15) & ~uint64_t{15});
  L.tsize = ef->mod_tokens + ef->mod_tokens_stride * d.num_dc_groups;
  ef->has_alpha = alpha ? 1 : 0;
  ef->alpha_tokens = L.tsize;
  if (alpha) L.tsize += static_cast<uint64_t>(d.num_groups) * 65536;
  L.num_ac_clusters = num_ac_clusters;
  L.num_leaves = num_leaves;
  return L;
}

// ---- coefficient orders (lib/jxl/enc_coeff_order.cc)
// Which varblocks enter the statistics when only every other one does (:92-111): xorshift128+ with libjxl's seed,
// one draw per varblock in group order.
inline std::vector<uint8_t> MakeOrderSampleBits(size_t n) {
  std::vector<uint8_t> bits(n);
  const uint64_t threshold = static_cast<uint64_t>((std::numeric_limits<uint64_t>::max() >> 32) * 0.5);
  uint64_t s[2] = {0x94D049BB133111EBull, 0xBF58476D1CE4E5B9ull};
  for (size_t i = 0; i < n; i++) {
    uint64_t s1 = s[0];
    const uint64_t s0 = s[1];
    const uint64_t b = s1 + s0;
    s[0] = s0;
    s1 ^= s1 << 23;
    s1 ^= s0 ^ (s1 >> 18) ^ (s0 >> 5);
    s[1] = s1;
    bits[i] = (b >> 32) <= threshold ? 1 : 0;
  }
  return bits;
}

struct CustomOrders {
  uint32_t used = 0;                                // bit ord: the order is transmitted
  std::vector<uint16_t> order[kNumCustomOrders][3];  // coefficient positions in scan order
};

// The sort of ComputeCoeffOrder (:160-238): positions in natural order, stably sorted by their quantised zero count.
// `mask` = orders that occur in the frame, `zero_counts` = the device's counters.
inline CustomOrders ComputeCustomOrders(uint32_t mask, const int32_t* zero_counts, uint32_t xblocks, uint32_t yblocks) {
  CustomOrders co;
  uint32_t customize = mask & ((1u << kNumCustomOrders) - 1);
  if (xblocks < 5 && yblocks < 5) customize = 0;  // default orders for small images (:72-74)
  const SharedVarDCTTables& sh = SharedVarDCTTables::Get();
  for (uint32_t ord = 0; ord < kNumCustomOrders; ord++) {
    if (!(customize & (1u << ord))) continue;
    const StrategyInfo si = GetStrategyInfo(kOrderFirstStrategy[ord]);
    const uint32_t sz = CustomOrderSize(ord);
    uint32_t lcx = si.cx, lcy = si.cy;
    if (lcy > lcx) std::swap(lcx, lcy);
    const uint16_t* natural = sh.opool.data() + sh.order_off[ord];
    bool is_nondefault = false;
    for (uint32_t c = 0; c < 3; c++) {
      struct PosAndCount { uint32_t pos, count; };
      std::vector<PosAndCount> pv(sz);
      const float inv_sqrt_sz = 1.0f / std::sqrt(static_cast<float>(sz));
      for (uint32_t i = 0; i < sz; i++) {
        const uint32_t pos = natural[i];
        const bool llf = pos % (8 * lcx) < lcx && pos / (8 * lcx) < lcy;
        const int32_t nz = llf ? -1 : zero_counts[CustomOrderBase(ord) + c * sz + pos];
        const float q = nz * inv_sqrt_sz + 0.1f;
        pv[i].pos = pos;
        pv[i].count = q <= 0.0f ? 0u : static_cast<uint32_t>(q);
      }
      std::stable_sort(pv.begin(), pv.end(), [](const PosAndCount& a, const PosAndCount& b) { return a.count < b.count; });
      co.order[ord][c].resize(sz);
      for (uint32_t i = 0; i < sz; i++) {
        co.order[ord][c][i] = static_cast<uint16_t>(pv[i].pos);
        is_nondefault |= natural[i] != pv[i].pos;
      }
    }
    if (is_nondefault) {
      co.used |= 1u << ord;
    } else {
      for (uint32_t c = 0; c < 3; c++) co.order[ord][c].clear();
    }
  }
  return co;
}

// EncodeCoeffOrders (:293-337): the permutations relative to the natural orders as Lehmer codes, one ANS stream.
inline void WriteCoeffOrders(BitWriter& w, const CustomOrders& co) {
  WriteU32(w, co.used, Val(0x5F), Val(0x13), Val(0), Bits(13));
  if (co.used == 0) return;
  const SharedVarDCTTables& sh = SharedVarDCTTables::Get();
  std::vector<std::pair<uint32_t, uint32_t>> tokens;
  for (uint32_t ord = 0; ord < kNumCustomOrders; ord++) {
    if (!(co.used & (1u << ord))) continue;
    const StrategyInfo si = GetStrategyInfo(kOrderFirstStrategy[ord]);
    const uint32_t llf = static_cast<uint32_t>(si.cx) * si.cy, sz = 64 * llf;
    const uint16_t* natural = sh.opool.data() + sh.order_off[ord];
    std::vector<uint32_t> lut(sz);
    for (uint32_t i = 0; i < sz; i++) lut[natural[i]] = i;
    for (uint32_t c = 0; c < 3; c++) {
      // Lehmer code (lib/jxl/lehmer_code.h): lehmer[i] = how many not yet used values are below value i,
      // = value - (used values below it), counted with a Fenwick tree
      std::vector<uint32_t> lehmer(sz), fen(sz + 1, 0);
      for (uint32_t i = 0; i < sz; i++) {
        const uint32_t v = lut[co.order[ord][c][i]];
        uint32_t below = 0;
        for (uint32_t k = v; k > 0; k -= k & (~k + 1)) below += fen[k];
        lehmer[i] = v - below;
        for (uint32_t k = v + 1; k <= sz; k += k & (~k + 1)) fen[k]++;
      }
      uint32_t end = sz;
      while (end > llf && lehmer[end - 1] == 0) end--;
      tokens.push_back({CoeffOrderContext(sz), end - llf});
      uint32_t last = 0;
      for (uint32_t i = llf; i < end; i++) {
        tokens.push_back({CoeffOrderContext(last), lehmer[i]});
        last = lehmer[i];
      }
    }
  }
  const std::vector<uint8_t> eight = {0, 1, 2, 3, 4, 5, 6, 7};
  WriteHostStream(w, 8, eight, 8, tokens);
}

// After the tokenisation kernels: the two host-written sections and the code tables for the emit kernels.
struct EncGlobals {
  BitWriter dc_global, ac_global;
  EncCode mod_code, ac_code;
};

inline void BuildEncGlobals(const EncParams& p, const EncLayout& L, const EncTree& tree, const std::vector<uint8_t>& ac_cluster_of,
                            uint32_t global_scale, uint32_t quant_dc, const uint32_t* mod_hist, const uint32_t* ac_hist,
                            const CustomOrders& orders, EncGlobals* g,
                            const std::vector<std::pair<uint32_t, uint32_t>>* global_alpha = nullptr) {
  BitWriter& d = g->dc_global;
  d.Write(1, 1);  // default DC quantisation
  WriteU32(d, global_scale, BitsOffset(11, 1), BitsOffset(11, 2049), BitsOffset(12, 4097), BitsOffset(16, 8193));
  WriteU32(d, quant_dc, Val(16), BitsOffset(5, 1), BitsOffset(8, 1), BitsOffset(16, 1));
  d.Write(1, 1);  // default block context map
  d.Write(1, 1);  // default colour correlation
  d.Write(1, 1);  // global MA tree
  std::vector<uint8_t> six = {0, 1, 2, 3, 4, 5};
  WriteHostStream(d, 6, six, 6, tree.tokens);
  std::vector<uint8_t> leaves(tree.num_leaves);
  for (uint32_t i = 0; i < tree.num_leaves; i++) leaves[i] = static_cast<uint8_t>(i);
  WriteCodeHeader(d, leaves, tree.num_leaves, mod_hist, &g->mod_code);
  if (p.alpha) {
    // the global Modular stream of the extra channel: GroupHeader (global tree, default weighted-predictor header, no
    // transforms); the alpha samples follow when the image fits one group (`global_alpha`), else they travel per AC group
    d.Write(4, 0x3);
    if (global_alpha != nullptr) WriteHostTokens(d, g->mod_code, leaves, *global_alpha);
  }
  BitWriter& a = g->ac_global;
  a.Write(1, 1);  // default quantisation matrices
  a.Write(CeilLog2(L.dim.num_groups), 0);  // one set of histograms
  WriteCoeffOrders(a, orders);
  WriteCodeHeader(a, ac_cluster_of, L.num_ac_clusters, ac_hist, &g->ac_code);
}

// A device-written section: `nbits` bits starting at bit `first` of `words`.
struct EncSection {
  const uint32_t* words;
  uint64_t first, nbits;
  uint64_t first2 = 0, nbits2 = 0;  // a second piece that follows the first bit-wise (the halves of a DC-group section)
};

// Final codestream from the sections.
inline std::vector<uint8_t> AssembleCodestream(const EncParams& p, const EncLayout& L, const EncGlobals& g,
                                               const std::vector<EncSection>& dcg, const std::vector<EncSection>& acg) {
  std::vector<std::vector<uint8_t>> sections;
  auto bytes_of = [](const EncSection& s) {
    BitWriter w;
    w.AppendWordBits(s.words, s.first, s.nbits);
    if (s.nbits2) w.AppendWordBits(s.words, s.first2, s.nbits2);
    w.ZeroPadToByte();
    return w.Bytes();
  };
  if (L.dim.num_groups == 1) {
    BitWriter all;
    all.AppendBits(g.dc_global.Bytes().data(), g.dc_global.BitsWritten());
    all.AppendWordBits(dcg[0].words, dcg[0].first, dcg[0].nbits);
    if (dcg[0].nbits2) all.AppendWordBits(dcg[0].words, dcg[0].first2, dcg[0].nbits2);
    all.AppendBits(g.ac_global.Bytes().data(), g.ac_global.BitsWritten());
    all.AppendWordBits(acg[0].words, acg[0].first, acg[0].nbits);
    sections.push_back(all.Bytes());
  } else {
    sections.push_back(g.dc_global.Bytes());
    for (const auto& s : dcg) sections.push_back(bytes_of(s));
    sections.push_back(g.ac_global.Bytes());
    for (const auto& s : acg) sections.push_back(bytes_of(s));
  }
  BitWriter out;
  WriteImageHeaders(out, L.dim.xsize, L.dim.ysize, p.alpha);
  WriteFrameHeader(out, p);
  std::vector<size_t> sizes;
  for (const auto& s : sections) sizes.push_back(s.size());
  WriteToc(out, sizes);
  for (const auto& s : sections) out.AppendBytes(s.data(), s.size());
  return out.Bytes();
}

inline float SrgbToLinearHost(float v) { return v <= 0.04045f ? v / 12.92f : std::pow((v + 0.055f) / 1.055f, 2.4f); }

}  // namespace jxlb

#endif  // JXLB_ENC_HOST_H_
