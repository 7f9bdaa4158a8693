// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
//
// A plain CPU VarDCT *encoder*: the stream generator behind the decoder parity tests and
// the 4K bench frames (the reference ships no VarDCT fixture larger than 40x50 and libjxl
// cannot be built here), and the CPU statement of the encode rows of SURVEY.md 8a (E1, E2,
// E8, E10-E14) that the CUDA encoder will be checked against. It writes conforming
// codestreams the way libjxl does, without libjxl's rate-distortion heuristics:
//   headers            lib/jxl/enc_fields.cc, lib/jxl/frame_header.cc:206-440, lib/jxl/enc_toc.cc
//   RGB -> XYB         lib/jxl/enc_xyb.cc:41-104, lib/jxl/cms/opsin_params.h
//   forward transforms lib/jxl/enc_transforms-inl.h (ComputeScaledDCT; the 8x8 special
//                      transforms by inverting the decoder's basis numerically)
//   quantisation       lib/jxl/enc_group.cc:46-90, :370-524 (no adaptive dead zone)
//   tokenisation       lib/jxl/enc_entropy_coder.cc:148-244 (mirror of DecodeACVarBlock)
//   histograms + rANS  lib/jxl/enc_ans.cc:253-364 (EncodeCounts), :1731-1816 (WriteTokens)
//   DC / AC metadata   lib/jxl/enc_modular.cc (global tree: fixed weighted-predictor DC tree + AC-metadata tree)
#ifndef JXLO_ENCODE_H_
#define JXLO_ENCODE_H_

#include <algorithm>
#include <functional>
#include <map>
#include <random>

#include "jxlo_render.h"
#include "jxlo_vardct.h"
#include "jxlo_enc_acs.h"

namespace jxlo {

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
  std::vector<uint8_t>& Bytes() { return bytes_; }
  void Append(const std::vector<uint8_t>& other) {
    ZeroPadToByte();
    bytes_.insert(bytes_.end(), other.begin(), other.end());
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
  throw Error("jxlo: value not representable in U32 field");
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
    w.Write(2, 3);
    w.Write(12, v & 4095);
    v >>= 12;
    unsigned shift = 12;
    while (v > 0 && shift < 60) {
      w.Write(1, 1);
      w.Write(8, v & 255);
      v >>= 8;
      shift += 8;
    }
    if (shift == 60) {
      if (v > 0) {
        w.Write(1, 1);
        w.Write(4, v & 15);
      } else {
        w.Write(1, 0);
      }
    } else {
      w.Write(1, 0);
    }
  }
}

inline void WriteF16(BitWriter& w, float f) {
  w.Write(16, FloatToHalf(f));
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

inline uint32_t PackSigned(int32_t v) { return (static_cast<uint32_t>(v) << 1) ^ (v < 0 ? 0xFFFFFFFFu : 0u); }

// ---------------------------------------------------------------- entropy coding
struct Token {
  uint32_t ctx, value;
};

struct EncodedUint {
  uint32_t token, nbits, bits;
};

inline EncodedUint EncodeHybrid(const HybridUintConfig& c, uint32_t v) {  // HybridUintConfig::Encode, dec_ans.h:73-90
  if (v < c.split_token) return {v, 0, 0};
  const uint32_t n = FloorLog2(v);
  const uint32_t m = v - (1u << n);
  EncodedUint e;
  e.token = c.split_token + ((n - c.split_exponent) << (c.msb_in_token + c.lsb_in_token)) +
            ((m >> (n - c.msb_in_token)) << c.lsb_in_token) + (m & ((1u << c.lsb_in_token) - 1));
  e.nbits = n - c.msb_in_token - c.lsb_in_token;
  e.bits = (m >> c.lsb_in_token) & ((1ull << e.nbits) - 1);
  return e;
}

// Normalises counts to a sum of 4096 keeping every used symbol (enc_ans.cc:119-219, simplified).
inline std::vector<int32_t> NormalizeCounts(const std::vector<uint64_t>& counts) {
  uint64_t total = 0;
  size_t used = 0;
  for (uint64_t c : counts) {
    total += c;
    used += c != 0;
  }
  std::vector<int32_t> out(counts.size(), 0);
  if (total == 0) {
    out.assign(1, kAnsTabSize);
    return out;
  }
  JXLO_CHECK(used <= kAnsTabSize, "too many symbols");
  int64_t sum = 0;
  size_t largest = 0;
  for (size_t i = 0; i < counts.size(); i++) {
    if (!counts[i]) continue;
    int32_t v = static_cast<int32_t>((counts[i] * kAnsTabSize + total / 2) / total);
    if (v < 1) v = 1;
    out[i] = v;
    sum += v;
    if (counts[i] > counts[largest] || !counts[largest]) largest = i;
  }
  int64_t diff = static_cast<int64_t>(kAnsTabSize) - sum;
  // spread the correction over the largest entries
  while (diff != 0) {
    size_t best = 0;
    for (size_t i = 0; i < out.size(); i++)
      if (out[i] > out[best]) best = i;
    const int64_t step = diff > 0 ? diff : std::max<int64_t>(diff, -(out[best] - 1));
    JXLO_CHECK(step != 0, "cannot normalise histogram");
    out[best] += static_cast<int32_t>(step);
    diff -= step;
  }
  return out;
}

// EncodeCounts-compatible writer (reads back through ReadAnsHistogram).
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
  // shift = 13: full precision for every count. log = 3 -> three 1 bits, then (shift + 1) - 8 in 3 bits.
  w.Write(3, 7);
  w.Write(3, 14 - 8);
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
    const int32_t c = counts[i];
    w.Write(bitcount, (c - (1 << (code - 1))) >> (code - 1 - bitcount));
  }
}

// ---- prefix codes, writer side (lib/jxl/enc_huffman.cc:17-214, lib/jxl/enc_huffman_tree.cc): code lengths by a
// length-limited Huffman construction, serialised the way ReadPrefixCode (jxlo_entropy.h) reads them -- "simple" codes
// for up to four symbols, otherwise the code-length code + run-length symbols 16 / 17.
inline std::vector<uint8_t> HuffmanLengths(const std::vector<uint64_t>& counts, int limit) {
  const size_t n = counts.size();
  std::vector<uint8_t> len(n, 0);
  std::vector<uint64_t> c(counts);
  for (;;) {
    struct Node { uint64_t w; int l, r; };
    std::vector<Node> nodes;
    std::vector<int> live;
    for (size_t s = 0; s < n; s++)
      if (c[s]) {
        nodes.push_back({c[s], -1, static_cast<int>(s)});
        live.push_back(static_cast<int>(nodes.size()) - 1);
      }
    if (live.size() < 2) {
      for (size_t s = 0; s < n; s++) len[s] = c[s] ? 1 : 0;
      return len;
    }
    while (live.size() > 1) {  // (ties: lower node index first, so the result is deterministic)
      std::stable_sort(live.begin(), live.end(), [&](int a, int b) { return nodes[a].w < nodes[b].w; });
      const int a = live[0], b = live[1];
      nodes.push_back({nodes[a].w + nodes[b].w, a, b});
      live.erase(live.begin(), live.begin() + 2);
      live.push_back(static_cast<int>(nodes.size()) - 1);
    }
    int max_len = 0;
    std::vector<std::pair<int, int>> stack = {{live[0], 0}};
    while (!stack.empty()) {
      const auto [id, depth] = stack.back();
      stack.pop_back();
      if (nodes[id].l < 0) {
        len[nodes[id].r] = static_cast<uint8_t>(depth);
        max_len = std::max(max_len, depth);
      } else {
        stack.push_back({nodes[id].l, depth + 1});
        stack.push_back({nodes[id].r, depth + 1});
      }
    }
    if (max_len <= limit) return len;
    for (auto& v : c)
      if (v) v = std::max<uint64_t>(1, v >> 1);  // flatten and try again
  }
}

// Canonical code words (bit-reversed: the stream is read LSB first), as BuildPrefixTable assigns them.
inline std::vector<uint32_t> CanonicalCodes(const std::vector<uint8_t>& len) {
  uint32_t count[kPrefixMaxBits + 2] = {0}, next[kPrefixMaxBits + 2] = {0};
  for (uint8_t l : len) count[l]++;
  count[0] = 0;
  uint32_t code = 0;
  for (int l = 1; l <= kPrefixMaxBits; l++) {
    code = (code + count[l - 1]) << 1;
    next[l] = code;
  }
  std::vector<uint32_t> out(len.size(), 0);
  for (size_t s = 0; s < len.size(); s++)
    if (len[s]) out[s] = ReverseBits(next[len[s]]++, len[s]);
  return out;
}

// Writes the prefix code of one cluster; `len` receives the code lengths the decoder will derive (0 bits per symbol for
// an alphabet of one). `counts.size()` is the alphabet size announced in the stream.
inline void WritePrefixCode(BitWriter& w, const std::vector<uint64_t>& counts, std::vector<uint8_t>* len) {
  const size_t alphabet = counts.size();
  len->assign(alphabet, 0);
  if (alphabet <= 1) return;
  std::vector<uint32_t> used;
  for (size_t s = 0; s < alphabet; s++)
    if (counts[s]) used.push_back(static_cast<uint32_t>(s));
  if (used.empty()) used.push_back(0);
  const uint32_t max_bits = FloorLog2(static_cast<uint32_t>(alphabet - 1)) + 1;
  if (used.size() <= 4) {  // simple code
    w.Write(2, 1);
    w.Write(2, used.size() - 1);
    // lengths by rank: 2 symbols 1,1; 3 symbols 1,2,2; 4 symbols 2,2,2,2 or 1,2,3,3 -- most frequent first
    std::vector<uint32_t> order(used);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return counts[a] > counts[b]; });
    bool tree_select = false;
    if (order.size() == 4) tree_select = counts[order[0]] > counts[order[2]] + counts[order[3]];
    if (order.size() == 4 && !tree_select) std::sort(order.begin(), order.end());
    if (order.size() == 3) std::sort(order.begin() + 1, order.end());
    if (order.size() == 4 && tree_select) std::sort(order.begin() + 2, order.end());
    if (order.size() == 2) std::sort(order.begin(), order.end());
    for (uint32_t sym : order) w.Write(max_bits, sym);
    if (order.size() == 4) w.Write(1, tree_select ? 1 : 0);
    static const uint8_t kLens[5][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {1, 1, 0, 0}, {1, 2, 2, 0}, {2, 2, 2, 2}};
    static const uint8_t kLens4[4] = {1, 2, 3, 3};
    for (size_t i = 0; i < order.size(); i++) (*len)[order[i]] = order.size() == 4 && tree_select ? kLens4[i] : kLens[order.size()][i];
    return;
  }
  *len = HuffmanLengths(counts, kPrefixMaxBits);
  // trailing zero lengths are implied once the code is complete
  size_t last = alphabet;
  while (last > 0 && (*len)[last - 1] == 0) last--;
  // run-length symbols: 16 = repeat the previous non-zero length 3..6 times, 17 = 3..10 zeros (one symbol per run;
  // consecutive 16s / 17s would extend the run multiplicatively, a literal in between keeps them independent)
  struct Rle { uint8_t sym, extra; };
  std::vector<Rle> seq;
  uint8_t prev = 8;
  for (size_t i = 0; i < last;) {
    const uint8_t v = (*len)[i];
    size_t run = 1;
    while (i + run < last && (*len)[i + run] == v) run++;
    const bool last_was_rle = !seq.empty() && seq.back().sym >= 16;
    if (v == 0 && run >= 3 && !(last_was_rle && seq.back().sym == 17)) {
      const size_t r = std::min<size_t>(run, 10);
      seq.push_back({17, static_cast<uint8_t>(r - 3)});
      i += r;
    } else if (v != 0 && v == prev && run >= 3 && !(last_was_rle && seq.back().sym == 16)) {
      const size_t r = std::min<size_t>(run, 6);
      seq.push_back({16, static_cast<uint8_t>(r - 3)});
      i += r;
    } else {
      seq.push_back({v, 0});
      if (v) prev = v;
      i++;
    }
  }
  std::vector<uint64_t> cl_counts(18, 0);
  for (const Rle& r : seq) cl_counts[r.sym]++;
  std::vector<uint8_t> cl = HuffmanLengths(cl_counts, 5);
  size_t cl_used = 0;
  for (uint8_t l : cl) cl_used += l != 0;
  static const uint8_t kOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  static const uint8_t kClBits[6] = {0, 7, 3, 2, 1, 15}, kClLen[6] = {2, 4, 3, 2, 2, 4};
  w.Write(2, 0);  // hskip
  int space = 32;
  for (int i = 0; i < 18 && space > 0; i++) {
    const uint8_t v = cl[kOrder[i]];
    w.Write(kClLen[v], kClBits[v]);
    if (v) space -= 32 >> v;
  }
  JXLO_CHECK(cl_used == 1 || space == 0, "internal: code-length code is not complete");
  const std::vector<uint32_t> cl_code = CanonicalCodes(cl);
  for (const Rle& r : seq) {
    if (cl_used > 1) w.Write(cl[r.sym], cl_code[r.sym]);
    if (r.sym == 16) w.Write(2, r.extra);
    if (r.sym == 17) w.Write(3, r.extra);
  }
}

struct EntropyOptions {
  bool use_prefix = false;   // prefix codes instead of ANS
  bool lz77 = false;         // LZ77 length / distance tokens (lib/jxl/enc_ans.cc:1040-1490, a greedy matcher here)
  uint32_t lz77_min_symbol = 224, lz77_min_length = 3;
};

// The 120 special distances of lib/jxl/dec_ans.h:121-143 (also in SymbolReader::SpecialDistance).
inline int LZ77SpecialDistance(uint32_t i, uint32_t dist_mult) {
  static const int8_t k[120][2] = {
      {0, 1},  {1, 0},  {1, 1},  {-1, 1}, {0, 2},  {2, 0},  {1, 2},  {-1, 2}, {2, 1},  {-2, 1}, {2, 2},  {-2, 2}, {0, 3},  {3, 0},  {1, 3},
      {-1, 3}, {3, 1},  {-3, 1}, {2, 3},  {-2, 3}, {3, 2},  {-3, 2}, {0, 4},  {4, 0},  {1, 4},  {-1, 4}, {4, 1},  {-4, 1}, {3, 3},  {-3, 3},
      {2, 4},  {-2, 4}, {4, 2},  {-4, 2}, {0, 5},  {3, 4},  {-3, 4}, {4, 3},  {-4, 3}, {5, 0},  {1, 5},  {-1, 5}, {5, 1},  {-5, 1}, {2, 5},
      {-2, 5}, {5, 2},  {-5, 2}, {4, 4},  {-4, 4}, {3, 5},  {-3, 5}, {5, 3},  {-5, 3}, {0, 6},  {6, 0},  {1, 6},  {-1, 6}, {6, 1},  {-6, 1},
      {2, 6},  {-2, 6}, {6, 2},  {-6, 2}, {4, 5},  {-4, 5}, {5, 4},  {-5, 4}, {3, 6},  {-3, 6}, {6, 3},  {-6, 3}, {0, 7},  {7, 0},  {1, 7},
      {-1, 7}, {5, 5},  {-5, 5}, {7, 1},  {-7, 1}, {4, 6},  {-4, 6}, {6, 4},  {-6, 4}, {2, 7},  {-2, 7}, {7, 2},  {-7, 2}, {3, 7},  {-3, 7},
      {7, 3},  {-7, 3}, {5, 6},  {-5, 6}, {6, 5},  {-6, 5}, {8, 0},  {4, 7},  {-4, 7}, {7, 4},  {-7, 4}, {8, 1},  {8, 2},  {6, 6},  {-6, 6},
      {8, 3},  {5, 7},  {-5, 7}, {7, 5},  {-7, 5}, {8, 4},  {6, 7},  {-6, 7}, {7, 6},  {-7, 6}, {8, 5},  {7, 7},  {-7, 7}, {8, 6},  {8, 7}};
  const int d = k[i][0] + static_cast<int>(dist_mult) * k[i][1];
  return d > 1 ? d : 1;
}

// A complete entropy-coded stream over `num_ctx` contexts: header (LZ77 parameters, context map, ANS histograms or
// prefix codes) + symbols. `cluster_of[ctx]` must use ids 0..n-1 without gaps. With LZ77 the distance context gets a
// cluster of its own.
class EntropyEncoder {
 public:
  EntropyEncoder(size_t num_ctx, std::vector<uint8_t> cluster_of, EntropyOptions opt = EntropyOptions())
      : num_ctx_(num_ctx), cluster_of_(std::move(cluster_of)), opt_(opt) {
    JXLO_CHECK(cluster_of_.size() == num_ctx_, "bad cluster map");
    num_clusters_ = 1;
    for (uint8_t c : cluster_of_) num_clusters_ = std::max<uint32_t>(num_clusters_, c + 1);
    if (opt_.lz77) {
      dist_cluster_ = num_clusters_++;
      cluster_of_.push_back(static_cast<uint8_t>(dist_cluster_));
    }
    cfg_ = HybridUintConfig(4, 2, 0);
    length_cfg_ = HybridUintConfig(0, 0, 0);
  }

  // Pass 1: statistics over every token that will be written with this code.
  // `dist_mult`: what the decoder will pass to its symbol reader for this stream (widest channel of a Modular stream,
  // else 0); only the LZ77 matcher looks at it.
  void Count(const std::vector<Token>& tokens, uint32_t dist_mult = 0) {
    if (hist_.empty()) hist_.assign(num_clusters_, std::vector<uint64_t>());
    auto add = [&](uint32_t cluster, uint32_t symbol) {
      std::vector<uint64_t>& h = hist_[cluster];
      if (h.size() <= symbol) h.resize(symbol + 1, 0);
      h[symbol]++;
    };
    if (!opt_.lz77) {
      for (const Token& t : tokens) add(cluster_of_[t.ctx], EncodeHybrid(cfg_, t.value).token);
      return;
    }
    for (const Item& it : Lz77(tokens, dist_mult)) {
      if (it.len == 0) {
        const uint32_t tok = EncodeHybrid(cfg_, it.value).token;
        JXLO_CHECK(tok < opt_.lz77_min_symbol, "literal token collides with the LZ77 length symbols");
        add(cluster_of_[it.ctx], tok);
      } else {
        add(cluster_of_[it.ctx], opt_.lz77_min_symbol + EncodeHybrid(length_cfg_, it.len - opt_.lz77_min_length).token);
        add(dist_cluster_, EncodeHybrid(cfg_, it.value).token);
      }
    }
  }

  void WriteHeader(BitWriter& w) {
    if (hist_.empty()) hist_.assign(num_clusters_, std::vector<uint64_t>());
    if (opt_.lz77) {
      w.Write(1, 1);
      WriteU32(w, opt_.lz77_min_symbol, Val(224), Val(512), Val(4096), BitsOffset(15, 8));
      WriteU32(w, opt_.lz77_min_length, Val(3), Val(4), BitsOffset(2, 5), BitsOffset(8, 9));
      WriteUintConfig(w, length_cfg_, 8);
    } else {
      w.Write(1, 0);  // no LZ77
    }
    if (cluster_of_.size() > 1) WriteContextMap(w);
    if (opt_.use_prefix) {
      w.Write(1, 1);
      for (uint32_t c = 0; c < num_clusters_; c++) WriteUintConfig(w, cfg_, kPrefixMaxBits);
      for (uint32_t c = 0; c < num_clusters_; c++) {  // alphabet sizes (VarLenUint16 of size - 1)
        const uint32_t v = static_cast<uint32_t>(std::max<size_t>(1, hist_[c].size())) - 1;
        if (v == 0) {
          w.Write(1, 0);
        } else {
          w.Write(1, 1);
          const unsigned n = FloorLog2(v);
          w.Write(4, n);
          w.Write(n, v - (1u << n));
        }
      }
      plen_.assign(num_clusters_, std::vector<uint8_t>());
      pcode_.assign(num_clusters_, std::vector<uint32_t>());
      for (uint32_t c = 0; c < num_clusters_; c++) {
        std::vector<uint64_t> counts = hist_[c];
        if (counts.empty()) counts.assign(1, 0);
        WritePrefixCode(w, counts, &plen_[c]);
        pcode_[c] = CanonicalCodes(plen_[c]);
      }
      return;
    }
    w.Write(1, 0);  // ANS, not prefix codes
    size_t max_alphabet = 1;
    for (auto& h : hist_) max_alphabet = std::max(max_alphabet, h.size());
    log_alpha_ = 5;
    while ((size_t{1} << log_alpha_) < max_alphabet) log_alpha_++;
    JXLO_CHECK(log_alpha_ <= 8, "alphabet too large");
    w.Write(2, log_alpha_ - 5);
    for (uint32_t c = 0; c < num_clusters_; c++) WriteUintConfig(w, cfg_, log_alpha_);
    const uint32_t ts = 1u << log_alpha_;
    alias_.assign(static_cast<size_t>(num_clusters_) * ts, AliasEntry{});
    freq_.assign(num_clusters_, std::vector<int32_t>());
    reverse_.assign(num_clusters_, std::vector<std::vector<uint16_t>>());
    for (uint32_t c = 0; c < num_clusters_; c++) {
      std::vector<int32_t> counts = NormalizeCounts(hist_[c]);
      WriteAnsHistogram(w, counts);
      freq_[c] = counts;
      BuildAliasTable(counts, log_alpha_, &alias_[static_cast<size_t>(c) * ts]);
      // reverse map from the decoder's own lookup (ANSBuildInfoTable, enc_ans.cc:44-68)
      reverse_[c].assign(counts.size(), std::vector<uint16_t>());
      for (size_t s = 0; s < counts.size(); s++) reverse_[c][s].assign(counts[s], 0);
      const uint32_t log_entry = kAnsLogTabSize - log_alpha_;
      for (uint32_t res = 0; res < kAnsTabSize; res++) {
        const AliasEntry& e = alias_[static_cast<size_t>(c) * ts + (res >> log_entry)];
        const uint32_t pos = res & ((1u << log_entry) - 1);
        const bool right = pos >= e.cutoff;
        const uint32_t sym = right ? e.right_value : (res >> log_entry);
        const uint32_t offset = (right ? e.offsets1 : 0) + pos;
        if (sym < reverse_[c].size() && offset < reverse_[c][sym].size()) reverse_[c][sym][offset] = res;
      }
    }
  }

  // Pass 2: one stream. ANS: 32-bit state + interleaved refills and raw bits, built back to front; prefix codes: the
  // symbols in order.
  void WriteTokens(BitWriter& w, const std::vector<Token>& tokens, uint32_t dist_mult = 0) const {
    // the symbols of the stream in decoding order: (cluster, symbol, raw bits)
    struct Sym { uint32_t cluster, symbol, nbits, bits; };
    std::vector<Sym> syms;
    syms.reserve(tokens.size());
    if (!opt_.lz77) {
      for (const Token& t : tokens) {
        const EncodedUint e = EncodeHybrid(cfg_, t.value);
        syms.push_back({cluster_of_[t.ctx], e.token, e.nbits, e.bits});
      }
    } else {
      for (const Item& it : Lz77(tokens, dist_mult)) {
        if (it.len == 0) {
          const EncodedUint e = EncodeHybrid(cfg_, it.value);
          syms.push_back({cluster_of_[it.ctx], e.token, e.nbits, e.bits});
        } else {
          const EncodedUint l = EncodeHybrid(length_cfg_, it.len - opt_.lz77_min_length);
          syms.push_back({cluster_of_[it.ctx], opt_.lz77_min_symbol + l.token, l.nbits, l.bits});
          const EncodedUint d = EncodeHybrid(cfg_, it.value);
          syms.push_back({dist_cluster_, d.token, d.nbits, d.bits});
        }
      }
    }
    if (opt_.use_prefix) {
      for (const Sym& s : syms) {
        JXLO_CHECK(s.symbol < plen_[s.cluster].size(), "token outside the prefix code");
        w.Write(plen_[s.cluster][s.symbol], pcode_[s.cluster][s.symbol]);
        w.Write(s.nbits, s.bits);
      }
      return;
    }
    struct Out { uint32_t nbits, bits; };
    std::vector<Out> out;
    out.reserve(syms.size() * 2);
    uint32_t state = kAnsSignature << 16;
    for (size_t i = syms.size(); i-- > 0;) {
      const Sym& s = syms[i];
      const uint32_t c = s.cluster;
      if (s.nbits) out.push_back({s.nbits, s.bits});
      JXLO_CHECK(s.symbol < freq_[c].size() && freq_[c][s.symbol] > 0, "token outside the histogram");
      const uint32_t f = freq_[c][s.symbol];
      if ((state >> (32 - kAnsLogTabSize)) >= f) {
        out.push_back({16, state & 0xFFFF});
        state >>= 16;
      }
      state = ((state / f) << kAnsLogTabSize) | reverse_[c][s.symbol][state % f];
    }
    w.Write(32, state);
    for (size_t i = out.size(); i-- > 0;) w.Write(out[i].nbits, out[i].bits);
  }

 private:
  // One literal (len == 0: ctx, value) or one copy (len >= min_length values, `value` = the distance symbol's value).
  struct Item { uint32_t ctx, value, len; };

  // Greedy matcher over the value sequence: at every position the longest match among a few candidate distances
  // (runs, the row above and its neighbours when the stream has a distance multiplier).
  std::vector<Item> Lz77(const std::vector<Token>& tokens, uint32_t mult) const {
    std::vector<Item> items;
    const size_t n = tokens.size();
    std::vector<uint32_t> cand = {1, 2, 3, 4, 7, 16};
    if (mult) {
      for (int dx = -2; dx <= 2; dx++) cand.push_back(mult + dx);
      cand.push_back(2 * mult);
    }
    for (size_t i = 0; i < n;) {
      size_t best_len = 0;
      uint32_t best_d = 0;
      for (uint32_t d : cand) {
        if (d == 0 || d > i || d > kLZ77Window) continue;
        size_t l = 0;
        while (i + l < n && l < 65535 && tokens[i + l].value == tokens[i + l - d].value) l++;
        if (l > best_len) {
          best_len = l;
          best_d = d;
        }
      }
      if (best_len >= std::max<uint32_t>(opt_.lz77_min_length, 3)) {
        uint32_t dsym;
        if (mult == 0) {
          dsym = best_d - 1;
        } else {
          dsym = best_d - 1 + kNumSpecialDistances;
          for (uint32_t k = 0; k < kNumSpecialDistances; k++)
            if (LZ77SpecialDistance(k, mult) == static_cast<int>(best_d)) {
              dsym = k;
              break;
            }
        }
        items.push_back({tokens[i].ctx, dsym, static_cast<uint32_t>(best_len)});
        i += best_len;
      } else {
        items.push_back({tokens[i].ctx, tokens[i].value, 0});
        i++;
      }
    }
    return items;
  }

  static void WriteUintConfig(BitWriter& w, const HybridUintConfig& c, uint32_t log_alpha) {
    w.Write(CeilLog2(log_alpha + 1), c.split_exponent);
    if (c.split_exponent == log_alpha) return;
    w.Write(CeilLog2(c.split_exponent + 1), c.msb_in_token);
    w.Write(CeilLog2(c.split_exponent - c.msb_in_token + 1), c.lsb_in_token);
  }

  void WriteContextMap(BitWriter& w) {
    const size_t n = cluster_of_.size();
    if (num_clusters_ == 1) {
      w.Write(1, 1);  // simple
      w.Write(2, 0);  // zero bits per entry
      return;
    }
    if (num_clusters_ <= 8 && n < 64) {
      const unsigned bits = CeilLog2(num_clusters_);
      w.Write(1, 1);
      w.Write(2, bits);
      for (uint8_t c : cluster_of_) w.Write(bits, c);
      return;
    }
    w.Write(1, 0);  // not simple
    w.Write(1, 0);  // no move-to-front
    std::vector<Token> toks;
    toks.reserve(n);
    for (uint8_t c : cluster_of_) toks.push_back({0, c});
    EntropyEncoder nested(1, std::vector<uint8_t>(1, 0));
    nested.Count(toks);
    nested.WriteHeader(w);
    nested.WriteTokens(w, toks);
  }

  size_t num_ctx_;
  std::vector<uint8_t> cluster_of_;
  EntropyOptions opt_;
  uint32_t num_clusters_ = 1, dist_cluster_ = 0;
  HybridUintConfig cfg_, length_cfg_;
  uint32_t log_alpha_ = 5;
  std::vector<std::vector<uint64_t>> hist_;
  std::vector<AliasEntry> alias_;
  std::vector<std::vector<int32_t>> freq_;
  std::vector<std::vector<std::vector<uint16_t>>> reverse_;
  std::vector<std::vector<uint8_t>> plen_;
  std::vector<std::vector<uint32_t>> pcode_;
};

// ---------------------------------------------------------------- Modular sub-streams (global tree)
// The DC and AC-metadata streams of a VarDCT frame share one global MA tree, like libjxl's
// encoder builds it at its default effort (lib/jxl/enc_modular.cc:1185-1215 and MergeTrees): the
// root splits on the stream id (static property 1); DC streams use the fixed weighted-predictor tree
// (PredefinedTree kWPFixedDC, lib/jxl/modular/encoding/enc_encoding.cc:66-97, :266-273) and AC-metadata
// streams the kACMeta tree (:218-265).
struct EncTreeNode {
  int property = -1;  // -1: leaf
  int32_t splitval = 0;
  int left = -1, right = -1;  // indices into the builder's node list (property > splitval ? left : right)
  uint32_t predictor = 0;
};

struct GlobalTree {
  Tree tree;                  // in decoder (breadth-first) order, leaves numbered in that order
  std::vector<Token> tokens;  // the tree itself, contexts of lib/jxl/modular/encoding/ma_common.h:13-22
  size_t num_leaves = 0;
};

inline int EncTreeFixed(std::vector<EncTreeNode>* n, int property, const std::vector<int32_t>& cutoffs, size_t begin,
                        size_t end, uint32_t predictor) {
  const int id = static_cast<int>(n->size());
  n->push_back(EncTreeNode());
  if (begin >= end) {
    (*n)[id].predictor = predictor;
    return id;
  }
  const size_t split = (begin + end) / 2;
  (*n)[id].property = property;
  (*n)[id].splitval = cutoffs[split];
  const int l = EncTreeFixed(n, property, cutoffs, split + 1, end, predictor);
  const int r = EncTreeFixed(n, property, cutoffs, begin, split, predictor);
  (*n)[id].left = l;
  (*n)[id].right = r;
  return id;
}

inline GlobalTree BuildGlobalTree(uint32_t num_dc_groups, int dc_tree = 0) {
  std::vector<EncTreeNode> n;
  auto leaf = [&](uint32_t pred) {
    n.push_back(EncTreeNode());
    n.back().predictor = pred;
    return static_cast<int>(n.size() - 1);
  };
  auto split = [&](int prop, int32_t val, int l, int r) {
    EncTreeNode e;
    e.property = prop;
    e.splitval = val;
    e.left = l;
    e.right = r;
    n.push_back(e);
    return static_cast<int>(n.size() - 1);
  };
  static const std::vector<int32_t> kCutoffs = {-500, -392, -255, -191, -127, -95, -63, -47, -31, -23, -15, -11, -7, -4, -3, -1, 0,
                                                1, 3, 5, 7, 11, 15, 23, 31, 47, 63, 95, 127, 191, 255, 392, 500};
  const int dc = dc_tree == 0 ? EncTreeFixed(&n, kWPProp, kCutoffs, 0, kCutoffs.size(), kPredWeighted)
                              : EncTreeFixed(&n, 9, kCutoffs, 0, kCutoffs.size(), kPredGradient);
  // AC metadata: channel 0 / 1 = chroma-from-luma maps, 2 = (strategy row, quant row), 3 = EPF sharpness
  auto four = [&](int prop, uint32_t pred) {  // splits at 11, 5, 3 on `prop`
    const int hi = split(prop, 11, leaf(pred), leaf(pred));
    const int lo = split(prop, 3, leaf(pred), leaf(pred));
    return split(prop, 5, hi, lo);
  };
  const int qf = four(7, kPredLeft), acs = four(7, kPredZero);
  const int acs_qf = split(2, 0, qf, acs);  // y > 0: quant field row
  const int epf_hi = split(7, 3, leaf(kPredZero), leaf(kPredZero)), epf_lo = split(7, 3, leaf(kPredZero), leaf(kPredZero));
  const int epf = split(6, 3, epf_hi, epf_lo);
  const int c23 = split(0, 2, epf, acs_qf);
  const int c01 = split(0, 0, leaf(kPredGradient), leaf(kPredGradient));
  const int meta = split(0, 1, c23, c01);
  const int root = split(1, static_cast<int32_t>(num_dc_groups), meta, dc);
  // breadth-first serialisation = the order ReadTree assigns child positions in
  GlobalTree g;
  std::vector<int> queue = {root};
  for (size_t k = 0; k < queue.size(); k++) {
    const EncTreeNode& e = n[queue[k]];
    TreeNode t{};
    if (e.property < 0) {
      t.property = -1;
      t.predictor = e.predictor;
      t.multiplier = 1;
      t.lchild = g.num_leaves++;
      g.tokens.push_back({1, 0});
      g.tokens.push_back({2, e.predictor});
      g.tokens.push_back({3, 0});
      g.tokens.push_back({4, 0});
      g.tokens.push_back({5, 0});
    } else {
      t.property = e.property;
      t.splitval = e.splitval;
      t.lchild = queue.size();
      t.rchild = queue.size() + 1;
      t.multiplier = 1;
      queue.push_back(e.left);
      queue.push_back(e.right);
      g.tokens.push_back({1, static_cast<uint32_t>(e.property + 1)});
      g.tokens.push_back({0, PackSigned(e.splitval)});
    }
    g.tree.push_back(t);
  }
  return g;
}

// Tokens of one Modular sub-stream under the global tree: the mirror of DecodeChannel (jxlo_modular.h).
inline void TokenizeModularStream(const Tree& tree, uint32_t stream_id, const std::vector<Channel>& channels,
                                  std::vector<Token>* toks) {
  const WPHeader wp_header;
  for (size_t chan = 0; chan < channels.size(); chan++) {
    const Channel& ch = channels[chan];
    if (ch.w == 0 || ch.h == 0) continue;
    std::vector<int32_t> props(kNumNonrefProps, 0);
    WPState wp(wp_header, ch.w);
    const int w = ch.w;
    for (int y = 0; y < ch.h; y++) {
      const int32_t* row = ch.Row(y);
      const int32_t* prev = y ? ch.Row(y - 1) : nullptr;
      const int32_t* prevprev = y > 1 ? ch.Row(y - 2) : nullptr;
      props[0] = static_cast<int32_t>(chan);
      props[1] = static_cast<int32_t>(stream_id);
      props[2] = y;
      props[9] = 0;
      for (int x = 0; x < w; x++) {
        const Neighbors n = LoadNeighbors(row, prev, prevprev, x, y, w);
        props[3] = x;
        props[4] = static_cast<int32_t>(n.top > 0 ? n.top : -n.top);
        props[5] = static_cast<int32_t>(n.left > 0 ? n.left : -n.left);
        props[6] = static_cast<int32_t>(n.top);
        props[7] = static_cast<int32_t>(n.left);
        props[8] = static_cast<int32_t>(n.left - props[9]);
        props[9] = static_cast<int32_t>(n.left + n.top - n.topleft);
        props[10] = static_cast<int32_t>(n.left - n.topleft);
        props[11] = static_cast<int32_t>(n.topleft - n.top);
        props[12] = static_cast<int32_t>(n.top - n.topright);
        props[13] = static_cast<int32_t>(n.top - n.toptop);
        props[14] = static_cast<int32_t>(n.left - n.leftleft);
        const int64_t wp_pred = wp.Predict(x, y, w, n.top, n.left, n.topright, n.topleft, n.toptop, &props[kWPProp]);
        size_t pos = 0;
        while (tree[pos].property >= 0) pos = props[tree[pos].property] > tree[pos].splitval ? tree[pos].lchild : tree[pos].rchild;
        const int64_t guess = PredictOne(tree[pos].predictor, n, wp_pred);
        toks->push_back({tree[pos].lchild, PackSigned(static_cast<int32_t>(row[x] - guess))});
        wp.Update(row[x], x, y, w);
      }
    }
  }
}

// GroupHeader of a sub-stream that uses the global tree, followed by its symbols.
inline void WriteModularStream(BitWriter& w, const EntropyEncoder& code, const std::vector<Channel>& channels,
                               const std::vector<Token>& toks) {
  bool any = false;
  for (const Channel& c : channels) any |= c.w > 0 && c.h > 0;
  if (channels.empty()) return;
  w.Write(1, 1);  // use_global_tree
  w.Write(1, 1);  // default weighted-predictor header
  w.Write(2, 0);  // no transforms
  if (!any) return;
  code.WriteTokens(w, toks);
}

// ---------------------------------------------------------------- forward transforms
// TransformFromPixels: pixels (stride) -> coefficients in the decoder's layout: libjxl's forward transforms
// (jxlo_enc_acs.h, lib/jxl/enc_transforms-inl.h:462-660).
inline void TransformFromPixels(int strategy, const float* pixels, size_t stride, float* coeffs, float* scratch) {
  TransformFromPixelsRef(strategy, pixels, stride, coeffs, scratch);
}

// ---------------------------------------------------------------- colour
inline float SrgbToLinear(float v) {
  return v <= 0.04045f ? v / 12.92f : std::pow((v + 0.055f) / 1.055f, 2.4f);
}

// CubeRootAndAdd, lib/jxl/base/fast_math-inl.h:177-224: cbrt(x) + add by Newton-Raphson on 1 / cbrt.
inline float CubeRootAndAdd(float x, float add) {
  const float k1_3 = 1.0f / 3, k4_3 = 4.0f / 3;
  const float xa_3 = k1_3 * x;
  int32_t m1;
  std::memcpy(&m1, &x, 4);
  const int32_t m2 = m1 == 0 ? 0 : 0x54800000 - (m1 >> 23) * 0x002AAAAA;
  float r;
  std::memcpy(&r, &m2, 4);
  for (int i = 0; i < 3; i++) {
    const float r2 = r * r;
    r = std::fmaf(-xa_3, r2 * r2, k4_3 * r);
  }
  float r2 = r * r;
  r = std::fmaf(k1_3, std::fmaf(-x, r2 * r2, r), r);
  r2 = r * r;
  return std::fmaf(r2, x, add);
}

constexpr float kNegOpsinBiasCbrt = -0.15595420054924863f;  // -cbrt(0.0037930732552754493)

// lib/jxl/enc_xyb.cc:41-104 with the default opsin matrix (intensity target 255).
inline void LinearRgbToXyb(float r, float g, float b, float* x, float* y, float* bb) {
  const float bias = 0.0037930732552754493f;
  const float kM[9] = {0.30f, 1.0f - 0.078f - 0.30f, 0.078f, 0.23f, 1.0f - 0.078f - 0.23f, 0.078f,
                       0.24342268924547819f, 0.20476744424496821f, 1.0f - 0.24342268924547819f - 0.20476744424496821f};
  float mixed[3];
  for (int i = 0; i < 3; i++) {
    mixed[i] = kM[3 * i] * r + kM[3 * i + 1] * g + kM[3 * i + 2] * b + bias;
    if (mixed[i] < 0) mixed[i] = 0;
    mixed[i] = CubeRootAndAdd(mixed[i], kNegOpsinBiasCbrt);
  }
  *x = 0.5f * (mixed[0] - mixed[1]);
  *y = 0.5f * (mixed[0] + mixed[1]);
  *bb = mixed[2];
}

// ---------------------------------------------------------------- the encoder
// Seeded random splines written into DC global (decoder coverage of lib/jxl/splines.cc): `count` splines inside an
// xsize x ysize frame, in the token order Splines::Decode reads (:607-650; QuantizedSpline::Decode :547-596).
inline void WriteRandomSplines(BitWriter& w, uint32_t count, uint32_t xsize, uint32_t ysize, uint32_t seed,
                               const EntropyOptions& eopt = EntropyOptions());

struct EncodeParams {
  float distance = 1.0f;
  // 0: DCT8 only; 1: seeded random mix of all 27 strategies (decoder coverage);
  // 2: variance heuristic over {8x8, 16x16, 32x32, 16x8, 8x16, 64x64}; 100 + s: strategy s wherever it fits;
  // 3: libjxl's effort-7 entropy-estimate search (jxlo_enc_acs.h: AcStrategyHeuristics::ProcessRect per 64x64 tile
  //    after the DCT8 chroma-from-luma pre-pass, lib/jxl/enc_heuristics.cc:1150-1166); needs adaptive_quant
  int strategy_mode = 2;
  uint32_t seed = 1;
  bool gab = true;
  uint32_t epf_iters = 2;
  bool dc_smoothing = true;
  bool random_side_info = false;  // random CfL factors, quant field and EPF sharpness (decoder coverage)
  uint32_t x_qm_scale = 3, b_qm_scale = 2;
  uint32_t num_passes = 1;        // > 1: coefficients split by magnitude shift (passes.shift)
  // DC streams: 0 = fixed weighted-predictor tree (libjxl's default effort, kWPFixedDC), 1 = fixed gradient tree
  // (kGradientFixedDC, what libjxl writes for decoding_speed_tier >= 1): contexts from property 9, no WP state.
  int dc_tree = 0;
  // Encoder-side heuristics that change which (valid) stream is written, not how a decoder reads it; the committed
  // 4K decode fixtures predate them and are regenerated with both off.
  bool gab_inverse = true;    // inverse Gaborish before the transforms when the frame signals Gaborish (E4)
  bool coeff_orders = true;   // coefficient orders from zero counts (E9); false: natural orders
  bool cfl = true;            // chroma-from-luma factors per 64x64 tile fitted to the AC coefficients (E6); false: 0
  // libjxl's effort-7 quantiser (E3, E7): InitialQuantField on the XYB planes before inverse Gaborish, global scale and
  // DC quantiser from ComputeGlobalScaleAndQuant(InitialQuantDC(d), 0.39 / d, 0), AdjustQuantField over each varblock,
  // SetQuantFieldRect. false (and with random_side_info): one constant raw quant value under this file's own scale.
  bool adaptive_quant = true;
  bool prefix_codes = false;  // every entropy-coded stream of the frame uses prefix codes instead of ANS (decoder coverage)
  // Frame upsampling (decoder coverage of lib/jxl/render_pipeline/stage_upsampling.cc): the pixels handed to the encoder
  // are the low-resolution frame, the image header declares `upsampling` times their size.
  uint32_t upsampling = 1;
  uint32_t orientation = 1;  // ImageMetadata::orientation (decoder coverage of the write stage's undo_orientation)
  // One 8-bit alpha extra channel (xsize * ysize samples), coded losslessly in the frame's Modular sub-streams under the
  // global tree the way libjxl places them (lib/jxl/enc_modular.cc:1258-1500, lib/jxl/dec_frame.cc:315-365, :478-560):
  // the global stream when the channel fits one group, else one stream per AC group behind that group's coefficients.
  const uint8_t* alpha = nullptr;
  uint32_t splines = 0;  // > 0: this many seeded random splines (frame flag kSplines, drawn over the decoded colour)
};

inline void WriteRandomSplines(BitWriter& w, uint32_t count, uint32_t xsize, uint32_t ysize, uint32_t seed, const EntropyOptions& eopt) {
  std::mt19937 rng(seed * 2654435761u + 97u);
  auto rnd = [&](int lo, int hi) { return lo + static_cast<int>(rng() % static_cast<uint32_t>(hi - lo + 1)); };
  std::vector<Token> t;
  t.push_back({2, count - 1});  // kNumSplinesContext
  int last_x = 0, last_y = 0;
  for (uint32_t i = 0; i < count; i++) {  // kStartingPositionContext
    const int x = rnd(4, static_cast<int>(xsize) - 5), y = rnd(4, static_cast<int>(ysize) - 5);
    if (i == 0) {
      t.push_back({1, static_cast<uint32_t>(x)});
      t.push_back({1, static_cast<uint32_t>(y)});
    } else {
      t.push_back({1, PackSigned(x - last_x)});
      t.push_back({1, PackSigned(y - last_y)});
    }
    last_x = x;
    last_y = y;
  }
  t.push_back({0, PackSigned(rnd(-2, 3))});  // kQuantizationAdjustmentContext
  for (uint32_t i = 0; i < count; i++) {
    const uint32_t n = static_cast<uint32_t>(rnd(1, 5));
    t.push_back({3, n});  // kNumControlPointsContext
    int dx = 0, dy = 0;
    for (uint32_t k = 0; k < n; k++) {  // double deltas; the running delta never becomes (0, 0)
      int ddx = rnd(-9, 9), ddy = rnd(-9, 9);
      if (dx + ddx == 0 && dy + ddy == 0) ddx += 3;
      dx += ddx;
      dy += ddy;
      t.push_back({4, PackSigned(ddx)});
      t.push_back({4, PackSigned(ddy)});
    }
    for (int c = 0; c < 4; c++)  // X, Y, B, sigma: a strong DC term, a few small harmonics (kDCTContext)
      for (int k = 0; k < 32; k++) {
        int v = 0;
        if (k == 0) v = c == 3 ? rnd(6, 30) : rnd(-60, 60);
        else if (k < 4) v = rnd(-3, 3);
        t.push_back({5, PackSigned(v)});
      }
  }
  EntropyEncoder code(6, {0, 1, 2, 3, 4, 5}, eopt);
  code.Count(t);
  code.WriteHeader(w);
  code.WriteTokens(w, t);
}

struct EncoderStats {
  size_t num_groups = 0, num_varblocks = 0, bytes = 0;
  size_t strategy_count[27] = {0};
};

inline void WriteImageHeaders(BitWriter& w, uint32_t xsize, uint32_t ysize, uint32_t orientation = 1, bool alpha = false) {
  w.Write(16, 0x0AFF);
  // SizeHeader
  w.Write(1, 0);  // not "small"
  WriteU32(w, ysize, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  w.Write(3, 0);  // no fixed aspect ratio
  WriteU32(w, xsize, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  if (orientation == 1 && !alpha) {
    w.Write(1, 1);  // ImageMetadata all_default: 8-bit sRGB, XYB encoded, no extra channels
  } else {          // the same image with an orientation (extra_fields) and / or an alpha channel (lib/jxl/image_metadata.cc:258-330)
    w.Write(1, 0);  // not all_default
    const bool extra_fields = orientation != 1;
    w.Write(1, extra_fields ? 1 : 0);
    if (extra_fields) {
      w.Write(3, orientation - 1);
      w.Write(1, 0);  // no intrinsic size
      w.Write(1, 0);  // no preview
      w.Write(1, 0);  // no animation
    }
    w.Write(1, 0);  // integer samples
    WriteU32(w, 8, Val(8), Val(10), Val(12), BitsOffset(6, 1));
    w.Write(1, 1);  // modular_16_bit_buffer_sufficient
    WriteU32(w, alpha ? 1 : 0, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(12, 1));
    if (alpha) w.Write(1, 1);  // ExtraChannelInfo all_default: 8-bit alpha
    w.Write(1, 1);  // xyb_encoded
    w.Write(1, 1);  // ColorEncoding all_default (sRGB)
    if (extra_fields) w.Write(1, 1);  // ToneMapping all_default
    WriteU64(w, 0);  // extensions
  }
  w.Write(1, 1);  // CustomTransformData all_default
  w.ZeroPadToByte();
}

inline void WriteFrameHeader(BitWriter& w, const EncodeParams& p) {
  w.Write(1, 0);  // not all_default
  w.Write(2, kRegularFrame);
  w.Write(1, 0);  // VarDCT
  WriteU64(w, (p.dc_smoothing ? uint64_t{0} : uint64_t{kFlagSkipAdaptiveDCSmoothing}) | (p.splines ? uint64_t{kFlagSplines} : uint64_t{0}));
  WriteU32(w, p.upsampling, Val(1), Val(2), Val(4), Val(8));  // upsampling
  if (p.alpha) WriteU32(w, p.upsampling, Val(1), Val(2), Val(4), Val(8));  // of the extra channel
  w.Write(3, p.x_qm_scale);
  w.Write(3, p.b_qm_scale);
  WriteU32(w, p.num_passes, Val(1), Val(2), Val(3), BitsOffset(3, 4));
  if (p.num_passes != 1) {
    WriteU32(w, 0, Val(0), Val(1), Val(2), BitsOffset(1, 3));  // num_downsample
    for (uint32_t i = 0; i + 1 < p.num_passes; i++) w.Write(2, p.num_passes - 1 - i);  // shift
  }
  w.Write(1, 0);  // no custom size or origin
  WriteU32(w, 0, Val(0), Val(1), Val(2), BitsOffset(2, 3));  // blend mode kReplace
  if (p.alpha) WriteU32(w, 0, Val(0), Val(1), Val(2), BitsOffset(2, 3));  // of the extra channel
  w.Write(1, 1);  // is_last
  WriteU32(w, 0, Val(0), Bits(4), BitsOffset(5, 16), BitsOffset(10, 48));  // name
  // LoopFilter
  if (p.gab && p.epf_iters == 2) {
    w.Write(1, 1);  // all_default
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

inline void WriteToc(BitWriter& w, const std::vector<std::vector<uint8_t>>& sections) {
  w.Write(1, 0);  // not permuted
  w.ZeroPadToByte();
  for (const auto& s : sections)
    WriteU32(w, s.size(), Bits(10), BitsOffset(14, 1024), BitsOffset(22, 17408), BitsOffset(30, 4211712));
  w.ZeroPadToByte();
}

// Static clustering of the 495 * 15 AC contexts of the default block context map.
inline std::vector<uint8_t> ACContextClusters(const BlockCtxMap& bctx) {
  const uint32_t n = bctx.NumACContexts();
  std::vector<uint8_t> cl(n, 0);
  const uint32_t nz_end = bctx.num_ctxs * kNonZeroBuckets;
  for (uint32_t ctx = 0; ctx < n; ctx++) {
    if (ctx < nz_end) {
      const uint32_t bucket = ctx / bctx.num_ctxs, block_ctx = ctx % bctx.num_ctxs;
      cl[ctx] = static_cast<uint8_t>((block_ctx < 7 ? 0 : 10) + std::min<uint32_t>(9, bucket / 4));
    } else {
      const uint32_t rel = ctx - nz_end;
      const uint32_t block_ctx = rel / kZeroDensityContextCount, zdc = rel % kZeroDensityContextCount;
      const uint32_t idx = zdc >> 1, prev = zdc & 1;
      cl[ctx] = static_cast<uint8_t>(20 + (block_ctx < 7 ? 0 : 60) + std::min<uint32_t>(29, idx / 8) * 2 + prev);
    }
  }
  // compact ids so that every cluster id is in use
  std::vector<int> remap(256, -1);
  int next = 0;
  for (uint8_t& c : cl) {
    if (remap[c] < 0) remap[c] = next++;
    c = static_cast<uint8_t>(remap[c]);
  }
  return cl;
}

// FindBestMultiplier, lib/jxl/enc_chroma_from_luma.cc:118-175 with `fast == false` (what libjxl's effort 7 passes,
// lib/jxl/enc_heuristics.cc:1176-1182): Newton iterations on f'(x), f = 1/3 sum((|colour residual| + 1)^2 - 1) +
// distance_mul * x^2 * num, derivatives by central differences (CFLFunction::Compute, :41-112). The sums run over
// 8 SIMD lanes (element i in lane i % 8) that are added in Highway's AVX2 SumOfLanes order.
inline float CflSumOfLanes(const float l[8]) { return ((l[0] + l[4]) + (l[2] + l[6])) + ((l[1] + l[5]) + (l[3] + l[7])); }
inline int32_t FindBestMultiplier(const float* values_m, const float* values_s, size_t num, float base, float distance_mul) {
  if (num == 0) return 0;
  const float kInvColorFactor = 1.0f / 84, kCoeff = 1.0f / 3, kThres = 100.0f, eps = 100, kClamp = 20.0f;
  const float coeffx2 = kCoeff * 2.0f;
  float x = 0;
  for (size_t iter = 0; iter < 20; iter++) {
    const float first_derivative = 2 * distance_mul * num * x;
    const float first_derivative_peps = 2 * distance_mul * num * (x + eps);
    const float first_derivative_meps = 2 * distance_mul * num * (x - eps);
    float fd[8] = {0}, fdpe[8] = {0}, fdme[8] = {0};
    const float xpe = x + eps, xme = x - eps;
    for (size_t i = 0; i < num; i++) {
      const float a = kInvColorFactor * values_m[i];
      const float b = base * values_m[i] - values_s[i];
      const float v = std::fmaf(a, x, b), vpe = std::fmaf(a, xpe, b), vme = std::fmaf(a, xme, b);
      const float av = std::fabs(v), avpe = std::fabs(vpe), avme = std::fabs(vme);
      const float acoeffx2 = coeffx2 * a;
      float d = acoeffx2 * (av + 1.0f), dpe = acoeffx2 * (avpe + 1.0f), dme = acoeffx2 * (avme + 1.0f);
      d = v < 0.0f ? 0.0f - d : d;
      dpe = vpe < 0.0f ? 0.0f - dpe : dpe;
      dme = vme < 0.0f ? 0.0f - dme : dme;
      const bool above = av >= kThres;  // (one mask for all three, as in the reference)
      fd[i % 8] = fd[i % 8] + (above ? 0.0f : d);
      fdpe[i % 8] = fdpe[i % 8] + (above ? 0.0f : dpe);
      fdme[i % 8] = fdme[i % 8] + (above ? 0.0f : dme);
    }
    const float dfpeps = first_derivative_peps + CflSumOfLanes(fdpe);
    const float dfmeps = first_derivative_meps + CflSumOfLanes(fdme);
    const float df = first_derivative + CflSumOfLanes(fd);
    const float ddf = (dfpeps - dfmeps) / (2 * eps);
    const float kExperimentalInsignificantStabilizer = 0.85f;
    const float step = df / (ddf + kExperimentalInsignificantStabilizer);
    x -= std::min(kClamp, std::max(-kClamp, step));
    if (std::fabs(step) < 3e-3f) break;
  }
  const float towards_zero = 2.6f;
  if (x >= towards_zero) {
    x -= towards_zero;
  } else if (x <= -towards_zero) {
    x += towards_zero;
  } else {
    x = 0;
  }
  return static_cast<int32_t>(std::max(-128.0f, std::min(127.0f, std::roundf(x))));
}

inline int kOrderFirstStrategyOf(uint32_t ord) {
  for (int o = 0; o < kNumStrategies; o++)
    if (kStrategyOrder[o] == static_cast<int>(ord)) return o;
  return -1;
}

// GaborishInverse, lib/jxl/enc_gaborish.cc:21-70 (weights) + Symmetric5, lib/jxl/convolve_symmetric5.cc:28-118
// (summation order; mirrored borders at the size of the padded image), with mul = {1, 1, 1} as
// lib/jxl/enc_heuristics.cc:1121-1131 calls it.
struct GaborishInverseWeights {
  float c, r, R, d, D, L;  // lower-right quadrant:  c r R / r d L / R L D
};
inline GaborishInverseWeights MakeGaborishInverseWeights(float mul) {
  static const float kGaborish[5] = {-0.09495815671340026, -0.041031725066768575, 0.013710004822696948,
                                     0.006510206083837737, -0.0014789063378272242};
  double sum = 1.0 + mul * 4 * (kGaborish[0] + kGaborish[1] + kGaborish[2] + kGaborish[4] + 2 * kGaborish[3]);
  if (sum < 1e-5) sum = 1e-5;
  const float normalize = static_cast<float>(1.0 / sum);
  const float normalize_mul = mul * normalize;
  return GaborishInverseWeights{normalize, normalize_mul * kGaborish[0], normalize_mul * kGaborish[2],
                                normalize_mul * kGaborish[1], normalize_mul * kGaborish[4], normalize_mul * kGaborish[3]};
}
inline int64_t EncMirror(int64_t x, int64_t size) {
  while (x < 0 || x >= size) x = x < 0 ? -x - 1 : 2 * size - 1 - x;
  return x;
}
inline void GaborishInverse(Plane xyb[3]) {
  const GaborishInverseWeights k = MakeGaborishInverseWeights(1.0f);
  for (int c = 0; c < 3; c++) {
    const Plane in = xyb[c];
    const int64_t w = in.w, h = in.h;
    auto row_sum = [&](int64_t x, int64_t y, float wx0, float wx1, float wx2) {
      const float* row = in.Row(EncMirror(y, h));
      const float sum_2 = wx2 * (row[EncMirror(x - 2, w)] + row[EncMirror(x + 2, w)]);
      const float sum_1 = wx1 * (row[EncMirror(x - 1, w)] + row[EncMirror(x + 1, w)]);
      const float sum_0 = wx0 * row[x];
      return sum_2 + (sum_1 + sum_0);
    };
    for (int64_t y = 0; y < h; y++)
      for (int64_t x = 0; x < w; x++) {
        float sum0 = row_sum(x, y, k.c, k.r, k.R);
        sum0 += row_sum(x, y - 2, k.R, k.L, k.D);
        float sum1 = row_sum(x, y + 2, k.R, k.L, k.D);
        sum0 += row_sum(x, y - 1, k.r, k.d, k.L);
        sum1 += row_sum(x, y + 1, k.r, k.d, k.L);
        xyb[c].Row(y)[x] = sum0 + sum1;
      }
  }
}

// ---------------------------------------------------------------- adaptive quantisation (E3)
// AdaptiveQuantizationMap / InitialQuantField, lib/jxl/enc_adaptive_quantization.cc:78-716, :1255-1276, as libjxl's
// effort 7 runs it on the XYB planes *before* inverse Gaborish (lib/jxl/enc_heuristics.cc:1104-1117). Restated per
// 64x64 tile like the reference (ComputeTile, :468-630), because which pixels of a tile row take the SIMD form of the
// Laplacian (`0.25 * ((r + l) + (t + b))`) and which the scalar form (`0.25 * (((t + b) + l) + r)`) follows from the
// tile's own x range (8 lanes: the x86 AVX2 target, as for the chroma-from-luma sums). Plain C++ expressions of the
// reference are evaluated without contraction, MulAdd is fmaf. The two masking images that only the entropy-estimate
// AcStrategy search reads (mask, mask1x1) are not produced.
namespace aq {
constexpr float kSGmul = 226.77216153508914f;
constexpr float kSGmul2 = 1.0f / 73.377132366608819f;
constexpr float kLog2 = 0.693147181f;
constexpr float kSGRetMul = kSGmul2 * 18.6580932135f * kLog2;
constexpr float kSGVOffset = 7.7825991679894591f;

template <bool invert>
inline float RatioOfDerivatives(float v) {  // :117-136
  const float kEpsilon = 1e-2;
  v = v < 0.0f ? 0.0f : v;
  const float kNumMul = kSGRetMul * 3 * kSGmul;
  const float kVOffset = kSGVOffset * kLog2 + kEpsilon;
  const float kDenMul = kLog2 * kSGmul;
  const float v2 = v * v;
  const float num = std::fmaf(kNumMul, v2, kEpsilon);
  const float den = std::fmaf(kDenMul * v, v2, kVOffset);
  return invert ? num / den : den / num;
}

inline float MaskingSqrt(float v) {  // :341-348
  const float kLogOffset = 27.505837037000106f;
  const float kMul = 211.66567973503678f;
  const float mul_v = static_cast<float>(kMul * 1e8);
  return 0.25f * std::sqrt(std::fmaf(v, std::sqrt(mul_v), kLogOffset));
}

inline float ComputeMask(float out_val) {  // :84-108
  const float kBase = -0.7647f, kMul4 = 9.4708735624378946f, kMul2 = 17.35036561631863f;
  const float kOffset2 = 302.59587815579727f, kMul3 = 6.7943250517376494f, kOffset3 = 3.7179635626140772f;
  const float kOffset4 = 0.25f * kOffset3, kMul0 = 0.80061762862741759f;
  const float v1 = std::max(out_val * kMul0, 1e-3f);
  const float v2 = 1.0f / (v1 + kOffset2);
  const float v3 = 1.0f / std::fmaf(v1, v1, kOffset3);
  const float v4 = 1.0f / std::fmaf(v1, v1, kOffset4);
  return kBase + std::fmaf(kMul4, v4, std::fmaf(kMul2, v2, kMul3 * v3));
}

inline float SumOfLanes8(const float l[8]) { return ((l[0] + l[4]) + (l[2] + l[6])) + ((l[1] + l[5]) + (l[3] + l[7])); }

inline float HfModulation(const Plane& py, size_t x, size_t y, float out_val) {  // :250-304
  const float valmin_y = 0.0206f;
  float sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t dy = 0; dy < 8; dy++) {
    const float* row = py.Row(y + dy) + x;
    const float* next = dy == 7 ? row : py.Row(y + dy + 1) + x;
    for (size_t i = 0; i < 8; i++) {
      const float right = i == 7 ? 0.0f : std::min(valmin_y, std::fabs(row[i] - row[i + 1]));
      sum[i] = sum[i] + right;
      sum[i] = sum[i] + std::min(valmin_y, std::fabs(row[i] - next[i]));
    }
  }
  const float kMul_y = -0.38f;
  float s = SumOfLanes8(sum);
  s *= kMul_y;
  const float kOffset = 0.42f;
  s += kOffset;
  return s + out_val;
}

inline float GammaModulation(const Plane& px, const Plane& py, size_t x, size_t y, float out_val) {  // :169-200
  const float kBias = 0.16f;
  float ratio[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t dy = 0; dy < 8; dy++) {
    const float* rx = px.Row(y + dy) + x;
    const float* ry = py.Row(y + dy) + x;
    for (size_t i = 0; i < 8; i++) {
      const float iny = ry[i] + kBias, inx = rx[i];
      ratio[i] = ratio[i] + RatioOfDerivatives<true>(iny - inx);
      ratio[i] = ratio[i] + RatioOfDerivatives<true>(iny + inx);
    }
  }
  const float overall = SumOfLanes8(ratio) * (0.5f / 64);
  const float kGamma = 0.1005613337192697f;
  return std::fmaf(kGamma, FastLog2f(overall), out_val);
}

inline float BlueModulation(const Plane& px, const Plane& py, const Plane& pb, size_t x, size_t y, float out_val) {  // :211-247
  const float kLimit = 0.027121074570634722f, kOffset = 0.084381641171960495f;
  float sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t dy = 0; dy < 8; dy++) {
    const float* rx = px.Row(y + dy) + x;
    const float* ry = py.Row(y + dy) + x;
    const float* rb = pb.Row(y + dy) + x;
    for (size_t i = 0; i < 8; i++) {
      const float p_y_effective = (ry[i] + kOffset) + std::fabs(rx[i]);
      sum[i] = sum[i] + (rb[i] > p_y_effective ? std::min(rb[i] - p_y_effective, kLimit) : 0.0f);
    }
  }
  const float kMul = 0.14207000358439159f;
  float s = SumOfLanes8(sum);
  if (s >= 32 * kLimit) s = 64 * kLimit - s;
  const float kMaxLimit = 15.398788439047934f;
  if (s >= kMaxLimit * kLimit) s = kMaxLimit * kLimit;
  s *= kMul;
  return s + out_val;
}

inline void StoreMin4(float v, float& min0, float& min1, float& min2, float& min3) {  // :356-375
  if (v < min3) {
    if (v < min0) {
      min3 = min2, min2 = min1, min1 = min0, min0 = v;
    } else if (v < min1) {
      min3 = min2, min2 = min1, min1 = v;
    } else if (v < min2) {
      min3 = min2, min2 = v;
    } else {
      min3 = v;
    }
  }
}

// The initial quantisation field of one frame: xsize_blocks x ysize_blocks floats, row major.
inline std::vector<float> InitialQuantField(const Plane xyb[3], float butteraugli_target, float rescale) {
  const size_t xsize = xyb[1].w, ysize = xyb[1].h;  // whole blocks
  const size_t W = xsize / 8, H = ysize / 8;
  const float kAcQuant = 0.725f;
  const float scale = kAcQuant / butteraugli_target * rescale;  // :1273-1275
  std::vector<float> aq_map(W * H, 0.0f);
  const float match_gamma_offset = 0.019f, limit = 0.2f;
  // FuzzyErosion weights (:380-410)
  float fmul = 0.0f;
  if (butteraugli_target < 2.0f) fmul = (2.0f - butteraugli_target) * (1.0f / 2.0f);
  float kMul0 = 0.125f + fmul * 0.0f, kMul1 = 0.10f + fmul * -0.10f, kMul2 = 0.09f + fmul * -0.09f, kMul3 = 0.06f + fmul * -0.06f;
  const float kTotal = 0.29959705784054957f;
  const float norm = kTotal / (kMul0 + kMul1 + kMul2 + kMul3);
  kMul0 *= norm, kMul1 *= norm, kMul2 *= norm, kMul3 *= norm;
  // PerBlockModulations (:306-339)
  const float base_level = 0.48f * scale;
  float dampen = 1.0f;
  if (butteraugli_target >= 2.0f) {
    dampen = 1.0f - ((butteraugli_target - 2.0f) / (14.0f - 2.0f));
    if (dampen < 0) dampen = 0;
  }
  const float mul = scale * dampen, add = (1.0f - dampen) * base_level;

  std::vector<float> diff_row(64 + 8), pre(18 * 18);
  for (size_t ty = 0; ty < DivCeil(H, size_t{8}); ty++)
    for (size_t tx = 0; tx < DivCeil(W, size_t{8}); tx++) {
      const size_t bx0 = tx * 8, by0 = ty * 8, bx1 = std::min(W, bx0 + 8), by1 = std::min(H, by0 + 8);
      size_t x_start = bx0 * 8, x_end = bx1 * 8, y_start = by0 * 8, y_end = by1 * 8;
      if (x_start != 0) x_start -= 4;
      if (x_end != xsize) x_end += 4;
      if (y_start != 0) y_start -= 4;
      if (y_end != ysize) y_end += 4;
      const size_t pw = (x_end - x_start) / 4, ph = (y_end - y_start) / 4;
      const Plane& Y = xyb[1];
      for (size_t y = y_start; y < y_end; y++) {
        const size_t y2 = y + 1 < ysize ? y + 1 : y, y1 = y > 0 ? y - 1 : y;
        const float *row_in = Y.Row(y), *row_in1 = Y.Row(y1), *row_in2 = Y.Row(y2);
        auto finish = [&](size_t x, float base) {
          const float gammac = RatioOfDerivatives<false>(row_in[x] + match_gamma_offset);
          float diff = gammac * (row_in[x] - base);
          diff *= diff;
          if (diff >= limit) diff = limit;
          diff = MaskingSqrt(diff);
          if ((y % 4) != 0) {
            diff_row[x - x_start] += diff;
          } else {
            diff_row[x - x_start] = diff;
          }
        };
        auto scalar_pixel = [&](size_t x) {
          const size_t x2 = x + 1 < xsize ? x + 1 : x, x1 = x > 0 ? x - 1 : x;
          finish(x, 0.25f * (row_in2[x] + row_in1[x] + row_in[x1] + row_in[x2]));
        };
        size_t x = x_start;
        if (x_start == 0) {
          scalar_pixel(x_start);
          ++x;
        }
        for (; x + 1 + 8 < x_end; x += 8)
          for (size_t i = x; i < x + 8; i++)
            finish(i, 0.25f * ((row_in[i + 1] + row_in[i - 1]) + (row_in2[i] + row_in1[i])));
        for (; x < x_end; ++x) scalar_pixel(x);
        if (y % 4 == 3) {
          float* out = pre.data() + ((y - y_start) / 4) * pw;
          for (size_t i = 0; i < pw; i++)
            out[i] = (diff_row[i * 4] + diff_row[i * 4 + 1] + diff_row[i * 4 + 2] + diff_row[i * 4 + 3]) * 0.25f;
        }
      }
      // FuzzyErosion (:380-451): the 4 smallest of each 3x3 neighbourhood, 2x2 cells summed per block
      const size_t fx0 = x_start % 8 == 0 ? 0 : 1, fy0 = y_start % 8 == 0 ? 0 : 1;
      const size_t fw = (bx1 - bx0) * 2, fh = (by1 - by0) * 2;
      for (size_t fy = 0; fy < fh; fy++) {
        const size_t y = fy + fy0;
        const size_t ym1 = y >= 1 ? y - 1 : y, yp1 = y + 1 < ph ? y + 1 : y;
        const float *rowt = pre.data() + ym1 * pw, *row = pre.data() + y * pw, *rowb = pre.data() + yp1 * pw;
        float* row_out = aq_map.data() + (by0 + fy / 2) * W + bx0;
        for (size_t fx = 0; fx < fw; fx++) {
          const size_t x = fx + fx0;
          const size_t xm1 = x >= 1 ? x - 1 : x, xp1 = x + 1 < pw ? x + 1 : x;
          float min0 = row[x], min1 = row[xm1], min2 = row[xp1], min3 = rowt[xm1];
          if (min0 > min1) std::swap(min0, min1);
          if (min0 > min2) std::swap(min0, min2);
          if (min0 > min3) std::swap(min0, min3);
          if (min1 > min2) std::swap(min1, min2);
          if (min1 > min3) std::swap(min1, min3);
          if (min2 > min3) std::swap(min2, min3);
          StoreMin4(rowt[x], min0, min1, min2, min3);
          StoreMin4(rowt[xp1], min0, min1, min2, min3);
          StoreMin4(rowb[xm1], min0, min1, min2, min3);
          StoreMin4(rowb[x], min0, min1, min2, min3);
          StoreMin4(rowb[xp1], min0, min1, min2, min3);
          const float v = kMul0 * min0 + kMul1 * min1 + kMul2 * min2 + kMul3 * min3;
          if (fx % 2 == 0 && fy % 2 == 0) {
            row_out[fx / 2] = v;
          } else {
            row_out[fx / 2] += v;
          }
        }
      }
      for (size_t by = by0; by < by1; by++)
        for (size_t bx = bx0; bx < bx1; bx++) {
          float out_val = ComputeMask(aq_map[by * W + bx]);
          out_val = HfModulation(xyb[1], bx * 8, by * 8, out_val);
          out_val = GammaModulation(xyb[0], xyb[1], bx * 8, by * 8, out_val);
          out_val = BlueModulation(xyb[0], xyb[1], xyb[2], bx * 8, by * 8, out_val);
          aq_map[by * W + bx] = FastPow2f(out_val * 1.442695041f) * mul + add;
        }
    }
  return aq_map;
}

inline float InitialQuantDC(float butteraugli_target) {  // :1255-1267
  const float kDcMul = 0.3f, kDcQuantPow = 0.83f, kDcQuant = 1.095924047623553f;
  const float butteraugli_target_dc = std::max<float>(
      0.5f * butteraugli_target, std::min<float>(butteraugli_target, kDcMul * std::pow((1.0f / kDcMul) * butteraugli_target, kDcQuantPow)));
  return std::min(kDcQuant / butteraugli_target_dc, 50.f);
}

// Quantizer::ComputeGlobalScaleAndQuant, lib/jxl/quantizer.cc:39-69
inline void ComputeGlobalScaleAndQuant(float quant_dc, float quant_median, float quant_median_absd, int* global_scale,
                                       int* quant_dc_out) {
  float scale = 65536 * (quant_median - quant_median_absd) / 5.0f;
  if (scale < 1) scale = 1;
  if (scale > (1 << 15)) scale = 1 << 15;
  int new_global_scale = static_cast<int>(scale);
  const int scaled_quant_dc = static_cast<int>(quant_dc * 4096 * 1.6);
  if (new_global_scale > scaled_quant_dc) {
    new_global_scale = scaled_quant_dc;
    if (new_global_scale <= 0) new_global_scale = 1;
  }
  *global_scale = new_global_scale;
  const float inv_global_scale = static_cast<float>(1.0 * 65536 / new_global_scale);
  float fval = quant_dc * inv_global_scale + 0.5f;
  fval = std::min<float>(1 << 16, fval);
  *quant_dc_out = static_cast<int>(fval);
}
}  // namespace aq

// ---------------------------------------------------------------- quantisation of a varblock as libjxl's effort 7 does it (E8)
// AdjustQuantBlockAC / QuantizeBlockAC / QuantizeRoundtripYBlockAC, lib/jxl/enc_group.cc:46-368, driven as in
// ComputeCoefficients (:455-493). `xsize` >= `ysize` are the covered blocks in coefficient-layout order. The inverse
// dequantisation matrix is taken as 1 / matrix (as for the chroma-from-luma fit). Plain C++ of the reference
// (float / double mixes included) is evaluated as written, without contraction.
namespace e8 {
inline void AdjustQuantBlockAC(float scale, size_t c, float qm_multiplier, int quant_kind, size_t xsize, size_t ysize,
                               float* thresholds, const float* block_in, const float* dm, int32_t* quant) {
  const uint32_t kPartialBlockKinds = (1u << 1) | (1u << 2) | (1u << 3) | (1u << 12) | (1u << 13) | (1u << 14) | (1u << 15) |
                                      (1u << 16) | (1u << 17);  // IDENTITY, DCT2X2, DCT4X4, DCT4X8, DCT8X4, AFV0..3
  if ((1u << quant_kind) & kPartialBlockKinds) return;
  const float qac = scale * (*quant);
  if (xsize > 1 || ysize > 1) {
    for (int i = 0; i < 4; ++i) {
      thresholds[i] -= std::min(std::max(0.003f * xsize * ysize, 0.f), 0.08f);
      if (thresholds[i] < 0.54) thresholds[i] = 0.54;
    }
  }
  float sum_of_highest_freq_row_and_column = 0, sum_of_error = 0, sum_of_vals = 0;
  float hfNonZeros[4] = {}, hfMaxError[4] = {};
  for (size_t y = 0; y < ysize * 8; y++) {
    for (size_t x = 0; x < xsize * 8; x++) {
      const size_t pos = y * 8 * xsize + x;
      if (x < xsize && y < ysize) continue;
      const size_t hfix = (static_cast<size_t>(y >= ysize * 8 / 2) * 2 + static_cast<size_t>(x >= xsize * 8 / 2));
      const float val = block_in[pos] * ((1.0f / dm[pos]) * qac * qm_multiplier);
      const float v = (std::abs(val) < thresholds[hfix]) ? 0 : rintf(val);
      const float error = std::abs(val - v);
      sum_of_error += error;
      sum_of_vals += std::abs(v);
      if (c == 1 && v == 0) {
        if (hfMaxError[hfix] < error) hfMaxError[hfix] = error;
      }
      if (v != 0.0f) {
        hfNonZeros[hfix] += std::abs(v);
        const bool in_corner = y >= 7 * ysize && x >= 7 * xsize;
        const bool on_border = y == ysize * 8 - 1 || x == xsize * 8 - 1;
        const bool in_larger_corner = x >= 4 * xsize && y >= 4 * ysize;
        if (in_corner || (on_border && in_larger_corner)) sum_of_highest_freq_row_and_column += std::abs(val);
      }
    }
  }
  if (c == 1 && sum_of_vals * 8 < xsize * ysize) {
    const double kLimit = 0.46, kMul = 0.9999;
    const int32_t orig_quant = *quant;
    int32_t new_quant = *quant;
    for (int i = 1; i < 4; ++i) {
      if (hfNonZeros[i] == 0.0 && hfMaxError[i] > kLimit) {
        new_quant = orig_quant + 1;
        break;
      }
    }
    *quant = new_quant;
    if (hfNonZeros[3] == 0.0 && hfMaxError[3] > kLimit) {
      thresholds[3] = kMul * hfMaxError[3] * new_quant / orig_quant;
    } else if ((hfNonZeros[1] == 0.0 && hfMaxError[1] > kLimit) || (hfNonZeros[2] == 0.0 && hfMaxError[2] > kLimit)) {
      thresholds[1] = kMul * std::max(hfMaxError[1], hfMaxError[2]) * new_quant / orig_quant;
      thresholds[2] = thresholds[1];
    } else if (hfNonZeros[0] == 0.0 && hfMaxError[0] > kLimit) {
      thresholds[0] = kMul * hfMaxError[0] * new_quant / orig_quant;
    }
  }
  {
    const float all = hfNonZeros[0] + hfNonZeros[1] + hfNonZeros[2] + hfNonZeros[3] + 1;
    const float mul[3] = {70, 30, 60};
    if (mul[c] * sum_of_highest_freq_row_and_column >= all) {
      *quant += mul[c] * sum_of_highest_freq_row_and_column / all;
      if (*quant >= 256) *quant = 256 - 1;
    }
  }
  if (quant_kind == 0) {
    if (hfNonZeros[0] + hfNonZeros[1] + hfNonZeros[2] + hfNonZeros[3] < 11) {
      *quant += 1;
      if (*quant >= 256) *quant = 256 - 1;
    }
  }
  {
    static const double kMul1[4][3] = {{0.22080615753848404, 0.45797479824262011, 0.29859235095977965},
                                       {0.70109486510286834, 0.16185281305512639, 0.14387691730035473},
                                       {0.114985964456218638, 0.44656840441027695, 0.10587658215149048},
                                       {0.46849665264409396, 0.41239077937781954, 0.088667407767185444}};
    static const double kMul2[4][3] = {{0.27450281941822197, 1.1255766549984996, 0.98950459134128388},
                                       {0.4652168675598285, 0.40945807983455818, 0.36581899811751367},
                                       {0.28034972424715715, 0.9182653201929738, 1.5581531543057416},
                                       {0.26873118114033728, 0.68863712390392484, 1.2082185408666786}};
    const double kQuantNormalizer = 2.2942708343284721;
    sum_of_error *= kQuantNormalizer;
    sum_of_vals *= kQuantNormalizer;
    if (quant_kind >= 4) {  // >= DCT16X16
      int ix = 3;
      if (quant_kind == 10 || quant_kind == 11) {
        ix = 1;
      } else if (quant_kind == 4) {
        ix = 0;
      } else if (quant_kind == 5) {
        ix = 2;
      }
      const double limit = kMul1[ix][c] * xsize * ysize * 8 * 8 + kMul2[ix][c] * sum_of_vals;
      int step = sum_of_error / limit;
      if (step >= 2) step = 2;
      if (step < 0) step = 0;
      if (sum_of_error > limit) {
        *quant += step;
        if (*quant >= 256) *quant = 256 - 1;
      }
    }
  }
  {
    const int32_t div = (xsize * ysize);
    int32_t activity = (static_cast<int32_t>(hfNonZeros[0]) + div / 2) / div;
    const int32_t orig_qp_limit = std::max(4, *quant / 2);
    for (int i = 1; i < 4; ++i) activity = std::min(activity, (static_cast<int32_t>(hfNonZeros[i]) + div / 2) / div);
    if (activity >= 15) activity = 15;
    int32_t qp = *quant - activity;
    if (c == 1) {
      for (int i = 1; i < 4; ++i) thresholds[i] += 0.01 * activity;
    }
    if (qp < orig_qp_limit) qp = orig_qp_limit;
    *quant = qp;
  }
}

inline void QuantizeBlockAC(float scale, size_t c, float qm_multiplier, size_t xsize, size_t ysize, float* thresholds,
                            const float* block_in, const float* dm, int32_t quant, int32_t* block_out) {
  const float qac = scale * quant;
  if (c != 1 && xsize * ysize >= 4) {
    for (int i = 0; i < 4; ++i) {
      thresholds[i] -= 0.00744f * xsize * ysize;
      if (thresholds[i] < 0.5) thresholds[i] = 0.5;
    }
  }
  const float quantv = qac * qm_multiplier;
  for (size_t y = 0; y < ysize * 8; y++) {
    const size_t yfix = static_cast<size_t>(y >= ysize * 8 / 2) * 2;
    for (size_t x = 0; x < xsize * 8; x++) {
      const size_t off = y * 8 * xsize + x;
      const float threshold = thresholds[yfix + static_cast<size_t>(x >= xsize * 8 / 2)];
      const float q = (1.0f / dm[off]) * quantv;
      const float val = q * block_in[off];
      block_out[off] = std::abs(val) >= threshold ? static_cast<int32_t>(rintf(val)) : 0;
    }
  }
}
}  // namespace e8

inline std::vector<uint8_t> EncodeVarDCT(const uint8_t* rgb, uint32_t xsize, uint32_t ysize, const EncodeParams& p,
                                         EncoderStats* stats = nullptr) {
  JXLO_CHECK(xsize > 0 && ysize > 0, "empty image");
  // ---- frame geometry
  FrameHeader fh;
  fh.is_modular = false;
  fh.xsize = xsize;
  fh.ysize = ysize;
  fh.x_qm_scale = p.x_qm_scale;
  fh.b_qm_scale = p.b_qm_scale;
  fh.flags = p.dc_smoothing ? uint64_t{0} : uint64_t{kFlagSkipAdaptiveDCSmoothing};
  fh.lf = DefaultLoopFilter();
  fh.lf.gab = p.gab;
  fh.lf.epf_iters = p.epf_iters;
  fh.passes.num_passes = p.num_passes;
  for (uint32_t i = 0; i + 1 < p.num_passes; i++) fh.passes.shift[i] = p.num_passes - 1 - i;
  const FrameDimensions dim = ToFrameDimensions(fh);
  const size_t W = dim.xsize_blocks, H = dim.ysize_blocks;
  const size_t PW = W * 8, PH = H * 8;
  std::mt19937 rng(p.seed);

  // ---- RGB8 -> XYB planes, edge-replicated to whole blocks
  Plane xyb[3] = {Plane(PW, PH), Plane(PW, PH), Plane(PW, PH)};
  {
    float lut[256];
    for (int i = 0; i < 256; i++) lut[i] = SrgbToLinear(i / 255.0f);
    for (size_t y = 0; y < PH; y++) {
      const size_t sy = std::min<size_t>(y, ysize - 1);
      for (size_t x = 0; x < PW; x++) {
        const size_t sx = std::min<size_t>(x, xsize - 1);
        const uint8_t* px = rgb + (sy * xsize + sx) * 3;
        LinearRgbToXyb(lut[px[0]], lut[px[1]], lut[px[2]], &xyb[0].Row(y)[x], &xyb[1].Row(y)[x], &xyb[2].Row(y)[x]);
      }
    }
  }

  // ---- initial quantisation field (E3), from the planes before inverse Gaborish (lib/jxl/enc_heuristics.cc:1104-1117)
  const bool adaptive = p.adaptive_quant && !p.random_side_info;
  std::vector<float> quant_field;
  if (adaptive) quant_field = aq::InitialQuantField(xyb, p.gab ? p.distance : p.distance * 0.62f, 1.0f);
  const bool acs_search = p.strategy_mode == 3;
  JXLO_CHECK(!acs_search || adaptive, "the AcStrategy search needs the adaptive quantisation field");
  std::vector<float> mask1x1;  // the per-pixel masking image of the search, also from the planes before inverse Gaborish
  if (acs_search) mask1x1 = acs::Mask1x1(xyb[1].d.data(), PW, PH, [](float v) { return aq::RatioOfDerivatives<false>(v); });
  if (adaptive && std::getenv("JXLO_DEBUG_AQ")) {
    std::vector<float> v = quant_field;
    std::sort(v.begin(), v.end());
    std::fprintf(stderr, "O aq: min %g p10 %g median %g p90 %g max %g first %a %a %a\n", v.front(), v[v.size() / 10], v[v.size() / 2],
                 v[v.size() * 9 / 10], v.back(), quant_field[0], quant_field[1], quant_field[W]);
  }

  // ---- inverse Gaborish: the 5x5 sharpening that the decoder's Gaborish smoothing undoes
  if (p.gab && p.gab_inverse) GaborishInverse(xyb);

  // ---- global quantiser: dequant step = table * inv_global_scale / raw_quant
  // quant value ~ 0.79 / distance as in libjxl's InitialQuantField target
  const float quant_ac = 0.79f / std::max(0.1f, p.distance);
  const int base_raw = 16;  // typical raw quant field value
  int global_scale = std::max(1, std::min(65535 + 8192, static_cast<int>(quant_ac * 65536 / base_raw + 0.5f)));
  int quant_dc = std::max(1, std::min(65536, static_cast<int>(0.9f / std::max(0.1f, p.distance) * 65536 / global_scale + 0.5f)));
  if (adaptive) {  // lib/jxl/enc_heuristics.cc:1055, :1115-1116
    const float q = 0.39 / p.distance;
    aq::ComputeGlobalScaleAndQuant(aq::InitialQuantDC(p.distance), q, 0, &global_scale, &quant_dc);
  }
  const float inv_global_scale = 1.0 * 65536 / global_scale;
  const float dc_quant[3] = {1.0f / 4096, 1.0f / 512, 1.0f / 256};
  float mul_dc[3];
  for (int c = 0; c < 3; c++) mul_dc[c] = (inv_global_scale / quant_dc) * dc_quant[c];
  const float x_dm = std::pow(1 / (1.25f), p.x_qm_scale - 2.0f), b_dm = std::pow(1 / (1.25f), p.b_qm_scale - 2.0f);

  std::vector<float> cfl_tables[17];
  auto table_for_cfl = [&](int strategy) -> const std::vector<float>& {
    const int t = kStrategyToQuantTable[strategy];
    if (cfl_tables[t].empty()) cfl_tables[t] = ComputeQuantTable(LibraryEncoding(t), t);
    return cfl_tables[t];
  };
  // ---- side information per block
  std::vector<uint8_t> acs(W * H, 0xFF), sharp(W * H, 4);
  std::vector<int32_t> raw_quant(W * H, 0);
  const size_t cmw = DivCeil(W, size_t{8}), cmh = DivCeil(H, size_t{8});
  std::vector<int8_t> ytox(cmw * cmh, 0), ytob(cmw * cmh, 0);
  if (p.random_side_info) {
    for (auto& v : ytox) v = static_cast<int8_t>(static_cast<int>(rng() % 21) - 10);
    for (auto& v : ytob) v = static_cast<int8_t>(static_cast<int>(rng() % 21) - 10);
    for (auto& v : sharp) v = rng() % 8;
  }
  auto block_variance = [&](size_t bx, size_t by, size_t nx, size_t ny) {
    double s = 0, s2 = 0;
    const size_t n = nx * ny * 64;
    for (size_t y = by * 8; y < (by + ny) * 8; y++)
      for (size_t x = bx * 8; x < (bx + nx) * 8; x++) {
        const double v = xyb[1].Row(y)[x];
        s += v;
        s2 += v * v;
      }
    return s2 / n - (s / n) * (s / n);
  };
  auto fits = [&](size_t bx, size_t by, int s) {
    const size_t cx = kCoveredX[s], cy = kCoveredY[s];
    if (bx + cx > W || by + cy > H) return false;
    if (bx / 32 != (bx + cx - 1) / 32 || by / 32 != (by + cy - 1) / 32) return false;  // stay inside one group
    for (size_t y = by; y < by + cy; y++)
      for (size_t x = bx; x < bx + cx; x++)
        if (acs[y * W + x] != 0xFF) return false;
    return true;
  };
  if (acs_search) {
    // per 64x64 tile: chroma-from-luma factors under 8x8 DCTs without quantisation weighting (CfLHeuristics::ComputeTile
    // with ac_strategy == nullptr, lib/jxl/enc_chroma_from_luma.cc:196-342), then the search with those factors
    acs::Config config;
    acs::InitConfig(&config, p.distance);
    std::vector<float> mats[kNumStrategies], inv_mats[kNumStrategies];
    for (int st = 0; st <= kDCT32X64; st++) {  // all transforms up to 64x64 (AcStrategyHeuristics::Init)
      const int t = kStrategyToQuantTable[st];
      mats[st] = ComputeQuantTable(LibraryEncoding(t), t);
      inv_mats[st] = ComputeQuantWeights(LibraryEncoding(t), t);
      size_t lx = kCoveredX[st], ly = kCoveredY[st];
      if (ly > lx) std::swap(lx, ly);
      const size_t size = 64u * lx * ly;
      for (int c = 0; c < 3; c++)
        for (size_t y = 0; y < ly; y++)
          for (size_t x = 0; x < lx; x++) inv_mats[st][c * size + y * 8 * lx + x] = 0;  // (LLF corner, quant_weights.cc:337-349)
      config.matrix[st] = &mats[st];
      config.inv_matrix[st] = &inv_mats[st];
    }
    config.quant_field = quant_field.data();
    config.quant_stride = W;
    config.mask1x1 = mask1x1.data();
    config.mask_stride = PW;
    config.mask1x1_xsize = PW;
    for (int c = 0; c < 3; c++) config.src[c] = xyb[c].d.data();
    config.src_stride = PW;
    acs::AcsImage image{acs.data(), W, H};
    std::vector<float> byx(4096), bxx(4096), byb(4096), bbb(4096), blk(3 * 4096), scr(5 * 4096 + 1024);
    for (size_t ty = 0; ty < cmh; ty++)
      for (size_t tx = 0; tx < cmw; tx++) {
        const size_t x0 = tx * 8, y0 = ty * 8, x1 = std::min(W, x0 + 8), y1 = std::min(H, y0 + 8);
        size_t num_ac = 0;
        for (size_t by = y0; by < y1; by++)
          for (size_t bx = x0; bx < x1; bx++) {
            for (int c = 0; c < 3; c++) TransformFromPixelsRef(kDCT, xyb[c].Row(by * 8) + bx * 8, PW, blk.data() + c * 64, scr.data());
            for (int c = 0; c < 3; c++) blk[c * 64] = 0;
            for (size_t i = 0; i < 64; i++) {
              const float qqm_x = 1.0f * inv_mats[kDCT][i], qqm_b = 1.0f * inv_mats[kDCT][128 + i];
              byx[num_ac] = blk[64 + i] * qqm_x;
              bxx[num_ac] = blk[i] * qqm_x;
              byb[num_ac] = blk[64 + i] * qqm_b;
              bbb[num_ac] = blk[128 + i] * qqm_b;
              num_ac++;
            }
          }
        const int32_t pre_x = FindBestMultiplier(byx.data(), bxx.data(), num_ac, 0.0f, 1e-9f);
        const int32_t pre_b = FindBestMultiplier(byb.data(), bbb.data(), num_ac, 1.0f, 1e-9f);
        const float cmap_factors[3] = {0.0f + static_cast<int8_t>(pre_x) * (1.0f / 84), 0.0f, 1.0f + static_cast<int8_t>(pre_b) * (1.0f / 84)};
        acs::ProcessRectACS(p.distance, config, x0, y0, x1 - x0, y1 - y0, cmap_factors, blk.data(), scr.data(), &image);
      }
  }
  size_t num_varblocks = 0;
  for (size_t by = 0; by < H; by++) {
    for (size_t bx = 0; bx < W; bx++) {
      if (acs_search) {
        if (!(acs[by * W + bx] & 1)) continue;  // (the search has filled the map)
      } else if (acs[by * W + bx] != 0xFF) {
        continue;
      }
      int s = acs_search ? acs[by * W + bx] >> 1 : kDCT;
      if (p.strategy_mode == 1) {
        // all 27 strategies; the larger ones only where they are aligned to their own size
        for (int attempt = 0; attempt < 4; attempt++) {
          const int cand = rng() % 27;
          const size_t cx = kCoveredX[cand], cy = kCoveredY[cand];
          if (cx * cy > 16 && (rng() % 4) != 0) continue;  // keep huge blocks rare
          if (bx % std::min<size_t>(cx, 8) == 0 && by % std::min<size_t>(cy, 8) == 0 && fits(bx, by, cand)) {
            s = cand;
            break;
          }
        }
      } else if (p.strategy_mode >= 100) {  // one forced strategy wherever it fits (decoder coverage per transform)
        const int cand = p.strategy_mode - 100;
        JXLO_CHECK(cand < kNumStrategies, "bad forced strategy");
        if (bx % kCoveredX[cand] == 0 && by % kCoveredY[cand] == 0 && fits(bx, by, cand)) s = cand;
      } else if (p.strategy_mode == 2) {
        static const int kCands[] = {kDCT64X64, kDCT32X32, kDCT16X16, kDCT16X8, kDCT8X16};
        static const double kThresh[] = {2e-6, 1e-5, 6e-5, 1.5e-4, 1.5e-4};
        for (int i = 0; i < 5; i++) {
          const int cand = kCands[i];
          const size_t cx = kCoveredX[cand], cy = kCoveredY[cand];
          if (bx % cx || by % cy || !fits(bx, by, cand)) continue;
          if (block_variance(bx, by, cx, cy) < kThresh[i] * p.distance) {
            s = cand;
            break;
          }
        }
      }
      for (size_t y = 0; y < kCoveredY[s] && !acs_search; y++)
        for (size_t x = 0; x < kCoveredX[s]; x++)
          acs[(by + y) * W + bx + x] = static_cast<uint8_t>((s << 1) | ((x | y) == 0 ? 1 : 0));
      int rq = base_raw;
      if (p.random_side_info) rq = 8 + rng() % 24;
      if (adaptive) {
        // AdjustQuantField (lib/jxl/enc_adaptive_quantization.cc:1203-1253) and Quantizer::SetQuantFieldRect
        // (lib/jxl/quantizer.cc:71-82) for this varblock
        const size_t cx = kCoveredX[s], cy = kCoveredY[s];
        float mean_max_mixer = 1.0f;
        if (p.distance > 1.54138f) mean_max_mixer = std::max(0.0f, mean_max_mixer - (p.distance - 1.54138f) * 0.56391f);
        float max = quant_field[by * W + bx], mean = 0.0f;
        for (size_t iy = 0; iy < cy; iy++)
          for (size_t ix = 0; ix < cx; ix++) {
            mean += quant_field[(by + iy) * W + bx + ix];
            max = std::max(quant_field[(by + iy) * W + bx + ix], max);
          }
        mean /= cy * cx;
        if (cy * cx >= 4) {
          max *= mean_max_mixer;
          max += (1.0f - mean_max_mixer) * mean;
        }
        rq = static_cast<int32_t>(std::max(1.0f, std::min<float>(max * inv_global_scale + 0.5f, 256.0f)));
      }
      raw_quant[by * W + bx] = rq;
      num_varblocks++;
      if (stats) stats->strategy_count[s]++;
    }
  }

  if (std::getenv("JXLO_DEBUG_ACS")) {
    size_t count[kNumStrategies] = {0};
    for (size_t i = 0; i < W * H; i++)
      if (acs[i] & 1) count[acs[i] >> 1]++;
    std::fprintf(stderr, "O acs:");
    for (int st = 0; st < kNumStrategies; st++)
      if (count[st]) std::fprintf(stderr, " %d:%zu", st, count[st]);
    std::fprintf(stderr, "\n");
  }
  // ---- chroma-from-luma map (E6): CfLHeuristics::ComputeTile, lib/jxl/enc_chroma_from_luma.cc:196-342, as libjxl runs it
  // after the block sizes are known (lib/jxl/enc_heuristics.cc:1176-1182): per 64x64 tile the AC coefficients of the
  // varblocks that lie inside the tile, weighted by the quantisation step, and one robust fit per chroma channel.
  // (InvMatrix is taken as 1 / Matrix; the DC-level factors stay at their defaults.)
  if (p.cfl && !p.random_side_info) {
    const float scale = global_scale * (1.0f / 65536);
    std::vector<float> byx(4096), bxx(4096), byb(4096), bbb(4096), blk(3 * 4096), scr(3 * 4096 + 1024);
    for (size_t ty = 0; ty < cmh; ty++)
      for (size_t tx = 0; tx < cmw; tx++) {
        const size_t x0 = tx * 8, y0 = ty * 8, x1 = std::min(W, x0 + 8), y1 = std::min(H, y0 + 8);
        size_t num_ac = 0;
        for (size_t by = y0; by < y1; by++)
          for (size_t bx = x0; bx < x1; bx++) {
            const uint8_t a = acs[by * W + bx];
            if (!(a & 1)) continue;
            const int st = a >> 1;
            size_t cx = kCoveredX[st], cy = kCoveredY[st];
            if (cx + x0 > x1 || cy + y0 > y1) continue;  // (the reference's test: blocks larger than the tile)
            const size_t size = cx * cy * 64;
            for (int c = 0; c < 3; c++)
              TransformFromPixels(st, xyb[c].Row(by * 8) + bx * 8, PW, blk.data() + c * size, scr.data());
            if (cy > cx) std::swap(cx, cy);
            for (size_t iy = 0; iy < cy; iy++)
              for (size_t ix = 0; ix < cx; ix++)
                for (int c = 0; c < 3; c++) blk[c * size + cx * 8 * iy + ix] = 0;
            const std::vector<float>& dm = table_for_cfl(st);
            const float q = scale * 128.0f * raw_quant[by * W + bx];
            for (size_t i = 0; i < size; i++) {
              const float qqm_x = q * (1.0f / dm[i]), qqm_b = q * (1.0f / dm[2 * size + i]);
              byx[num_ac] = blk[size + i] * qqm_x;
              bxx[num_ac] = blk[i] * qqm_x;
              byb[num_ac] = blk[size + i] * qqm_b;
              bbb[num_ac] = blk[2 * size + i] * qqm_b;
              num_ac++;
            }
          }
        ytox[ty * cmw + tx] = static_cast<int8_t>(FindBestMultiplier(byx.data(), bxx.data(), num_ac, 0.0f, 1e-9f));
        ytob[ty * cmw + tx] = static_cast<int8_t>(FindBestMultiplier(byb.data(), bbb.data(), num_ac, 1.0f, 1e-9f));
        if (std::getenv("JXLO_DEBUG_CFL"))
          std::fprintf(stderr, "O tile %zu %zu: n=%zu x=%d b=%d v0=%a %a %a %a\n", tx, ty, num_ac, ytox[ty * cmw + tx], ytob[ty * cmw + tx],
                       byx[70], bxx[70], byb[70], bbb[70]);
      }
  }

  // ---- DC: block means, quantised with chroma-from-luma on DC (default factors: X 0, B 1)
  std::vector<Channel> dcq = {Channel(W, H), Channel(W, H), Channel(W, H)};  // [0] = Y, [1] = X, [2] = B
  Plane dc_rec[3] = {Plane(W, H), Plane(W, H), Plane(W, H)};
  for (size_t by = 0; by < H; by++)
    for (size_t bx = 0; bx < W; bx++) {
      float mean[3];
      for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int y = 0; y < 8; y++)
          for (int x = 0; x < 8; x++) s += xyb[c].Row(by * 8 + y)[bx * 8 + x];
        mean[c] = static_cast<float>(s / 64);
      }
      const int32_t qy = static_cast<int32_t>(std::lrintf(mean[1] / mul_dc[1]));
      const float ry = qy * mul_dc[1];
      const int32_t qx = static_cast<int32_t>(std::lrintf((mean[0] - 0.0f * ry) / mul_dc[0]));
      const int32_t qb = static_cast<int32_t>(std::lrintf((mean[2] - 1.0f * ry) / mul_dc[2]));
      dcq[0].Row(by)[bx] = qy;
      dcq[1].Row(by)[bx] = qx;
      dcq[2].Row(by)[bx] = qb;
      dc_rec[1].Row(by)[bx] = ry;
      dc_rec[0].Row(by)[bx] = std::fmaf(ry, 0.0f, qx * mul_dc[0]);
      dc_rec[2].Row(by)[bx] = std::fmaf(ry, 1.0f, qb * mul_dc[2]);
    }

  // ---- AC: transform, quantise, tokenise per group and pass
  const BlockCtxMap bctx;
  const std::vector<uint8_t> clusters = ACContextClusters(bctx);
  std::vector<float> tables[17];
  auto table_for = [&](int strategy) -> const std::vector<float>& {
    const int t = kStrategyToQuantTable[strategy];
    if (tables[t].empty()) tables[t] = ComputeQuantTable(LibraryEncoding(t), t);
    return tables[t];
  };
  std::vector<uint32_t> natural[13];
  auto order_for = [&](int strategy) -> const std::vector<uint32_t>& {
    const int ord = kStrategyOrder[strategy];
    if (natural[ord].empty()) {
      natural[ord].resize(64u * kCoveredX[strategy] * kCoveredY[strategy]);
      NaturalCoeffOrder(strategy, natural[ord].data());
    }
    return natural[ord];
  };
  const size_t num_groups = dim.num_groups, num_passes = p.num_passes;
  std::vector<std::vector<Token>> group_tokens(num_groups * num_passes);
  const float* biases = kDefaultQuantBias;
  std::vector<float> coeff(3 * 65536), scratch(3 * 65536 + 1024);
  // quantised coefficients of every varblock (3 * size values at its first block), kept for the order statistics
  std::vector<std::vector<int32_t>> qblocks(W * H);
  for (size_t g = 0; g < num_groups; g++) {
    const BlockRect r = BlockGroupRect(dim, g);
    for (size_t by = 0; by < r.ys; by++) {
      for (size_t bx = 0; bx < r.xs; bx++) {
        const size_t pos = (r.y0 + by) * W + r.x0 + bx;
        const uint8_t a = acs[pos];
        if (!(a & 1)) continue;
        const int s = a >> 1;
        const size_t cx = kCoveredX[s], cy = kCoveredY[s];
        const size_t covered = cx * cy, size = covered * 64;
        const std::vector<float>& dm = table_for(s);
        qblocks[pos].assign(3 * size, 0);
        int32_t* quantized = qblocks[pos].data();
        const float sd_base = inv_global_scale / raw_quant[pos];
        const float sd[3] = {sd_base * x_dm, sd_base, sd_base * b_dm};
        const size_t tile = ((r.y0 + by) / 8) * cmw + (r.x0 + bx) / 8;
        const float x_cc = 0.0f + ytox[tile] * (1.0f / 84), b_cc = 1.0f + ytob[tile] * (1.0f / 84);
        for (int c = 0; c < 3; c++)
          TransformFromPixels(s, xyb[c].Row((r.y0 + by) * 8) + (r.x0 + bx) * 8, PW, coeff.data() + c * size, scratch.data());
        if (adaptive) {
          // QuantizeRoundtripYBlockAC + the X / B quantisation of ComputeCoefficients (lib/jxl/enc_group.cc:319-368, :455-491)
          const float scale = global_scale * (1.0f / 65536);
          const size_t lx = std::max(cx, cy), ly = std::min(cx, cy);
          const float qm_mul[3] = {std::pow(1.25f, p.x_qm_scale - 2.0f), 1.0f, std::pow(1.25f, p.b_qm_scale - 2.0f)};
          const int32_t quant_orig = raw_quant[pos];
          int32_t max_quant = 0;
          float thres_y[4] = {0.58f, 0.64f, 0.64f, 0.64f};
          for (int c : {1, 0, 2}) {
            float thres[4] = {0.58f, 0.64f, 0.64f, 0.64f};
            int32_t quant = quant_orig;
            e8::AdjustQuantBlockAC(scale, c, qm_mul[c], s, lx, ly, thres, coeff.data() + c * size, dm.data() + c * size, &quant);
            if (c == 1)
              for (int k = 0; k < 4; k++) thres_y[k] = thres[k];
            max_quant = std::max(quant, max_quant);
          }
          const int32_t quant = max_quant;
          e8::QuantizeBlockAC(scale, 1, 1.0f, lx, ly, thres_y, coeff.data() + size, dm.data() + size, quant, quantized + size);
          const float inv_qac = inv_global_scale / quant;
          for (size_t k = 0; k < size; k++) {
            const float dq_y = (AdjustQuantBias(1, quantized[size + k], biases) * dm[size + k]) * inv_qac;
            coeff[k] = std::fmaf(-x_cc, dq_y, coeff[k]);
            coeff[2 * size + k] = std::fmaf(-b_cc, dq_y, coeff[2 * size + k]);
          }
          for (int c : {0, 2}) {
            float thres[4] = {0.58f, 0.62f, 0.62f, 0.62f};
            e8::QuantizeBlockAC(scale, c, qm_mul[c], lx, ly, thres, coeff.data() + c * size, dm.data() + c * size, quant,
                                quantized + c * size);
          }
          raw_quant[pos] = quant;
        }
        // Y first (the decoder adds ratio * dequantised Y to X and B)
        for (size_t k = 0; k < size && !adaptive; k++) {
          const int32_t q = static_cast<int32_t>(std::lrintf(coeff[size + k] / (dm[size + k] * sd[1])));
          quantized[size + k] = q;
          const float dq_y = AdjustQuantBias(1, q, biases) * (dm[size + k] * sd[1]);
          quantized[k] = static_cast<int32_t>(std::lrintf((coeff[k] - x_cc * dq_y) / (dm[k] * sd[0])));
          quantized[2 * size + k] =
              static_cast<int32_t>(std::lrintf((coeff[2 * size + k] - b_cc * dq_y) / (dm[2 * size + k] * sd[2])));
        }
        // the lowest frequencies come from the DC image
        const size_t lcx = std::max(cx, cy), lcy = std::min(cx, cy);
        for (int c = 0; c < 3; c++)
          for (size_t y = 0; y < lcy; y++)
            for (size_t x = 0; x < lcx; x++) quantized[c * size + y * lcx * 8 + x] = 0;
      }
    }
  }

  // ---- coefficient orders (E9): ComputeUsedOrders + ComputeCoeffOrder, lib/jxl/enc_coeff_order.cc:47-238, as libjxl's
  // effort 7 (kSquirrel) runs them; single-pass frames only (with more passes the natural orders stay)
  std::vector<uint32_t> custom[13][3];
  uint32_t used_orders = 0;
  if (p.coeff_orders && num_passes == 1) {
    uint32_t used_acs = 0, customize = 0;
    for (size_t i = 0; i < W * H; i++) {
      if (acs[i] == 0xFF) continue;
      const uint32_t ord = kStrategyOrder[acs[i] >> 1];
      used_acs |= 1u << ord;
      if (ord <= 6) customize |= 1u << ord;  // no custom orders for blocks larger than 32x32
    }
    if (W < 5 && H < 5) customize = 0;  // default orders for small images
    if (customize != 0) {
      std::vector<int32_t> num_zeros[13][3];
      const double block_fraction = customize == 1 ? 0.5 : 1.0;  // only 8x8 DCTs: every other block is enough
      const uint64_t threshold = static_cast<uint64_t>((std::numeric_limits<uint64_t>::max() >> 32) * block_fraction);
      uint64_t rs[2] = {0x94D049BB133111EBull, 0xBF58476D1CE4E5B9ull};
      auto use_sample = [&]() {  // xorshift128+
        uint64_t s1 = rs[0];
        const uint64_t s0 = rs[1];
        const uint64_t bits = s1 + s0;
        rs[0] = s0;
        s1 ^= s1 << 23;
        s1 ^= s0 ^ (s1 >> 18) ^ (s0 >> 5);
        rs[1] = s1;
        return (bits >> 32) <= threshold;
      };
      for (size_t g = 0; g < num_groups; g++) {
        const BlockRect r = BlockGroupRect(dim, g);
        for (size_t by = 0; by < r.ys; by++)
          for (size_t bx = 0; bx < r.xs; bx++) {
            const size_t pos = (r.y0 + by) * W + r.x0 + bx;
            if (!(acs[pos] & 1)) continue;
            if (!use_sample()) continue;
            const int s = acs[pos] >> 1;
            const uint32_t ord = kStrategyOrder[s];
            size_t cx = kCoveredX[s], cy = kCoveredY[s];
            const size_t size = cx * cy * 64;
            if (cy > cx) std::swap(cx, cy);
            for (int c = 0; c < 3; c++) {
              std::vector<int32_t>& nzv = num_zeros[ord][c];
              if (nzv.empty()) nzv.assign(size, 0);
              const int32_t* q = qblocks[pos].data() + c * size;
              for (size_t k = 0; k < size; k++) nzv[k] += q[k] == 0 ? 1 : 0;
              for (size_t iy = 0; iy < cy; iy++)  // the lowest frequencies stay first
                for (size_t ix = 0; ix < cx; ix++) nzv[iy * 8 * cx + ix] = -1;
            }
          }
      }
      used_orders = customize;
      for (int o = 0; o < kNumStrategies; o++) {
        const uint32_t ord = kStrategyOrder[o];
        if (!(used_orders & (1u << ord)) || !custom[ord][0].empty()) continue;
        if (!(used_acs & (1u << ord))) continue;
        const size_t sz = 64u * kCoveredX[o] * kCoveredY[o];
        const std::vector<uint32_t>& natural_order = order_for(o);
        bool is_nondefault = false;
        for (int c = 0; c < 3; c++) {
          struct PosAndCount { uint32_t pos, count; };
          std::vector<PosAndCount> pv(sz);
          const float inv_sqrt_sz = 1.0f / std::sqrt(static_cast<float>(sz));
          for (size_t i = 0; i < sz; i++) {
            const uint32_t pos = natural_order[i];
            pv[i].pos = pos;
            const int32_t nzc = num_zeros[ord][c].empty() ? 0 : num_zeros[ord][c][pos];
            const float q = nzc * inv_sqrt_sz + 0.1f;  // quantised counts: a less permuted order
            pv[i].count = q <= 0.0f ? 0u : static_cast<uint32_t>(q);
          }
          std::stable_sort(pv.begin(), pv.end(), [](const PosAndCount& a, const PosAndCount& b) { return a.count < b.count; });
          custom[ord][c].resize(sz);
          for (size_t i = 0; i < sz; i++) {
            custom[ord][c][i] = pv[i].pos;
            is_nondefault |= natural_order[i] != pv[i].pos;
          }
        }
        if (!is_nondefault) {
          used_orders &= ~(1u << ord);
          for (int c = 0; c < 3; c++) custom[ord][c].clear();
        }
      }
      // (an order bit whose strategies do not occur keeps no custom order)
      for (uint32_t ord = 0; ord < 13; ord++)
        if ((used_orders & (1u << ord)) && custom[ord][0].empty()) used_orders &= ~(1u << ord);
    }
  }

  // ---- tokens per group and pass
  for (size_t g = 0; g < num_groups; g++) {
    const BlockRect r = BlockGroupRect(dim, g);
    std::vector<std::vector<int32_t>> nzeros(num_passes, std::vector<int32_t>(3 * 32 * 32, 0));
    for (size_t by = 0; by < r.ys; by++) {
      for (size_t bx = 0; bx < r.xs; bx++) {
        const size_t pos = (r.y0 + by) * W + r.x0 + bx;
        const uint8_t a = acs[pos];
        if (!(a & 1)) continue;
        const int s = a >> 1;
        const size_t cx = kCoveredX[s], cy = kCoveredY[s];
        const size_t covered = cx * cy, size = covered * 64, log2c = kLog2Covered[s];
        const int32_t* quantized = qblocks[pos].data();
        // tokens, channel order Y, X, B; pass i carries (value >> shift_i) - (what earlier passes carried)
        for (int c : {1, 0, 2}) {
          for (size_t pass = 0; pass < num_passes; pass++) {
            const uint32_t shift = fh.passes.shift[pass];
            const uint32_t prev_shift = pass == 0 ? 32 : fh.passes.shift[pass - 1];
            auto pass_value = [&](int32_t v) -> int32_t {
              // sum over passes of (part << shift) must equal v: part_i = (v >> s_i) - ((v >> s_{i-1}) << (s_{i-1} - s_i))
              const int64_t hi = prev_shift >= 32 ? 0 : (static_cast<int64_t>(v) >> prev_shift);
              const int64_t cur = static_cast<int64_t>(v) >> shift;
              return static_cast<int32_t>(cur - (prev_shift >= 32 ? 0 : (hi << (prev_shift - shift))));
            };
            std::vector<Token>& toks = group_tokens[pass * num_groups + g];
            int32_t* row_nz = &nzeros[pass][(c * 32 + by) * 32];
            const int32_t* row_top = by == 0 ? nullptr : row_nz - 32;
            const int32_t predicted = PredictFromTopAndLeft(row_top, row_nz, bx, 32);
            const size_t ord = kStrategyOrder[s];
            const std::vector<uint32_t>& order = custom[ord][c].empty() ? order_for(s) : custom[ord][c];
            const size_t block_ctx = bctx.Context(0, raw_quant[(r.y0 + by) * W + r.x0 + bx], ord, c);
            size_t nz = 0;
            for (size_t k = covered; k < size; k++) nz += pass_value(quantized[c * size + order[k]]) != 0;
            toks.push_back({bctx.NonZeroContext(predicted, block_ctx), static_cast<uint32_t>(nz)});
            for (size_t y = 0; y < cy; y++)
              for (size_t x = 0; x < cx; x++) row_nz[bx + x + y * 32] = (nz + covered - 1) >> log2c;
            const size_t histo_offset = bctx.ZeroDensityContextsOffset(block_ctx);
            size_t prev = (nz > size / 16 ? 0 : 1);
            for (size_t k = covered; k < size && nz != 0; ++k) {
              const size_t nzl = (nz + covered - 1) >> log2c;
              const size_t ctx = histo_offset + (kCoeffNumNonzeroContext[nzl] + kCoeffFreqContext[k >> log2c]) * 2 + prev;
              const int32_t v = pass_value(quantized[c * size + order[k]]);
              const uint32_t u = PackSigned(v);
              toks.push_back({static_cast<uint32_t>(ctx), u});
              prev = u != 0;
              nz -= prev;
            }
          }
        }
      }
    }
  }

  // ---- sections
  std::vector<std::vector<uint8_t>> sections;
  auto finish = [&](BitWriter& w) {
    w.ZeroPadToByte();
    sections.push_back(w.Bytes());
  };
  // Modular sub-streams (DC, AC metadata) under one global tree and one entropy code
  const GlobalTree gtree = BuildGlobalTree(dim.num_dc_groups, p.dc_tree);
  std::vector<std::vector<Channel>> dc_chans(dim.num_dc_groups), meta_chans(dim.num_dc_groups);
  std::vector<std::vector<Token>> dc_toks(dim.num_dc_groups), meta_toks(dim.num_dc_groups);
  std::vector<size_t> meta_count(dim.num_dc_groups, 0);
  for (size_t g = 0; g < dim.num_dc_groups; g++) {
    const BlockRect r = DCGroupRect(dim, g);
    std::vector<Channel> ch = {Channel(r.xs, r.ys), Channel(r.xs, r.ys), Channel(r.xs, r.ys)};
    for (int c = 0; c < 3; c++)
      for (size_t y = 0; y < r.ys; y++)
        for (size_t x = 0; x < r.xs; x++) ch[c].Row(y)[x] = dcq[c].Row(r.y0 + y)[r.x0 + x];
    size_t count = 0;
    for (size_t y = 0; y < r.ys; y++)
      for (size_t x = 0; x < r.xs; x++) count += acs[(r.y0 + y) * W + r.x0 + x] & 1;
    const size_t cx0 = r.x0 >> 3, cy0 = r.y0 >> 3, cw = (r.xs + 7) >> 3, chh = (r.ys + 7) >> 3;
    std::vector<Channel> meta = {Channel(cw, chh, 3, 3), Channel(cw, chh, 3, 3), Channel(count, 2), Channel(r.xs, r.ys)};
    for (size_t y = 0; y < chh; y++)
      for (size_t x = 0; x < cw; x++) {
        meta[0].Row(y)[x] = ytox[(cy0 + y) * cmw + cx0 + x];
        meta[1].Row(y)[x] = ytob[(cy0 + y) * cmw + cx0 + x];
      }
    size_t num = 0;
    for (size_t y = 0; y < r.ys; y++)
      for (size_t x = 0; x < r.xs; x++) {
        const size_t pos = (r.y0 + y) * W + r.x0 + x;
        meta[3].Row(y)[x] = sharp[pos];
        if (!(acs[pos] & 1)) continue;
        meta[2].Row(0)[num] = acs[pos] >> 1;
        meta[2].Row(1)[num] = raw_quant[pos] - 1;
        num++;
      }
    TokenizeModularStream(gtree.tree, StreamVarDCTDC(dim, g), ch, &dc_toks[g]);
    TokenizeModularStream(gtree.tree, StreamACMetadata(dim, g), meta, &meta_toks[g]);
    dc_chans[g] = std::move(ch);
    meta_chans[g] = std::move(meta);
    meta_count[g] = count;
  }
  // The alpha channel: one Modular channel of the frame's size. It travels in the global stream when it fits one group,
  // else in the per-group streams that follow the coefficients of each AC group (single pass: shift bracket 0 .. 2).
  std::vector<Channel> alpha_global;                // what the global stream's header is followed by
  std::vector<std::vector<Channel>> alpha_group(num_groups);
  std::vector<Token> alpha_global_toks;
  std::vector<std::vector<Token>> alpha_group_toks(num_groups);
  if (p.alpha) {
    JXLO_CHECK(num_passes == 1 && p.upsampling == 1, "alpha: single-pass frames without upsampling only");
    if (xsize <= dim.group_dim && ysize <= dim.group_dim) {
      Channel a(xsize, ysize);
      for (uint32_t y = 0; y < ysize; y++)
        for (uint32_t x = 0; x < xsize; x++) a.Row(y)[x] = p.alpha[static_cast<size_t>(y) * xsize + x];
      alpha_global.push_back(std::move(a));
      TokenizeModularStream(gtree.tree, 0, alpha_global, &alpha_global_toks);
    } else {
      alpha_global.push_back(Channel(0, 0));  // the header is written, no channel is small enough for the global stream
      for (size_t g = 0; g < num_groups; g++) {
        const size_t gx = g % dim.xsize_groups, gy = g / dim.xsize_groups;
        const size_t x0 = gx * dim.group_dim, y0 = gy * dim.group_dim;
        const size_t w = std::min<size_t>(dim.group_dim, xsize - x0), h = std::min<size_t>(dim.group_dim, ysize - y0);
        Channel a(w, h);
        for (size_t y = 0; y < h; y++)
          for (size_t x = 0; x < w; x++) a.Row(y)[x] = p.alpha[(y0 + y) * xsize + x0 + x];
        alpha_group[g].push_back(std::move(a));
        TokenizeModularStream(gtree.tree, StreamModularAC(dim, g, 0), alpha_group[g], &alpha_group_toks[g]);
      }
    }
  }
  EntropyOptions eopt;
  eopt.use_prefix = p.prefix_codes;
  EntropyEncoder tree_code(6, {0, 1, 2, 3, 4, 5}, eopt);
  tree_code.Count(gtree.tokens);
  std::vector<uint8_t> leaf_clusters(gtree.num_leaves);
  for (size_t i = 0; i < gtree.num_leaves; i++) leaf_clusters[i] = static_cast<uint8_t>(i);
  EntropyEncoder modular_code(gtree.num_leaves, leaf_clusters, eopt);
  for (size_t g = 0; g < dim.num_dc_groups; g++) {
    modular_code.Count(dc_toks[g]);
    modular_code.Count(meta_toks[g]);
  }
  modular_code.Count(alpha_global_toks);
  for (size_t g = 0; g < num_groups; g++) modular_code.Count(alpha_group_toks[g]);
  BitWriter dc_global;
  {
    if (p.splines) WriteRandomSplines(dc_global, p.splines, xsize, ysize, p.seed, eopt);
    dc_global.Write(1, 1);  // default DC quantisation
    WriteU32(dc_global, global_scale, BitsOffset(11, 1), BitsOffset(11, 2049), BitsOffset(12, 4097), BitsOffset(16, 8193));
    WriteU32(dc_global, quant_dc, Val(16), BitsOffset(5, 1), BitsOffset(8, 1), BitsOffset(16, 1));
    dc_global.Write(1, 1);  // default block context map
    dc_global.Write(1, 1);  // default colour correlation
    dc_global.Write(1, 1);  // global MA tree
    tree_code.WriteHeader(dc_global);
    tree_code.WriteTokens(dc_global, gtree.tokens);
    modular_code.WriteHeader(dc_global);
    WriteModularStream(dc_global, modular_code, alpha_global, alpha_global_toks);  // (nothing without extra channels)
  }
  std::vector<BitWriter> dc_groups(dim.num_dc_groups);
  for (size_t g = 0; g < dim.num_dc_groups; g++) {
    BitWriter& w = dc_groups[g];
    const BlockRect r = DCGroupRect(dim, g);
    w.Write(2, 0);  // extra_precision
    WriteModularStream(w, modular_code, dc_chans[g], dc_toks[g]);
    // (no Modular DC-group channels)
    w.Write(CeilLog2(r.xs * r.ys), meta_count[g] - 1);
    WriteModularStream(w, modular_code, meta_chans[g], meta_toks[g]);
  }
  BitWriter ac_global;
  std::vector<EntropyEncoder> pass_codes;
  {
    ac_global.Write(1, 1);  // default quantisation matrices
    ac_global.Write(CeilLog2(num_groups), 0);  // one set of histograms
    for (size_t pass = 0; pass < num_passes; pass++) {
      WriteU32(ac_global, used_orders, Val(0x5F), Val(0x13), Val(0), Bits(13));
      if (used_orders != 0) {
        // EncodeCoeffOrders, lib/jxl/enc_coeff_order.cc:293-337: per order and channel the permutation relative to the
        // natural order as a Lehmer code (trailing zeros dropped), one ANS stream over the 8 permutation contexts
        std::vector<Token> ptoks;
        for (int o = 0; o < kNumStrategies; o++) {
          const uint32_t ord = kStrategyOrder[o];
          if (!(used_orders & (1u << ord)) || kOrderFirstStrategyOf(ord) != o) continue;
          const size_t llf = kCoveredX[o] * kCoveredY[o], sz = 64 * llf;
          const std::vector<uint32_t>& nat = order_for(o);
          std::vector<uint32_t> lut(sz);
          for (size_t i = 0; i < sz; i++) lut[nat[i]] = i;
          for (int c = 0; c < 3; c++) {
            std::vector<uint32_t> zz(sz), lehmer(sz, 0);
            for (size_t i = 0; i < sz; i++) zz[i] = lut[custom[ord][c][i]];
            // Lehmer code: how many later elements are smaller (lib/jxl/lehmer_code.h)
            std::vector<uint32_t> avail(sz);
            for (size_t i = 0; i < sz; i++) avail[i] = i;
            for (size_t i = 0; i < sz; i++) {
              const auto it = std::lower_bound(avail.begin(), avail.end(), zz[i]);
              lehmer[i] = static_cast<uint32_t>(it - avail.begin());
              avail.erase(it);
            }
            size_t end = sz;
            while (end > llf && lehmer[end - 1] == 0) end--;
            ptoks.push_back({CoeffOrderContext(sz), static_cast<uint32_t>(end - llf)});
            uint32_t last = 0;
            for (size_t i = llf; i < end; i++) {
              ptoks.push_back({CoeffOrderContext(last), lehmer[i]});
              last = lehmer[i];
            }
          }
        }
        EntropyEncoder perm_code(8, {0, 1, 2, 3, 4, 5, 6, 7}, eopt);
        perm_code.Count(ptoks);
        perm_code.WriteHeader(ac_global);
        perm_code.WriteTokens(ac_global, ptoks);
      }
      pass_codes.emplace_back(bctx.NumACContexts(), clusters, eopt);
      for (size_t g = 0; g < num_groups; g++) pass_codes.back().Count(group_tokens[pass * num_groups + g]);
      pass_codes.back().WriteHeader(ac_global);
    }
  }
  std::vector<BitWriter> ac_groups(num_groups * num_passes);
  for (size_t pass = 0; pass < num_passes; pass++)
    for (size_t g = 0; g < num_groups; g++) {
      pass_codes[pass].WriteTokens(ac_groups[pass * num_groups + g], group_tokens[pass * num_groups + g]);
      if (pass == 0) WriteModularStream(ac_groups[g], modular_code, alpha_group[g], alpha_group_toks[g]);
    }

  if (num_groups == 1 && num_passes == 1) {
    BitWriter all = dc_global;
    auto append_bits = [&](BitWriter& src) {
      const size_t n = src.BitsWritten();
      const std::vector<uint8_t>& b = src.Bytes();
      for (size_t i = 0; i < n; i++) all.Write(1, (b[i >> 3] >> (i & 7)) & 1);
    };
    append_bits(dc_groups[0]);
    append_bits(ac_global);
    append_bits(ac_groups[0]);
    finish(all);
  } else {
    finish(dc_global);
    for (auto& w : dc_groups) finish(w);
    finish(ac_global);
    for (auto& w : ac_groups) finish(w);
  }

  BitWriter out;
  WriteImageHeaders(out, xsize * p.upsampling, ysize * p.upsampling, p.orientation, p.alpha != nullptr);
  WriteFrameHeader(out, p);
  WriteToc(out, sections);
  for (const auto& s : sections) out.Append(s);
  if (stats) {
    stats->num_groups = num_groups;
    stats->num_varblocks = num_varblocks;
    stats->bytes = out.Bytes().size();
  }
  return out.Bytes();
}

}  // namespace jxlo

#endif  // JXLO_ENCODE_H_
