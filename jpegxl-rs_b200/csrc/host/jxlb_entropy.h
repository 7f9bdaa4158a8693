// jxl_b200 host parser (product code; runs on the CPU in front of the CUDA kernels).
// Host side of the entropy coder: parses histogram / prefix-code / context-map headers
// into tables the device kernels consume. Symbol decoding of the small host-parsed
// streams (context maps, MA trees, permutations) also runs here.
//
// Entropy-coded stream decoding: histogram / prefix-code headers, context map,
// rANS + alias table, hybrid-uint, LZ77. Restates
//   lib/jxl/dec_ans.cc:51-187 (ReadHistogram), :189-262 (DecodeANSCodes),
//   :264-300 (uint configs), :336-368 (DecodeHistograms),
//   lib/jxl/ans_common.cc:54-160 (InitAliasTable), lib/jxl/ans_common.h:91-138,
//   lib/jxl/dec_ans.h:168-195 (rANS step), :223-255 (hybrid uint), :286-343 (LZ77),
//   lib/jxl/dec_context_map.cc:46-96, lib/jxl/dec_huffman.cc:23-240.
#ifndef JXLB_ENTROPY_H_
#define JXLB_ENTROPY_H_

#include <algorithm>
#include <vector>

#include "jxlb_bits.h"

namespace jxlb {

constexpr int kAnsLogTabSize = 12;
constexpr uint32_t kAnsTabSize = 1u << kAnsLogTabSize;
constexpr uint32_t kAnsSignature = 0x13;  // lib/jxl/ans_params.h:32
constexpr int kPrefixMaxBits = 15;
constexpr size_t kLZ77Window = 1u << 20;
constexpr uint32_t kNumSpecialDistances = 120;

struct HybridUintConfig {
  uint32_t split_exponent = 0, split_token = 1, msb_in_token = 0, lsb_in_token = 0;
  HybridUintConfig() = default;
  HybridUintConfig(uint32_t se, uint32_t msb, uint32_t lsb)
      : split_exponent(se), split_token(1u << se), msb_in_token(msb), lsb_in_token(lsb) {}
};

struct AliasEntry {
  uint8_t cutoff;
  uint8_t right_value;
  uint16_t freq0;
  uint16_t offsets1;
  uint16_t freq1_xor_freq0;
};

// Prefix code as a two-level lookup table (8 root bits).
// Entry: low 16 bits = symbol or sub-table offset, bits 16..23 = code length
// (for a link: number of index bits of the sub-table), bit 31 = link.
struct PrefixTable {
  std::vector<uint32_t> t;
};

struct EntropyCode {
  bool lz77_enabled = false;
  uint32_t lz77_min_symbol = 224, lz77_min_length = 3;
  HybridUintConfig lz77_length_cfg;
  uint32_t lz77_dist_cluster = 0;
  std::vector<uint8_t> ctx_map;  // context -> cluster
  uint32_t num_clusters = 1;
  bool use_prefix = false;
  uint32_t log_alpha_size = 5;
  std::vector<HybridUintConfig> cfg;  // per cluster
  std::vector<AliasEntry> alias;      // num_clusters << log_alpha_size
  std::vector<PrefixTable> prefix;    // per cluster
  std::vector<int> degenerate;        // per cluster: the only symbol, or -1
  uint32_t max_num_bits = 0;
};

inline int ReadVarLenUint8(BitReader& br) {
  if (!br.Read(1)) return 0;
  int n = br.Read(3);
  return n == 0 ? 1 : static_cast<int>(br.Read(n)) + (1 << n);
}
inline int ReadVarLenUint16(BitReader& br) {
  if (!br.Read(1)) return 0;
  int n = br.Read(4);
  return n == 0 ? 1 : static_cast<int>(br.Read(n)) + (1 << n);
}

// lib/jxl/dec_ans.cc:264-290
inline HybridUintConfig ReadUintConfig(BitReader& br, uint32_t log_alpha_size) {
  uint32_t se = br.Read(CeilLog2(log_alpha_size + 1));
  uint32_t msb = 0, lsb = 0;
  if (se != log_alpha_size) {
    msb = br.Read(CeilLog2(se + 1));
    JXLB_CHECK(msb <= se, "bad uint config");
    lsb = br.Read(CeilLog2(se - msb + 1));
  }
  JXLB_CHECK(msb + lsb <= se, "bad uint config");
  return HybridUintConfig(se, msb, lsb);
}

// ---- ANS histogram (lib/jxl/dec_ans.cc:51-187) ----
inline uint32_t PopulationCountPrecision(uint32_t logcount, uint32_t shift) {
  int r = std::min<int>(logcount, static_cast<int>(shift) -
                                      static_cast<int>((kAnsLogTabSize - logcount) >> 1));
  return r < 0 ? 0 : r;
}

inline std::vector<int32_t> ReadAnsHistogram(BitReader& br) {
  const int range = kAnsTabSize;
  std::vector<int32_t> counts;
  if (br.Read(1)) {  // 1 or 2 symbols
    int n = br.Read(1) + 1;
    int sym[2] = {0, 0};
    for (int i = 0; i < n; i++) sym[i] = ReadVarLenUint8(br);
    counts.assign(std::max(sym[0], sym[1]) + 1, 0);
    if (n == 1) {
      counts[sym[0]] = range;
    } else {
      JXLB_CHECK(sym[0] != sym[1], "histogram: duplicate symbol");
      counts[sym[0]] = br.Read(kAnsLogTabSize);
      counts[sym[1]] = range - counts[sym[0]];
    }
    return counts;
  }
  if (br.Read(1)) {  // flat
    int n = ReadVarLenUint8(br) + 1;
    counts.assign(n, range / n);
    for (int i = 0; i < range % n; i++) counts[i]++;
    return counts;
  }
  // general: unary-coded shift, then per-symbol log-counts through a fixed prefix code.
  int log = 0;
  const int upper = FloorLog2(kAnsLogTabSize + 1);
  for (; log < upper; log++) {
    if (!br.Read(1)) break;
  }
  uint32_t shift = (br.Read(log) | (1u << log)) - 1;
  JXLB_CHECK(shift <= kAnsLogTabSize + 1, "histogram: bad shift");
  int length = ReadVarLenUint8(br) + 3;
  counts.assign(length, 0);
  // The fixed code for log-counts: (code bits LSB-first, length) -> value.
  // Derived from the 128-entry lookup at lib/jxl/dec_ans.cc:101-118.
  static const uint8_t kLen[14] = {5, 4, 4, 4, 4, 4, 3, 3, 3, 3, 3, 6, 7, 7};
  static const uint8_t kCode[14] = {17, 11, 15, 3, 9, 7, 4, 2, 5, 6, 0, 33, 1, 65};
  std::vector<int> logcounts(length, 0), same(length, 0);
  int omit_log = -1, omit_pos = -1;
  for (int i = 0; i < length; i++) {
    uint32_t w = br.Peek(7);
    int v = -1;
    for (int s = 0; s < 14; s++) {
      if ((w & ((1u << kLen[s]) - 1)) == kCode[s]) {
        v = s;
        break;
      }
    }
    JXLB_CHECK(v >= 0, "histogram: bad log-count code");
    br.Skip(kLen[v]);
    logcounts[i] = v;
    if (v == kAnsLogTabSize + 1) {  // RLE
      int rle = ReadVarLenUint8(br);
      same[i] = rle + 5;
      i += rle + 3;
      continue;
    }
    if (v > omit_log) {
      omit_log = v;
      omit_pos = i;
    }
  }
  JXLB_CHECK(omit_pos >= 0, "histogram: no omit position");
  JXLB_CHECK(!(omit_pos + 1 < length && logcounts[omit_pos + 1] == kAnsLogTabSize + 1),
             "histogram: RLE after omitted symbol");
  int prev = 0, numsame = 0, total = 0;
  for (int i = 0; i < length; i++) {
    if (same[i]) {
      numsame = same[i] - 1;
      prev = i > 0 ? counts[i - 1] : 0;
    }
    if (numsame > 0) {
      counts[i] = prev;
      numsame--;
    } else {
      int code = logcounts[i];
      if (i == omit_pos || code == 0) continue;
      if (code == 1) {
        counts[i] = 1;
      } else {
        int bitcount = PopulationCountPrecision(code - 1, shift);
        counts[i] = (1 << (code - 1)) + (br.Read(bitcount) << (code - 1 - bitcount));
      }
    }
    total += counts[i];
  }
  counts[omit_pos] = range - total;
  JXLB_CHECK(counts[omit_pos] > 0, "histogram: bad omitted count");
  return counts;
}

// lib/jxl/ans_common.cc:54-160. The construction order is normative: the
// encoder's reverse map is derived from the same table.
inline void BuildAliasTable(std::vector<int32_t> dist, uint32_t log_alpha, AliasEntry* a) {
  const uint32_t table_size = 1u << log_alpha;
  while (!dist.empty() && dist.back() == 0) dist.pop_back();
  if (dist.empty()) dist.push_back(kAnsTabSize);
  JXLB_CHECK(dist.size() <= table_size, "alias: alphabet too large");
  const uint32_t entry_size = kAnsTabSize >> log_alpha;
  int single = -1;
  uint32_t sum = 0;
  for (size_t s = 0; s < dist.size(); s++) {
    sum += dist[s];
    if (dist[s] == static_cast<int32_t>(kAnsTabSize)) single = static_cast<int>(s);
  }
  JXLB_CHECK(sum == kAnsTabSize, "alias: histogram does not sum to 4096");
  if (single >= 0) {
    for (uint32_t i = 0; i < table_size; i++) {
      a[i].right_value = static_cast<uint8_t>(single);
      a[i].cutoff = 0;
      a[i].offsets1 = static_cast<uint16_t>(entry_size * i);
      a[i].freq0 = 0;
      a[i].freq1_xor_freq0 = kAnsTabSize;
    }
    return;
  }
  std::vector<uint32_t> under, over, cut(table_size, 0);
  for (size_t i = 0; i < dist.size(); i++) {
    cut[i] = dist[i];
    if (cut[i] > entry_size) over.push_back(i);
    else if (cut[i] < entry_size) under.push_back(i);
  }
  for (uint32_t i = dist.size(); i < table_size; i++) under.push_back(i);
  std::vector<uint32_t> right(table_size, 0), off1(table_size, 0);
  while (!over.empty()) {
    uint32_t o = over.back();
    over.pop_back();
    JXLB_CHECK(!under.empty(), "alias: inconsistent histogram");
    uint32_t u = under.back();
    under.pop_back();
    uint32_t by = entry_size - cut[u];
    cut[o] -= by;
    right[u] = o;
    off1[u] = cut[o];
    if (cut[o] < entry_size) under.push_back(o);
    else if (cut[o] > entry_size) over.push_back(o);
  }
  for (uint32_t i = 0; i < table_size; i++) {
    if (cut[i] == entry_size) {
      right[i] = i;
      off1[i] = 0;
      a[i].cutoff = 0;
    } else {
      off1[i] -= cut[i];
      a[i].cutoff = static_cast<uint8_t>(cut[i]);
    }
    a[i].right_value = static_cast<uint8_t>(right[i]);
    a[i].offsets1 = static_cast<uint16_t>(off1[i]);
    uint32_t f0 = i < dist.size() ? dist[i] : 0;
    uint32_t f1 = right[i] < dist.size() ? dist[right[i]] : 0;
    a[i].freq0 = static_cast<uint16_t>(f0);
    a[i].freq1_xor_freq0 = static_cast<uint16_t>(f1 ^ f0);
  }
}

// ---- prefix codes (lib/jxl/dec_huffman.cc, lib/jxl/huffman_table.cc) ----
constexpr uint32_t kPrefixRootBits = 8;
constexpr uint32_t kPrefixLink = 0x80000000u;

inline uint32_t ReverseBits(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
  return r;
}

// Canonical code: shorter codes first, ties by symbol value; the stream stores
// codes MSB-first but is read LSB-first, hence the bit reversal.
inline void BuildPrefixTable(const std::vector<uint8_t>& len, uint32_t root_bits,
                             std::vector<uint32_t>* out) {
  const size_t n = len.size();
  int nonzero = 0, last = 0;
  for (size_t s = 0; s < n; s++) {
    if (len[s]) {
      nonzero++;
      last = static_cast<int>(s);
    }
  }
  const uint32_t root_size = 1u << root_bits;
  out->assign(root_size, 0);
  if (nonzero == 0) return;  // symbol 0, zero bits
  if (nonzero == 1) {
    std::fill(out->begin(), out->end(), static_cast<uint32_t>(last));
    return;
  }
  uint32_t count[kPrefixMaxBits + 2] = {0}, next[kPrefixMaxBits + 2] = {0};
  for (size_t s = 0; s < n; s++) count[len[s]]++;
  count[0] = 0;
  uint32_t code = 0;
  for (int l = 1; l <= kPrefixMaxBits; l++) {
    code = (code + count[l - 1]) << 1;
    next[l] = code;
  }
  std::vector<uint32_t> rev(n, 0);
  for (size_t s = 0; s < n; s++) {
    if (len[s]) rev[s] = ReverseBits(next[len[s]]++, len[s]);
  }
  // widths of the second-level tables, keyed by the root prefix
  std::vector<uint8_t> sub_bits(root_size, 0);
  for (size_t s = 0; s < n; s++) {
    if (len[s] > root_bits) {
      uint32_t low = rev[s] & (root_size - 1);
      sub_bits[low] = std::max<uint8_t>(sub_bits[low], len[s] - root_bits);
    }
  }
  std::vector<uint32_t> sub_off(root_size, 0);
  for (uint32_t i = 0; i < root_size; i++) {
    if (sub_bits[i]) {
      sub_off[i] = out->size();
      out->resize(out->size() + (size_t{1} << sub_bits[i]), 0);
      (*out)[i] = kPrefixLink | (static_cast<uint32_t>(sub_bits[i]) << 16) | sub_off[i];
      JXLB_CHECK(sub_off[i] < 65536, "prefix table too large");
    }
  }
  for (size_t s = 0; s < n; s++) {
    uint32_t l = len[s];
    if (!l) continue;
    if (l <= root_bits) {
      for (uint32_t k = rev[s]; k < root_size; k += 1u << l)
        (*out)[k] = (l << 16) | static_cast<uint32_t>(s);
    } else {
      uint32_t low = rev[s] & (root_size - 1);
      uint32_t hi = rev[s] >> root_bits;
      uint32_t sl = l - root_bits;
      for (uint32_t k = hi; k < (1u << sub_bits[low]); k += 1u << sl)
        (*out)[sub_off[low] + k] = (l << 16) | static_cast<uint32_t>(s);
    }
  }
}

inline uint32_t ReadPrefixSymbol(const PrefixTable& pt, BitReader& br) {
  uint64_t w = br.Window();
  uint32_t e = pt.t[w & ((1u << kPrefixRootBits) - 1)];
  if (e & kPrefixLink) {
    uint32_t bits = (e >> 16) & 0xFF;
    e = pt.t[(e & 0xFFFF) + ((w >> kPrefixRootBits) & ((1u << bits) - 1))];
  }
  br.Skip((e >> 16) & 0xFF);
  return e & 0xFFFF;
}

// lib/jxl/dec_huffman.cc:188-240 and :23-101
inline void ReadPrefixCode(BitReader& br, uint32_t alphabet_size, PrefixTable* pt) {
  if (alphabet_size <= 1) {
    pt->t.assign(1u << kPrefixRootBits, 0);
    return;
  }
  uint32_t hskip = br.Read(2);
  std::vector<uint8_t> len(alphabet_size, 0);
  if (hskip == 1) {  // simple code: 1..4 explicit symbols
    uint32_t max_bits = FloorLog2(alphabet_size - 1) + 1;
    uint32_t nsym = br.Read(2) + 1;
    uint32_t sym[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < nsym; i++) {
      sym[i] = br.Read(max_bits);
      JXLB_CHECK(sym[i] < alphabet_size, "prefix: symbol out of range");
    }
    for (uint32_t i = 0; i + 1 < nsym; i++)
      for (uint32_t j = i + 1; j < nsym; j++)
        JXLB_CHECK(sym[i] != sym[j], "prefix: duplicate symbol");
    bool tree_select = nsym == 4 ? br.Read(1) : false;
    switch (nsym) {
      case 1: len[sym[0]] = 0; break;  // zero-bit code, handled below
      case 2: len[sym[0]] = len[sym[1]] = 1; break;
      case 3: len[sym[0]] = 1; len[sym[1]] = len[sym[2]] = 2; break;
      case 4:
        if (!tree_select) {
          len[sym[0]] = len[sym[1]] = len[sym[2]] = len[sym[3]] = 2;
        } else {
          len[sym[0]] = 1; len[sym[1]] = 2; len[sym[2]] = len[sym[3]] = 3;
        }
        break;
    }
    if (nsym == 1) {
      pt->t.assign(1u << kPrefixRootBits, sym[0]);
      return;
    }
    BuildPrefixTable(len, kPrefixRootBits, &pt->t);
    return;
  }
  // complex code: code-length code first
  static const uint8_t kOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  // fixed code for the code-length-code lengths: value -> (bits LSB-first, length)
  // 0:'00' 3:'10'(LSB-first 01) 4:'01'... expressed as a 16-entry lookup on 4 peeked bits.
  static const uint8_t kClLen[16] = {2, 2, 2, 3, 2, 2, 2, 4, 2, 2, 2, 3, 2, 2, 2, 4};
  static const uint8_t kClVal[16] = {0, 4, 3, 2, 0, 4, 3, 1, 0, 4, 3, 2, 0, 4, 3, 5};
  std::vector<uint8_t> cl(18, 0);
  int space = 32, num_codes = 0;
  for (uint32_t i = hskip; i < 18 && space > 0; i++) {
    uint32_t w = br.Peek(4);
    br.Skip(kClLen[w]);
    uint8_t v = kClVal[w];
    cl[kOrder[i]] = v;
    if (v) {
      space -= 32 >> v;
      num_codes++;
    }
  }
  JXLB_CHECK(num_codes == 1 || space == 0, "prefix: bad code-length code");
  std::vector<uint32_t> cl_table;
  BuildPrefixTable(cl, 5, &cl_table);
  uint32_t symbol = 0, prev_len = 8, repeat = 0, repeat_len = 0;
  int sp = 32768;
  while (symbol < alphabet_size && sp > 0) {
    uint32_t e = cl_table[br.Peek(5)];
    br.Skip((e >> 16) & 0xFF);
    uint32_t v = e & 0xFFFF;
    if (v < 16) {
      repeat = 0;
      len[symbol++] = static_cast<uint8_t>(v);
      if (v) {
        prev_len = v;
        sp -= 32768 >> v;
      }
    } else {
      uint32_t extra = v - 14;
      uint32_t new_len = v == 16 ? prev_len : 0;
      if (repeat_len != new_len) {
        repeat = 0;
        repeat_len = new_len;
      }
      uint32_t old = repeat;
      if (repeat > 0) repeat = (repeat - 2) << extra;
      repeat += br.Read(extra) + 3;
      uint32_t delta = repeat - old;
      JXLB_CHECK(symbol + delta <= alphabet_size, "prefix: repeat overflows alphabet");
      for (uint32_t k = 0; k < delta; k++) len[symbol++] = static_cast<uint8_t>(repeat_len);
      if (repeat_len) sp -= static_cast<int>(delta << (15 - repeat_len));
    }
  }
  JXLB_CHECK(sp == 0, "prefix: code is not complete");
  BuildPrefixTable(len, kPrefixRootBits, &pt->t);
}

inline void UpdateMaxNumBits(EntropyCode* c, size_t cluster, uint32_t symbol) {
  const HybridUintConfig* cfg = &c->cfg[cluster];
  if (c->lz77_enabled && c->lz77_dist_cluster != cluster && symbol >= c->lz77_min_symbol) {
    symbol -= c->lz77_min_symbol;
    cfg = &c->lz77_length_cfg;
  }
  if (symbol < cfg->split_token) {
    c->max_num_bits = std::max(c->max_num_bits, cfg->split_exponent);
    return;
  }
  uint32_t in_token = cfg->msb_in_token + cfg->lsb_in_token;
  uint32_t extra = cfg->split_exponent - in_token + ((symbol - cfg->split_token) >> in_token);
  c->max_num_bits = std::max(c->max_num_bits, in_token + extra + 1);
}

void ReadEntropyCode(BitReader& br, size_t num_contexts, EntropyCode* code,
                     bool disallow_lz77 = false);

class SymbolReader {
 public:
  SymbolReader(const EntropyCode* code, BitReader& br, uint32_t distance_multiplier = 0)
      : c_(code) {
    state_ = code->use_prefix ? (kAnsSignature << 16) : br.Read(32);
    if (code->lz77_enabled) {
      window_.assign(kLZ77Window, 0);
      num_special_ = distance_multiplier == 0 ? 0 : kNumSpecialDistances;
      dist_mult_ = distance_multiplier;
    }
  }

  uint32_t ReadSymbol(uint32_t cluster, BitReader& br) {
    if (c_->use_prefix) return ReadPrefixSymbol(c_->prefix[cluster], br);
    const uint32_t log_entry = kAnsLogTabSize - c_->log_alpha_size;
    const uint32_t res = state_ & (kAnsTabSize - 1);
    const AliasEntry& e = c_->alias[(cluster << c_->log_alpha_size) + (res >> log_entry)];
    const uint32_t pos = res & ((1u << log_entry) - 1);
    const bool right = pos >= e.cutoff;
    const uint32_t sym = right ? e.right_value : (res >> log_entry);
    const uint32_t offset = (right ? e.offsets1 : 0) + pos;
    const uint32_t freq = right ? (e.freq0 ^ e.freq1_xor_freq0) : e.freq0;
    state_ = freq * (state_ >> kAnsLogTabSize) + offset;
    if (state_ < (1u << 16)) state_ = (state_ << 16) | br.Read(16);
    return sym;
  }

  static uint32_t ReadHybrid(const HybridUintConfig& cfg, uint32_t token, BitReader& br) {
    if (token < cfg.split_token) return token;
    uint32_t in_token = cfg.msb_in_token + cfg.lsb_in_token;
    uint32_t nbits = (cfg.split_exponent - in_token + ((token - cfg.split_token) >> in_token)) & 31;
    uint32_t low = token & ((1u << cfg.lsb_in_token) - 1);
    token >>= cfg.lsb_in_token;
    uint32_t bits = br.Read(nbits);
    uint32_t hi = (1u << cfg.msb_in_token) | (token & ((1u << cfg.msb_in_token) - 1));
    return (((hi << nbits) | bits) << cfg.lsb_in_token) | low;
  }

  // `cluster` is already mapped through the context map.
  uint32_t ReadUintClustered(uint32_t cluster, BitReader& br) {
    if (!c_->lz77_enabled) return ReadHybrid(c_->cfg[cluster], ReadSymbol(cluster, br), br);
    if (num_to_copy_ > 0) return CopyOne();
    uint32_t token = ReadSymbol(cluster, br);
    if (token >= c_->lz77_min_symbol) {
      num_to_copy_ = ReadHybrid(c_->lz77_length_cfg, token - c_->lz77_min_symbol, br) +
                     c_->lz77_min_length;
      uint32_t dtok = ReadSymbol(c_->lz77_dist_cluster, br);
      uint64_t distance = ReadHybrid(c_->cfg[c_->lz77_dist_cluster], dtok, br);
      if (distance < num_special_) {
        distance = SpecialDistance(distance);
      } else {
        distance = distance + 1 - num_special_;
      }
      if (distance > num_decoded_) distance = num_decoded_;
      if (distance > kLZ77Window) distance = kLZ77Window;
      copy_pos_ = num_decoded_ - distance;
      if (distance == 0) {
        size_t fill = std::min<size_t>(num_to_copy_, kLZ77Window);
        std::fill(window_.begin(), window_.begin() + fill, 0);
      }
      if (num_to_copy_ < c_->lz77_min_length) return 0;  // overflow guard, dec_ans.h:330
      return CopyOne();
    }
    uint32_t v = ReadHybrid(c_->cfg[cluster], token, br);
    window_[(num_decoded_++) & (kLZ77Window - 1)] = v;
    return v;
  }

  uint32_t ReadUint(uint32_t ctx, BitReader& br) { return ReadUintClustered(c_->ctx_map[ctx], br); }

  bool FinalStateOk() const { return state_ == (kAnsSignature << 16); }

 private:
  uint32_t CopyOne() {
    uint32_t v = window_[(copy_pos_++) & (kLZ77Window - 1)];
    num_to_copy_--;
    window_[(num_decoded_++) & (kLZ77Window - 1)] = v;
    return v;
  }
  // lib/jxl/dec_ans.h:121-143
  uint64_t SpecialDistance(uint64_t i) const {
    static const int8_t k[120][2] = {
        {0, 1},  {1, 0},  {1, 1},  {-1, 1}, {0, 2},  {2, 0},  {1, 2},  {-1, 2}, {2, 1},  {-2, 1},
        {2, 2},  {-2, 2}, {0, 3},  {3, 0},  {1, 3},  {-1, 3}, {3, 1},  {-3, 1}, {2, 3},  {-2, 3},
        {3, 2},  {-3, 2}, {0, 4},  {4, 0},  {1, 4},  {-1, 4}, {4, 1},  {-4, 1}, {3, 3},  {-3, 3},
        {2, 4},  {-2, 4}, {4, 2},  {-4, 2}, {0, 5},  {3, 4},  {-3, 4}, {4, 3},  {-4, 3}, {5, 0},
        {1, 5},  {-1, 5}, {5, 1},  {-5, 1}, {2, 5},  {-2, 5}, {5, 2},  {-5, 2}, {4, 4},  {-4, 4},
        {3, 5},  {-3, 5}, {5, 3},  {-5, 3}, {0, 6},  {6, 0},  {1, 6},  {-1, 6}, {6, 1},  {-6, 1},
        {2, 6},  {-2, 6}, {6, 2},  {-6, 2}, {4, 5},  {-4, 5}, {5, 4},  {-5, 4}, {3, 6},  {-3, 6},
        {6, 3},  {-6, 3}, {0, 7},  {7, 0},  {1, 7},  {-1, 7}, {5, 5},  {-5, 5}, {7, 1},  {-7, 1},
        {4, 6},  {-4, 6}, {6, 4},  {-6, 4}, {2, 7},  {-2, 7}, {7, 2},  {-7, 2}, {3, 7},  {-3, 7},
        {7, 3},  {-7, 3}, {5, 6},  {-5, 6}, {6, 5},  {-6, 5}, {8, 0},  {4, 7},  {-4, 7}, {7, 4},
        {-7, 4}, {8, 1},  {8, 2},  {6, 6},  {-6, 6}, {8, 3},  {5, 7},  {-5, 7}, {7, 5},  {-7, 5},
        {8, 4},  {6, 7},  {-6, 7}, {7, 6},  {-7, 6}, {8, 5},  {7, 7},  {-7, 7}, {8, 6},  {8, 7}};
    int d = k[i][0] + static_cast<int>(dist_mult_) * k[i][1];
    return d > 1 ? d : 1;
  }

  const EntropyCode* c_;
  uint32_t state_;
  std::vector<uint32_t> window_;
  uint32_t num_to_copy_ = 0;
  uint64_t copy_pos_ = 0, num_decoded_ = 0;
  uint32_t num_special_ = 0, dist_mult_ = 0;
};

// lib/jxl/dec_context_map.cc:46-96
inline void ReadContextMap(BitReader& br, std::vector<uint8_t>* map, uint32_t* num_clusters) {
  if (br.Read(1)) {  // simple
    uint32_t bits = br.Read(2);
    for (auto& m : *map) m = bits ? br.Read(bits) : 0;
  } else {
    bool use_mtf = br.Read(1);
    EntropyCode nested;
    ReadEntropyCode(br, 1, &nested, /*disallow_lz77=*/map->size() <= 2);
    SymbolReader reader(&nested, br);
    uint32_t maxsym = 0;
    for (auto& m : *map) {
      uint32_t s = reader.ReadUint(0, br);
      maxsym = std::max(maxsym, s);
      m = static_cast<uint8_t>(s);
    }
    JXLB_CHECK(maxsym < 256, "context map: cluster id too large");
    JXLB_CHECK(reader.FinalStateOk(), "context map: bad ANS final state");
    if (use_mtf) {
      uint8_t mtf[256];
      for (int i = 0; i < 256; i++) mtf[i] = static_cast<uint8_t>(i);
      for (auto& m : *map) {
        uint8_t idx = m;
        uint8_t v = mtf[idx];
        m = v;
        for (; idx; idx--) mtf[idx] = mtf[idx - 1];
        mtf[0] = v;
      }
    }
  }
  uint32_t n = *std::max_element(map->begin(), map->end()) + 1;
  std::vector<bool> seen(n, false);
  for (uint8_t m : *map) seen[m] = true;
  for (bool s : seen) JXLB_CHECK(s, "context map: unused cluster");
  *num_clusters = n;
}

// lib/jxl/dec_ans.cc:336-368
inline void ReadEntropyCode(BitReader& br, size_t num_contexts, EntropyCode* code,
                            bool disallow_lz77) {
  code->lz77_enabled = br.Read(1);
  if (code->lz77_enabled) {
    code->lz77_min_symbol = ReadU32(br, Val(224), Val(512), Val(4096), BitsOffset(15, 8));
    code->lz77_min_length = ReadU32(br, Val(3), Val(4), BitsOffset(2, 5), BitsOffset(8, 9));
    num_contexts++;
    code->lz77_length_cfg = ReadUintConfig(br, 8);
  }
  JXLB_CHECK(!(code->lz77_enabled && disallow_lz77), "LZ77 not allowed here");
  code->ctx_map.assign(num_contexts, 0);
  code->num_clusters = 1;
  if (num_contexts > 1) ReadContextMap(br, &code->ctx_map, &code->num_clusters);
  code->lz77_dist_cluster = code->ctx_map.back();
  code->use_prefix = br.Read(1);
  code->log_alpha_size = code->use_prefix ? kPrefixMaxBits : br.Read(2) + 5;
  code->cfg.resize(code->num_clusters);
  for (auto& c : code->cfg) c = ReadUintConfig(br, code->log_alpha_size);
  code->degenerate.assign(code->num_clusters, -1);
  code->max_num_bits = 0;
  if (code->use_prefix) {
    code->prefix.resize(code->num_clusters);
    std::vector<uint32_t> alphabet(code->num_clusters);
    for (auto& a : alphabet) a = ReadVarLenUint16(br) + 1;
    for (uint32_t c = 0; c < code->num_clusters; c++) {
      ReadPrefixCode(br, alphabet[c], &code->prefix[c]);
      const auto& t = code->prefix[c].t;
      for (uint32_t i = 0; i < t.size(); i++)
        if (!(t[i] & kPrefixLink)) UpdateMaxNumBits(code, c, t[i] & 0xFFFF);
    }
  } else {
    const uint32_t ts = 1u << code->log_alpha_size;
    code->alias.resize(static_cast<size_t>(code->num_clusters) * ts);
    for (uint32_t c = 0; c < code->num_clusters; c++) {
      std::vector<int32_t> counts = ReadAnsHistogram(br);
      JXLB_CHECK(counts.size() <= ts, "histogram: alphabet too large");
      while (!counts.empty() && counts.back() == 0) counts.pop_back();
      for (size_t s = 0; s < counts.size(); s++)
        if (counts[s]) UpdateMaxNumBits(code, c, s);
      int deg = counts.empty() ? 0 : static_cast<int>(counts.size()) - 1;
      for (int s = 0; s < deg; s++) {
        if (counts[s]) {
          deg = -1;
          break;
        }
      }
      code->degenerate[c] = deg;
      BuildAliasTable(counts, code->log_alpha_size, &code->alias[static_cast<size_t>(c) * ts]);
    }
  }
  br.CheckInBounds();
}

}  // namespace jxlb

#endif  // JXLB_ENTROPY_H_
