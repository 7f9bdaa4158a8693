// TEST INFRASTRUCTURE: libjxl's self-contained known-answer methods restated against BOTH host
// parsers -- the oracle's (jxlo::) and the product's (jxlb::, csrc/host/) -- because the two share most of their text:
//   lib/jxl/fields_test.cc, lib/jxl/bit_reader_test.cc   random U32 / U64 / F16 / Enum / raw-bit round trips
//   lib/jxl/ans_test.cc:27-170, :202-298                random token streams through ANS, prefix codes and LZ77, final state
//   lib/jxl/coeff_order_test.cc, lib/jxl/toc_test.cc     Lehmer-coded permutations, (permuted) section tables
// The streams are written by the oracle's writers (oracle/jxlo_encode.h) from a seeded generator and must read back
// value for value. Each function returns 0 or 1 + the index of the first mismatch; -1 on an exception.
#include <cstring>
#include <random>

#include "../../jpegxl-rs_b200/csrc/host/jxlb_headers.h"
#include "../../oracle/jxlo_decode.h"
#include "../../oracle/jxlo_encode.h"

namespace {

// ---- fields
struct FieldOp {
  int kind;  // 0 raw bits, 1 U32, 2 U64, 3 F16, 4 Enum, 5 bool, 6 align to byte
  uint32_t nbits;
  uint64_t value;
  int dist;
};

std::vector<FieldOp> MakeFieldOps(uint32_t seed, size_t n) {
  std::mt19937_64 rng(seed);
  std::vector<FieldOp> ops;
  for (size_t i = 0; i < n; i++) {
    FieldOp op{};
    op.kind = static_cast<int>(rng() % 7);
    switch (op.kind) {
      case 0:
        op.nbits = 1 + rng() % 32;
        op.value = rng() & ((uint64_t{1} << op.nbits) - 1);
        break;
      case 1: {
        op.dist = static_cast<int>(rng() % 3);
        const uint32_t sel = rng() % 4;
        if (op.dist == 0) {  // Bits(10), BitsOffset(14, 1024), BitsOffset(22, 17408), BitsOffset(30, 4211712): TOC sizes
          static const uint32_t nb[4] = {10, 14, 22, 30}, off[4] = {0, 1024, 17408, 4211712};
          op.value = off[sel] + (rng() & ((1u << nb[sel]) - 1));
        } else if (op.dist == 1) {  // Val(0), Val(1), BitsOffset(4, 2), BitsOffset(8, 18)
          op.value = sel == 0 ? 0 : sel == 1 ? 1 : sel == 2 ? 2 + rng() % 16 : 18 + rng() % 256;
        } else {  // BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1): image sizes
          static const uint32_t nb[4] = {9, 13, 18, 30};
          op.value = 1 + (rng() & ((1u << nb[sel]) - 1));
        }
        break;
      }
      case 2: {
        const uint32_t sel = rng() % 5;
        op.value = sel == 0 ? 0 : sel == 1 ? 1 + rng() % 16 : sel == 2 ? 17 + rng() % 256 : sel == 3 ? rng() >> (rng() % 60) : rng();
        break;
      }
      case 3: {
        // a finite half-precision value: sign, exponent 0 .. 30, mantissa
        const uint32_t h = ((rng() & 1) << 15) | ((rng() % 31) << 10) | (rng() & 0x3FF);
        op.value = h;
        break;
      }
      case 4:
        op.value = rng() % 4 == 0 ? 18 + rng() % 64 : rng() % 18;
        break;
      case 5:
        op.value = rng() & 1;
        break;
      default:
        break;
    }
    ops.push_back(op);
  }
  return ops;
}

float HalfBitsToFloat(uint32_t h) {
  const uint32_t sign = h >> 15, exp = (h >> 10) & 31, mant = h & 0x3FF;
  float v;
  if (exp == 0) {
    v = std::ldexp(static_cast<float>(mant), -24);
  } else {
    v = std::ldexp(static_cast<float>(mant | 0x400), static_cast<int>(exp) - 25);
  }
  return sign ? -v : v;
}

template <class NS_BitReader, class Ops>
long ReadFieldOps(NS_BitReader& br, const std::vector<FieldOp>& ops, const Ops& o) {
  for (size_t i = 0; i < ops.size(); i++) {
    const FieldOp& op = ops[i];
    bool ok = true;
    switch (op.kind) {
      case 0: ok = br.Read(op.nbits) == op.value; break;
      case 1: ok = o.U32(br, op.dist) == op.value; break;
      case 2: ok = o.U64(br) == op.value; break;
      case 3: ok = o.F16(br) == HalfBitsToFloat(static_cast<uint32_t>(op.value)); break;
      case 4: ok = o.Enum(br) == op.value; break;
      case 5: ok = br.ReadBool() == (op.value != 0); break;
      default: br.AlignToByte(); break;
    }
    if (!ok) return static_cast<long>(i) + 1;
  }
  return 0;
}

struct OracleOps {
  uint32_t U32(jxlo::BitReader& br, int dist) const {
    using namespace jxlo;
    if (dist == 0) return ReadU32(br, Bits(10), BitsOffset(14, 1024), BitsOffset(22, 17408), BitsOffset(30, 4211712));
    if (dist == 1) return ReadU32(br, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(8, 18));
    return ReadU32(br, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  }
  uint64_t U64(jxlo::BitReader& br) const { return jxlo::ReadU64(br); }
  float F16(jxlo::BitReader& br) const { return jxlo::ReadF16(br); }
  uint32_t Enum(jxlo::BitReader& br) const { return jxlo::ReadEnum(br); }
};
struct ProductOps {
  uint32_t U32(jxlb::BitReader& br, int dist) const {
    using namespace jxlb;
    if (dist == 0) return ReadU32(br, Bits(10), BitsOffset(14, 1024), BitsOffset(22, 17408), BitsOffset(30, 4211712));
    if (dist == 1) return ReadU32(br, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(8, 18));
    return ReadU32(br, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  }
  uint64_t U64(jxlb::BitReader& br) const { return jxlb::ReadU64(br); }
  float F16(jxlb::BitReader& br) const { return jxlb::ReadF16(br); }
  uint32_t Enum(jxlb::BitReader& br) const { return jxlb::ReadEnum(br); }
};

// ---- token streams
std::vector<jxlo::Token> MakeTokens(uint32_t seed, uint32_t num_ctx, size_t n, uint32_t value_bits, bool runs) {
  std::mt19937 rng(seed);
  std::vector<jxlo::Token> t;
  while (t.size() < n) {
    const uint32_t ctx = rng() % num_ctx;
    // skewed magnitudes: mostly small, sometimes up to value_bits bits
    const uint32_t bits = rng() % 4 == 0 ? 1 + rng() % value_bits : 1 + rng() % 4;
    const uint32_t v = rng() & ((bits >= 32 ? 0u : (1u << bits)) - 1u);
    size_t rep = 1;
    if (runs && rng() % 8 == 0) rep = 3 + rng() % 40;
    for (size_t k = 0; k < rep && t.size() < n; k++) t.push_back({runs && rng() % 2 ? ctx : static_cast<uint32_t>(rng() % num_ctx), v});
  }
  if (runs && n > 400) {  // matches at the distances a Modular stream sees (the row above and around it)
    for (size_t k = 0; k < 60; k++) t[n - 80 + k].value = t[n - 80 + k - 97].value;
    for (size_t k = 0; k < 40; k++) t[n - 200 + k].value = t[n - 200 + k - 99].value;
  }
  return t;
}

template <class Code, class Reader, class BR>
long DecodeTokens(BR& br, size_t num_ctx, const std::vector<jxlo::Token>& toks, uint32_t dist_mult) {
  Code code;
  ReadEntropyCode(br, num_ctx, &code);
  Reader reader(&code, br, dist_mult);
  for (size_t i = 0; i < toks.size(); i++)
    if (reader.ReadUint(toks[i].ctx, br) != toks[i].value) return static_cast<long>(i) + 1;
  if (!reader.FinalStateOk()) return static_cast<long>(toks.size()) + 1;
  br.CheckInBounds();
  return 0;
}

}  // namespace

extern "C" {

// which: 0 = the oracle's readers, 1 = the product's host readers.
long jxlb_kat_fields(uint32_t seed, size_t n, int which) {
  try {
    const std::vector<FieldOp> ops = MakeFieldOps(seed, n);
    jxlo::BitWriter w;
    for (const FieldOp& op : ops) {
      using namespace jxlo;
      switch (op.kind) {
        case 0: w.Write(op.nbits, op.value); break;
        case 1:
          if (op.dist == 0) WriteU32(w, op.value, Bits(10), BitsOffset(14, 1024), BitsOffset(22, 17408), BitsOffset(30, 4211712));
          else if (op.dist == 1) WriteU32(w, op.value, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(8, 18));
          else WriteU32(w, op.value, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
          break;
        case 2: WriteU64(w, op.value); break;
        case 3: w.Write(16, op.value); break;
        case 4: WriteU32(w, op.value, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(6, 18)); break;
        case 5: w.Write(1, op.value); break;
        default: w.ZeroPadToByte(); break;
      }
    }
    w.ZeroPadToByte();
    const std::vector<uint8_t>& b = w.Bytes();
    if (which == 0) {
      jxlo::BitReader br(b.data(), b.size());
      const long r = ReadFieldOps(br, ops, OracleOps());
      br.CheckInBounds();
      return r;
    }
    jxlb::BitReader br(b.data(), b.size());
    const long r = ReadFieldOps(br, ops, ProductOps());
    br.CheckInBounds();
    return r;
  } catch (const std::exception&) {
    return -1;
  }
}

// mode: bit 0 prefix codes, bit 1 LZ77. Clusters: ctx % num_clusters.
long jxlb_kat_entropy(uint32_t seed, uint32_t num_ctx, uint32_t num_clusters, size_t n, int mode, uint32_t dist_mult,
                      uint32_t value_bits, int which) {
  try {
    const std::vector<jxlo::Token> toks = MakeTokens(seed, num_ctx, n, value_bits, (mode & 2) != 0);
    std::vector<uint8_t> clusters(num_ctx);
    for (uint32_t c = 0; c < num_ctx; c++) clusters[c] = static_cast<uint8_t>(c % num_clusters);
    jxlo::EntropyOptions opt;
    opt.use_prefix = (mode & 1) != 0;
    opt.lz77 = (mode & 2) != 0;
    if (opt.use_prefix && opt.lz77 && (seed & 1)) opt.lz77_min_symbol = 512;
    jxlo::EntropyEncoder enc(num_ctx, clusters, opt);
    enc.Count(toks, dist_mult);
    jxlo::BitWriter w;
    enc.WriteHeader(w);
    enc.WriteTokens(w, toks, dist_mult);
    w.ZeroPadToByte();
    const std::vector<uint8_t>& b = w.Bytes();
    if (which == 0) {
      jxlo::BitReader br(b.data(), b.size());
      return DecodeTokens<jxlo::EntropyCode, jxlo::SymbolReader>(br, num_ctx, toks, dist_mult);
    }
    jxlb::BitReader br(b.data(), b.size());
    return DecodeTokens<jxlb::EntropyCode, jxlb::SymbolReader>(br, num_ctx, toks, dist_mult);
  } catch (const std::exception&) {
    return -1;
  }
}

// A random permutation of `size` elements (the first `skip` fixed), Lehmer-coded like EncodeCoeffOrders
// (lib/jxl/enc_coeff_order.cc:293-337) / EncodePermutation (lib/jxl/enc_toc.cc), read back through ReadPermutationStream.
static std::vector<jxlo::Token> PermutationTokens(const std::vector<uint32_t>& perm, size_t skip) {
  const size_t size = perm.size();
  std::vector<uint32_t> lehmer(size, 0), avail(size);
  for (size_t i = 0; i < size; i++) avail[i] = i;
  for (size_t i = 0; i < size; i++) {
    const auto it = std::lower_bound(avail.begin(), avail.end(), perm[i]);
    lehmer[i] = static_cast<uint32_t>(it - avail.begin());
    avail.erase(it);
  }
  size_t end = size;
  while (end > skip && lehmer[end - 1] == 0) end--;
  std::vector<jxlo::Token> toks;
  toks.push_back({jxlo::CoeffOrderContext(size), static_cast<uint32_t>(end - skip)});
  uint32_t last = 0;
  for (size_t i = skip; i < end; i++) {
    toks.push_back({jxlo::CoeffOrderContext(last), lehmer[i]});
    last = lehmer[i];
  }
  return toks;
}

static std::vector<uint32_t> RandomPermutation(uint32_t seed, size_t size, size_t skip) {
  std::mt19937 rng(seed);
  std::vector<uint32_t> perm(size);
  for (size_t i = 0; i < size; i++) perm[i] = i;
  // mostly local swaps (what coefficient orders look like) + a few far ones; the tail stays in place
  const size_t active = skip + (size - skip) * (1 + rng() % 4) / 4;
  for (size_t i = skip; i + 1 < active; i++) {
    const size_t span = rng() % 16 == 0 ? active - i : std::min<size_t>(active - i, 1 + rng() % 6);
    std::swap(perm[i], perm[i + rng() % span]);
  }
  return perm;
}

long jxlb_kat_permutation(uint32_t seed, size_t size, size_t skip, int which) {
  try {
    const std::vector<uint32_t> perm = RandomPermutation(seed, size, skip);
    const std::vector<jxlo::Token> toks = PermutationTokens(perm, skip);
    jxlo::EntropyEncoder enc(8, {0, 1, 2, 3, 4, 5, 6, 7});
    enc.Count(toks);
    jxlo::BitWriter w;
    enc.WriteHeader(w);
    enc.WriteTokens(w, toks);
    w.ZeroPadToByte();
    const std::vector<uint8_t>& b = w.Bytes();
    std::vector<uint32_t> got(size, 0);
    if (which == 0) {
      jxlo::BitReader br(b.data(), b.size());
      jxlo::ReadPermutationStream(br, skip, size, got.data());
      br.CheckInBounds();
    } else {
      jxlb::BitReader br(b.data(), b.size());
      jxlb::ReadPermutationStream(br, skip, size, got.data());
      br.CheckInBounds();
    }
    for (size_t i = 0; i < size; i++)
      if (got[i] != perm[i]) return static_cast<long>(i) + 1;
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

// Section table with random sizes over all four U32 ranges, optionally permuted (lib/jxl/toc_test.cc, lib/jxl/enc_toc.cc).
long jxlb_kat_toc(uint32_t seed, size_t entries, int permuted, int which) {
  try {
    std::mt19937 rng(seed);
    std::vector<uint32_t> sizes(entries);
    for (auto& s : sizes) {
      static const uint32_t nb[4] = {10, 14, 22, 30}, off[4] = {0, 1024, 17408, 4211712};
      const uint32_t sel = rng() % 16 == 0 ? 3 : rng() % 3;
      s = off[sel] + (rng() & ((1u << std::min<uint32_t>(nb[sel], 24)) - 1));
    }
    std::vector<uint32_t> perm;
    jxlo::BitWriter w;
    if (permuted) {
      perm = RandomPermutation(seed + 1, entries, 0);
      w.Write(1, 1);
      const std::vector<jxlo::Token> toks = PermutationTokens(perm, 0);
      jxlo::EntropyEncoder enc(8, {0, 1, 2, 3, 4, 5, 6, 7});
      enc.Count(toks);
      enc.WriteHeader(w);
      enc.WriteTokens(w, toks);
    } else {
      w.Write(1, 0);
    }
    w.ZeroPadToByte();
    using namespace jxlo;
    for (uint32_t s : sizes) WriteU32(w, s, Bits(10), BitsOffset(14, 1024), BitsOffset(22, 17408), BitsOffset(30, 4211712));
    w.ZeroPadToByte();
    const std::vector<uint8_t>& b = w.Bytes();
    // expected: logical section j lives at bitstream slot perm[j]
    std::vector<uint64_t> pre(entries, 0);
    uint64_t off = 0;
    for (size_t i = 0; i < entries; i++) {
      pre[i] = off;
      off += sizes[i];
    }
    auto check = [&](const auto& toc) -> long {
      if (toc.total != off) return static_cast<long>(entries) + 1;
      for (size_t j = 0; j < entries; j++) {
        const size_t slot = perm.empty() ? j : perm[j];
        if (toc.offsets[j] != pre[slot] || toc.logical_size[j] != sizes[slot]) return static_cast<long>(j) + 1;
      }
      return 0;
    };
    if (which == 0) {
      jxlo::BitReader br(b.data(), b.size());
      return check(jxlo::ReadToc(br, entries));
    }
    jxlb::BitReader br(b.data(), b.size());
    return check(jxlb::ReadToc(br, entries));
  } catch (const std::exception&) {
    return -1;
  }
}

// Bit reader: reads of every width at every alignment, peeks, skips, position accounting and the zero-fill behind
// the end (lib/jxl/bit_reader_test.cc). Returns 0 when both properties hold for reader `which`.
long jxlb_kat_bit_reader(uint32_t seed, size_t nbytes, int which) {
  try {
    std::mt19937 rng(seed);
    std::vector<uint8_t> bytes(nbytes);
    for (auto& v : bytes) v = static_cast<uint8_t>(rng());
    auto bit = [&](size_t i) -> uint64_t { return i / 8 < nbytes ? (bytes[i / 8] >> (i % 8)) & 1u : 0u; };
    auto run = [&](auto& br) -> long {
      size_t pos = 0;
      for (size_t step = 0; pos + 64 < nbytes * 8; step++) {
        const uint32_t n = rng() % 33;
        uint64_t want = 0;
        for (uint32_t k = 0; k < n; k++) want |= bit(pos + k) << k;
        if (rng() % 3 == 0) {
          if (n && br.Peek(n) != want) return static_cast<long>(step) + 1;
          br.Skip(n);
        } else if (br.Read(n) != want) {
          return static_cast<long>(step) + 1;
        }
        pos += n;
        if (br.BitPos() != pos) return static_cast<long>(step) + 1;
      }
      return 0;
    };
    if (which == 0) {
      jxlo::BitReader br(bytes.data(), bytes.size());
      return run(br);
    }
    jxlb::BitReader br(bytes.data(), bytes.size());
    return run(br);
  } catch (const std::exception&) {
    return -1;
  }
}

}  // extern "C"
