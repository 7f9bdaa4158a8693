// jxl_b200 host planner, VarDCT frames: everything of a lossy frame that is O(KB) and
// serial -- quantiser, block context map, chroma-from-luma base, dequantisation tables,
// coefficient orders, per-pass entropy codes -- is parsed here; the per-sample loops (DC and
// AC-metadata Modular streams, AC coefficients, dequantisation, inverse transforms, loop
// filters, colour) run in the CUDA kernels. Follows
//   lib/jxl/dec_frame.cc:61-77, :266-339, :367-476 (section contents),
//   lib/jxl/quantizer.{h,cc}, lib/jxl/quant_weights.cc:42-355, :367-520,
//   lib/jxl/base/fast_math-inl.h:46-90, lib/jxl/entropy_coder.cc:25-60, lib/jxl/ac_context.h,
//   lib/jxl/chroma_from_luma.{h,cc}, lib/jxl/coeff_order.{h,cc}, lib/jxl/ac_strategy.cc:24-82,
//   lib/jxl/dec_modular.cc:397-532 (which Modular streams a DC group section holds),
//   lib/jxl/dec_xyb.cc:192-330 (output encoding), lib/jxl/loop_filter.cc.
#ifndef JXLB_VARDCT_PLAN_H_
#define JXLB_VARDCT_PLAN_H_

#include <cmath>
#include <mutex>

#include "../kernels/jxlb_vardct_desc.h"
#include "jxlb_plan.h"

namespace jxlb {

#include "jxlb_tables.inc"

// ---------------------------------------------------------------- fast math (host side of the tables)
inline float HEvalRational2(float x, const float p[3], const float q[3]) {
  float yp = p[2], yq = q[2];
  yp = std::fmaf(yp, x, p[1]);
  yq = std::fmaf(yq, x, q[1]);
  yp = std::fmaf(yp, x, p[0]);
  yq = std::fmaf(yq, x, q[0]);
  return yp / yq;
}

inline float HFastLog2f(float x) {
  static const float p[3] = {-1.8503833400518310E-06f, 1.4287160470083755E+00f, 7.4245873327820566E-01f};
  static const float q[3] = {9.9032814277590719E-01f, 1.0096718572241148E+00f, 1.7409343003366853E-01f};
  int32_t x_bits;
  std::memcpy(&x_bits, &x, 4);
  const int32_t exp_bits = x_bits - 0x3f2aaaab;
  const int32_t exp_shifted = exp_bits >> 23;
  const int32_t mant_bits = x_bits - static_cast<int32_t>(static_cast<uint32_t>(exp_shifted) << 23);
  float mantissa;
  std::memcpy(&mantissa, &mant_bits, 4);
  return HEvalRational2(mantissa - 1.0f, p, q) + static_cast<float>(exp_shifted);
}

inline float HFastPow2f(float x) {
  const float floorx = std::floor(x);
  const int32_t e = static_cast<int32_t>(static_cast<uint32_t>(static_cast<int32_t>(floorx) + 127) << 23);
  float exp;
  std::memcpy(&exp, &e, 4);
  const float frac = x - floorx;
  float num = frac + 1.01749063e+01f;
  num = std::fmaf(num, frac, 4.88687798e+01f);
  num = std::fmaf(num, frac, 9.85506591e+01f);
  num = num * exp;
  float den = std::fmaf(frac, 2.10242958e-01f, -2.22328856e-02f);
  den = std::fmaf(den, frac, -1.94414990e+01f);
  den = std::fmaf(den, frac, 9.85506633e+01f);
  return num / den;
}

inline float HFastPowf(float base, float exponent) { return HFastPow2f(HFastLog2f(base) * exponent); }

// ---------------------------------------------------------------- quantisation tables
constexpr int kNumQuantKinds = 17;
static const int kQuantSizeX[17] = {1, 1, 1, 1, 2, 4, 1, 1, 2, 1, 1, 8, 4, 16, 8, 32, 16};
static const int kQuantSizeY[17] = {1, 1, 1, 1, 2, 4, 2, 4, 4, 1, 1, 8, 8, 16, 16, 32, 32};
enum QuantMode { kQLib = 0, kQID = 1, kQDCT2 = 2, kQDCT4 = 3, kQDCT4X8 = 4, kQAFV = 5, kQDCT = 6, kQRAW = 7 };
constexpr float kTinyWeight = 1e-8f;
constexpr float kSqrt2H = 1.41421356237f;

struct BandParams {
  int num_bands = 0;
  float bands[3][17] = {};
};

struct QuantTableSpec {
  int mode = kQLib;
  BandParams dct, dct4x4;
  float idweights[3][3] = {}, dct2weights[3][6] = {}, dct4mul[3][2] = {}, dct4x8mul[3] = {}, afv[3][9] = {};
  float qtable_den = 0;       // raw mode (lib/jxl/quant_weights.cc:339-346): weight = 1 / (qtable_den * qtable[i])
  std::vector<int32_t> qtable;
};

inline QuantTableSpec LibrarySpec(int kind) {
  const QuantLibEntry& e = kQuantLibrary[kind];
  QuantTableSpec q;
  q.mode = e.mode;
  q.dct.num_bands = e.num_bands;
  std::memcpy(q.dct.bands, e.bands, sizeof(e.bands));
  q.dct4x4.num_bands = e.num_bands2;
  std::memcpy(q.dct4x4.bands, e.bands2, sizeof(e.bands2));
  for (int c = 0; c < 3; c++) {
    for (int i = 0; i < 3; i++) q.idweights[c][i] = e.extra[c][i];
    for (int i = 0; i < 6; i++) q.dct2weights[c][i] = e.extra[c][i];
    for (int i = 0; i < 2; i++) q.dct4mul[c][i] = e.extra[c][i];
    q.dct4x8mul[c] = e.extra[c][0];
    for (int i = 0; i < 9; i++) q.afv[c][i] = e.extra[c][i];
  }
  return q;
}

inline void ReadBandParams(BitReader& br, BandParams* p) {
  p->num_bands = br.Read(4) + 1;
  for (int c = 0; c < 3; c++) {
    for (int i = 0; i < p->num_bands; i++) p->bands[c][i] = ReadF16(br);
    JXLB_CHECK(p->bands[c][0] >= kTinyWeight, "distance band seed too small");
    p->bands[c][0] *= 64.0f;
  }
}

inline float BandMult(float v) { return v > 0.0f ? 1.0f + v : 1.0f / (1.0f - v); }

// Radial band interpolation of the weights of a rows x cols DCT (GetQuantWeights).
inline void BandWeights(size_t rows, size_t cols, const BandParams& p, float* out) {
  const size_t nb = p.num_bands;
  for (size_t c = 0; c < 3; c++) {
    float bands[17] = {p.bands[c][0]};
    JXLB_CHECK(bands[0] >= kTinyWeight, "invalid distance bands");
    for (size_t i = 1; i < nb; i++) {
      bands[i] = bands[i - 1] * BandMult(p.bands[c][i]);
      JXLB_CHECK(bands[i] >= kTinyWeight, "invalid distance bands");
    }
    const float scale = (nb - 1) / (kSqrt2H + 1e-6f);
    const float rcpcol = scale / (cols - 1), rcprow = scale / (rows - 1);
    for (uint32_t y = 0; y < rows; y++) {
      const float dy = y * rcprow, dy2 = dy * dy;
      for (uint32_t x = 0; x < cols; x++) {
        const float dx = (static_cast<float>(x & ~3u) + static_cast<float>(x & 3u)) * rcpcol;
        const float dist = std::sqrt(std::fmaf(dx, dx, dy2));
        float weight = bands[0];
        if (nb > 1) {
          const int32_t idx = static_cast<int32_t>(dist);
          const float frac = dist - static_cast<float>(idx);
          const float a = bands[idx], b = bands[idx + 1];
          weight = a * HFastPowf(b / a, frac);
        }
        out[c * cols * rows + y * cols + x] = weight;
      }
    }
  }
}

// Dequantisation multipliers (1 / weight) of table kind `kind`: 3 channels x 64 * sx * sy.
inline std::vector<float> BuildDequantTable(const QuantTableSpec& q, int kind) {
  const size_t wrows = 8 * kQuantSizeX[kind], wcols = 8 * kQuantSizeY[kind], num = wrows * wcols;
  std::vector<float> w(3 * num, 0.0f);
  auto need64 = [&]() { JXLB_CHECK(num == 64, "quant mode does not fit this table"); };
  switch (q.mode) {
    case kQID:
      need64();
      for (size_t c = 0; c < 3; c++) {
        for (int i = 0; i < 64; i++) w[64 * c + i] = q.idweights[c][0];
        w[64 * c + 1] = w[64 * c + 8] = q.idweights[c][1];
        w[64 * c + 9] = q.idweights[c][2];
      }
      break;
    case kQDCT2:
      need64();
      for (size_t c = 0; c < 3; c++) {
        float* t = &w[c * 64];
        const float* v = q.dct2weights[c];
        t[0] = 0xBAD;
        t[1] = t[8] = v[0];
        t[9] = v[1];
        for (int y = 0; y < 8; y++)
          for (int x = 0; x < 8; x++) {
            if (x < 2 && y < 2) continue;
            const int m = std::max(x, y);
            const bool diag = std::min(x, y) >= (m < 4 ? 2 : 4);
            t[y * 8 + x] = m < 4 ? (diag ? v[3] : v[2]) : (diag ? v[5] : v[4]);
          }
      }
      break;
    case kQDCT4: {
      need64();
      float w4[3 * 16];
      BandWeights(4, 4, q.dct, w4);
      for (size_t c = 0; c < 3; c++) {
        for (size_t y = 0; y < 8; y++)
          for (size_t x = 0; x < 8; x++) w[c * 64 + y * 8 + x] = w4[c * 16 + (y / 2) * 4 + (x / 2)];
        w[c * 64 + 1] /= q.dct4mul[c][0];
        w[c * 64 + 8] /= q.dct4mul[c][0];
        w[c * 64 + 9] /= q.dct4mul[c][1];
      }
      break;
    }
    case kQDCT4X8: {
      need64();
      float w48[3 * 32];
      BandWeights(4, 8, q.dct, w48);
      for (size_t c = 0; c < 3; c++) {
        for (size_t y = 0; y < 8; y++)
          for (size_t x = 0; x < 8; x++) w[c * 64 + y * 8 + x] = w48[c * 32 + (y / 2) * 8 + x];
        w[c * 64 + 8] /= q.dct4x8mul[c];
      }
      break;
    }
    case kQDCT:
      BandWeights(wrows, wcols, q.dct, w.data());
      break;
    case kQAFV: {
      need64();
      static const float kFreqs[16] = {0xBAD, 0xBAD, 0.8517778890324296, 5.37778436506804, 0xBAD, 0xBAD,
                                       4.734747904497923, 5.449245381693219, 1.6598270267479331, 4, 7.275749096817861,
                                       10.423227632456525, 2.662932286148962, 7.630657783650829, 8.962388608184032,
                                       12.97166202570235};
      float w48[3 * 32], w44[3 * 16];
      BandWeights(4, 8, q.dct, w48);
      BandWeights(4, 4, q.dct4x4, w44);
      constexpr float lo = 0.8517778890324296;
      constexpr float hi = 12.97166202570235f - lo + 1e-6f;
      for (size_t c = 0; c < 3; c++) {
        float bands[4] = {q.afv[c][5]};
        JXLB_CHECK(bands[0] >= kTinyWeight, "invalid AFV bands");
        for (size_t i = 1; i < 4; i++) {
          bands[i] = bands[i - 1] * BandMult(q.afv[c][i + 5]);
          JXLB_CHECK(bands[i] >= kTinyWeight, "invalid AFV bands");
        }
        float* t = &w[c * 64];
        t[0] = 1;
        t[1 * 8 + 0] = q.afv[c][0];
        t[0 * 8 + 1] = q.afv[c][1];
        t[2 * 8 + 0] = q.afv[c][2];
        t[0 * 8 + 2] = q.afv[c][3];
        t[2 * 8 + 2] = q.afv[c][4];
        for (size_t y = 0; y < 4; y++)
          for (size_t x = 0; x < 4; x++) {
            if (x < 2 && y < 2) continue;
            const float pos = (kFreqs[y * 4 + x] - lo) * (4 - 1) / hi;
            const size_t idx = static_cast<size_t>(pos);
            JXLB_CHECK(idx + 1 < 4, "AFV interpolation out of range");
            const float a = bands[idx], b = bands[idx + 1];
            t[(2 * y) * 8 + 2 * x] = a * HFastPowf(b / a, pos - idx);
          }
        for (size_t y = 0; y < 4; y++)
          for (size_t x = 0; x < 8; x++)
            if (x || y) t[(2 * y + 1) * 8 + x] = w48[c * 32 + y * 8 + x];
        for (size_t y = 0; y < 4; y++)
          for (size_t x = 0; x < 4; x++)
            if (x || y) t[(2 * y) * 8 + 2 * x + 1] = w44[c * 16 + y * 4 + x];
      }
      break;
    }
    case kQRAW:
      JXLB_CHECK(q.qtable.size() == 3 * num, "invalid raw quantisation table");
      for (size_t i = 0; i < 3 * num; i++) w[i] = 1.0f / (q.qtable_den * q.qtable[i]);
      break;
    default:
      throw Error("invalid quantisation table mode");
  }
  std::vector<float> out(3 * num);
  for (size_t i = 0; i < 3 * num; i++) {
    JXLB_CHECK(!(w[i] >= 1.0f / kTinyWeight) && !(w[i] < kTinyWeight), "invalid quantisation table");
    out[i] = 1.0f / w[i];
  }
  return out;
}

// `read_raw(br, kind, q)` fills q->qtable_den / q->qtable for the raw mode: the table is a Modular-coded image
// (lib/jxl/dec_modular.cc:765-812) whose samples come back from the device (ProbeCtx).
template <typename ReadRaw>
inline void ReadQuantTableSpec(BitReader& br, int kind, QuantTableSpec* q, const ReadRaw& read_raw) {
  const int blocks = kQuantSizeX[kind] * kQuantSizeY[kind];
  const int mode = br.Read(3);
  auto small_only = [&]() { JXLB_CHECK(blocks == 1, "invalid quant mode"); };
  auto nonzero = [&](float v) { JXLB_CHECK(std::fabs(v) >= kTinyWeight, "quantiser parameter too small"); };
  switch (mode) {
    case kQLib:
      break;
    case kQID:
      small_only();
      for (int c = 0; c < 3; c++)
        for (int i = 0; i < 3; i++) {
          q->idweights[c][i] = ReadF16(br);
          nonzero(q->idweights[c][i]);
          q->idweights[c][i] *= 64;
        }
      break;
    case kQDCT2:
      small_only();
      for (int c = 0; c < 3; c++)
        for (int i = 0; i < 6; i++) {
          q->dct2weights[c][i] = ReadF16(br);
          nonzero(q->dct2weights[c][i]);
          q->dct2weights[c][i] *= 64;
        }
      break;
    case kQDCT4X8:
      small_only();
      for (int c = 0; c < 3; c++) {
        q->dct4x8mul[c] = ReadF16(br);
        nonzero(q->dct4x8mul[c]);
      }
      ReadBandParams(br, &q->dct);
      break;
    case kQDCT4:
      small_only();
      for (int c = 0; c < 3; c++)
        for (int i = 0; i < 2; i++) {
          q->dct4mul[c][i] = ReadF16(br);
          nonzero(q->dct4mul[c][i]);
        }
      ReadBandParams(br, &q->dct);
      break;
    case kQAFV:
      small_only();
      for (int c = 0; c < 3; c++) {
        for (int i = 0; i < 9; i++) q->afv[c][i] = ReadF16(br);
        for (int i = 0; i < 6; i++) q->afv[c][i] *= 64;
      }
      ReadBandParams(br, &q->dct);
      ReadBandParams(br, &q->dct4x4);
      break;
    case kQDCT:
      ReadBandParams(br, &q->dct);
      break;
    default:
      read_raw(br, kind, q);
      break;
  }
  q->mode = mode;
}

// ---------------------------------------------------------------- coefficient orders
// Zig-zag over the square of side max(cx, cy) * 8 keeping every (max / min)-th row, with the
// lowest cx * cy frequencies first.
inline void NaturalOrder(uint32_t strategy, std::vector<uint16_t>* out) {
  const StrategyInfo si = GetStrategyInfo(strategy);
  size_t cx = si.cx, cy = si.cy;
  if (cy > cx) std::swap(cx, cy);
  out->assign(cx * cy * 64, 0);
  const size_t ratio = cx / cy, mask = ratio - 1, shift = CeilLog2(ratio);
  const size_t side = cx * 8;
  size_t cur = cx * cy;
  auto visit = [&](size_t x, size_t y, bool llf_possible) {
    if ((y & mask) != 0) return;
    y >>= shift;
    size_t slot;
    if (llf_possible && x < cx && y < cy) {
      slot = y * cx + x;
    } else {
      slot = cur++;
    }
    (*out)[slot] = static_cast<uint16_t>(y * side + x);
  };
  for (size_t d = 0; d < side; d++)  // upper-left triangle, diagonal d
    for (size_t j = 0; j <= d; j++) {
      size_t x = j, y = d - j;
      if (d % 2) std::swap(x, y);
      visit(x, y, true);
    }
  for (size_t d = side - 1; d-- > 0;)  // lower-right triangle
    for (size_t j = 0; j <= d; j++) {
      size_t x = side - 1 - (d - j), y = side - 1 - j;
      if (d % 2) std::swap(x, y);
      visit(x, y, false);
    }
}

// First strategy of every coefficient-order class, and the size of its order in blocks.
static const uint8_t kOrderFirstStrategy[13] = {0, 1, 4, 5, 6, 8, 10, 18, 19, 21, 22, 24, 25};

// Tables that do not depend on the frame: the 17 library dequantisation tables and the 13
// natural coefficient orders. Built once per process, uploaded at the head of every batch's
// fpool / opool.
static const uint16_t kCoeffFreqContext[64] = {
    0xBAD, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 15, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 21, 21, 22, 22,
    23, 23, 23, 23, 24, 24, 24, 24, 25, 25, 25, 25, 26, 26, 26, 26, 27, 27, 27, 27, 28, 28, 28, 28, 29, 29, 29, 29, 30, 30, 30, 30};
static const uint16_t kCoeffNumNonzeroContext[64] = {
    0xBAD, 0, 31, 62, 62, 93, 93, 93, 93, 123, 123, 123, 123, 152, 152, 152, 152, 152, 152, 152, 152, 180, 180, 180, 180, 180,
    180, 180, 180, 180, 180, 180, 180, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206,
    206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206};

struct SharedVarDCTTables {
  std::vector<float> fpool;
  uint32_t table_off[17];
  uint32_t wc_off = 0, llf_off = 0, afv_off = 0;
  std::vector<uint16_t> opool;
  uint32_t order_off[13];
  std::vector<uint32_t> upool;  // packed StrategyInfo x 27, kCoeffFreqContext, kCoeffNumNonzeroContext
  uint32_t sinfo_off = 0, ctxtab_off = 27;
  static const SharedVarDCTTables& Get() {
    static SharedVarDCTTables* t = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
      t = new SharedVarDCTTables();
      for (int k = 0; k < kNumQuantKinds; k++) {
        t->table_off[k] = t->fpool.size();
        std::vector<float> tab = BuildDequantTable(LibrarySpec(k), k);
        t->fpool.insert(t->fpool.end(), tab.begin(), tab.end());
      }
      t->wc_off = t->fpool.size();
      t->fpool.insert(t->fpool.end(), kWcMultipliers, kWcMultipliers + 254);
      t->llf_off = t->fpool.size();
      t->fpool.insert(t->fpool.end(), kResampleToLLF, kResampleToLLF + 63);
      t->afv_off = t->fpool.size();
      for (int j = 0; j < 16; j++) t->fpool.insert(t->fpool.end(), kAFVBasis[j], kAFVBasis[j] + 16);
      t->fpool.resize((t->fpool.size() + 3) & ~size_t{3}, 0.0f);
      for (uint32_t st = 0; st < kNumStrategies; st++) t->upool.push_back(PackStrategyInfo(GetStrategyInfo(st)));
      for (int i = 0; i < 64; i++) t->upool.push_back(kCoeffFreqContext[i]);
      for (int i = 0; i < 64; i++) t->upool.push_back(kCoeffNumNonzeroContext[i]);
      for (int o = 0; o < 13; o++) {
        t->order_off[o] = t->opool.size();
        std::vector<uint16_t> ord;
        NaturalOrder(kOrderFirstStrategy[o], &ord);
        t->opool.insert(t->opool.end(), ord.begin(), ord.end());
      }
    });
    return *t;
  }
};

constexpr uint32_t kSharedFlag = 0x80000000u;  // table_off / order offsets that refer to the shared tables

// ---------------------------------------------------------------- block context map
static const uint8_t kDefaultBlockCtx[39] = {0, 1, 2, 2, 3, 3, 4, 5, 6, 6, 6, 6, 6, 7, 8, 9, 9, 10, 11, 12,
                                             13, 14, 14, 14, 14, 14, 7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14};

struct HostBlockCtx {
  std::vector<int32_t> dc_thr[3];
  std::vector<uint32_t> qf_thr;
  std::vector<uint8_t> ctx_map;
  uint32_t num_ctxs = 15, num_dc_ctxs = 1;
  HostBlockCtx() { ctx_map.assign(kDefaultBlockCtx, kDefaultBlockCtx + 39); }
  uint32_t NumACContexts() const { return num_ctxs * (37 + 458); }
};

inline void ReadBlockCtx(BitReader& br, HostBlockCtx* m) {
  if (br.ReadBool()) return;  // default
  m->num_dc_ctxs = 1;
  for (int j = 0; j < 3; j++) {
    m->dc_thr[j].resize(br.Read(4));
    m->num_dc_ctxs *= m->dc_thr[j].size() + 1;
    for (int32_t& t : m->dc_thr[j]) t = UnpackSigned(ReadU32(br, Bits(4), BitsOffset(8, 16), BitsOffset(16, 272), BitsOffset(32, 65808)));
  }
  m->qf_thr.resize(br.Read(4));
  for (uint32_t& t : m->qf_thr) t = ReadU32(br, Bits(2), BitsOffset(3, 4), BitsOffset(5, 12), BitsOffset(8, 44)) + 1;
  JXLB_CHECK(m->num_dc_ctxs * (m->qf_thr.size() + 1) <= 64, "block context map too big");
  m->ctx_map.assign(3 * kNumOrders * m->num_dc_ctxs * (m->qf_thr.size() + 1), 0);
  uint32_t n = 1;
  ReadContextMap(br, &m->ctx_map, &n);
  m->num_ctxs = n;
  JXLB_CHECK(m->num_ctxs <= 16, "too many block contexts");
}

// ---------------------------------------------------------------- output colour
inline bool CanOutputToEncoding(const ColorEncoding& c) {
  if (c.want_icc) return false;
  if (!c.have_gamma && c.transfer_function != kTFPQ && c.transfer_function != kTFSRGB && c.transfer_function != kTFLinear &&
      c.transfer_function != kTFHLG && c.transfer_function != kTFDCI && c.transfer_function != kTF709)
    return false;
  if (c.IsGray() && c.white_point != 1) return false;
  return true;
}

inline void FillOutputColor(const ImageMetadata& meta, DevVFrame* vf) {
  ColorEncoding c = meta.color;
  if (meta.xyb_encoded && !CanOutputToEncoding(c)) {
    const bool grey = c.IsGray();
    c = ColorEncoding();
    c.color_space = grey ? kGray : kRGB;
    c.transfer_function = kTFLinear;
  }
  JXLB_CHECK(c.IsGray() || (c.primaries == 1 && c.white_point == 1) || !meta.xyb_encoded,
             "unsupported: XYB output to non-sRGB primaries / white point");
  for (int i = 0; i < 3; i++) {
    vf->opsin_bias[i] = meta.opsin_biases[i];
    vf->opsin_bias_cbrt[i] = cbrtf(meta.opsin_biases[i]);
  }
  float inv[9];
  for (int i = 0; i < 9; i++) inv[i] = meta.inverse_opsin[i];
  if (c.IsGray()) {
    const float lum[3] = {0.2126, 0.7152, 0.0722};
    float tmp[9];
    for (int x = 0; x < 3; x++) {
      const double t[3] = {inv[0 * 3 + x], inv[1 * 3 + x], inv[2 * 3 + x]};
      for (int y = 0; y < 3; y++) tmp[y * 3 + x] = static_cast<float>(lum[0] * t[0] + lum[1] * t[1] + lum[2] * t[2]);
    }
    std::memcpy(inv, tmp, sizeof(inv));
  }
  for (int i = 0; i < 9; i++) vf->inv_mat[i] = inv[i] * (255.0f / meta.intensity_target);
  if (c.have_gamma || c.transfer_function == kTFDCI) {
    vf->tf = 2;
    vf->inv_gamma = c.have_gamma ? static_cast<float>(c.gamma * (1.0 / 10000000)) : 1.0f / 2.6f;
  } else if (c.transfer_function == kTFLinear) {
    vf->tf = 0;
  } else if (c.transfer_function == kTFSRGB) {
    vf->tf = 1;
  } else {
    throw Error("unsupported: PQ / HLG / BT.709 transfer function");
  }
}

}  // namespace jxlb

#endif  // JXLB_VARDCT_PLAN_H_
