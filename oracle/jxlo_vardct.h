// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
//
// VarDCT frames: global parameters, DC + side-information sub-streams, AC entropy
// decoding, dequantisation, chroma-from-luma and the inverse transforms. Restates
//   lib/jxl/dec_frame.cc:61-77, :266-355, :367-476 (section contents),
//   lib/jxl/quantizer.{h,cc} (global scale, DC steps), lib/jxl/quantizer-inl.h:34-71,
//   lib/jxl/quant_weights.cc:42-355 (table synthesis), :367-520 (table decoding),
//   lib/jxl/base/fast_math-inl.h:46-90 (FastLog2f / FastPow2f / FastPowf),
//   lib/jxl/entropy_coder.cc:25-60 (block context map), lib/jxl/ac_context.h,
//   lib/jxl/chroma_from_luma.{h,cc}, lib/jxl/compressed_dc.cc:124-290,
//   lib/jxl/dec_modular.cc:397-532 (DC and AC-metadata streams), lib/jxl/epf.cc:39-147,
//   lib/jxl/coeff_order.{h,cc}, lib/jxl/ac_strategy.cc:24-82,
//   lib/jxl/dec_group.cc:98-166, :168-442, :454-527, :534-645.
//
// Where libjxl's result depends on the SIMD target, the oracle follows the x86 FMA
// builds with two documented exceptions: AdjustQuantBias uses an exact reciprocal
// (the reference calls the 12-bit rcpps approximation, lib/jxl/quantizer-inl.h:66-68),
// and per-vector decisions (EPF's sigma test) are taken per pixel.
#ifndef JXLO_VARDCT_H_
#define JXLO_VARDCT_H_

#include <array>
#include <functional>

#include "jxlo_dct.h"
#include "jxlo_frame.h"

namespace jxlo {

// ---------------------------------------------------------------- fast math
inline float EvalRational2(float x, const float p[3], const float q[3]) {
  float yp = p[2], yq = q[2];
  yp = std::fmaf(yp, x, p[1]);
  yq = std::fmaf(yq, x, q[1]);
  yp = std::fmaf(yp, x, p[0]);
  yq = std::fmaf(yq, x, q[0]);
  return yp / yq;
}

inline float FastLog2f(float x) {
  static const float p[3] = {-1.8503833400518310E-06f, 1.4287160470083755E+00f, 7.4245873327820566E-01f};
  static const float q[3] = {9.9032814277590719E-01f, 1.0096718572241148E+00f, 1.7409343003366853E-01f};
  int32_t x_bits;
  std::memcpy(&x_bits, &x, 4);
  const int32_t exp_bits = x_bits - 0x3f2aaaab;
  const int32_t exp_shifted = exp_bits >> 23;
  const int32_t mant_bits = x_bits - static_cast<int32_t>(static_cast<uint32_t>(exp_shifted) << 23);
  float mantissa;
  std::memcpy(&mantissa, &mant_bits, 4);
  const float exp_val = static_cast<float>(exp_shifted);
  return EvalRational2(mantissa - 1.0f, p, q) + exp_val;
}

inline float FastPow2f(float x) {
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

inline float FastPowf(float base, float exponent) { return FastPow2f(FastLog2f(base) * exponent); }

// ---------------------------------------------------------------- quantisation tables
constexpr int kNumQuantTablesI = 17;
static const int kRequiredSizeX[17] = {1, 1, 1, 1, 2, 4, 1, 1, 2, 1, 1, 8, 4, 16, 8, 32, 16};
static const int kRequiredSizeY[17] = {1, 1, 1, 1, 2, 4, 2, 4, 4, 1, 1, 8, 8, 16, 16, 32, 32};
// lib/jxl/quant_weights.h:343-353
static const uint8_t kStrategyToQuantTable[27] = {0, 1, 2, 3, 4, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10, 10,
                                                  11, 12, 12, 13, 14, 14, 15, 16, 16};
enum QuantMode { kQuantLib = 0, kQuantID = 1, kQuantDCT2 = 2, kQuantDCT4 = 3, kQuantDCT4X8 = 4, kQuantAFV = 5,
                 kQuantDCT = 6, kQuantRAW = 7 };

struct DctParams {
  int num_bands = 0;
  float bands[3][17] = {};
};

struct QuantEncoding {
  int mode = kQuantLib;
  DctParams dct, dct4x4;          // dct: main (or 4x8 for AFV); dct4x4: AFV only
  float idweights[3][3] = {};
  float dct2weights[3][6] = {};
  float dct4mul[3][2] = {};
  float dct4x8mul[3] = {};
  float afv[3][9] = {};
  float qtable_den = 0;
  std::vector<int> qtable;
};

inline QuantEncoding LibraryEncoding(int table) {
  const QuantLibEntry& e = kQuantLibrary[table];
  QuantEncoding q;
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

constexpr float kAlmostZero = 1e-8f;

inline void ReadDctParams(BitReader& br, DctParams* p) {  // quant_weights.cc:367-381
  p->num_bands = br.Read(4) + 1;
  for (int c = 0; c < 3; c++) {
    for (int i = 0; i < p->num_bands; i++) p->bands[c][i] = ReadF16(br);
    JXLO_CHECK(p->bands[c][0] >= kAlmostZero, "distance band seed too small");
    p->bands[c][0] *= 64.0f;
  }
}

inline float QuantMult(float v) { return v > 0.0f ? 1.0f + v : 1.0f / (1.0f - v); }

// GetQuantWeights, quant_weights.cc:127-160
inline void GetQuantWeights(size_t rows, size_t cols, const DctParams& p, float* out) {
  const size_t num_bands = p.num_bands;
  for (size_t c = 0; c < 3; c++) {
    float bands[17] = {p.bands[c][0]};
    JXLO_CHECK(bands[0] >= kAlmostZero, "invalid distance bands");
    for (size_t i = 1; i < num_bands; i++) {
      bands[i] = bands[i - 1] * QuantMult(p.bands[c][i]);
      JXLO_CHECK(bands[i] >= kAlmostZero, "invalid distance bands");
    }
    const float scale = (num_bands - 1) / (kSqrt2f + 1e-6f);
    const float rcpcol = scale / (cols - 1);
    const float rcprow = scale / (rows - 1);
    for (uint32_t y = 0; y < rows; y++) {
      const float dy = y * rcprow;
      const float dy2 = dy * dy;
      for (uint32_t x = 0; x < cols; x++) {
        const float dx = (static_cast<float>(x & ~3u) + static_cast<float>(x & 3u)) * rcpcol;
        const float scaled_distance = std::sqrt(std::fmaf(dx, dx, dy2));
        float weight;
        if (num_bands == 1) {
          weight = bands[0];
        } else {
          const int32_t idx = static_cast<int32_t>(scaled_distance);
          const float frac = scaled_distance - static_cast<float>(idx);
          const float a = bands[idx], b = bands[idx + 1];
          weight = a * FastPowf(b / a, frac);
        }
        out[c * cols * rows + y * cols + x] = weight;
      }
    }
  }
}

// ComputeQuantTable, quant_weights.cc:162-355. Returns the dequantisation
// multipliers (1 / weight), 3 * num values, channel-major.
// The encoder-side weights of a quantisation table (what libjxl keeps as InvMatrix before it zeroes the LLF corner,
// lib/jxl/quant_weights.cc:322-349); the dequantisation table is 1 / weights (ComputeQuantTable below).
inline std::vector<float> ComputeQuantWeights(const QuantEncoding& enc, int table) {
  const size_t wrows = 8 * kRequiredSizeX[table], wcols = 8 * kRequiredSizeY[table];
  const size_t num = wrows * wcols;
  std::vector<float> weights(3 * num, 0.0f);
  switch (enc.mode) {
    case kQuantID:
      JXLO_CHECK(num == 64, "bad quant mode for table");
      for (size_t c = 0; c < 3; c++) {
        for (int i = 0; i < 64; i++) weights[64 * c + i] = enc.idweights[c][0];
        weights[64 * c + 1] = enc.idweights[c][1];
        weights[64 * c + 8] = enc.idweights[c][1];
        weights[64 * c + 9] = enc.idweights[c][2];
      }
      break;
    case kQuantDCT2:
      JXLO_CHECK(num == 64, "bad quant mode for table");
      for (size_t c = 0; c < 3; c++) {
        const size_t start = c * 64;
        const float* w = enc.dct2weights[c];
        weights[start] = 0xBAD;
        weights[start + 1] = weights[start + 8] = w[0];
        weights[start + 9] = w[1];
        for (size_t y = 0; y < 2; y++)
          for (size_t x = 0; x < 2; x++) {
            weights[start + y * 8 + x + 2] = w[2];
            weights[start + (y + 2) * 8 + x] = w[2];
          }
        for (size_t y = 0; y < 2; y++)
          for (size_t x = 0; x < 2; x++) weights[start + (y + 2) * 8 + x + 2] = w[3];
        for (size_t y = 0; y < 4; y++)
          for (size_t x = 0; x < 4; x++) {
            weights[start + y * 8 + x + 4] = w[4];
            weights[start + (y + 4) * 8 + x] = w[4];
          }
        for (size_t y = 0; y < 4; y++)
          for (size_t x = 0; x < 4; x++) weights[start + (y + 4) * 8 + x + 4] = w[5];
      }
      break;
    case kQuantDCT4: {
      JXLO_CHECK(num == 64, "bad quant mode for table");
      float w4[3 * 16];
      GetQuantWeights(4, 4, enc.dct, w4);
      for (size_t c = 0; c < 3; c++) {
        for (size_t y = 0; y < 8; y++)
          for (size_t x = 0; x < 8; x++) weights[c * num + y * 8 + x] = w4[c * 16 + (y / 2) * 4 + (x / 2)];
        weights[c * num + 1] /= enc.dct4mul[c][0];
        weights[c * num + 8] /= enc.dct4mul[c][0];
        weights[c * num + 9] /= enc.dct4mul[c][1];
      }
      break;
    }
    case kQuantDCT4X8: {
      JXLO_CHECK(num == 64, "bad quant mode for table");
      float w48[3 * 32];
      GetQuantWeights(4, 8, enc.dct, w48);
      for (size_t c = 0; c < 3; c++) {
        for (size_t y = 0; y < 8; y++)
          for (size_t x = 0; x < 8; x++) weights[c * num + y * 8 + x] = w48[c * 32 + (y / 2) * 8 + x];
        weights[c * num + 8] /= enc.dct4x8mul[c];
      }
      break;
    }
    case kQuantDCT:
      GetQuantWeights(wrows, wcols, enc.dct, weights.data());
      break;
    case kQuantRAW:
      JXLO_CHECK(enc.qtable.size() == 3 * num, "invalid raw quant table");
      for (size_t i = 0; i < 3 * num; i++) weights[i] = 1.f / (enc.qtable_den * enc.qtable[i]);
      break;
    case kQuantAFV: {
      JXLO_CHECK(num == 64, "bad quant mode for table");
      static const float kFreqs[16] = {
          0xBAD, 0xBAD, 0.8517778890324296, 5.37778436506804, 0xBAD, 0xBAD, 4.734747904497923, 5.449245381693219,
          1.6598270267479331, 4, 7.275749096817861, 10.423227632456525, 2.662932286148962, 7.630657783650829,
          8.962388608184032, 12.97166202570235};
      float w48[3 * 32], w44[3 * 16];
      GetQuantWeights(4, 8, enc.dct, w48);
      GetQuantWeights(4, 4, enc.dct4x4, w44);
      constexpr float lo = 0.8517778890324296;
      constexpr float hi = 12.97166202570235f - lo + 1e-6f;
      for (size_t c = 0; c < 3; c++) {
        float bands[4];
        bands[0] = enc.afv[c][5];
        JXLO_CHECK(bands[0] >= kAlmostZero, "invalid AFV bands");
        for (size_t i = 1; i < 4; i++) {
          bands[i] = bands[i - 1] * QuantMult(enc.afv[c][i + 5]);
          JXLO_CHECK(bands[i] >= kAlmostZero, "invalid AFV bands");
        }
        const size_t start = c * 64;
        auto set_weight = [&](size_t x, size_t y, float val) { weights[start + y * 8 + x] = val; };
        weights[start] = 1;
        set_weight(0, 1, enc.afv[c][0]);
        set_weight(1, 0, enc.afv[c][1]);
        set_weight(0, 2, enc.afv[c][2]);
        set_weight(2, 0, enc.afv[c][3]);
        set_weight(2, 2, enc.afv[c][4]);
        for (size_t y = 0; y < 4; y++)
          for (size_t x = 0; x < 4; x++) {
            if (x < 2 && y < 2) continue;
            // Interpolate(), quant_weights.cc:86-94
            const float pos = kFreqs[y * 4 + x] - lo;
            const float scaled_pos = pos * (4 - 1) / hi;
            const size_t idx = static_cast<size_t>(scaled_pos);
            JXLO_CHECK(idx + 1 < 4, "AFV interpolation out of range");
            const float a = bands[idx], b = bands[idx + 1];
            set_weight(2 * x, 2 * y, a * FastPowf(b / a, scaled_pos - idx));
          }
        for (size_t y = 0; y < 4; y++)
          for (size_t x = 0; x < 8; x++) {
            if (x == 0 && y == 0) continue;
            weights[c * num + (2 * y + 1) * 8 + x] = w48[c * 32 + y * 8 + x];
          }
        for (size_t y = 0; y < 4; y++)
          for (size_t x = 0; x < 4; x++) {
            if (x == 0 && y == 0) continue;
            weights[c * num + (2 * y) * 8 + 2 * x + 1] = w44[c * 16 + y * 4 + x];
          }
      }
      break;
    }
    default:
      throw Error("jxlo: unresolved quant table mode");
  }
  for (size_t i = 0; i < 3 * num; i++)
    JXLO_CHECK(!(weights[i] >= 1.0f / kAlmostZero) && !(weights[i] < kAlmostZero), "invalid quantization table");
  return weights;
}

inline std::vector<float> ComputeQuantTable(const QuantEncoding& enc, int table) {
  std::vector<float> out_table = ComputeQuantWeights(enc, table);
  for (float& v : out_table) v = 1.0f / v;
  return out_table;
}

inline void ReadQuantEncoding(BitReader& br, int table, QuantEncoding* q,
                              const std::function<void(BitReader&, int, int, int, std::vector<int>*)>& read_raw) {
  const int required = kRequiredSizeX[table] * kRequiredSizeY[table];
  const int mode = br.Read(3);
  switch (mode) {
    case kQuantLib:
      break;  // kCeilLog2NumPredefinedTables = 0 bits
    case kQuantID:
      JXLO_CHECK(required == 1, "invalid quant mode");
      for (int c = 0; c < 3; c++)
        for (int i = 0; i < 3; i++) {
          q->idweights[c][i] = ReadF16(br);
          JXLO_CHECK(std::fabs(q->idweights[c][i]) >= kAlmostZero, "ID quantizer too small");
          q->idweights[c][i] *= 64;
        }
      break;
    case kQuantDCT2:
      JXLO_CHECK(required == 1, "invalid quant mode");
      for (int c = 0; c < 3; c++)
        for (int i = 0; i < 6; i++) {
          q->dct2weights[c][i] = ReadF16(br);
          JXLO_CHECK(std::fabs(q->dct2weights[c][i]) >= kAlmostZero, "quantizer too small");
          q->dct2weights[c][i] *= 64;
        }
      break;
    case kQuantDCT4X8:
      JXLO_CHECK(required == 1, "invalid quant mode");
      for (int c = 0; c < 3; c++) {
        q->dct4x8mul[c] = ReadF16(br);
        JXLO_CHECK(std::fabs(q->dct4x8mul[c]) >= kAlmostZero, "DCT4X8 multiplier too small");
      }
      ReadDctParams(br, &q->dct);
      break;
    case kQuantDCT4:
      JXLO_CHECK(required == 1, "invalid quant mode");
      for (int c = 0; c < 3; c++)
        for (int i = 0; i < 2; i++) {
          q->dct4mul[c][i] = ReadF16(br);
          JXLO_CHECK(std::fabs(q->dct4mul[c][i]) >= kAlmostZero, "DCT4 multiplier too small");
        }
      ReadDctParams(br, &q->dct);
      break;
    case kQuantAFV:
      JXLO_CHECK(required == 1, "invalid quant mode");
      for (int c = 0; c < 3; c++) {
        for (int i = 0; i < 9; i++) q->afv[c][i] = ReadF16(br);
        for (int i = 0; i < 6; i++) q->afv[c][i] *= 64;
      }
      ReadDctParams(br, &q->dct);
      ReadDctParams(br, &q->dct4x4);
      break;
    case kQuantDCT:
      ReadDctParams(br, &q->dct);
      break;
    case kQuantRAW: {
      q->qtable_den = ReadF16(br);
      JXLO_CHECK(q->qtable_den >= kAlmostZero, "invalid qtable_den");
      read_raw(br, table, 8 * kRequiredSizeX[table], 8 * kRequiredSizeY[table], &q->qtable);
      break;
    }
  }
  q->mode = mode;
}

// ---------------------------------------------------------------- block context map
static const uint8_t kDefaultBlockCtxMap[39] = {0, 1, 2, 2, 3, 3, 4, 5, 6, 6, 6, 6, 6, 7, 8, 9, 9, 10, 11, 12,
                                                13, 14, 14, 14, 14, 14, 7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14};
constexpr uint32_t kNumOrders = 13;
constexpr uint32_t kNonZeroBuckets = 37;
constexpr uint32_t kZeroDensityContextCount = 458;
constexpr uint32_t kZeroDensityContextLimit = 474;
static const uint8_t kStrategyOrder[27] = {0, 1, 1, 1, 2, 3, 4, 4, 5, 5, 6, 6, 1, 1, 1, 1, 1, 1, 7, 8, 8, 9, 10, 10, 11, 12, 12};
static const uint16_t kCoeffOrderOffset[40] = {0, 1, 2, 3, 4, 5, 6, 10, 14, 18, 34, 50, 66, 68, 70, 72, 76, 80, 84, 92,
                                               100, 108, 172, 236, 300, 332, 364, 396, 652, 908, 1164, 1292, 1420, 1548,
                                               2572, 3596, 4620, 5132, 5644, 6156};
static const uint16_t kCoeffFreqContext[64] = {
    0xBAD, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 15, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 21, 21, 22, 22,
    23, 23, 23, 23, 24, 24, 24, 24, 25, 25, 25, 25, 26, 26, 26, 26, 27, 27, 27, 27, 28, 28, 28, 28, 29, 29, 29, 29, 30, 30, 30, 30};
static const uint16_t kCoeffNumNonzeroContext[64] = {
    0xBAD, 0, 31, 62, 62, 93, 93, 93, 93, 123, 123, 123, 123, 152, 152, 152, 152, 152, 152, 152, 152, 180, 180, 180, 180, 180,
    180, 180, 180, 180, 180, 180, 180, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206,
    206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206};

struct BlockCtxMap {
  std::vector<int> dc_thresholds[3];
  std::vector<uint32_t> qf_thresholds;
  std::vector<uint8_t> ctx_map;
  size_t num_ctxs = 0, num_dc_ctxs = 1;
  BlockCtxMap() {
    ctx_map.assign(kDefaultBlockCtxMap, kDefaultBlockCtxMap + 39);
    num_ctxs = 15;
  }
  size_t Context(int dc_idx, uint32_t qf, size_t ord, size_t c) const {
    size_t qf_idx = 0;
    for (uint32_t t : qf_thresholds)
      if (qf > t) qf_idx++;
    size_t idx = c < 2 ? c ^ 1 : 2;
    idx = idx * kNumOrders + ord;
    idx = idx * (qf_thresholds.size() + 1) + qf_idx;
    idx = idx * num_dc_ctxs + dc_idx;
    return ctx_map[idx];
  }
  uint32_t ZeroDensityContextsOffset(uint32_t block_ctx) const {
    return static_cast<uint32_t>(num_ctxs * kNonZeroBuckets + kZeroDensityContextCount * block_ctx);
  }
  uint32_t NumACContexts() const { return static_cast<uint32_t>(num_ctxs * (kNonZeroBuckets + kZeroDensityContextCount)); }
  uint32_t NonZeroContext(uint32_t non_zeros, uint32_t block_ctx) const {
    if (non_zeros >= 64) non_zeros = 64;
    const uint32_t ctx = non_zeros < 8 ? non_zeros : 4 + non_zeros / 2;
    return static_cast<uint32_t>(ctx * num_ctxs + block_ctx);
  }
};

inline void ReadBlockCtxMap(BitReader& br, BlockCtxMap* m) {  // entropy_coder.cc:25-60
  if (br.Read(1)) {
    *m = BlockCtxMap();
    return;
  }
  m->num_dc_ctxs = 1;
  for (int j = 0; j < 3; j++) {
    m->dc_thresholds[j].resize(br.Read(4));
    m->num_dc_ctxs *= m->dc_thresholds[j].size() + 1;
    for (int& t : m->dc_thresholds[j])
      t = UnpackSigned(ReadU32(br, Bits(4), BitsOffset(8, 16), BitsOffset(16, 272), BitsOffset(32, 65808)));
  }
  m->qf_thresholds.resize(br.Read(4));
  for (uint32_t& t : m->qf_thresholds) t = ReadU32(br, Bits(2), BitsOffset(3, 4), BitsOffset(5, 12), BitsOffset(8, 44)) + 1;
  JXLO_CHECK(m->num_dc_ctxs * (m->qf_thresholds.size() + 1) <= 64, "block context map too big");
  m->ctx_map.assign(3 * kNumOrders * m->num_dc_ctxs * (m->qf_thresholds.size() + 1), 0);
  uint32_t n = 1;
  ReadContextMap(br, &m->ctx_map, &n);
  m->num_ctxs = n;
  JXLO_CHECK(m->num_ctxs <= 16, "too many block contexts");
}

// ---------------------------------------------------------------- natural coefficient order
// lib/jxl/ac_strategy.cc:24-82
inline void NaturalCoeffOrder(int strategy, uint32_t* out) {
  size_t cx = kCoveredX[strategy], cy = kCoveredY[strategy];
  if (cy > cx) std::swap(cx, cy);  // CoefficientLayout: cx >= cy
  const size_t xs = cx / cy, xsm = xs - 1, xss = CeilLog2(xs);
  size_t cur = cx * cy;
  for (size_t i = 0; i < cx * 8; i++) {
    for (size_t j = 0; j <= i; j++) {
      size_t x = j, y = i - j;
      if (i % 2) std::swap(x, y);
      if ((y & xsm) != 0) continue;
      y >>= xss;
      size_t val;
      if (x < cx && y < cy) {
        val = y * cx + x;
      } else {
        val = cur++;
      }
      out[val] = y * cx * 8 + x;
    }
  }
  for (size_t ip = cx * 8 - 1; ip > 0; ip--) {
    const size_t i = ip - 1;
    for (size_t j = 0; j <= i; j++) {
      size_t x = cx * 8 - 1 - (i - j), y = cx * 8 - 1 - j;
      if (i % 2) std::swap(x, y);
      if ((y & xsm) != 0) continue;
      y >>= xss;
      out[cur++] = y * cx * 8 + x;
    }
  }
}

// ---------------------------------------------------------------- state
struct PassCode {
  std::vector<uint32_t> orders;  // kCoeffOrderLimit * 64
  EntropyCode code;
};

struct VarDCTState {
  const FrameHeader& fh;
  FrameDimensions dim;
  const ImageMetadata& meta;
  float dc_quant[3];
  // Quantizer
  int global_scale = 1, quant_dc_q = 1;
  float global_scale_float = 0, inv_global_scale = 0, inv_quant_dc = 0, mul_dc[3] = {0, 0, 0};
  BlockCtxMap bctx;
  // ColorCorrelation
  uint32_t color_factor = 84;
  float color_scale = 1.0f / 84, base_x = 0.0f, base_b = 1.0f;
  int ytox_dc = 0, ytob_dc = 0;
  float dc_factors[3] = {0, 0, 0};
  size_t cmap_w = 0, cmap_h = 0;
  std::vector<int8_t> ytox_map, ytob_map;
  // per-block side information (xsize_blocks x ysize_blocks)
  std::vector<uint8_t> acs;       // (strategy << 1) | is_first, 0xFF = unset
  std::vector<uint8_t> epf_sharpness, quant_dc;
  std::vector<int32_t> raw_quant;
  std::vector<float> inv_sigma;   // 1 / sigma per block
  Plane dc[3];
  uint32_t used_acs = 0;
  // AC global
  QuantEncoding encodings[17];
  std::vector<float> tables[17];
  uint32_t num_histograms = 1;
  std::vector<PassCode> passes;
  float x_dm_multiplier = 1, b_dm_multiplier = 1;
  // accumulated coefficients per group (multi-pass) and output planes
  std::vector<std::vector<int32_t>> group_coeffs;
  Plane pix[3];

  VarDCTState(const FrameHeader& f, const FrameDimensions& d, const ImageMetadata& m) : fh(f), dim(d), meta(m) {
    const size_t nb = dim.xsize_blocks * dim.ysize_blocks;
    acs.assign(nb, 0xFF);
    epf_sharpness.assign(nb, 0);
    quant_dc.assign(nb, 0);
    raw_quant.assign(nb, 0);
    inv_sigma.assign(nb, 0.0f);
    cmap_w = DivCeil(dim.xsize_blocks, size_t{8});
    cmap_h = DivCeil(dim.ysize_blocks, size_t{8});
    ytox_map.assign(cmap_w * cmap_h, 0);
    ytob_map.assign(cmap_w * cmap_h, 0);
    for (int c = 0; c < 3; c++) {
      dc[c] = Plane(dim.xsize_blocks >> fh.HShift(c), dim.ysize_blocks >> fh.VShift(c));
      pix[c] = Plane((dim.xsize_blocks >> fh.HShift(c)) * 8, (dim.ysize_blocks >> fh.VShift(c)) * 8);
    }
    // lib/jxl/dec_cache.h:161-162
    x_dm_multiplier = std::pow(1 / (1.25f), f.x_qm_scale - 2.0f);
    b_dm_multiplier = std::pow(1 / (1.25f), f.b_qm_scale - 2.0f);
    group_coeffs.resize(dim.num_groups);
    for (auto& e : encodings) e = QuantEncoding();
  }
  float YtoXRatio(int f) const { return base_x + f * color_scale; }
  float YtoBRatio(int f) const { return base_b + f * color_scale; }
};

// DecodeGlobalDCInfo, lib/jxl/dec_frame.cc:61-77
inline void VarDCTReadGlobalDC(BitReader& br, VarDCTState* s) {
  s->global_scale = ReadU32(br, BitsOffset(11, 1), BitsOffset(11, 2049), BitsOffset(12, 4097), BitsOffset(16, 8193));
  s->quant_dc_q = ReadU32(br, Val(16), BitsOffset(5, 1), BitsOffset(8, 1), BitsOffset(16, 1));
  // RecomputeFromGlobalScale, lib/jxl/quantizer.h:82-90 (kGlobalScaleDenom = 1 << 16)
  s->global_scale_float = s->global_scale * (1.0 / 65536);
  s->inv_global_scale = 1.0 * 65536 / s->global_scale;
  s->inv_quant_dc = s->inv_global_scale / s->quant_dc_q;
  for (int c = 0; c < 3; c++) s->mul_dc[c] = s->inv_quant_dc * s->dc_quant[c];
  ReadBlockCtxMap(br, &s->bctx);
  // ColorCorrelation::DecodeDC, lib/jxl/chroma_from_luma.cc:20-41
  if (!br.Read(1)) {
    s->color_factor = ReadU32(br, Val(84), Val(256), BitsOffset(8, 2), BitsOffset(16, 258));
    s->color_scale = 1.0f / s->color_factor;
    s->base_x = ReadF16(br);
    JXLO_CHECK(std::fabs(s->base_x) <= 4.0f, "base X correlation out of range");
    s->base_b = ReadF16(br);
    JXLO_CHECK(std::fabs(s->base_b) <= 4.0f, "base B correlation out of range");
    s->ytox_dc = static_cast<int>(br.Read(8)) - 128;
    s->ytob_dc = static_cast<int>(br.Read(8)) - 128;
  }
  s->dc_factors[0] = s->YtoXRatio(s->ytox_dc);
  s->dc_factors[2] = s->YtoBRatio(s->ytob_dc);
}

// DC group rectangle in blocks.
struct BlockRect { size_t x0, y0, xs, ys; };
inline BlockRect DCGroupRect(const FrameDimensions& dim, size_t g) {
  const size_t gx = g % dim.xsize_dc_groups, gy = g / dim.xsize_dc_groups;
  BlockRect r{gx * dim.group_dim, gy * dim.group_dim, dim.group_dim, dim.group_dim};
  r.xs = std::min(r.xs, dim.xsize_blocks - r.x0);
  r.ys = std::min(r.ys, dim.ysize_blocks - r.y0);
  return r;
}
inline BlockRect BlockGroupRect(const FrameDimensions& dim, size_t g) {
  const size_t gx = g % dim.xsize_groups, gy = g / dim.xsize_groups;
  const size_t gd = dim.group_dim >> 3;
  BlockRect r{gx * gd, gy * gd, gd, gd};
  r.xs = std::min(r.xs, dim.xsize_blocks - r.x0);
  r.ys = std::min(r.ys, dim.ysize_blocks - r.y0);
  return r;
}

// DecodeVarDCTDC + DequantDC, lib/jxl/dec_modular.cc:397-435, lib/jxl/compressed_dc.cc:197-290
inline void VarDCTReadDCGroup(BitReader& br, VarDCTState* s, ModularFrameState* ms, size_t g) {
  const BlockRect r = DCGroupRect(s->dim, g);
  const FrameHeader& fh = s->fh;
  const uint32_t extra_precision = br.Read(2);
  const float mul = 1.0f / (1 << extra_precision);
  ModImage image;
  image.w = r.xs;
  image.h = r.ys;
  image.bitdepth = ms->full.bitdepth;
  for (int c = 0; c < 3; c++) image.ch.emplace_back(r.xs, r.ys);
  for (int c = 0; c < 3; c++) {
    Channel& ch = image.ch[c < 2 ? c ^ 1 : c];
    ch.Resize(ch.w >> fh.HShift(c), ch.h >> fh.VShift(c));
  }
  ModularOptions opt;
  ModularDecode(br, image, StreamVarDCTDC(s->dim, g), opt, ms->has_tree ? &ms->tree : nullptr,
                ms->has_tree ? &ms->code : nullptr);
  UndoTransforms(image, WPHeader());
  if (fh.Is444()) {
    const float fac_x = s->mul_dc[0] * mul, fac_y = s->mul_dc[1] * mul, fac_b = s->mul_dc[2] * mul;
    for (size_t y = 0; y < r.ys; y++) {
      const int32_t* qx = image.ch[1].Row(y);
      const int32_t* qy = image.ch[0].Row(y);
      const int32_t* qb = image.ch[2].Row(y);
      float* ox = s->dc[0].Row(r.y0 + y) + r.x0;
      float* oy = s->dc[1].Row(r.y0 + y) + r.x0;
      float* ob = s->dc[2].Row(r.y0 + y) + r.x0;
      for (size_t x = 0; x < r.xs; x++) {
        const float in_x = static_cast<float>(qx[x]) * fac_x;
        const float in_y = static_cast<float>(qy[x]) * fac_y;
        const float in_b = static_cast<float>(qb[x]) * fac_b;
        oy[x] = in_y;
        ox[x] = std::fmaf(in_y, s->dc_factors[0], in_x);
        ob[x] = std::fmaf(in_y, s->dc_factors[2], in_b);
      }
    }
  } else {
    for (int c : {1, 0, 2}) {
      const size_t x0 = r.x0 >> fh.HShift(c), y0 = r.y0 >> fh.VShift(c);
      const size_t xs = r.xs >> fh.HShift(c), ys = r.ys >> fh.VShift(c);
      const float fac = s->mul_dc[c] * mul;
      const Channel& ch = image.ch[c < 2 ? c ^ 1 : c];
      for (size_t y = 0; y < ys; y++)
        for (size_t x = 0; x < xs; x++) s->dc[c].Row(y0 + y)[x0 + x] = static_cast<float>(ch.Row(y)[x]) * fac;
    }
  }
  if (s->bctx.num_dc_ctxs > 1) {
    for (size_t y = 0; y < r.ys; y++) {
      const int32_t* qx = image.ch[1].Row(y >> fh.VShift(0));
      const int32_t* qy = image.ch[0].Row(y >> fh.VShift(1));
      const int32_t* qb = image.ch[2].Row(y >> fh.VShift(2));
      for (size_t x = 0; x < r.xs; x++) {
        int bx = 0, by = 0, bb = 0;
        for (int t : s->bctx.dc_thresholds[0])
          if (qx[x >> fh.HShift(0)] > t) bx++;
        for (int t : s->bctx.dc_thresholds[1])
          if (qy[x >> fh.HShift(1)] > t) by++;
        for (int t : s->bctx.dc_thresholds[2])
          if (qb[x >> fh.HShift(2)] > t) bb++;
        int bucket = bx;
        bucket *= s->bctx.dc_thresholds[2].size() + 1;
        bucket += bb;
        bucket *= s->bctx.dc_thresholds[1].size() + 1;
        bucket += by;
        s->quant_dc[(r.y0 + y) * s->dim.xsize_blocks + r.x0 + x] = bucket;
      }
    }
  }
}

// ComputeSigma, lib/jxl/epf.cc:39-147 (the mirrored one-block border of the reference's
// sigma image equals clamping the block coordinate, which RenderFrame does on lookup).
inline void ComputeSigma(VarDCTState* s, const BlockRect& r) {
  const LoopFilter& lf = s->fh.lf;
  constexpr float kInvSigmaNum = -1.1715728752538099024f;
  const size_t W = s->dim.xsize_blocks;
  for (size_t by = 0; by < r.ys; by++) {
    for (size_t bx = 0; bx < r.xs; bx++) {
      const size_t pos = (r.y0 + by) * W + r.x0 + bx;
      const uint8_t a = s->acs[pos];
      if (!(a & 1)) continue;
      const int strategy = a >> 1;
      const float sigma_quant = lf.epf_quant_mul / (s->global_scale_float * s->raw_quant[pos] * kInvSigmaNum);
      for (size_t iy = 0; iy < kCoveredY[strategy]; iy++)
        for (size_t ix = 0; ix < kCoveredX[strategy]; ix++) {
          float sigma = sigma_quant * lf.epf_sharp_lut[s->epf_sharpness[pos + ix + iy * W]];
          sigma = std::min(-1e-4f, sigma);
          s->inv_sigma[pos + ix + iy * W] = 1.0f / sigma;
        }
    }
  }
}

// DecodeAcMetadata, lib/jxl/dec_modular.cc:437-532
inline void VarDCTReadACMetadata(BitReader& br, VarDCTState* s, ModularFrameState* ms, size_t g) {
  const BlockRect r = DCGroupRect(s->dim, g);
  const size_t upper_bound = r.xs * r.ys;
  const size_t count = br.Read(CeilLog2(upper_bound)) + 1;
  ModImage image;
  image.w = r.xs;
  image.h = r.ys;
  image.bitdepth = ms->full.bitdepth;
  const size_t cx0 = r.x0 >> 3, cy0 = r.y0 >> 3, cw = (r.xs + 7) >> 3, chh = (r.ys + 7) >> 3;
  image.ch.emplace_back(cw, chh, 3, 3);
  image.ch.emplace_back(cw, chh, 3, 3);
  image.ch.emplace_back(count, 2, 0, 0);
  image.ch.emplace_back(r.xs, r.ys, 0, 0);
  ModularOptions opt;
  ModularDecode(br, image, StreamACMetadata(s->dim, g), opt, ms->has_tree ? &ms->tree : nullptr,
                ms->has_tree ? &ms->code : nullptr);
  UndoTransforms(image, WPHeader());
  auto clamp8 = [](int32_t v) { return static_cast<int8_t>(std::min(127, std::max(-128, v))); };
  for (size_t y = 0; y < chh; y++)
    for (size_t x = 0; x < cw; x++) {
      s->ytox_map[(cy0 + y) * s->cmap_w + cx0 + x] = clamp8(image.ch[0].Row(y)[x]);
      s->ytob_map[(cy0 + y) * s->cmap_w + cx0 + x] = clamp8(image.ch[1].Row(y)[x]);
    }
  size_t num = 0;
  const bool is444 = s->fh.Is444();
  const size_t W = s->dim.xsize_blocks;
  const size_t xlim = std::min(W, r.x0 + r.xs), ylim = std::min(s->dim.ysize_blocks, r.y0 + r.ys);
  const size_t gdb = s->dim.group_dim >> 3;  // kGroupDimInBlocks
  const int32_t* row_in_1 = image.ch[2].Row(0);
  const int32_t* row_in_2 = image.ch[2].Row(1);
  for (size_t iy = 0; iy < r.ys; iy++) {
    const size_t y = r.y0 + iy;
    const int32_t* row_in_3 = image.ch[3].Row(iy);
    for (size_t ix = 0; ix < r.xs; ix++) {
      const size_t x = r.x0 + ix;
      const int sharpness = row_in_3[ix];
      JXLO_CHECK(sharpness >= 0 && sharpness < 8, "corrupted sharpness field");
      s->epf_sharpness[y * W + x] = sharpness;
      if (s->acs[y * W + x] != 0xFF) continue;
      JXLO_CHECK(num < count, "corrupted AC metadata stream");
      const int raw = row_in_1[num];
      JXLO_CHECK(raw >= 0 && raw < kNumStrategies, "invalid AC strategy");
      s->used_acs |= 1u << raw;
      const size_t cbx = kCoveredX[raw], cby = kCoveredY[raw];
      JXLO_CHECK(!((cbx > 1 || cby > 1) && !is444), "AC strategy not compatible with chroma subsampling");
      const size_t next_x_ac = (x / gdb + 1) * gdb, next_y_ac = (y / gdb + 1) * gdb;
      JXLO_CHECK(x + cbx <= next_x_ac && x + cbx <= xlim, "invalid AC strategy, x overflow");
      JXLO_CHECK(y + cby <= next_y_ac && y + cby <= ylim, "invalid AC strategy, y overflow");
      for (size_t jy = 0; jy < cby; jy++)
        for (size_t jx = 0; jx < cbx; jx++) {
          uint8_t& e = s->acs[(y + jy) * W + x + jx];
          JXLO_CHECK(e == 0xFF, "invalid AC strategy: block overlap");
          e = static_cast<uint8_t>((raw << 1) | ((jy | jx) == 0 ? 1 : 0));
        }
      s->raw_quant[y * W + x] = 1 + std::max<int32_t>(0, std::min(255, row_in_2[num]));
      num++;
    }
  }
  if (s->fh.lf.epf_iters > 0) ComputeSigma(s, r);
}

// AdaptiveDCSmoothing, lib/jxl/compressed_dc.cc:124-195
inline void VarDCTFinalizeDC(VarDCTState* s) {
  if (s->fh.flags & kFlagSkipAdaptiveDCSmoothing) return;
  const size_t xsize = s->dc[0].w, ysize = s->dc[0].h;
  if (ysize <= 2 || xsize <= 2) return;
  const float w1 = 0.20345139757231578f, w2 = 0.0334829185968739f;
  const float w0 = 1.0f - 4.0f * (w1 + w2);
  Plane sm[3] = {s->dc[0], s->dc[1], s->dc[2]};
  for (size_t y = 1; y + 1 < ysize; y++) {
    for (size_t x = 1; x + 1 < xsize; x++) {
      float mc[3], smv[3];
      float gap = 0.5f;
      for (int c = 0; c < 3; c++) {
        const float* rt = s->dc[c].Row(y - 1);
        const float* rm = s->dc[c].Row(y);
        const float* rb = s->dc[c].Row(y + 1);
        mc[c] = rm[x];
        const float corner = (rt[x - 1] + rt[x + 1]) + (rb[x - 1] + rb[x + 1]);
        const float side = (rm[x - 1] + rm[x + 1]) + (rt[x] + rb[x]);
        smv[c] = std::fmaf(corner, w2, std::fmaf(side, w1, mc[c] * w0));
        gap = std::max(gap, std::fabs((mc[c] - smv[c]) / s->mul_dc[c]));
      }
      float factor = std::fmaf(-4.0f, gap, 3.0f);
      if (factor < 0.0f) factor = 0.0f;
      for (int c = 0; c < 3; c++) sm[c].Row(y)[x] = std::fmaf(smv[c] - mc[c], factor, mc[c]);
    }
  }
  for (int c = 0; c < 3; c++) s->dc[c] = std::move(sm[c]);
}

// ProcessACGlobal, lib/jxl/dec_frame.cc:367-430; DecodeCoeffOrders, lib/jxl/coeff_order.cc:99-160
inline void VarDCTReadGlobalAC(BitReader& br, VarDCTState* s, ModularFrameState* ms) {
  const bool all_default = br.Read(1);
  if (!all_default) {
    auto read_raw = [&](BitReader& r, int table, int w, int h, std::vector<int>* out) {
      // ModularFrameDecoder::DecodeQuantTable, lib/jxl/dec_modular.cc:765-812
      ModImage image;
      image.w = w;
      image.h = h;
      image.bitdepth = 8;
      for (int c = 0; c < 3; c++) image.ch.emplace_back(w, h);
      ModularOptions opt;
      ModularDecode(r, image, StreamQuantTable(s->dim, table), opt, ms->has_tree ? &ms->tree : nullptr,
                    ms->has_tree ? &ms->code : nullptr);
      UndoTransforms(image, WPHeader());
      out->assign(static_cast<size_t>(w) * h * 3, 0);
      for (int c = 0; c < 3; c++)
        for (int y = 0; y < h; y++)
          for (int x = 0; x < w; x++) {
            const int v = image.ch[c].Row(y)[x];
            JXLO_CHECK(v > 0, "invalid raw quantization table");
            (*out)[static_cast<size_t>(c) * w * h + y * w + x] = v;
          }
    };
    for (int i = 0; i < kNumQuantTablesI; i++) ReadQuantEncoding(br, i, &s->encodings[i], read_raw);
  }
  // EnsureComputed for the used strategies
  uint32_t kind_mask = 0;
  for (int i = 0; i < kNumStrategies; i++)
    if (s->used_acs & (1u << i)) kind_mask |= 1u << kStrategyToQuantTable[i];
  for (int t = 0; t < kNumQuantTablesI; t++) {
    if (!(kind_mask & (1u << t))) continue;
    const QuantEncoding enc = s->encodings[t].mode == kQuantLib ? LibraryEncoding(t) : s->encodings[t];
    s->tables[t] = ComputeQuantTable(enc, t);
  }
  const size_t num_histo_bits = CeilLog2(s->dim.num_groups);
  s->num_histograms = 1 + br.Read(num_histo_bits);
  const size_t num_passes = s->fh.passes.num_passes;
  s->passes.resize(num_passes);
  uint32_t acs_mask = 0;
  for (int o = 0; o < kNumStrategies; o++)
    if (s->used_acs & (1u << o)) acs_mask |= 1u << kStrategyOrder[o];
  for (size_t p = 0; p < num_passes; p++) {
    PassCode& pc = s->passes[p];
    pc.orders.assign(static_cast<size_t>(6156) * 64, 0);
    const uint32_t used_orders = ReadU32(br, Val(0x5F), Val(0x13), Val(0), Bits(13));  // kOrderEnc
    EntropyCode perm_code;
    std::unique_ptr<SymbolReader> reader;
    if (used_orders != 0) {
      ReadEntropyCode(br, 8, &perm_code);
      reader.reset(new SymbolReader(&perm_code, br));
    }
    uint32_t computed = 0;
    std::vector<uint32_t> natural, perm;
    for (int o = 0; o < kNumStrategies; o++) {
      const uint32_t ord = kStrategyOrder[o];
      if (computed & (1u << ord)) continue;
      computed |= 1u << ord;
      const bool used = (acs_mask & (1u << ord)) != 0;
      const size_t llf = static_cast<size_t>(kCoveredX[o]) * kCoveredY[o];
      const size_t size = 64 * llf;
      if (used || (used_orders & (1u << ord))) {
        natural.resize(size);
        NaturalCoeffOrder(o, natural.data());
      }
      if ((used_orders & (1u << ord)) == 0) {
        if (used)
          for (int c = 0; c < 3; c++)
            std::memcpy(&pc.orders[static_cast<size_t>(kCoeffOrderOffset[3 * ord + c]) * 64], natural.data(), size * 4);
      } else {
        for (int c = 0; c < 3; c++) {
          perm.resize(size);
          ReadPermutation(br, *reader, llf, size, perm.data());
          if (!used) continue;
          uint32_t* dest = &pc.orders[static_cast<size_t>(kCoeffOrderOffset[3 * ord + c]) * 64];
          for (size_t k = 0; k < size; k++) dest[k] = natural[perm[k]];
        }
      }
    }
    if (used_orders) JXLO_CHECK(reader->FinalStateOk(), "coefficient orders: bad ANS final state");
    const size_t num_contexts = static_cast<size_t>(s->num_histograms) * s->bctx.NumACContexts();
    ReadEntropyCode(br, num_contexts, &pc.code);
    // dec_frame.cc:406-408: padding for the unchecked read in the coefficient loop
    pc.code.ctx_map.resize(num_contexts + kZeroDensityContextLimit - kZeroDensityContextCount, 0);
  }
}

inline int32_t PredictFromTopAndLeft(const int32_t* row_top, const int32_t* row, size_t x, int32_t default_val) {
  if (x == 0) return row_top == nullptr ? default_val : row_top[x];
  if (row_top == nullptr) return row[x - 1];
  return (row_top[x] + row[x - 1] + 1) / 2;
}

// AdjustQuantBias, lib/jxl/quantizer-inl.h:34-71 (exact reciprocal, see header note)
inline float AdjustQuantBias(int c, int32_t quant_i, const float* biases) {
  const float quant = static_cast<float>(quant_i);
  const float abs_quant = std::fabs(quant);
  if (abs_quant < 1.125f) {
    if (!(abs_quant > 0.0f)) return 0.0f;
    return quant_i < 0 ? -biases[c] : biases[c];
  }
  return std::fmaf(-biases[3], 1.0f / quant, quant);
}

// One (group, pass) section: DecodeGroup / DecodeGroupImpl / DecodeACVarBlock.
// Coefficients are accumulated per group; the last pass dequantises and transforms.
inline void VarDCTReadACGroup(BitReader& br, VarDCTState* s, size_t g, size_t pass) {
  const FrameHeader& fh = s->fh;
  const BlockRect r = BlockGroupRect(s->dim, g);
  const size_t W = s->dim.xsize_blocks;
  const size_t num_passes = fh.passes.num_passes;
  PassCode& pc = s->passes[pass];
  const size_t histo_selector_bits = CeilLog2(s->num_histograms);
  size_t cur_histogram = 0;
  if (histo_selector_bits != 0) cur_histogram = br.Read(histo_selector_bits);
  JXLO_CHECK(cur_histogram < s->num_histograms, "invalid histogram selector");
  const size_t ctx_offset = cur_histogram * s->bctx.NumACContexts();
  SymbolReader reader(&pc.code, br);
  const uint32_t shift = fh.passes.shift[pass];
  std::vector<int32_t>& coeffs = s->group_coeffs[g];
  const size_t group_area = static_cast<size_t>(s->dim.group_dim) * s->dim.group_dim;
  if (coeffs.empty()) coeffs.assign(3 * group_area, 0);
  // num_nzeroes: 32 x 32 per channel
  const size_t nz_stride = s->dim.group_dim >> 3;
  std::vector<int32_t> nzeros(3 * nz_stride * nz_stride, 0);
  const uint32_t hshift[3] = {fh.HShift(0), fh.HShift(1), fh.HShift(2)};
  const uint32_t vshift[3] = {fh.VShift(0), fh.VShift(1), fh.VShift(2)};
  size_t offset = 0;
  for (size_t by = 0; by < r.ys; by++) {
    const int32_t* qf_row = &s->raw_quant[(r.y0 + by) * W + r.x0];
    const uint8_t* qdc_row = &s->quant_dc[(r.y0 + by) * W + r.x0];
    const uint8_t* acs_row = &s->acs[(r.y0 + by) * W + r.x0];
    for (size_t bx = 0; bx < r.xs;) {
      const uint8_t a = acs_row[bx];
      JXLO_CHECK(a != 0xFF, "AC strategy not set for a block");
      const int strategy = a >> 1;
      const size_t llf_x = kCoveredX[strategy];
      if (!(a & 1)) {
        bx += llf_x;
        continue;
      }
      const size_t log2_covered = kLog2Covered[strategy];
      const size_t covered = size_t{1} << log2_covered;
      const size_t size = covered * 64;
      for (int c : {1, 0, 2}) {
        const size_t sbx = bx >> hshift[c], sby = by >> vshift[c];
        if ((sbx << hshift[c]) != bx || (sby << vshift[c]) != by) continue;
        int32_t* block = coeffs.data() + c * group_area + offset;
        int32_t* row_nz = &nzeros[(c * nz_stride + sby) * nz_stride];
        const int32_t* row_nz_top = sby == 0 ? nullptr : row_nz - nz_stride;
        // DecodeACVarBlock (bx := sbx, lbx := bx as in GetBlockFromBitstream::LoadBlock)
        const int32_t predicted = PredictFromTopAndLeft(row_nz_top, row_nz, sbx, 32);
        const size_t ord = kStrategyOrder[strategy];
        const uint32_t* order = &pc.orders[static_cast<size_t>(kCoeffOrderOffset[3 * ord + c]) * 64];
        const size_t block_ctx = s->bctx.Context(qdc_row[bx], qf_row[sbx], ord, c);
        const uint32_t nzero_ctx = s->bctx.NonZeroContext(predicted, block_ctx) + ctx_offset;
        size_t nz = reader.ReadUint(nzero_ctx, br);
        JXLO_CHECK(nz <= size - covered, "invalid AC: too many non-zeros");
        for (size_t y = 0; y < kCoveredY[strategy]; y++)
          for (size_t x = 0; x < kCoveredX[strategy]; x++)
            row_nz[sbx + x + y * nz_stride] = (nz + covered - 1) >> log2_covered;
        const size_t histo_offset = ctx_offset + s->bctx.ZeroDensityContextsOffset(block_ctx);
        size_t prev = (nz > size / 16 ? 0 : 1);
        for (size_t k = covered; k < size && nz != 0; ++k) {
          const size_t nzl = (nz + covered - 1) >> log2_covered;
          const size_t kk = k >> log2_covered;
          const size_t ctx = histo_offset + (kCoeffNumNonzeroContext[nzl] + kCoeffFreqContext[kk]) * 2 + prev;
          const size_t u_coeff = reader.ReadUint(ctx, br);
          const size_t magnitude = u_coeff >> 1;
          const size_t neg_sign = (~u_coeff) & 1;
          const int32_t coeff = static_cast<int32_t>(static_cast<uint32_t>((magnitude ^ (neg_sign - 1)) << shift));
          block[order[k]] += coeff;
          prev = static_cast<size_t>(u_coeff != 0);
          nz -= prev;
        }
        JXLO_CHECK(nz == 0, "invalid AC: non-zeros left at the end of a block");
      }
      offset += size;
      bx += llf_x;
    }
  }
  JXLO_CHECK(reader.FinalStateOk(), "AC group: bad ANS final state");
  if (pass + 1 != num_passes) return;

  // ---- dequantise + inverse transforms (DecodeGroupImpl with draw == kDraw) ----
  std::vector<float> block(3 * group_area), scratch(3 * group_area + 1024);
  const float* biases = s->meta.quant_biases;
  static const float kNoBias[4] = {1, 1, 1, 0};
  if (getenv("JXLO_NO_QUANT_BIAS")) biases = kNoBias;  // diagnostic only (tests: JPEG coefficient check)
  offset = 0;
  for (size_t by = 0; by < r.ys; by++) {
    const size_t ty = (r.y0 + by) / 8;
    for (size_t bx = 0; bx < r.xs;) {
      const size_t pos = (r.y0 + by) * W + r.x0 + bx;
      const uint8_t a = s->acs[pos];
      const int strategy = a >> 1;
      const size_t llf_x = kCoveredX[strategy];
      if (!(a & 1)) {
        bx += llf_x;
        continue;
      }
      const size_t covered = size_t{1} << kLog2Covered[strategy];
      const size_t size = covered * 64;
      const size_t abs_tx = (r.x0 + bx) / 8;
      const float x_cc_mul = s->YtoXRatio(s->ytox_map[ty * s->cmap_w + abs_tx]);
      const float b_cc_mul = s->YtoBRatio(s->ytob_map[ty * s->cmap_w + abs_tx]);
      // DequantBlock, dec_group.cc:139-166
      const float scaled_dequant_s = s->inv_global_scale / s->raw_quant[pos];
      const float sd[3] = {scaled_dequant_s * s->x_dm_multiplier, scaled_dequant_s, scaled_dequant_s * s->b_dm_multiplier};
      const std::vector<float>& dm = s->tables[kStrategyToQuantTable[strategy]];
      JXLO_CHECK(dm.size() == 3 * size, "internal: dequant table missing");
      for (size_t k = 0; k < size; k++) {
        const float x_mul = dm[k] * sd[0], y_mul = dm[size + k] * sd[1], b_mul = dm[2 * size + k] * sd[2];
        const int32_t qx = coeffs[0 * group_area + offset + k];
        const int32_t qy = coeffs[1 * group_area + offset + k];
        const int32_t qb = coeffs[2 * group_area + offset + k];
        const float dq_x_cc = AdjustQuantBias(0, qx, biases) * x_mul;
        const float dq_y = AdjustQuantBias(1, qy, biases) * y_mul;
        const float dq_b_cc = AdjustQuantBias(2, qb, biases) * b_mul;
        block[k] = std::fmaf(x_cc_mul, dq_y, dq_x_cc);
        block[size + k] = dq_y;
        block[2 * size + k] = std::fmaf(b_cc_mul, dq_y, dq_b_cc);
      }
      for (int c = 0; c < 3; c++) {
        const size_t sbx = (r.x0 + bx) >> hshift[c], sby = (r.y0 + by) >> vshift[c];
        LowestFrequenciesFromDC(strategy, s->dc[c].Row(sby) + sbx, s->dc[c].w, block.data() + c * size);
      }
      for (int c : {1, 0, 2}) {
        const size_t abx = r.x0 + bx, aby = r.y0 + by;
        const size_t sbx = abx >> hshift[c], sby = aby >> vshift[c];
        if ((sbx << hshift[c]) != abx || (sby << vshift[c]) != aby) continue;
        TransformToPixels(strategy, block.data() + c * size, s->pix[c].Row(sby * 8) + sbx * 8, s->pix[c].w,
                          scratch.data());
      }
      offset += size;
      bx += llf_x;
    }
  }
  std::vector<int32_t>().swap(coeffs);
}

}  // namespace jxlo

#endif  // JXLO_VARDCT_H_
