// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
//
// libjxl's forward transforms and its effort-7 AcStrategy search, restated:
//   lib/jxl/enc_transforms-inl.h:30-58   ReinterpretingIDCT, :60-95 DCT2TopBlock, :97-395 AFVDCT4x4,
//                                        :401-460 AFVTransformFromPixels, :462-660 TransformFromPixels,
//                                        :662-790 DCFromLowestFrequencies
//   lib/jxl/enc_ac_strategy.cc:299-358   MultiBlockTransformCrosses{Horizontal,Vertical}Boundary
//                              :361-494  EstimateEntropy
//                              :496-599  FindBest8x8Transform
//                              :601-636  TryMergeAcs, :638-680 helpers
//                              :686-808  FindBestFirstLevelDivisionForSquare
//                              :810-1048 ProcessRectACS
//                              :1059-1117 AcStrategyHeuristics::Init (the three cost weights)
//   lib/jxl/enc_adaptive_quantization.cc:493-525 (1x1 masking image), :636-667 Blur1x1Masking
// Lane structure: the x86 AVX2 target (8 float lanes, Highway's SumOfLanes order), as for the chroma-from-luma sums.
// Plain C++ expressions of the reference are evaluated without contraction; MulAdd is fmaf.
// Two deliberate, documented deviations, both so that the CUDA encoder can reproduce the search bit for bit:
//   * log1p of the masking image is evaluated by a fixed double-precision series (the reference calls libm's log1p);
//   * the 8th root of the information loss is three IEEE square roots in double (the reference calls pow(x, 1 / 8.0)).
// Either differs from the libm result by at most an ulp of double before the value is rounded to float.
#ifndef JXLO_ENC_ACS_H_
#define JXLO_ENC_ACS_H_

#include <cmath>
#include <limits>

#include "jxlo_vardct.h"

namespace jxlo {

// ---------------------------------------------------------------- forward transforms (enc_transforms-inl.h)
inline void DCT2TopBlock(int S, const float* block, size_t stride, float* out) {
  float temp[64];
  const int num_2x2 = S / 2;
  for (int y = 0; y < num_2x2; y++)
    for (int x = 0; x < num_2x2; x++) {
      const float c00 = block[y * 2 * stride + x * 2], c01 = block[y * 2 * stride + x * 2 + 1];
      const float c10 = block[(y * 2 + 1) * stride + x * 2], c11 = block[(y * 2 + 1) * stride + x * 2 + 1];
      float r00 = c00 + c01 + c10 + c11, r01 = c00 + c01 - c10 - c11, r10 = c00 - c01 + c10 - c11, r11 = c00 - c01 - c10 + c11;
      r00 *= 0.25f;
      r01 *= 0.25f;
      r10 *= 0.25f;
      r11 *= 0.25f;
      temp[y * 8 + x] = r00;
      temp[y * 8 + num_2x2 + x] = r01;
      temp[(y + num_2x2) * 8 + x] = r10;
      temp[(y + num_2x2) * 8 + num_2x2 + x] = r11;
    }
  for (int y = 0; y < S; y++)
    for (int x = 0; x < S; x++) out[y * 8 + x] = temp[y * 8 + x];
}

// k4x4AFVBasisTranspose is the transpose of the decoder's basis (checked value for value against the table).
inline void AFVDCT4x4(const float* pixels, float* coeffs) {
  for (int i = 0; i < 16; i++) {
    float scalar = 0.0f;
    for (int j = 0; j < 16; j++) scalar = std::fmaf(pixels[j], kAFVBasis[i][j], scalar);
    coeffs[i] = scalar;
  }
}

inline void AFVTransformFromPixels(int afv_kind, const float* pixels, size_t pixels_stride, float* coefficients) {
  float scratch_space[4 * 8 * 5];
  const int afv_x = afv_kind & 1, afv_y = afv_kind / 2;
  float block[4 * 8] = {};
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 4; ix++)
      block[(afv_y == 1 ? 3 - iy : iy) * 4 + (afv_x == 1 ? 3 - ix : ix)] = pixels[(iy + 4 * afv_y) * pixels_stride + ix + 4 * afv_x];
  float coeff[4 * 4];
  AFVDCT4x4(block, coeff);
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 4; ix++) coefficients[iy * 2 * 8 + ix * 2] = coeff[iy * 4 + ix];
  ScaledDCT(4, 4, pixels + afv_y * 4 * pixels_stride + (afv_x == 1 ? 0 : 4), pixels_stride, block, scratch_space);
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 8; ix++) coefficients[iy * 2 * 8 + ix * 2 + 1] = block[iy * 4 + ix];
  ScaledDCT(4, 8, pixels + (afv_y == 1 ? 0 : 4) * pixels_stride, pixels_stride, block, scratch_space);
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 8; ix++) coefficients[(1 + iy * 2) * 8 + ix] = block[iy * 8 + ix];
  const float block00 = coefficients[0] * 0.25f, block01 = coefficients[1], block10 = coefficients[8];
  coefficients[0] = (block00 + block01 + 2 * block10) * 0.25f;
  coefficients[1] = (block00 - block01) * 0.5f;
  coefficients[8] = (block00 + block01 - 2 * block10) * 0.25f;
}

// (The loop `for ix < 8` over a 4x4 block above reads block[iy * 4 + ix] for ix up to 7 exactly like the reference
// does -- rows overlap; every coefficient position it writes twice ends with the value of the later write.)

// TransformFromPixels: pixels (stride) -> coefficients in the decoder's layout. scratch: 4 * covered pixels floats.
inline void TransformFromPixelsRef(int strategy, const float* pixels, size_t pixels_stride, float* coefficients, float* scratch) {
  switch (strategy) {
    case kIDENTITY: {
      for (int y = 0; y < 2; y++)
        for (int x = 0; x < 2; x++) {
          float block_dc = 0;
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) block_dc += pixels[(y * 4 + iy) * pixels_stride + x * 4 + ix];
          block_dc *= 1.0f / 16;
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) {
              if (ix == 1 && iy == 1) continue;
              coefficients[(y + iy * 2) * 8 + x + ix * 2] =
                  pixels[(y * 4 + iy) * pixels_stride + x * 4 + ix] - pixels[(y * 4 + 1) * pixels_stride + x * 4 + 1];
            }
          coefficients[(y + 2) * 8 + x + 2] = coefficients[y * 8 + x];
          coefficients[y * 8 + x] = block_dc;
        }
      const float block00 = coefficients[0], block01 = coefficients[1], block10 = coefficients[8], block11 = coefficients[9];
      coefficients[0] = (block00 + block01 + block10 + block11) * 0.25f;
      coefficients[1] = (block00 + block01 - block10 - block11) * 0.25f;
      coefficients[8] = (block00 - block01 + block10 - block11) * 0.25f;
      coefficients[9] = (block00 - block01 - block10 + block11) * 0.25f;
      return;
    }
    case kDCT8X4: {
      for (int x = 0; x < 2; x++) {
        float block[4 * 8];
        ScaledDCT(8, 4, pixels + x * 4, pixels_stride, block, scratch);
        for (int iy = 0; iy < 4; iy++)
          for (int ix = 0; ix < 8; ix++) coefficients[(x + iy * 2) * 8 + ix] = block[iy * 8 + ix];
      }
      const float block0 = coefficients[0], block1 = coefficients[8];
      coefficients[0] = (block0 + block1) * 0.5f;
      coefficients[8] = (block0 - block1) * 0.5f;
      return;
    }
    case kDCT4X8: {
      for (int y = 0; y < 2; y++) {
        float block[4 * 8];
        ScaledDCT(4, 8, pixels + y * 4 * pixels_stride, pixels_stride, block, scratch);
        for (int iy = 0; iy < 4; iy++)
          for (int ix = 0; ix < 8; ix++) coefficients[(y + iy * 2) * 8 + ix] = block[iy * 8 + ix];
      }
      const float block0 = coefficients[0], block1 = coefficients[8];
      coefficients[0] = (block0 + block1) * 0.5f;
      coefficients[8] = (block0 - block1) * 0.5f;
      return;
    }
    case kDCT4X4: {
      for (int y = 0; y < 2; y++)
        for (int x = 0; x < 2; x++) {
          float block[4 * 4];
          ScaledDCT(4, 4, pixels + y * 4 * pixels_stride + x * 4, pixels_stride, block, scratch);
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) coefficients[(y + iy * 2) * 8 + x + ix * 2] = block[iy * 4 + ix];
        }
      const float block00 = coefficients[0], block01 = coefficients[1], block10 = coefficients[8], block11 = coefficients[9];
      coefficients[0] = (block00 + block01 + block10 + block11) * 0.25f;
      coefficients[1] = (block00 + block01 - block10 - block11) * 0.25f;
      coefficients[8] = (block00 - block01 + block10 - block11) * 0.25f;
      coefficients[9] = (block00 - block01 - block10 + block11) * 0.25f;
      return;
    }
    case kDCT2X2:
      DCT2TopBlock(8, pixels, pixels_stride, coefficients);
      DCT2TopBlock(4, coefficients, 8, coefficients);
      DCT2TopBlock(2, coefficients, 8, coefficients);
      return;
    case kAFV0: case kAFV1: case kAFV2: case kAFV3:
      AFVTransformFromPixels(strategy - kAFV0, pixels, pixels_stride, coefficients);
      return;
    default:
      ScaledDCT(kCoveredY[strategy] * 8, kCoveredX[strategy] * 8, pixels, pixels_stride, coefficients, scratch);
      return;
  }
}

// DCFromLowestFrequencies: the DC value of each 8x8 block a varblock covers, from its lowest-frequency coefficients
// (ReinterpretingIDCT: the LLF corner rescaled to a covered_y x covered_x DCT, then the small inverse transform).
inline void DCFromLowestFrequencies(int strategy, const float* block, float* dc, size_t dc_stride) {
  if (!IsPlainDCT(strategy) || strategy == kDCT) {
    dc[0] = block[0];
    return;
  }
  const int ROWS = kCoveredY[strategy], COLS = kCoveredX[strategy];
  const size_t input_stride = 8 * std::max(ROWS, COLS);
  float small[32 * 32] = {};
  if (ROWS < COLS) {
    for (int y = 0; y < ROWS; y++)
      for (int x = 0; x < COLS; x++)
        small[y * COLS + x] = block[y * input_stride + x] * kResampleFromLLF[ROWS - 1 + y] * kResampleFromLLF[COLS - 1 + x];
  } else {
    for (int y = 0; y < COLS; y++)
      for (int x = 0; x < ROWS; x++)
        small[y * ROWS + x] = block[y * input_stride + x] * kResampleFromLLF[COLS - 1 + y] * kResampleFromLLF[ROWS - 1 + x];
  }
  float scratch[32 * 32 + 3 * 32];
  ScaledIDCT(ROWS, COLS, small, dc, dc_stride, scratch);
}

// ---------------------------------------------------------------- the 1x1 masking image
// log1p(x) for x >= 0 by its atanh series in double, fixed operation order (see the header comment).
inline double Log1pSeries(double x) {
  const double z = x / (2.0 + x), z2 = z * z;
  double sum = 0.0;
  for (int k = 30; k >= 0; k--) sum = sum * z2 + 1.0 / (2 * k + 1);
  return 2.0 * z * sum;
}

namespace acs {

inline float SumOfLanes8(const float l[8]) { return ((l[0] + l[4]) + (l[2] + l[6])) + ((l[1] + l[5]) + (l[3] + l[7])); }

// enc_adaptive_quantization.cc:493-525 over the whole (block-padded) plane, then Blur1x1Masking (:636-667): a
// Symmetric5 convolution (convolve_symmetric5.cc:28-118) with mirrored borders.
template <class RatioFn>
inline std::vector<float> Mask1x1(const float* y_plane, size_t xsize, size_t ysize, RatioFn ratio_of_derivatives) {
  const float match_gamma_offset = 0.019f;
  std::vector<float> m(xsize * ysize);
  for (size_t y = 0; y < ysize; y++) {
    const size_t y2 = y + 1 < ysize ? y + 1 : y, y1 = y > 0 ? y - 1 : y;
    const float *row_in = y_plane + y * xsize, *row_in1 = y_plane + y1 * xsize, *row_in2 = y_plane + y2 * xsize;
    for (size_t x = 0; x < xsize; x++) {
      const size_t x2 = x + 1 < xsize ? x + 1 : x, x1 = x > 0 ? x - 1 : x;
      const float base = 0.25f * (row_in2[x] + row_in1[x] + row_in[x1] + row_in[x2]);
      const float gammac = ratio_of_derivatives(row_in[x] + match_gamma_offset);
      float diff = std::fabs(gammac * (row_in[x] - base));
      diff = static_cast<float>(Log1pSeries(diff));
      m[y * xsize + x] = 1.0f / (diff + 0.01f);
    }
  }
  static const float kFilter[5] = {static_cast<float>(0.25647067633737227), static_cast<float>(0.2050056912354399075),
                                   static_cast<float>(0.154082048668497307), static_cast<float>(0.08149576591362004441),
                                   static_cast<float>(0.0512750104812308467)};
  double sum = 1.0 + 4 * (kFilter[0] + kFilter[1] + kFilter[2] + kFilter[4] + 2 * kFilter[3]);
  if (sum < 1e-5) sum = 1e-5;
  const float normalize = static_cast<float>(1.0 / sum);
  // WeightsSymmetric5 {c, r, R, d, D, L} = {1, f0, f2, f1, f4, f3} * normalize
  const float wc = normalize, wr = normalize * kFilter[0], wR = normalize * kFilter[2], wd = normalize * kFilter[1],
              wD = normalize * kFilter[4], wL = normalize * kFilter[3];
  std::vector<float> out(xsize * ysize);
  const int64_t w = xsize, h = ysize;
  auto mirror = [](int64_t v, int64_t size) {
    while (v < 0 || v >= size) v = v < 0 ? -v - 1 : 2 * size - 1 - v;
    return v;
  };
  auto row_sum = [&](int64_t x, int64_t y, float wx0, float wx1, float wx2) {
    const float* row = m.data() + mirror(y, h) * w;
    const float sum_2 = wx2 * (row[mirror(x - 2, w)] + row[mirror(x + 2, w)]);
    const float sum_1 = wx1 * (row[mirror(x - 1, w)] + row[mirror(x + 1, w)]);
    const float sum_0 = wx0 * row[x];
    return sum_2 + (sum_1 + sum_0);
  };
  for (int64_t y = 0; y < h; y++)
    for (int64_t x = 0; x < w; x++) {
      float sum0 = row_sum(x, y, wc, wr, wR);
      sum0 += row_sum(x, y - 2, wR, wL, wD);
      float sum1 = row_sum(x, y + 2, wR, wL, wD);
      sum0 += row_sum(x, y - 1, wr, wd, wL);
      sum1 += row_sum(x, y + 1, wr, wd, wL);
      out[y * w + x] = sum0 + sum1;
    }
  return out;
}

// ---------------------------------------------------------------- the search
struct Config {
  // dequantisation matrices ("Matrix") and the encoder-side weights ("InvMatrix": 1 / Matrix's source values with the
  // LLF corner zeroed, quant_weights.cc:322-349) per strategy: 3 * size floats each
  const std::vector<float>* matrix[kNumStrategies] = {};
  const std::vector<float>* inv_matrix[kNumStrategies] = {};
  const float* quant_field = nullptr;  // per block, W blocks per row
  size_t quant_stride = 0;
  const float* mask1x1 = nullptr;      // per pixel
  size_t mask_stride = 0, mask1x1_xsize = 0;
  const float* src[3] = {nullptr, nullptr, nullptr};  // XYB planes after inverse Gaborish
  size_t src_stride = 0;
  float info_loss_multiplier = 0, cost_delta = 0, zeros_mul = 0;
  float channel_mul[3] = {0, 0, 0};
  float Quant(size_t bx, size_t by) const { return quant_field[by * quant_stride + bx]; }
};

inline void InitConfig(Config* c, float butteraugli_distance) {  // AcStrategyHeuristics::Init, :1094-1115
  c->info_loss_multiplier = 1.2;
  c->zeros_mul = 9.3089059022677905;
  c->cost_delta = 10.833273317067883;
  static const float kBias = 0.13731742964354549;
  const float ratio = (butteraugli_distance + kBias) / (1.0f + kBias);
  static const float kPow1 = 0.33677806662454718, kPow2 = 0.50990926717963703, kPow3 = 0.36702940662370243;
  c->info_loss_multiplier *= std::pow(ratio, kPow1);
  c->zeros_mul *= std::pow(ratio, kPow2);
  c->cost_delta *= std::pow(ratio, kPow3);
  static const double kChannelMul[3] = {std::pow(10.2, 8.0), std::pow(1.0, 8.0), std::pow(1.03, 8.0)};
  for (int i = 0; i < 3; i++) c->channel_mul[i] = static_cast<float>(kChannelMul[i]);
}

inline uint32_t CeilLog2Nonzero(size_t v) { return v <= 1 ? 0 : FloorLog2(v - 1) + 1; }

// The strategy map of the frame while it is being chosen: (strategy << 1) | is_first per block, 0xFF = not set.
struct AcsImage {
  uint8_t* d;
  size_t w, h;
  bool IsFirst(size_t x, size_t y) const { return (d[y * w + x] & 1) != 0; }
  int Strategy(size_t x, size_t y) const { return d[y * w + x] >> 1; }
  void Set(size_t x, size_t y, int s) {
    for (size_t iy = 0; iy < kCoveredY[s]; iy++)
      for (size_t ix = 0; ix < kCoveredX[s]; ix++) d[(y + iy) * w + x + ix] = static_cast<uint8_t>((s << 1) | ((ix | iy) == 0 ? 1 : 0));
  }
};

inline bool CrossesHorizontalBoundary(const AcsImage& a, size_t start_x, size_t y, size_t end_x) {
  if (start_x >= a.w || y >= a.h) return false;
  if (y % 8 == 0) return false;
  end_x = std::min(end_x, a.w);
  const size_t start_x_limit = start_x & ~size_t{7};
  while (start_x != start_x_limit && !a.IsFirst(start_x, y)) --start_x;
  for (size_t x = start_x; x < end_x;) {
    if (a.IsFirst(x, y)) {
      x += kCoveredX[a.Strategy(x, y)];
    } else {
      return true;
    }
  }
  return false;
}

inline bool CrossesVerticalBoundary(const AcsImage& a, size_t x, size_t start_y, size_t end_y) {
  if (x >= a.w || start_y >= a.h) return false;
  if (x % 8 == 0) return false;
  end_y = std::min(end_y, a.h);
  const size_t start_y_limit = start_y & ~size_t{7};
  while (start_y != start_y_limit && !a.IsFirst(x, start_y)) --start_y;
  for (size_t y = start_y; y < end_y;) {
    if (a.IsFirst(x, y)) {
      y += kCoveredY[a.Strategy(x, y)];
    } else {
      return true;
    }
  }
  return false;
}

// `block`: 3 * size floats; `full_scratch`: kMaxCoeffArea (the quantisation error `mem`) + transform scratch.
constexpr size_t kAcsMaxCoeffArea = 64 * 64;  // (the search stops at 64x64)
inline float EstimateEntropy(int s, float entropy_mul, size_t x, size_t y, const Config& config, const float* cmap_factors,
                             float* block, float* full_scratch) {
  float entropy = 0.0f;
  float* mem = full_scratch;
  float* scratch_space = full_scratch + kAcsMaxCoeffArea;
  const size_t cbx = kCoveredX[s], cby = kCoveredY[s];
  const size_t num_blocks = cbx * cby;
  const size_t size = num_blocks * 64;
  for (size_t c = 0; c < 3; c++)
    TransformFromPixelsRef(s, config.src[c] + y * config.src_stride + x, config.src_stride, block + size * c, scratch_space);
  float quant_norm16 = 0;
  if (num_blocks == 1) {
    quant_norm16 = config.Quant(x / 8, y / 8);
  } else if (num_blocks == 2) {
    if (cby == 2) {
      quant_norm16 = std::max(config.Quant(x / 8, y / 8), config.Quant(x / 8, y / 8 + 1));
    } else {
      quant_norm16 = std::max(config.Quant(x / 8, y / 8), config.Quant(x / 8 + 1, y / 8));
    }
  } else {
    for (size_t iy = 0; iy < cby; iy++)
      for (size_t ix = 0; ix < cbx; ix++) {
        float qval = config.Quant(x / 8 + ix, y / 8 + iy);
        qval *= qval;
        qval *= qval;
        qval *= qval;
        quant_norm16 += qval * qval;
      }
    quant_norm16 /= num_blocks;
    quant_norm16 = FastPowf(quant_norm16, 1.0f / 16.0f);
  }
  const float quant = quant_norm16;
  float loss[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t c = 0; c < 3; c++) {
    const float* inv_matrix = config.inv_matrix[s]->data() + c * size;
    const float* matrix = config.matrix[s]->data() + c * size;
    const float cmap_factor = cmap_factors[c];
    float entropy_v[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nzeros_v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < size; i += 8)
      for (size_t l = 0; l < 8; l++) {
        const float in = block[c * size + i + l];
        const float in_y = block[size + i + l] * cmap_factor;
        const float im = inv_matrix[i + l];
        const float val = (in - in_y) * (im * quant);
        const float rval = std::nearbyintf(val);
        const float diff = val - rval;
        mem[i + l] = matrix[i + l] * diff;
        const float q = std::fabs(rval);
        entropy_v[l] = std::sqrt(q) + entropy_v[l];
        nzeros_v[l] = nzeros_v[l] + (q == 0.0f ? 0.0f : 1.0f);
      }
    {
      float lossc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      TransformToPixels(s, mem, block, cbx * 8, scratch_space);
      for (size_t iy = 0; iy < cby; iy++)
        for (size_t ix = 0; ix < cbx; ix++)
          for (size_t dy = 0; dy < 8; ++dy)
            for (size_t dx = 0; dx < 8; dx++) {
              float in = block[(iy * 8 + dy) * (cbx * 8) + ix * 8 + dx];
              if (x + ix * 8 + 8 <= config.mask1x1_xsize) {
                const float masku = std::fabs(config.mask1x1[(y + iy * 8 + dy) * config.mask_stride + x + ix * 8 + dx]);
                in = masku * in;
                in = in * in;
                in = in * in;
                in = in * in;
                lossc[dx] = lossc[dx] + in;
              }
            }
      for (int l = 0; l < 8; l++) {
        lossc[l] = config.channel_mul[c] * lossc[l];
        loss[l] = loss[l] + lossc[l];
      }
    }
    entropy += config.cost_delta * SumOfLanes8(entropy_v);
    const size_t num_nzeros = static_cast<size_t>(SumOfLanes8(nzeros_v));
    const size_t nbits = CeilLog2Nonzero(num_nzeros + 1) + 1;
    entropy += config.zeros_mul * (CeilLog2Nonzero(nbits + 17) + nbits);
  }
  const double mean_loss = SumOfLanes8(loss) / (num_blocks * 64);
  const float loss_scalar =
      static_cast<float>(std::sqrt(std::sqrt(std::sqrt(mean_loss))) * static_cast<double>(num_blocks * 64) / quant_norm16);
  entropy *= entropy_mul;
  entropy += config.info_loss_multiplier * loss_scalar;
  return entropy;
}

inline int FindBest8x8Transform(size_t x, size_t y, int encoding_speed_tier, float butteraugli_target, const Config& config,
                                const float* cmap_factors, float* block, float* scratch, float* entropy_out) {
  struct TransformTry8x8 {
    int type;
    int encoding_speed_tier_max_limit;
    double entropy_mul;
  };
  static const TransformTry8x8 kTransforms8x8[] = {
      {kDCT, 9, 0.8},
      {kDCT4X4, 5, 1.08},
      {kDCT2X2, 5, 0.95},
      {kDCT4X8, 4, 0.85931637428340035},
      {kDCT8X4, 4, 0.85931637428340035},
      {kIDENTITY, 5, 1.0427542510634957},
      {kAFV0, 4, 0.81779489591359944},
      {kAFV1, 4, 0.81779489591359944},
      {kAFV2, 4, 0.81779489591359944},
      {kAFV3, 4, 0.81779489591359944},
  };
  double best = 1e30;
  int best_tx = kTransforms8x8[0].type;
  for (const TransformTry8x8& tx : kTransforms8x8) {
    if (tx.encoding_speed_tier_max_limit < encoding_speed_tier) continue;
    float entropy_mul = tx.entropy_mul / kTransforms8x8[0].entropy_mul;
    if ((tx.type == kDCT2X2 || tx.type == kIDENTITY) && butteraugli_target < 5.0) {
      static const float kFavor2X2AtHighQuality = 0.4;
      const float weight = std::pow((5.0f - butteraugli_target) / 5.0f, 2.0);
      entropy_mul -= kFavor2X2AtHighQuality * weight;
    }
    if ((tx.type != kDCT && tx.type != kDCT2X2 && tx.type != kIDENTITY) && butteraugli_target > 4.0) {
      static const float kAvoidEntropyOfTransforms = 0.5;
      float mul = 1.0;
      if (butteraugli_target < 12.0) mul *= (12.0 - 4.0) / (butteraugli_target - 4.0);
      entropy_mul += kAvoidEntropyOfTransforms * mul;
    }
    const float entropy = EstimateEntropy(tx.type, entropy_mul, x, y, config, cmap_factors, block, scratch);
    if (entropy < best) {
      best_tx = tx.type;
      best = entropy;
    }
  }
  *entropy_out = best;
  return best_tx;
}

inline void SetEntropyForTransform(size_t cx, size_t cy, int s, float entropy, float* entropy_estimate) {
  for (size_t dy = 0; dy < kCoveredY[s]; ++dy)
    for (size_t dx = 0; dx < kCoveredX[s]; ++dx) entropy_estimate[(cy + dy) * 8 + cx + dx] = 0.0;
  entropy_estimate[cy * 8 + cx] = entropy;
}

inline void TryMergeAcs(int s, size_t bx, size_t by, size_t cx, size_t cy, const Config& config, const float* cmap_factors,
                        AcsImage* ac_strategy, float entropy_mul, uint8_t candidate_priority, uint8_t* priority,
                        float* entropy_estimate, float* block, float* scratch) {
  float entropy_current = 0;
  for (size_t iy = 0; iy < kCoveredY[s]; ++iy)
    for (size_t ix = 0; ix < kCoveredX[s]; ++ix) {
      if (priority[(cy + iy) * 8 + (cx + ix)] >= candidate_priority) return;
      entropy_current += entropy_estimate[(cy + iy) * 8 + (cx + ix)];
    }
  const float entropy_candidate = EstimateEntropy(s, entropy_mul, (bx + cx) * 8, (by + cy) * 8, config, cmap_factors, block, scratch);
  if (entropy_candidate >= entropy_current) return;
  for (size_t iy = 0; iy < kCoveredY[s]; iy++)
    for (size_t ix = 0; ix < kCoveredX[s]; ix++) {
      entropy_estimate[(cy + iy) * 8 + cx + ix] = 0;
      priority[(cy + iy) * 8 + cx + ix] = candidate_priority;
    }
  ac_strategy->Set(bx + cx, by + cy, s);
  entropy_estimate[cy * 8 + cx] = entropy_candidate;
}

inline int AcsSquare(size_t blocks) { return blocks == 2 ? kDCT16X16 : (blocks == 4 ? kDCT32X32 : kDCT64X64); }
inline int AcsVerticalSplit(size_t blocks) { return blocks == 2 ? kDCT16X8 : (blocks == 4 ? kDCT32X16 : kDCT64X32); }
inline int AcsHorizontalSplit(size_t blocks) { return blocks == 2 ? kDCT8X16 : (blocks == 4 ? kDCT16X32 : kDCT32X64); }

inline void FindBestFirstLevelDivisionForSquare(size_t blocks, bool allow_square_transform, size_t bx, size_t by, size_t cx,
                                                size_t cy, const Config& config, const float* cmap_factors, AcsImage* ac_strategy,
                                                float entropy_mul_JXK, float entropy_mul_JXJ, float* entropy_estimate, float* block,
                                                float* scratch) {
  const size_t blocks_half = blocks / 2;
  const int acs_rawJXK = AcsVerticalSplit(blocks), acs_rawKXJ = AcsHorizontalSplit(blocks), acs_rawJXJ = AcsSquare(blocks);
  const AcsImage& a = *ac_strategy;
  if (CrossesHorizontalBoundary(a, bx + cx, by + cy, bx + cx + blocks) ||
      CrossesHorizontalBoundary(a, bx + cx, by + cy + blocks, bx + cx + blocks) ||
      CrossesVerticalBoundary(a, bx + cx, by + cy, by + cy + blocks) ||
      CrossesVerticalBoundary(a, bx + cx + blocks, by + cy, by + cy + blocks)) {
    return;
  }
  const bool allow_JXK = !CrossesVerticalBoundary(a, bx + cx + blocks_half, by + cy, by + cy + blocks);
  const bool allow_KXJ = !CrossesHorizontalBoundary(a, bx + cx, by + cy + blocks_half, bx + cx + blocks);
  float entropy[2][2] = {};
  for (size_t dy = 0; dy < blocks; ++dy)
    for (size_t dx = 0; dx < blocks; ++dx) entropy[dy / blocks_half][dx / blocks_half] += entropy_estimate[(cy + dy) * 8 + (cx + dx)];
  const float kMax = std::numeric_limits<float>::max();
  float entropy_JXK_left = kMax, entropy_JXK_right = kMax, entropy_KXJ_top = kMax, entropy_KXJ_bottom = kMax, entropy_JXJ = kMax;
  if (allow_JXK) {
    if (a.Strategy(bx + cx + 0, by + cy) != acs_rawJXK)
      entropy_JXK_left = EstimateEntropy(acs_rawJXK, entropy_mul_JXK, (bx + cx + 0) * 8, (by + cy + 0) * 8, config, cmap_factors, block, scratch);
    if (a.Strategy(bx + cx + blocks_half, by + cy) != acs_rawJXK)
      entropy_JXK_right = EstimateEntropy(acs_rawJXK, entropy_mul_JXK, (bx + cx + blocks_half) * 8, (by + cy + 0) * 8, config, cmap_factors, block, scratch);
  }
  if (allow_KXJ) {
    if (a.Strategy(bx + cx, by + cy) != acs_rawKXJ)
      entropy_KXJ_top = EstimateEntropy(acs_rawKXJ, entropy_mul_JXK, (bx + cx + 0) * 8, (by + cy + 0) * 8, config, cmap_factors, block, scratch);
    if (a.Strategy(bx + cx, by + cy + blocks_half) != acs_rawKXJ)
      entropy_KXJ_bottom = EstimateEntropy(acs_rawKXJ, entropy_mul_JXK, (bx + cx + 0) * 8, (by + cy + blocks_half) * 8, config, cmap_factors, block, scratch);
  }
  if (allow_square_transform)
    entropy_JXJ = EstimateEntropy(acs_rawJXJ, entropy_mul_JXJ, (bx + cx + 0) * 8, (by + cy + 0) * 8, config, cmap_factors, block, scratch);
  const float costJxN = std::min(entropy_JXK_left, entropy[0][0] + entropy[1][0]) + std::min(entropy_JXK_right, entropy[0][1] + entropy[1][1]);
  const float costNxJ = std::min(entropy_KXJ_top, entropy[0][0] + entropy[0][1]) + std::min(entropy_KXJ_bottom, entropy[1][0] + entropy[1][1]);
  if (entropy_JXJ < costJxN && entropy_JXJ < costNxJ) {
    ac_strategy->Set(bx + cx, by + cy, acs_rawJXJ);
    SetEntropyForTransform(cx, cy, acs_rawJXJ, entropy_JXJ, entropy_estimate);
  } else if (costJxN < costNxJ) {
    if (entropy_JXK_left < entropy[0][0] + entropy[1][0]) {
      ac_strategy->Set(bx + cx, by + cy, acs_rawJXK);
      SetEntropyForTransform(cx, cy, acs_rawJXK, entropy_JXK_left, entropy_estimate);
    }
    if (entropy_JXK_right < entropy[0][1] + entropy[1][1]) {
      ac_strategy->Set(bx + cx + blocks_half, by + cy, acs_rawJXK);
      SetEntropyForTransform(cx + blocks_half, cy, acs_rawJXK, entropy_JXK_right, entropy_estimate);
    }
  } else {
    if (entropy_KXJ_top < entropy[0][0] + entropy[0][1]) {
      ac_strategy->Set(bx + cx, by + cy, acs_rawKXJ);
      SetEntropyForTransform(cx, cy, acs_rawKXJ, entropy_KXJ_top, entropy_estimate);
    }
    if (entropy_KXJ_bottom < entropy[1][0] + entropy[1][1]) {
      ac_strategy->Set(bx + cx, by + cy + blocks_half, acs_rawKXJ);
      SetEntropyForTransform(cx, cy + blocks_half, acs_rawKXJ, entropy_KXJ_bottom, entropy_estimate);
    }
  }
}

// One 64x64 tile (rect in blocks: bx, by, xs <= 8, ys <= 8) at libjxl's effort 7 (speed tier kSquirrel = 3, decoding
// speed tier 0). `block`: 3 * 4096 floats, `scratch`: 4096 + 4 * 4096 floats.
inline void ProcessRectACS(float butteraugli_target, const Config& config, size_t bx, size_t by, size_t xs, size_t ys,
                           const float cmap_factors[3], float* block, float* scratch, AcsImage* ac_strategy) {
  const int speed_tier = 3, decoding_speed_tier = 0;
  float entropy_estimate[64] = {};
  static const float k8x8mul1 = -0.4, k8x8mul2 = 1.0, k8x8base = 1.4;
  const float mul8x8 = k8x8mul2 + k8x8mul1 / (butteraugli_target + k8x8base);
  for (size_t iy = 0; iy < ys; iy++)
    for (size_t ix = 0; ix < xs; ix++) {
      float entropy = 0.0;
      const int best_of_8x8s = FindBest8x8Transform(8 * (bx + ix), 8 * (by + iy), speed_tier, butteraugli_target, config, cmap_factors,
                                                    block, scratch, &entropy);
      ac_strategy->Set(bx + ix, by + iy, best_of_8x8s);
      entropy_estimate[iy * 8 + ix] = entropy * mul8x8;
    }
  struct MergeTry {
    int type;
    uint8_t priority, decoding_speed_tier_max_limit, encoding_speed_tier_max_limit;
    float entropy_mul;
  };
  const float entropy_mul16X8 = 1.25, entropy_mul16X16 = 1.35, entropy_mul16X32 = 1.5, entropy_mul32X32 = 1.5, entropy_mul64X32 = 2.26,
              entropy_mul64X64 = 2.26;
  // (the reference's array has three more value-initialised entries: 8x8 DCTs at priority 0, which TryMergeAcs rejects
  // at its first test)
  const MergeTry kTransformsForMerge[6] = {
      {kDCT16X8, 2, 4, 5, entropy_mul16X8},   {kDCT8X16, 2, 4, 5, entropy_mul16X8},   {kDCT16X32, 4, 4, 4, entropy_mul16X32},
      {kDCT32X16, 4, 4, 4, entropy_mul16X32}, {kDCT64X32, 6, 1, 3, entropy_mul64X32}, {kDCT32X64, 6, 1, 3, entropy_mul64X32},
  };
  uint8_t priority[64] = {};
  const bool enable_32x32 = decoding_speed_tier < 4;
  for (const MergeTry& tx : kTransformsForMerge) {
    if (tx.decoding_speed_tier_max_limit < decoding_speed_tier) continue;
    const size_t tcx = kCoveredX[tx.type], tcy = kCoveredY[tx.type];
    for (size_t cy = 0; cy + tcy - 1 < ys; cy += tcy)
      for (size_t cx = 0; cx + tcx - 1 < xs; cx += tcx) {
        if (cy + 7 < ys && cx + 7 < xs) {
          if (decoding_speed_tier < 4 && tx.type == kDCT32X64) {
            if ((cy | cx) % 8 == 0)
              FindBestFirstLevelDivisionForSquare(8, true, bx, by, cx, cy, config, cmap_factors, ac_strategy, tx.entropy_mul,
                                                  entropy_mul64X64, entropy_estimate, block, scratch);
            continue;
          } else if (tx.type == kDCT32X16) {
            continue;
          }
        }
        if ((tx.type == kDCT16X32 && cy % 4 != 0) || (tx.type == kDCT32X16 && cx % 4 != 0)) continue;
        if (cy + 3 < ys && cx + 3 < xs) {
          if (tx.type == kDCT16X32) {
            if ((cy | cx) % 4 == 0)
              FindBestFirstLevelDivisionForSquare(4, enable_32x32, bx, by, cx, cy, config, cmap_factors, ac_strategy, tx.entropy_mul,
                                                  entropy_mul32X32, entropy_estimate, block, scratch);
            continue;
          } else if (tx.type == kDCT32X16) {
            continue;
          }
        }
        if ((tx.type == kDCT16X32 && cy % 4 != 0) || (tx.type == kDCT32X16 && cx % 4 != 0)) continue;
        if (cy + 1 < ys && cx + 1 < xs) {
          if (tx.type == kDCT8X16) {
            if ((cy | cx) % 2 == 0)
              FindBestFirstLevelDivisionForSquare(2, true, bx, by, cx, cy, config, cmap_factors, ac_strategy, tx.entropy_mul,
                                                  entropy_mul16X16, entropy_estimate, block, scratch);
            continue;
          } else if (tx.type == kDCT16X8) {
            continue;
          }
        }
        if ((tx.type == kDCT8X16 && cy % 2 == 1) || (tx.type == kDCT16X8 && cx % 2 == 1)) continue;
        TryMergeAcs(tx.type, bx, by, cx, cy, config, cmap_factors, ac_strategy, tx.entropy_mul, tx.priority, priority,
                    entropy_estimate, block, scratch);
      }
  }
  // (speed tier kSquirrel < kHare: the non-aligned passes run)
  for (size_t cy = 0; cy + 1 < ys; ++cy)
    for (size_t cx = 0; cx + 1 < xs; ++cx)
      if ((cy | cx) % 2 != 0)
        FindBestFirstLevelDivisionForSquare(2, true, bx, by, cx, cy, config, cmap_factors, ac_strategy, entropy_mul16X8,
                                            entropy_mul16X16, entropy_estimate, block, scratch);
  const size_t step = 2;  // speed_tier >= kTortoise
  for (size_t cy = 0; cy + 3 < ys; cy += step)
    for (size_t cx = 0; cx + 3 < xs; cx += step) {
      if ((cy | cx) % 4 == 0) continue;
      FindBestFirstLevelDivisionForSquare(4, enable_32x32, bx, by, cx, cy, config, cmap_factors, ac_strategy, entropy_mul16X32,
                                          entropy_mul32X32, entropy_estimate, block, scratch);
    }
}

}  // namespace acs
}  // namespace jxlo

#endif  // JXLO_ENC_ACS_H_
