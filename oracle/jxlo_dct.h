// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
//
// Scalar restatement of libjxl's separable scaled DCT / IDCT and of the 27
// inverse block transforms:
//   lib/jxl/dct-inl.h:50-165 (CoeffBundle), :167-222 (DCT1DImpl / IDCT1DImpl),
//   :282-335 (ComputeScaledDCT / ComputeScaledIDCT), lib/jxl/dct_scales.h (constants),
//   lib/jxl/dec_transforms-inl.h:30-59 (ReinterpretingDCT), :61-88 (IDCT2TopBlock),
//   :90-449 (AFV), :451-684 (TransformToPixels), :686-813 (LowestFrequenciesFromDC).
//
// Highway's MulAdd / NegMulAdd are single-rounding FMAs on every x86 target libjxl
// ships (AVX2, AVX-512) and on NEON; the oracle uses std::fmaf for them and plain
// float operations elsewhere (compiled with -ffp-contract=off), i.e. it restates the
// arithmetic of those builds. Every 1-D transform of the reference is lane-wise, so a
// scalar column-at-a-time restatement performs the same operations in the same order.
#ifndef JXLO_DCT_H_
#define JXLO_DCT_H_

#include <cmath>
#include <cstring>
#include <vector>

#include "jxlo_bits.h"

namespace jxlo {

#include "jxlo_tables.inc"

constexpr float kSqrt2f = 1.41421356237f;  // lib/jxl/base/common.h kSqrt2

inline const float* WcMultipliers(int n) { return kWcMultipliers + (n / 2 - 2); }

// IDCT1DImpl<N, 1>: `from` / `to` strided columns, tmp needs 2 * N floats.
inline void IDCT1D(int n, const float* from, size_t from_stride, float* to, size_t to_stride, float* tmp) {
  if (n == 1) {
    to[0] = from[0];
    return;
  }
  if (n == 2) {
    const float a = from[0], b = from[from_stride];
    to[0] = a + b;
    to[to_stride] = a - b;
    return;
  }
  const int h = n / 2;
  for (int i = 0; i < h; i++) tmp[i] = from[2 * i * from_stride];          // ForwardEvenOdd
  for (int i = h; i < n; i++) tmp[i] = from[(2 * (i - h) + 1) * from_stride];
  IDCT1D(h, tmp, 1, tmp, 1, tmp + n);
  for (int i = h - 1; i > 0; i--) tmp[h + i] = tmp[h + i] + tmp[h + i - 1];  // BTranspose
  tmp[h] = tmp[h] * kSqrt2f;
  IDCT1D(h, tmp + h, 1, tmp + h, 1, tmp + n);
  const float* mul = WcMultipliers(n);
  for (int i = 0; i < h; i++) {  // MultiplyAndAdd
    const float in1 = tmp[i], in2 = tmp[h + i];
    to[i * to_stride] = std::fmaf(mul[i], in2, in1);
    to[(n - i - 1) * to_stride] = std::fmaf(-mul[i], in2, in1);
  }
}

// DCT1DImpl<N, 1> on a contiguous column (in place), tmp needs 2 * N floats.
inline void DCT1DInPlace(int n, float* mem, float* tmp) {
  if (n == 1) return;
  if (n == 2) {
    const float a = mem[0], b = mem[1];
    mem[0] = a + b;
    mem[1] = a - b;
    return;
  }
  const int h = n / 2;
  for (int i = 0; i < h; i++) tmp[i] = mem[i] + mem[n - 1 - i];  // AddReverse
  DCT1DInPlace(h, tmp, tmp + n);
  for (int i = 0; i < h; i++) tmp[h + i] = mem[i] - mem[n - 1 - i];  // SubReverse
  const float* mul = WcMultipliers(n);
  for (int i = 0; i < h; i++) tmp[h + i] = tmp[h + i] * mul[i];  // Multiply
  DCT1DInPlace(h, tmp + h, tmp + n);
  tmp[h] = std::fmaf(tmp[h], kSqrt2f, tmp[h + 1]);  // B
  for (int i = 1; i + 1 < h; i++) tmp[h + i] = tmp[h + i] + tmp[h + i + 1];
  for (int i = 0; i < h; i++) {  // InverseEvenOdd
    mem[2 * i] = tmp[i];
    mem[2 * i + 1] = tmp[h + i];
  }
}

// DCT1D<N, M>: N-point DCT down each of the M columns, scaled by 1 / N.
inline void DCT1DColumns(int n, int m, const float* from, size_t from_stride, float* to, size_t to_stride,
                         float* tmp /* 3 * n */) {
  const float scale = 1.0f / n;
  for (int col = 0; col < m; col++) {
    for (int i = 0; i < n; i++) tmp[i] = from[i * from_stride + col];
    DCT1DInPlace(n, tmp, tmp + n);
    for (int i = 0; i < n; i++) to[i * to_stride + col] = scale * tmp[i];
  }
}

inline void IDCT1DColumns(int n, int m, const float* from, size_t from_stride, float* to, size_t to_stride,
                          float* tmp /* 2 * n */) {
  for (int col = 0; col < m; col++) IDCT1D(n, from + col, from_stride, to + col, to_stride, tmp);
}

inline void TransposeBlock(int rows, int cols, const float* from, size_t from_stride, float* to, size_t to_stride) {
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) to[j * to_stride + i] = from[i * from_stride + j];
}

// ComputeScaledDCT<ROWS, COLS>. Output layout: min(R, C) rows x max(R, C) columns.
inline void ScaledDCT(int R, int C, const float* from, size_t from_stride, float* to, float* scratch /* R*C + 3*max */) {
  float* block = scratch;
  float* tmp = scratch + R * C;
  if (R < C) {
    DCT1DColumns(R, C, from, from_stride, block, C, tmp);
    TransposeBlock(R, C, block, C, to, R);
    DCT1DColumns(C, R, to, R, block, R, tmp);
    TransposeBlock(C, R, block, R, to, C);
  } else {
    DCT1DColumns(R, C, from, from_stride, to, C, tmp);
    TransposeBlock(R, C, to, C, block, R);
    DCT1DColumns(C, R, block, R, to, R, tmp);
  }
}

// ComputeScaledIDCT<ROWS, COLS>; `from` (min x max layout) is clobbered.
inline void ScaledIDCT(int R, int C, float* from, float* to, size_t to_stride, float* scratch /* R*C + 2*max */) {
  float* block = scratch;
  float* tmp = scratch + R * C;
  if (R < C) {
    TransposeBlock(R, C, from, C, block, R);
    IDCT1DColumns(C, R, block, R, from, R, tmp);
    TransposeBlock(C, R, from, R, block, C);
    IDCT1DColumns(R, C, block, C, to, to_stride, tmp);
  } else {
    IDCT1DColumns(C, R, from, R, block, R, tmp);
    TransposeBlock(C, R, block, R, from, C);
    IDCT1DColumns(R, C, from, C, to, to_stride, tmp);
  }
}

// ---------------------------------------------------------------- AC strategies
// lib/jxl/ac_strategy.h:32-80, :148-174
constexpr int kNumStrategies = 27;
static const uint8_t kCoveredX[27] = {1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32};
static const uint8_t kCoveredY[27] = {1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16};
static const uint8_t kLog2Covered[27] = {0, 0, 0, 0, 2, 4, 1, 1, 2, 2, 3, 3, 0, 0, 0, 0, 0, 0, 6, 5, 5, 8, 7, 7, 10, 9, 9};
enum Strategy {
  kDCT = 0, kIDENTITY = 1, kDCT2X2 = 2, kDCT4X4 = 3, kDCT16X16 = 4, kDCT32X32 = 5, kDCT16X8 = 6, kDCT8X16 = 7,
  kDCT32X8 = 8, kDCT8X32 = 9, kDCT32X16 = 10, kDCT16X32 = 11, kDCT4X8 = 12, kDCT8X4 = 13, kAFV0 = 14, kAFV1 = 15,
  kAFV2 = 16, kAFV3 = 17, kDCT64X64 = 18, kDCT64X32 = 19, kDCT32X64 = 20, kDCT128X128 = 21, kDCT128X64 = 22,
  kDCT64X128 = 23, kDCT256X256 = 24, kDCT256X128 = 25, kDCT128X256 = 26
};
inline bool IsPlainDCT(int s) { return s == kDCT || (s >= kDCT16X16 && s <= kDCT16X32) || s >= kDCT64X64; }

// lib/jxl/dec_transforms-inl.h:61-88
inline void IDCT2TopBlock(int S, const float* block, size_t stride_out, float* out) {
  float temp[64];
  const int num_2x2 = S / 2;
  for (int y = 0; y < num_2x2; y++) {
    for (int x = 0; x < num_2x2; x++) {
      const float c00 = block[y * 8 + x];
      const float c01 = block[y * 8 + num_2x2 + x];
      const float c10 = block[(y + num_2x2) * 8 + x];
      const float c11 = block[(y + num_2x2) * 8 + num_2x2 + x];
      temp[y * 2 * 8 + x * 2] = c00 + c01 + c10 + c11;
      temp[y * 2 * 8 + x * 2 + 1] = c00 + c01 - c10 - c11;
      temp[(y * 2 + 1) * 8 + x * 2] = c00 - c01 + c10 - c11;
      temp[(y * 2 + 1) * 8 + x * 2 + 1] = c00 - c01 - c10 + c11;
    }
  }
  for (int y = 0; y < S; y++)
    for (int x = 0; x < S; x++) out[y * stride_out + x] = temp[y * 8 + x];
}

// lib/jxl/dec_transforms-inl.h:380-392
inline void AFVIDCT4x4(const float* coeffs, float* pixels) {
  for (int i = 0; i < 16; i++) {
    float pixel = 0.0f;
    for (int j = 0; j < 16; j++) pixel = std::fmaf(coeffs[j], kAFVBasis[j][i], pixel);
    pixels[i] = pixel;
  }
}

// lib/jxl/dec_transforms-inl.h:394-449
inline void AFVTransformToPixels(int afv_kind, const float* coefficients, float* pixels, size_t pixels_stride) {
  float scratch[4 * 8 * 4];
  const int afv_x = afv_kind & 1, afv_y = afv_kind / 2;
  float dcs[3];
  const float block00 = coefficients[0], block01 = coefficients[1], block10 = coefficients[8];
  dcs[0] = (block00 + block10 + block01) * 4.0f;
  dcs[1] = (block00 + block10 - block01);
  dcs[2] = block00 - block10;
  float coeff[16];
  coeff[0] = dcs[0];
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 4; ix++) {
      if (ix == 0 && iy == 0) continue;
      coeff[iy * 4 + ix] = coefficients[iy * 2 * 8 + ix * 2];
    }
  float block[4 * 8];
  AFVIDCT4x4(coeff, block);
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 4; ix++)
      pixels[(iy + afv_y * 4) * pixels_stride + afv_x * 4 + ix] =
          block[(afv_y == 1 ? 3 - iy : iy) * 4 + (afv_x == 1 ? 3 - ix : ix)];
  block[0] = dcs[1];
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 4; ix++) {
      if (ix == 0 && iy == 0) continue;
      block[iy * 4 + ix] = coefficients[iy * 2 * 8 + ix * 2 + 1];
    }
  ScaledIDCT(4, 4, block, pixels + afv_y * 4 * pixels_stride + (afv_x == 1 ? 0 : 4), pixels_stride, scratch);
  block[0] = dcs[2];
  for (int iy = 0; iy < 4; iy++)
    for (int ix = 0; ix < 8; ix++) {
      if (ix == 0 && iy == 0) continue;
      block[iy * 8 + ix] = coefficients[(1 + iy * 2) * 8 + ix];
    }
  ScaledIDCT(4, 8, block, pixels + (afv_y == 1 ? 0 : 4) * pixels_stride, pixels_stride, scratch);
}

// lib/jxl/dec_transforms-inl.h:451-684. `coefficients` may be clobbered.
// scratch: 3 * covered pixels floats.
inline void TransformToPixels(int strategy, float* coefficients, float* pixels, size_t pixels_stride, float* scratch) {
  switch (strategy) {
    case kIDENTITY: {
      float dcs[4];
      const float b00 = coefficients[0], b01 = coefficients[1], b10 = coefficients[8], b11 = coefficients[9];
      dcs[0] = b00 + b01 + b10 + b11;
      dcs[1] = b00 + b01 - b10 - b11;
      dcs[2] = b00 - b01 + b10 - b11;
      dcs[3] = b00 - b01 - b10 + b11;
      for (int y = 0; y < 2; y++) {
        for (int x = 0; x < 2; x++) {
          const float block_dc = dcs[y * 2 + x];
          float residual_sum = 0;
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) {
              if (ix == 0 && iy == 0) continue;
              residual_sum += coefficients[(y + iy * 2) * 8 + x + ix * 2];
            }
          pixels[(4 * y + 1) * pixels_stride + 4 * x + 1] = block_dc - residual_sum * (1.0f / 16);
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) {
              if (ix == 1 && iy == 1) continue;
              pixels[(y * 4 + iy) * pixels_stride + x * 4 + ix] =
                  coefficients[(y + iy * 2) * 8 + x + ix * 2] + pixels[(4 * y + 1) * pixels_stride + 4 * x + 1];
            }
          pixels[y * 4 * pixels_stride + x * 4] =
              coefficients[(y + 2) * 8 + x + 2] + pixels[(4 * y + 1) * pixels_stride + 4 * x + 1];
        }
      }
      return;
    }
    case kDCT8X4: {
      const float block0 = coefficients[0], block1 = coefficients[8];
      const float dcs[2] = {block0 + block1, block0 - block1};
      for (int x = 0; x < 2; x++) {
        float block[4 * 8];
        block[0] = dcs[x];
        for (int iy = 0; iy < 4; iy++)
          for (int ix = 0; ix < 8; ix++) {
            if (ix == 0 && iy == 0) continue;
            block[iy * 8 + ix] = coefficients[(x + iy * 2) * 8 + ix];
          }
        ScaledIDCT(8, 4, block, pixels + x * 4, pixels_stride, scratch);
      }
      return;
    }
    case kDCT4X8: {
      const float block0 = coefficients[0], block1 = coefficients[8];
      const float dcs[2] = {block0 + block1, block0 - block1};
      for (int y = 0; y < 2; y++) {
        float block[4 * 8];
        block[0] = dcs[y];
        for (int iy = 0; iy < 4; iy++)
          for (int ix = 0; ix < 8; ix++) {
            if (ix == 0 && iy == 0) continue;
            block[iy * 8 + ix] = coefficients[(y + iy * 2) * 8 + ix];
          }
        ScaledIDCT(4, 8, block, pixels + y * 4 * pixels_stride, pixels_stride, scratch);
      }
      return;
    }
    case kDCT4X4: {
      float dcs[4];
      const float b00 = coefficients[0], b01 = coefficients[1], b10 = coefficients[8], b11 = coefficients[9];
      dcs[0] = b00 + b01 + b10 + b11;
      dcs[1] = b00 + b01 - b10 - b11;
      dcs[2] = b00 - b01 + b10 - b11;
      dcs[3] = b00 - b01 - b10 + b11;
      for (int y = 0; y < 2; y++)
        for (int x = 0; x < 2; x++) {
          float block[16];
          block[0] = dcs[y * 2 + x];
          for (int iy = 0; iy < 4; iy++)
            for (int ix = 0; ix < 4; ix++) {
              if (ix == 0 && iy == 0) continue;
              block[iy * 4 + ix] = coefficients[(y + iy * 2) * 8 + x + ix * 2];
            }
          ScaledIDCT(4, 4, block, pixels + y * 4 * pixels_stride + x * 4, pixels_stride, scratch);
        }
      return;
    }
    case kDCT2X2: {
      float coeffs[64];
      std::memcpy(coeffs, coefficients, sizeof(coeffs));
      IDCT2TopBlock(2, coeffs, 8, coeffs);
      IDCT2TopBlock(4, coeffs, 8, coeffs);
      IDCT2TopBlock(8, coeffs, 8, coeffs);
      for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) pixels[y * pixels_stride + x] = coeffs[y * 8 + x];
      return;
    }
    case kAFV0: case kAFV1: case kAFV2: case kAFV3:
      AFVTransformToPixels(strategy - kAFV0, coefficients, pixels, pixels_stride);
      return;
    default:
      ScaledIDCT(kCoveredY[strategy] * 8, kCoveredX[strategy] * 8, coefficients, pixels, pixels_stride, scratch);
      return;
  }
}

// lib/jxl/dec_transforms-inl.h:686-813 (+ ReinterpretingDCT :30-59)
inline void LowestFrequenciesFromDC(int strategy, const float* dc, size_t dc_stride, float* llf) {
  if (!IsPlainDCT(strategy) || strategy == kDCT) {
    llf[0] = dc[0];
    return;
  }
  const int R = kCoveredY[strategy], C = kCoveredX[strategy];
  const size_t out_stride = 8 * std::max(R, C);
  float block[32 * 32];
  float scratch[32 * 32 + 3 * 32];
  ScaledDCT(R, C, dc, dc_stride, block, scratch);
  if (R < C) {
    for (int y = 0; y < R; y++)
      for (int x = 0; x < C; x++)
        llf[y * out_stride + x] = block[y * C + x] * kResampleToLLF[R - 1 + y] * kResampleToLLF[C - 1 + x];
  } else {
    for (int y = 0; y < C; y++)
      for (int x = 0; x < R; x++)
        llf[y * out_stride + x] = block[y * R + x] * kResampleToLLF[C - 1 + y] * kResampleToLLF[R - 1 + x];
  }
}

}  // namespace jxlo

#endif  // JXLO_DCT_H_
