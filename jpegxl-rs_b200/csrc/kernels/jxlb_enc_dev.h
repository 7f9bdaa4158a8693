// jxl_b200 device code of the VarDCT encode path (the mirror of jxlb_vardct_dev.h).
//
//   DevEncXybPixel        RGB8 -> XYB (sRGB table, opsin mix, cube root)      lib/jxl/enc_xyb.cc:41-104, :226-275
//   DevEncStrategyGroup   AcStrategy per 256x256 group (variance heuristic)     lib/jxl/enc_ac_strategy.cc (role of)
//   DevEncNumberBlocks    raster numbering of the varblocks of a DC group       lib/jxl/enc_modular.cc AddACMetadata
//   DevEncDcBlock         DC = block mean, quantised with CfL on DC             lib/jxl/enc_cache.cc:207-228, compressed_dc.cc
//   DevEncVarblock        forward transform + quantisation (Y round trip, CfL)  lib/jxl/enc_group.cc:46-90, :370-524
//   DevEncTokenizeGroup   coefficient tokens + histogram counts                 lib/jxl/enc_entropy_coder.cc:148-244
//   DevEncModularToken    DC / AC-metadata tokens under the fixed global tree   lib/jxl/modular/encoding/enc_encoding.cc
//   DevRansEmit           rANS writer (reverse pass)                            lib/jxl/enc_ans.cc:1731-1816, enc_ans.h:53-70
//
// The heuristics are those of the oracle's plain encoder (oracle/jxlo_encode.h), not libjxl's rate-distortion
// search; every statement is arranged so that the bytes produced equal the oracle's.
#ifndef JXLB_ENC_DEV_H_
#define JXLB_ENC_DEV_H_

#include "jxlb_enc_desc.h"
#include "jxlb_vardct_dev.h"

namespace jxlb {

struct DevEPools {
  const uint8_t* bytes_in;   // RGB8 inputs
  float* farena;
  int32_t* iarena;
  uint8_t* barena;
  uint2* tokens;
  const float* srgb_lut;     // 256 entries: sRGB byte -> linear
  const float* fpool;        // shared VarDCT tables (dequantisation tables, WcMultipliers)
  const uint16_t* opool;     // natural coefficient orders
  const uint16_t* opool_custom;  // per-frame custom orders (DevEFrame::custom_order), uploaded after the statistics
  const uint8_t* sample_bits;    // sample_bits[n]: the n-th varblock (group order) enters the statistics when only half do
  const uint32_t* upool;     // packed StrategyInfo, coefficient context tables
  uint32_t table_off[17], order_off[13];
  uint32_t wc_off, sinfo_off, ctxtab_off;
  const uint8_t* ac_cluster_of;  // 15 * 495 contexts -> cluster
  const DevEncTreeNode* tree;
  uint32_t num_dc_groups_for_tree;
};

JXLB_HD int32_t DevRoundToInt(float v) {
#if defined(__CUDA_ARCH__)
  return __float2int_rn(v);
#else
  return static_cast<int32_t>(lrintf(v));
#endif
}

JXLB_HD float DevCubeRootAndAdd(float x, float add) {  // lib/jxl/base/fast_math-inl.h:177-224
  const float k1_3 = 1.0f / 3, k4_3 = 4.0f / 3;
  const float xa_3 = k1_3 * x;
  const int32_t m1 = DevFloatToBits(x);
  const int32_t m2 = m1 == 0 ? 0 : 0x54800000 - (m1 >> 23) * 0x002AAAAA;
  float r = DevBitsToFloat(m2);
  for (int i = 0; i < 3; i++) {
    const float r2 = r * r;
    r = fmaf(-xa_3, r2 * r2, k4_3 * r);
  }
  float r2 = r * r;
  r = fmaf(k1_3, fmaf(-x, r2 * r2, r), r);
  r2 = r * r;
  return fmaf(r2, x, add);
}

// One sample position of the padded planes.
JXLB_HD void DevEncXybPixel(const DevEPools& E, const DevEFrame& ef, uint32_t x, uint32_t y) {
  const uint32_t PW = ef.xblocks * 8;
  const uint32_t sx = x < ef.xsize ? x : ef.xsize - 1, sy = y < ef.ysize ? y : ef.ysize - 1;
  const uint8_t* px = E.bytes_in + ef.rgb + (static_cast<size_t>(sy) * ef.xsize + sx) * (ef.has_alpha ? 4 : 3);
  const float r = E.srgb_lut[px[0]], g = E.srgb_lut[px[1]], b = E.srgb_lut[px[2]];
  const float bias = 0.0037930732552754493f;
  const float kNegBiasCbrt = -0.15595420054924863f;
  const float m00 = 0.30f, m01 = 1.0f - 0.078f - 0.30f, m02 = 0.078f;
  const float m10 = 0.23f, m11 = 1.0f - 0.078f - 0.23f, m12 = 0.078f;
  const float m20 = 0.24342268924547819f, m21 = 0.20476744424496821f, m22 = 1.0f - 0.24342268924547819f - 0.20476744424496821f;
  float mixed0 = m00 * r + m01 * g + m02 * b + bias;
  float mixed1 = m10 * r + m11 * g + m12 * b + bias;
  float mixed2 = m20 * r + m21 * g + m22 * b + bias;
  if (mixed0 < 0) mixed0 = 0;
  if (mixed1 < 0) mixed1 = 0;
  if (mixed2 < 0) mixed2 = 0;
  mixed0 = DevCubeRootAndAdd(mixed0, kNegBiasCbrt);
  mixed1 = DevCubeRootAndAdd(mixed1, kNegBiasCbrt);
  mixed2 = DevCubeRootAndAdd(mixed2, kNegBiasCbrt);
  const size_t at = static_cast<size_t>(y) * PW + x;
  const uint64_t* dst = ef.gab ? ef.xyb_raw : ef.xyb;
  E.farena[dst[0] + at] = 0.5f * (mixed0 - mixed1);
  E.farena[dst[1] + at] = 0.5f * (mixed0 + mixed1);
  E.farena[dst[2] + at] = mixed2;
}

// GaborishInverse (lib/jxl/enc_gaborish.cc:21-70): the symmetric 5x5 sharpening of lib/jxl/convolve_symmetric5.cc:28-118
// with its summation order and mirrored borders at the size of the padded planes. One sample of channel c.
JXLB_HD int32_t DevEncMirror(int32_t v, int32_t size) {
  while (v < 0 || v >= size) v = v < 0 ? -v - 1 : 2 * size - 1 - v;
  return v;
}
template <bool INTERIOR>
JXLB_HD void DevEncGaborishInvPixel(const DevEPools& E, const DevEFrame& ef, uint32_t c, int32_t x, int32_t y) {
  const int32_t PW = static_cast<int32_t>(ef.xblocks * 8), PH = static_cast<int32_t>(ef.yblocks * 8);
  const float* in = E.farena + ef.xyb_raw[c];
  const float wc = ef.gabinv_w[0], wr = ef.gabinv_w[1], wR = ef.gabinv_w[2], wd = ef.gabinv_w[3], wD = ef.gabinv_w[4],
              wL = ef.gabinv_w[5];
  const int32_t xm2 = INTERIOR ? x - 2 : DevEncMirror(x - 2, PW), xm1 = INTERIOR ? x - 1 : DevEncMirror(x - 1, PW);
  const int32_t xp1 = INTERIOR ? x + 1 : DevEncMirror(x + 1, PW), xp2 = INTERIOR ? x + 2 : DevEncMirror(x + 2, PW);
  auto row_sum = [&](int32_t yy, float wx0, float wx1, float wx2) {
    const float* row = in + static_cast<size_t>(INTERIOR ? yy : DevEncMirror(yy, PH)) * PW;
    const float sum_2 = wx2 * (row[xm2] + row[xp2]);
    const float sum_1 = wx1 * (row[xm1] + row[xp1]);
    const float sum_0 = wx0 * row[x];
    return sum_2 + (sum_1 + sum_0);
  };
  float sum0 = row_sum(y, wc, wr, wR);
  sum0 += row_sum(y - 2, wR, wL, wD);
  float sum1 = row_sum(y + 2, wR, wL, wD);
  sum0 += row_sum(y - 1, wr, wd, wL);
  sum1 += row_sum(y + 1, wr, wd, wL);
  E.farena[ef.xyb[c] + static_cast<size_t>(y) * PW + x] = sum0 + sum1;
}

// ---- adaptive quantisation: InitialQuantField / AdaptiveQuantizationMap, lib/jxl/enc_adaptive_quantization.cc:78-716,
// one 64x64 tile (ComputeTile, :468-630) per call, `nt` cooperating threads. The planes are those before inverse
// Gaborish. `sm`: kAqSmemFloats floats. Which pixels of a tile row take the SIMD form of the Laplacian and which the
// scalar form (different association of the four neighbours) follows from the tile's own x range with 8 lanes
// (AVX2), as in the oracle. The masking images of the AcStrategy search (mask, mask1x1) are not produced.
constexpr uint32_t kAqSmemFloats = 72 * 72 + 18 * 18 + 16 * 16 + 64 * 24;
constexpr float kAqSGmul = 226.77216153508914f;
constexpr float kAqSGmul2 = 1.0f / 73.377132366608819f;
constexpr float kAqLog2 = 0.693147181f;
constexpr float kAqSGRetMul = kAqSGmul2 * 18.6580932135f * kAqLog2;
constexpr float kAqSGVOffset = 7.7825991679894591f;

template <bool INVERT>
JXLB_HD float DevAqRatioOfDerivatives(float v) {  // :117-136
  const float kEpsilon = 1e-2;
  v = v < 0.0f ? 0.0f : v;
  const float kNumMul = kAqSGRetMul * 3 * kAqSGmul;
  const float kVOffset = kAqSGVOffset * kAqLog2 + kEpsilon;
  const float kDenMul = kAqLog2 * kAqSGmul;
  const float v2 = v * v;
  const float num = fmaf(kNumMul, v2, kEpsilon);
  const float den = fmaf(kDenMul * v, v2, kVOffset);
  return INVERT ? num / den : den / num;
}

JXLB_HD float DevFastLog2f(float v) {  // lib/jxl/base/fast_math-inl.h:46-66
  const float lp[3] = {-1.8503833400518310E-06f, 1.4287160470083755E+00f, 7.4245873327820566E-01f};
  const float lq[3] = {9.9032814277590719E-01f, 1.0096718572241148E+00f, 1.7409343003366853E-01f};
  const int32_t x_bits = DevFloatToBits(v);
  const int32_t exp_bits = x_bits - 0x3f2aaaab;
  const int32_t exp_shifted = exp_bits >> 23;
  const float mantissa = DevBitsToFloat(x_bits - static_cast<int32_t>(static_cast<uint32_t>(exp_shifted) << 23));
  return DevRational2(mantissa - 1.0f, lp, lq) + static_cast<float>(exp_shifted);
}

JXLB_HD float DevFastPow2f(float x) {  // lib/jxl/base/fast_math-inl.h:69-84
  const float floorx = floorf(x);
  const float exp = DevBitsToFloat(static_cast<int32_t>(static_cast<uint32_t>(static_cast<int32_t>(floorx) + 127) << 23));
  const float frac = x - floorx;
  float num = frac + 1.01749063e+01f;
  num = fmaf(num, frac, 4.88687798e+01f);
  num = fmaf(num, frac, 9.85506591e+01f);
  num = num * exp;
  float den = fmaf(frac, 2.10242958e-01f, -2.22328856e-02f);
  den = fmaf(den, frac, -1.94414990e+01f);
  den = fmaf(den, frac, 9.85506633e+01f);
  return num / den;
}

JXLB_HD void DevAqStoreMin4(float v, float& min0, float& min1, float& min2, float& min3) {  // :356-375
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

JXLB_HD void DevAqSwapIfGreater(float& a, float& b) {
  if (a > b) {
    const float t = a;
    a = b;
    b = t;
  }
}

template <int SCOPE>
JXLB_HD void DevEncAqTile(const DevEPools& E, const DevEFrame& ef, uint32_t tx, uint32_t ty, uint32_t tid, uint32_t nt, float* sm) {
  const uint32_t W = ef.xblocks, H = ef.yblocks, xsize = W * 8, ysize = H * 8;
  const uint32_t bx0 = tx * 8, by0 = ty * 8, bx1 = bx0 + 8 < W ? bx0 + 8 : W, by1 = by0 + 8 < H ? by0 + 8 : H;
  uint32_t x_start = bx0 * 8, x_end = bx1 * 8, y_start = by0 * 8, y_end = by1 * 8;
  if (x_start != 0) x_start -= 4;
  if (x_end != xsize) x_end += 4;
  if (y_start != 0) y_start -= 4;
  if (y_end != ysize) y_end += 4;
  const uint32_t dw = x_end - x_start, dh = y_end - y_start, pw = dw / 4, ph = dh / 4;
  float* diff = sm;                 // [dh][72]
  float* pre = sm + 72 * 72;        // [ph][18]
  float* fz = pre + 18 * 18;        // [16][16]
  float* lanes = fz + 16 * 16;      // [64 blocks][3 modulations][8 lanes]
  const float* P[3] = {E.farena + (ef.gab ? ef.xyb_raw[0] : ef.xyb[0]), E.farena + (ef.gab ? ef.xyb_raw[1] : ef.xyb[1]),
                       E.farena + (ef.gab ? ef.xyb_raw[2] : ef.xyb[2])};
  // pixels [simd_begin, simd_end) of a row go through the 8-lane loop (`for (; x + 1 + Lanes < x_end; x += Lanes)`)
  const uint32_t simd_begin = x_start == 0 ? 1 : x_start;
  const uint32_t simd_iters = x_end > simd_begin + 9 ? (x_end - simd_begin - 9 + 7) / 8 : 0;
  const uint32_t simd_end = simd_begin + 8 * simd_iters;
  const float kMaskMul = sqrtf(static_cast<float>(211.66567973503678f * 1e8));
  for (uint32_t i = tid; i < dw * dh; i += nt) {
    const uint32_t x = x_start + i % dw, y = y_start + i / dw;
    const uint32_t y2 = y + 1 < ysize ? y + 1 : y, y1 = y > 0 ? y - 1 : y;
    const float* row_in = P[1] + static_cast<size_t>(y) * xsize;
    const float* row_in1 = P[1] + static_cast<size_t>(y1) * xsize;
    const float* row_in2 = P[1] + static_cast<size_t>(y2) * xsize;
    float base;
    if (x >= simd_begin && x < simd_end) {
      base = 0.25f * ((row_in[x + 1] + row_in[x - 1]) + (row_in2[x] + row_in1[x]));
    } else {
      const uint32_t x2 = x + 1 < xsize ? x + 1 : x, x1 = x > 0 ? x - 1 : x;
      base = 0.25f * (((row_in2[x] + row_in1[x]) + row_in[x1]) + row_in[x2]);
    }
    const float gammac = DevAqRatioOfDerivatives<false>(row_in[x] + 0.019f);
    float d = gammac * (row_in[x] - base);
    d = d * d;
    if (d >= 0.2f) d = 0.2f;
    diff[(i / dw) * 72 + i % dw] = 0.25f * sqrtf(fmaf(d, kMaskMul, 27.505837037000106f));  // MaskingSqrt, :341-348
  }
  CoopSync<SCOPE>();
  for (uint32_t i = tid; i < pw * ph; i += nt) {  // 4x4 sums: rows accumulate first, then the four columns
    const uint32_t cx = i % pw, cy = i / pw;
    float col[4];
    for (uint32_t k = 0; k < 4; k++) {
      const float* d0 = diff + (cy * 4) * 72 + cx * 4 + k;
      col[k] = ((d0[0] + d0[72]) + d0[144]) + d0[216];
    }
    pre[cy * 18 + cx] = (((col[0] + col[1]) + col[2]) + col[3]) * 0.25f;
  }
  CoopSync<SCOPE>();
  // FuzzyErosion (:380-451)
  const uint32_t fx0 = x_start % 8 == 0 ? 0 : 1, fy0 = y_start % 8 == 0 ? 0 : 1;
  const uint32_t fw = (bx1 - bx0) * 2, fh = (by1 - by0) * 2;
  for (uint32_t i = tid; i < fw * fh; i += nt) {
    const uint32_t x = i % fw + fx0, y = i / fw + fy0;
    const uint32_t xm1 = x >= 1 ? x - 1 : x, xp1 = x + 1 < pw ? x + 1 : x;
    const uint32_t ym1 = y >= 1 ? y - 1 : y, yp1 = y + 1 < ph ? y + 1 : y;
    const float *rowt = pre + ym1 * 18, *row = pre + y * 18, *rowb = pre + yp1 * 18;
    float min0 = row[x], min1 = row[xm1], min2 = row[xp1], min3 = rowt[xm1];
    DevAqSwapIfGreater(min0, min1);
    DevAqSwapIfGreater(min0, min2);
    DevAqSwapIfGreater(min0, min3);
    DevAqSwapIfGreater(min1, min2);
    DevAqSwapIfGreater(min1, min3);
    DevAqSwapIfGreater(min2, min3);
    DevAqStoreMin4(rowt[x], min0, min1, min2, min3);
    DevAqStoreMin4(rowt[xp1], min0, min1, min2, min3);
    DevAqStoreMin4(rowb[xm1], min0, min1, min2, min3);
    DevAqStoreMin4(rowb[x], min0, min1, min2, min3);
    DevAqStoreMin4(rowb[xp1], min0, min1, min2, min3);
    fz[(i / fw) * 16 + i % fw] =
        ((ef.aq_erosion[0] * min0 + ef.aq_erosion[1] * min1) + ef.aq_erosion[2] * min2) + ef.aq_erosion[3] * min3;
  }
  // per-block modulations (:169-304): lane `l` of block `b` sums its column of the 8x8 block, row after row
  const uint32_t nbx = bx1 - bx0, nby = by1 - by0;
  for (uint32_t i = tid; i < nbx * nby * 8; i += nt) {
    const uint32_t l = i % 8, b = i / 8;
    const uint32_t px = (bx0 + b % nbx) * 8 + l, py = (by0 + b / nbx) * 8;
    float hf = 0.0f, gamma = 0.0f, blue = 0.0f;
    for (uint32_t dy = 0; dy < 8; dy++) {
      const size_t at = static_cast<size_t>(py + dy) * xsize + px;
      const float vy = P[1][at], vx = P[0][at], vb = P[2][at];
      const float right = l == 7 ? 0.0f : fminf(0.0206f, fabsf(vy - P[1][at + 1]));
      const float down = dy == 7 ? 0.0f : fminf(0.0206f, fabsf(vy - P[1][at + xsize]));
      hf = hf + right;
      hf = hf + down;
      const float iny = vy + 0.16f;
      gamma = gamma + DevAqRatioOfDerivatives<true>(iny - vx);
      gamma = gamma + DevAqRatioOfDerivatives<true>(iny + vx);
      const float p_y_effective = (vy + 0.084381641171960495f) + fabsf(vx);
      blue = blue + (vb > p_y_effective ? fminf(vb - p_y_effective, 0.027121074570634722f) : 0.0f);
    }
    lanes[b * 24 + l] = hf;
    lanes[b * 24 + 8 + l] = gamma;
    lanes[b * 24 + 16 + l] = blue;
  }
  CoopSync<SCOPE>();
  for (uint32_t b = tid; b < nbx * nby; b += nt) {
    const uint32_t ix = b % nbx, iy = b / nbx;
    const float* f = fz + (iy * 2) * 16 + ix * 2;
    const float eroded = ((f[0] + f[1]) + f[16]) + f[17];
    // ComputeMask (:84-108)
    const float kOffset3 = 3.7179635626140772f, kOffset4 = 0.25f * kOffset3;
    const float v1 = fmaxf(eroded * 0.80061762862741759f, 1e-3f);
    const float v2 = 1.0f / (v1 + 302.59587815579727f);
    const float v3 = 1.0f / fmaf(v1, v1, kOffset3);
    const float v4 = 1.0f / fmaf(v1, v1, kOffset4);
    float out_val = -0.7647f + fmaf(9.4708735624378946f, v4, fmaf(17.35036561631863f, v2, 6.7943250517376494f * v3));
    const float* l = lanes + b * 24;
    float s = ((l[0] + l[4]) + (l[2] + l[6])) + ((l[1] + l[5]) + (l[3] + l[7]));  // HfModulation
    s = s * -0.38f;
    s = s + 0.42f;
    out_val = s + out_val;
    const float overall = (((l[8] + l[12]) + (l[10] + l[14])) + ((l[9] + l[13]) + (l[11] + l[15]))) * (0.5f / 64);  // GammaModulation
    out_val = fmaf(0.1005613337192697f, DevFastLog2f(overall), out_val);
    const float kLimit = 0.027121074570634722f, kMaxLimit = 15.398788439047934f;  // BlueModulation
    s = ((l[16] + l[20]) + (l[18] + l[22])) + ((l[17] + l[21]) + (l[19] + l[23]));
    if (s >= 32 * kLimit) s = 64 * kLimit - s;
    if (s >= kMaxLimit * kLimit) s = kMaxLimit * kLimit;
    s = s * 0.14207000358439159f;
    out_val = s + out_val;
    E.farena[ef.quant_field + static_cast<size_t>(by0 + iy) * W + bx0 + ix] = DevFastPow2f(out_val * 1.442695041f) * ef.aq_mul + ef.aq_add;
  }
  CoopSync<SCOPE>();
}

// AcStrategy of one 64x64-pixel tile (8x8 blocks at block (x0, y0)): greedy raster scan, large smooth blocks first
// (serial: a choice depends on which blocks are still free). Every candidate is aligned to its own size and at most
// 64x64, so it lies inside one tile: the tiles of a frame are independent (one thread each) and the result is that of
// a raster scan over the whole group. acs must be 0xFF on entry.
JXLB_HD void DevEncStrategyTile(const DevEPools& E, const DevEFrame& ef, uint32_t x0, uint32_t y0) {
  const uint32_t W = ef.xblocks, H = ef.yblocks, PW = W * 8;
  const uint32_t xs = W - x0 < 8 ? W - x0 : 8, ys = H - y0 < 8 ? H - y0 : 8;
  uint8_t* acs = E.barena + ef.acs;
  const float* yplane = E.farena + ef.xyb[1];
  const int kCands[5] = {18, 5, 4, 6, 7};  // 64x64, 32x32, 16x16, 16x8, 8x16
  const double kThresh[5] = {2e-6, 1e-5, 6e-5, 1.5e-4, 1.5e-4};
  for (uint32_t iy = 0; iy < ys; iy++) {
    for (uint32_t ix = 0; ix < xs; ix++) {
      const uint32_t bx = x0 + ix, by = y0 + iy;
      const size_t pos = static_cast<size_t>(by) * W + bx;
      if (acs[pos] != 0xFF) continue;
      uint32_t s = 0, scx = 1, scy = 1;
      if (ef.strategy_mode == 2) {
        for (int i = 0; i < 5; i++) {
          const StrategyInfo si = UnpackStrategyInfo(E.upool[E.sinfo_off + kCands[i]]);
          const uint32_t cx = si.cx, cy = si.cy;
          if (bx % cx || by % cy) continue;
          if (bx + cx > W || by + cy > H) continue;
          if (bx / 32 != (bx + cx - 1) / 32 || by / 32 != (by + cy - 1) / 32) continue;
          bool free_cells = true;
          for (uint32_t y = 0; y < cy && free_cells; y++)
            for (uint32_t x = 0; x < cx; x++)
              if (acs[pos + static_cast<size_t>(y) * W + x] != 0xFF) {
                free_cells = false;
                break;
              }
          if (!free_cells) continue;
          double sum = 0, sum2 = 0;
          const uint32_t n = cx * cy * 64;
          for (uint32_t y = by * 8; y < (by + cy) * 8; y++)
            for (uint32_t x = bx * 8; x < (bx + cx) * 8; x++) {
              const double v = yplane[static_cast<size_t>(y) * PW + x];
              sum += v;
              sum2 += v * v;
            }
          const double var = sum2 / n - (sum / n) * (sum / n);
          if (var < kThresh[i] * ef.distance) {
            s = kCands[i];
            scx = cx;
            scy = cy;
            break;
          }
        }
      }
      for (uint32_t y = 0; y < scy; y++)
        for (uint32_t x = 0; x < scx; x++)
          acs[pos + static_cast<size_t>(y) * W + x] = static_cast<uint8_t>((s << 1) | ((x | y) == 0 ? 1 : 0));
      int32_t rq = 16;
      if (ef.adaptive) {
        // AdjustQuantField (lib/jxl/enc_adaptive_quantization.cc:1203-1253) + Quantizer::SetQuantFieldRect
        // (lib/jxl/quantizer.cc:71-82) for this varblock
        const float* qf = E.farena + ef.quant_field + pos;
        float max = qf[0], mean = 0.0f;
        for (uint32_t y = 0; y < scy; y++)
          for (uint32_t x = 0; x < scx; x++) {
            const float v = qf[static_cast<size_t>(y) * W + x];
            mean = mean + v;
            max = v > max ? v : max;
          }
        mean = mean / static_cast<float>(scy * scx);
        if (scy * scx >= 4) {
          max = max * ef.aq_mixer;
          max = max + (1.0f - ef.aq_mixer) * mean;
        }
        const float val = fmaxf(1.0f, fminf(max * ef.inv_global_scale + 0.5f, 256.0f));
        rq = static_cast<int32_t>(val);
      }
      E.barena[ef.raw_quant + pos] = static_cast<uint8_t>(rq - 1);
    }
  }
}

// The tiles of one 256x256 group (host emulation of the kernel's grid).
JXLB_HD void DevEncStrategyGroup(const DevEPools& E, const DevEFrame& ef, uint32_t g) {
  const uint32_t x0 = (g % ef.xgroups) * 32, y0 = (g / ef.xgroups) * 32;
  for (uint32_t ty = 0; ty < 4; ty++)
    for (uint32_t tx = 0; tx < 4; tx++)
      if (x0 + tx * 8 < ef.xblocks && y0 + ty * 8 < ef.yblocks) DevEncStrategyTile(E, ef, x0 + tx * 8, y0 + ty * 8);
}

// Numbers the varblocks of DC group g in raster order of their top-left blocks.
JXLB_HD void DevEncNumberBlocks(const DevEPools& E, const DevEFrame& ef, uint32_t g) {
  const uint32_t W = ef.xblocks, H = ef.yblocks;
  const uint32_t x0 = (g % ef.xdcgroups) * 256, y0 = (g / ef.xdcgroups) * 256;
  const uint32_t xs = W - x0 < 256 ? W - x0 : 256, ys = H - y0 < 256 ? H - y0 : 256;
  const uint8_t* acs = E.barena + ef.acs;
  int32_t* first_index = E.iarena + ef.first_index;
  int32_t* block_of_num = E.iarena + ef.block_of_num + static_cast<size_t>(g) * 65536;
  uint32_t num = 0;
  for (uint32_t y = 0; y < ys; y++)
    for (uint32_t x = 0; x < xs; x++) {
      const size_t pos = static_cast<size_t>(y0 + y) * W + x0 + x;
      if (!(acs[pos] & 1)) continue;
      first_index[pos] = static_cast<int32_t>(num);
      block_of_num[num] = static_cast<int32_t>(pos);
      num++;
    }
  E.iarena[ef.dcg_count + g] = static_cast<int32_t>(num);
}

JXLB_HD void DevEncDcBlock(const DevEPools& E, const DevEFrame& ef, uint32_t bx, uint32_t by) {
  const uint32_t W = ef.xblocks, PW = W * 8;
  float mean[3];
  for (int c = 0; c < 3; c++) {
    const float* p = E.farena + ef.xyb[c] + static_cast<size_t>(by) * 8 * PW + bx * 8;
    double s = 0;
    for (int y = 0; y < 8; y++)
      for (int x = 0; x < 8; x++) s += p[static_cast<size_t>(y) * PW + x];
    mean[c] = static_cast<float>(s / 64);
  }
  const int32_t qy = DevRoundToInt(mean[1] / ef.mul_dc[1]);
  const float ry = static_cast<float>(qy) * ef.mul_dc[1];
  const int32_t qx = DevRoundToInt((mean[0] - 0.0f * ry) / ef.mul_dc[0]);
  const int32_t qb = DevRoundToInt((mean[2] - 1.0f * ry) / ef.mul_dc[2]);
  const size_t pos = static_cast<size_t>(by) * W + bx;
  E.iarena[ef.dcq[0] + pos] = qy;
  E.iarena[ef.dcq[1] + pos] = qx;
  E.iarena[ef.dcq[2] + pos] = qb;
}

// AdjustQuantBlockAC (lib/jxl/enc_group.cc:92-316) for channel c of one varblock: a serial pass over the coefficients
// in layout order (the running sums are float additions in that order), then the quant / dead-zone decisions. `coef`:
// the channel's float coefficients in the varblock's footprint (row stride PW, C coefficients per footprint row).
// xsize >= ysize: covered blocks in coefficient-layout order. The inverse matrix is 1 / dm (as for the CfL fit).
JXLB_HD void DevEncAdjustQuantBlockAC(float scale, uint32_t c, float qm_multiplier, uint32_t quant_kind, uint32_t xsize, uint32_t ysize,
                                      float* thresholds, const float* coef, uint32_t PW, uint32_t C, const float* dm, int32_t* quant) {
  const uint32_t kPartialBlockKinds = (1u << 1) | (1u << 2) | (1u << 3) | (1u << 12) | (1u << 13) | (1u << 14) | (1u << 15) |
                                      (1u << 16) | (1u << 17);
  if ((1u << quant_kind) & kPartialBlockKinds) return;
  const float qac = scale * static_cast<float>(*quant);
  if (xsize > 1 || ysize > 1) {
    for (int i = 0; i < 4; ++i) {
      thresholds[i] = thresholds[i] - fminf(fmaxf(0.003f * static_cast<float>(xsize) * static_cast<float>(ysize), 0.f), 0.08f);
      if (static_cast<double>(thresholds[i]) < 0.54) thresholds[i] = static_cast<float>(0.54);
    }
  }
  float sum_of_highest_freq_row_and_column = 0, sum_of_error = 0, sum_of_vals = 0;
  float hfNonZeros[4] = {0, 0, 0, 0}, hfMaxError[4] = {0, 0, 0, 0};
  for (uint32_t y = 0; y < ysize * 8; y++) {
    for (uint32_t x = 0; x < xsize * 8; x++) {
      const uint32_t pos = y * 8 * xsize + x;
      if (x < xsize && y < ysize) continue;
      const uint32_t hfix = (y >= ysize * 8 / 2 ? 2u : 0u) + (x >= xsize * 8 / 2 ? 1u : 0u);
      const float in = coef[static_cast<size_t>(pos / C) * PW + pos % C];
      const float val = in * (((1.0f / dm[pos]) * qac) * qm_multiplier);
      const float v = (fabsf(val) < thresholds[hfix]) ? 0.0f : rintf(val);
      const float error = fabsf(val - v);
      sum_of_error = sum_of_error + error;
      sum_of_vals = sum_of_vals + fabsf(v);
      if (c == 1 && v == 0) {
        if (hfMaxError[hfix] < error) hfMaxError[hfix] = error;
      }
      if (v != 0.0f) {
        hfNonZeros[hfix] = hfNonZeros[hfix] + fabsf(v);
        const bool in_corner = y >= 7 * ysize && x >= 7 * xsize;
        const bool on_border = y == ysize * 8 - 1 || x == xsize * 8 - 1;
        const bool in_larger_corner = x >= 4 * xsize && y >= 4 * ysize;
        if (in_corner || (on_border && in_larger_corner))
          sum_of_highest_freq_row_and_column = sum_of_highest_freq_row_and_column + fabsf(val);
      }
    }
  }
  if (c == 1 && sum_of_vals * 8 < static_cast<float>(xsize * ysize)) {
    const double kLimit = 0.46, kMul = 0.9999;
    const int32_t orig_quant = *quant;
    int32_t new_quant = *quant;
    for (int i = 1; i < 4; ++i) {
      if (hfNonZeros[i] == 0.0f && static_cast<double>(hfMaxError[i]) > kLimit) {
        new_quant = orig_quant + 1;
        break;
      }
    }
    *quant = new_quant;
    if (hfNonZeros[3] == 0.0f && static_cast<double>(hfMaxError[3]) > kLimit) {
      thresholds[3] = static_cast<float>(kMul * static_cast<double>(hfMaxError[3]) * new_quant / orig_quant);
    } else if ((hfNonZeros[1] == 0.0f && static_cast<double>(hfMaxError[1]) > kLimit) ||
               (hfNonZeros[2] == 0.0f && static_cast<double>(hfMaxError[2]) > kLimit)) {
      thresholds[1] = static_cast<float>(kMul * static_cast<double>(fmaxf(hfMaxError[1], hfMaxError[2])) * new_quant / orig_quant);
      thresholds[2] = thresholds[1];
    } else if (hfNonZeros[0] == 0.0f && static_cast<double>(hfMaxError[0]) > kLimit) {
      thresholds[0] = static_cast<float>(kMul * static_cast<double>(hfMaxError[0]) * new_quant / orig_quant);
    }
  }
  {
    const float all = hfNonZeros[0] + hfNonZeros[1] + hfNonZeros[2] + hfNonZeros[3] + 1;
    const float mul = c == 0 ? 70.0f : (c == 1 ? 30.0f : 60.0f);
    if (mul * sum_of_highest_freq_row_and_column >= all) {
      *quant = static_cast<int32_t>(static_cast<float>(*quant) + mul * sum_of_highest_freq_row_and_column / all);
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
    const double kMul1[4][3] = {{0.22080615753848404, 0.45797479824262011, 0.29859235095977965},
                                {0.70109486510286834, 0.16185281305512639, 0.14387691730035473},
                                {0.114985964456218638, 0.44656840441027695, 0.10587658215149048},
                                {0.46849665264409396, 0.41239077937781954, 0.088667407767185444}};
    const double kMul2[4][3] = {{0.27450281941822197, 1.1255766549984996, 0.98950459134128388},
                                {0.4652168675598285, 0.40945807983455818, 0.36581899811751367},
                                {0.28034972424715715, 0.9182653201929738, 1.5581531543057416},
                                {0.26873118114033728, 0.68863712390392484, 1.2082185408666786}};
    const double kQuantNormalizer = 2.2942708343284721;
    sum_of_error = static_cast<float>(static_cast<double>(sum_of_error) * kQuantNormalizer);
    sum_of_vals = static_cast<float>(static_cast<double>(sum_of_vals) * kQuantNormalizer);
    if (quant_kind >= 4) {
      int ix = 3;
      if (quant_kind == 10 || quant_kind == 11) {
        ix = 1;
      } else if (quant_kind == 4) {
        ix = 0;
      } else if (quant_kind == 5) {
        ix = 2;
      }
      const double limit = kMul1[ix][c] * xsize * ysize * 8 * 8 + kMul2[ix][c] * static_cast<double>(sum_of_vals);
      int step = static_cast<int>(static_cast<double>(sum_of_error) / limit);
      if (step >= 2) step = 2;
      if (step < 0) step = 0;
      if (static_cast<double>(sum_of_error) > limit) {
        *quant += step;
        if (*quant >= 256) *quant = 256 - 1;
      }
    }
  }
  {
    const int32_t div = static_cast<int32_t>(xsize * ysize);
    int32_t activity = (static_cast<int32_t>(hfNonZeros[0]) + div / 2) / div;
    const int32_t orig_qp_limit = 4 > *quant / 2 ? 4 : *quant / 2;
    for (int i = 1; i < 4; ++i) {
      const int32_t a = (static_cast<int32_t>(hfNonZeros[i]) + div / 2) / div;
      activity = activity < a ? activity : a;
    }
    if (activity >= 15) activity = 15;
    int32_t qp = *quant - activity;
    if (c == 1) {
      for (int i = 1; i < 4; ++i) thresholds[i] = static_cast<float>(static_cast<double>(thresholds[i]) + 0.01 * activity);
    }
    if (qp < orig_qp_limit) qp = orig_qp_limit;
    *quant = qp;
  }
}

// Channel c of the varblock at (bx, by): AdjustQuantBlockAC on the float coefficients k_enc_coeffs<0> left in the xyb_raw
// planes. One thread per (varblock, channel): the statistics pass is serial per channel, the varblocks are not. Results:
// the Y thresholds (c == 1) and the maximum of the three channels' adjusted quants (the int arena is zeroed per batch).
JXLB_HD void DevEncAdjustVarblockChannel(const DevEPools& E, const DevEFrame& ef, uint32_t bx, uint32_t by, uint32_t c) {
  const uint32_t W = ef.xblocks, PW = W * 8;
  const size_t pos = static_cast<size_t>(by) * W + bx;
  const uint8_t a = E.barena[ef.acs + pos];
  if (!ef.adaptive || !(a & 1) || a == 0xFF) return;
  const uint32_t s = a >> 1;
  const StrategyInfo si = UnpackStrategyInfo(E.upool[E.sinfo_off + s]);
  const uint32_t Rb = si.cy, Cb = si.cx, C = 8 * Cb, N = 64 * Rb * Cb;
  const uint32_t lcx = Cb > Rb ? Cb : Rb, lcy = Cb > Rb ? Rb : Cb;
  const size_t origin = static_cast<size_t>(by) * 8 * PW + static_cast<size_t>(bx) * 8;
  const float* dm = E.fpool + E.table_off[si.table];
  const float scale = ef.cfl_scale128 * (1.0f / 128.0f);  // Quantizer::Scale() (exact: a power-of-two factor)
  float thres[4] = {0.58f, 0.64f, 0.64f, 0.64f};
  int32_t quant = static_cast<int32_t>(E.barena[ef.raw_quant + pos]) + 1;
  const float qm_mul = c == 0 ? ef.x_qm_mul : (c == 1 ? 1.0f : ef.b_qm_mul);
  DevEncAdjustQuantBlockAC(scale, c, qm_mul, s, lcx, lcy, thres, E.farena + ef.xyb_raw[c] + origin, PW, C, dm + c * N, &quant);
  if (c == 1) {
    const size_t stride = (static_cast<size_t>(W) * ef.yblocks + 15) & ~size_t{15};
    for (uint32_t k = 0; k < 4; k++) E.farena[ef.adj_thres + k * stride + pos] = thres[k];
  }
#if defined(__CUDA_ARCH__)
  atomicMax(E.iarena + ef.adj_quant + pos, quant);
#else
  if (E.iarena[ef.adj_quant + pos] < quant) E.iarena[ef.adj_quant + pos] = quant;
#endif
}

// One varblock (plain DCT strategies), in two steps around the chroma-from-luma fit. MODE 0: forward transform, the
// float coefficients go to the xyb_raw planes (layout order inside the varblock's footprint); `buf`: 4 * 64 * covered
// floats. MODE 1: quantisation of those coefficients (Y round trip, chroma relative to decoded Y with the tile's factors).
template <int SCOPE, int MODE>
JXLB_HD void DevEncVarblock(const DevEPools& E, const DevEFrame& ef, uint32_t bx, uint32_t by, uint32_t s, float* buf,
                            uint32_t tid, uint32_t nt) {
  const StrategyInfo si = UnpackStrategyInfo(E.upool[E.sinfo_off + s]);
  const uint32_t Rb = si.cy, Cb = si.cx, R = 8 * Rb, C = 8 * Cb, N = R * C;
  const uint32_t W = ef.xblocks, PW = W * 8;
  const float* wc = E.fpool + E.wc_off;
  const size_t origin = static_cast<size_t>(by) * 8 * PW + static_cast<size_t>(bx) * 8;
  float* dct[3] = {E.farena + ef.xyb_raw[0] + origin, E.farena + ef.xyb_raw[1] + origin, E.farena + ef.xyb_raw[2] + origin};
  if (MODE == 0) {
  float* ch[3] = {buf, buf + N, buf + 2 * static_cast<size_t>(N)};
  float* scratch = buf + 3 * static_cast<size_t>(N);
  const float* F[3];
  for (uint32_t c = 0; c < 3; c++) {
    const float* px = E.farena + ef.xyb[c] + origin;
    for (uint32_t i = tid; i < N; i += nt) ch[c][i] = px[static_cast<size_t>(i / C) * PW + i % C];
    CoopSync<SCOPE>();
    float* r1 = CoopDCT<SCOPE>(R, C, 1, C, ch[c], scratch, wc, tid, nt);            // along y for every column
    float* r2 = CoopDCT<SCOPE>(C, R, C, 1, r1, r1 == ch[c] ? scratch : ch[c], wc, tid, nt);  // along x for every row
    if (r2 != ch[c]) {
      for (uint32_t i = tid; i < N; i += nt) ch[c][i] = r2[i];
      CoopSync<SCOPE>();
    }
    F[c] = ch[c];  // F[yfreq * C + xfreq]
  }
  for (uint32_t k = tid; k < N; k += nt) {
    const uint32_t fi = R < C ? k : (k % R) * C + k / R;  // coefficient layout: min(R, C) rows x max(R, C) columns
    const size_t at = static_cast<size_t>(k / C) * PW + k % C;
    dct[0][at] = F[0][fi];
    dct[1][at] = F[1][fi];
    dct[2][at] = F[2][fi];
  }
  CoopSync<SCOPE>();
  return;
  }
  const size_t tile = static_cast<size_t>(by / 8) * ef.cmw + bx / 8;
  const float x_cc = 0.0f + static_cast<float>(reinterpret_cast<const int8_t*>(E.barena + ef.ytox)[tile]) * (1.0f / 84);
  const float b_cc = 1.0f + static_cast<float>(reinterpret_cast<const int8_t*>(E.barena + ef.ytob)[tile]) * (1.0f / 84);
  const float* dm = E.fpool + E.table_off[si.table];
  const uint32_t lcx = Cb > Rb ? Cb : Rb, lcy = Cb > Rb ? Rb : Cb;
  int32_t* out[3] = {E.iarena + ef.coef[0] + origin, E.iarena + ef.coef[1] + origin, E.iarena + ef.coef[2] + origin};
  uint8_t* raw_quant = E.barena + ef.raw_quant + static_cast<size_t>(by) * W + bx;
  if (ef.adaptive) {
    // libjxl's effort-7 quantisation (QuantizeRoundtripYBlockAC + ComputeCoefficients, lib/jxl/enc_group.cc:319-368,
    // :455-491): the per-channel statistics pass ran in k_enc_adjust (DevEncAdjustVarblockChannel); here every
    // coefficient on its own: Y with its dead-zone thresholds, the round trip, chroma relative to decoded Y.
    const float scale = ef.cfl_scale128 * (1.0f / 128.0f);  // Quantizer::Scale() (exact: a power-of-two factor)
    const size_t bpos = static_cast<size_t>(by) * W + bx;
    const size_t tstride = (static_cast<size_t>(W) * ef.yblocks + 15) & ~size_t{15};
    const float thres_y[4] = {E.farena[ef.adj_thres + bpos], E.farena[ef.adj_thres + tstride + bpos],
                              E.farena[ef.adj_thres + 2 * tstride + bpos], E.farena[ef.adj_thres + 3 * tstride + bpos]};
    const int32_t quant = E.iarena[ef.adj_quant + bpos];  // (k_enc_adjust: max over the three channels)
    float thres_c[4] = {0.58f, 0.62f, 0.62f, 0.62f};
    if (lcx * lcy >= 4) {
      for (int i = 0; i < 4; ++i) {
        thres_c[i] = thres_c[i] - 0.00744f * static_cast<float>(lcx) * static_cast<float>(lcy);
        if (static_cast<double>(thres_c[i]) < 0.5) thres_c[i] = 0.5f;
      }
    }
    const float qac = scale * static_cast<float>(quant);
    const float inv_qac = ef.inv_global_scale / static_cast<float>(quant);
    const float quantv[3] = {qac * ef.x_qm_mul, qac * 1.0f, qac * ef.b_qm_mul};
    for (uint32_t k = tid; k < N; k += nt) {
      const size_t at = static_cast<size_t>(k / C) * PW + k % C;
      const uint32_t ly = k / (lcx * 8), lx = k % (lcx * 8);
      const uint32_t hfix = (ly >= lcy * 8 / 2 ? 2u : 0u) + (lx >= lcx * 8 / 2 ? 1u : 0u);
      const float val_y = ((1.0f / dm[N + k]) * quantv[1]) * dct[1][at];
      int32_t qy = fabsf(val_y) >= thres_y[hfix] ? static_cast<int32_t>(rintf(val_y)) : 0;
      const float dq_y = (DevAdjustQuantBias(1, qy, ef.biases) * dm[N + k]) * inv_qac;
      const float in_x = fmaf(-x_cc, dq_y, dct[0][at]);
      const float in_b = fmaf(-b_cc, dq_y, dct[2][at]);
      const float val_x = ((1.0f / dm[k]) * quantv[0]) * in_x;
      const float val_b = ((1.0f / dm[2 * N + k]) * quantv[2]) * in_b;
      int32_t qx = fabsf(val_x) >= thres_c[hfix] ? static_cast<int32_t>(rintf(val_x)) : 0;
      int32_t qb = fabsf(val_b) >= thres_c[hfix] ? static_cast<int32_t>(rintf(val_b)) : 0;
      if (ly < lcy && lx < lcx) qx = qy = qb = 0;  // the lowest frequencies come from the DC image
      out[0][at] = qx;
      out[1][at] = qy;
      out[2][at] = qb;
    }
    if (tid == 0) *raw_quant = static_cast<uint8_t>(quant - 1);
    CoopSync<SCOPE>();
    return;
  }
  const float sd_base = ef.inv_global_scale / static_cast<float>(*raw_quant + 1);
  const float sd0 = sd_base * ef.x_dm, sd1 = sd_base, sd2 = sd_base * ef.b_dm;
  for (uint32_t k = tid; k < N; k += nt) {
    const size_t from = static_cast<size_t>(k / C) * PW + k % C;
    const float y_mul = dm[N + k] * sd1;
    int32_t qy = DevRoundToInt(dct[1][from] / y_mul);
    const float dq_y = DevAdjustQuantBias(1, qy, ef.biases) * y_mul;
    int32_t qx = DevRoundToInt((dct[0][from] - x_cc * dq_y) / (dm[k] * sd0));
    int32_t qb = DevRoundToInt((dct[2][from] - b_cc * dq_y) / (dm[2 * N + k] * sd2));
    const uint32_t ly = k / (lcx * 8), lx = k % (lcx * 8);
    if (ly < lcy && lx < lcx) qx = qy = qb = 0;  // the lowest frequencies come from the DC image
    const size_t at = static_cast<size_t>(k / C) * PW + k % C;
    out[0][at] = qx;
    out[1][at] = qy;
    out[2][at] = qb;
  }
  CoopSync<SCOPE>();
}

// Chroma-from-luma factors of one 64x64 tile: CfLHeuristics::ComputeTile + FindBestMultiplier (fast == false),
// lib/jxl/enc_chroma_from_luma.cc:41-175, :196-342, over the float coefficients MODE 0 left behind. `vals`: 4 * 4096
// floats (luma and chroma coefficients weighted for X and for B), `red`: 64 floats. The three sums of a Newton step
// run as 8 lanes each (element i in lane i % 8, added in increasing i), combined in Highway's AVX2 SumOfLanes order --
// 2 channels x 3 sums x 8 lanes = 48 independent chains for the threads of the CTA.
template <int SCOPE>
JXLB_HD void DevEncCflTile(const DevEPools& E, const DevEFrame& ef, uint32_t tx, uint32_t ty, uint32_t tid, uint32_t nt,
                           float* vals, float* red) {
  int8_t* ytox = reinterpret_cast<int8_t*>(E.barena + ef.ytox);
  int8_t* ytob = reinterpret_cast<int8_t*>(E.barena + ef.ytob);
  const size_t tile = static_cast<size_t>(ty) * ef.cmw + tx;
  if (!ef.cfl) {
    if (tid == 0) ytox[tile] = ytob[tile] = 0;
    return;
  }
  const uint32_t W = ef.xblocks, H = ef.yblocks, PW = W * 8;
  const uint32_t x0 = tx * 8, y0 = ty * 8, x1 = x0 + 8 < W ? x0 + 8 : W, y1 = y0 + 8 < H ? y0 + 8 : H;
  const uint8_t* acs = E.barena + ef.acs;
  float* v_m[2] = {vals, vals + 8192};          // luma weighted for X, for B
  float* v_s[2] = {vals + 4096, vals + 12288};  // X, B
  uint32_t num_ac = 0;
  for (uint32_t by = y0; by < y1; by++)
    for (uint32_t bx = x0; bx < x1; bx++) {
      const uint8_t a = acs[static_cast<size_t>(by) * W + bx];
      if (!(a & 1)) continue;
      const StrategyInfo si = UnpackStrategyInfo(E.upool[E.sinfo_off + (a >> 1)]);
      if (si.cx + x0 > x1 || si.cy + y0 > y1) continue;  // (the reference's test: blocks larger than the tile)
      const uint32_t N = 64u * si.cx * si.cy, C = si.cx * 8u;
      const uint32_t lcx = si.cx > si.cy ? si.cx : si.cy, lcy = si.cx > si.cy ? si.cy : si.cx;
      const float* dm = E.fpool + E.table_off[si.table];
      const size_t origin = static_cast<size_t>(by) * 8 * PW + static_cast<size_t>(bx) * 8;
      const float cfl_q = ef.cfl_scale128 * static_cast<float>(E.barena[ef.raw_quant + static_cast<size_t>(by) * W + bx] + 1);
      for (uint32_t k = tid; k < N; k += nt) {
        const size_t at = origin + static_cast<size_t>(k / C) * PW + k % C;
        const bool llf = k / (lcx * 8) < lcy && k % (lcx * 8) < lcx;
        const float cy = llf ? 0.0f : E.farena[ef.xyb_raw[1] + at];
        const float cxv = llf ? 0.0f : E.farena[ef.xyb_raw[0] + at];
        const float cb = llf ? 0.0f : E.farena[ef.xyb_raw[2] + at];
        const float qqm_x = cfl_q * (1.0f / dm[k]), qqm_b = cfl_q * (1.0f / dm[2 * N + k]);
        v_m[0][num_ac + k] = cy * qqm_x;
        v_s[0][num_ac + k] = cxv * qqm_x;
        v_m[1][num_ac + k] = cy * qqm_b;
        v_s[1][num_ac + k] = cb * qqm_b;
      }
      num_ac += N;
    }
  // red[0..1]: x per channel, red[2..3]: done flags, red[8 + 24 * ch + 8 * which + lane]: lane sums
  for (uint32_t i = tid; i < 4; i += nt) red[i] = 0.0f;
  CoopSync<SCOPE>();
  const float kInvColorFactor = 1.0f / 84, kThres = 100.0f, eps = 100, kClamp = 20.0f, distance_mul = 1e-9f;
  const float coeffx2 = (1.0f / 3) * 2.0f;
  for (uint32_t iter = 0; iter < 20 && num_ac != 0; iter++) {
    for (uint32_t chain = tid; chain < 48; chain += nt) {
      const uint32_t ch = chain / 24, which = (chain / 8) % 3, lane = chain % 8;
      if (red[2 + ch] != 0.0f) continue;
      const float base = ch == 0 ? 0.0f : 1.0f;
      const float x = red[ch];
      const float xw = which == 0 ? x : (which == 1 ? x + eps : x - eps);
      const float* vm = v_m[ch];
      const float* vs = v_s[ch];
      float acc = 0.0f;
      for (uint32_t i = lane; i < num_ac; i += 8) {
        const float a = kInvColorFactor * vm[i];
        const float b = base * vm[i] - vs[i];
        const float v = fmaf(a, x, b), vw = fmaf(a, xw, b);
        const float acoeffx2 = coeffx2 * a;
        float d = acoeffx2 * (fabsf(vw) + 1.0f);
        d = vw < 0.0f ? 0.0f - d : d;
        acc = acc + (fabsf(v) >= kThres ? 0.0f : d);
      }
      red[8 + chain] = acc;
    }
    CoopSync<SCOPE>();
    for (uint32_t ch = tid; ch < 2; ch += nt) {
      if (red[2 + ch] != 0.0f) continue;
      const float x = red[ch];
      const float* l = red + 8 + 24 * ch;
      const float sum0 = ((l[0] + l[4]) + (l[2] + l[6])) + ((l[1] + l[5]) + (l[3] + l[7]));
      const float sum1 = ((l[8] + l[12]) + (l[10] + l[14])) + ((l[9] + l[13]) + (l[11] + l[15]));
      const float sum2 = ((l[16] + l[20]) + (l[18] + l[22])) + ((l[17] + l[21]) + (l[19] + l[23]));
      const float df = 2 * distance_mul * num_ac * x + sum0;
      const float dfpeps = 2 * distance_mul * num_ac * (x + eps) + sum1;
      const float dfmeps = 2 * distance_mul * num_ac * (x - eps) + sum2;
      const float ddf = (dfpeps - dfmeps) / (2 * eps);
      const float step = df / (ddf + 0.85f);
      red[ch] = x - (step > kClamp ? kClamp : (step < -kClamp ? -kClamp : step));
      if (fabsf(step) < 3e-3f) red[2 + ch] = 1.0f;
    }
    CoopSync<SCOPE>();
    if (red[2] != 0.0f && red[3] != 0.0f) break;  // (uniform: both flags are read after the barrier)
  }
  for (uint32_t ch = tid; ch < 2; ch += nt) {
    float x = num_ac == 0 ? 0.0f : red[ch];
    const float towards_zero = 2.6f;
    if (x >= towards_zero) {
      x -= towards_zero;
    } else if (x <= -towards_zero) {
      x += towards_zero;
    } else {
      x = 0;
    }
    float r = roundf(x);
    r = r < -128.0f ? -128.0f : (r > 127.0f ? 127.0f : r);
    (ch == 0 ? ytox : ytob)[tile] = static_cast<int8_t>(static_cast<int32_t>(r));
  }
}

JXLB_HD uint32_t DevPackSigned(int32_t v) { return (static_cast<uint32_t>(v) << 1) ^ (v < 0 ? 0xFFFFFFFFu : 0u); }

// HybridUintConfig(4, 2, 0)
JXLB_HD void DevHybrid420(uint32_t v, uint32_t* token, uint32_t* nbits, uint32_t* bits) {
  if (v < 16) {
    *token = v;
    *nbits = 0;
    *bits = 0;
    return;
  }
  const uint32_t n = DevFloorLog2(v), m = v - (1u << n);
  *token = 16 + ((n - 4) << 2) + (m >> (n - 2));
  *nbits = n - 2;
  *bits = m & ((1u << (n - 2)) - 1);
}

JXLB_HD void DevCountToken(uint32_t* hist, uint32_t cluster, uint32_t value) {
  uint32_t token, nbits, bits;
  DevHybrid420(value, &token, &nbits, &bits);
#if defined(__CUDA_ARCH__)
  atomicAdd(hist + cluster * 256 + token, 1u);
#else
  hist[cluster * 256 + token]++;
#endif
}

// ---- coefficient-order statistics (ComputeUsedOrders + the counting loop of ComputeCoeffOrder,
// lib/jxl/enc_coeff_order.cc:47-160). Step 1, one thread per group: varblocks of the group and the orders they use.
JXLB_HD void DevEncGroupOrders(const DevEPools& E, const DevEFrame& ef, uint32_t g) {
  const uint32_t W = ef.xblocks, H = ef.yblocks;
  const uint32_t x0 = (g % ef.xgroups) * 32, y0 = (g / ef.xgroups) * 32;
  const uint32_t xs = W - x0 < 32 ? W - x0 : 32, ys = H - y0 < 32 ? H - y0 : 32;
  const uint8_t* acs = E.barena + ef.acs;
  uint32_t count = 0, mask = 0;
  for (uint32_t by = 0; by < ys; by++)
    for (uint32_t bx = 0; bx < xs; bx++) {
      const uint8_t a = acs[static_cast<size_t>(y0 + by) * W + x0 + bx];
      if (!(a & 1)) continue;
      count++;
      mask |= 1u << UnpackStrategyInfo(E.upool[E.sinfo_off + (a >> 1)]).order;
    }
  E.iarena[ef.group_first + g] = static_cast<int32_t>(count);
#if defined(__CUDA_ARCH__)
  atomicOr(reinterpret_cast<uint32_t*>(E.iarena + ef.order_mask), mask);
#else
  E.iarena[ef.order_mask] |= static_cast<int32_t>(mask);
#endif
}

// Step 2, one CTA per group (`nt` cooperating threads, `local`: kCustomOrderCounters + 1024 words of shared memory):
// zero coefficients per (order, channel, position). libjxl walks the varblocks in group order and, when only 8x8
// DCT orders can be customised, takes every other one by a fixed xorshift128+ sequence: sample_bits[rank].
template <int SCOPE>
JXLB_HD void DevEncOrderStatsGroup(const DevEPools& E, const DevEFrame& ef, uint32_t g, uint32_t tid, uint32_t nt, uint32_t* local) {
  const uint32_t W = ef.xblocks, H = ef.yblocks, PW = W * 8;
  const uint32_t x0 = (g % ef.xgroups) * 32, y0 = (g / ef.xgroups) * 32;
  const uint32_t xs = W - x0 < 32 ? W - x0 : 32, ys = H - y0 < 32 ? H - y0 : 32;
  const uint8_t* acs = E.barena + ef.acs;
  uint32_t* counts = local;
  uint32_t* rank = local + kCustomOrderCounters;  // per cell of the group: rank of its varblock, or 0xFFFFFFFF
  for (uint32_t i = tid; i < kCustomOrderCounters; i += nt) counts[i] = 0;
  const bool half = (static_cast<uint32_t>(E.iarena[ef.order_mask]) & 0x7Fu) == 1u;
  if (tid == 0) {
    uint32_t r = 0;
    for (uint32_t q = 0; q < g; q++) r += static_cast<uint32_t>(E.iarena[ef.group_first + q]);
    for (uint32_t by = 0; by < ys; by++)
      for (uint32_t bx = 0; bx < xs; bx++) {
        const bool first = (acs[static_cast<size_t>(y0 + by) * W + x0 + bx] & 1) != 0;
        rank[by * 32 + bx] = first ? r : 0xFFFFFFFFu;
        r += first ? 1 : 0;
      }
  }
  CoopSync<SCOPE>();
  for (uint32_t cell = 0; cell < xs * ys; cell++) {  // (uniform loop: the threads split each varblock's coefficients)
    const uint32_t bx = cell % xs, by = cell / xs;
    const uint32_t r = rank[by * 32 + bx];
    if (r == 0xFFFFFFFFu) continue;
    if (half && !E.sample_bits[r]) continue;
    const size_t pos = static_cast<size_t>(y0 + by) * W + x0 + bx;
    const StrategyInfo si = UnpackStrategyInfo(E.upool[E.sinfo_off + (acs[pos] >> 1)]);
    if (si.order >= kNumCustomOrders) continue;
    const uint32_t size = 64u << si.log2_covered, C = si.cx * 8u;
    uint32_t lcx = si.cx, lcy = si.cy;  // the lowest frequencies: lcy rows of lcx in the (cx >= cy) coefficient layout
    if (lcy > lcx) {
      const uint32_t t = lcx;
      lcx = lcy;
      lcy = t;
    }
    for (uint32_t i = tid; i < 3 * size; i += nt) {
      const uint32_t c = i / size, p = i % size;
      if (p % (8 * lcx) < lcx && p / (8 * lcx) < lcy) continue;  // LLF: pinned to the front by the host
      const int32_t* coef = E.iarena + ef.coef[c] + static_cast<size_t>(y0 + by) * 8 * PW + static_cast<size_t>(x0 + bx) * 8;
      if (coef[static_cast<size_t>(p / C) * PW + p % C] == 0) {
#if defined(__CUDA_ARCH__)
        atomicAdd(counts + CustomOrderBase(si.order) + c * size + p, 1u);
#else
        counts[CustomOrderBase(si.order) + c * size + p]++;
#endif
      }
    }
  }
  CoopSync<SCOPE>();
  for (uint32_t i = tid; i < kCustomOrderCounters; i += nt) {
    if (counts[i] == 0) continue;
#if defined(__CUDA_ARCH__)
    atomicAdd(reinterpret_cast<uint32_t*>(E.iarena + ef.zero_counts) + i, counts[i]);
#else
    E.iarena[ef.zero_counts + i] += static_cast<int32_t>(counts[i]);
#endif
  }
}

// The coefficient order of (order class, channel): the frame's custom one if the host made one, else the natural one.
JXLB_HD const uint16_t* DevEncOrder(const DevEPools& E, const DevEFrame& ef, uint32_t ord, uint32_t c) {
  if (ord < kNumCustomOrders && ef.custom_order[3 * ord + c] != 0xFFFFFFFFu) return E.opool_custom + ef.custom_order[3 * ord + c];
  return E.opool + E.order_off[ord];
}

// ---- AC tokenisation in three data-parallel steps. The context of a coefficient depends on the number of
// non-zeros still to come and on whether the previous coefficient (in scan order) was zero: both follow from a
// scan of the block alone, so every (varblock, channel) is tokenised independently once the token offsets are known.
//   1. DevEncBlockStats     per (varblock, channel): non-zero count, token count, non-zero bucket of its cells
//   2. DevEncTokenOffsets   per group: prefix sum of the token counts in stream order (raster varblocks, Y X B)
//   3. DevEncBlockTokens    per (varblock, channel): the tokens at their final slots + histogram counts
// iarena planes used (xblocks * yblocks each, per channel c): blk_nz[c] (non-zero count, first blocks),
// blk_ntok[c] (token count, then token offset), blk_bucket[c] (per cell: non-zero bucket of the covering varblock).
JXLB_HD void DevEncBlockStats(const DevEPools& E, const DevEFrame& ef, uint32_t bx, uint32_t by, uint32_t c) {
  const uint32_t W = ef.xblocks, PW = W * 8;
  const size_t pos = static_cast<size_t>(by) * W + bx;
  const uint8_t a = E.barena[ef.acs + pos];
  if (!(a & 1)) return;
  const StrategyInfo si = UnpackStrategyInfo(E.upool[E.sinfo_off + (a >> 1)]);
  const uint32_t log2c = si.log2_covered, covered = 1u << log2c, size = covered * 64, C = si.cx * 8u;
  const uint16_t* order = DevEncOrder(E, ef, si.order, c);
  const int32_t* coef = E.iarena + ef.coef[c] + static_cast<size_t>(by) * 8 * PW + static_cast<size_t>(bx) * 8;
  uint32_t nz = 0, last = 0;
  for (uint32_t k = covered; k < size; k++) {
    const uint32_t p = JXLB_LDG(order + k);
    if (coef[static_cast<size_t>(p / C) * PW + p % C] != 0) {
      nz++;
      last = k;
    }
  }
  E.iarena[ef.blk_nz[c] + pos] = static_cast<int32_t>(nz);
  E.iarena[ef.blk_ntok[c] + pos] = static_cast<int32_t>(nz == 0 ? 1 : 2 + last - covered);
  const int32_t bucket = static_cast<int32_t>((nz + covered - 1) >> log2c);
  for (uint32_t y = 0; y < si.cy; y++)
    for (uint32_t x = 0; x < si.cx; x++) E.iarena[ef.blk_bucket[c] + pos + static_cast<size_t>(y) * W + x] = bucket;
}

// blk_ntok: counts -> offsets inside the group's token region; group_tokens[g] = total.
JXLB_HD void DevEncTokenOffsets(const DevEPools& E, const DevEFrame& ef, uint32_t g) {
  const uint32_t W = ef.xblocks, H = ef.yblocks;
  const uint32_t x0 = (g % ef.xgroups) * 32, y0 = (g / ef.xgroups) * 32;
  const uint32_t xs = W - x0 < 32 ? W - x0 : 32, ys = H - y0 < 32 ? H - y0 : 32;
  const uint8_t* acs = E.barena + ef.acs;
  uint32_t acc = 0;
  for (uint32_t by = 0; by < ys; by++)
    for (uint32_t bx = 0; bx < xs; bx++) {
      const size_t pos = static_cast<size_t>(y0 + by) * W + x0 + bx;
      if (!(acs[pos] & 1)) continue;
      for (uint32_t ci = 0; ci < 3; ci++) {
        const uint32_t c = ci == 0 ? 1 : (ci == 1 ? 0 : 2);
        const uint32_t n = static_cast<uint32_t>(E.iarena[ef.blk_ntok[c] + pos]);
        E.iarena[ef.blk_ntok[c] + pos] = static_cast<int32_t>(acc);
        acc += n;
      }
    }
  E.iarena[ef.group_tokens + g] = static_cast<int32_t>(acc);
}

JXLB_HD void DevEncBlockTokens(const DevEPools& E, const DevEFrame& ef, uint32_t bx, uint32_t by, uint32_t c,
                               const uint16_t* freq_ctx, const uint16_t* nnz_ctx) {
  const uint8_t kBlockCtx[39] = {0, 1, 2, 2, 3, 3, 4, 5, 6, 6, 6, 6, 6, 7, 8, 9, 9, 10, 11, 12,
                                 13, 14, 14, 14, 14, 14, 7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14};
  const uint32_t W = ef.xblocks, PW = W * 8;
  const size_t pos = static_cast<size_t>(by) * W + bx;
  const uint8_t a = E.barena[ef.acs + pos];
  if (!(a & 1)) return;
  const StrategyInfo si = UnpackStrategyInfo(E.upool[E.sinfo_off + (a >> 1)]);
  const uint32_t log2c = si.log2_covered, covered = 1u << log2c, size = covered * 64, C = si.cx * 8u, ord = si.order;
  const uint16_t* order = DevEncOrder(E, ef, ord, c);
  const int32_t* coef = E.iarena + ef.coef[c] + static_cast<size_t>(by) * 8 * PW + static_cast<size_t>(bx) * 8;
  const uint32_t g = (by / 32) * ef.xgroups + bx / 32;
  uint2* tok = E.tokens + ef.ac_tokens + static_cast<size_t>(g) * 3 * 65536 + static_cast<uint32_t>(E.iarena[ef.blk_ntok[c] + pos]);
  uint32_t* hist = reinterpret_cast<uint32_t*>(E.iarena + ef.ac_hist);
  const uint32_t num_ctxs = 15;
  const uint32_t gbx = bx % 32, gby = by % 32;  // position inside the group: prediction does not cross groups
  const int32_t* bucket_plane = E.iarena + ef.blk_bucket[c] + pos;
  uint32_t predicted;
  if (gbx == 0) {
    predicted = gby == 0 ? 32 : static_cast<uint32_t>(bucket_plane[-static_cast<ptrdiff_t>(W)]);
  } else if (gby == 0) {
    predicted = static_cast<uint32_t>(bucket_plane[-1]);
  } else {
    predicted = (static_cast<uint32_t>(bucket_plane[-static_cast<ptrdiff_t>(W)]) + static_cast<uint32_t>(bucket_plane[-1]) + 1) / 2;
  }
  const uint32_t block_ctx = kBlockCtx[(c < 2 ? c ^ 1 : 2) * kNumOrders + ord];
  uint32_t nz = static_cast<uint32_t>(E.iarena[ef.blk_nz[c] + pos]);
  uint32_t bucket = predicted >= 64 ? 64 : predicted;
  bucket = bucket < 8 ? bucket : 4 + bucket / 2;
  const uint32_t nz_cluster = E.ac_cluster_of[bucket * num_ctxs + block_ctx];
  uint32_t ntok = 0;
  tok[ntok++] = make_uint2(nz_cluster, nz);  // tokens carry the cluster of their context
  DevCountToken(hist, nz_cluster, nz);
  const uint32_t histo_offset = num_ctxs * 37 + 458 * block_ctx;
  uint32_t prev = nz > size / 16 ? 0 : 1;
  for (uint32_t k = covered; k < size && nz != 0; k++) {
    const uint32_t nzl = (nz + covered - 1) >> log2c;
    const uint32_t ctx = histo_offset + (nnz_ctx[nzl] + freq_ctx[k >> log2c]) * 2 + prev;
    const uint32_t p = JXLB_LDG(order + k);
    const uint32_t u = DevPackSigned(coef[static_cast<size_t>(p / C) * PW + p % C]);
    const uint32_t cluster = E.ac_cluster_of[ctx];
    tok[ntok++] = make_uint2(cluster, u);
    DevCountToken(hist, cluster, u);
    prev = u != 0;
    nz -= prev;
  }
}

// ---- Modular sub-streams (DC, AC metadata) under the fixed global tree: every sample's context and prediction
// depend only on already-known neighbours, so the tokens are produced in parallel, sample i at slot i.
struct DevModChan {
  // value of sample (x, y): kind 0 = quantised DC plane `plane` at (x0 + x, y0 + y); 1 = constant `konst`;
  // 2 = the (strategy, quant - 1) rows of the DC group's varblock list
  uint32_t kind, w, h;
  int32_t konst;
  uint64_t plane;
};

JXLB_HD int32_t DevModValue(const DevEPools& E, const DevEFrame& ef, const DevModChan& ch, uint32_t g, uint32_t x0,
                            uint32_t y0, uint32_t x, uint32_t y) {
  if (ch.kind == 0) return E.iarena[ch.plane + static_cast<size_t>(y0 + y) * ef.xblocks + x0 + x];
  if (ch.kind == 1) return ch.konst;
  if (ch.kind == 4)  // alpha of the input pixel (x0 + x, y0 + y), x0 / y0 in pixels
    return E.bytes_in[ef.rgb + (static_cast<size_t>(y0 + y) * ef.xsize + x0 + x) * 4 + 3];
  if (ch.kind == 3)  // chroma-from-luma map (int8 per tile; the DC group starts at tile (x0 / 8, y0 / 8))
    return reinterpret_cast<const int8_t*>(E.barena + ch.plane)[static_cast<size_t>((y0 >> 3) + y) * ef.cmw + (x0 >> 3) + x];
  const int32_t pos = E.iarena[ef.block_of_num + static_cast<size_t>(g) * 65536 + x];
  if (y == 1) return E.barena[ef.raw_quant + pos];  // raw quant - 1
  return E.barena[ef.acs + pos] >> 1;
}

// Token of sample (x, y) of channel `chan` (index inside its stream's image) of stream `stream_id`.
JXLB_HD uint2 DevEncModularToken(const DevEPools& E, const DevEFrame& ef, const DevModChan& ch, uint32_t chan,
                                 uint32_t stream_id, uint32_t g, uint32_t x0, uint32_t y0, uint32_t x, uint32_t y) {
  const int32_t v = DevModValue(E, ef, ch, g, x0, y0, x, y);
  const int32_t left = x ? DevModValue(E, ef, ch, g, x0, y0, x - 1, y) : (y ? DevModValue(E, ef, ch, g, x0, y0, x, y - 1) : 0);
  const int32_t top = y ? DevModValue(E, ef, ch, g, x0, y0, x, y - 1) : left;
  const int32_t topleft = (x && y) ? DevModValue(E, ef, ch, g, x0, y0, x - 1, y - 1) : left;
  int32_t props[12];
  props[0] = static_cast<int32_t>(chan);
  props[1] = static_cast<int32_t>(stream_id);
  props[2] = static_cast<int32_t>(y);
  props[3] = static_cast<int32_t>(x);
  props[4] = top > 0 ? top : -top;
  props[5] = left > 0 ? left : -left;
  props[6] = top;
  props[7] = left;
  props[8] = 0;  // (the fixed tree never tests property 8)
  props[9] = left + top - topleft;
  props[10] = left - topleft;
  props[11] = topleft - top;
  uint32_t pos = 0;
  DevEncTreeNode n = E.tree[0];
  while (n.prop >= 0) {
    pos = props[n.prop] > n.a ? n.l : n.r;
    n = E.tree[pos];
  }
  int32_t guess = 0;
  if (n.a == 1) guess = left;
  else if (n.a == 2) guess = top;
  else if (n.a == 5) guess = DevClampedGradient(left, top, topleft);
  return make_uint2(n.l, DevPackSigned(v - guess));
}

// Alpha sample of pixel `i` of the frame -> its token in its group's alpha stream (DevEFrame::alpha_tokens), counted in
// the Modular histograms. The stream is the global one (id 0) when the image fits one group, else ModularAC(g, pass 0);
// neighbours inside the group's crop, the fixed global tree decides context and predictor like for every other stream.
JXLB_HD void DevEncAlphaSample(const DevEPools& E, const DevEFrame& ef, uint64_t i) {
  const uint32_t y = static_cast<uint32_t>(i / ef.xsize), x = static_cast<uint32_t>(i - static_cast<uint64_t>(y) * ef.xsize);
  const bool global_only = ef.xsize <= 256 && ef.ysize <= 256;
  const uint32_t gx = x >> 8, gy = y >> 8, lx = x & 255, ly = y & 255, g = gy * ef.xgroups + gx;
  const uint32_t gw = ef.xsize - (gx << 8) < 256 ? ef.xsize - (gx << 8) : 256;
  const uint32_t gh = ef.ysize - (gy << 8) < 256 ? ef.ysize - (gy << 8) : 256;
  DevModChan ch;
  ch.kind = 4;
  ch.w = gw;
  ch.h = gh;
  ch.konst = 0;
  ch.plane = 0;
  const uint32_t ndc = ef.xdcgroups * ef.ydcgroups;
  const uint32_t stream_id = global_only ? 0 : 1 + 3 * ndc + 17 + g;  // ModularStreamId::ModularAC(g, pass 0)
  const uint2 t = DevEncModularToken(E, ef, ch, 0, stream_id, g, gx << 8, gy << 8, lx, ly);
  E.tokens[ef.alpha_tokens + static_cast<uint64_t>(g) * 65536 + ly * gw + lx] = t;
  DevCountToken(reinterpret_cast<uint32_t*>(E.iarena + ef.mod_hist), t.x, t.y);
}

// ---- rANS writer. Tables of one code: fs[cluster][256] = freq | first reverse slot << 16, reverse[cluster][4096].
struct DevEncCode {
  const uint32_t* fs;
  const uint16_t* reverse;
};

JXLB_HD void DevWriteBitsAt(uint32_t* words, uint64_t pos, uint32_t nbits, uint32_t value) {
  if (nbits == 0) return;
  const uint32_t sh = static_cast<uint32_t>(pos & 31);
  words[pos >> 5] |= value << sh;
  if (sh + nbits > 32) words[(pos >> 5) + 1] |= value >> (32 - sh);
}

// Bits written back to front: `acc` holds the bits of [cursor, cursor + nacc); whole aligned words leave with a plain
// store, only the two ends of the range are merged into words that hold other data.
struct DevBackWriter {
  uint32_t* words;
  uint64_t cursor;
  uint64_t acc;
  uint32_t nacc;
  JXLB_HD void Init(uint32_t* w, uint64_t end_pos) {
    words = w;
    cursor = end_pos;
    acc = 0;
    nacc = 0;
  }
  JXLB_HD void Put(uint32_t nbits, uint32_t value) {  // nbits <= 32
    acc = (acc << nbits) | value;
    nacc += nbits;
    cursor -= nbits;
    uint64_t hi = cursor + nacc;
    const uint32_t r = static_cast<uint32_t>(hi & 31);
    if (r != 0 && nacc >= r) {  // (only until the top end reaches a word boundary)
      words[hi >> 5] |= static_cast<uint32_t>(acc >> (nacc - r)) & ((1u << r) - 1);
      nacc -= r;
      acc &= (uint64_t{1} << nacc) - 1;
      hi -= r;
    }
    if ((hi & 31) == 0 && nacc >= 32) {
      words[(hi >> 5) - 1] = static_cast<uint32_t>(acc >> (nacc - 32));
      nacc -= 32;
      acc &= (uint64_t{1} << nacc) - 1;
    }
  }
  JXLB_HD void Finish() {  // the remaining low bits share their word with whatever precedes the range
    if (nacc == 0) return;
    const uint64_t hi = cursor + nacc;
    const uint32_t r = static_cast<uint32_t>(hi & 31);
    if (r != 0 && nacc > r) {  // spans the boundary below an unaligned top: split
      words[hi >> 5] |= static_cast<uint32_t>(acc >> (nacc - r)) & ((1u << r) - 1);
      nacc -= r;
      acc &= (uint64_t{1} << nacc) - 1;
    }
    words[cursor >> 5] |= static_cast<uint32_t>(acc) << (cursor & 31);
    nacc = 0;
  }
};

// Pushes the `n` tokens (cluster, value) onto the rANS stream in reverse order (the decoder pops them forwards),
// writing back to front through `w`; ends with the 32-bit state. Everything that does not depend on the coder
// state (the next token, its hybrid split, its frequency entry) is fetched one token ahead.
JXLB_HD void DevRansPush(const uint2* tok, uint32_t n, const DevEncCode& code, DevBackWriter& w) {
  uint32_t state = 0x13u << 16;
  uint2 t = n ? tok[n - 1] : make_uint2(0, 0);
  uint2 raw_ahead = n > 1 ? tok[n - 2] : make_uint2(0, 0);  // the token after next: its load has a whole iteration to land
  uint32_t token = 0, nbits = 0, bits = 0, fs = 0;
  if (n) {
    DevHybrid420(t.y, &token, &nbits, &bits);
    fs = JXLB_LDG(code.fs + t.x * 256 + token);
  }
  for (uint32_t i = n; i-- > 0;) {
    const uint32_t cluster = t.x, cur_nbits = nbits, cur_bits = bits, cur_fs = fs;
    if (i > 0) {
      t = raw_ahead;
      if (i > 1) raw_ahead = tok[i - 2];
      DevHybrid420(t.y, &token, &nbits, &bits);
      fs = JXLB_LDG(code.fs + t.x * 256 + token);
    }
    if (cur_nbits) w.Put(cur_nbits, cur_bits);
    const uint32_t f = cur_fs & 0xFFFF;
    if ((state >> 20) >= f) {
      w.Put(16, state & 0xFFFF);
      state >>= 16;
    }
    const uint32_t q = state / f, r = state - q * f;
    state = (q << 12) | JXLB_LDG(code.reverse + cluster * 4096 + (cur_fs >> 16) + r);
  }
  w.Put(32, state);
}

// Forward 8-point DCT of one line in registers: the statements of CoopDCT (jxlb_vardct_dev.h) for n = 8, element by
// element in the same order (AddReverse / SubReverse + Multiply at 8 and 4, the 2-point butterflies, B + InverseEvenOdd
// at 4 and 8, the 1 / 8 scale), so the coefficients are bit-identical to the generic path.
JXLB_HD void DevFwdDct8(float* v, const float* wc) {
  const float w0 = wc[0], w1 = wc[1], w2 = wc[2], w3 = wc[3], w4 = wc[4], w5 = wc[5];
  const float a0 = v[0] + v[7], a1 = v[1] + v[6], a2 = v[2] + v[5], a3 = v[3] + v[4];
  const float a4 = (v[0] - v[7]) * w2, a5 = (v[1] - v[6]) * w3, a6 = (v[2] - v[5]) * w4, a7 = (v[3] - v[4]) * w5;
  const float b0 = a0 + a3, b1 = a1 + a2, b2 = (a0 - a3) * w0, b3 = (a1 - a2) * w1;
  const float b4 = a4 + a7, b5 = a5 + a6, b6 = (a4 - a7) * w0, b7 = (a5 - a6) * w1;
  const float c0 = b0 + b1, c1 = b0 - b1, c2 = b2 + b3, c3 = b2 - b3;
  const float c4 = b4 + b5, c5 = b4 - b5, c6 = b6 + b7, c7 = b6 - b7;
  const float d1 = fmaf(c2, kDevSqrt2, c3), d5 = fmaf(c6, kDevSqrt2, c7);  // d0 = c0, d2 = c1, d3 = c3, d4 = c4, d6 = c5, d7 = c7
  v[0] = 0.125f * c0;
  v[2] = 0.125f * d1;
  v[4] = 0.125f * c1;
  v[6] = 0.125f * c3;
  v[1] = 0.125f * fmaf(c4, kDevSqrt2, d5);
  v[3] = 0.125f * (d5 + c5);
  v[5] = 0.125f * (c5 + c7);
  v[7] = 0.125f * c7;
}

#if defined(__CUDACC__)
// MODE 0 of DevEncVarblock for a DCT8X8 varblock, by one warp: lane = channel * 8 + line (24 lanes). Columns, then rows
// through 3 x 72 floats of `buf`; coefficient (yfreq, xfreq) goes to row xfreq, column yfreq of the block's footprint
// (the layout DevEncVarblock writes for R == C).
__device__ __forceinline__ void DevEncDct8Warp(const DevEPools& E, const DevEFrame& ef, uint32_t bx, uint32_t by, float* buf, uint32_t lane) {
  const uint32_t PW = ef.xblocks * 8;
  const size_t origin = static_cast<size_t>(by) * 8 * PW + static_cast<size_t>(bx) * 8;
  const float* wc = E.fpool + E.wc_off;
  const uint32_t c = lane >> 3, t = lane & 7;
  float v[8];
  if (lane < 24) {
    const float* px = E.farena + ef.xyb[c] + origin + t;
#pragma unroll
    for (uint32_t y = 0; y < 8; y++) v[y] = px[static_cast<size_t>(y) * PW];
    DevFwdDct8(v, wc);
#pragma unroll
    for (uint32_t y = 0; y < 8; y++) buf[c * 72 + y * 9 + t] = v[y];
  }
  __syncwarp();
  if (lane < 24) {
#pragma unroll
    for (uint32_t x = 0; x < 8; x++) v[x] = buf[c * 72 + t * 9 + x];
    DevFwdDct8(v, wc);
    float* out = E.farena + ef.xyb_raw[c] + origin + t;
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) out[static_cast<size_t>(k) * PW] = v[k];
  }
  __syncwarp();
}

// Bits into a zero-initialised region shared with other writers of the warp: nbits <= 32, value < 2^nbits.
__device__ __forceinline__ void DevOrBits(uint32_t* words, uint64_t pos, uint32_t nbits, uint32_t value) {
  if (nbits == 0) return;
  const uint32_t sh = static_cast<uint32_t>(pos & 31);
  atomicOr(words + (pos >> 5), value << sh);
  if (sh + nbits > 32) atomicOr(words + (pos >> 5) + 1, value >> (32 - sh));
}

// The same stream written by a warp, back to front from bit `*cursor` (uniform over the lanes, updated). The 32 lanes
// fetch 32 tokens at a time (one coalesced load, from the end), split them, fetch their frequency entries and the
// reciprocal of the frequency in parallel, one chunk ahead of the coder; lane 0 owns the coder state and takes the
// tokens of a chunk from the lanes by shuffle. On its serial chain stay only the renormalisation test, the division
// (multiply by the reciprocal + one correction) and the reverse-table load that depend on the state: the renormalised
// 16 bits leave as fire-and-forget atomics, and every lane writes the extra bits of its own token after the chunk,
// at the position that follows from a prefix sum of the bit counts and the chunk's renormalisation mask.
// (One thread per section left 31 lanes of a warp to other sections with other branch histories: the warp ran them
// one after the other, every token load was a dependent miss of its own, and the bit writer sat on the chain.)
// kSmem: `srev` holds the reverse tables of some clusters in shared memory (4096 entries each), `sslot[cluster]` says
// which (0xFF: not staged, read from global memory): the load on the chain then is a shared-memory load.
// `stage`: 32 uint4 of shared memory of this warp. The chain itself runs on ALL lanes (the same state everywhere, the
// chunk's tokens read from `stage` by broadcast): no divergent region per token, one shared-memory load instead of
// three shuffles -- a lone warp issues an instruction every ~ 5 cycles, so the chain's length is its instruction count.
template <bool kSmem = false>
__device__ __forceinline__ void DevRansPushWarp(const uint2* tok, uint32_t n, const DevEncCode& code, uint32_t* words, uint64_t* cursor,
                                                uint32_t lane, uint4* stage, const uint16_t* srev = nullptr,
                                                const uint8_t* sslot = nullptr) {
  uint32_t state = 0x13u << 16;
  uint64_t pos = *cursor;
  auto fetch = [&](uint32_t done, uint32_t* cn, uint32_t* bt, uint32_t* fs, uint32_t* rcp) {
    *cn = 0;
    *bt = 0;
    *fs = 1;
    *rcp = 0;
    if (done + lane < n) {
      const uint2 t = tok[n - 1 - done - lane];
      uint32_t token, nbits;
      DevHybrid420(t.y, &token, &nbits, bt);
      *cn = t.x | (nbits << 16);
      if (kSmem) *cn |= static_cast<uint32_t>(sslot[t.x]) << 8;  // (clusters are below 256)
      *fs = JXLB_LDG(code.fs + t.x * 256 + token);
      *rcp = 0xFFFFFFFFu / (*fs & 0xFFFF);  // >= 2^32 / f - 1: the quotient estimate is at most one short
    }
  };
  uint32_t cn, bt, fs, rcp;
  fetch(0, &cn, &bt, &fs, &rcp);
  for (uint32_t done = 0; done < n; done += 32) {
    uint32_t ncn, nbt, nfs, nrcp;
    fetch(done + 32, &ncn, &nbt, &nfs, &nrcp);  // in flight while this chunk is coded
    const uint32_t cnt = n - done < 32 ? n - done : 32;
    // bits of the tokens before mine in the chunk (inclusive prefix sum of the extra-bit counts)
    const uint32_t my_nbits = cn >> 16;
    uint32_t incl = my_nbits;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += v;
    }
    stage[lane] = make_uint4(cn, fs, rcp, 0);
    __syncwarp();
    uint32_t mask = 0;  // tokens of the chunk after which the state was renormalised
    uint64_t p = pos;   // write position
    for (uint32_t k = 0; k < cnt; k++) {
      const uint4 t = stage[k];
      p -= t.x >> 16;
      const uint32_t f = t.y & 0xFFFF;
      if ((state >> 20) >= f) {  // (uniform)
        p -= 16;
        if (lane == 0) DevOrBits(words, p, 16, state & 0xFFFF);
        state >>= 16;
        mask |= 1u << k;
      }
      uint32_t q = __umulhi(state, t.z), r = state - q * f;
      if (r >= f) {
        q++;
        r -= f;
      }
      const uint32_t at = (t.y >> 16) + r, slot = (t.x >> 8) & 0xFF;
      uint32_t rev;
      if (kSmem && slot != 0xFF) {
        rev = srev[slot * 4096 + at];
      } else {
        rev = JXLB_LDG(code.reverse + (t.x & 0xFF) * 4096 + at);
      }
      state = (q << 12) | rev;
    }
    __syncwarp();  // (the next chunk overwrites `stage`)
    if (lane < cnt && my_nbits != 0)
      DevOrBits(words, pos - incl - 16ull * __popc(mask & ((1u << lane) - 1)), my_nbits, bt);
    pos -= __shfl_sync(0xFFFFFFFFu, incl, 31) + 16ull * __popc(mask);
    cn = ncn;
    bt = nbt;
    fs = nfs;
    rcp = nrcp;
  }
  pos -= 32;
  if (lane == 0) DevOrBits(words, pos, 32, state);
  *cursor = pos;
}

// Header bits between the streams of a section, written by lane 0 (the cursor stays uniform).
__device__ __forceinline__ void DevWarpPut(uint32_t* words, uint64_t* cursor, uint32_t nbits, uint32_t value, uint32_t lane) {
  *cursor -= nbits;
  if (lane == 0) DevOrBits(words, *cursor, nbits, value);
}
#endif

// Token slots of DC group g's Modular streams: [DC Y | DC X | DC B | ytox | ytob | strategy row, quant row | sharpness].
struct DevDcGroupLayout {
  uint32_t x0, y0, xs, ys, cw, chh, count;
  uint32_t dc_tokens, meta_tokens;  // counts
};

JXLB_HD DevDcGroupLayout DevDcGroupGeometry(const DevEPools& E, const DevEFrame& ef, uint32_t g) {
  DevDcGroupLayout L;
  L.x0 = (g % ef.xdcgroups) * 256;
  L.y0 = (g / ef.xdcgroups) * 256;
  L.xs = ef.xblocks - L.x0 < 256 ? ef.xblocks - L.x0 : 256;
  L.ys = ef.yblocks - L.y0 < 256 ? ef.yblocks - L.y0 : 256;
  L.cw = (L.xs + 7) >> 3;
  L.chh = (L.ys + 7) >> 3;
  L.count = static_cast<uint32_t>(E.iarena[ef.dcg_count + g]);
  L.dc_tokens = 3 * L.xs * L.ys;
  L.meta_tokens = 2 * L.cw * L.chh + 2 * L.count + L.xs * L.ys;
  return L;
}

// Token `i` of DC group g (0 <= i < dc_tokens + meta_tokens), written to its slot and counted.
JXLB_HD void DevEncModularSample(const DevEPools& E, const DevEFrame& ef, uint32_t g, const DevDcGroupLayout& L, uint32_t i) {
  const uint32_t ndc = ef.xdcgroups * ef.ydcgroups;
  DevModChan ch;
  uint32_t chan, stream_id, local;
  if (i < L.dc_tokens) {
    chan = i / (L.xs * L.ys);
    local = i - chan * L.xs * L.ys;
    ch.kind = 0;
    ch.w = L.xs;
    ch.h = L.ys;
    ch.konst = 0;
    ch.plane = ef.dcq[chan];
    stream_id = 1 + g;
  } else {
    uint32_t j = i - L.dc_tokens;
    stream_id = 1 + 2 * ndc + g;
    const uint32_t n01 = L.cw * L.chh;
    ch.plane = 0;
    if (j < 2 * n01) {
      chan = j / n01;
      local = j - chan * n01;
      ch.kind = 3;
      ch.konst = 0;
      ch.plane = chan == 0 ? ef.ytox : ef.ytob;
      ch.w = L.cw;
      ch.h = L.chh;
    } else if (j < 2 * n01 + 2 * L.count) {
      chan = 2;
      local = j - 2 * n01;
      ch.kind = 2;
      ch.konst = 0;
      ch.w = L.count;
      ch.h = 2;
    } else {
      chan = 3;
      local = j - 2 * n01 - 2 * L.count;
      ch.kind = 1;
      ch.konst = 4;  // EPF sharpness
      ch.w = L.xs;
      ch.h = L.ys;
    }
  }
  const uint32_t x = local % ch.w, y = local / ch.w;
  const uint2 t = DevEncModularToken(E, ef, ch, chan, stream_id, g, L.x0, L.y0, x, y);
  E.tokens[ef.mod_tokens + ef.mod_tokens_stride * g + i] = t;
  DevCountToken(reinterpret_cast<uint32_t*>(E.iarena + ef.mod_hist), t.x, t.y);
}

// The DC group section -- extra_precision, DC stream, varblock count, AC-metadata stream -- written back to front
// so that it ends at bit `end_pos` of `words`. Returns the position of its first bit.
JXLB_HD uint64_t DevEncEmitDcGroup(const DevEPools& E, const DevEFrame& ef, uint32_t g, const DevEncCode& code, uint32_t* words,
                                   uint64_t end_pos) {
  const DevDcGroupLayout L = DevDcGroupGeometry(E, ef, g);
  const uint2* tok = E.tokens + ef.mod_tokens + ef.mod_tokens_stride * g;
  DevBackWriter w;
  w.Init(words, end_pos);
  DevRansPush(tok + L.dc_tokens, L.meta_tokens, code, w);
  w.Put(4, 0x3);  // use_global_tree = 1, default WP header = 1, no transforms
  uint32_t count_bits = 0;
  while ((1u << count_bits) < L.xs * L.ys) count_bits++;
  if (count_bits) w.Put(count_bits, L.count - 1);
  DevRansPush(tok, L.dc_tokens, code, w);
  w.Put(4, 0x3);
  w.Put(2, 0);  // extra_precision
  w.Finish();
  return w.cursor;
}

// An AC group section ending at bit `end_pos`; returns the position of its first bit. `alpha_tok` (else null): the
// `alpha_n` tokens of the group's extra-channel stream, which follows the coefficients: GroupHeader (global tree, default
// weighted-predictor header, no transforms), then its symbols under the Modular code.
JXLB_HD uint64_t DevEncEmitAcGroup(const uint2* tok, uint32_t n, const DevEncCode& code, uint32_t* words, uint64_t end_pos,
                                   const uint2* alpha_tok = nullptr, uint32_t alpha_n = 0, const DevEncCode* mod_code = nullptr) {
  DevBackWriter w;
  w.Init(words, end_pos);
  if (alpha_tok != nullptr) {
    DevRansPush(alpha_tok, alpha_n, *mod_code, w);
    w.Put(4, 0x3);
  }
  DevRansPush(tok, n, code, w);
  w.Finish();
  return w.cursor;
}

#if defined(__CUDACC__)
// The two section writers above, a warp per section (DevRansPushWarp); returns the position of the first bit.
// A DC-group section is two chained streams -- [extra_precision, GroupHeader, DC samples] then [varblock count,
// GroupHeader, AC metadata] -- of up to 196 K and 200 K tokens: by far the longest serial chains of a frame. The two
// halves are written by two warps into regions of their own (`half` 0 / 1) and joined bit-wise by the host assembly.
__device__ __forceinline__ uint64_t DevEncEmitDcGroupWarp(const DevEPools& E, const DevEFrame& ef, uint32_t g, const DevEncCode& code,
                                                          uint32_t* words, uint64_t end_pos, uint32_t lane, uint32_t half,
                                                          uint4* stage, const uint16_t* srev, const uint8_t* sslot) {
  const DevDcGroupLayout L = DevDcGroupGeometry(E, ef, g);
  const uint2* tok = E.tokens + ef.mod_tokens + ef.mod_tokens_stride * g;
  uint64_t cursor = end_pos;
  if (half == 1) {
    DevRansPushWarp<true>(tok + L.dc_tokens, L.meta_tokens, code, words, &cursor, lane, stage, srev, sslot);
    DevWarpPut(words, &cursor, 4, 0x3, lane);  // use_global_tree = 1, default WP header = 1, no transforms
    uint32_t count_bits = 0;
    while ((1u << count_bits) < L.xs * L.ys) count_bits++;
    if (count_bits) DevWarpPut(words, &cursor, count_bits, L.count - 1, lane);
    return cursor;
  }
  DevRansPushWarp<true>(tok, L.dc_tokens, code, words, &cursor, lane, stage, srev, sslot);
  DevWarpPut(words, &cursor, 4, 0x3, lane);
  DevWarpPut(words, &cursor, 2, 0, lane);  // extra_precision
  return cursor;
}

__device__ __forceinline__ uint64_t DevEncEmitAcGroupWarp(const uint2* tok, uint32_t n, const DevEncCode& code, uint32_t* words,
                                                          uint64_t end_pos, uint32_t lane, uint4* stage, const uint2* alpha_tok = nullptr,
                                                          uint32_t alpha_n = 0, const DevEncCode* mod_code = nullptr) {
  uint64_t cursor = end_pos;
  if (alpha_tok != nullptr) {
    DevRansPushWarp(alpha_tok, alpha_n, *mod_code, words, &cursor, lane, stage);
    DevWarpPut(words, &cursor, 4, 0x3, lane);
  }
  DevRansPushWarp(tok, n, code, words, &cursor, lane, stage);
  return cursor;
}
#endif

}  // namespace jxlb

#endif  // JXLB_ENC_DEV_H_
