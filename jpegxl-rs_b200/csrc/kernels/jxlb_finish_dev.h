// jxl_b200 device code: inverse Modular transforms and sample output.
//
// Restates (as data-parallel element functions)
//   lib/jxl/modular/transform/rct.cc:21-147     InvRCT (7 types x 6 permutations),
//   lib/jxl/modular/transform/palette.cc:15-176 InvPalette, palette.h:54-130 GetPaletteValue,
//   lib/jxl/modular/transform/squeeze.cc:104-357 InvHSqueeze / InvVSqueeze, squeeze.h:60-82,
//   lib/jxl/dec_modular.cc:534-708 int -> float,
//   lib/jxl/render_pipeline/stage_write.cc:51-113 float -> u8/u16/f16/f32 (+ 8x8 ordered dither).
#ifndef JXLB_FINISH_DEV_H_
#define JXLB_FINISH_DEV_H_

#include "jxlb_modular_dev.h"

namespace jxlb {

JXLB_HD int32_t DevAdd32(int32_t a, int32_t b) {
  return static_cast<int32_t>(static_cast<uint32_t>(a) + static_cast<uint32_t>(b));
}

// One pixel of the inverse reversible colour transform, in place.
JXLB_HD void DevInvRCTPixel(int32_t* p0, int32_t* p1, int32_t* p2, size_t i, uint32_t rct_type) {
  const uint32_t perm = rct_type / 7, custom = rct_type % 7;
  const int32_t a = p0[i], b = p1[i], c = p2[i];
  int32_t o0, o1, o2;
  if (custom == 6) {
    const int32_t tmp = DevAdd32(a, -(c >> 1));
    o1 = DevAdd32(c, tmp);
    o2 = DevAdd32(tmp, -(b >> 1));
    o0 = DevAdd32(o2, b);
  } else {
    const uint32_t second = custom >> 1, third = custom & 1;
    int32_t S = b, T = c;
    if (third) T = DevAdd32(T, a);
    if (second == 1) S = DevAdd32(S, a);
    else if (second == 2) S = DevAdd32(S, DevAdd32(a, T) >> 1);
    o0 = a; o1 = S; o2 = T;
  }
  int32_t* dst[3] = {p0, p1, p2};
  dst[perm % 3][i] = o0;
  dst[(perm + 1 + perm / 3) % 3][i] = o1;
  dst[(perm + 2 - perm / 3) % 3][i] = o2;
}

JXLB_HD int32_t DevPaletteValue(const int32_t* pal, int pal_w, int index, int c, int bit_depth) {
  if (index < 0) {
    if (c >= 3) return 0;
    const int16_t kDelta[72][3] = {
      {0, 0, 0}, {4, 4, 4}, {11, 0, 0}, {0, 0, -13}, {0, -12, 0}, {-10, -10, -10}, {-18, -18, -18}, {-27, -27, -27},
      {-18, -18, 0}, {0, 0, -32}, {-32, 0, 0}, {-37, -37, -37}, {0, -32, -32}, {24, 24, 45}, {50, 50, 50},
      {-45, -24, -24}, {-24, -45, -45}, {0, -24, -24}, {-34, -34, 0}, {-24, 0, -24}, {-45, -45, -24}, {64, 64, 64},
      {-32, 0, -32}, {0, -32, 0}, {-32, 0, 32}, {-24, -45, -24}, {45, 24, 45}, {24, -24, -45}, {-45, -24, 24},
      {80, 80, 80}, {64, 0, 0}, {0, 0, -64}, {0, -64, -64}, {-24, -24, 45}, {96, 96, 96}, {64, 64, 0}, {45, -24, -24},
      {34, -34, 0}, {112, 112, 112}, {24, -45, -45}, {45, 45, -24}, {0, -32, 32}, {24, -24, 45}, {0, 96, 96},
      {45, -24, 24}, {24, -45, -24}, {-24, -45, 24}, {0, -64, 0}, {96, 0, 0}, {128, 128, 128}, {64, 0, 64},
      {144, 144, 144}, {96, 96, 0}, {-36, -36, 36}, {45, -24, -45}, {45, -45, -24}, {0, 0, -96}, {0, 128, 128},
      {0, 96, 0}, {45, 24, -45}, {-128, 0, 0}, {24, -45, 24}, {-45, 24, -45}, {64, 0, -64}, {64, -64, -64},
      {96, 0, 96}, {45, -45, 24}, {24, 45, -45}, {64, 64, -64}, {128, 128, 0}, {0, 0, -128}, {-24, 45, -45}};
    index = -(index + 1);
    index %= 143;
    const int e = (index + 1) >> 1;
    const int32_t v = kDelta[e][c];
    int32_t r = (index & 1) ? v : -v;
    if (bit_depth > 8) r *= 1 << (bit_depth - 8);
    return r;
  } else if (pal_w <= index && index < pal_w + 64) {
    if (c >= 3) return 0;
    index -= pal_w;
    index >>= c * 2;
    const int sh = bit_depth - 3 > 0 ? bit_depth - 3 : 0;
    return static_cast<int32_t>((static_cast<uint64_t>(index % 4) * ((uint64_t{1} << bit_depth) - 1)) >> 2) + (1 << sh);
  } else if (pal_w + 64 <= index) {
    if (c >= 3) return 0;
    index -= pal_w + 64;
    if (c == 1) index /= 5;
    if (c == 2) index /= 25;
    return static_cast<int32_t>((static_cast<uint64_t>(index % 5) * ((uint64_t{1} << bit_depth) - 1)) >> 2);
  }
  return pal[static_cast<size_t>(c) * pal_w + index];
}

JXLB_HD int64_t DevSmoothTendency(int64_t B, int64_t a, int64_t n) {
  int64_t diff = 0;
  if (B >= a && a >= n) {
    diff = (4 * B - 3 * n - a + 6) / 12;
    if (diff - (diff & 1) > 2 * (B - a)) diff = 2 * (B - a) + 1;
    if (diff + (diff & 1) > 2 * (a - n)) diff = 2 * (a - n);
  } else if (B <= a && a <= n) {
    diff = (4 * B - 3 * n - a - 6) / 12;
    if (diff + (diff & 1) < 2 * (B - a)) diff = 2 * (B - a) - 1;
    if (diff - (diff & 1) < 2 * (a - n)) diff = 2 * (a - n);
  }
  return diff;
}

// Executes op `op` cooperatively: worker `tid` of `nthreads`. The caller
// synchronises the workers between ops.
JXLB_HD void DevRunOp(const DevPools& P, const DevOp& op, uint32_t tid, uint32_t nthreads) {
  switch (op.kind) {
    case kOpRCT: {
      const DevPlane a = P.planes[op.a], b = P.planes[op.b], c = P.planes[op.c];
      const size_t n = static_cast<size_t>(a.w) * a.h;
      int32_t* p0 = P.arena + a.off;
      int32_t* p1 = P.arena + b.off;
      int32_t* p2 = P.arena + c.off;
      for (size_t i = tid; i < n; i += nthreads) DevInvRCTPixel(p0, p1, p2, i, op.p0);
      break;
    }
    case kOpPalette: {
      const DevPlane idx = P.planes[op.a], pal = P.planes[op.b];
      const int nb = static_cast<int>(op.p0), bit_depth = static_cast<int>(op.p1);
      const uint32_t nb_deltas = op.p2, predictor = op.p3;
      const int32_t* palp = P.arena + pal.off;
      const int pal_w = static_cast<int>(pal.w);
      const size_t n = static_cast<size_t>(idx.w) * idx.h;
      int32_t* out0 = P.arena + idx.off;
      if (nb_deltas == 0 && predictor == 0) {
        for (size_t i = tid; i < n; i += nthreads) {
          int index = out0[i];
          if (nb == 1) index = index < 0 ? 0 : (index > pal_w - 1 ? pal_w - 1 : index);
          out0[i] = DevPaletteValue(palp, pal_w, index, 0, bit_depth);
          for (int c = 1; c < nb; c++) {
            const DevPlane o = P.planes[op.c + c - 1];
            P.arena[o.off + i] = DevPaletteValue(palp, pal_w, index, c, bit_depth);
          }
        }
      } else {
        // Delta palette: prediction from already reconstructed neighbours, serial per
        // channel. op.pad holds a scratch plane with a copy slot for the indices;
        // one worker per channel.
        const DevPlane copy = P.planes[op.pad];
        int32_t* indices = P.arena + copy.off;
        if (tid < static_cast<uint32_t>(nb)) {
          const int c = static_cast<int>(tid);
          int32_t* chan = c == 0 ? out0 : P.arena + P.planes[op.c + c - 1].off;
          const int w = static_cast<int>(idx.w), h = static_cast<int>(idx.h);
          DevWPPlain wp{};
          const bool use_wp = predictor == 6;
          // WP scratch for this op lives behind the index copy.
          int32_t* scratch = indices + n + static_cast<size_t>(c) * 10 * (w + 2);
          uint32_t lut[64];
          if (use_wp) {
            for (uint32_t k = 0; k < 64; k++) lut[k] = (1u << 24) / (k + 1);
            wp.InitPlain(op.wp_params, scratch, w, lut);
            wp.Reset(w);
          }
          for (int y = 0; y < h; y++) {
            int32_t* row = chan + static_cast<size_t>(y) * w;
            const int32_t* prev = y ? row - w : nullptr;
            const int32_t* prevprev = y > 1 ? row - 2 * w : nullptr;
            for (int x = 0; x < w; x++) {
              const int index = indices[static_cast<size_t>(y) * w + x];
              int64_t val = DevPaletteValue(palp, pal_w, index, c, bit_depth);
              const DevNeighbors nbh = DevLoadNeighbors(row, prev, prevprev, x, y, w);
              int64_t wp_pred = 0;
              if (use_wp) wp_pred = wp.Predict(x, y, w, nbh.top, nbh.left, nbh.topright, nbh.topleft, nbh.toptop, nullptr);
              if (index < static_cast<int32_t>(nb_deltas)) val += DevPredictOne(predictor, nbh, wp_pred);
              row[x] = static_cast<int32_t>(val);
              if (use_wp) wp.Update(row[x], x, y);
            }
          }
        }
      }
      break;
    }
    case kOpCopy: {
      const DevPlane s = P.planes[op.a], d = P.planes[op.b];
      const size_t n = static_cast<size_t>(s.w) * s.h;
      for (size_t i = tid; i < n; i += nthreads) {
        const uint32_t y = static_cast<uint32_t>(i / s.w), x = static_cast<uint32_t>(i % s.w);
        P.arena[d.off + static_cast<size_t>(op.p1 + y) * d.w + op.p0 + x] = P.arena[s.off + i];
      }
      break;
    }
    case kOpHSqueeze: {
      const DevPlane avg = P.planes[op.a], res = P.planes[op.b], out = P.planes[op.c];
      for (uint32_t y = tid; y < avg.h; y += nthreads) {
        const int32_t* pa = P.arena + avg.off + static_cast<size_t>(y) * avg.w;
        const int32_t* pr = P.arena + res.off + static_cast<size_t>(y) * res.w;
        int32_t* po = P.arena + out.off + static_cast<size_t>(y) * out.w;
        for (uint32_t x = 0; x < res.w; x++) {
          const int64_t a = pa[x];
          const int64_t next_avg = x + 1 < avg.w ? pa[x + 1] : a;
          const int64_t left = x ? po[(x << 1) - 1] : a;
          const int64_t diff = pr[x] + DevSmoothTendency(left, a, next_avg);
          const int64_t A = a + diff / 2;
          po[x << 1] = static_cast<int32_t>(A);
          po[(x << 1) + 1] = static_cast<int32_t>(A - diff);
        }
        if (out.w & 1) po[out.w - 1] = pa[avg.w - 1];
      }
      break;
    }
    case kOpVSqueeze: {
      const DevPlane avg = P.planes[op.a], res = P.planes[op.b], out = P.planes[op.c];
      for (uint32_t x = tid; x < avg.w; x += nthreads) {
        for (uint32_t y = 0; y < res.h; y++) {
          const int64_t a = P.arena[avg.off + static_cast<size_t>(y) * avg.w + x];
          const int64_t next_avg = P.arena[avg.off + static_cast<size_t>(y + 1 < avg.h ? y + 1 : y) * avg.w + x];
          const int64_t top = y > 0 ? P.arena[out.off + static_cast<size_t>((y << 1) - 1) * out.w + x] : a;
          const int64_t diff = P.arena[res.off + static_cast<size_t>(y) * res.w + x] + DevSmoothTendency(top, a, next_avg);
          const int64_t o = a + diff / 2;
          P.arena[out.off + static_cast<size_t>(y << 1) * out.w + x] = static_cast<int32_t>(o);
          P.arena[out.off + static_cast<size_t>((y << 1) + 1) * out.w + x] = static_cast<int32_t>(o - diff);
        }
        if (out.h & 1) P.arena[out.off + static_cast<size_t>(out.h - 1) * out.w + x] = P.arena[avg.off + static_cast<size_t>(avg.h - 1) * avg.w + x];
      }
      break;
    }
    default: break;
  }
}

// ---------------------------------------------------------------- output
JXLB_HD float DevDither(uint32_t x, uint32_t y) {
  // lib/jxl/render_pipeline/stage_write.cc:51-84: kDither[i] = ((bayer8x8 + 0.5) / 64) - 0.5. The 8x8 Bayer
  // index is 16 * f(x0, y0) + 4 * f(x1, y1) + f(x2, y2) with f(a, b) = 2 * (a ^ b) + b over the bits of x, y.
  const uint32_t b = y & 7, t = (x ^ y) & 7;
  const uint32_t bayer = ((t & 1) << 5) | ((b & 1) << 4) | ((t & 2) << 2) | ((b & 2) << 1) | ((t & 4) >> 1) | ((b & 4) >> 2);
  return (static_cast<float>(bayer) + 0.5f) * (1.0f / 64.0f) - 0.5f;
}

JXLB_HD float DevIntToFloat(int32_t in, int bits, int exp_bits) {
  union { uint32_t u; float f; } cv;
  if (bits == 32) {
    cv.u = static_cast<uint32_t>(in);
    return cv.f;
  }
  const int exp_bias = (1 << (exp_bits - 1)) - 1;
  uint32_t f = static_cast<uint32_t>(in);
  const uint32_t signbit = (f >> (bits - 1)) & 1;
  f &= (1u << (bits - 1)) - 1;
  if (f == 0) return signbit ? -0.0f : 0.0f;
  const int mant_bits = bits - exp_bits - 1;
  int exp = static_cast<int>(f >> mant_bits);
  int mantissa = static_cast<int>((f & ((1u << mant_bits) - 1)) << (23 - mant_bits));
  if (exp == 0 && exp_bits < 8) {
    exp = 1;
    while ((mantissa & 0x800000) == 0) {
      mantissa <<= 1;
      exp--;
    }
    mantissa &= 0x7fffff;
  }
  exp += 127 - exp_bias;
  cv.u = (signbit << 31) | (static_cast<uint32_t>(exp) << 23) | static_cast<uint32_t>(mantissa);
  return cv.f;
}

JXLB_HD uint16_t DevFloatToHalf(float f) {
  union { uint32_t u; float f; } cv;
  cv.f = f;
  const uint32_t b = cv.u;
  const uint32_t sign = (b >> 16) & 0x8000;
  const int32_t exp = static_cast<int32_t>((b >> 23) & 0xFF) - 127 + 15;
  uint32_t mant = b & 0x7FFFFF;
  if (((b >> 23) & 0xFF) == 0xFF) return static_cast<uint16_t>(sign | 0x7C00 | (mant ? 0x200 : 0));
  if (exp >= 31) return static_cast<uint16_t>(sign | 0x7C00);
  if (exp <= 0) {
    if (exp < -10) return static_cast<uint16_t>(sign);
    mant |= 0x800000;
    const uint32_t shift = 14 - exp;
    uint32_t h = mant >> shift;
    const uint32_t rem = mant & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) h++;
    return static_cast<uint16_t>(sign | h);
  }
  uint32_t h = (static_cast<uint32_t>(exp) << 10) | (mant >> 13);
  const uint32_t rem = mant & 0x1FFF;
  if (rem > 0x1000 || (rem == 0x1000 && (h & 1))) h++;
  return static_cast<uint16_t>(sign | h);
}

JXLB_HD int32_t DevRoundHalfEven(float v) {
#if defined(__CUDA_ARCH__)
  return __float2int_rn(v);
#else
  // Host compile (tests only): nearbyint under the default rounding mode.
  return static_cast<int32_t>(__builtin_nearbyintf(v));
#endif
}

// The float sample of output channel c at (x, y).
JXLB_HD float DevSampleFloat(const DevPools& P, const DevFrameOut& fo, uint32_t c, uint32_t x, uint32_t y) {
  if (fo.plane[c] == kNoPlane) return 1.0f;
  const DevPlane pl = P.planes[fo.plane[c]];
  const int32_t v = P.arena[pl.off + static_cast<size_t>(y) * pl.w + x];
  if (fo.is_float[c]) return DevIntToFloat(v, fo.is_float[c] & 0xFF, fo.is_float[c] >> 8);
#if defined(__CUDA_ARCH__)
  return __fmul_rn(static_cast<float>(v), fo.factor[c]);
#else
  return static_cast<float>(v) * fo.factor[c];
#endif
}

// FastErff, lib/jxl/base/fast_math-inl.h:126-146.
JXLB_HD float DevFastErff(float x) {
  const bool xle0 = x <= 0.0f;
  const float absx = fabsf(x);
  const float d1 = fmaf(absx, 7.77394369e-02f, 2.05260015e-04f);
  const float d2 = fmaf(d1, absx, 2.32120216e-01f);
  const float d3 = fmaf(d2, absx, 2.77820801e-01f);
  const float d4 = fmaf(d3, absx, 1.0f);
  const float d5 = d4 * d4;
  const float inv = 1.0f / d5;
  const float r = fmaf(-inv, inv, 1.0f);
  return xle0 ? -r : r;
}

// The splines' contribution to pixel (x, y) of a Modular frame, added to the three colour samples v[0..2] in the draw
// order of the row's segment list (lib/jxl/splines.cc:78-113 DrawSegment, :167-175 DrawSegments): every segment whose
// column span holds x adds colour * sigma / 4 * intensity * (erf(...) - erf(...))^2.
JXLB_HD void DevSplineAdd(const uint32_t* rows, const uint32_t* idx, const float* seg, uint32_t xsize, uint32_t x, uint32_t y,
                          float* v) {
  const float fx = static_cast<float>(static_cast<int32_t>(x)), fy = static_cast<float>(y);
  for (uint32_t i = rows[y]; i < rows[y + 1]; i++) {
    const float* s = seg + static_cast<size_t>(idx[i]) * kSplineSegmentWords;
    union { float f; int32_t i; } xa, xb;
    xa.f = s[7];
    xb.f = s[8];
    const int32_t lo = xa.i > 0 ? xa.i : 0, hi = xb.i < static_cast<int32_t>(xsize) ? xb.i : static_cast<int32_t>(xsize);
    if (static_cast<int32_t>(x) < lo || static_cast<int32_t>(x) >= hi) continue;
    const float dx = fx - s[0], dy = fy - s[1];
    const float distance = sqrtf(fmaf(dx, dx, dy * dy));
    const float f = DevFastErff(fmaf(distance, 0.5f, 0.353553391f) * s[2]) - DevFastErff(fmaf(distance, 0.5f, -0.353553391f) * s[2]);
    const float li = s[3] * (f * f);
    v[0] = fmaf(s[4], li, v[0]);
    v[1] = fmaf(s[5], li, v[1]);
    v[2] = fmaf(s[6], li, v[2]);
  }
}

JXLB_HD void DevSplinePixel(const DevPools& P, const DevFrameOut& fo, uint32_t x, uint32_t y, float* v) {
  DevSplineAdd(P.spl_idx + fo.spl_rows, P.spl_idx + fo.spl_idx, P.spl_seg + fo.spl_seg, fo.xsize, x, y, v);
}

// Converts and stores one pixel (all channels).
JXLB_HD void DevWritePixel(const DevPools& P, const DevFrameOut& fo, uint8_t* out, uint32_t x, uint32_t y) {
  uint32_t dx = x, dy = y, orow = y, ocol = x;  // dither position, store position
  if (fo.orient != 0) DevOrient(fo.orient, fo.xsize, fo.ysize, x, y, &dx, &dy, &orow, &ocol);
  uint8_t* row = out + fo.out_off + fo.stride * orow;
  float spl[3] = {0.0f, 0.0f, 0.0f};
  if (fo.has_splines) {  // (three colour channels: checked by the planner)
    for (uint32_t c = 0; c < 3; c++) spl[c] = DevSampleFloat(P, fo, c, x, y);
    DevSplinePixel(P, fo, x, y, spl);
  }
  for (uint32_t c = 0; c < fo.num_channels; c++) {
    float v = (fo.has_splines && c < 3) ? spl[c] : DevSampleFloat(P, fo, c, x, y);
    const size_t idx = static_cast<size_t>(ocol) * fo.num_channels + c;
    if (fo.data_type == 2 || fo.data_type == 3) {
      const float mul = fo.data_type == 2 ? 255.0f : 65535.0f;
#if defined(__CUDA_ARCH__)
      v = __fmul_rn(v, mul);
      if (fo.data_type == 2) v = __fadd_rn(v, DevDither(dx, dy));
#else
      v = v * mul;
      if (fo.data_type == 2) v = v + DevDither(dx, dy);
#endif
      if (!(v >= 0.0f)) v = 0.0f;
      if (v > mul) v = mul;
      const int32_t r = DevRoundHalfEven(v);
      if (fo.data_type == 2) {
        row[idx] = static_cast<uint8_t>(r);
      } else {
        uint16_t u = static_cast<uint16_t>(r);
        if (fo.big_endian) u = static_cast<uint16_t>((u >> 8) | (u << 8));
        row[idx * 2] = static_cast<uint8_t>(u & 0xFF);
        row[idx * 2 + 1] = static_cast<uint8_t>(u >> 8);
      }
    } else if (fo.data_type == 5) {
      uint16_t u = DevFloatToHalf(v);
      if (fo.big_endian) u = static_cast<uint16_t>((u >> 8) | (u << 8));
      row[idx * 2] = static_cast<uint8_t>(u & 0xFF);
      row[idx * 2 + 1] = static_cast<uint8_t>(u >> 8);
    } else {
      union { uint32_t u; float f; } cv;
      cv.f = v;
      uint32_t u = cv.u;
      if (fo.big_endian) u = (u >> 24) | ((u >> 8) & 0xFF00) | ((u << 8) & 0xFF0000) | (u << 24);
      row[idx * 4] = static_cast<uint8_t>(u);
      row[idx * 4 + 1] = static_cast<uint8_t>(u >> 8);
      row[idx * 4 + 2] = static_cast<uint8_t>(u >> 16);
      row[idx * 4 + 3] = static_cast<uint8_t>(u >> 24);
    }
  }
}

}  // namespace jxlb

#endif  // JXLB_FINISH_DEV_H_
