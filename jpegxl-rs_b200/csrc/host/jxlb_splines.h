// jxl_b200 host side of the splines: the quantised splines of DC global, their dequantisation
// and the draw cache -- one segment per pixel of arc length, a list of segments per image row -- which the output
// kernel then evaluates per pixel (kernels/jxlb_finish_dev.h DevSplinePixel). O(KB) of bit-serial and float work per
// frame, like the rest of the planner. Restates libjxl 0.11.2:
//   lib/jxl/splines.cc:277-308 DecodeAllStartingPoints, :310-366 DrawCentripetalCatmullRomSpline, :368-400
//   ForEachEquallySpacedPoint, :456-545 QuantizedSpline::Dequantize, :547-596 QuantizedSpline::Decode, :607-650
//   Splines::Decode, :667-735 Splines::InitializeDrawCache, :43-76 ContinuousIDCT (8 float lanes, AVX2 SumOfLanes
//   order), :115-165 ComputeSegments (JXL_HIGH_PRECISION = 1), :177-200 SegmentsFromPoints;
//   lib/jxl/base/fast_math-inl.h:94-122 FastCosf, :126-146 FastErff.
// (Same text as the oracle's restatement up to the prefix, like the other host parsers: both are pinned on the
// reference's fixture 2bit.jxl -- the decoded drawing -- not on each other.)
#ifndef JXLB_SPLINES_H_
#define JXLB_SPLINES_H_

#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>

#include "jxlb_entropy.h"

namespace jxlb {

struct SplinePoint {
  float x = 0, y = 0;
};

struct QuantizedSpline {
  std::vector<std::pair<int64_t, int64_t>> control_points;  // double deltas
  int color_dct[3][32] = {};
  int sigma_dct[32] = {};
};

struct SplineSegment {
  float center_x, center_y, maximum_distance, inv_sigma, sigma_over_4_times_intensity, color[3];
};

struct SplineState {
  int32_t quantization_adjustment = 0;
  std::vector<SplinePoint> starting_points;
  std::vector<QuantizedSpline> splines;
  // draw cache
  std::vector<SplineSegment> segments;
  std::vector<uint32_t> segment_indices, segment_y_start;
  bool Any() const { return !splines.empty(); }
};

enum SplineContext { kSplQuantizationAdjustment = 0, kSplStartingPosition, kSplNumSplines, kSplNumControlPoints, kSplControlPoints,
                     kSplDCT, kNumSplineContexts };

inline void CheckSplinePos(int64_t x, int64_t y) {
  const int64_t lim = int64_t{1} << 23;
  JXLB_CHECK(x < lim && x > -lim && y < lim && y > -lim, "spline coordinates out of bounds");
}

inline void ReadSplines(BitReader& br, size_t num_pixels, SplineState* s) {
  EntropyCode code;
  ReadEntropyCode(br, kNumSplineContexts, &code);
  SymbolReader reader(&code, br);
  size_t num_splines = reader.ReadUint(kSplNumSplines, br);
  const size_t max_control_points = std::min<size_t>(size_t{1} << 20, num_pixels / 2);
  JXLB_CHECK(num_splines <= max_control_points && num_splines + 1 <= max_control_points, "too many splines");
  num_splines++;
  int64_t last_x = 0, last_y = 0;
  for (size_t i = 0; i < num_splines; i++) {
    int64_t x = reader.ReadUint(kSplStartingPosition, br);
    int64_t y = reader.ReadUint(kSplStartingPosition, br);
    if (i != 0) {
      x = UnpackSigned(static_cast<uint32_t>(x)) + last_x;
      y = UnpackSigned(static_cast<uint32_t>(y)) + last_y;
    }
    CheckSplinePos(x, y);
    SplinePoint p;
    p.x = static_cast<float>(x);
    p.y = static_cast<float>(y);
    s->starting_points.push_back(p);
    last_x = x;
    last_y = y;
  }
  s->quantization_adjustment = UnpackSigned(reader.ReadUint(kSplQuantizationAdjustment, br));
  size_t total_control_points = num_splines;
  for (size_t i = 0; i < num_splines; i++) {
    QuantizedSpline q;
    const size_t n = reader.ReadUint(kSplNumControlPoints, br);
    JXLB_CHECK(n <= max_control_points, "too many control points");
    total_control_points += n;
    JXLB_CHECK(total_control_points <= max_control_points, "too many control points");
    q.control_points.resize(n);
    const int64_t delta_limit = int64_t{1} << 30;
    for (auto& cp : q.control_points) {
      cp.first = UnpackSigned(reader.ReadUint(kSplControlPoints, br));
      cp.second = UnpackSigned(reader.ReadUint(kSplControlPoints, br));
      JXLB_CHECK(cp.first < delta_limit && cp.first > -delta_limit && cp.second < delta_limit && cp.second > -delta_limit,
                 "spline delta-delta is out of bounds");
    }
    auto decode_dct = [&](int dct[32]) {
      for (int k = 0; k < 32; k++) {
        dct[k] = UnpackSigned(reader.ReadUint(kSplDCT, br));
        JXLB_CHECK(dct[k] != INT32_MIN, "the weird number in spline DCT");
      }
    };
    for (auto& dct : q.color_dct) decode_dct(dct);
    decode_dct(q.sigma_dct);
    s->splines.push_back(std::move(q));
  }
  JXLB_CHECK(reader.FinalStateOk(), "splines: bad ANS final state");
}

// ---- fast math (lib/jxl/base/fast_math-inl.h)
inline float SplFastCosf(float x) {
  const double kPi = 3.14159265358979323846264338327950288;
  const float pi2 = static_cast<float>(kPi * 2.0f), pi2_inv = static_cast<float>(0.5f / kPi);
  const float npi2 = std::floor(x * pi2_inv) * pi2;
  const float xmodpi2 = x - npi2;
  const float x_pi = std::min(xmodpi2, pi2 - xmodpi2);
  const bool above_pihalf = x_pi >= static_cast<float>(kPi / 2.0f);
  const float x_pihalf = above_pihalf ? static_cast<float>(kPi) - x_pi : x_pi;
  const float xs = x_pihalf * 0.25f;
  const float x2 = xs * xs;
  const float x4 = x2 * x2;
  const float pre = std::fmaf(x4, 0.06960438f, std::fmaf(x2, -0.84087373f, 1.68179268f));
  const float s1 = std::fmaf(pre, pre, -1.414213562f);
  const float s2 = std::fmaf(s1, s1, -1.0f);
  return above_pihalf ? -s2 : s2;
}

inline float SplFastErff(float x) {
  const bool xle0 = x <= 0.0f;
  const float absx = std::fabs(x);
  const float d1 = std::fmaf(absx, 7.77394369e-02f, 2.05260015e-04f);
  const float d2 = std::fmaf(d1, absx, 2.32120216e-01f);
  const float d3 = std::fmaf(d2, absx, 2.77820801e-01f);
  const float d4 = std::fmaf(d3, absx, 1.0f);
  const float d5 = d4 * d4;
  const float inv = 1.0f / d5;
  const float r = std::fmaf(-inv, inv, 1.0f);
  return xle0 ? -r : r;
}

inline float SplContinuousIDCT(const float dct[32], float t) {
  const double kPi = 3.14159265358979323846264338327950288;
  const float kSqrt2 = 1.41421356237f;
  float lanes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const float tandhalf = t + 0.5f;
  for (int i = 0; i < 32; i += 8)
    for (int l = 0; l < 8; l++) {
      const float mult = static_cast<float>(kPi / 32 * (i + l));
      const float c = SplFastCosf(mult * tandhalf);
      const float local = dct[i + l] * c;
      lanes[l] = std::fmaf(kSqrt2, local, lanes[l]);
    }
  return ((lanes[0] + lanes[4]) + (lanes[2] + lanes[6])) + ((lanes[1] + lanes[5]) + (lanes[3] + lanes[7]));
}

// ---- geometry
inline void SplCatmullRom(std::vector<SplinePoint> points, std::vector<SplinePoint>* result) {
  if (points.empty()) return;
  if (points.size() == 1) {
    result->push_back(points[0]);
    return;
  }
  const int kNumPoints = 16;
  auto P = [](float x, float y) {
    SplinePoint p;
    p.x = x;
    p.y = y;
    return p;
  };
  points.insert(points.begin(), P(points[0].x + (points[0].x - points[1].x), points[0].y + (points[0].y - points[1].y)));
  {
    const SplinePoint a = points[points.size() - 1], b = points[points.size() - 2];
    points.push_back(P(a.x + (a.x - b.x), a.y + (a.y - b.y)));
  }
  for (size_t start = 0; start < points.size() - 3; ++start) {
    const SplinePoint* p = &points[start];
    result->push_back(p[1]);
    float d[3], t[4];
    t[0] = 0;
    for (int k = 0; k < 3; ++k) {
      d[k] = std::sqrt(hypotf(p[k + 1].x - p[k].x, p[k + 1].y - p[k].y));
      t[k + 1] = t[k] + d[k];
    }
    for (int i = 1; i < kNumPoints; ++i) {
      const float tt = d[0] + (static_cast<float>(i) / kNumPoints) * d[1];
      SplinePoint a[3];
      for (int k = 0; k < 3; ++k) {
        const float f = (tt - t[k]) / d[k];
        a[k] = P(p[k].x + f * (p[k + 1].x - p[k].x), p[k].y + f * (p[k + 1].y - p[k].y));
      }
      SplinePoint b[2];
      for (int k = 0; k < 2; ++k) {
        const float f = (tt - t[k]) / (d[k] + d[k + 1]);
        b[k] = P(a[k].x + f * (a[k + 1].x - a[k].x), a[k].y + f * (a[k + 1].y - a[k].y));
      }
      const float f = (tt - t[1]) / d[1];
      result->push_back(P(b[0].x + f * (b[1].x - b[0].x), b[0].y + f * (b[1].y - b[0].y)));
    }
  }
  result->push_back(points[points.size() - 2]);
}

inline void SplEquallySpaced(const std::vector<SplinePoint>& points, std::vector<std::pair<SplinePoint, float>>* out) {
  const float kDist = 1.0f;  // kDesiredRenderingDistance
  SplinePoint current = points.front();
  out->emplace_back(current, kDist);
  size_t next = 0;
  while (next != points.size()) {
    const SplinePoint* previous = &current;
    float arclength_from_previous = 0.f;
    for (;;) {
      if (next == points.size()) {
        out->emplace_back(*previous, arclength_from_previous);
        return;
      }
      const float dx = points[next].x - previous->x, dy = points[next].y - previous->y;
      const float arclength_to_next = std::sqrt(dx * dx + dy * dy);
      if (arclength_from_previous + arclength_to_next >= kDist) {
        const float f = (kDist - arclength_from_previous) / arclength_to_next;
        SplinePoint c;
        c.x = previous->x + f * dx;
        c.y = previous->y + f * dy;
        current = c;
        out->emplace_back(current, kDist);
        break;
      }
      arclength_from_previous += arclength_to_next;
      previous = &points[next];
      ++next;
    }
  }
}

inline uint32_t SplCeilLog2Nonzero(uint64_t v) {
  uint32_t f = 63 - __builtin_clzll(v);
  return (v & (v - 1)) ? f + 1 : f;
}

// Splines::InitializeDrawCache: y_to_x / y_to_b = ColorCorrelation::YtoXRatio(0) / YtoBRatio(0) (the base correlations).
inline void InitSplineDrawCache(size_t image_xsize, size_t image_ysize, float y_to_x, float y_to_b, SplineState* s) {
  const float kChannelWeight[4] = {0.0042f, 0.075f, 0.07f, .3333f};
  const float kSqrt0_5 = 0.70710678118f;
  std::vector<std::pair<size_t, size_t>> segments_by_y;
  uint64_t total_estimated_area_reached = 0;
  const uint64_t kOne = 1, image_size = static_cast<uint64_t>(image_xsize) * image_ysize;
  const uint64_t area_limit = std::min(1024 * image_size + (kOne << 32), kOne << 42);
  const float inv_quant = s->quantization_adjustment >= 0 ? 1.f / (1.f + .125f * s->quantization_adjustment)
                                                          : (1.f - .125f * s->quantization_adjustment);
  for (size_t si = 0; si < s->splines.size(); si++) {
    const QuantizedSpline& q = s->splines[si];
    // ---- Dequantize
    std::vector<SplinePoint> control_points;
    const float px = std::roundf(s->starting_points[si].x), py = std::roundf(s->starting_points[si].y);
    CheckSplinePos(static_cast<int64_t>(px), static_cast<int64_t>(py));
    int current_x = static_cast<int>(px), current_y = static_cast<int>(py);
    auto push = [&](int x, int y) {
      SplinePoint p;
      p.x = static_cast<float>(x);
      p.y = static_cast<float>(y);
      control_points.push_back(p);
    };
    push(current_x, current_y);
    int current_delta_x = 0, current_delta_y = 0;
    uint64_t manhattan_distance = 0;
    for (const auto& point : q.control_points) {
      current_delta_x += static_cast<int>(point.first);
      current_delta_y += static_cast<int>(point.second);
      manhattan_distance += std::abs(current_delta_x) + std::abs(current_delta_y);
      JXLB_CHECK(manhattan_distance <= area_limit, "spline: too large manhattan distance");
      CheckSplinePos(current_delta_x, current_delta_y);
      current_x += current_delta_x;
      current_y += current_delta_y;
      CheckSplinePos(current_x, current_y);
      push(current_x, current_y);
    }
    float color_dct[3][32], sigma_dct[32];
    for (int c = 0; c < 3; ++c)
      for (int i = 0; i < 32; ++i) {
        const float inv_dct_factor = (i == 0) ? kSqrt0_5 : 1.0f;
        color_dct[c][i] = q.color_dct[c][i] * inv_dct_factor * kChannelWeight[c] * inv_quant;
      }
    for (int i = 0; i < 32; ++i) {
      color_dct[0][i] += y_to_x * color_dct[1][i];
      color_dct[2][i] += y_to_b * color_dct[1][i];
    }
    uint64_t width_estimate = 0;
    uint64_t color[3] = {};
    for (int c = 0; c < 3; ++c)
      for (int i = 0; i < 32; ++i) color[c] += static_cast<uint64_t>(std::ceil(inv_quant * std::abs(q.color_dct[c][i])));
    color[0] += static_cast<uint64_t>(std::ceil(std::abs(y_to_x))) * color[1];
    color[2] += static_cast<uint64_t>(std::ceil(std::abs(y_to_b))) * color[1];
    const uint64_t max_color = std::max({color[1], color[0], color[2]});
    const uint64_t logcolor = std::max<uint64_t>(kOne, SplCeilLog2Nonzero(kOne + max_color));
    const float weight_limit =
        std::ceil(std::sqrt((static_cast<float>(area_limit) / logcolor) / std::max<size_t>(1, manhattan_distance)));
    for (int i = 0; i < 32; ++i) {
      const float inv_dct_factor = (i == 0) ? kSqrt0_5 : 1.0f;
      sigma_dct[i] = q.sigma_dct[i] * inv_dct_factor * kChannelWeight[3] * inv_quant;
      const float weight_f = std::ceil(inv_quant * std::abs(q.sigma_dct[i]));
      const uint64_t weight = static_cast<uint64_t>(std::min(weight_limit, std::max(1.0f, weight_f)));
      width_estimate += weight * weight * logcolor;
    }
    total_estimated_area_reached += width_estimate * manhattan_distance;
    JXLB_CHECK(total_estimated_area_reached <= area_limit, "spline: too large total estimated area");
    for (size_t k = 0; k + 1 < control_points.size(); k++)
      JXLB_CHECK(!(std::fabs(control_points[k].x - control_points[k + 1].x) < 1e-3f &&
                   std::fabs(control_points[k].y - control_points[k + 1].y) < 1e-3f),
                 "identical successive control points in spline");
    // ---- points to draw, segments
    std::vector<SplinePoint> intermediate;
    SplCatmullRom(control_points, &intermediate);
    std::vector<std::pair<SplinePoint, float>> points_to_draw;
    SplEquallySpaced(intermediate, &points_to_draw);
    const float arc_length = (points_to_draw.size() - 2) * 1.0f + points_to_draw.back().second;
    if (arc_length <= 0.f) continue;
    const float inv_arc_length = 1.0f / arc_length;
    int k = 0;
    for (const auto& ptd : points_to_draw) {
      const SplinePoint& point = ptd.first;
      const float multiplier = ptd.second;
      const float progress_along_arc = std::min(1.f, (k * 1.0f) * inv_arc_length);
      ++k;
      float col[3];
      for (size_t c = 0; c < 3; ++c) col[c] = SplContinuousIDCT(color_dct[c], (32 - 1) * progress_along_arc);
      const float sigma = SplContinuousIDCT(sigma_dct, (32 - 1) * progress_along_arc);
      // ComputeSegments
      if (!(std::isfinite(sigma) && sigma != 0.0f && std::isfinite(1.0f / sigma) && std::isfinite(multiplier))) continue;
      const float kDistanceExp = 5;
      float max_color_f = 0.01f;
      for (size_t c = 0; c < 3; c++) max_color_f = std::max(max_color_f, std::abs(col[c] * multiplier));
      const float maximum_distance = std::sqrt(-2 * sigma * sigma * (std::log(0.1) * kDistanceExp - std::log(max_color_f)));
      SplineSegment seg;
      seg.center_y = point.y;
      seg.center_x = point.x;
      for (int c = 0; c < 3; c++) seg.color[c] = col[c];
      seg.inv_sigma = 1.0f / sigma;
      seg.sigma_over_4_times_intensity = .25f * sigma * multiplier;
      seg.maximum_distance = maximum_distance;
      const int64_t y0 = std::llround(point.y - maximum_distance);
      const int64_t y1 = std::llround(point.y + maximum_distance) + 1;
      for (int64_t y = std::max<int64_t>(y0, 0); y < y1; y++) segments_by_y.emplace_back(static_cast<size_t>(y), s->segments.size());
      s->segments.push_back(seg);
    }
  }
  std::sort(segments_by_y.begin(), segments_by_y.end());
  s->segment_indices.resize(segments_by_y.size());
  s->segment_y_start.assign(image_ysize + 1, 0);
  for (size_t i = 0; i < segments_by_y.size(); i++) {
    s->segment_indices[i] = static_cast<uint32_t>(segments_by_y[i].second);
    const size_t y = segments_by_y[i].first;
    if (y < image_ysize) s->segment_y_start[y + 1]++;
  }
  for (size_t y = 0; y < image_ysize; y++) s->segment_y_start[y + 1] += s->segment_y_start[y];
}

// DrawSegment for one pixel of row y: the contribution of `seg` is added to v[0..2] when x lies in the segment's span.
inline void SplDrawPixel(const SplineSegment& seg, int64_t x, size_t y, int64_t x0, int64_t x1, float* v0, float* v1, float* v2) {
  const int64_t xa = std::max<int64_t>(x0, std::llround(seg.center_x - seg.maximum_distance));
  const int64_t xb = std::min<int64_t>(x1, std::llround(seg.center_x + seg.maximum_distance) + 1);
  if (x < xa || x >= xb) return;
  const float dx = static_cast<float>(x) - seg.center_x;
  const float dy = static_cast<float>(y) - seg.center_y;
  const float sqd = std::fmaf(dx, dx, dy * dy);
  const float distance = std::sqrt(sqd);
  const float one_over_2s2 = 0.353553391f;
  const float f = SplFastErff(std::fmaf(distance, 0.5f, one_over_2s2) * seg.inv_sigma) -
                  SplFastErff(std::fmaf(distance, 0.5f, -one_over_2s2) * seg.inv_sigma);
  const float local_intensity = seg.sigma_over_4_times_intensity * (f * f);
  *v0 = std::fmaf(seg.color[0], local_intensity, *v0);
  *v1 = std::fmaf(seg.color[1], local_intensity, *v1);
  *v2 = std::fmaf(seg.color[2], local_intensity, *v2);
}

}  // namespace jxlb

#endif  // JXLB_SPLINES_H_
