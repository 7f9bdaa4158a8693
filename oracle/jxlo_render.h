// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
//
// Render stages applied to a decoded frame, in the order of
// PassesDecoderState::PreparePipeline (lib/jxl/dec_cache.cc:103-347):
//   Gaborish          lib/jxl/render_pipeline/stage_gaborish.cc:22-100
//   EPF 0 / 1 / 2     lib/jxl/render_pipeline/stage_epf.cc:43-500, lib/jxl/epf.cc
//   patches           lib/jxl/dec_patch_dictionary.cc:28-200, :316-360, lib/jxl/blending.cc:42-160
//   XYB -> linear     lib/jxl/render_pipeline/stage_xyb.cc, lib/jxl/dec_xyb-inl.h:37-83,
//                     lib/jxl/dec_xyb.cc:192-330 (OutputEncodingInfo)
//   YCbCr -> RGB      lib/jxl/render_pipeline/stage_ycbcr.cc
//   from linear       lib/jxl/render_pipeline/stage_from_linear.cc, lib/jxl/cms/transfer_functions-inl.h:219-242
// Whole-frame restatement of the row-based pipeline: borders are mirrored at the frame
// size (lib/jxl/render_pipeline/low_memory_render_pipeline.cc:466-492, lib/jxl/image_ops.h:184-195).
#ifndef JXLO_RENDER_H_
#define JXLO_RENDER_H_

#include "jxlo_frame.h"
#include "jxlo_vardct.h"
#include "jxlo_splines.h"

namespace jxlo {

// ---------------------------------------------------------------- patches
enum PatchBlendMode { kPatchNone = 0, kPatchReplace = 1, kPatchAdd = 2, kPatchMul = 3, kPatchBlendAbove = 4,
                      kPatchBlendBelow = 5, kPatchAlphaWeightedAddAbove = 6, kPatchAlphaWeightedAddBelow = 7 };
struct PatchBlending { uint32_t mode = 0, alpha_channel = 0; bool clamp = false; };
struct PatchRef { uint32_t ref, x0, y0, xsize, ysize; };
struct PatchPos { uint32_t x, y, ref_idx; };
struct FeatureState {
  std::vector<PatchRef> refs;
  std::vector<PatchPos> positions;
  std::vector<PatchBlending> blendings;  // (1 + num_extra) per position
  size_t blend_stride = 1;
  SplineState splines;
};

inline void ReadPatches(BitReader& br, const FrameDimensions& dim, const ImageMetadata& meta, const CodestreamState& cs,
                        FeatureState* f) {
  const size_t num_ec = meta.extra.size();
  f->blend_stride = num_ec + 1;
  EntropyCode code;
  ReadEntropyCode(br, 10, &code);
  SymbolReader reader(&code, br);
  auto read_num = [&](uint32_t ctx) { return reader.ReadUint(ctx, br); };
  const size_t xsize = dim.xsize_padded, ysize = dim.ysize_padded;
  const size_t num_ref_patch = read_num(0);
  const size_t max_ref_patches = 1024 + xsize * ysize / 4;
  const size_t max_patches = max_ref_patches * 4;
  JXLO_CHECK(num_ref_patch <= max_ref_patches, "too many patches");
  size_t total = 0;
  for (size_t id = 0; id < num_ref_patch; id++) {
    PatchRef rp;
    rp.ref = read_num(1);
    JXLO_CHECK(rp.ref < 4 && cs.reference_valid[rp.ref] && !cs.reference[rp.ref].planes.empty(), "invalid patch reference frame");
    JXLO_CHECK(cs.reference[rp.ref].is_xyb, "patches cannot use frames saved post colour transform");
    const Plane& rpl = cs.reference[rp.ref].planes[0];
    rp.x0 = read_num(3);
    rp.y0 = read_num(3);
    rp.xsize = read_num(2) + 1;
    rp.ysize = read_num(2) + 1;
    JXLO_CHECK(rp.x0 + rp.xsize <= static_cast<size_t>(rpl.w) && rp.y0 + rp.ysize <= static_cast<size_t>(rpl.h),
               "invalid patch position in reference frame");
    size_t id_count = read_num(7);
    JXLO_CHECK(id_count <= max_patches, "too many patches");
    id_count++;
    total += id_count;
    JXLO_CHECK(total <= max_patches, "too many patches");
    const bool choose_alpha = num_ec > 1;
    for (size_t i = 0; i < id_count; i++) {
      PatchPos pos;
      pos.ref_idx = f->refs.size();
      if (i == 0) {
        pos.x = read_num(4);
        pos.y = read_num(4);
      } else {
        const int64_t dx = UnpackSigned(read_num(6));
        JXLO_CHECK(!(dx < 0 && static_cast<uint64_t>(-dx) > f->positions.back().x), "negative patch x");
        pos.x = f->positions.back().x + dx;
        const int64_t dy = UnpackSigned(read_num(6));
        JXLO_CHECK(!(dy < 0 && static_cast<uint64_t>(-dy) > f->positions.back().y), "negative patch y");
        pos.y = f->positions.back().y + dy;
      }
      JXLO_CHECK(pos.x + rp.xsize <= xsize && pos.y + rp.ysize <= ysize, "patch outside the frame");
      for (size_t j = 0; j < f->blend_stride; j++) {
        PatchBlending info;
        info.mode = read_num(5);
        JXLO_CHECK(info.mode < 8, "invalid patch blend mode");
        const bool uses_alpha = info.mode >= kPatchBlendAbove;
        if (uses_alpha && choose_alpha) {
          info.alpha_channel = read_num(8);
          JXLO_CHECK(info.alpha_channel < num_ec, "invalid patch alpha channel");
        }
        if (uses_alpha || info.mode == kPatchMul) info.clamp = read_num(9) != 0;
        f->blendings.push_back(info);
      }
      f->positions.push_back(pos);
    }
    f->refs.push_back(rp);
  }
  JXLO_CHECK(reader.FinalStateOk(), "patches: bad ANS final state");
}

inline void ApplyPatches(const FeatureState& f, const CodestreamState& cs, std::vector<Plane>* planes) {
  const size_t nplanes = planes->size();
  for (size_t p = 0; p < f.positions.size(); p++) {
    const PatchPos& pos = f.positions[p];
    const PatchRef& rp = f.refs[pos.ref_idx];
    const FrameBuffer& ref = cs.reference[rp.ref];
    for (size_t c = 0; c < nplanes; c++) {
      const PatchBlending& b = f.blendings[p * f.blend_stride + (c < 3 ? 0 : 1 + (c - 3))];
      JXLO_CHECK(b.mode <= kPatchMul, "alpha patch blending is not supported by the oracle");
      if (b.mode == kPatchNone) continue;
      Plane& out = (*planes)[c];
      const Plane& fg = ref.planes[c];
      for (uint32_t iy = 0; iy < rp.ysize; iy++) {
        const uint32_t y = pos.y + iy;
        if (y >= static_cast<uint32_t>(out.h)) continue;
        for (uint32_t ix = 0; ix < rp.xsize; ix++) {
          const uint32_t x = pos.x + ix;
          if (x >= static_cast<uint32_t>(out.w)) continue;
          const float v = fg.Row(rp.y0 + iy)[rp.x0 + ix];
          float& o = out.Row(y)[x];
          if (b.mode == kPatchReplace) o = v;
          else if (b.mode == kPatchAdd) o = o + v;
          else o = o * (b.clamp ? std::min(1.0f, std::max(0.0f, v)) : v);
        }
      }
    }
  }
}

// ---------------------------------------------------------------- loop filters
inline int64_t Mirror(int64_t x, int64_t xsize) {
  while (x < 0 || x >= xsize) x = x < 0 ? -x - 1 : 2 * xsize - 1 - x;
  return x;
}

struct MirroredPlane {
  const Plane& p;
  float At(int64_t x, int64_t y) const { return p.Row(Mirror(y, p.h))[Mirror(x, p.w)]; }
};

inline void Gaborish(const LoopFilter& lf, Plane planes[3]) {
  const float gw[3][2] = {{lf.gab_x_weight1, lf.gab_x_weight2}, {lf.gab_y_weight1, lf.gab_y_weight2},
                          {lf.gab_b_weight1, lf.gab_b_weight2}};
  for (int c = 0; c < 3; c++) {
    float w[3] = {1.0f, gw[c][0], gw[c][1]};
    const float div = w[0] + 4 * (w[1] + w[2]);
    const float mul = 1.0f / div;
    for (float& v : w) v *= mul;
    const Plane in = planes[c];
    const MirroredPlane m{in};
    for (int y = 0; y < in.h; y++) {
      float* out = planes[c].Row(y);
      for (int x = 0; x < in.w; x++) {
        const float sum0 = m.At(x, y);
        const float sum1 = (m.At(x - 1, y) + m.At(x + 1, y)) + (m.At(x, y - 1) + m.At(x, y + 1));
        const float sum2 = (m.At(x - 1, y - 1) + m.At(x + 1, y - 1)) + (m.At(x - 1, y + 1) + m.At(x + 1, y + 1));
        out[x] = std::fmaf(sum2, w[2], std::fmaf(sum1, w[1], sum0 * w[0]));
      }
    }
  }
}

constexpr float kMinSigma = -3.90524291751269967465540850526868f;

struct SigmaLookup {
  const VarDCTState* vs;      // null for Modular frames
  float modular_inv_sigma;    // kInvSigmaNum / epf_sigma_for_modular
  float At(int x, int y) const {
    if (!vs) return modular_inv_sigma;
    const size_t bx = std::min<size_t>(x / 8, vs->dim.xsize_blocks - 1), by = std::min<size_t>(y / 8, vs->dim.ysize_blocks - 1);
    return vs->inv_sigma[by * vs->dim.xsize_blocks + bx];
  }
};

inline float SadMul(int x, int y, float sm, float bsm) {
  const int iy = y % 8, ix = x % 8;
  if (iy == 0 || iy == 7) return bsm;
  return (ix == 0 || ix == 7) ? bsm : sm;
}

// stage: 0, 1 or 2
inline void EPFStage(int stage, const LoopFilter& lf, const SigmaLookup& sigma, Plane planes[3]) {
  const Plane in[3] = {planes[0], planes[1], planes[2]};
  const MirroredPlane m[3] = {{in[0]}, {in[1]}, {in[2]}};
  float sm;
  if (stage == 0) sm = lf.epf_pass0_sigma_scale * 1.65;
  else if (stage == 1) sm = 1.65f;
  else sm = lf.epf_pass2_sigma_scale * 1.65;
  const float bsm = sm * lf.epf_border_sad_mul;
  const int W = in[0].w, H = in[0].h;
  for (int y = 0; y < H; y++) {
    for (int x = 0; x < W; x++) {
      const float row_sigma = sigma.At(x, y);
      if (row_sigma < kMinSigma) continue;  // output = input
      const float inv_sigma = row_sigma * SadMul(x, y, sm, bsm);
      float w = 1.0f;
      float X = m[0].At(x, y), Y = m[1].At(x, y), B = m[2].At(x, y);
      auto add_pixel = [&](int dx, int dy, float sad) {
        float weight = std::fmaf(sad, inv_sigma, 1.0f);
        if (weight < 0.0f) weight = 0.0f;
        w = w + weight;
        X = std::fmaf(weight, m[0].At(x + dx, y + dy), X);
        Y = std::fmaf(weight, m[1].At(x + dx, y + dy), Y);
        B = std::fmaf(weight, m[2].At(x + dx, y + dy), B);
      };
      if (stage == 0) {
        static const int sads_off[12][2] = {{-2, 0}, {-1, -1}, {-1, 0}, {-1, 1}, {0, -2}, {0, -1},
                                            {0, 1},  {0, 2},   {1, -1}, {1, 0},  {1, 1},  {2, 0}};  // {row, col}
        static const int plus_off[5][2] = {{0, 0}, {-1, 0}, {0, -1}, {1, 0}, {0, 1}};
        float sads[12] = {0};
        for (int c = 0; c < 3; c++) {
          const float scale = lf.epf_channel_scale[c];
          for (int i = 0; i < 12; i++) {
            float sad = 0.0f;
            for (const auto& off : plus_off) {
              const float r11 = m[c].At(x + off[1], y + off[0]);
              const float c11 = m[c].At(x + sads_off[i][1] + off[1], y + sads_off[i][0] + off[0]);
              sad = sad + std::fabs(r11 - c11);
            }
            sads[i] = std::fmaf(sad, scale, sads[i]);
          }
        }
        for (int i = 0; i < 12; i++) add_pixel(sads_off[i][1], sads_off[i][0], sads[i]);
      } else if (stage == 1) {
        float sad0 = 0, sad1 = 0, sad2 = 0, sad3 = 0;
        for (int c = 0; c < 3; c++) {
          auto P = [&](int col, int row) { return m[c].At(x + col - 2, y + row - 2); };  // pCR: column C, row R
          const float p20 = P(2, 0), p21 = P(2, 1);
          float sad0c = std::fabs(p20 - p21);
          const float p11 = P(1, 1);
          float sad1c = std::fabs(p11 - p21);
          const float p31 = P(3, 1);
          float sad2c = std::fabs(p31 - p21);
          const float p02 = P(0, 2), p12 = P(1, 2);
          sad1c = sad1c + std::fabs(p02 - p12);
          sad0c = sad0c + std::fabs(p11 - p12);
          const float p22 = P(2, 2);
          float t = std::fabs(p12 - p22);
          sad1c = sad1c + t;
          sad2c = sad2c + t;
          t = std::fabs(p22 - p21);
          float sad3c = t;
          sad0c = sad0c + t;
          const float p32 = P(3, 2);
          sad0c = sad0c + std::fabs(p31 - p32);
          t = std::fabs(p22 - p32);
          sad1c = sad1c + t;
          sad2c = sad2c + t;
          const float p42 = P(4, 2);
          sad2c = sad2c + std::fabs(p42 - p32);
          const float p13 = P(1, 3);
          sad3c = sad3c + std::fabs(p13 - p12);
          const float p23 = P(2, 3);
          t = std::fabs(p22 - p23);
          sad0c = sad0c + t;
          sad3c = sad3c + t;
          sad1c = sad1c + std::fabs(p13 - p23);
          const float p33 = P(3, 3);
          sad2c = sad2c + std::fabs(p33 - p23);
          sad3c = sad3c + std::fabs(p33 - p32);
          const float p24 = P(2, 4);
          sad3c = sad3c + std::fabs(p24 - p23);
          const float scale = lf.epf_channel_scale[c];
          sad0 = std::fmaf(sad0c, scale, sad0);
          sad1 = std::fmaf(sad1c, scale, sad1);
          sad2 = std::fmaf(sad2c, scale, sad2);
          sad3 = std::fmaf(sad3c, scale, sad3);
        }
        add_pixel(0, -1, sad0);
        add_pixel(-1, 0, sad1);
        add_pixel(1, 0, sad2);
        add_pixel(0, 1, sad3);
      } else {
        const float rx = X, ry = Y, rb = B;
        auto add2 = [&](int dx, int dy) {
          const float cx = m[0].At(x + dx, y + dy), cy = m[1].At(x + dx, y + dy), cb = m[2].At(x + dx, y + dy);
          float sad = std::fabs(cx - rx) * lf.epf_channel_scale[0];
          sad = std::fmaf(std::fabs(cy - ry), lf.epf_channel_scale[1], sad);
          sad = std::fmaf(std::fabs(cb - rb), lf.epf_channel_scale[2], sad);
          add_pixel(dx, dy, sad);
        };
        add2(0, -1);
        add2(-1, 0);
        add2(1, 0);
        add2(0, 1);
      }
      const float inv_w = 1.0f / w;
      planes[0].Row(y)[x] = X * inv_w;
      planes[1].Row(y)[x] = Y * inv_w;
      planes[2].Row(y)[x] = B * inv_w;
    }
  }
}

// ---------------------------------------------------------------- colour
struct OutputColor {
  bool linear_srgb_fallback = false;  // XYB image whose tagged encoding cannot be produced
  float inverse_matrix[9];
  float opsin_biases[3], opsin_biases_cbrt[3];
  uint32_t tf = kTFSRGB;  // effective transfer function
  bool have_gamma = false;
  float inverse_gamma = 1.0f;
};

inline bool CanOutputTo(const ColorEncoding& c) {  // CanOutputToColorEncoding, dec_xyb.cc:205-220
  if (c.want_icc) return false;
  if (!c.have_gamma && c.transfer_function != kTFPQ && c.transfer_function != kTFSRGB && c.transfer_function != kTFLinear &&
      c.transfer_function != kTFHLG && c.transfer_function != kTFDCI && c.transfer_function != kTF709)
    return false;
  if (c.IsGray() && c.white_point != 1) return false;
  return true;
}

inline OutputColor MakeOutputColor(const ImageMetadata& meta) {  // OutputEncodingInfo, dec_xyb.cc:222-330
  OutputColor o;
  ColorEncoding c = meta.color;
  if (meta.xyb_encoded && !CanOutputTo(c)) {
    const bool grey = c.IsGray();
    c = ColorEncoding();
    c.color_space = grey ? kGray : kRGB;
    c.transfer_function = kTFLinear;
    o.linear_srgb_fallback = true;
  }
  JXLO_CHECK(c.IsGray() || (c.primaries == 1 && c.white_point == 1) || !meta.xyb_encoded,
             "non-sRGB primaries / white point are not supported by the oracle");
  for (int i = 0; i < 3; i++) {
    o.opsin_biases[i] = meta.opsin_biases[i];
    o.opsin_biases_cbrt[i] = cbrtf(meta.opsin_biases[i]);
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
  for (int i = 0; i < 9; i++) o.inverse_matrix[i] = inv[i] * (255.0f / meta.intensity_target);
  o.have_gamma = c.have_gamma;
  o.tf = c.transfer_function;
  o.inverse_gamma = c.have_gamma ? static_cast<float>(c.gamma * (1.0 / 10000000)) : (c.transfer_function == kTFDCI ? 1.0f / 2.6f : 1.0f);
  return o;
}

inline void XybToRgb(const OutputColor& o, float x, float y, float b, float* r, float* g, float* bl) {
  float gamma_r = y + x, gamma_g = y - x, gamma_b = b;
  gamma_r = gamma_r - o.opsin_biases_cbrt[0];
  gamma_g = gamma_g - o.opsin_biases_cbrt[1];
  gamma_b = gamma_b - o.opsin_biases_cbrt[2];
  const float r2 = gamma_r * gamma_r, g2 = gamma_g * gamma_g, b2 = gamma_b * gamma_b;
  const float mixed_r = std::fmaf(r2, gamma_r, o.opsin_biases[0]);
  const float mixed_g = std::fmaf(g2, gamma_g, o.opsin_biases[1]);
  const float mixed_b = std::fmaf(b2, gamma_b, o.opsin_biases[2]);
  const float* m = o.inverse_matrix;
  float lr = m[0] * mixed_r, lg = m[3] * mixed_r, lb = m[6] * mixed_r;
  lr = std::fmaf(m[1], mixed_g, lr);
  lg = std::fmaf(m[4], mixed_g, lg);
  lb = std::fmaf(m[7], mixed_g, lb);
  lr = std::fmaf(m[2], mixed_b, lr);
  lg = std::fmaf(m[5], mixed_b, lg);
  lb = std::fmaf(m[8], mixed_b, lb);
  *r = lr;
  *g = lg;
  *bl = lb;
}

inline float SrgbFromLinear(float v) {  // TF_SRGB::EncodedFromDisplay
  static const float p[5] = {-5.135152395e-04f, 5.287254571e-03f, 3.903842876e-01f, 1.474205315e+00f, 7.352629620e-01f};
  static const float q[5] = {1.004519624e-02f, 3.036675394e-01f, 1.340816930e+00f, 9.258482155e-01f, 2.424867759e-02f};
  const float x = std::fabs(v);
  const float linear = x * 12.92f;
  const float s = std::sqrt(x);
  float yp = p[4], yq = q[4];
  for (int i = 3; i >= 0; i--) {
    yp = std::fmaf(yp, s, p[i]);
    yq = std::fmaf(yq, s, q[i]);
  }
  const float poly = yp / yq;
  const float magnitude = x > 0.0031308f ? poly : linear;
  return std::copysign(magnitude, v);
}

inline float FromLinear(const OutputColor& o, float v) {  // stage_from_linear.cc:52-104
  if (o.have_gamma || o.tf == kTFDCI) return v <= 1e-5f ? 0.0f : FastPowf(v, o.inverse_gamma);
  if (o.tf == kTFLinear) return v;
  if (o.tf == kTFSRGB) return SrgbFromLinear(v);
  throw Error("jxlo: transfer function not supported by the oracle (PQ / HLG / 709)");
}

// ---------------------------------------------------------------- upsampling
// lib/jxl/render_pipeline/stage_upsampling.cc:28-170: every input pixel becomes N x N outputs, each a 5 x 5 weighted sum
// of the (mirrored) neighbourhood -- MulAdd in the order iy = -2 .. 2, ix = -2 .. 2 -- clamped to the neighbourhood's
// minimum and maximum. The N / 2 x N / 2 distinct kernels come from the upper triangle `weights` by symmetry
// (constructor, :33-47; Kernel<N>, :93-112).
inline Plane Upsample(const Plane& in, int N, const float* weights) {
  const int half = N / 2;
  std::vector<float> kernel(static_cast<size_t>(4) * 4 * 5 * 5, 0.0f);
  auto K = [&](int a, int b, int c, int d) -> float& { return kernel[((a * 4 + b) * 5 + c) * 5 + d]; };
  for (int i = 0; i < 5 * half; i++)
    for (int j = 0; j < 5 * half; j++) {
      const int y = std::min(i, j), x = std::max(i, j);
      K(j / 5, i / 5, j % 5, i % 5) = weights[5 * half * y - y * (y - 1) / 2 + x - y];
    }
  auto kernel_at = [&](int x, int y, int ix, int iy) {
    ix += 2;
    iy += 2;
    if (N == 2) return K(0, 0, y % 2 ? 4 - iy : iy, x % 2 ? 4 - ix : ix);
    if (N == 4) return K(y % 4 < 2 ? y % 2 : 1 - y % 2, x % 4 < 2 ? x % 2 : 1 - x % 2, y % 4 < 2 ? iy : 4 - iy, x % 4 < 2 ? ix : 4 - ix);
    return K(y % 8 < 4 ? y % 4 : 3 - y % 4, x % 8 < 4 ? x % 4 : 3 - x % 4, y % 8 < 4 ? iy : 4 - iy, x % 8 < 4 ? ix : 4 - ix);
  };
  Plane out(in.w * N, in.h * N);
  const MirroredPlane m{in};
  for (int y = 0; y < in.h; y++)
    for (int x = 0; x < in.w; x++)
      for (int oy = 0; oy < N; oy++)
        for (int ox = 0; ox < N; ox++) {
          float result = 0.0f, mn = in.Row(y)[x], mx = mn;
          for (int iy = -2; iy <= 2; iy++)
            for (int ix = -2; ix <= 2; ix++) {
              const float v = m.At(x + ix, y + iy);
              result = std::fmaf(kernel_at(ox, oy, ix, iy), v, result);
              mn = std::min(v, mn);
              mx = std::max(v, mx);
            }
          out.Row(y * N + oy)[x * N + ox] = std::min(std::max(result, mn), mx);  // hwy Clamp(v, lo, hi) = Min(Max(lo, v), hi)
        }
  return out;
}

inline const float* UpsamplingWeights(const ImageMetadata& meta, uint32_t factor) {
  if (factor == 2) return (meta.custom_weights_mask & 1) ? meta.up2.data() : kDefaultUpsampling2;
  if (factor == 4) return (meta.custom_weights_mask & 2) ? meta.up4.data() : kDefaultUpsampling4;
  return (meta.custom_weights_mask & 4) ? meta.up8.data() : kDefaultUpsampling8;
}

// ---------------------------------------------------------------- the frame
inline void VarDCTToPixels(VarDCTState* vs, std::vector<Plane>* planes) {
  const FrameHeader& fh = vs->fh;
  JXLO_CHECK(fh.Is444(), "chroma-subsampled VarDCT frames are not supported by the oracle");
  for (int c = 0; c < 3; c++) {
    Plane out(vs->dim.xsize, vs->dim.ysize);
    for (int y = 0; y < out.h; y++) std::memcpy(out.Row(y), vs->pix[c].Row(y), sizeof(float) * out.w);
    (*planes)[c] = std::move(out);
  }
}

// After this call `planes` hold either XYB (is_xyb, frame saved before the colour
// transform) or non-linear output samples.
inline void RenderFrame(const FrameHeader& fh, const FrameDimensions& dim, const CodestreamState& cs, VarDCTState* vs,
                        const FeatureState& feat, std::vector<Plane>* planes, bool* is_xyb) {
  for (uint32_t u : fh.ec_upsampling) JXLO_CHECK(u == fh.upsampling, "extra channels upsampled differently from the colour channels are not supported by the oracle");
  JXLO_CHECK(fh.upsampling == 1 || fh.frame_type != kReferenceOnly, "upsampled reference-only frames are not supported by the oracle");
  const LoopFilter& lf = fh.lf;
  if (lf.gab) Gaborish(lf, planes->data());
  if (lf.epf_iters > 0) {
    SigmaLookup sigma{vs, -1.1715728752538099024f / lf.epf_sigma_for_modular};
    if (lf.epf_iters >= 3) EPFStage(0, lf, sigma, planes->data());
    EPFStage(1, lf, sigma, planes->data());
    if (lf.epf_iters >= 2) EPFStage(2, lf, sigma, planes->data());
  }
  if (fh.flags & kFlagPatches) ApplyPatches(feat, cs, planes);
  if ((fh.flags & kFlagSplines) && !feat.splines.segments.empty()) {
    // SplineStage (stage_splines.cc:24-45): every row gets the segments of its list, in list order, over its whole width
    const SplineState& sp = feat.splines;
    Plane& s0 = (*planes)[0];
    Plane& s1 = (*planes)[1];
    Plane& s2 = (*planes)[2];
    for (int y = 0; y < s0.h && static_cast<size_t>(y) + 1 < sp.segment_y_start.size(); y++)
      for (uint32_t i = sp.segment_y_start[y]; i < sp.segment_y_start[y + 1]; i++) {
        const SplineSegment& seg = sp.segments[sp.segment_indices[i]];
        for (int x = 0; x < s0.w; x++) SplDrawPixel(seg, x, static_cast<size_t>(y), 0, s0.w, &s0.Row(y)[x], &s1.Row(y)[x], &s2.Row(y)[x]);
      }
  }
  if (fh.upsampling > 1) {  // (dec_cache.cc:103-347: after patches / splines, before noise and the colour transform)
    for (Plane& p : *planes) p = Upsample(p, static_cast<int>(fh.upsampling), UpsamplingWeights(cs.meta, fh.upsampling));
  }
  *is_xyb = false;
  const bool can_reference = fh.CanBeReferenced() || fh.frame_type == kReferenceOnly;
  if (can_reference && fh.save_before_color_transform) {
    *is_xyb = fh.color_transform == kCTXYB;
    if (fh.frame_type == kReferenceOnly) return;
    JXLO_CHECK(fh.color_transform != kCTXYB, "frames that are both shown and saved before the colour transform are not supported");
  }
  Plane& p0 = (*planes)[0];
  Plane& p1 = (*planes)[1];
  Plane& p2 = (*planes)[2];
  if (fh.color_transform == kCTYCbCr) {
    const float c128 = 128.0f / 255, crcr = 1.402f, cgcb = -0.114f * 1.772f / 0.587f, cgcr = -0.299f * 1.402f / 0.587f,
                cbcb = 1.772f;
    for (int y = 0; y < p0.h; y++)
      for (int x = 0; x < p0.w; x++) {
        const float yv = p1.Row(y)[x] + c128, cb = p0.Row(y)[x], cr = p2.Row(y)[x];
        p0.Row(y)[x] = std::fmaf(crcr, cr, yv);
        p1.Row(y)[x] = std::fmaf(cgcr, cr, std::fmaf(cgcb, cb, yv));
        p2.Row(y)[x] = std::fmaf(cbcb, cb, yv);
      }
  } else if (fh.color_transform == kCTXYB) {
    const OutputColor oc = MakeOutputColor(cs.meta);
    for (int y = 0; y < p0.h; y++)
      for (int x = 0; x < p0.w; x++) {
        float r, g, b;
        XybToRgb(oc, p0.Row(y)[x], p1.Row(y)[x], p2.Row(y)[x], &r, &g, &b);
        p0.Row(y)[x] = FromLinear(oc, r);
        p1.Row(y)[x] = FromLinear(oc, g);
        p2.Row(y)[x] = FromLinear(oc, b);
      }
  }
}

}  // namespace jxlo

#endif  // JXLO_RENDER_H_
