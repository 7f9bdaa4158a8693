// jxl_b200 host parser (product code; runs on the CPU in front of the CUDA kernels).
// Codestream / frame headers and the TOC (host-side parse).
//
// Codestream / frame headers and the TOC. Restates
//   lib/jxl/headers.cc:124-198 (SizeHeader, PreviewHeader, AnimationHeader),
//   lib/jxl/image_metadata.cc:20-60 (BitDepth), :74-200 (CustomTransformData),
//   :203-260 (ExtraChannelInfo), :278-344 (ImageMetadata), :346-398 (opsin, tone mapping),
//   lib/jxl/color_encoding_internal.cc:94-213, lib/jxl/frame_header.cc:25-440,
//   lib/jxl/loop_filter.cc:16-99, lib/jxl/toc.cc:23-66, lib/jxl/coeff_order.cc:36-76,
//   lib/jxl/frame_dimensions.h:34-62.
#ifndef JXLB_HEADERS_H_
#define JXLB_HEADERS_H_

#include <string>
#include <vector>

#include "jxlb_entropy.h"

namespace jxlb {

struct BitDepth {
  bool floating_point = false;
  uint32_t bits = 8;
  uint32_t exp_bits = 0;
};

inline BitDepth ReadBitDepth(BitReader& br) {
  BitDepth d;
  d.floating_point = br.ReadBool();
  if (!d.floating_point) {
    d.bits = ReadU32(br, Val(8), Val(10), Val(12), BitsOffset(6, 1));
    d.exp_bits = 0;
    JXLB_CHECK(d.bits <= 31, "bits_per_sample too large");
  } else {
    d.bits = ReadU32(br, Val(32), Val(16), Val(24), BitsOffset(6, 1));
    d.exp_bits = br.Read(4) + 1;
    JXLB_CHECK(d.exp_bits >= 2 && d.exp_bits <= 8, "bad exponent bits");
    int mant = static_cast<int>(d.bits) - d.exp_bits - 1;
    JXLB_CHECK(mant >= 2 && mant <= 23, "bad float sample bits");
  }
  return d;
}

enum ExtraChannelType { kAlpha = 0, kDepth = 1, kSpotColor = 2, kSelectionMask = 3, kBlack = 4, kCFA = 5, kThermal = 6, kOptional = 16 };

struct ExtraChannelInfo {
  uint32_t type = kAlpha;
  BitDepth bit_depth;
  uint32_t dim_shift = 0;
  std::string name;
  bool alpha_associated = false;
  float spot_color[4] = {0, 0, 0, 0};
  uint32_t cfa_channel = 1;
};

inline std::string ReadName(BitReader& br) {
  uint32_t n = ReadU32(br, Val(0), Bits(4), BitsOffset(5, 16), BitsOffset(10, 48));
  std::string s(n, '\0');
  for (uint32_t i = 0; i < n; i++) s[i] = static_cast<char>(br.Read(8));
  return s;
}

inline ExtraChannelInfo ReadExtraChannelInfo(BitReader& br) {
  ExtraChannelInfo e;
  if (br.ReadBool()) return e;  // all_default: 8-bit alpha
  e.type = ReadEnum(br);
  e.bit_depth = ReadBitDepth(br);
  e.dim_shift = ReadU32(br, Val(0), Val(3), Val(4), BitsOffset(3, 1));
  JXLB_CHECK((1u << e.dim_shift) <= 8, "dim_shift too large");
  e.name = ReadName(br);
  if (e.type == kAlpha) e.alpha_associated = br.ReadBool();
  if (e.type == kSpotColor)
    for (float& c : e.spot_color) c = ReadF16(br);
  if (e.type == kCFA) e.cfa_channel = ReadU32(br, Val(1), Bits(2), BitsOffset(4, 3), BitsOffset(8, 19));
  JXLB_CHECK(e.type <= kThermal || e.type == kOptional, "unknown extra channel type");
  return e;
}

enum ColorSpace { kRGB = 0, kGray = 1, kXYB = 2, kUnknownCS = 3 };
enum TransferFunction { kTF709 = 1, kTFUnknown = 2, kTFLinear = 8, kTFSRGB = 13, kTFPQ = 16, kTFDCI = 17, kTFHLG = 18 };

struct ColorEncoding {
  bool want_icc = false;
  uint32_t color_space = kRGB;
  uint32_t white_point = 1;  // D65
  uint32_t primaries = 1;    // sRGB
  bool have_gamma = false;
  uint32_t gamma = 0;  // exponent * 1e7
  uint32_t transfer_function = kTFSRGB;
  uint32_t rendering_intent = 1;  // relative
  int32_t white_xy[2] = {0, 0};
  int32_t prim_xy[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  bool IsGray() const { return color_space == kGray; }
};

inline void ReadCustomXY(BitReader& br, int32_t xy[2]) {
  for (int i = 0; i < 2; i++) {
    uint32_t u = ReadU32(br, Bits(19), BitsOffset(19, 524288), BitsOffset(20, 1048576), BitsOffset(21, 2097152));
    xy[i] = UnpackSigned(u);
  }
}

inline ColorEncoding ReadColorEncoding(BitReader& br) {
  ColorEncoding c;
  if (br.ReadBool()) return c;  // all_default: sRGB
  c.want_icc = br.ReadBool();
  c.color_space = ReadEnum(br);
  if (c.want_icc) return c;
  if (c.color_space != kXYB) {  // !ImplicitWhitePoint
    c.white_point = ReadEnum(br);
    if (c.white_point == 2) ReadCustomXY(br, c.white_xy);
  } else {
    c.white_point = 1;
  }
  if (c.color_space != kGray && c.color_space != kXYB) {  // HasPrimaries
    c.primaries = ReadEnum(br);
    if (c.primaries == 2)
      for (auto& p : c.prim_xy) ReadCustomXY(br, p);
  }
  if (c.color_space != kXYB) {  // tf not implicit
    c.have_gamma = br.ReadBool();
    if (c.have_gamma) {
      c.gamma = br.Read(24);
      JXLB_CHECK(c.gamma <= 10000000 && static_cast<uint64_t>(c.gamma) * 8192 >= 10000000, "bad gamma");
    } else {
      c.transfer_function = ReadEnum(br);
    }
  } else {
    c.have_gamma = true;
    c.gamma = 3333333;
  }
  c.rendering_intent = ReadEnum(br);
  return c;
}

struct SizeHeader {
  uint32_t xsize = 0, ysize = 0;
};

inline uint32_t AspectRatioX(uint32_t ratio, uint32_t ysize) {
  static const uint32_t num[7] = {1, 12, 4, 3, 16, 5, 2};
  static const uint32_t den[7] = {1, 10, 3, 2, 9, 4, 1};
  return static_cast<uint32_t>(static_cast<uint64_t>(ysize) * num[ratio - 1] / den[ratio - 1]);
}

inline SizeHeader ReadSizeHeader(BitReader& br) {
  SizeHeader s;
  bool small = br.ReadBool();
  if (small) {
    s.ysize = (br.Read(5) + 1) * 8;
  } else {
    s.ysize = ReadU32(br, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  }
  uint32_t ratio = br.Read(3);
  if (ratio == 0) {
    if (small) {
      s.xsize = (br.Read(5) + 1) * 8;
    } else {
      s.xsize = ReadU32(br, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
    }
  } else {
    s.xsize = AspectRatioX(ratio, s.ysize);
  }
  return s;
}

inline SizeHeader ReadPreviewHeader(BitReader& br) {
  SizeHeader s;
  bool div8 = br.ReadBool();
  if (div8) {
    s.ysize = 8 * ReadU32(br, Val(16), Val(32), BitsOffset(5, 1), BitsOffset(9, 33));
  } else {
    s.ysize = ReadU32(br, BitsOffset(6, 1), BitsOffset(8, 65), BitsOffset(10, 321), BitsOffset(12, 1345));
  }
  uint32_t ratio = br.Read(3);
  if (ratio == 0) {
    if (div8) {
      s.xsize = 8 * ReadU32(br, Val(16), Val(32), BitsOffset(5, 1), BitsOffset(9, 33));
    } else {
      s.xsize = ReadU32(br, BitsOffset(6, 1), BitsOffset(8, 65), BitsOffset(10, 321), BitsOffset(12, 1345));
    }
  } else {
    s.xsize = AspectRatioX(ratio, s.ysize);
  }
  return s;
}

// lib/jxl/cms/opsin_params.h:34-53 and lib/jxl/quantizer.h:52-57
static const float kDefaultInverseOpsin[9] = {
    11.031566901960783f, -9.866943921568629f, -0.16462299647058826f,
    -3.254147380392157f, 4.418770392156863f,  -0.16462299647058826f,
    -3.6588512862745097f, 2.7129230470588235f, 1.9459282392156863f};
static const float kNegOpsinBias[3] = {-0.0037930732552754493f, -0.0037930732552754493f, -0.0037930732552754493f};
static const float kDefaultQuantBias[4] = {1.0f - 0.05465007330715401f, 1.0f - 0.07005449891748593f,
                                          1.0f - 0.049935103337343655f, 0.145f};

struct ImageMetadata {
  uint32_t orientation = 1;
  bool have_intrinsic_size = false, have_preview = false, have_animation = false;
  SizeHeader intrinsic_size, preview_size;
  uint32_t tps_num = 100, tps_den = 1, num_loops = 0;
  bool have_timecodes = false;
  BitDepth bit_depth;
  bool modular_16_bit_buffer_sufficient = true;
  std::vector<ExtraChannelInfo> extra;
  bool xyb_encoded = true;
  ColorEncoding color;
  float intensity_target = 255.0f, min_nits = 0.0f, linear_below = 0.0f;
  bool relative_to_max_display = false;
  // CustomTransformData
  float inverse_opsin[9];
  float opsin_biases[3];
  float quant_biases[4];
  uint32_t custom_weights_mask = 0;
  std::vector<float> up2, up4, up8;
  ImageMetadata() {
    for (int i = 0; i < 9; i++) inverse_opsin[i] = kDefaultInverseOpsin[i];
    for (int i = 0; i < 3; i++) opsin_biases[i] = kNegOpsinBias[i];
    for (int i = 0; i < 4; i++) quant_biases[i] = kDefaultQuantBias[i];
  }
  int AlphaIndex() const {
    for (size_t i = 0; i < extra.size(); i++)
      if (extra[i].type == kAlpha) return static_cast<int>(i);
    return -1;
  }
};

inline void ReadImageMetadata(BitReader& br, ImageMetadata* m) {
  if (br.ReadBool()) return;  // all_default
  bool extra_fields = br.ReadBool();
  if (extra_fields) {
    m->orientation = br.Read(3) + 1;
    m->have_intrinsic_size = br.ReadBool();
    if (m->have_intrinsic_size) m->intrinsic_size = ReadSizeHeader(br);
    m->have_preview = br.ReadBool();
    if (m->have_preview) m->preview_size = ReadPreviewHeader(br);
    m->have_animation = br.ReadBool();
    if (m->have_animation) {
      m->tps_num = ReadU32(br, Val(100), Val(1000), BitsOffset(10, 1), BitsOffset(30, 1));
      m->tps_den = ReadU32(br, Val(1), Val(1001), BitsOffset(8, 1), BitsOffset(10, 1));
      m->num_loops = ReadU32(br, Val(0), Bits(3), Bits(16), Bits(32));
      m->have_timecodes = br.ReadBool();
    }
  }
  m->bit_depth = ReadBitDepth(br);
  m->modular_16_bit_buffer_sufficient = br.ReadBool();
  uint32_t nec = ReadU32(br, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(12, 1));
  m->extra.resize(nec);
  for (auto& e : m->extra) e = ReadExtraChannelInfo(br);
  m->xyb_encoded = br.ReadBool();
  m->color = ReadColorEncoding(br);
  if (extra_fields) {
    if (!br.ReadBool()) {  // tone mapping not all_default
      m->intensity_target = ReadF16(br);
      JXLB_CHECK(m->intensity_target > 0, "bad intensity target");
      m->min_nits = ReadF16(br);
      m->relative_to_max_display = br.ReadBool();
      m->linear_below = ReadF16(br);
    }
  }
  SkipExtensions(br);
}

inline void ReadCustomTransformData(BitReader& br, ImageMetadata* m) {
  if (br.ReadBool()) return;  // all_default
  if (m->xyb_encoded) {
    if (!br.ReadBool()) {  // OpsinInverseMatrix not all_default
      for (int i = 0; i < 9; i++) m->inverse_opsin[i] = ReadF16(br);
      for (int i = 0; i < 3; i++) m->opsin_biases[i] = ReadF16(br);
      for (int i = 0; i < 4; i++) m->quant_biases[i] = ReadF16(br);
    }
  }
  m->custom_weights_mask = br.Read(3);
  if (m->custom_weights_mask & 1) {
    m->up2.resize(15);
    for (float& w : m->up2) w = ReadF16(br);
  }
  if (m->custom_weights_mask & 2) {
    m->up4.resize(55);
    for (float& w : m->up4) w = ReadF16(br);
  }
  if (m->custom_weights_mask & 4) {
    m->up8.resize(210);
    for (float& w : m->up8) w = ReadF16(br);
  }
}

// ---------------------------------------------------------------- frame header
enum FrameType { kRegularFrame = 0, kDCFrame = 1, kReferenceOnly = 2, kSkipProgressive = 3 };
enum FrameFlags : uint64_t {
  kFlagNoise = 1, kFlagPatches = 2, kFlagSplines = 16, kFlagUseDcFrame = 32, kFlagSkipAdaptiveDCSmoothing = 128
};
enum ColorTransform { kCTXYB = 0, kCTNone = 1, kCTYCbCr = 2 };
enum BlendMode { kReplace = 0, kAdd = 1, kBlend = 2, kAlphaWeightedAdd = 3, kMul = 4 };

struct BlendingInfo {
  uint32_t mode = kReplace, alpha_channel = 0, source = 0;
  bool clamp = false;
};

struct LoopFilter {
  bool gab = true, gab_custom = false;
  float gab_x_weight1 = 1.1f * 0.104699568f, gab_x_weight2 = 1.1f * 0.055680538f;
  float gab_y_weight1 = 1.1f * 0.104699568f, gab_y_weight2 = 1.1f * 0.055680538f;
  float gab_b_weight1 = 1.1f * 0.104699568f, gab_b_weight2 = 1.1f * 0.055680538f;
  uint32_t epf_iters = 2;
  bool epf_sharp_custom = false, epf_weight_custom = false, epf_sigma_custom = false;
  float epf_sharp_lut[8];
  float epf_channel_scale[3] = {40.0f, 5.0f, 3.5f};
  float epf_pass1_zeroflush = 0.45f, epf_pass2_zeroflush = 0.6f;
  float epf_quant_mul = 0.46f, epf_pass0_sigma_scale = 0.9f, epf_pass2_sigma_scale = 6.5f;
  float epf_border_sad_mul = 2.0f / 3.0f;
  float epf_sigma_for_modular = 1.0f;
  LoopFilter() {
    for (int i = 0; i < 8; i++) epf_sharp_lut[i] = static_cast<float>(i) / 7.0f;
  }
};

// The F16 defaults of the gaborish weights are stored as the float product
// computed in double (lib/jxl/loop_filter.cc:28-53: `1.1 * 0.104699568f`).
inline float GabDefault1() { return static_cast<float>(1.1 * 0.104699568f); }
inline float GabDefault2() { return static_cast<float>(1.1 * 0.055680538f); }

inline LoopFilter DefaultLoopFilter() {
  LoopFilter lf;
  lf.gab_x_weight1 = lf.gab_y_weight1 = lf.gab_b_weight1 = GabDefault1();
  lf.gab_x_weight2 = lf.gab_y_weight2 = lf.gab_b_weight2 = GabDefault2();
  return lf;
}

inline LoopFilter ReadLoopFilter(BitReader& br, bool is_modular) {
  LoopFilter lf = DefaultLoopFilter();
  if (br.ReadBool()) return lf;  // all_default
  lf.gab = br.ReadBool();
  if (lf.gab) {
    lf.gab_custom = br.ReadBool();
    if (lf.gab_custom) {
      lf.gab_x_weight1 = ReadF16(br);
      lf.gab_x_weight2 = ReadF16(br);
      lf.gab_y_weight1 = ReadF16(br);
      lf.gab_y_weight2 = ReadF16(br);
      lf.gab_b_weight1 = ReadF16(br);
      lf.gab_b_weight2 = ReadF16(br);
    }
  }
  lf.epf_iters = br.Read(2);
  if (lf.epf_iters > 0) {
    if (!is_modular) {
      lf.epf_sharp_custom = br.ReadBool();
      if (lf.epf_sharp_custom)
        for (float& v : lf.epf_sharp_lut) v = ReadF16(br);
    }
    lf.epf_weight_custom = br.ReadBool();
    if (lf.epf_weight_custom) {
      for (float& v : lf.epf_channel_scale) v = ReadF16(br);
      lf.epf_pass1_zeroflush = ReadF16(br);
      lf.epf_pass2_zeroflush = ReadF16(br);
    }
    lf.epf_sigma_custom = br.ReadBool();
    if (lf.epf_sigma_custom) {
      if (!is_modular) lf.epf_quant_mul = ReadF16(br);
      lf.epf_pass0_sigma_scale = ReadF16(br);
      lf.epf_pass2_sigma_scale = ReadF16(br);
      lf.epf_border_sad_mul = ReadF16(br);
    }
    if (is_modular) {
      lf.epf_sigma_for_modular = ReadF16(br);
      JXLB_CHECK(lf.epf_sigma_for_modular >= 1e-8f, "EPF sigma too small");
    }
  }
  SkipExtensions(br);
  return lf;
}

struct Passes {
  uint32_t num_passes = 1, num_downsample = 0;
  uint32_t shift[11] = {0}, downsample[4] = {0}, last_pass[4] = {0};
  // lib/jxl/frame_header.h:268-284
  void DownsamplingBracket(size_t pass, int* min_shift, int* max_shift) const {
    *max_shift = 2;
    *min_shift = 3;
    for (size_t i = 0;; i++) {
      for (uint32_t j = 0; j < num_downsample; j++) {
        if (i == last_pass[j]) {
          if (downsample[j] == 8) *min_shift = 3;
          if (downsample[j] == 4) *min_shift = 2;
          if (downsample[j] == 2) *min_shift = 1;
          if (downsample[j] == 1) *min_shift = 0;
        }
      }
      if (i == num_passes - 1) *min_shift = 0;
      if (i == pass) return;
      *max_shift = *min_shift - 1;
    }
  }
};

struct FrameHeader {
  uint32_t frame_type = kRegularFrame;
  bool is_modular = false;
  uint64_t flags = 0;
  uint32_t color_transform = kCTXYB;
  uint32_t chroma_mode[3] = {0, 0, 0};
  uint32_t upsampling = 1;
  std::vector<uint32_t> ec_upsampling;
  uint32_t group_size_shift = 1;
  uint32_t x_qm_scale = 3, b_qm_scale = 2;
  Passes passes;
  uint32_t dc_level = 0;
  bool custom_size_or_origin = false;
  int32_t x0 = 0, y0 = 0;
  uint32_t xsize = 0, ysize = 0;  // of this frame (after resolving defaults)
  BlendingInfo blending;
  std::vector<BlendingInfo> ec_blending;
  uint32_t duration = 0, timecode = 0;
  bool is_last = true;
  uint32_t save_as_reference = 0;
  bool save_before_color_transform = false;
  std::string name;
  LoopFilter lf;

  static const uint8_t kHShift[4];
  static const uint8_t kVShift[4];
  uint32_t MaxHShift() const { return std::max({kHShift[chroma_mode[0]], kHShift[chroma_mode[1]], kHShift[chroma_mode[2]]}); }
  uint32_t MaxVShift() const { return std::max({kVShift[chroma_mode[0]], kVShift[chroma_mode[1]], kVShift[chroma_mode[2]]}); }
  uint32_t HShift(int c) const { return MaxHShift() - kHShift[chroma_mode[c]]; }
  uint32_t VShift(int c) const { return MaxVShift() - kVShift[chroma_mode[c]]; }
  bool Is444() const { return MaxHShift() == 0 && MaxVShift() == 0; }
  bool CanBeReferenced() const {
    return !is_last && frame_type != kDCFrame && (duration == 0 || save_as_reference != 0);
  }
};
inline const uint8_t FrameHeader::kHShift[4] = {0, 1, 1, 0};
inline const uint8_t FrameHeader::kVShift[4] = {0, 1, 0, 1};

inline BlendingInfo ReadBlendingInfo(BitReader& br, size_t num_ec, bool partial) {
  BlendingInfo b;
  b.mode = ReadU32(br, Val(kReplace), Val(kAdd), Val(kBlend), BitsOffset(2, 3));
  JXLB_CHECK(b.mode <= kMul, "bad blend mode");
  bool uses_alpha = num_ec > 0 && (b.mode == kBlend || b.mode == kAlphaWeightedAdd);
  if (uses_alpha) {
    b.alpha_channel = ReadU32(br, Val(0), Val(1), Val(2), BitsOffset(3, 3));
    JXLB_CHECK(b.alpha_channel < num_ec, "bad blend alpha channel");
  }
  if (uses_alpha || b.mode == kMul) b.clamp = br.ReadBool();
  if (b.mode != kReplace || partial) b.source = ReadU32(br, Val(0), Val(1), Val(2), Val(3));
  return b;
}

inline void ReadFrameHeader(BitReader& br, const SizeHeader& size, const ImageMetadata& m,
                            bool is_preview, FrameHeader* f) {
  const size_t num_ec = m.extra.size();
  const uint32_t def_x = is_preview ? m.preview_size.xsize : size.xsize;
  const uint32_t def_y = is_preview ? m.preview_size.ysize : size.ysize;
  f->xsize = def_x;
  f->ysize = def_y;
  f->ec_upsampling.assign(num_ec, 1);
  f->ec_blending.assign(num_ec, BlendingInfo());
  f->color_transform = m.xyb_encoded ? kCTXYB : kCTNone;
  f->lf = DefaultLoopFilter();
  if (br.ReadBool()) {  // all_default
    f->x_qm_scale = 3;
    f->b_qm_scale = 2;
    return;
  }
  f->frame_type = br.Read(2);
  f->is_modular = br.ReadBool();
  f->flags = ReadU64(br);
  if (!m.xyb_encoded) f->color_transform = br.ReadBool() ? kCTYCbCr : kCTNone;
  if (f->color_transform == kCTYCbCr && !(f->flags & kFlagUseDcFrame)) {
    for (auto& c : f->chroma_mode) c = br.Read(2);
  }
  if (!(f->flags & kFlagUseDcFrame)) {
    f->upsampling = ReadU32(br, Val(1), Val(2), Val(4), Val(8));
    for (size_t i = 0; i < num_ec; i++) {
      uint32_t u = ReadU32(br, Val(1), Val(2), Val(4), Val(8));
      u <<= m.extra[i].dim_shift;
      JXLB_CHECK(u >= f->upsampling && u <= 8, "bad extra-channel upsampling");
      f->ec_upsampling[i] = u;
    }
  }
  if (f->is_modular) f->group_size_shift = br.Read(2);
  if (!f->is_modular && f->color_transform == kCTXYB) {
    f->x_qm_scale = br.Read(3);
    f->b_qm_scale = br.Read(3);
  } else {
    f->x_qm_scale = f->b_qm_scale = 2;
  }
  if (f->frame_type != kReferenceOnly) {
    Passes& p = f->passes;
    p.num_passes = ReadU32(br, Val(1), Val(2), Val(3), BitsOffset(3, 4));
    if (p.num_passes != 1) {
      p.num_downsample = ReadU32(br, Val(0), Val(1), Val(2), BitsOffset(1, 3));
      JXLB_CHECK(p.num_downsample <= 4 && p.num_downsample <= p.num_passes, "bad num_downsample");
      for (uint32_t i = 0; i + 1 < p.num_passes; i++) p.shift[i] = br.Read(2);
      p.shift[p.num_passes - 1] = 0;
      for (uint32_t i = 0; i < p.num_downsample; i++) p.downsample[i] = ReadU32(br, Val(1), Val(2), Val(4), Val(8));
      for (uint32_t i = 0; i < p.num_downsample; i++) p.last_pass[i] = ReadU32(br, Val(0), Val(1), Val(2), Bits(3));
    }
  }
  if (f->frame_type == kDCFrame) f->dc_level = ReadU32(br, Val(1), Val(2), Val(3), Val(4));
  bool partial = false;
  if (f->frame_type != kDCFrame) {
    f->custom_size_or_origin = br.ReadBool();
    if (f->custom_size_or_origin) {
      auto rd = [&]() { return ReadU32(br, Bits(8), BitsOffset(11, 256), BitsOffset(14, 2304), BitsOffset(30, 18688)); };
      if (f->frame_type == kRegularFrame || f->frame_type == kSkipProgressive) {
        f->x0 = UnpackSigned(rd());
        f->y0 = UnpackSigned(rd());
      }
      f->xsize = rd();
      f->ysize = rd();
      JXLB_CHECK(f->xsize != 0 && f->ysize != 0, "empty frame");
      if (f->frame_type == kRegularFrame || f->frame_type == kSkipProgressive) {
        partial |= f->x0 > 0 || f->y0 > 0;
        partial |= static_cast<int64_t>(f->xsize) + f->x0 < static_cast<int64_t>(def_x);
        partial |= static_cast<int64_t>(f->ysize) + f->y0 < static_cast<int64_t>(def_y);
      }
    }
  }
  if (f->frame_type == kRegularFrame || f->frame_type == kSkipProgressive) {
    f->blending = ReadBlendingInfo(br, num_ec, partial);
    for (auto& b : f->ec_blending) b = ReadBlendingInfo(br, num_ec, partial);
    if (m.have_animation) {
      f->duration = ReadU32(br, Val(0), Val(1), Bits(8), Bits(32));
      if (m.have_timecodes) f->timecode = br.Read(32);
    }
    f->is_last = br.ReadBool();
  } else {
    f->is_last = false;
  }
  if (f->frame_type != kDCFrame && !f->is_last) f->save_as_reference = ReadU32(br, Val(0), Val(1), Val(2), Val(3));
  if (f->frame_type != kDCFrame) {
    if (f->CanBeReferenced() && f->blending.mode == kReplace && !partial &&
        (f->frame_type == kRegularFrame || f->frame_type == kSkipProgressive)) {
      f->save_before_color_transform = br.ReadBool();
    } else if (f->frame_type == kReferenceOnly) {
      f->save_before_color_transform = br.ReadBool();
    }
  } else {
    f->save_before_color_transform = true;
  }
  f->name = ReadName(br);
  f->lf = ReadLoopFilter(br, f->is_modular);
  SkipExtensions(br);
}

// lib/jxl/frame_dimensions.h:34-62
struct FrameDimensions {
  size_t xsize, ysize, xsize_upsampled, ysize_upsampled;
  size_t xsize_padded, ysize_padded, xsize_blocks, ysize_blocks;
  size_t xsize_groups, ysize_groups, xsize_dc_groups, ysize_dc_groups;
  size_t num_groups, num_dc_groups, group_dim, dc_group_dim;
  void Set(size_t xs, size_t ys, size_t group_size_shift, size_t max_hshift, size_t max_vshift,
           bool modular_mode, size_t upsampling) {
    group_dim = size_t{128} << group_size_shift;
    dc_group_dim = group_dim * 8;
    xsize_upsampled = xs;
    ysize_upsampled = ys;
    xsize = DivCeil(xs, upsampling);
    ysize = DivCeil(ys, upsampling);
    xsize_blocks = DivCeil(xsize, size_t{8} << max_hshift) << max_hshift;
    ysize_blocks = DivCeil(ysize, size_t{8} << max_vshift) << max_vshift;
    xsize_padded = xsize_blocks * 8;
    ysize_padded = ysize_blocks * 8;
    if (modular_mode) {
      xsize_padded = xsize;
      ysize_padded = ysize;
    }
    xsize_groups = DivCeil(xsize, group_dim);
    ysize_groups = DivCeil(ysize, group_dim);
    xsize_dc_groups = DivCeil(xsize_blocks, group_dim);
    ysize_dc_groups = DivCeil(ysize_blocks, group_dim);
    num_groups = xsize_groups * ysize_groups;
    num_dc_groups = xsize_dc_groups * ysize_dc_groups;
  }
};

inline FrameDimensions ToFrameDimensions(const FrameHeader& f) {
  FrameDimensions d;
  d.Set(f.xsize, f.ysize, f.group_size_shift, f.MaxHShift(), f.MaxVShift(), f.is_modular, f.upsampling);
  return d;
}

// ---------------------------------------------------------------- permutation + TOC
// lib/jxl/coeff_order.cc:36-76 + lib/jxl/lehmer_code.h (decode)
inline uint32_t CoeffOrderContext(uint32_t v) {
  if (v == 0) return 0;
  return std::min<uint32_t>(FloorLog2(v) + 1, 7);
}

inline void ReadPermutation(BitReader& br, SymbolReader& reader, size_t skip, size_t size, uint32_t* order) {
  std::vector<uint32_t> lehmer(size, 0);
  uint32_t end = reader.ReadUint(CoeffOrderContext(size), br) + skip;
  JXLB_CHECK(end <= size, "bad permutation size");
  uint32_t last = 0;
  for (size_t i = skip; i < end; i++) {
    lehmer[i] = reader.ReadUint(CoeffOrderContext(last), br);
    last = lehmer[i];
    JXLB_CHECK(lehmer[i] < size - i, "bad lehmer code");
  }
  // order[i] = the lehmer[i]-th not-yet-used element
  std::vector<uint32_t> avail(size);
  for (size_t i = 0; i < size; i++) avail[i] = i;
  for (size_t i = 0; i < size; i++) {
    order[i] = avail[lehmer[i]];
    avail.erase(avail.begin() + lehmer[i]);
  }
}

inline void ReadPermutationStream(BitReader& br, size_t skip, size_t size, uint32_t* order) {
  EntropyCode code;
  ReadEntropyCode(br, 8, &code);
  SymbolReader reader(&code, br);
  ReadPermutation(br, reader, skip, size, order);
  JXLB_CHECK(reader.FinalStateOk(), "permutation: bad ANS final state");
}

struct Toc {
  std::vector<uint32_t> sizes;     // in bitstream order
  std::vector<size_t> offsets;     // offsets[i] = byte offset (from the end of the TOC) of logical section i
  std::vector<uint32_t> logical_size;
  size_t total = 0;
};

inline size_t NumTocEntries(size_t num_groups, size_t num_dc_groups, size_t num_passes) {
  if (num_groups == 1 && num_passes == 1) return 1;
  return 2 + num_dc_groups + num_passes * num_groups;
}

inline Toc ReadToc(BitReader& br, size_t entries) {
  JXLB_CHECK(entries <= 65536, "too many TOC entries");
  Toc t;
  std::vector<uint32_t> perm;
  if (br.ReadBool()) {
    perm.resize(entries);
    ReadPermutationStream(br, 0, entries, perm.data());
  }
  br.AlignToByte();
  t.sizes.resize(entries);
  for (auto& s : t.sizes) s = ReadU32(br, Bits(10), BitsOffset(14, 1024), BitsOffset(22, 17408), BitsOffset(30, 4211712));
  br.AlignToByte();
  br.CheckInBounds();
  // lib/jxl/toc.cc:70-105: logical section j lives at bitstream slot perm[j].
  std::vector<size_t> pre(entries, 0);
  size_t off = 0;
  for (size_t i = 0; i < entries; i++) {
    pre[i] = off;
    off += t.sizes[i];
  }
  t.offsets.assign(entries, 0);
  t.logical_size.assign(entries, 0);
  for (size_t j = 0; j < entries; j++) {
    size_t slot = perm.empty() ? j : perm[j];
    t.offsets[j] = pre[slot];
    t.logical_size[j] = t.sizes[slot];
  }
  t.total = off;
  return t;
}

}  // namespace jxlb

#endif  // JXLB_HEADERS_H_
