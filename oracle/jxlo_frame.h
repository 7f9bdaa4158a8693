// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
//
// Frame decoding: section walk, global / per-group Modular streams, hand-off
// to the VarDCT path and the render stages. Restates
//   lib/jxl/dec_frame.cc:133-212 (InitFrame), :266-339 (DC global, DC group),
//   :367-555 (AC global, AC group), :568-731 (ProcessSections),
//   lib/jxl/dec_modular.cc:179-288 (DecodeGlobalInfo), :301-395 (DecodeGroup),
//   :534-708 (ModularImageToDecodedRect), :710-761 (FinalizeDecoding).
#ifndef JXLO_FRAME_H_
#define JXLO_FRAME_H_

#include <memory>

#include "jxlo_headers.h"
#include "jxlo_modular.h"

namespace jxlo {

struct Plane {
  int w = 0, h = 0;
  std::vector<float> d;
  Plane() = default;
  Plane(int w_, int h_) : w(w_), h(h_), d(static_cast<size_t>(w_) * h_, 0.0f) {}
  float* Row(int y) { return d.data() + static_cast<size_t>(y) * w; }
  const float* Row(int y) const { return d.data() + static_cast<size_t>(y) * w; }
};

// A decoded frame before blending: 3 colour planes + one per extra channel, at
// the frame's upsampled size.
struct FrameBuffer {
  std::vector<Plane> planes;
  bool is_xyb = false;  // planes hold XYB (saved before colour transform)
};

struct VarDCTState;  // jxlo_vardct.h

struct CodestreamState {
  SizeHeader size;
  ImageMetadata meta;
  FrameBuffer reference[4];
  bool reference_valid[4] = {false, false, false, false};
  std::vector<Plane> dc_frames[4];
};

// Stream numbering, lib/jxl/dec_modular.h:44-68.
constexpr uint32_t kNumQuantTables = 17;
inline uint32_t StreamGlobal() { return 0; }
inline uint32_t StreamVarDCTDC(const FrameDimensions&, size_t g) { return 1 + g; }
inline uint32_t StreamModularDC(const FrameDimensions& d, size_t g) { return 1 + d.num_dc_groups + g; }
inline uint32_t StreamACMetadata(const FrameDimensions& d, size_t g) { return 1 + 2 * d.num_dc_groups + g; }
inline uint32_t StreamQuantTable(const FrameDimensions& d, size_t i) { return 1 + 3 * d.num_dc_groups + i; }
inline uint32_t StreamModularAC(const FrameDimensions& d, size_t g, size_t pass) {
  return 1 + 3 * d.num_dc_groups + kNumQuantTables + d.num_groups * pass + g;
}

struct ModularFrameState {
  Tree tree;
  EntropyCode code;
  bool has_tree = false;
  ModImage full;
  GroupHeader global_header;
  bool do_color = false;
  size_t nb_color = 0;
};

// lib/jxl/dec_modular.cc:179-288
inline void ModularDecodeGlobal(BitReader& br, const FrameHeader& fh, const FrameDimensions& dim,
                                const ImageMetadata& meta, ModularFrameState* ms) {
  bool is_gray = meta.color.IsGray();
  size_t nb_chans = (is_gray && fh.color_transform == kCTNone) ? 1 : 3;
  ms->do_color = fh.is_modular;
  size_t nb_extra = meta.extra.size();
  ms->has_tree = br.ReadBool();
  if (ms->has_tree) {
    size_t limit = std::min<size_t>(size_t{1} << 22, 1024 + dim.xsize * dim.ysize * (nb_chans + nb_extra) / 16);
    ReadTree(br, &ms->tree, limit);
    ReadEntropyCode(br, (ms->tree.size() + 1) / 2, &ms->code);
  }
  if (!ms->do_color) nb_chans = 0;
  ms->nb_color = nb_chans;
  if (meta.bit_depth.bits >= 32 && ms->do_color && fh.color_transform != kCTXYB) {
    JXLO_CHECK(meta.bit_depth.bits == 32 && meta.bit_depth.floating_point, "unsupported 32-bit integer samples");
  }
  ModImage& gi = ms->full;
  gi.w = dim.xsize;
  gi.h = dim.ysize;
  gi.bitdepth = meta.bit_depth.bits;
  gi.nb_meta = 0;
  gi.ch.clear();
  for (size_t c = 0; c < nb_chans + nb_extra; c++) gi.ch.emplace_back(dim.xsize, dim.ysize);
  if (fh.color_transform == kCTYCbCr) {
    for (size_t c = 0; c < nb_chans; c++) {
      gi.ch[c].hshift = fh.HShift(c);
      gi.ch[c].vshift = fh.VShift(c);
      gi.ch[c].Resize(DivCeil(dim.xsize, size_t{1} << gi.ch[c].hshift), DivCeil(dim.ysize, size_t{1} << gi.ch[c].vshift));
    }
  }
  for (size_t ec = 0, c = nb_chans; ec < nb_extra; ec++, c++) {
    size_t ecups = fh.ec_upsampling[ec];
    gi.ch[c].Resize(DivCeil(dim.xsize_upsampled, ecups), DivCeil(dim.ysize_upsampled, ecups));
    gi.ch[c].hshift = gi.ch[c].vshift = static_cast<int>(CeilLog2(ecups)) - static_cast<int>(CeilLog2(fh.upsampling));
  }
  ModularOptions opt;
  opt.max_chan_size = dim.group_dim;
  opt.group_dim = dim.group_dim;
  ms->global_header = ModularDecode(br, gi, StreamGlobal(), opt, ms->has_tree ? &ms->tree : nullptr,
                                    ms->has_tree ? &ms->code : nullptr);
}

// lib/jxl/dec_modular.cc:301-395 (always the full-image path; the per-group
// shortcut there produces the same samples).
inline void ModularDecodeGroup(BitReader& br, const FrameDimensions& dim, ModularFrameState* ms, size_t x0, size_t y0,
                               size_t xs, size_t ys, int min_shift, int max_shift, uint32_t stream_id) {
  ModImage& full = ms->full;
  ModImage gi;
  gi.w = xs;
  gi.h = ys;
  gi.bitdepth = full.bitdepth;
  size_t c = full.nb_meta;
  for (; c < full.ch.size(); c++) {
    const Channel& fc = full.ch[c];
    if (static_cast<size_t>(fc.w) > dim.group_dim || static_cast<size_t>(fc.h) > dim.group_dim) break;
  }
  size_t beginc = c;
  struct Dest { size_t c; int x, y, w, h; };
  std::vector<Dest> dests;
  for (; c < full.ch.size(); c++) {
    const Channel& fc = full.ch[c];
    int shift = std::min(fc.hshift, fc.vshift);
    if (shift > max_shift || shift < min_shift) continue;
    int rx = x0 >> fc.hshift, ry = y0 >> fc.vshift;
    int rw = xs >> fc.hshift, rh = ys >> fc.vshift;
    if (rx >= fc.w || ry >= fc.h) continue;
    rw = std::min(rw, fc.w - rx);
    rh = std::min(rh, fc.h - ry);
    if (rw <= 0 || rh <= 0) continue;
    gi.ch.emplace_back(rw, rh, fc.hshift, fc.vshift);
    dests.push_back(Dest{c, rx, ry, rw, rh});
  }
  (void)beginc;
  if (gi.ch.empty()) return;
  ModularOptions opt;
  GroupHeader hdr = ModularDecode(br, gi, stream_id, opt, ms->has_tree ? &ms->tree : nullptr,
                                  ms->has_tree ? &ms->code : nullptr);
  UndoTransforms(gi, hdr.wp);
  JXLO_CHECK(gi.ch.size() == dests.size(), "modular group: channel count changed");
  for (size_t i = 0; i < dests.size(); i++) {
    const Dest& d = dests[i];
    JXLO_CHECK(gi.ch[i].w == d.w && gi.ch[i].h == d.h, "modular group: channel size changed");
    Channel& fc = full.ch[d.c];
    for (int y = 0; y < d.h; y++) std::memcpy(fc.Row(d.y + y) + d.x, gi.ch[i].Row(y), sizeof(int32_t) * d.w);
  }
}

// int -> float for custom float samples (lib/jxl/dec_modular.cc:104-160)
inline float IntToFloat(int32_t in, int bits, int exp_bits) {
  if (bits == 32) {
    float f;
    std::memcpy(&f, &in, 4);
    return f;
  }
  int exp_bias = (1 << (exp_bits - 1)) - 1;
  uint32_t f = static_cast<uint32_t>(in);
  int signbit = (f >> (bits - 1)) & 1;
  f &= (1u << (bits - 1)) - 1;
  if (f == 0) return signbit ? -0.0f : 0.0f;
  int exp = static_cast<int>(f >> (bits - exp_bits - 1));
  int mant_bits = bits - exp_bits - 1;
  int mant_shift = 23 - mant_bits;
  int mantissa = static_cast<int>((f & ((1u << mant_bits) - 1)) << mant_shift);
  if (exp == 0 && exp_bits < 8) {  // subnormal in the source format
    exp = 1;
    while ((mantissa & 0x800000) == 0) {
      mantissa <<= 1;
      exp--;
    }
    mantissa &= 0x7fffff;
  }
  exp -= exp_bias;
  exp += 127;
  JXLO_CHECK(exp >= 0, "float sample underflow");
  uint32_t out = (signbit ? 0x80000000u : 0) | (static_cast<uint32_t>(exp) << 23) | static_cast<uint32_t>(mantissa);
  float r;
  std::memcpy(&r, &out, 4);
  return r;
}

// lib/jxl/dec_modular.cc:534-708, applied to the whole frame at once.
// `planes` must hold 3 + num_extra planes; colour planes are only written when
// the frame is Modular-encoded. dc_quant = DequantMatrices::DCQuants().
inline void ModularToFloat(const FrameHeader& fh, const ImageMetadata& meta, ModularFrameState* ms,
                           const float dc_quant[3], std::vector<Plane>* planes) {
  ModImage& gi = ms->full;
  UndoTransforms(gi, ms->global_header.wp);
  size_t c = 0;
  auto convert = [&](const Channel& in, Plane* out, double factor, bool fp, int bits, int exp_bits) {
    *out = Plane(in.w, in.h);
    for (int y = 0; y < in.h; y++) {
      const int32_t* ri = in.Row(y);
      float* ro = out->Row(y);
      for (int x = 0; x < in.w; x++) {
        if (fp) {
          ro[x] = IntToFloat(ri[x], bits, exp_bits);
        } else if (gi.bitdepth < 23) {
          ro[x] = static_cast<float>(ri[x]) * static_cast<float>(factor);
        } else {
          ro[x] = static_cast<float>(ri[x] * factor);
        }
      }
    }
  };
  if (ms->do_color) {
    const bool rgb_from_gray = meta.color.IsGray() && fh.color_transform == kCTNone;
    const bool fp = meta.bit_depth.floating_point && fh.color_transform != kCTXYB;
    for (; c < 3; c++) {
      double factor = gi.bitdepth < 32 ? 1.0 / ((1u << gi.bitdepth) - 1) : 0;
      size_t c_in = c;
      if (fh.color_transform == kCTXYB) {
        factor = dc_quant[c];
        if (c < 2) c_in = 1 - c;
      } else if (rgb_from_gray) {
        c_in = 0;
      }
      JXLO_CHECK(c_in < gi.ch.size(), "modular: missing colour channel");
      const Channel& in = gi.ch[c_in];
      JXLO_CHECK(in.w && in.h, "modular: empty colour channel");
      if (fh.color_transform == kCTXYB && c == 2) {
        Plane& out = (*planes)[2];
        out = Plane(in.w, in.h);
        const Channel& iy = gi.ch[0];
        const float f = static_cast<float>(factor);
        for (int y = 0; y < in.h; y++)
          for (int x = 0; x < in.w; x++)
            out.Row(y)[x] = static_cast<float>(in.Row(y)[x] + iy.Row(y)[x]) * f;  // MultiplySum
      } else {
        convert(in, &(*planes)[c], factor, fp, meta.bit_depth.bits, meta.bit_depth.exp_bits);
        if (rgb_from_gray) {
          (*planes)[1] = (*planes)[0];
          (*planes)[2] = (*planes)[0];
          break;
        }
      }
    }
    if (rgb_from_gray) c = 1;
  }
  for (size_t ec = 0; ec < meta.extra.size(); ec++, c++) {
    const ExtraChannelInfo& eci = meta.extra[ec];
    bool fp = eci.bit_depth.floating_point;
    double factor = fp ? 0 : 1.0 / ((1u << eci.bit_depth.bits) - 1);
    JXLO_CHECK(c < gi.ch.size(), "modular: missing extra channel");
    convert(gi.ch[c], &(*planes)[3 + ec], factor, fp, eci.bit_depth.bits, eci.bit_depth.exp_bits);
  }
}

}  // namespace jxlo

#endif  // JXLO_FRAME_H_
