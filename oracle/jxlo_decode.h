// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
//
// Top level: container, codestream headers, frame loop, blending onto the
// canvas, conversion to the caller's pixel format. Restates
//   lib/jxl/decode.cc:969-1080 (header sequence), :1149-1500 (frame loop),
//   lib/jxl/dec_frame.cc:568-731 (section order),
//   lib/jxl/render_pipeline/stage_write.cc:51-113 (dither + rounding), :380-430.
#ifndef JXLO_DECODE_H_
#define JXLO_DECODE_H_

#include <cmath>

#include "jxlo_frame.h"
#include "jxlo_vardct.h"
#include "jxlo_render.h"

namespace jxlo {

// ISO-BMFF container -> bare codestream (concatenated jxlc / jxlp payloads).
inline std::vector<uint8_t> ExtractCodestream(const uint8_t* data, size_t size) {
  static const uint8_t kSig[12] = {0, 0, 0, 0xC, 'J', 'X', 'L', ' ', 0xD, 0xA, 0x87, 0xA};
  if (size >= 2 && data[0] == 0xFF && data[1] == 0x0A) return std::vector<uint8_t>(data, data + size);
  JXLO_CHECK(size >= 12 && std::memcmp(data, kSig, 12) == 0, "not a JPEG XL file");
  std::vector<uint8_t> cs;
  size_t pos = 0;
  while (pos + 8 <= size) {
    uint64_t box = (uint64_t{data[pos]} << 24) | (data[pos + 1] << 16) | (data[pos + 2] << 8) | data[pos + 3];
    const uint8_t* type = data + pos + 4;
    size_t hdr = 8;
    if (box == 1) {
      JXLO_CHECK(pos + 16 <= size, "truncated box header");
      box = 0;
      for (int i = 0; i < 8; i++) box = (box << 8) | data[pos + 8 + i];
      hdr = 16;
    }
    size_t end = box == 0 ? size : pos + box;
    JXLO_CHECK(end <= size && end >= pos + hdr, "bad box size");
    if (!std::memcmp(type, "jxlc", 4)) {
      cs.insert(cs.end(), data + pos + hdr, data + end);
    } else if (!std::memcmp(type, "jxlp", 4)) {
      JXLO_CHECK(end >= pos + hdr + 4, "bad jxlp box");
      cs.insert(cs.end(), data + pos + hdr + 4, data + end);
    }
    pos = end;
  }
  JXLO_CHECK(!cs.empty(), "container without codestream");
  return cs;
}

struct DecodedImage {
  uint32_t xsize = 0, ysize = 0;
  ImageMetadata meta;
  std::vector<Plane> planes;  // 3 colour (R,G,B or grey replicated) + extra channels, non-linear, canvas size
  std::vector<std::string> frame_info;
};

// One frame: header, TOC, sections, render. Returns the frame header.
inline FrameHeader DecodeFrame(BitReader& br, CodestreamState* cs, bool is_preview, std::vector<Plane>* out_planes,
                               bool* out_is_xyb, std::string* info) {
  br.AlignToByte();
  FrameHeader fh;
  ReadFrameHeader(br, cs->size, cs->meta, is_preview, &fh);
  FrameDimensions dim = ToFrameDimensions(fh);
  const size_t num_passes = fh.passes.num_passes;
  const size_t entries = NumTocEntries(dim.num_groups, dim.num_dc_groups, num_passes);
  Toc toc = ReadToc(br, entries);
  const size_t base = br.BitPos() / 8;
  JXLO_CHECK(base + toc.total <= br.Size(), "truncated frame");
  const uint8_t* data = br.Data();
  if (info) {
    *info = std::string(fh.is_modular ? "modular" : "vardct") + " type=" + std::to_string(fh.frame_type) + " " +
            std::to_string(fh.xsize) + "x" + std::to_string(fh.ysize) + " groups=" + std::to_string(dim.num_groups) +
            " flags=" + std::to_string(fh.flags) + " gab=" + std::to_string(fh.lf.gab) +
            " epf=" + std::to_string(fh.lf.epf_iters) + " passes=" + std::to_string(num_passes);
  }
  JXLO_CHECK(!(fh.flags & kFlagUseDcFrame), "DC frames are not supported by the oracle");
  JXLO_CHECK(fh.frame_type != kDCFrame, "DC frames are not supported by the oracle");

  ModularFrameState ms;
  std::unique_ptr<VarDCTState> vs;
  if (!fh.is_modular) vs.reset(new VarDCTState(fh, dim, cs->meta));
  FeatureState feat;
  float dc_quant[3] = {1.0f / 4096, 1.0f / 512, 1.0f / 256};

  auto section = [&](size_t i) { return BitReader(data + base + toc.offsets[i], toc.logical_size[i]); };
  auto dc_global = [&](BitReader& r) {
    if (fh.flags & kFlagPatches) ReadPatches(r, dim, cs->meta, *cs, &feat);
    if (fh.flags & kFlagSplines) ReadSplines(r, dim.xsize * dim.ysize, &feat.splines);
    JXLO_CHECK(!(fh.flags & kFlagNoise), "noise is not supported by the oracle");
    if (!r.ReadBool()) {  // DequantMatrices::DecodeDC, lib/jxl/quant_weights.cc:507-520
      for (int c = 0; c < 3; c++) {
        dc_quant[c] = ReadF16(r) * (1.0f / 128.0f);
        JXLO_CHECK(dc_quant[c] >= 1e-8f, "bad dc quant");
      }
    }
    if (vs) {
      for (int c = 0; c < 3; c++) vs->dc_quant[c] = dc_quant[c];
      VarDCTReadGlobalDC(r, vs.get());
    }
    // (the draw cache uses the base colour correlation: lib/jxl/dec_frame.cc:300-305; defaults in a Modular frame)
    if (fh.flags & kFlagSplines)
      InitSplineDrawCache(dim.xsize_upsampled, dim.ysize_upsampled, vs ? vs->base_x : 0.0f, vs ? vs->base_b : 1.0f, &feat.splines);
    ModularDecodeGlobal(r, fh, dim, cs->meta, &ms);
  };
  auto dc_group = [&](BitReader& r, size_t g) {
    if (vs) VarDCTReadDCGroup(r, vs.get(), &ms, g);
    size_t gx = g % dim.xsize_dc_groups, gy = g / dim.xsize_dc_groups;
    ModularDecodeGroup(r, dim, &ms, gx * dim.dc_group_dim, gy * dim.dc_group_dim, dim.dc_group_dim, dim.dc_group_dim, 3,
                       1000, StreamModularDC(dim, g));
    if (vs) VarDCTReadACMetadata(r, vs.get(), &ms, g);
  };
  auto ac_global = [&](BitReader& r) {
    if (vs) VarDCTReadGlobalAC(r, vs.get(), &ms);
  };
  auto ac_group = [&](BitReader& r, size_t g, size_t pass) {
    if (vs) VarDCTReadACGroup(r, vs.get(), g, pass);
    int min_shift, max_shift;
    fh.passes.DownsamplingBracket(pass, &min_shift, &max_shift);
    size_t gx = g % dim.xsize_groups, gy = g / dim.xsize_groups;
    ModularDecodeGroup(r, dim, &ms, gx * dim.group_dim, gy * dim.group_dim, dim.group_dim, dim.group_dim, min_shift,
                       max_shift, StreamModularAC(dim, g, pass));
  };

  if (entries == 1) {
    BitReader r = section(0);
    dc_global(r);
    dc_group(r, 0);
    if (vs) VarDCTFinalizeDC(vs.get());
    ac_global(r);
    ac_group(r, 0, 0);
    r.CheckInBounds();
  } else {
    {
      BitReader r = section(0);
      dc_global(r);
      r.CheckInBounds();
    }
    for (size_t g = 0; g < dim.num_dc_groups; g++) {
      BitReader r = section(1 + g);
      dc_group(r, g);
      r.CheckInBounds();
    }
    if (vs) VarDCTFinalizeDC(vs.get());
    {
      BitReader r = section(1 + dim.num_dc_groups);
      ac_global(r);
      r.CheckInBounds();
    }
    for (size_t pass = 0; pass < num_passes; pass++) {
      for (size_t g = 0; g < dim.num_groups; g++) {
        BitReader r = section(2 + dim.num_dc_groups + pass * dim.num_groups + g);
        ac_group(r, g, pass);
        r.CheckInBounds();
      }
    }
  }
  br.Skip((base + toc.total) * 8 - br.BitPos());

  // planes: 3 colour + extras
  std::vector<Plane> planes(3 + cs->meta.extra.size());
  ModularToFloat(fh, cs->meta, &ms, dc_quant, &planes);
  if (vs) VarDCTToPixels(vs.get(), &planes);
  RenderFrame(fh, dim, *cs, vs.get(), feat, &planes, out_is_xyb);
  *out_planes = std::move(planes);
  return fh;
}

inline DecodedImage DecodeCodestream(const uint8_t* file, size_t file_size) {
  std::vector<uint8_t> cs_bytes = ExtractCodestream(file, file_size);
  BitReader br(cs_bytes.data(), cs_bytes.size());
  JXLO_CHECK(br.Read(16) == 0x0AFF, "bad codestream signature");
  CodestreamState cs;
  cs.size = ReadSizeHeader(br);
  ReadImageMetadata(br, &cs.meta);
  ReadCustomTransformData(br, &cs.meta);
  br.CheckInBounds();
  JXLO_CHECK(!cs.meta.color.want_icc, "embedded ICC profiles are not supported by the oracle");
  DecodedImage img;
  img.xsize = cs.size.xsize;
  img.ysize = cs.size.ysize;
  img.meta = cs.meta;
  const size_t nplanes = 3 + cs.meta.extra.size();
  img.planes.assign(nplanes, Plane(img.xsize, img.ysize));
  if (cs.meta.have_preview) {
    std::vector<Plane> p;
    bool xyb;
    DecodeFrame(br, &cs, true, &p, &xyb, nullptr);
  }
  for (;;) {
    std::vector<Plane> planes;
    bool is_xyb = false;
    std::string info;
    FrameHeader fh = DecodeFrame(br, &cs, false, &planes, &is_xyb, &info);
    img.frame_info.push_back(info);
    if (fh.CanBeReferenced() || fh.frame_type == kReferenceOnly) {
      FrameBuffer& ref = cs.reference[fh.save_as_reference];
      ref.planes = planes;
      ref.is_xyb = is_xyb;
      cs.reference_valid[fh.save_as_reference] = true;
    }
    if (fh.frame_type == kRegularFrame || fh.frame_type == kSkipProgressive) {
      JXLO_CHECK(!is_xyb, "internal: regular frame left in XYB");
      // Blending (lib/jxl/blending.cc): only kReplace is restated.
      JXLO_CHECK(fh.blending.mode == kReplace, "only kReplace blending is supported by the oracle");
      for (size_t c = 0; c < nplanes; c++) {
        const Plane& src = planes[c];
        Plane& dst = img.planes[c];
        for (int y = 0; y < src.h; y++) {
          int dy = y + fh.y0;
          if (dy < 0 || dy >= dst.h) continue;
          for (int x = 0; x < src.w; x++) {
            int dx = x + fh.x0;
            if (dx < 0 || dx >= dst.w) continue;
            dst.Row(dy)[dx] = src.Row(y)[x];
          }
        }
      }
    }
    if (fh.is_last) break;
  }
  return img;
}

// ---------------------------------------------------------------- output
enum DataType { kTypeFloat = 0, kTypeUint8 = 2, kTypeUint16 = 3, kTypeFloat16 = 5 };
enum Endianness { kNativeEndian = 0, kLittleEndian = 1, kBigEndian = 2 };

// lib/jxl/render_pipeline/stage_write.cc:51-84 (one 8x8 period of the table)
static const float kDither8x8[64] = {
    -0.4921875f, 0.0078125f,  -0.3671875f, 0.1328125f,  -0.4609375f, 0.0390625f,  -0.3359375f, 0.1640625f,
    0.2578125f,  -0.2421875f, 0.3828125f,  -0.1171875f, 0.2890625f,  -0.2109375f, 0.4140625f,  -0.0859375f,
    -0.3046875f, 0.1953125f,  -0.4296875f, 0.0703125f,  -0.2734375f, 0.2265625f,  -0.3984375f, 0.1015625f,
    0.4453125f,  -0.0546875f, 0.3203125f,  -0.1796875f, 0.4765625f,  -0.0234375f, 0.3515625f,  -0.1484375f,
    -0.4453125f, 0.0546875f,  -0.3203125f, 0.1796875f,  -0.4765625f, 0.0234375f,  -0.3515625f, 0.1484375f,
    0.3046875f,  -0.1953125f, 0.4296875f,  -0.0703125f, 0.2734375f,  -0.2265625f, 0.3984375f,  -0.1015625f,
    -0.2578125f, 0.2421875f,  -0.3828125f, 0.1171875f,  -0.2890625f, 0.2109375f,  -0.4140625f, 0.0859375f,
    0.4921875f,  -0.0078125f, 0.3671875f,  -0.1328125f, 0.4609375f,  -0.0390625f, 0.3359375f,  -0.1640625f};

inline uint16_t FloatToHalf(float f) {  // round-to-nearest-even, like hwy DemoteTo(float16)
  uint32_t b;
  std::memcpy(&b, &f, 4);
  uint32_t sign = (b >> 16) & 0x8000;
  int32_t exp = static_cast<int32_t>((b >> 23) & 0xFF) - 127 + 15;
  uint32_t mant = b & 0x7FFFFF;
  if (((b >> 23) & 0xFF) == 0xFF) return sign | 0x7C00 | (mant ? 0x200 : 0);
  if (exp >= 31) return sign | 0x7C00;
  if (exp <= 0) {
    if (exp < -10) return sign;
    mant |= 0x800000;
    uint32_t shift = 14 - exp;
    uint32_t h = mant >> shift;
    uint32_t rem = mant & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) h++;
    return sign | h;
  }
  uint32_t h = (exp << 10) | (mant >> 13);
  uint32_t rem = mant & 0x1FFF;
  if (rem > 0x1000 || (rem == 0x1000 && (h & 1))) h++;
  return sign | h;
}

inline size_t BytesPerSample(uint32_t data_type) {
  return data_type == kTypeUint8 ? 1 : (data_type == kTypeFloat ? 4 : 2);
}

inline size_t OutputStride(uint32_t xsize, uint32_t num_channels, uint32_t data_type, size_t align) {
  size_t row = static_cast<size_t>(xsize) * num_channels * BytesPerSample(data_type);
  if (align > 1) row = DivCeil(row, align) * align;
  return row;
}

// WriteToOutputStage with the scalar dither semantics (pattern indexed by the
// absolute pixel position). num_channels: 1 grey, 2 grey+alpha, 3 RGB, 4 RGBA.
// undo_orientation (stage_write.cc:131-135, :163-172, :345-366, :271-288; libjxl's default, JxlDecoderSetKeepOrientation
// false): the row is flipped in y, the pixels of a row in x -- the dither pattern is indexed by the flipped position --
// and a transposing orientation writes pixel (x', y') at row x', column y'.
inline void WritePixels(const DecodedImage& img, uint32_t num_channels, uint32_t data_type, uint32_t endianness,
                        size_t align, uint8_t* out, bool undo_orientation = false) {
  const uint32_t num_color = num_channels < 3 ? 1 : 3;
  const bool want_alpha = num_channels == 2 || num_channels == 4;
  const int alpha_idx = img.meta.AlphaIndex();
  const uint32_t o = undo_orientation ? img.meta.orientation : 1;
  const bool flip_x = o == 2 || o == 3 || o == 8 || o == 7, flip_y = o == 4 || o == 3 || o == 6 || o == 7;
  const bool transpose = o >= 5;
  const size_t stride = OutputStride(transpose ? img.ysize : img.xsize, num_channels, data_type, align);
  const bool swap = (endianness == kBigEndian);  // host is little-endian
  const uint32_t bits = data_type == kTypeUint8 ? 8 : 16;
  const float mul = static_cast<float>((1u << bits) - 1);
  for (uint32_t sy = 0; sy < img.ysize; sy++) {
    for (uint32_t sx = 0; sx < img.xsize; sx++) {
      const uint32_t x = flip_x ? img.xsize - 1 - sx : sx, y = flip_y ? img.ysize - 1 - sy : sy;
      uint8_t* row = out + stride * (transpose ? x : y);
      for (uint32_t c = 0; c < num_channels; c++) {
        float v;
        if (c < num_color) {
          v = img.planes[c].Row(sy)[sx];
        } else if (want_alpha && alpha_idx >= 0) {
          v = img.planes[3 + alpha_idx].Row(sy)[sx];
        } else {
          v = 1.0f;
        }
        size_t idx = static_cast<size_t>(transpose ? y : x) * num_channels + c;
        if (data_type == kTypeUint8 || data_type == kTypeUint16) {
          v = v * mul;
          if (data_type == kTypeUint8) v += kDither8x8[(y % 8) * 8 + (x % 8)];
          if (!(v >= 0.0f)) v = 0.0f;  // Clamp(Zero, v, mul); NaN -> 0
          if (v > mul) v = mul;
          long r = std::lrintf(v);  // round-half-to-even under the default rounding mode
          if (data_type == kTypeUint8) {
            row[idx] = static_cast<uint8_t>(r);
          } else {
            uint16_t u = static_cast<uint16_t>(r);
            if (swap) u = static_cast<uint16_t>((u >> 8) | (u << 8));
            std::memcpy(row + idx * 2, &u, 2);
          }
        } else if (data_type == kTypeFloat16) {
          uint16_t u = FloatToHalf(v);
          if (swap) u = static_cast<uint16_t>((u >> 8) | (u << 8));
          std::memcpy(row + idx * 2, &u, 2);
        } else {
          uint32_t u;
          std::memcpy(&u, &v, 4);
          if (swap) u = __builtin_bswap32(u);
          std::memcpy(row + idx * 4, &u, 4);
        }
      }
    }
  }
}

}  // namespace jxlo

#endif  // JXLO_DECODE_H_
