// jxl_b200 host planner, file level: container, codestream headers, frame header,
// TOC and the section walk (lib/jxl/decode.cc:969-1080, lib/jxl/dec_frame.cc:568-731).
// Produces a FramePlan for the CUDA kernels, or throws jxlb::Error naming the
// unsupported feature -- there is no CPU decode path in the product.
#ifndef JXLB_FRAME_PLAN_H_
#define JXLB_FRAME_PLAN_H_

#include "jxlb_plan.h"
#include "jxlb_splines.h"
#include "jxlb_vardct_frame.h"

namespace jxlb {

// Locates the codestream inside `data` without copying when possible.
struct CodestreamView {
  const uint8_t* data = nullptr;
  size_t size = 0;
  std::vector<uint8_t> storage;  // only for multi-box (jxlp) files
};

inline CodestreamView FindCodestream(const uint8_t* data, size_t size) {
  static const uint8_t kSig[12] = {0, 0, 0, 0xC, 'J', 'X', 'L', ' ', 0xD, 0xA, 0x87, 0xA};
  CodestreamView v;
  if (size >= 2 && data[0] == 0xFF && data[1] == 0x0A) {
    v.data = data;
    v.size = size;
    return v;
  }
  JXLB_CHECK(size >= 12 && std::memcmp(data, kSig, 12) == 0, "not a JPEG XL file");
  size_t pos = 0;
  int parts = 0;
  while (pos + 8 <= size) {
    uint64_t box = (uint64_t{data[pos]} << 24) | (data[pos + 1] << 16) | (data[pos + 2] << 8) | data[pos + 3];
    const uint8_t* type = data + pos + 4;
    size_t hdr = 8;
    if (box == 1) {
      JXLB_CHECK(pos + 16 <= size, "truncated box header");
      box = 0;
      for (int i = 0; i < 8; i++) box = (box << 8) | data[pos + 8 + i];
      hdr = 16;
    }
    size_t end = box == 0 ? size : pos + box;
    JXLB_CHECK(end <= size && end >= pos + hdr, "bad box size");
    bool c = !std::memcmp(type, "jxlc", 4), p = !std::memcmp(type, "jxlp", 4);
    if (c || p) {
      size_t skip = p ? 4 : 0;
      JXLB_CHECK(end >= pos + hdr + skip, "bad jxlp box");
      if (parts == 0) {
        v.data = data + pos + hdr + skip;
        v.size = end - (pos + hdr + skip);
      } else {
        if (parts == 1) v.storage.assign(v.data, v.data + v.size);
        v.storage.insert(v.storage.end(), data + pos + hdr + skip, data + end);
      }
      parts++;
    }
    pos = end;
  }
  JXLB_CHECK(parts > 0, "container without codestream");
  if (parts > 1) {
    v.data = v.storage.data();
    v.size = v.storage.size();
  }
  return v;
}

struct BasicInfo {
  uint32_t xsize = 0, ysize = 0;
  ImageMetadata meta;
};

inline BasicInfo ReadBasicInfo(const uint8_t* cs, size_t size, BitReader* out_br = nullptr) {
  BitReader br(cs, size);
  JXLB_CHECK(br.Read(16) == 0x0AFF, "bad codestream signature");
  BasicInfo bi;
  SizeHeader sh = ReadSizeHeader(br);
  bi.xsize = sh.xsize;
  bi.ysize = sh.ysize;
  ReadImageMetadata(br, &bi.meta);
  ReadCustomTransformData(br, &bi.meta);
  br.CheckInBounds();
  if (out_br) *out_br = br;
  return bi;
}

constexpr uint32_t kNumQuantTables = 17;

// Plans the (single) frame of a lossless / non-XYB Modular codestream.
// `cs` must stay alive until the batch has been uploaded; DevStream::bit_pos is
// relative to cs.
inline void PlanCodestream(const uint8_t* cs, size_t cs_size, const PixelFormat& fmt, FramePlan* plan, ProbeCtx* pc = nullptr) {
  BitReader br;
  BasicInfo bi = ReadBasicInfo(cs, cs_size, &br);
  const ImageMetadata& meta = bi.meta;
  plan->xsize = bi.xsize;
  plan->ysize = bi.ysize;
  plan->meta = meta;
  plan->pixels = static_cast<uint64_t>(bi.xsize) * bi.ysize;
  JXLB_CHECK(!meta.color.want_icc, "unsupported: embedded ICC profile");
  JXLB_CHECK(!meta.have_preview, "unsupported: preview frame");
  SizeHeader size;
  size.xsize = bi.xsize;
  size.ysize = bi.ysize;
  RefSlot refs[4];
  // Frames: any number of reference-only frames (kReferenceOnly: Modular, saved before the colour transform; what
  // libjxl's encoder writes for patches, lib/jxl/enc_patch_dictionary.cc), then the one regular frame.
  for (int frame_index = 0;; frame_index++) {
  JXLB_CHECK(frame_index < 16, "unsupported: more than 16 frames");
  br.AlignToByte();
  FrameHeader fh;
  ReadFrameHeader(br, size, meta, false, &fh);
  const bool is_ref = fh.frame_type == kReferenceOnly;
  if (is_ref) {
    JXLB_CHECK(fh.is_modular && fh.save_before_color_transform && fh.color_transform == kCTXYB && meta.extra.empty(),
               "unsupported: reference frame that is not an XYB Modular frame");
    JXLB_CHECK(fh.flags == 0, "unsupported: image features inside a reference frame");
  } else {
    JXLB_CHECK(fh.frame_type == kRegularFrame && fh.is_last, "unsupported: multi-frame codestream (animation / layers)");
    JXLB_CHECK(!fh.is_modular || fh.color_transform == kCTNone, "unsupported: XYB / YCbCr Modular frame");
    JXLB_CHECK(!fh.custom_size_or_origin, "unsupported: cropped frame");
    JXLB_CHECK(!fh.is_modular || !(fh.flags & (kFlagPatches | kFlagNoise | kFlagUseDcFrame)),
               "unsupported: patches / noise / DC frame on a Modular frame");
    JXLB_CHECK(!(fh.is_modular && (fh.flags & kFlagSplines)) || (!meta.color.IsGray() && fmt.num_channels >= 3),
               "unsupported: splines on a grey image");
    JXLB_CHECK(fh.blending.mode == kReplace, "unsupported: blending");
  }
  JXLB_CHECK(fh.upsampling == 1 || (!fh.is_modular && !is_ref), "unsupported: upsampling of a Modular or reference-only frame");
  for (uint32_t u : fh.ec_upsampling) JXLB_CHECK(u == 1, "unsupported: extra-channel upsampling");
  JXLB_CHECK(!fh.is_modular || (!fh.lf.gab && fh.lf.epf_iters == 0), "unsupported: loop filter on a Modular frame");
  FrameDimensions dim = ToFrameDimensions(fh);
  const size_t num_passes = fh.passes.num_passes;
  const size_t entries = NumTocEntries(dim.num_groups, dim.num_dc_groups, num_passes);
  Toc toc = ReadToc(br, entries);
  const size_t base = br.BitPos() / 8;
  JXLB_CHECK(base + toc.total <= cs_size, "truncated frame");

  if (!fh.is_modular) {
    PlanVarDCTFrame(cs, cs_size, fh, dim, meta, toc, base, fmt, plan, pc, refs);
    DevFrameOut& fo = plan->out;
    fo = DevFrameOut{};
    fo.xsize = bi.xsize;
    fo.ysize = bi.ysize;
    fo.num_channels = fmt.num_channels;
    fo.data_type = fmt.data_type;
    fo.orient = OrientBits(meta.orientation, fmt);
    fo.stride = OutputStride((fo.orient & 4) ? bi.ysize : bi.xsize, fmt);
    fo.vardct = 1;
    for (uint32_t c = 0; c < 4; c++) fo.plane[c] = kNoPlane;
    return;
  }
  FramePlanner planner(plan);
  const bool is_gray = meta.color.IsGray() && fh.color_transform == kCTNone;
  const size_t nb_chans = is_gray ? 1 : 3;
  const size_t nb_extra = meta.extra.size();

  auto section = [&](size_t i, uint64_t* bit_base) {
    *bit_base = (base + toc.offsets[i]) * 8;
    return BitReader(cs + base + toc.offsets[i], toc.logical_size[i]);
  };

  HImage full;
  full.bitdepth = meta.bit_depth.bits;
  HostTree global_tree;
  GroupHeader global_header;

  float dc_quant[3] = {1.0f / 4096, 1.0f / 512, 1.0f / 256};
  SplineState splines;
  auto dc_global = [&](BitReader& r, uint64_t bit_base) {
    if (fh.flags & kFlagSplines) {  // (lib/jxl/dec_frame.cc:286-305; the base colour correlation of a Modular frame is the default)
      ReadSplines(r, dim.xsize * dim.ysize, &splines);
      InitSplineDrawCache(dim.xsize_upsampled, dim.ysize_upsampled, 0.0f, 1.0f, &splines);
    }
    if (!r.ReadBool()) {  // DequantMatrices::DecodeDC (lib/jxl/quant_weights.cc:507-520): the XYB scale of Modular frames
      for (int c = 0; c < 3; c++) {
        dc_quant[c] = ReadF16(r) * (1.0f / 128.0f);
        JXLB_CHECK(dc_quant[c] >= 1e-8f, "bad DC quantisation step");
      }
    }
    bool has_tree = r.ReadBool();
    if (has_tree) {
      size_t limit = std::min<size_t>(size_t{1} << 22, 1024 + dim.xsize * dim.ysize * (nb_chans + nb_extra) / 16);
      global_tree = planner.ReadTreeAndCode(r, limit);
    }
    if (meta.bit_depth.bits >= 32) {
      JXLB_CHECK(meta.bit_depth.bits == 32 && meta.bit_depth.floating_point, "unsupported 32-bit integer samples");
    }
    for (size_t c = 0; c < nb_chans + nb_extra; c++) {
      HChan ch;
      ch.w = dim.xsize;
      ch.h = dim.ysize;
      ch.plane = planner.NewPlane(ch.w, ch.h);
      full.ch.push_back(ch);
    }
    global_header = planner.PlanStream(r, bit_base, full, 0, dim.group_dim, global_tree);
  };

  std::vector<DevOp> ops = plan->ops;  // (continues after earlier frames) group programs first, then global levels
  auto group = [&](BitReader& r, uint64_t bit_base, size_t x0, size_t y0, size_t xs, size_t ys, int min_shift,
                   int max_shift, uint32_t stream_id) {
    HImage gi;
    gi.bitdepth = full.bitdepth;
    size_t c = full.nb_meta;
    for (; c < full.ch.size(); c++) {
      if (static_cast<size_t>(full.ch[c].w) > dim.group_dim || static_cast<size_t>(full.ch[c].h) > dim.group_dim) break;
    }
    struct Dest { size_t c; int x, y, w, h; };
    std::vector<Dest> dests;
    for (; c < full.ch.size(); c++) {
      const HChan& fc = full.ch[c];
      int shift = std::min(fc.hshift, fc.vshift);
      if (shift > max_shift || shift < min_shift) continue;
      int rx = x0 >> fc.hshift, ry = y0 >> fc.vshift;
      int rw = xs >> fc.hshift, rh = ys >> fc.vshift;
      if (rx >= fc.w || ry >= fc.h) continue;
      rw = std::min(rw, fc.w - rx);
      rh = std::min(rh, fc.h - ry);
      if (rw <= 0 || rh <= 0) continue;
      HChan gc;
      gc.w = rw;
      gc.h = rh;
      gc.hshift = fc.hshift;
      gc.vshift = fc.vshift;
      gc.plane = planner.NewPlane(rw, rh);
      gi.ch.push_back(gc);
      dests.push_back(Dest{c, rx, ry, rw, rh});
    }
    if (gi.ch.empty()) return;
    GroupHeader hdr = planner.PlanStream(r, bit_base, gi, stream_id, 0xFFFFFF, global_tree);
    DevProgram prog;
    prog.op_begin = ops.size();
    planner.EmitInverse(gi, hdr.wp, &ops);
    JXLB_CHECK(gi.ch.size() == dests.size(), "modular group: channel count changed");
    for (size_t i = 0; i < dests.size(); i++) {
      JXLB_CHECK(gi.ch[i].w == dests[i].w && gi.ch[i].h == dests[i].h, "modular group: channel size changed");
      DevOp cp{};
      cp.kind = kOpCopy;
      cp.a = gi.ch[i].plane;
      cp.b = full.ch[dests[i].c].plane;
      cp.p0 = dests[i].x;
      cp.p1 = dests[i].y;
      ops.push_back(cp);
    }
    prog.op_end = ops.size();
    plan->group_programs.push_back(prog);
  };
  auto dc_group = [&](BitReader& r, uint64_t bit_base, size_t g) {
    size_t gx = g % dim.xsize_dc_groups, gy = g / dim.xsize_dc_groups;
    group(r, bit_base, gx * dim.dc_group_dim, gy * dim.dc_group_dim, dim.dc_group_dim, dim.dc_group_dim, 3, 1000,
          1 + dim.num_dc_groups + g);
  };
  auto ac_group = [&](BitReader& r, uint64_t bit_base, size_t g, size_t pass) {
    int min_shift, max_shift;
    fh.passes.DownsamplingBracket(pass, &min_shift, &max_shift);
    size_t gx = g % dim.xsize_groups, gy = g / dim.xsize_groups;
    group(r, bit_base, gx * dim.group_dim, gy * dim.group_dim, dim.group_dim, dim.group_dim, min_shift, max_shift,
          1 + 3 * dim.num_dc_groups + kNumQuantTables + dim.num_groups * pass + g);
  };

  uint64_t bb;
  if (entries == 1) {
    // One section holds everything: the sub-streams follow each other bit by bit, so
    // a later one starts where the device decode of the earlier one ends. With a
    // single group every channel fits the global stream (or none does), hence at
    // most one of them carries samples; anything else cannot be planned on the host.
    const size_t before = plan->streams.size();
    BitReader r = section(0, &bb);
    dc_global(r, bb);
    const size_t after_global = plan->streams.size();
    dc_group(r, bb, 0);
    ac_group(r, bb, 0, 0);
    JXLB_CHECK(after_global == before || plan->streams.size() == after_global,
               "unsupported: single-section frame with chained Modular streams");
    JXLB_CHECK(plan->streams.size() - before <= 1, "unsupported: single-section frame with chained Modular streams");
  } else {
    {
      BitReader r = section(0, &bb);
      dc_global(r, bb);
    }
    for (size_t g = 0; g < dim.num_dc_groups; g++) {
      BitReader r = section(1 + g, &bb);
      dc_group(r, bb, g);
    }
    for (size_t pass = 0; pass < num_passes; pass++) {
      for (size_t g = 0; g < dim.num_groups; g++) {
        BitReader r = section(2 + dim.num_dc_groups + pass * dim.num_groups + g, &bb);
        ac_group(r, bb, g, pass);
      }
    }
  }
  // global inverse transforms: one op per level
  std::vector<DevOp> global_ops;
  planner.EmitInverse(full, global_header.wp, &global_ops);
  for (const DevOp& op : global_ops) {
    DevProgram lvl;
    lvl.op_begin = ops.size();
    ops.push_back(op);
    lvl.op_end = ops.size();
    plan->frame_levels.push_back(lvl);
  }
  plan->ops = ops;
  JXLB_CHECK(full.ch.size() == nb_chans + nb_extra, "modular: unexpected channel count after transforms");

  if (is_ref) {
    // ModularImageToDecodedRect for an XYB frame (lib/jxl/dec_modular.cc:534-708): three float planes in the
    // codestream's part of farena, kept for the patches of the frame that follows.
    DevRefFrame rf{};
    rf.plane_y = full.ch[0].plane;
    rf.plane_x = full.ch[1].plane;
    rf.plane_b = full.ch[2].plane;
    rf.w = dim.xsize;
    rf.h = dim.ysize;
    for (int c = 0; c < 3; c++) {
      JXLB_CHECK(full.ch[c].w == static_cast<int>(dim.xsize) && full.ch[c].h == static_cast<int>(dim.ysize),
                 "unsupported: subsampled channel in a reference frame");
      rf.factor[c] = dc_quant[c];
      rf.dst[c] = plan->v.farena_size;
      plan->v.farena_size += (static_cast<uint64_t>(rf.w) * rf.h + 3) & ~uint64_t{3};
    }
    plan->v.ref_frames.push_back(rf);
    RefSlot& slot = refs[fh.save_as_reference & 3];
    slot.valid = true;
    slot.w = rf.w;
    slot.h = rf.h;
    for (int c = 0; c < 3; c++) slot.off[c] = rf.dst[c];
    br.SeekTo((base + toc.total) * 8);
    continue;
  }

  // output mapping (lib/jxl/dec_modular.cc:534-708 + stage_write)
  DevFrameOut& fo = plan->out;
  fo = DevFrameOut{};
  fo.xsize = bi.xsize;
  fo.ysize = bi.ysize;
  fo.num_channels = fmt.num_channels;
  fo.data_type = fmt.data_type;
  fo.big_endian = fmt.endianness == 2;
  fo.orient = OrientBits(meta.orientation, fmt);
  fo.stride = OutputStride((fo.orient & 4) ? bi.ysize : bi.xsize, fmt);
  if (!splines.segments.empty()) {  // the draw cache for DevSplinePixel
    fo.has_splines = 1;
    PackSplineDrawCache(splines, bi.ysize, plan, &fo.spl_seg, &fo.spl_rows, &fo.spl_idx);
  }
  const uint32_t num_color = fmt.num_channels < 3 ? 1 : 3;
  const bool want_alpha = fmt.num_channels == 2 || fmt.num_channels == 4;
  const int alpha = meta.AlphaIndex();
  for (uint32_t c = 0; c < fmt.num_channels; c++) {
    const BitDepth* bd = &meta.bit_depth;
    size_t src;
    if (c < num_color) {
      src = is_gray ? 0 : c;
    } else if (want_alpha && alpha >= 0) {
      src = nb_chans + alpha;
      bd = &meta.extra[alpha].bit_depth;
    } else {
      fo.plane[c] = kNoPlane;
      continue;
    }
    const HChan& ch = full.ch[src];
    JXLB_CHECK(ch.w == static_cast<int>(bi.xsize) && ch.h == static_cast<int>(bi.ysize), "unsupported: subsampled channel");
    fo.plane[c] = ch.plane;
    if (bd->floating_point) {
      fo.is_float[c] = bd->bits | (bd->exp_bits << 8);
    } else {
      JXLB_CHECK(full.bitdepth < 23 && bd->bits < 23, "unsupported: integer samples wider than 22 bits");
      fo.factor[c] = static_cast<float>(1.0 / ((1u << bd->bits) - 1));
    }
  }
  return;
  }  // frames
}

}  // namespace jxlb

#endif  // JXLB_FRAME_PLAN_H_
