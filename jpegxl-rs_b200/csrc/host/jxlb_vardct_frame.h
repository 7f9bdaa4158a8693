// jxl_b200 host planner: section walk of one VarDCT frame (lib/jxl/dec_frame.cc:568-731,
// :266-339, :367-476) producing the device descriptors of jxlb_vardct_desc.h.
#ifndef JXLB_VARDCT_FRAME_H_
#define JXLB_VARDCT_FRAME_H_

#include "jxlb_vardct_plan.h"

namespace jxlb {

struct PixelFormat {
  uint32_t num_channels = 4;
  uint32_t data_type = 2;   // JxlDataType
  uint32_t endianness = 0;  // JxlEndianness
  size_t align = 0;
  bool keep_orientation = false;  // JxlDecoderSetKeepOrientation: leave the image as coded
};

// The write stage's undo_orientation (lib/jxl/render_pipeline/stage_write.cc:271-288) as bits: 1 flip x, 2 flip y,
// 4 transpose (pixel (x', y') after the flips is stored at row x', column y').
inline uint32_t OrientBits(uint32_t orientation, const PixelFormat& f) {
  const uint32_t o = f.keep_orientation ? 1 : orientation;
  return ((o == 2 || o == 3 || o == 8 || o == 7) ? 1u : 0u) | ((o == 4 || o == 3 || o == 6 || o == 7) ? 2u : 0u) | (o >= 5 ? 4u : 0u);
}

inline size_t BytesPerSample(uint32_t data_type) { return data_type == 2 ? 1 : (data_type == 0 ? 4 : 2); }
inline size_t OutputStride(uint32_t xsize, const PixelFormat& f) {
  size_t row = static_cast<size_t>(xsize) * f.num_channels * BytesPerSample(f.data_type);
  if (f.align > 1) row = DivCeil(row, f.align) * f.align;
  return row;
}

// The spline draw cache as the device reads it (DevSplineAdd): kSplineSegmentWords floats per segment (the column span
// precomputed as two int32 in float words), ysize + 1 row offsets, the per-row index lists. Offsets into the plan's pools.
inline void PackSplineDrawCache(const SplineState& splines, uint32_t ysize, FramePlan* plan, uint64_t* seg_off, uint64_t* rows_off,
                                uint64_t* idx_off) {
  *seg_off = plan->spl_seg.size();
  for (const SplineSegment& sg : splines.segments) {
    auto bits = [](int64_t v) {
      const int32_t c = static_cast<int32_t>(std::min<int64_t>(std::max<int64_t>(v, INT32_MIN), INT32_MAX));
      float f;
      std::memcpy(&f, &c, 4);
      return f;
    };
    const float w[kSplineSegmentWords] = {sg.center_x, sg.center_y, sg.inv_sigma, sg.sigma_over_4_times_intensity,
                                          sg.color[0], sg.color[1], sg.color[2],
                                          bits(std::llround(sg.center_x - sg.maximum_distance)),
                                          bits(std::llround(sg.center_x + sg.maximum_distance) + 1), 0.0f};
    plan->spl_seg.insert(plan->spl_seg.end(), w, w + kSplineSegmentWords);
  }
  *rows_off = plan->spl_idx.size();
  JXLB_CHECK(splines.segment_y_start.size() == static_cast<size_t>(ysize) + 1, "internal: spline row table");
  plan->spl_idx.insert(plan->spl_idx.end(), splines.segment_y_start.begin(), splines.segment_y_start.end());
  *idx_off = plan->spl_idx.size();
  plan->spl_idx.insert(plan->spl_idx.end(), splines.segment_indices.begin(), splines.segment_indices.end());
}

// Plans the sections of a VarDCT frame. `br` is positioned after the TOC, `base` is the byte
// offset of the first section inside `cs`.
inline void PlanVarDCTFrame(const uint8_t* cs, size_t cs_size, const FrameHeader& fh, const FrameDimensions& dim,
                            const ImageMetadata& meta, const Toc& toc, size_t base, const PixelFormat& fmt, FramePlan* plan,
                            ProbeCtx* pc = nullptr, const RefSlot* refs = nullptr) {
  JXLB_CHECK(fh.Is444(), "unsupported: chroma-subsampled VarDCT frame");
  JXLB_CHECK(fh.passes.num_passes <= kMaxPasses, "too many passes");
  const size_t num_passes = fh.passes.num_passes;
  // Extra channels (alpha) of a VarDCT frame are a Modular image of their own inside the frame's sections
  // (lib/jxl/dec_frame.cc:266-365, :478-560; lib/jxl/dec_modular.cc:188-395): the global stream in DC global, then one
  // stream per AC group behind the group's coefficients for the channels larger than a group.
  const size_t nb_extra = meta.extra.size();
  if (nb_extra != 0) {
    plan->late_programs0 = plan->group_programs.size();
    plan->late_levels0 = plan->frame_levels.size();
    JXLB_CHECK(num_passes == 1, "unsupported: multi-pass VarDCT frame with extra channels");
    JXLB_CHECK(fh.upsampling == 1, "unsupported: upsampled VarDCT frame with extra channels");
    JXLB_CHECK(!(fh.flags & kFlagPatches), "unsupported: patches in a VarDCT frame with extra channels");
    for (const auto& bl : fh.ec_blending) JXLB_CHECK(bl.mode == kReplace, "unsupported: blending");
    for (const auto& e : meta.extra) JXLB_CHECK(e.dim_shift == 0, "unsupported: subsampled extra channel");
  }
  // A frame with one group and one pass has a single section in which DC global, the DC group, AC global and the AC
  // group follow each other bit by bit (lib/jxl/dec_frame.cc:597-677): `whole` walks it, and the positions that only
  // the device can know (the end of the DC / AC-metadata chain) come from the probe rounds.
  const bool single = toc.offsets.size() == 1;
  BitReader whole(cs + base + toc.offsets[0], toc.logical_size[0]);
  plan->is_vardct = true;
  VarDCTPlan& v = plan->v;
  DevVFrame& vf = v.vf;
  vf = DevVFrame{};
  JXLB_CHECK(!(fh.flags & (kFlagNoise | kFlagUseDcFrame)), "unsupported: noise / DC frame");
  FramePlanner planner(plan);
  const size_t W = dim.xsize_blocks, H = dim.ysize_blocks, nb = W * H;

  auto section = [&](size_t i, uint64_t* bit_base) {
    if (single) i = 0;
    *bit_base = (base + toc.offsets[i]) * 8;
    return single ? whole : BitReader(cs + base + toc.offsets[i], toc.logical_size[i]);
  };
  uint64_t bb = 0;

  // ---- DC global (lib/jxl/dec_frame.cc:266-313, :61-77)
  float dc_quant[3] = {1.0f / 4096, 1.0f / 512, 1.0f / 256};
  HostBlockCtx bctx;
  HostTree global_tree;
  HImage full;  // the extra channels
  GroupHeader global_header;
  {
    BitReader r = section(0, &bb);
    if (fh.flags & kFlagPatches) {
      // PatchDictionary::Decode, lib/jxl/dec_patch_dictionary.cc:28-200 (contexts: :21-32 of the header)
      JXLB_CHECK(refs != nullptr, "patches without reference frames");
      EntropyCode code;
      ReadEntropyCode(r, 10, &code);
      SymbolReader reader(&code, r);
      auto read_num = [&](uint32_t ctx) { return reader.ReadUint(ctx, r); };
      const size_t num_ref_patch = read_num(0);
      const size_t max_ref_patches = 1024 + dim.xsize_padded * dim.ysize_padded / 4, max_patches = max_ref_patches * 4;
      JXLB_CHECK(num_ref_patch <= max_ref_patches, "too many patches");
      size_t total = 0;
      uint32_t last_x = 0, last_y = 0;
      for (size_t id = 0; id < num_ref_patch; id++) {
        const uint32_t ref = read_num(1);
        JXLB_CHECK(ref < 4 && refs[ref].valid, "invalid patch reference frame");
        const RefSlot& slot = refs[ref];
        DevPatch p{};
        for (int c = 0; c < 3; c++) p.src[c] = slot.off[c];
        p.src_w = slot.w;
        p.x0 = read_num(3);
        p.y0 = read_num(3);
        p.xsize = read_num(2) + 1;
        p.ysize = read_num(2) + 1;
        JXLB_CHECK(static_cast<uint64_t>(p.x0) + p.xsize <= slot.w && static_cast<uint64_t>(p.y0) + p.ysize <= slot.h,
                   "invalid patch position in reference frame");
        size_t id_count = read_num(7);
        JXLB_CHECK(id_count <= max_patches, "too many patches");
        id_count++;
        total += id_count;
        JXLB_CHECK(total <= max_patches, "too many patches");
        for (size_t i = 0; i < id_count; i++) {
          if (i == 0) {
            p.x = read_num(4);
            p.y = read_num(4);
          } else {
            const int64_t dx = UnpackSigned(read_num(6));
            JXLB_CHECK(!(dx < 0 && static_cast<uint64_t>(-dx) > last_x), "negative patch x");
            p.x = static_cast<uint32_t>(last_x + dx);
            const int64_t dy = UnpackSigned(read_num(6));
            JXLB_CHECK(!(dy < 0 && static_cast<uint64_t>(-dy) > last_y), "negative patch y");
            p.y = static_cast<uint32_t>(last_y + dy);
          }
          last_x = p.x;
          last_y = p.y;
          JXLB_CHECK(static_cast<uint64_t>(p.x) + p.xsize <= dim.xsize_padded && static_cast<uint64_t>(p.y) + p.ysize <= dim.ysize_padded,
                     "patch outside the frame");
          // one blending per colour + one per extra channel; VarDCT frames with extra channels are refused above
          p.mode = read_num(5);
          JXLB_CHECK(p.mode < 8, "invalid patch blend mode");
          JXLB_CHECK(p.mode <= 3, "unsupported: alpha patch blending");
          p.clamp = p.mode == 3 ? (read_num(9) != 0) : 0;
          v.patches.push_back(p);
        }
      }
      JXLB_CHECK(reader.FinalStateOk(), "patches: bad ANS final state");
    }
    SplineState splines;
    if (fh.flags & kFlagSplines) ReadSplines(r, dim.xsize * dim.ysize, &splines);  // (dec_frame.cc:286-292)
    if (!r.ReadBool()) {
      for (int c = 0; c < 3; c++) {
        dc_quant[c] = ReadF16(r) * (1.0f / 128.0f);
        JXLB_CHECK(dc_quant[c] >= 1e-8f, "bad DC quantisation step");
      }
    }
    const uint32_t global_scale = ReadU32(r, BitsOffset(11, 1), BitsOffset(11, 2049), BitsOffset(12, 4097), BitsOffset(16, 8193));
    const uint32_t quant_dc = ReadU32(r, Val(16), BitsOffset(5, 1), BitsOffset(8, 1), BitsOffset(16, 1));
    vf.global_scale_f = global_scale * (1.0 / 65536);
    vf.inv_global_scale = 1.0 * 65536 / global_scale;
    const float inv_quant_dc = vf.inv_global_scale / quant_dc;
    for (int c = 0; c < 3; c++) {
      vf.mul_dc[c] = inv_quant_dc * dc_quant[c];
      vf.inv_mul_dc[c] = 1.0f / vf.mul_dc[c];
    }
    ReadBlockCtx(r, &bctx);
    uint32_t color_factor = 84;
    vf.base_x = 0.0f;
    vf.base_b = 1.0f;
    int ytox_dc = 0, ytob_dc = 0;
    if (!r.ReadBool()) {  // ColorCorrelation::DecodeDC, lib/jxl/chroma_from_luma.cc:20-41
      color_factor = ReadU32(r, Val(84), Val(256), BitsOffset(8, 2), BitsOffset(16, 258));
      vf.base_x = ReadF16(r);
      JXLB_CHECK(std::fabs(vf.base_x) <= 4.0f, "base X correlation out of range");
      vf.base_b = ReadF16(r);
      JXLB_CHECK(std::fabs(vf.base_b) <= 4.0f, "base B correlation out of range");
      ytox_dc = static_cast<int>(r.Read(8)) - 128;
      ytob_dc = static_cast<int>(r.Read(8)) - 128;
    }
    vf.color_scale = 1.0f / color_factor;
    vf.cfl_dc_x = vf.base_x + ytox_dc * vf.color_scale;
    vf.cfl_dc_b = vf.base_b + ytob_dc * vf.color_scale;
    if (fh.flags & kFlagSplines) {  // the draw cache wants the base correlation (dec_frame.cc:299-305)
      InitSplineDrawCache(dim.xsize_upsampled, dim.ysize_upsampled, vf.base_x, vf.base_b, &splines);
      if (!splines.segments.empty()) {
        vf.has_splines = 1;
        PackSplineDrawCache(splines, dim.ysize_upsampled, plan, &vf.spl_seg, &vf.spl_rows, &vf.spl_idx);
      }
    }
    if (r.ReadBool()) {
      const size_t limit = std::min<size_t>(size_t{1} << 22, 1024 + dim.xsize * dim.ysize * std::max<size_t>(3, nb_extra) / 16);
      global_tree = planner.ReadTreeAndCode(r, limit);
    }
    if (nb_extra != 0) {  // the global Modular stream of the extra channels (ModularFrameDecoder::DecodeGlobalInfo)
      full.bitdepth = meta.bit_depth.bits;
      for (size_t c = 0; c < nb_extra; c++) {
        HChan ch;
        ch.w = dim.xsize;
        ch.h = dim.ysize;
        ch.plane = planner.NewPlane(ch.w, ch.h);
        full.ch.push_back(ch);
      }
      const size_t before = plan->streams.size();
      global_header = planner.PlanStream(r, bb, full, 0, dim.group_dim, global_tree);
      if (single && plan->streams.size() > before) {  // the DC group starts where the device stops reading this stream
        const ProbeResult& res = planner.DeviceResult(pc, plan->streams.size() - 1, {});
        JXLB_CHECK(res.end_bit >= bb && res.end_bit <= bb + r.Size() * 8, "global Modular stream: read past end of section");
        r.SeekTo(res.end_bit - bb);
      }
    }
    r.CheckInBounds();
    whole = r;
  }

  // ---- DC groups: quantised DC + AC metadata, two chained Modular streams decoded by one thread
  vf.dcg_index = v.upool.size();
  for (size_t g = 0; g < dim.num_dc_groups; g++) {
    BitReader r = section(1 + g, &bb);
    const size_t gx = g % dim.xsize_dc_groups, gy = g / dim.xsize_dc_groups;
    const size_t x0 = gx * dim.group_dim, y0 = gy * dim.group_dim;  // in blocks (2048 px / 8 = group_dim)
    const size_t xs = std::min(dim.group_dim, W - x0), ys = std::min(dim.group_dim, H - y0);
    const uint32_t extra_precision = r.Read(2);
    GroupHeader hdr = ReadGroupHeader(r);
    r.CheckInBounds();
    JXLB_CHECK(hdr.use_global_tree && global_tree.valid, "unsupported: VarDCT DC stream with a local MA tree");
    JXLB_CHECK(hdr.transforms.empty(), "unsupported: transforms in a VarDCT DC stream");
    JXLB_CHECK(!global_tree.lz77, "unsupported: LZ77 in VarDCT DC / AC-metadata streams");
    DevStream st{};
    st.bit_pos = bb + r.BitPos();
    st.bit_end = bb + r.Size() * 8;
    st.code = global_tree.code;
    st.stream_id = 1 + g;
    st.chan_begin = plan->chans.size();
    st.dist_multiplier = xs;
    hdr.wp.Pack(st.wp_params);
    st.num_props = global_tree.num_props;
    st.lz77_slot = 0xFFFFFFFFu;
    const int want_refs = (static_cast<int>(global_tree.num_props) - 16) / 4;
    struct Ch { uint32_t w, h, plane; bool dyn; };
    std::vector<Ch> image;
    auto add_channels = [&](const std::vector<Ch>& chs, uint32_t stream_id, uint32_t count_bits) {
      const size_t first = image.size();
      for (size_t i = 0; i < chs.size(); i++) {
        DevChannel dc{};
        dc.plane = chs[i].plane;
        dc.prop0 = i;
        dc.ref_off = plan->refs.size();
        for (int j = static_cast<int>(i) - 1; j >= 0 && static_cast<int>(dc.ref_count) < want_refs; j--) {
          if (chs[j].dyn || chs[i].dyn || chs[j].w != chs[i].w || chs[j].h != chs[i].h) continue;
          plan->refs.push_back(chs[j].plane);
          dc.ref_count++;
        }
        bool ch_wp = false;
        dc.tree_off = planner.PruneTree(*global_tree.nodes, static_cast<int32_t>(i), static_cast<int32_t>(stream_id), &ch_wp);
        dc.uses_wp = ch_wp;
        if (ch_wp) st.uses_wp = 1;
        dc.dyn = chs[i].dyn;
        if (ch_wp && !chs[i].dyn) planner.TryWpLut(&dc, dc.ref_count != 0);
        if (!ch_wp && !FramePlanner::NoNwLut()) planner.TryNwLut(&dc);
        planner.BuildCoopLut(&dc);
        if (i == 0 && first != 0) {
          dc.preamble = 1;
          dc.count_bits = count_bits;
        }
        plan->chans.push_back(dc);
        image.push_back(chs[i]);
      }
    };
    std::vector<Ch> dcs, metas;
    for (int c = 0; c < 3; c++) dcs.push_back(Ch{static_cast<uint32_t>(xs), static_cast<uint32_t>(ys), planner.NewPlane(xs, ys), false});
    add_channels(dcs, 1 + g, 0);
    const uint32_t cw = (xs + 7) >> 3, chh = (ys + 7) >> 3, cap = xs * ys;
    metas.push_back(Ch{cw, chh, planner.NewPlane(cw, chh), false});
    metas.push_back(Ch{cw, chh, planner.NewPlane(cw, chh), false});
    metas.push_back(Ch{cap, 2, planner.NewPlane(cap, 2, 4), true});
    metas.push_back(Ch{static_cast<uint32_t>(xs), static_cast<uint32_t>(ys), planner.NewPlane(xs, ys), false});
    add_channels(metas, 1 + 2 * dim.num_dc_groups + g, CeilLog2(cap));
    st.chan_end = plan->chans.size();
    st.max_w = xs;
    plan->wp_width = std::max<uint32_t>(plan->wp_width, xs);
    plan->streams.push_back(st);
    for (int i = 0; i < 7; i++) v.upool.push_back(image[i].plane);
    v.upool.push_back(extra_precision);
    if (single) {  // AC global starts where the device stops reading this chain
      const ProbeResult& res = planner.DeviceResult(pc, plan->streams.size() - 1, {});
      JXLB_CHECK(res.end_bit >= bb && res.end_bit <= bb + whole.Size() * 8, "DC group: read past end of section");
      whole.SeekTo(res.end_bit - bb);
    }
  }

  // ---- AC global (lib/jxl/dec_frame.cc:367-476)
  const SharedVarDCTTables& shared = SharedVarDCTTables::Get();
  {
    BitReader r = section(1 + dim.num_dc_groups, &bb);
    const bool all_default = r.ReadBool();
    for (int k = 0; k < kNumQuantKinds; k++) vf.table_off[k] = kSharedFlag | shared.table_off[k];
    // Raw tables (what a transcoded JPEG carries) are Modular-coded images: decoded by the device in a probe round.
    auto read_raw = [&](BitReader& rr, int kind, QuantTableSpec* q) {
      q->qtable_den = ReadF16(rr);
      JXLB_CHECK(q->qtable_den >= 1e-8f, "invalid qtable_den");
      HImage image;
      image.bitdepth = 8;
      const int tw = 8 * kQuantSizeX[kind], th = 8 * kQuantSizeY[kind];
      std::vector<uint32_t> planes;
      for (int c = 0; c < 3; c++) {
        HChan ch;
        ch.w = tw;
        ch.h = th;
        ch.plane = planner.NewPlane(tw, th);
        planes.push_back(ch.plane);
        image.ch.push_back(ch);
      }
      const size_t first = plan->streams.size();
      const GroupHeader hdr = planner.PlanStream(rr, bb, image, 1 + 3 * dim.num_dc_groups + kind, 0xFFFFFF, global_tree);
      JXLB_CHECK(hdr.transforms.empty(), "unsupported: transforms in a raw quantisation table");
      JXLB_CHECK(plan->streams.size() == first + 1, "raw quantisation table without samples");
      const ProbeResult& res = planner.DeviceResult(pc, first, planes);
      JXLB_CHECK(res.end_bit >= bb && res.end_bit <= bb + rr.Size() * 8, "raw quantisation table: read past end of section");
      rr.SeekTo(res.end_bit - bb);
      JXLB_CHECK(res.samples.size() == static_cast<size_t>(3) * tw * th, "internal: probe returned the wrong planes");
      q->qtable = res.samples;
      for (int32_t t : q->qtable) JXLB_CHECK(t > 0, "invalid raw quantisation table");
    };
    if (!all_default) {
      for (int k = 0; k < kNumQuantKinds; k++) {
        QuantTableSpec spec;
        ReadQuantTableSpec(r, k, &spec, read_raw);
        if (spec.mode == kQLib) continue;
        vf.table_off[k] = v.fpool.size();
        std::vector<float> tab = BuildDequantTable(spec, k);
        v.fpool.insert(v.fpool.end(), tab.begin(), tab.end());
      }
    }
    vf.num_histograms = 1 + r.Read(CeilLog2(dim.num_groups));
    vf.order_index = v.upool.size();
    v.upool.resize(v.upool.size() + num_passes * 39, 0);
    const uint32_t num_ac_ctx = bctx.NumACContexts();
    for (size_t p = 0; p < num_passes; p++) {
      const uint32_t used_orders = ReadU32(r, Val(0x5F), Val(0x13), Val(0), Bits(13));
      EntropyCode perm_code;
      std::unique_ptr<SymbolReader> reader;
      if (used_orders != 0) {
        ReadEntropyCode(r, 8, &perm_code);
        reader.reset(new SymbolReader(&perm_code, r));
      }
      std::vector<uint32_t> perm;
      for (uint32_t ord = 0; ord < kNumOrders; ord++) {
        const StrategyInfo si = GetStrategyInfo(kOrderFirstStrategy[ord]);
        const size_t llf = static_cast<size_t>(si.cx) * si.cy, size = 64 * llf;
        uint32_t* index = &v.upool[vf.order_index + p * 39 + 3 * ord];
        if (!(used_orders & (1u << ord))) {
          for (int c = 0; c < 3; c++) index[c] = kSharedFlag | shared.order_off[ord];
          continue;
        }
        const uint16_t* natural = shared.opool.data() + shared.order_off[ord];
        for (int c = 0; c < 3; c++) {
          perm.resize(size);
          ReadPermutation(r, *reader, llf, size, perm.data());
          v.upool[vf.order_index + p * 39 + 3 * ord + c] = v.opool.size();
          for (size_t k = 0; k < size; k++) v.opool.push_back(natural[perm[k]]);
        }
      }
      if (used_orders) JXLB_CHECK(reader->FinalStateOk(), "coefficient orders: bad ANS final state");
      EntropyCode code;
      ReadEntropyCode(r, static_cast<size_t>(vf.num_histograms) * num_ac_ctx, &code);
      JXLB_CHECK(!code.lz77_enabled, "unsupported: LZ77 in AC coefficient streams");
      vf.ac_code[p] = planner.AddCode(code);
      vf.ctx_map_off[p] = v.cpool.size();
      v.cpool.insert(v.cpool.end(), code.ctx_map.begin(), code.ctx_map.end());
      v.cpool.resize(v.cpool.size() + 16 + 3, 0);  // lib/jxl/dec_frame.cc:406-408 padding
      v.cpool.resize(v.cpool.size() & ~size_t{3});
    }
    r.CheckInBounds();
    whole = r;
  }

  // ---- AC groups: one stream per (pass, group)
  for (size_t p = 0; p < num_passes; p++) {
    for (size_t g = 0; g < dim.num_groups; g++) {
      const size_t idx = single ? 0 : 2 + dim.num_dc_groups + p * dim.num_groups + g;
      DevAcStream s{};
      s.bit_pos = (base + toc.offsets[idx]) * 8;
      s.bit_end = s.bit_pos + static_cast<uint64_t>(toc.logical_size[idx]) * 8;
      if (single) {
        s.bit_end = (base + toc.offsets[0]) * 8 + static_cast<uint64_t>(toc.logical_size[0]) * 8;
        s.bit_pos = (base + toc.offsets[0]) * 8 + whole.BitPos();
      }
      s.frame = 0;
      s.group = g;
      s.pass = p;
      s.tok_cap = std::min<uint64_t>(3 * 65536 * 2, 2 * static_cast<uint64_t>(toc.logical_size[idx]) + 512);
      s.tok_off = v.tok_size;
      v.tok_size += s.tok_cap;
      // the extra channels of this group (ModularStreamId::ModularAC): channels larger than a group, shift 0 .. 2
      if (nb_extra != 0) {
        HImage gi;
        gi.bitdepth = full.bitdepth;
        size_t c = full.nb_meta;
        for (; c < full.ch.size(); c++)
          if (static_cast<size_t>(full.ch[c].w) > dim.group_dim || static_cast<size_t>(full.ch[c].h) > dim.group_dim) break;
        struct Dest { size_t c; int x, y; };
        std::vector<Dest> dests;
        const size_t gx = g % dim.xsize_groups, gy = g / dim.xsize_groups;
        const size_t x0 = gx * dim.group_dim, y0 = gy * dim.group_dim;
        for (; c < full.ch.size(); c++) {
          const HChan& fc = full.ch[c];
          const int shift = std::min(fc.hshift, fc.vshift);
          JXLB_CHECK(shift < 3, "unsupported: DC-group Modular stream in a VarDCT frame (squeezed extra channel)");
          const int rx = static_cast<int>(x0 >> fc.hshift), ry = static_cast<int>(y0 >> fc.vshift);
          int rw = static_cast<int>(dim.group_dim >> fc.hshift), rh = static_cast<int>(dim.group_dim >> fc.vshift);
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
          dests.push_back(Dest{c, rx, ry});
        }
        if (!gi.ch.empty()) {
          s.chain_slot = ++plan->chain_slots;
          planner.PlanChainedStream(gi, 1 + 3 * dim.num_dc_groups + kNumQuantKinds + dim.num_groups * p + g, global_tree,
                                    s.bit_end, s.chain_slot);
          DevProgram prog;  // (no transforms in such a stream: the program is the copy into the frame's planes)
          prog.op_begin = plan->ops.size();
          for (size_t i = 0; i < dests.size(); i++) {
            DevOp cp{};
            cp.kind = kOpCopy;
            cp.a = gi.ch[i].plane;
            cp.b = full.ch[dests[i].c].plane;
            cp.p0 = dests[i].x;
            cp.p1 = dests[i].y;
            plan->ops.push_back(cp);
          }
          prog.op_end = plan->ops.size();
          plan->group_programs.push_back(prog);
        }
      }
      v.ac_streams.push_back(s);
    }
  }
  vf.alpha_plane = kNoPlane;
  if (nb_extra != 0) {
    // global inverse transforms of the extra channels, then the alpha plane for the output
    std::vector<DevOp> global_ops;
    planner.EmitInverse(full, global_header.wp, &global_ops);
    for (const DevOp& op : global_ops) {
      DevProgram lvl;
      lvl.op_begin = plan->ops.size();
      plan->ops.push_back(op);
      lvl.op_end = plan->ops.size();
      plan->frame_levels.push_back(lvl);
    }
    JXLB_CHECK(full.ch.size() == nb_extra, "modular: unexpected channel count after transforms");
    const int alpha = meta.AlphaIndex();
    if (alpha >= 0) {
      const HChan& ch = full.ch[alpha];
      const BitDepth& bd = meta.extra[alpha].bit_depth;
      JXLB_CHECK(ch.w == static_cast<int>(dim.xsize) && ch.h == static_cast<int>(dim.ysize), "unsupported: subsampled channel");
      JXLB_CHECK(!bd.floating_point && bd.bits < 23, "unsupported: float or wide alpha in a VarDCT frame");
      vf.alpha_plane = ch.plane;
      vf.alpha_factor = static_cast<float>(1.0 / ((1u << bd.bits) - 1));
    }
  }

  // ---- frame descriptor
  vf.xsize = dim.xsize;
  vf.ysize = dim.ysize;
  vf.xblocks = W;
  vf.yblocks = H;
  vf.xgroups = dim.xsize_groups;
  vf.ygroups = dim.ysize_groups;
  vf.xdcgroups = dim.xsize_dc_groups;
  vf.ydcgroups = dim.ysize_dc_groups;
  vf.cmw = DivCeil(W, size_t{8});
  vf.cmh = DivCeil(H, size_t{8});
  vf.num_passes = num_passes;
  for (size_t p = 0; p < num_passes; p++) vf.pass_shift[p] = fh.passes.shift[p];
  vf.skip_dc_smoothing = ((fh.flags & kFlagSkipAdaptiveDCSmoothing) != 0 || W <= 2 || H <= 2) ? 1 : 0;
  vf.gab = fh.lf.gab;
  vf.epf_iters = fh.lf.epf_iters;
  vf.patch_begin = 0;
  vf.patch_count = v.patches.size();
  uint64_t f = v.farena_size;  // (reference frames of this codestream come first)
  for (int c = 0; c < 3; c++) {
    vf.dc[c] = f;
    f += nb;
  }
  for (int c = 0; c < 3; c++) {
    vf.dc_final[c] = vf.skip_dc_smoothing ? vf.dc[c] : f;
    if (!vf.skip_dc_smoothing) f += nb;
  }
  vf.inv_sigma = f;
  f += nb;
  f = (f + 3) & ~uint64_t{3};
  vf.upsampling = fh.upsampling;
  vf.up_xsize = dim.xsize_upsampled;
  vf.up_ysize = dim.ysize_upsampled;
  if (fh.upsampling > 1) {
    // the upsampled planes of this frame (per frame, not per wave slot: such frames are rare) and the 5 x 5 kernels
    // expanded from the upper triangle of weights (stage_upsampling.cc:33-47)
    vf.up_stride = static_cast<uint32_t>((dim.xsize * fh.upsampling + 3) & ~size_t{3});
    for (int c = 0; c < 3; c++) {
      vf.up_pix[c] = f;
      f += static_cast<uint64_t>(vf.up_stride) * dim.ysize * fh.upsampling;
    }
    const uint32_t U = fh.upsampling, half = U / 2;
    const float* weights = U == 2 ? ((meta.custom_weights_mask & 1) ? meta.up2.data() : kDefaultUpsampling2)
                                  : (U == 4 ? ((meta.custom_weights_mask & 2) ? meta.up4.data() : kDefaultUpsampling4)
                                            : ((meta.custom_weights_mask & 4) ? meta.up8.data() : kDefaultUpsampling8));
    vf.up_kernel = v.fpool.size();
    v.fpool.resize(v.fpool.size() + 400, 0.0f);
    float* kernel = v.fpool.data() + vf.up_kernel;
    for (uint32_t i = 0; i < 5 * half; i++)
      for (uint32_t j = 0; j < 5 * half; j++) {
        const uint32_t y = std::min(i, j), x = std::max(i, j);
        kernel[(((j / 5) * 4 + i / 5) * 5 + j % 5) * 5 + i % 5] = weights[5 * half * y - y * (y - 1) / 2 + x - y];
      }
  }
  v.farena_size = f;
  v.pix_plane = static_cast<uint64_t>(W * 8) * (H * 8);
  uint64_t b = 0;
  vf.acs = b; b += nb;
  vf.qdc = b; b += nb;
  vf.sharp = b; b += nb;
  b = (b + 1) & ~uint64_t{1};
  vf.rawq = b; b += 2 * nb;
  vf.ytox = b; b += vf.cmw * vf.cmh;
  vf.ytob = b; b += vf.cmw * vf.cmh;
  v.barena_size = (b + 15) & ~uint64_t{15};
  vf.tok_start = 0;
  vf.tok_count = num_passes * 3 * nb;
  vf.blist = 2 * num_passes * 3 * nb;
  vf.blist_count = vf.blist + 2 * nb;
  v.uarena_size = vf.blist_count + dim.num_groups;
  // quantiser
  vf.x_dm = std::pow(1 / (1.25f), fh.x_qm_scale - 2.0f);  // lib/jxl/dec_cache.h:161-162
  vf.b_dm = std::pow(1 / (1.25f), fh.b_qm_scale - 2.0f);
  for (int i = 0; i < 4; i++) vf.biases[i] = meta.quant_biases[i];
  // block context map -> upool: dc thresholds (3 lists), qf thresholds, then the map bytes packed 4 per word
  vf.bctx_off = v.upool.size();
  vf.num_ctxs = bctx.num_ctxs;
  vf.num_dc_ctxs = bctx.num_dc_ctxs;
  vf.num_qf_thr = bctx.qf_thr.size();
  for (int j = 0; j < 3; j++) {
    vf.num_dc_thr[j] = bctx.dc_thr[j].size();
    for (int32_t t : bctx.dc_thr[j]) v.upool.push_back(static_cast<uint32_t>(t));
  }
  for (uint32_t t : bctx.qf_thr) v.upool.push_back(t);
  for (size_t i = 0; i < bctx.ctx_map.size(); i += 4) {
    uint32_t wv = 0;
    for (size_t k = 0; k < 4 && i + k < bctx.ctx_map.size(); k++) wv |= static_cast<uint32_t>(bctx.ctx_map[i + k]) << (8 * k);
    v.upool.push_back(wv);
  }
  // loop filter (lib/jxl/render_pipeline/stage_gaborish.cc:22-48, stage_epf.cc, lib/jxl/epf.cc)
  const LoopFilter& lf = fh.lf;
  const float gw[3][2] = {{lf.gab_x_weight1, lf.gab_x_weight2}, {lf.gab_y_weight1, lf.gab_y_weight2},
                          {lf.gab_b_weight1, lf.gab_b_weight2}};
  for (int c = 0; c < 3; c++) {
    float w[3] = {1.0f, gw[c][0], gw[c][1]};
    const float div = w[0] + 4 * (w[1] + w[2]);
    const float mul = 1.0f / div;
    for (int i = 0; i < 3; i++) vf.gab_w[c][i] = w[i] * mul;
  }
  vf.epf_sigma_scale[0] = lf.epf_pass0_sigma_scale * 1.65;
  vf.epf_sigma_scale[1] = 1.65f;
  vf.epf_sigma_scale[2] = lf.epf_pass2_sigma_scale * 1.65;
  vf.epf_border_sad_mul = lf.epf_border_sad_mul;
  for (int c = 0; c < 3; c++) vf.epf_channel_scale[c] = lf.epf_channel_scale[c];
  vf.epf_quant_mul = lf.epf_quant_mul;
  for (int i = 0; i < 8; i++) vf.epf_sharp_lut[i] = lf.epf_sharp_lut[i];
  // colour + output
  vf.color_transform = fh.color_transform;
  if (fh.color_transform == kCTXYB) FillOutputColor(meta, &vf);
  vf.out_channels = fmt.num_channels;
  vf.out_type = fmt.data_type;
  vf.out_big_endian = fmt.endianness == 2;
  vf.orient = OrientBits(meta.orientation, fmt);
  vf.out_stride = OutputStride((vf.orient & 4) ? dim.ysize_upsampled : dim.xsize_upsampled, fmt);
  if (fmt.num_channels < 3) JXLB_CHECK(meta.color.IsGray(), "grey output requested for a colour image");
}

}  // namespace jxlb

#endif  // JXLB_VARDCT_FRAME_H_
