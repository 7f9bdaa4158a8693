// TEST INFRASTRUCTURE: runs the __host__ __device__ bodies of the jxl_b200 kernels on
// the CPU so that the planner and the per-stream logic can be checked on a box
// without a GPU. The product library never does this (it fails without CUDA);
// nothing outside tests/ builds or loads this file.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../jpegxl-rs_b200/csrc/host/jxlb_batch.h"
#include "../../jpegxl-rs_b200/csrc/kernels/jxlb_modular_coop_dev.h"
#include "../../jpegxl-rs_b200/csrc/kernels/jxlb_finish_dev.h"
#include "../../jpegxl-rs_b200/csrc/kernels/jxlb_vardct_dev.h"

using namespace jxlb;

static uint64_t g_last_num_coop[2];  // streams of the last plan: one per warp (k_modular_decode_coop), total
static uint64_t g_last_plan_stats[3];  // channels, of which weighted-predictor LUT, of which (y, N, W) table

extern "C" {

// Streams of the last jxlb_emul_decode plan: decoded one per warp, total.
void jxlb_emul_last_coop_streams(uint64_t* out2) {
  out2[0] = g_last_num_coop[0];
  out2[1] = g_last_num_coop[1];
}

// Channel counts of the last jxlb_emul_decode plan: total, weighted-predictor LUT path, (y, N, W) table path.
void jxlb_emul_last_plan_stats(uint64_t* out3) {
  for (int i = 0; i < 3; i++) out3[i] = g_last_plan_stats[i];
}

// Decodes n files; returns 0 or a negative error. out must hold out_size bytes
// laid out like the batch output buffer (frames 256-byte aligned).
int jxlb_emul_decode(const uint8_t* const* files, const size_t* sizes, size_t n, uint32_t num_channels,
                     uint32_t data_type, uint32_t endianness, size_t align, uint8_t* out, size_t out_cap,
                     uint64_t* frame_offsets, char* err, size_t errlen) {
  try {
    const bool force_wide = (endianness & 0x100) != 0;  // test hook: run the 64-bit predictor path
    const bool generic_only = (endianness & 0x200) != 0;  // test hook: generic varblock path for every block
    const bool undo_orientation = (endianness & 0x400) != 0;  // libjxl's default output; the tests' default keeps the coded image
    endianness &= 0xFF;
    (void)generic_only;
    PixelFormat fmt;
    fmt.num_channels = num_channels;
    fmt.data_type = data_type;
    fmt.endianness = endianness;
    fmt.align = align;
    fmt.keep_orientation = !undo_orientation;
    // the Modular decode "kernel": every stream of `b`, one lane at a time
    // phase 0: the streams whose position the host knows; phase 1: the ones chained behind AC coefficient streams
    auto run_modular = [&](const BatchPlan& b, std::vector<int32_t>& arena, DevPools* pools_out, std::vector<uint64_t>* end_bits,
                           int phase = 0, const uint64_t* chain_pos = nullptr) {
      const size_t num_warps = b.warp_chans.size() + 1;
      std::vector<int32_t> wp(num_warps * 10 * (b.wp_width + 2) * 32 + 16, 0);
      std::vector<int32_t> ring(num_warps * 3 * b.wp_width * 32 + 16, 0);
      std::vector<int32_t> props(kDevMaxProps * 32, 0);
      uint32_t divlut[64];
      for (uint32_t i = 0; i < 64; i++) divlut[i] = (1u << 24) / (i + 1);
      std::vector<uint32_t> lz(static_cast<size_t>(b.lz77_slots) << 20, 0);
      DevPools P{};
      P.words = reinterpret_cast<const uint32_t*>(b.bytes.data());
      P.alias = b.alias.data();
      P.prefix = b.prefix.data();
      P.cfg = b.cfg.data();
      P.tree = b.tree.data();
      P.chans = b.chans.data();
      P.streams = b.streams.data();
      P.planes = b.planes.data();
      P.refs = b.refs.data();
      P.lut = b.lut.data();
      P.codes = b.codes.data();
      P.arena = arena.data();
      P.wp_scratch = wp.data();
      P.ring = ring.data();
      P.wp_width = b.wp_width;
      P.lz77 = lz.data();
      P.num_streams = b.streams.size();
      const uint32_t first = phase == 0 ? 0 : b.num_early, last = phase == 0 ? b.num_early : static_cast<uint32_t>(b.streams.size());
      const uint32_t ncoop = phase == 0 ? b.num_coop : b.late_coop, warp0 = phase == 0 ? 0 : b.early_warps;
      P.coop0 = first;
      P.stream0 = first + ncoop;
      P.chain_pos = chain_pos;
      P.spl_seg = b.spl_seg.data();
      P.spl_idx = b.spl_idx.data();
      P.warp_chans = b.warp_chans.data();
      P.warp_dims_off = b.warp_dims_off.data();
      P.warp_dims = b.warp_dims.data();
      if (end_bits && phase == 0) end_bits->assign(b.streams.size(), 0);
      // k_modular_decode_coop: one warp per stream; the host runs its single "lane" (jxlb_modular_coop_dev.h)
      std::vector<int32_t> coop_rows(7 * b.wp_width + 10, 0);
      for (uint32_t s = first; s < first + ncoop; s++) {
        uint64_t end = 0;
        const uint32_t st = force_wide || !b.narrow
                                ? DevDecodeModularStreamCoop<int64_t>(P, s, coop_rows.data(), coop_rows.data() + 2 * b.wp_width, b.wp_width, divlut, &end)
                                : DevDecodeModularStreamCoop<int32_t>(P, s, coop_rows.data(), coop_rows.data() + 2 * b.wp_width, b.wp_width, divlut, &end);
        if (st != 0) throw Error("stream " + std::to_string(s) + " (one per warp) failed with status " + std::to_string(st));
        if (end_bits) (*end_bits)[s] = end;
      }
      for (uint32_t s = first + ncoop; s < last; s++) {
        // same addressing as the kernel: warp = (s - stream0) / 32 (+ the early bundles in the late launch), lane = (s - stream0) % 32
        const uint32_t warp = warp0 + (s - first - ncoop) / 32, lane = (s - first - ncoop) % 32;
        DevLaneMem m;
        m.props = props.data() + lane;
        m.props_stride = 32;
        m.divlut = divlut;
        m.ring_w = b.wp_width;
        m.lane_stride = 32;
        m.ring = ring.data() + static_cast<size_t>(warp) * 3 * b.wp_width * 32 + lane;
        m.wp = wp.data() + static_cast<size_t>(warp) * 10 * (b.wp_width + 2) * 32 + lane;
        const uint32_t* dims = b.warp_dims.data() + b.warp_dims_off[warp];
        uint64_t end = 0;
        uint32_t st = force_wide || !b.narrow ? DevDecodeModularStream<int64_t>(P, s, m, dims, b.warp_chans[warp], true, &end)
                                              : DevDecodeModularStream<int32_t>(P, s, m, dims, b.warp_chans[warp], true, &end);
        if (st != 0) throw Error("stream " + std::to_string(s) + " failed with status " + std::to_string(st));
        if (end_bits) (*end_bits)[s] = end;
      }
      P.wp_scratch = nullptr;  // scratch dies with this call
      P.ring = nullptr;
      P.lz77 = nullptr;
      if (pools_out) *pools_out = P;
    };
    // probe rounds, as JxlB200DecoderSetInputBatch runs them on the device
    const ProbeFn probe = [&](const BatchPlan& pb, std::vector<uint64_t>* end_bits, std::vector<int32_t>* arena) {
      arena->assign(pb.arena_size + 16, 0);
      run_modular(pb, *arena, nullptr, end_bits);
    };
    BatchPlan b;
    PlanBatch(files, sizes, n, fmt, 2, &b, probe);
    g_last_num_coop[0] = b.num_coop;
    g_last_num_coop[1] = b.streams.size();
    g_last_plan_stats[0] = b.chans.size();
    g_last_plan_stats[1] = g_last_plan_stats[2] = 0;
    for (const DevChannel& c : b.chans) {
      g_last_plan_stats[1] += c.wp_lut ? 1 : 0;
      g_last_plan_stats[2] += c.nw_lut ? 1 : 0;
    }
    if (b.out_size > out_cap) throw Error("output buffer too small");
    std::vector<int32_t> arena(b.arena_size + 16, 0);
    DevPools P{};
    run_modular(b, arena, &P, nullptr);
    const uint32_t nt = 4;  // emulate a few cooperating workers
    for (const DevProgram& pr : b.group_programs)
      for (uint32_t o = pr.op_begin; o < pr.op_end; o++)
        for (uint32_t t = 0; t < nt; t++) DevRunOp(P, b.ops[o], t, nt);
    for (const auto& lvl : b.levels)
      for (const DevProgram& pr : lvl)
        for (uint32_t o = pr.op_begin; o < pr.op_end; o++)
          for (uint32_t t = 0; t < nt; t++) DevRunOp(P, b.ops[o], t, nt);
    for (size_t f = 0; f < b.frames.size(); f++) {
      const DevFrameOut& fo = b.frames[f];
      frame_offsets[f] = fo.out_off;
      if (fo.vardct) continue;
      for (uint32_t y = 0; y < fo.ysize; y++)
        for (uint32_t x = 0; x < fo.xsize; x++) DevWritePixel(P, fo, out, x, y);
    }
    if (!b.vframes.empty()) {
      // ---- VarDCT frames: the same device functions, one "thread" at a time
      const SharedVarDCTTables& sh = SharedVarDCTTables::Get();
      std::vector<float> farena(b.farena_size + 16, 0.0f);
      std::vector<uint8_t> barena(b.barena_size + 16, 0);
      std::vector<uint32_t> uarena(b.uarena_size + 16, 0);
      std::vector<uint32_t> tokens(b.tok_size + 16, 0);
      std::vector<uint32_t> ac_status(b.ac_streams.size() + 1, 0), ac_used(b.ac_streams.size() + 1, 0);
      size_t num_dcg = 0;
      for (const DevVFrame& vf : b.vframes) num_dcg += vf.xdcgroups * vf.ydcgroups;
      std::vector<uint32_t> dc_status(num_dcg + 1, 0);
      DevVPools V{};
      V.frames = b.vframes.data();
      V.streams = b.ac_streams.data();
      V.num_streams = b.ac_streams.size();
      V.fpool = b.fpool.data();
      V.opool = b.opool.data();
      V.cpool = b.cpool.data();
      V.upool = b.upool.data();
      V.farena = farena.data();
      V.barena = barena.data();
      V.uarena = uarena.data();
      V.tokens = tokens.data();
      V.ac_status = ac_status.data();
      V.ac_used = ac_used.data();
      V.dc_status = dc_status.data();
      V.wc_off = sh.wc_off;
      V.llf_off = sh.llf_off;
      V.afv_off = sh.afv_off;
      V.sinfo_off = sh.sinfo_off;
      V.ctxtab_off = sh.ctxtab_off;
      V.out = out;
      V.ref_frames = b.ref_frames.data();
      V.num_ref_frames = b.ref_frames.size();
      V.patches = b.patches.data();
      std::vector<uint64_t> chain_pos(b.chain_slots + 1, 0);
      V.chain_pos = chain_pos.data();
      V.arena = arena.data();
      V.planes = b.planes.data();
      V.spl_seg = b.spl_seg.data();
      V.spl_idx = b.spl_idx.data();
      for (const DevRefFrame& rf : b.ref_frames)  // k_ref_frames
        for (uint32_t i = 0; i < rf.w * rf.h; i++) DevRefFrameSample(P, V, rf, i);
      uint32_t dcg = 0;
      std::vector<uint32_t> occ(kDcOccWords);  // stands in for the kernel's shared memory
      for (uint32_t f = 0; f < b.vframes.size(); f++) {
        const DevVFrame& vf = b.vframes[f];
        std::vector<uint32_t> stage(kDcStageEntries);
        for (uint32_t g = 0; g < vf.xdcgroups * vf.ydcgroups; g++, dcg++)
          DevDcGroupFinish<0>(P, V, f, g, 0, 1, dcg, occ.data(), stage.data());
        if (!vf.skip_dc_smoothing)
          for (uint32_t y = 0; y < vf.yblocks; y++)
            for (uint32_t x = 0; x < vf.xblocks; x++) DevDcSmoothBlock(V, vf, x, y);
      }
      for (uint32_t i = 0; i < dcg; i++)
        if (dc_status[i]) throw Error("DC group " + std::to_string(i) + " failed with status " + std::to_string(dc_status[i]));
      uint16_t ctxtab[128];
      for (int i = 0; i < 128; i++) ctxtab[i] = static_cast<uint16_t>(b.upool[sh.ctxtab_off + i]);
      bool ac_plain = true;  // as JxlB200DecoderSetInputBatch decides
      for (const DevVFrame& vf : b.vframes)
        for (uint32_t p = 0; p < vf.num_passes; p++)
          if (b.codes[vf.ac_code[p]].use_prefix || b.codes[vf.ac_code[p]].lz77_enabled) ac_plain = false;
      for (const DevVFrame& vf : b.vframes)  // k_block_lists
        for (uint32_t g = 0; g < vf.xgroups * vf.ygroups; g++) DevBuildBlockList(V, vf, g);
      for (int attempt = 0; attempt < 2; attempt++) {
        bool overflow = false;
        V.streams = b.ac_streams.data();
        for (uint32_t s = 0; s < b.ac_streams.size(); s++) {
          uint8_t colnz[96] = {0};
          DevAcLaneMem m;
          m.colnz = colnz;
          m.stride = 1;
          m.freq_ctx = ctxtab;
          m.nnz_ctx = ctxtab + 64;
          const uint32_t st = ac_plain ? DevDecodeAcStream<true>(P, V, s, m, true) : DevDecodeAcStream<false>(P, V, s, m, true);
          if (st == kVTokenOverflow && attempt == 0) {
            overflow = true;
          } else if (st != 0) {
            throw Error("AC stream " + std::to_string(s) + " failed with status " + std::to_string(st));
          }
        }
        if (!overflow) break;
        // the product does the same in JxlB200DecoderWait: grow the token arena and decode again
        if (!GrowTokenCapacity(&b, ac_used.data())) throw Error("token overflow without growth");
        tokens.assign(b.tok_size + 16, 0);
        V.tokens = tokens.data();
      }
      // the extra channels of VarDCT frames: the Modular streams chained behind the AC coefficients (second launch),
      // their copies into the frame's planes, the global inverse transforms
      if (b.num_early != b.streams.size()) run_modular(b, arena, nullptr, nullptr, 1, chain_pos.data());
      for (const DevProgram& pr : b.late_group_programs)
        for (uint32_t o = pr.op_begin; o < pr.op_end; o++)
          for (uint32_t t = 0; t < nt; t++) DevRunOp(P, b.ops[o], t, nt);
      for (const auto& lvl : b.late_levels)
        for (const DevProgram& pr : lvl)
          for (uint32_t o = pr.op_begin; o < pr.op_end; o++)
            for (uint32_t t = 0; t < nt; t++) DevRunOp(P, b.ops[o], t, nt);
      for (uint32_t s = 0; s < 0 * b.ac_streams.size(); s++) {
        uint8_t colnz[96] = {0};
        DevAcLaneMem m;
        m.colnz = colnz;
        m.stride = 1;
        m.freq_ctx = ctxtab;
        m.nnz_ctx = ctxtab + 64;
        const uint32_t st = DevDecodeAcStream(P, V, s, m, true);
        if (st != 0) throw Error("AC stream " + std::to_string(s) + " failed with status " + std::to_string(st) + " (used " +
                                 std::to_string(ac_used[s]) + " of " + std::to_string(b.ac_streams[s].tok_cap) + " tokens)");
      }
      std::vector<float> buf(4 * 65536 + 64);
      for (uint32_t f = 0; f < b.vframes.size(); f++) {
        const DevVFrame& vf = b.vframes[f];
        for (uint32_t by = 0; by < vf.yblocks; by++)
          for (uint32_t bx = 0; bx < vf.xblocks; bx++) {
            const uint8_t a = barena[vf.acs + static_cast<size_t>(by) * vf.xblocks + bx];
            if (!(a & 1)) continue;
            const StrategyInfo si = GetStrategyInfo(a >> 1);
            if (si.plain_dct && si.cx * si.cy <= 16 && !generic_only) {
              DevVarblockFast<0, 32>(V, vf, bx, by, a >> 1, buf.data(), 0, 1);  // what k_dequant_idct runs for these
            } else if (si.plain_dct && si.cx * si.cy <= 64 && !generic_only) {
              DevVarblockFast<0, 64>(V, vf, bx, by, a >> 1, buf.data(), 0, 1);  // k_idct_mid
            } else if (!si.plain_dct && si.cx * si.cy == 1 && vf.num_passes == 1 && !generic_only) {
              // the special 8x8 transforms of single-pass frames (k_dequant_idct: four lanes per varblock)
              const DevBlockMeta meta = DevLoadBlockMeta(V, vf, static_cast<size_t>(by) * vf.xblocks + bx);
              DevVarblockSpecial<0>(V, vf, bx, by, a >> 1, buf.data(), 0, 1, meta, true);
            } else {
              DevVarblock<0>(V, vf, bx, by, a >> 1, buf.data(), 0, 1);
            }
          }
        if (DevRenderFused(vf) && std::getenv("JXLB_EMUL_UNFUSED") == nullptr) {  // k_render_fused, tile by tile
          // (the kernel takes the stride of the batch's largest halo; here the frame's own, or the large one on request)
          const uint32_t halo = DevRenderHalo(vf.gab, vf.epf_iters);
          const bool large = halo > 4 || std::getenv("JXLB_EMUL_LARGE_STRIDE") != nullptr;
          const uint32_t cap = (large ? kRtStrideLarge : kRtStrideSmall) * (kRtH + 2 * (large ? kRtMaxHalo : halo));
          std::vector<float> sm(6 * static_cast<size_t>(cap));
          const int H = static_cast<int>(halo);
          const int xsize = static_cast<int>(vf.xsize), ysize = static_cast<int>(vf.ysize);
          for (int ty0 = 0; ty0 < ysize; ty0 += kRtH)
            for (int tx0 = 0; tx0 < xsize; tx0 += kRtW) {
              std::fill(sm.begin(), sm.end(), std::numeric_limits<float>::quiet_NaN());  // a read of an unset cell shows
              const bool interior = tx0 - H >= 0 && ty0 - H >= 0 && tx0 + kRtW + H <= xsize && ty0 + kRtH + H <= ysize;
              if (large) {
                if (interior) DevRenderTile<0, true, kRtStrideLarge>(V, vf, tx0, ty0, 0, 1, sm.data(), cap);
                else DevRenderTile<0, false, kRtStrideLarge>(V, vf, tx0, ty0, 0, 1, sm.data(), cap);
              } else {
                if (interior) DevRenderTile<0, true, kRtStrideSmall>(V, vf, tx0, ty0, 0, 1, sm.data(), cap);
                else DevRenderTile<0, false, kRtStrideSmall>(V, vf, tx0, ty0, 0, 1, sm.data(), cap);
              }
            }
          continue;
        }
        uint32_t set = 0;
        if (vf.gab) {
          for (uint32_t c = 0; c < 3; c++)
            for (uint32_t y = 0; y < vf.ysize; y++)
              for (uint32_t x = 0; x < vf.xsize; x++) {
                const bool in = x >= 1 && y >= 1 && x + 1 < vf.xsize && y + 1 < vf.ysize;
                if (in) DevGaborishPixel<true>(V, vf, set, set ^ 1, c, x, y);
                else DevGaborishPixel<false>(V, vf, set, set ^ 1, c, x, y);
              }
          set ^= 1;
        }
        for (uint32_t stage = 0; stage < 3; stage++) {
          if (vf.epf_iters == 0 || (stage == 0 && vf.epf_iters < 3) || (stage == 2 && vf.epf_iters < 2)) continue;
          for (uint32_t y = 0; y < vf.ysize; y++)
            for (uint32_t x = 0; x < vf.xsize; x++) {
              const bool in = x >= 3 && y >= 3 && x + 3 < vf.xsize && y + 3 < vf.ysize;
              if (in) DevEpfPixel<true>(V, vf, stage, set, set ^ 1, x, y);
              else DevEpfPixel<false>(V, vf, stage, set, set ^ 1, x, y);
            }
          set ^= 1;
        }
        for (uint32_t k = 0; k < vf.patch_count; k++) {  // k_patches: in order, one patch after the other
          const DevPatch& pt = b.patches[vf.patch_begin + k];
          for (uint32_t iy = 0; iy < pt.ysize; iy++)
            for (uint32_t ix = 0; ix < pt.xsize; ix++) DevPatchPixel(V, vf, pt, set, ix, iy);
        }
        if (vf.upsampling > 1) {  // k_upsample, then k_color_write at the image's size
          for (uint32_t y = 0; y < vf.ysize * vf.upsampling; y++)
            for (uint32_t x = 0; x < vf.xsize * vf.upsampling; x++) DevUpsamplePixel(V, vf, set, x, y);
          for (uint32_t y = 0; y < vf.up_ysize; y++)
            for (uint32_t x = 0; x < vf.up_xsize; x++) DevColorPixel(V, vf, 0, x, y);
          continue;
        }
        const bool x4 = vf.out_type == 2 && vf.out_channels == 3 && vf.out_stride % 4 == 0 && vf.orient == 0 && !vf.has_splines;
        for (uint32_t y = 0; y < vf.ysize; y++)
          for (uint32_t x = 0; x < vf.xsize; x++) {
            if (x4 && x % 4 == 0 && x + 4 <= vf.xsize) {  // the kernel's vector path
              DevColorPixelsRgb8x4(V, vf, set, x, y);
              x += 3;
            } else {
              DevColorPixel(V, vf, set, x, y);
            }
          }
      }
    }
    return 0;
  } catch (const std::exception& e) {
    std::snprintf(err, errlen, "%s", e.what());
    return -1;
  }
}
}

// ---------------------------------------------------------------- encoder
#include "../../jpegxl-rs_b200/csrc/host/jxlb_enc_host.h"
#include "../../jpegxl-rs_b200/csrc/kernels/jxlb_enc_dev.h"
#include <cmath>
#include <random>
#include "../../jpegxl-rs_b200/csrc/host/jxlb_encl_host.h"
#include "../../jpegxl-rs_b200/csrc/kernels/jxlb_encl_dev.h"

extern "C" {

// Encodes one RGB8 image with the kernels' device functions run on the CPU; returns the size or -1.
long jxlb_emul_encode(const uint8_t* rgb, uint32_t xsize, uint32_t ysize, float distance, int strategy_mode, int gab,
                      uint32_t epf_iters, int dc_smoothing, uint8_t* out, size_t out_cap, char* err, size_t errlen) {
  try {
    EncParams p;
    p.distance = distance;
    p.strategy_mode = strategy_mode;
    p.gab = gab != 0;
    p.epf_iters = epf_iters;
    p.dc_smoothing = (dc_smoothing & 1) != 0;
    p.alpha = (dc_smoothing & 2) != 0;  // test hook: `rgb` holds interleaved RGBA8
    JXLB_CHECK(strategy_mode == 0 || strategy_mode == 2, "unsupported strategy mode");
    const SharedVarDCTTables& sh = SharedVarDCTTables::Get();
    uint32_t num_ac_clusters = 0;
    const std::vector<uint8_t> ac_cluster_of = AcContextClusters(&num_ac_clusters);
    DevEFrame ef{};
    uint32_t global_scale = 0, quant_dc = 0;
    FillQuantizer(p, &ef, &global_scale, &quant_dc);
    FrameHeader fh0;
    fh0.xsize = xsize;
    fh0.ysize = ysize;
    const EncTree tree = BuildEncTree(ToFrameDimensions(fh0).num_dc_groups);
    const EncLayout L = LayoutEncFrame(xsize, ysize, num_ac_clusters, tree.num_leaves, &ef, p.alpha);
    std::vector<float> farena(L.fsize + 16, 0.0f);
    std::vector<int32_t> iarena(L.isize + 16, 0);
    std::vector<uint8_t> barena(L.bsize + 16, 0xFF);
    std::vector<uint2> tokens(L.tsize + 16);
    float lut[256];
    for (int i = 0; i < 256; i++) lut[i] = SrgbToLinearHost(i / 255.0f);
    DevEPools E{};
    E.bytes_in = rgb;
    E.farena = farena.data();
    E.iarena = iarena.data();
    E.barena = barena.data();
    E.tokens = tokens.data();
    E.srgb_lut = lut;
    E.fpool = sh.fpool.data();
    E.opool = sh.opool.data();
    E.upool = sh.upool.data();
    for (int i = 0; i < 17; i++) E.table_off[i] = sh.table_off[i];
    for (int i = 0; i < 13; i++) E.order_off[i] = sh.order_off[i];
    E.wc_off = sh.wc_off;
    E.sinfo_off = sh.sinfo_off;
    E.ctxtab_off = sh.ctxtab_off;
    E.ac_cluster_of = ac_cluster_of.data();
    E.tree = tree.nodes.data();
    ef.rgb = 0;
    const FrameDimensions& d = L.dim;
    const uint32_t W = d.xsize_blocks, H = d.ysize_blocks;
    for (uint32_t y = 0; y < H * 8; y++)
      for (uint32_t x = 0; x < W * 8; x++) DevEncXybPixel(E, ef, x, y);
    if (ef.gab) {
      for (uint32_t c = 0; c < 3; c++)
        for (uint32_t y = 0; y < H * 8; y++)
          for (uint32_t x = 0; x < W * 8; x++) {
            const bool in = x >= 2 && y >= 2 && x + 2 < W * 8 && y + 2 < H * 8;
            if (in) DevEncGaborishInvPixel<true>(E, ef, c, x, y);
            else DevEncGaborishInvPixel<false>(E, ef, c, x, y);
          }
    }
    if (ef.adaptive) {  // k_enc_aq
      std::vector<float> aq_sm(kAqSmemFloats);
      for (uint32_t ty = 0; ty < (H + 7) / 8; ty++)
        for (uint32_t tx = 0; tx < (W + 7) / 8; tx++) DevEncAqTile<0>(E, ef, tx, ty, 0, 1, aq_sm.data());
      if (std::getenv("JXLO_DEBUG_AQ"))
        std::fprintf(stderr, "E aq: first %a %a %a\n", farena[ef.quant_field], farena[ef.quant_field + 1], farena[ef.quant_field + W]);
    }
    for (uint32_t g = 0; g < d.num_groups; g++) DevEncStrategyGroup(E, ef, g);
    for (uint32_t g = 0; g < d.num_dc_groups; g++) DevEncNumberBlocks(E, ef, g);
    for (uint32_t by = 0; by < H; by++)
      for (uint32_t bx = 0; bx < W; bx++) DevEncDcBlock(E, ef, bx, by);
    std::vector<float> buf(4 * 4096 + 64), cfl_vals(4 * 4096), red(64);
    for (uint32_t by = 0; by < H; by++)  // k_enc_coeffs<0>: forward transforms
      for (uint32_t bx = 0; bx < W; bx++) {
        const uint8_t a = barena[static_cast<size_t>(by) * W + bx];
        if (a & 1) DevEncVarblock<0, 0>(E, ef, bx, by, a >> 1, buf.data(), 0, 1);
      }
    for (uint32_t ty = 0; ty < ef.cmh; ty++)  // k_enc_cfl
      for (uint32_t tx = 0; tx < ef.cmw; tx++) {
        DevEncCflTile<0>(E, ef, tx, ty, 0, 1, cfl_vals.data(), red.data());
        if (std::getenv("JXLO_DEBUG_CFL"))
          std::fprintf(stderr, "E tile %u %u: x=%d b=%d v0=%a %a %a %a\n", tx, ty, static_cast<int8_t>(barena[ef.ytox + ty * ef.cmw + tx]),
                       static_cast<int8_t>(barena[ef.ytob + ty * ef.cmw + tx]), cfl_vals[70], cfl_vals[4096 + 70], cfl_vals[8192 + 70],
                       cfl_vals[12288 + 70]);
      }
    for (uint32_t c = 0; c < 3; c++)  // k_enc_adjust
      for (uint32_t by = 0; by < H; by++)
        for (uint32_t bx = 0; bx < W; bx++) DevEncAdjustVarblockChannel(E, ef, bx, by, c);
    for (uint32_t by = 0; by < H; by++)  // k_enc_coeffs<1>: quantisation
      for (uint32_t bx = 0; bx < W; bx++) {
        const uint8_t a = barena[static_cast<size_t>(by) * W + bx];
        if (a & 1) DevEncVarblock<0, 1>(E, ef, bx, by, a >> 1, buf.data(), 0, 1);
      }
    // coefficient orders: statistics kernels, then the host's sort, then the custom orders go back to the "device"
    CustomOrders orders;
    std::vector<uint16_t> custom_pool;
    const std::vector<uint8_t> sample_bits = MakeOrderSampleBits(static_cast<size_t>(W) * H);
    E.sample_bits = sample_bits.data();
    if (p.coeff_orders) {
      for (uint32_t g = 0; g < d.num_groups; g++) DevEncGroupOrders(E, ef, g);
      std::vector<uint32_t> local(kCustomOrderCounters + 1024);
      for (uint32_t g = 0; g < d.num_groups; g++) DevEncOrderStatsGroup<0>(E, ef, g, 0, 1, local.data());
      orders = ComputeCustomOrders(static_cast<uint32_t>(iarena[ef.order_mask]), iarena.data() + ef.zero_counts, W, H);
      for (uint32_t ord = 0; ord < kNumCustomOrders; ord++)
        for (uint32_t c = 0; c < 3; c++) {
          if (!(orders.used & (1u << ord))) continue;
          ef.custom_order[3 * ord + c] = custom_pool.size();
          custom_pool.insert(custom_pool.end(), orders.order[ord][c].begin(), orders.order[ord][c].end());
        }
    }
    E.opool_custom = custom_pool.data();
    uint16_t ctxtab[128];
    for (int i = 0; i < 128; i++) ctxtab[i] = static_cast<uint16_t>(sh.upool[sh.ctxtab_off + i]);
    for (uint32_t c = 0; c < 3; c++)
      for (uint32_t by = 0; by < H; by++)
        for (uint32_t bx = 0; bx < W; bx++) DevEncBlockStats(E, ef, bx, by, c);
    for (uint32_t g = 0; g < d.num_groups; g++) DevEncTokenOffsets(E, ef, g);
    for (uint32_t c = 0; c < 3; c++)
      for (uint32_t by = 0; by < H; by++)
        for (uint32_t bx = 0; bx < W; bx++) DevEncBlockTokens(E, ef, bx, by, c, ctxtab, ctxtab + 64);
    for (uint32_t g = 0; g < d.num_dc_groups; g++) {
      const DevDcGroupLayout gl = DevDcGroupGeometry(E, ef, g);
      for (uint32_t i = 0; i < gl.dc_tokens + gl.meta_tokens; i++) DevEncModularSample(E, ef, g, gl, i);
    }
    const bool alpha_global = p.alpha && xsize <= 256 && ysize <= 256;
    std::vector<std::pair<uint32_t, uint32_t>> global_alpha;
    if (p.alpha) {  // k_enc_alpha
      for (uint64_t i = 0; i < static_cast<uint64_t>(xsize) * ysize; i++) DevEncAlphaSample(E, ef, i);
      if (alpha_global)
        for (uint64_t i = 0; i < static_cast<uint64_t>(xsize) * ysize; i++)
          global_alpha.push_back({tokens[ef.alpha_tokens + i].x, tokens[ef.alpha_tokens + i].y});
    }
    EncGlobals G;
    const auto t_globals = std::chrono::steady_clock::now();
    BuildEncGlobals(p, L, tree, ac_cluster_of, global_scale, quant_dc, reinterpret_cast<uint32_t*>(iarena.data() + ef.mod_hist),
                    reinterpret_cast<uint32_t*>(iarena.data() + ef.ac_hist), orders, &G, alpha_global ? &global_alpha : nullptr);
    if (std::getenv("JXLB_EMUL_TIMING"))
      std::fprintf(stderr, "BuildEncGlobals: %.2f ms\n",
                   std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_globals).count());
    const std::vector<uint32_t> mod_fs = G.mod_code.Fs(), ac_fs = G.ac_code.Fs();
    DevEncCode mod{mod_fs.data(), G.mod_code.reverse.data()};
    DevEncCode ac{ac_fs.data(), G.ac_code.reverse.data()};
    std::vector<std::vector<uint32_t>> dc_words(d.num_dc_groups), ac_words(d.num_groups);
    std::vector<EncSection> dcg, acg;
    for (uint32_t g = 0; g < d.num_dc_groups; g++) {
      const DevDcGroupLayout gl = DevDcGroupGeometry(E, ef, g);
      const size_t cap = (static_cast<size_t>(gl.dc_tokens + gl.meta_tokens) * 6 + 64) / 4 + 4;
      dc_words[g].assign(cap, 0);
      const uint64_t end = (cap - 1) * 32 - 5;  // an unaligned end, like a shared region could have
      const uint64_t first = DevEncEmitDcGroup(E, ef, g, mod, dc_words[g].data(), end);
      dcg.push_back({dc_words[g].data(), first, end - first});
    }
    for (uint32_t g = 0; g < d.num_groups; g++) {
      const uint32_t n = static_cast<uint32_t>(iarena[ef.group_tokens + g]);
      const size_t cap = (static_cast<size_t>(n) * 6 + 64) / 4 + 4 + (p.alpha ? (65536 * 6 + 64) / 4 + 4 : 0);
      ac_words[g].assign(cap, 0);
      const uint64_t end = cap * 32;
      uint64_t first;
      if (p.alpha && !alpha_global) {  // as k_enc_emit
        const uint32_t gx = g % ef.xgroups, gy = g / ef.xgroups;
        const uint32_t gw = std::min<uint32_t>(256, xsize - (gx << 8)), gh = std::min<uint32_t>(256, ysize - (gy << 8));
        first = DevEncEmitAcGroup(tokens.data() + ef.ac_tokens + static_cast<size_t>(g) * 3 * 65536, n, ac, ac_words[g].data(), end,
                                  tokens.data() + ef.alpha_tokens + static_cast<size_t>(g) * 65536, gw * gh, &mod);
      } else {
        first = DevEncEmitAcGroup(tokens.data() + ef.ac_tokens + static_cast<size_t>(g) * 3 * 65536, n, ac, ac_words[g].data(), end);
      }
      acg.push_back({ac_words[g].data(), first, end - first});
    }
    const std::vector<uint8_t> cs = AssembleCodestream(p, L, G, dcg, acg);
    if (cs.size() > out_cap) throw Error("output buffer too small");
    std::memcpy(out, cs.data(), cs.size());
    return static_cast<long>(cs.size());
  } catch (const std::exception& e) {
    std::snprintf(err, errlen, "%s", e.what());
    return -1;
  }
}

// Lossless (Modular) encode of one image with the device functions of kernels/jxlb_encl_dev.h run on the CPU and the
// host code of host/jxlb_encl_host.h, in the order of JxlB200EncoderEncodeLosslessBatch; returns the size or -1.
// DevFwdDct8 (the register transform of the DCT8X8 fast path, kernels/jxlb_enc_dev.h) against the generic staged
// forward DCT (CoopDCT, n = 8) on `trials` seeded random lines, including large and tiny magnitudes: the number of
// lines whose eight coefficients are not bit-identical.
long jxlb_emul_fwd_dct8_mismatches(uint32_t seed, uint32_t trials) {
  const SharedVarDCTTables& sh = SharedVarDCTTables::Get();
  const float* wc = sh.fpool.data() + sh.wc_off;
  std::mt19937 rng(seed);
  long bad = 0;
  for (uint32_t t = 0; t < trials; t++) {
    float v[8], a[8], b[8];
    const float scale = std::ldexp(1.0f, static_cast<int>(rng() % 40) - 30);
    for (int i = 0; i < 8; i++) v[i] = a[i] = (static_cast<float>(rng() % 2000001) - 1000000.0f) * 1e-6f * scale;
    DevFwdDct8(v, wc);
    const float* r = CoopDCT<0>(8, 1, 8, 1, a, b, wc, 0, 1);
    if (std::memcmp(v, r, sizeof(v)) != 0) bad++;
  }
  return bad;
}

long jxlb_emul_encode_lossless(const void* pixels, uint32_t xsize, uint32_t ysize, uint32_t num_channels, uint32_t bits,
                               uint8_t* out, size_t out_cap, char* err, size_t errlen) {
  try {
    JXLB_CHECK(num_channels >= 1 && num_channels <= 4 && (bits == 8 || bits == 16) && xsize > 0 && ysize > 0, "bad arguments");
    const EnclTree tree = BuildEnclTree();
    EnclParams p;
    p.xsize = xsize;
    p.ysize = ysize;
    p.nch = num_channels;
    p.bits = bits;
    DevLFrame f{};
    f.xsize = xsize;
    f.ysize = ysize;
    f.xgroups = p.XGroups();
    f.ygroups = p.YGroups();
    f.nch = num_channels;
    f.bytes = bits / 8;
    const uint64_t px = static_cast<uint64_t>(xsize) * ysize;
    const uint32_t groups = f.xgroups * f.ygroups;
    std::vector<int32_t> planes(px * num_channels);
    std::vector<uint2> tokens(static_cast<size_t>(groups) * num_channels * kEnclGroupSamples);
    std::vector<uint32_t> hist(34 * 256, 0);
    std::vector<int32_t> cutoffs(kEnclCutoffValues, kEnclCutoffValues + 33);
    DevLPools L{};
    L.in = static_cast<const uint8_t*>(pixels);
    L.planes = planes.data();
    L.tokens = tokens.data();
    L.hist = hist.data();
    L.cutoffs = cutoffs.data();
    L.leaf_of = tree.leaf_of;
    for (uint64_t i = 0; i < px; i++) DevEnclPlanes(L, f, i);
    for (uint64_t i = 0; i < px * num_channels; i++) DevEnclToken(L, f, i);
    BitWriter global;
    EncCode code;
    WriteEnclGlobal(global, p, tree, hist.data(), &code);
    const std::vector<uint32_t> fs = code.Fs();
    const DevEncCode dcode{fs.data(), code.reverse.data()};
    std::vector<std::vector<uint32_t>> words(groups);
    std::vector<EncSection> secs;
    for (uint32_t g = 0; g < groups; g++) {
      const uint32_t gx = g % f.xgroups, gy = g / f.xgroups;
      const uint64_t gw = std::min<uint32_t>(kEnclGroupDim, xsize - gx * kEnclGroupDim), gh = std::min<uint32_t>(kEnclGroupDim, ysize - gy * kEnclGroupDim);
      const size_t cap = EnclSectionWords(gw * gh * num_channels);
      words[g].assign(cap + 1, 0);
      const uint64_t end = cap * 32;
      const uint64_t first = DevEnclEmitGroup(L, f, g, dcode, words[g].data(), end, !p.GlobalOnly());
      secs.push_back({words[g].data(), first, end - first});
    }
    const std::vector<uint8_t> cs = AssembleEncl(p, global, secs);
    if (cs.size() > out_cap) throw Error("output buffer too small");
    std::memcpy(out, cs.data(), cs.size());
    return static_cast<long>(cs.size());
  } catch (const std::exception& e) {
    std::snprintf(err, errlen, "%s", e.what());
    return -1;
  }
}
}
