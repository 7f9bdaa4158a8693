// jxl_b200: merges per-frame plans into the flat pools of one batch.
#ifndef JXLB_BATCH_H_
#define JXLB_BATCH_H_

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <thread>
#include <tuple>

#include "jxlb_frame_plan.h"

namespace jxlb {

struct BatchPlan {
  std::vector<uint8_t> bytes;  // all codestreams, each starting on a 4-byte boundary, zero padded
  // ... or, when the caller of PlanBatch provides the storage (pinned host memory the upload then reads directly,
  // filled by the planning threads): the same layout at `ext_bytes`, and `bytes` stays empty
  const uint8_t* ext_bytes = nullptr;
  uint64_t ext_bytes_size = 0;
  std::vector<DevAlias> alias;
  std::vector<uint32_t> prefix, cfg, refs;
  std::vector<uint16_t> lut;
  std::vector<DevTreeNode> tree;
  std::vector<DevCode> codes;
  std::vector<DevChannel> chans;
  std::vector<DevStream> streams;
  std::vector<DevPlane> planes;
  std::vector<DevOp> ops;
  std::vector<DevProgram> group_programs;
  std::vector<std::vector<DevProgram>> levels;  // levels[k] = k-th global op of every frame that has one
  std::vector<DevFrameOut> frames;
  std::vector<uint64_t> frame_out_size;
  uint64_t arena_size = 0, out_size = 0;
  uint32_t wp_slots = 0, wp_width = 0, lz77_slots = 0;
  uint32_t max_frame_pixels = 0;
  uint64_t total_pixels = 0;
  uint64_t compressed_bytes = 0;
  std::vector<BasicInfo> info;
  std::vector<uint32_t> warp_chans, warp_dims_off, warp_dims;
  uint32_t num_coop = 0;  // streams [0, num_coop) go one per warp (k_modular_decode_coop), the rest in lock-step bundles
  // Streams [num_early, size) are chained behind AC coefficient streams (DevStream::chain_slot): they are decoded by a
  // second Modular launch after the AC decode kernel, one per warp or in lock-step bundles like the early ones
  // (late_coop of them one per warp); the group programs and frame levels of VarDCT frames (the extra channels' copies
  // and inverse transforms) run after that launch.
  uint32_t num_early = 0, late_coop = 0, chain_slots = 0;
  uint32_t early_warps = 0;  // lock-step bundles of the early streams (warp_chans / warp_dims index of the first late bundle)
  std::vector<float> spl_seg;
  std::vector<uint32_t> spl_idx;
  std::vector<DevProgram> late_group_programs;
  std::vector<std::vector<DevProgram>> late_levels;
  bool narrow = true;  // every frame promises that 16-bit buffers suffice -> 32-bit predictor math
  // VarDCT frames
  std::vector<DevVFrame> vframes;
  std::vector<uint32_t> vframe_frame;  // vframes[k] is frame vframe_frame[k] of the batch
  std::vector<DevAcStream> ac_streams;
  // per (frame, pass): the ids of its streams in ac_streams, longest first (k_ac_decode_frame)
  bool any_upsampling = false;  // some lossy frame is upsampled: the per-pixel render kernels run for it
  std::vector<DevAcUnit> ac_units;
  std::vector<uint32_t> ac_unit_streams;
  std::vector<float> fpool;
  std::vector<uint16_t> opool;
  std::vector<uint8_t> cpool;
  std::vector<uint32_t> upool;
  uint64_t farena_size = 0, barena_size = 0, uarena_size = 0, tok_size = 0;
  std::vector<DevPatch> patches;
  std::vector<DevRefFrame> ref_frames;
  uint32_t max_patch_pixels = 0;
  uint64_t pix_plane_max = 0;  // floats per pixel plane slot
  uint32_t wave_frames = 1;    // VarDCT frames whose pixel planes are live at the same time
};

// Pixel planes are the largest intermediate (24 bytes per pixel for two sets of three float
// planes), so they are allocated for a wave of frames and reused: the entropy kernels run over
// the whole batch, the dequant / IDCT / filter / colour kernels wave by wave.
constexpr uint64_t kWavePixelBytes = uint64_t{4} << 30;

// Orders the streams so that the 32 lanes of a warp decode planes of the same shape
// (they run one loop nest in lock step) and derives the warp-uniform loop bounds.
// A stream the warp-cooperative kernel can decode (kernels/jxlb_modular_coop_dev.h): plain ANS without LZ77, every
// channel on one of the table paths (FramePlanner::BuildCoopLut).
inline bool CoopEligible(const BatchPlan& b, const DevStream& s) {
  const DevCode& c = b.codes[s.code];
  if (c.use_prefix || c.lz77_enabled || s.chan_end == s.chan_begin) return false;
  for (uint32_t k = s.chan_begin; k < s.chan_end; k++) {
    const DevChannel& ch = b.chans[k];
    if (!ch.coop) return false;
    if (ch.nw_lut != 0) continue;
    if (ch.wp_lut != 0 && ch.uses_wp != 0 && ch.ref_count == 0 && !ch.dyn) continue;
    return false;
  }
  return true;
}

inline void BundleStreams(BatchPlan* b) {
  // One warp per stream pays while the batch leaves the chip empty -- the kernel then takes as long as its longest chain
  // and a chain is 1.7 x (1024 chains) to 2.6 x (64 chains) faster that way -- but it holds 32 lanes' registers per
  // stream for that time; a big batch of lossy frames is bound by the register file shared with the per-pixel kernels
  // of the other handles, where the lock-step kernel (8 streams per warp) is the cheaper one. Hence a cap
  // of one warp per scheduler. (JXLB200_NO_COOP=1: never; JXLB200_COOP_MAX=n: another cap -- comparison runs, tests.)
  uint32_t coop_max = 4 * 148;
  if (const char* e = std::getenv("JXLB200_COOP_MAX")) coop_max = static_cast<uint32_t>(std::strtoul(e, nullptr, 10));
  bool coop_on = std::getenv("JXLB200_NO_COOP") == nullptr;
  if (coop_on) {
    uint32_t eligible = 0;
    for (const DevStream& s : b->streams) eligible += CoopEligible(*b, s) ? 1u : 0u;
    coop_on = eligible <= coop_max;
  }
  auto key = [&](const DevStream& s) {
    const uint32_t n = s.chan_end - s.chan_begin;
    uint32_t w = 0, h = 0;
    if (n) {
      const DevPlane& p = b->planes[b->chans[s.chan_end - 1].plane];
      w = p.w;
      h = p.h;
    }
    // early streams first, then the ones chained behind AC coefficient streams; in each part the one-per-warp streams first
    return std::make_tuple(s.chain_slot == 0 ? 1u : 0u, coop_on && CoopEligible(*b, s) ? 1u : 0u, n, w, h);
  };
  std::stable_sort(b->streams.begin(), b->streams.end(),
                   [&](const DevStream& x, const DevStream& y) { return key(x) > key(y); });
  b->num_coop = b->num_early = b->late_coop = 0;
  for (const DevStream& s : b->streams) {
    const auto k = key(s);
    if (std::get<0>(k) != 0) {
      b->num_early++;
      if (std::get<1>(k) != 0) b->num_coop++;
    } else if (std::get<1>(k) != 0) {
      b->late_coop++;
    }
  }
  b->warp_chans.clear();
  b->warp_dims_off.clear();
  b->warp_dims.clear();
  // lock-step bundles of 32 streams: [num_coop, num_early), then [num_early + late_coop, size)
  auto bundle = [&](size_t first, size_t last) {
    for (size_t s0 = first; s0 < last; s0 += 32) {
      const size_t s1 = std::min(last, s0 + 32);
      uint32_t nmax = 0;
      for (size_t s = s0; s < s1; s++) nmax = std::max(nmax, b->streams[s].chan_end - b->streams[s].chan_begin);
      b->warp_chans.push_back(nmax);
      b->warp_dims_off.push_back(b->warp_dims.size());
      for (uint32_t k = 0; k < nmax; k++) {
        uint32_t mw = 0, mh = 0;
        for (size_t s = s0; s < s1; s++) {
          const DevStream& st = b->streams[s];
          if (k >= st.chan_end - st.chan_begin) continue;
          const DevPlane& p = b->planes[b->chans[st.chan_begin + k].plane];
          if (p.w == 0 || p.h == 0) continue;
          mw = std::max(mw, p.w);
          mh = std::max(mh, p.h);
        }
        b->warp_dims.push_back(mw);
        b->warp_dims.push_back(mh);
      }
    }
  };
  bundle(b->num_coop, b->num_early);
  b->early_warps = b->warp_chans.size();
  bundle(b->num_early + b->late_coop, b->streams.size());
}

inline void MergeFrame(const FramePlan& f, const uint8_t* cs, size_t cs_size, BatchPlan* b, int64_t ext_base = -1) {
  uint64_t byte_base = b->bytes.size();
  if (ext_base >= 0) {  // the bytes already sit at ext_bytes + ext_base
    byte_base = static_cast<uint64_t>(ext_base);
  } else {
    b->bytes.insert(b->bytes.end(), cs, cs + cs_size);
    b->bytes.resize((b->bytes.size() + 3) & ~size_t{3}, 0);
  }
  b->compressed_bytes += cs_size;
  const uint32_t alias0 = b->alias.size(), prefix0 = b->prefix.size(), cfg0 = b->cfg.size(), refs0 = b->refs.size();
  const uint32_t tree0 = b->tree.size(), codes0 = b->codes.size(), chans0 = b->chans.size(), lut0 = b->lut.size();
  b->lut.insert(b->lut.end(), f.lut.begin(), f.lut.end());
  const uint32_t planes0 = b->planes.size(), ops0 = b->ops.size();
  const uint64_t arena0 = b->arena_size;
  b->alias.insert(b->alias.end(), f.alias.begin(), f.alias.end());
  b->prefix.insert(b->prefix.end(), f.prefix.begin(), f.prefix.end());
  b->cfg.insert(b->cfg.end(), f.cfg.begin(), f.cfg.end());
  for (uint32_t r : f.refs) b->refs.push_back(r + planes0);
  // tree children are relative to each pruned tree's root, no relocation needed
  b->tree.insert(b->tree.end(), f.tree.begin(), f.tree.end());
  for (DevCode c : f.codes) {
    c.alias_off += alias0;
    c.prefix_off += prefix0;
    c.cfg_off += cfg0;
    b->codes.push_back(c);
  }
  for (DevChannel c : f.chans) {
    c.plane += planes0;
    c.ref_off += refs0;
    c.tree_off += tree0;
    c.lut_off += lut0;
    c.coop_lut_off += lut0;
    c.coop_list_off += lut0;
    b->chans.push_back(c);
  }
  for (DevStream s : f.streams) {
    if (s.chain_slot == 0) s.bit_pos += byte_base * 8;
    if (s.chain_slot != 0) s.chain_slot += b->chain_slots;
    s.bit_end += byte_base * 8;
    s.code += codes0;
    s.chan_begin += chans0;
    s.chan_end += chans0;
    if (s.lz77_slot != 0xFFFFFFFFu) s.lz77_slot += b->lz77_slots;
    b->streams.push_back(s);
  }
  for (DevPlane p : f.planes) {
    p.off += arena0;
    b->planes.push_back(p);
  }
  for (DevOp o : f.ops) {
    o.a += planes0;
    o.b += planes0;
    if (o.kind != kOpCopy) o.c += planes0;
    if (o.kind == kOpPalette) o.pad += planes0;
    b->ops.push_back(o);
  }
  for (size_t k = 0; k < f.group_programs.size(); k++) {
    DevProgram p = f.group_programs[k];
    p.op_begin += ops0;
    p.op_end += ops0;
    (k >= f.late_programs0 ? b->late_group_programs : b->group_programs).push_back(p);
  }
  {
    const size_t early = std::min(f.frame_levels.size(), f.late_levels0);
    if (b->levels.size() < early) b->levels.resize(early);
    if (b->late_levels.size() < f.frame_levels.size() - early) b->late_levels.resize(f.frame_levels.size() - early);
    for (size_t k = 0; k < f.frame_levels.size(); k++) {
      DevProgram p = f.frame_levels[k];
      p.op_begin += ops0;
      p.op_end += ops0;
      if (k < early) {
        b->levels[k].push_back(p);
      } else {
        b->late_levels[k - early].push_back(p);
      }
    }
  }
  if (f.is_vardct) {
    const VarDCTPlan& v = f.v;
    if (b->vframes.empty()) {  // shared tables sit at offset 0 of the pools
      const SharedVarDCTTables& sh = SharedVarDCTTables::Get();
      JXLB_CHECK(b->fpool.empty() && b->opool.empty(), "internal: pools not empty");
      JXLB_CHECK(b->upool.empty(), "internal: pools not empty");
      b->fpool = sh.fpool;
      b->opool = sh.opool;
      b->upool = sh.upool;
    }
    const uint32_t fpool0 = b->fpool.size(), opool0 = b->opool.size(), cpool0 = b->cpool.size(), upool0 = b->upool.size();
    b->fpool.insert(b->fpool.end(), v.fpool.begin(), v.fpool.end());
    b->opool.insert(b->opool.end(), v.opool.begin(), v.opool.end());
    b->cpool.insert(b->cpool.end(), v.cpool.begin(), v.cpool.end());
    b->upool.insert(b->upool.end(), v.upool.begin(), v.upool.end());
    DevVFrame vf = v.vf;
    for (int c = 0; c < 3; c++) {
      vf.dc[c] += b->farena_size;
      vf.dc_final[c] += b->farena_size;
    }
    vf.inv_sigma += b->farena_size;
    vf.acs += b->barena_size;
    vf.qdc += b->barena_size;
    vf.sharp += b->barena_size;
    vf.rawq += b->barena_size;
    vf.ytox += b->barena_size;
    vf.ytob += b->barena_size;
    vf.tok_start += b->uarena_size;
    vf.tok_count += b->uarena_size;
    vf.blist += b->uarena_size;
    vf.blist_count += b->uarena_size;
    vf.bctx_off += upool0;
    vf.order_index += upool0;
    vf.dcg_index += upool0;
    for (uint32_t i = 0; i < vf.num_passes * 39; i++) {
      uint32_t& e = b->upool[vf.order_index + i];
      e = (e & kSharedFlag) ? (e & ~kSharedFlag) : e + opool0;
    }
    for (uint32_t g = 0; g < vf.xdcgroups * vf.ydcgroups; g++)
      for (uint32_t i = 0; i < 7; i++) b->upool[vf.dcg_index + 8 * g + i] += planes0;
    for (uint32_t k = 0; k < 17; k++)
      vf.table_off[k] = (vf.table_off[k] & kSharedFlag) ? (vf.table_off[k] & ~kSharedFlag) : vf.table_off[k] + fpool0;
    for (uint32_t p = 0; p < vf.num_passes; p++) {
      vf.ac_code[p] += codes0;
      vf.ctx_map_off[p] += cpool0;
    }
    vf.out_off = b->out_size;
    if (vf.alpha_plane != kNoPlane) vf.alpha_plane += planes0;
    if (vf.has_splines) {
      vf.spl_seg += b->spl_seg.size();
      vf.spl_rows += b->spl_idx.size();
      vf.spl_idx += b->spl_idx.size();
      b->spl_seg.insert(b->spl_seg.end(), f.spl_seg.begin(), f.spl_seg.end());
      b->spl_idx.insert(b->spl_idx.end(), f.spl_idx.begin(), f.spl_idx.end());
    }
    if (vf.upsampling > 1) {
      for (int c = 0; c < 3; c++) vf.up_pix[c] += b->farena_size;
      vf.up_kernel += fpool0;
      b->any_upsampling = true;
    }
    vf.patch_begin = b->patches.size();
    for (DevPatch p : v.patches) {
      for (int c = 0; c < 3; c++) p.src[c] += b->farena_size;
      b->max_patch_pixels = std::max(b->max_patch_pixels, p.xsize * p.ysize);
      b->patches.push_back(p);
    }
    for (DevRefFrame r : v.ref_frames) {
      r.plane_y += planes0;
      r.plane_x += planes0;
      r.plane_b += planes0;
      for (int c = 0; c < 3; c++) r.dst[c] += b->farena_size;
      b->ref_frames.push_back(r);
    }
    for (DevAcStream s : v.ac_streams) {
      if (s.chain_slot != 0) s.chain_slot += b->chain_slots;
      s.bit_pos += byte_base * 8;
      s.bit_end += byte_base * 8;
      s.frame = b->vframes.size();
      s.tok_off += b->tok_size;
      b->ac_streams.push_back(s);
    }
    b->farena_size += v.farena_size;
    b->barena_size += v.barena_size;
    b->uarena_size += v.uarena_size;
    b->tok_size += v.tok_size;
    b->pix_plane_max = std::max(b->pix_plane_max, v.pix_plane);
    b->vframe_frame.push_back(static_cast<uint32_t>(b->frames.size()));
    b->vframes.push_back(vf);
  }
  DevFrameOut fo = f.out;
  for (uint32_t c = 0; c < 4; c++)
    if (c < fo.num_channels && fo.plane[c] != kNoPlane) fo.plane[c] += planes0;
  fo.out_off = b->out_size;
  if (fo.has_splines) {
    fo.spl_seg += b->spl_seg.size();
    fo.spl_rows += b->spl_idx.size();
    fo.spl_idx += b->spl_idx.size();
    b->spl_seg.insert(b->spl_seg.end(), f.spl_seg.begin(), f.spl_seg.end());
    b->spl_idx.insert(b->spl_idx.end(), f.spl_idx.begin(), f.spl_idx.end());
  }
  const uint64_t osize = fo.stride * ((fo.orient & 4) ? fo.xsize : fo.ysize);
  b->out_size += (osize + 255) & ~uint64_t{255};
  b->frame_out_size.push_back(osize);
  b->frames.push_back(fo);
  b->arena_size += f.arena_size;
  b->wp_slots += f.wp_slots;
  b->wp_width = std::max(b->wp_width, f.wp_width);
  b->lz77_slots += f.lz77_slots;
  b->chain_slots += f.chain_slots;
  b->max_frame_pixels = std::max<uint32_t>(b->max_frame_pixels, f.pixels);
  b->total_pixels += f.pixels;
  if (!(f.meta.modular_16_bit_buffer_sufficient && f.meta.bit_depth.bits <= 16 && !f.meta.bit_depth.floating_point))
    b->narrow = false;
  for (const auto& e : f.meta.extra)
    if (e.bit_depth.bits > 16 || e.bit_depth.floating_point) b->narrow = false;
  BasicInfo bi;
  bi.xsize = f.xsize;
  bi.ysize = f.ysize;
  bi.meta = f.meta;
  b->info.push_back(bi);
}

// After a run in which some AC streams produced more tokens than their capacity: gives every
// stream at least what it used (`used[s]`, reported by the decode kernel) and lays the token
// arena out again. Returns false if nothing overflowed.
inline bool GrowTokenCapacity(BatchPlan* b, const uint32_t* used) {
  bool grew = false;
  for (size_t s = 0; s < b->ac_streams.size(); s++) {
    if (used[s] > b->ac_streams[s].tok_cap) {
      b->ac_streams[s].tok_cap = used[s] + 64;
      grew = true;
    }
  }
  if (!grew) return false;
  uint64_t off = 0;
  for (DevAcStream& s : b->ac_streams) {
    s.tok_off = off;
    off += s.tok_cap;
  }
  b->tok_size = off;
  return true;
}

// Runs the Modular streams of a probe batch on the device: `end_bits[s]` = first bit after stream s (absolute, in
// `probe.bytes`), `arena` = the sample arena afterwards. Throws on a failed stream.
using ProbeFn = std::function<void(const BatchPlan& probe, std::vector<uint64_t>* end_bits, std::vector<int32_t>* arena)>;

constexpr int kMaxProbeRounds = 24;  // one DC chain + at most 17 raw quantisation tables, with slack

// Plans `n` files on `threads` host threads and merges them in input order. Files whose plan depends on where the
// device stops reading a stream (ProbeCtx) are planned in rounds, one launch of `probe` per round for all of them.
// `bytes_alloc(size)`: storage for the batch's codestream bytes (see BatchPlan::ext_bytes), or empty.
using BytesAlloc = std::function<uint8_t*(size_t)>;

inline void PlanBatch(const uint8_t* const* files, const size_t* sizes, size_t n, const PixelFormat& fmt,
                      int threads, BatchPlan* batch, const ProbeFn& probe = ProbeFn(), const BytesAlloc& bytes_alloc = BytesAlloc()) {
  const bool timing = std::getenv("JXLB200_PLAN_TIMING") != nullptr;  // stderr: where the host time of a batch goes
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&](std::chrono::steady_clock::time_point t) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
  };
  double t_frames = 0;
  std::vector<FramePlan> plans(n);
  std::vector<CodestreamView> views(n);
  std::vector<std::string> errors(n);
  std::vector<ProbeCtx> ctx(n);
  std::vector<size_t> todo(n);
  for (size_t i = 0; i < n; i++) todo[i] = i;
  threads = std::max(1, std::min<int>(threads, static_cast<int>(n)));
  // Where every codestream goes in the byte pool is known as soon as the containers are opened: the planning threads
  // copy their file there themselves (in parallel, into memory that is already mapped) instead of one thread
  // appending 1 - 2 MB per file to a growing vector afterwards.
  std::vector<uint64_t> byte_off(n, 0);
  uint8_t* ext = nullptr;
  if (bytes_alloc) {
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) {
      try {
        views[i] = FindCodestream(files[i], sizes[i]);
      } catch (const std::exception& e) {
        throw Error("frame " + std::to_string(i) + ": " + e.what());
      }
      byte_off[i] = total;
      total += (views[i].size + 3) & ~size_t{3};
    }
    ext = bytes_alloc(total + 64);
    JXLB_CHECK(ext != nullptr, "out of memory (codestream staging)");
    std::memset(ext + total, 0, 64);  // read-ahead padding for the device bit reader
    batch->ext_bytes = ext;
    batch->ext_bytes_size = total + 64;
  }
  for (int round = 0; !todo.empty(); round++) {
    JXLB_CHECK(round < kMaxProbeRounds, "too many chained sub-streams");
    std::atomic<size_t> next{0};
    auto work = [&]() {
      for (;;) {
        const size_t k = next.fetch_add(1);
        if (k >= todo.size()) break;
        const size_t i = todo[k];
        try {
          if (round == 0 && !ext) views[i] = FindCodestream(files[i], sizes[i]);
          if (round == 0 && ext) {
            std::memcpy(ext + byte_off[i], views[i].data, views[i].size);
            std::memset(ext + byte_off[i] + views[i].size, 0, ((views[i].size + 3) & ~size_t{3}) - views[i].size);
          }
          plans[i] = FramePlan();
          ctx[i].next = 0;
          ctx[i].pending = false;
          PlanCodestream(views[i].data, views[i].size, fmt, &plans[i], probe ? &ctx[i] : nullptr);
        } catch (const ProbePending&) {
        } catch (const std::exception& e) {
          errors[i] = e.what();
          if (errors[i].empty()) errors[i] = "unknown error";
        }
      }
    };
    const auto t_round = std::chrono::steady_clock::now();
    if (threads == 1 || todo.size() == 1) {
      work();
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < std::min<int>(threads, todo.size()); t++) pool.emplace_back(work);
      for (auto& t : pool) t.join();
    }
    t_frames += since(t_round);
    for (size_t i = 0; i < n; i++) {
      if (!errors[i].empty()) throw Error("frame " + std::to_string(i) + ": " + errors[i]);
    }
    std::vector<size_t> pending;
    for (size_t i : todo)
      if (ctx[i].pending) pending.push_back(i);
    todo = pending;
    if (todo.empty()) break;
    // one probe batch for every file that waits for the device
    BatchPlan pb;
    std::vector<uint64_t> byte_base, plane_base;
    for (size_t i : todo) {
      JXLB_CHECK(plans[i].streams.size() == 1, "internal: a probe is exactly one stream");
      byte_base.push_back(pb.bytes.size());
      plane_base.push_back(pb.planes.size());
      MergeFrame(plans[i], views[i].data, views[i].size, &pb);
    }
    pb.bytes.resize(pb.bytes.size() + 64, 0);
    BundleStreams(&pb);
    std::vector<uint64_t> end_bits;
    std::vector<int32_t> arena;
    probe(pb, &end_bits, &arena);
    JXLB_CHECK(end_bits.size() == pb.streams.size() && arena.size() >= pb.arena_size, "internal: bad probe result");
    for (size_t s = 0; s < pb.streams.size(); s++) {
      const uint64_t byte = pb.streams[s].bit_pos / 8;
      const size_t k = std::upper_bound(byte_base.begin(), byte_base.end(), byte) - byte_base.begin() - 1;
      ProbeCtx& c = ctx[todo[k]];
      ProbeResult res;
      res.end_bit = end_bits[s] - byte_base[k] * 8;
      for (uint32_t pl : c.want_planes) {
        const DevPlane& dp = pb.planes[plane_base[k] + pl];
        res.samples.insert(res.samples.end(), arena.begin() + dp.off, arena.begin() + dp.off + static_cast<size_t>(dp.w) * dp.h);
      }
      c.done.push_back(std::move(res));
    }
  }
  const auto t_merge = std::chrono::steady_clock::now();
  {  // one allocation per pool instead of a doubling series
    size_t alias = 0, lut = 0, cpool = 0, tree = 0, chans = 0, planes = 0, streams = 0, acs = 0, prefix = 0, cfg = 0, opool = 0;
    for (const FramePlan& f : plans) {
      alias += f.alias.size();
      lut += f.lut.size();
      cpool += f.v.cpool.size();
      opool += f.v.opool.size();
      tree += f.tree.size();
      chans += f.chans.size();
      planes += f.planes.size();
      streams += f.streams.size();
      acs += f.v.ac_streams.size();
      prefix += f.prefix.size();
      cfg += f.cfg.size();
    }
    batch->alias.reserve(alias);
    batch->lut.reserve(lut);
    batch->cpool.reserve(cpool);
    batch->tree.reserve(tree);
    batch->chans.reserve(chans);
    batch->planes.reserve(planes);
    batch->streams.reserve(streams);
    batch->ac_streams.reserve(acs);
    batch->prefix.reserve(prefix);
    batch->cfg.reserve(cfg);
    const SharedVarDCTTables& sh = SharedVarDCTTables::Get();
    batch->opool.reserve(opool + sh.opool.size());
  }
  for (size_t i = 0; i < n; i++) MergeFrame(plans[i], views[i].data, views[i].size, batch, ext ? static_cast<int64_t>(byte_off[i]) : -1);
  if (!ext) batch->bytes.resize(batch->bytes.size() + 64, 0);  // read-ahead padding for the device bit reader
  BundleStreams(batch);
  if (!batch->vframes.empty()) {
    // pixel-plane slots for one wave of frames, after the per-frame planes of the float arena
    // Two sets of three planes (ping-pong of the per-pixel filter kernels) only when those kernels run: frames
    // without patches go through the fused render tile, which reads set 0 and writes output samples, so a wave holds
    // twice as many frames in the same memory (half as many launches and kernel tails per batch).
    const bool two_sets = !batch->patches.empty() || batch->any_upsampling || std::getenv("JXLB200_UNFUSED_RENDER") != nullptr ||
                          std::getenv("JXLB_EMUL_UNFUSED") != nullptr;
    const uint64_t per_frame = (two_sets ? 6 : 3) * batch->pix_plane_max;
    uint64_t wave_bytes = kWavePixelBytes;
    if (const char* e = std::getenv("JXLB200_WAVE_MB")) wave_bytes = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10)) << 20;  // tuning knob
    batch->wave_frames = static_cast<uint32_t>(
        std::max<uint64_t>(1, std::min<uint64_t>(batch->vframes.size(), wave_bytes / (per_frame * 4))));
    const uint64_t pix_base = batch->farena_size;
    batch->farena_size += per_frame * batch->wave_frames;
    for (size_t i = 0; i < batch->vframes.size(); i++) {
      const uint64_t slot = pix_base + (i % batch->wave_frames) * per_frame;
      for (int set = 0; set < 2; set++)
        for (int c = 0; c < 3; c++)
          batch->vframes[i].pix[set][c] = slot + ((two_sets ? set * 3 : 0) + c) * batch->pix_plane_max;
    }
    JXLB_CHECK(batch->tok_size < (uint64_t{1} << 32), "batch too large: token arena exceeds 2^32 entries");
    // Lanes of a warp run until their longest stream ends: put streams of similar length together, longest first
    // (measured: bundling by frame instead, for L1 locality of the alias tables, is 1.6x slower).
    std::stable_sort(batch->ac_streams.begin(), batch->ac_streams.end(), [](const DevAcStream& x, const DevAcStream& y) {
      return x.bit_end - x.bit_pos > y.bit_end - y.bit_pos;
    });
    // the same streams grouped by (frame, pass), each group still longest first
    std::vector<uint32_t> unit_of(batch->vframes.size() * kMaxPasses, 0xFFFFFFFFu);
    for (const DevAcStream& s : batch->ac_streams) {
      uint32_t& u = unit_of[s.frame * kMaxPasses + s.pass];
      if (u == 0xFFFFFFFFu) {
        u = batch->ac_units.size();
        batch->ac_units.push_back(DevAcUnit{0, 0, s.frame, s.pass});
      }
      batch->ac_units[u].count++;
    }
    uint32_t first = 0;
    for (DevAcUnit& u : batch->ac_units) {
      u.first = first;
      first += u.count;
      u.count = 0;
    }
    batch->ac_unit_streams.resize(batch->ac_streams.size());
    for (uint32_t i = 0; i < batch->ac_streams.size(); i++) {
      DevAcUnit& u = batch->ac_units[unit_of[batch->ac_streams[i].frame * kMaxPasses + batch->ac_streams[i].pass]];
      batch->ac_unit_streams[u.first + u.count++] = i;
    }
  }
  if (timing)
    std::fprintf(stderr, "PlanBatch: %zu files, %d threads: %.1f ms (per-file planning %.1f ms, merge %.1f ms)\n", n, threads,
                 since(t_begin), t_frames, since(t_merge));
}

}  // namespace jxlb

#endif  // JXLB_BATCH_H_
