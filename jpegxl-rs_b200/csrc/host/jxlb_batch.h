// jxl_b200: merges per-frame plans into the flat pools of one batch.
#ifndef JXLB_BATCH_H_
#define JXLB_BATCH_H_

#include <algorithm>
#include <atomic>
#include <thread>
#include <tuple>

#include "jxlb_frame_plan.h"

namespace jxlb {

struct BatchPlan {
  std::vector<uint8_t> bytes;  // all codestreams, each starting on a 4-byte boundary, zero padded
  std::vector<DevAlias> alias;
  std::vector<uint32_t> prefix, cfg, refs;
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
  bool narrow = true;  // every frame promises that 16-bit buffers suffice -> 32-bit predictor math
};

// Orders the streams so that the 32 lanes of a warp decode planes of the same shape
// (they run one loop nest in lock step) and derives the warp-uniform loop bounds.
inline void BundleStreams(BatchPlan* b) {
  auto key = [&](const DevStream& s) {
    const uint32_t n = s.chan_end - s.chan_begin;
    uint32_t w = 0, h = 0;
    if (n) {
      const DevPlane& p = b->planes[b->chans[s.chan_end - 1].plane];
      w = p.w;
      h = p.h;
    }
    return std::make_tuple(n, w, h);
  };
  std::stable_sort(b->streams.begin(), b->streams.end(),
                   [&](const DevStream& x, const DevStream& y) { return key(x) > key(y); });
  const size_t num_warps = (b->streams.size() + 31) / 32;
  b->warp_chans.assign(num_warps, 0);
  b->warp_dims_off.assign(num_warps, 0);
  b->warp_dims.clear();
  for (size_t wi = 0; wi < num_warps; wi++) {
    uint32_t nmax = 0;
    const size_t s0 = wi * 32, s1 = std::min(b->streams.size(), s0 + 32);
    for (size_t s = s0; s < s1; s++) nmax = std::max(nmax, b->streams[s].chan_end - b->streams[s].chan_begin);
    b->warp_chans[wi] = nmax;
    b->warp_dims_off[wi] = b->warp_dims.size();
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
}

inline void MergeFrame(const FramePlan& f, const uint8_t* cs, size_t cs_size, BatchPlan* b) {
  const uint64_t byte_base = b->bytes.size();
  b->bytes.insert(b->bytes.end(), cs, cs + cs_size);
  b->bytes.resize((b->bytes.size() + 3) & ~size_t{3}, 0);
  b->compressed_bytes += cs_size;
  const uint32_t alias0 = b->alias.size(), prefix0 = b->prefix.size(), cfg0 = b->cfg.size(), refs0 = b->refs.size();
  const uint32_t tree0 = b->tree.size(), codes0 = b->codes.size(), chans0 = b->chans.size();
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
    b->chans.push_back(c);
  }
  for (DevStream s : f.streams) {
    s.bit_pos += byte_base * 8;
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
  for (DevProgram p : f.group_programs) {
    p.op_begin += ops0;
    p.op_end += ops0;
    b->group_programs.push_back(p);
  }
  if (b->levels.size() < f.frame_levels.size()) b->levels.resize(f.frame_levels.size());
  for (size_t k = 0; k < f.frame_levels.size(); k++) {
    DevProgram p = f.frame_levels[k];
    p.op_begin += ops0;
    p.op_end += ops0;
    b->levels[k].push_back(p);
  }
  DevFrameOut fo = f.out;
  for (uint32_t c = 0; c < 4; c++)
    if (c < fo.num_channels && fo.plane[c] != kNoPlane) fo.plane[c] += planes0;
  fo.out_off = b->out_size;
  const uint64_t osize = fo.stride * fo.ysize;
  b->out_size += (osize + 255) & ~uint64_t{255};
  b->frame_out_size.push_back(osize);
  b->frames.push_back(fo);
  b->arena_size += f.arena_size;
  b->wp_slots += f.wp_slots;
  b->wp_width = std::max(b->wp_width, f.wp_width);
  b->lz77_slots += f.lz77_slots;
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

// Plans `n` files on `threads` host threads and merges them in input order.
inline void PlanBatch(const uint8_t* const* files, const size_t* sizes, size_t n, const PixelFormat& fmt,
                      int threads, BatchPlan* batch) {
  std::vector<FramePlan> plans(n);
  std::vector<CodestreamView> views(n);
  std::vector<std::string> errors(n);
  std::atomic<size_t> next{0};
  auto work = [&]() {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= n) break;
      try {
        views[i] = FindCodestream(files[i], sizes[i]);
        PlanCodestream(views[i].data, views[i].size, fmt, &plans[i]);
      } catch (const std::exception& e) {
        errors[i] = e.what();
        if (errors[i].empty()) errors[i] = "unknown error";
      }
    }
  };
  threads = std::max(1, std::min<int>(threads, static_cast<int>(n)));
  if (threads == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  for (size_t i = 0; i < n; i++) {
    if (!errors[i].empty()) throw Error("frame " + std::to_string(i) + ": " + errors[i]);
  }
  for (size_t i = 0; i < n; i++) MergeFrame(plans[i], views[i].data, views[i].size, batch);
  batch->bytes.resize(batch->bytes.size() + 64, 0);  // read-ahead padding for the device bit reader
  BundleStreams(batch);
}

}  // namespace jxlb

#endif  // JXLB_BATCH_H_
