// TEST INFRASTRUCTURE: runs the __host__ __device__ bodies of the jxl_b200 kernels on
// the CPU so that the planner and the per-stream logic can be checked on a box
// without a GPU. The product library never does this (it fails without CUDA);
// nothing outside tests/ builds or loads this file.
#include <cstdio>
#include <cstring>

#include "../../jpegxl-rs_b200/csrc/host/jxlb_batch.h"
#include "../../jpegxl-rs_b200/csrc/kernels/jxlb_finish_dev.h"

using namespace jxlb;

extern "C" {

// Decodes n files; returns 0 or a negative error. out must hold out_size bytes
// laid out like the batch output buffer (frames 256-byte aligned).
int jxlb_emul_decode(const uint8_t* const* files, const size_t* sizes, size_t n, uint32_t num_channels,
                     uint32_t data_type, uint32_t endianness, size_t align, uint8_t* out, size_t out_cap,
                     uint64_t* frame_offsets, char* err, size_t errlen) {
  try {
    const bool force_wide = (endianness & 0x100) != 0;  // test hook: run the 64-bit predictor path
    endianness &= 0xFF;
    PixelFormat fmt;
    fmt.num_channels = num_channels;
    fmt.data_type = data_type;
    fmt.endianness = endianness;
    fmt.align = align;
    BatchPlan b;
    PlanBatch(files, sizes, n, fmt, 2, &b);
    if (b.out_size > out_cap) throw Error("output buffer too small");
    std::vector<int32_t> arena(b.arena_size + 16, 0);
    const size_t num_warps = (b.streams.size() + 31) / 32;
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
    P.codes = b.codes.data();
    P.arena = arena.data();
    P.wp_scratch = wp.data();
    P.ring = ring.data();
    P.wp_width = b.wp_width;
    P.lz77 = lz.data();
    P.num_streams = b.streams.size();
    P.warp_chans = b.warp_chans.data();
    P.warp_dims_off = b.warp_dims_off.data();
    P.warp_dims = b.warp_dims.data();
    for (uint32_t s = 0; s < b.streams.size(); s++) {
      // same addressing as the kernel: warp = s / 32, lane = s % 32
      const uint32_t warp = s / 32, lane = s % 32;
      DevLaneMem m;
      m.props = props.data() + lane;
      m.props_stride = 32;
      m.divlut = divlut;
      m.ring_w = b.wp_width;
      m.lane_stride = 32;
      m.ring = ring.data() + static_cast<size_t>(warp) * 3 * b.wp_width * 32 + lane;
      m.wp = wp.data() + static_cast<size_t>(warp) * 10 * (b.wp_width + 2) * 32 + lane;
      const uint32_t* dims = b.warp_dims.data() + b.warp_dims_off[warp];
      uint32_t st = force_wide || !b.narrow ? DevDecodeModularStream<int64_t>(P, s, m, dims, b.warp_chans[warp], true)
                                            : DevDecodeModularStream<int32_t>(P, s, m, dims, b.warp_chans[warp], true);
      if (st != 0) throw Error("stream " + std::to_string(s) + " failed with status " + std::to_string(st));
    }
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
      for (uint32_t y = 0; y < fo.ysize; y++)
        for (uint32_t x = 0; x < fo.xsize; x++) DevWritePixel(P, fo, out, x, y);
    }
    return 0;
  } catch (const std::exception& e) {
    std::snprintf(err, errlen, "%s", e.what());
    return -1;
  }
}
}
