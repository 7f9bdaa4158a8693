// jxl_b200 device code: the lossless (Modular) encoder -- SURVEY.md 8f N4, jpegxl-rs `lossless(true)`
// (jpegxl-rs/src/encode.rs:143, :230-234 -> JxlEncoderSetFrameLossless).
//
// What libjxl's Modular encoder does per sample (lib/jxl/modular/encoding/enc_encoding.cc:296-520: properties, MA-tree
// leaf, predictor, residual token) under a FIXED tree: libjxl's own fixed gradient tree (enc_encoding.cc:274-282 -- the
// 33 cutoffs on property 9 = W + N - NW, Gradient predictor in every leaf) after the YCoCg-R reversible colour transform
// (lib/jxl/modular/transform/enc_rct.cc:17-68, rct_type 6), groups of kEnclGroupDim x kEnclGroupDim (lib/jxl/enc_modular.cc:1258-1500).
// With a fixed tree the context and the prediction of a sample only need its TRUE neighbours, which the encoder has:
// every token of the image is produced independently at its final slot (thread per sample); only the rANS emission of
// a group is a serial chain (thread per group, DevRansPush). No tree learning, palette or squeeze (libjxl's effort-7
// encoder learns a tree per image; its fixed trees are what it uses at the fast efforts).
// The oracle's EncodeModular(tree = 1, predictor = Gradient, rct = 6) is the byte-exact CPU statement of this file.
#ifndef JXLB_ENCL_DEV_H_
#define JXLB_ENCL_DEV_H_

#include "jxlb_enc_dev.h"
#include "jxlb_encl_const.h"

namespace jxlb {

constexpr uint32_t kEnclCutoffs = 33;

struct DevLFrame {
  uint32_t xsize, ysize, xgroups, ygroups;
  uint32_t nch;         // 1 grey, 2 grey + alpha, 3 RGB, 4 RGBA
  uint32_t bytes;       // per input sample: 1 or 2 (native endian)
  uint64_t in_off;      // byte offset of the interleaved input samples
  uint64_t plane_off;   // int32 index of channel 0's plane; channel c at plane_off + c * xsize * ysize
  uint64_t tok_off;     // uint2 index of the first token of group 0; group g at tok_off + g * nch * kEnclGroupSamples
  uint64_t hist_off;    // uint32 index: [34][256] token counts of the frame
  uint32_t sec_base;    // index of group 0's section in the offset / first-bit arrays
  uint32_t pad_;
  uint64_t code_off[2]; // this frame's code in the fs / reverse tables
};

struct DevLPools {
  const uint8_t* in;
  int32_t* planes;
  uint2* tokens;
  uint32_t* hist;
  const int32_t* cutoffs;   // kEnclCutoffs ascending split values of property 9
  const uint32_t* leaf_of;  // [kEnclCutoffs + 1]: number of cutoffs below the property -> leaf (= context = cluster)
};

JXLB_HD int32_t DevLoadSample(const DevLPools& L, const DevLFrame& f, uint64_t idx) {
  if (f.bytes == 1) return L.in[f.in_off + idx];
  return static_cast<int32_t>(L.in[f.in_off + 2 * idx]) | (static_cast<int32_t>(L.in[f.in_off + 2 * idx + 1]) << 8);
}

// Pixel `i` of the frame: interleaved samples -> planes, the colour channels through YCoCg-R (enc_rct.cc:44-52).
JXLB_HD void DevEnclPlanes(const DevLPools& L, const DevLFrame& f, uint64_t i) {
  const uint64_t n = static_cast<uint64_t>(f.xsize) * f.ysize;
  int32_t* p = L.planes + f.plane_off;
  if (f.nch >= 3) {
    const int32_t R = DevLoadSample(L, f, i * f.nch), G = DevLoadSample(L, f, i * f.nch + 1), B = DevLoadSample(L, f, i * f.nch + 2);
    const int32_t co = R - B;
    const int32_t tmp = B + (co >> 1);
    const int32_t cg = G - tmp;
    p[i] = tmp + (cg >> 1);
    p[n + i] = co;
    p[2 * n + i] = cg;
    if (f.nch == 4) p[3 * n + i] = DevLoadSample(L, f, i * 4 + 3);
  } else {
    for (uint32_t c = 0; c < f.nch; c++) p[c * n + i] = DevLoadSample(L, f, i * f.nch + c);
  }
}

// Sample `i` (0 <= i < nch * xsize * ysize, channel-major) -> its token at its slot of its group, counted.
// Neighbours with the edge rules of context_predict.h:496-504 INSIDE the group's crop of the channel.
JXLB_HD void DevEnclToken(const DevLPools& L, const DevLFrame& f, uint64_t i) {
  const uint64_t n = static_cast<uint64_t>(f.xsize) * f.ysize;
  const uint32_t c = static_cast<uint32_t>(i / n);
  const uint64_t at = i - c * n;
  const uint32_t y = static_cast<uint32_t>(at / f.xsize), x = static_cast<uint32_t>(at - static_cast<uint64_t>(y) * f.xsize);
  const uint32_t gx = x / kEnclGroupDim, gy = y / kEnclGroupDim, lx = x % kEnclGroupDim, ly = y % kEnclGroupDim;
  const uint32_t gw = f.xsize - gx * kEnclGroupDim < kEnclGroupDim ? f.xsize - gx * kEnclGroupDim : kEnclGroupDim;
  const uint32_t gh = f.ysize - gy * kEnclGroupDim < kEnclGroupDim ? f.ysize - gy * kEnclGroupDim : kEnclGroupDim;
  const int32_t* p = L.planes + f.plane_off + c * n;
  const int32_t v = p[at];
  const int32_t left = lx ? p[at - 1] : (ly ? p[at - f.xsize] : 0);
  const int32_t top = ly ? p[at - f.xsize] : left;
  const int32_t topleft = (lx && ly) ? p[at - f.xsize - 1] : left;
  const int32_t prop = left + top - topleft;
  uint32_t below = 0;
  for (uint32_t k = 0; k < kEnclCutoffs; k++) below += prop > L.cutoffs[k] ? 1u : 0u;
  const uint32_t cluster = L.leaf_of[below];
  const int32_t guess = DevClampedGradient(left, top, topleft);
  const int32_t r = v - guess;
  const uint32_t packed = (static_cast<uint32_t>(r) << 1) ^ (r < 0 ? 0xFFFFFFFFu : 0u);
  const uint32_t g = gy * f.xgroups + gx;
  L.tokens[f.tok_off + static_cast<uint64_t>(g) * f.nch * kEnclGroupSamples + static_cast<uint64_t>(c) * gw * gh + ly * gw + lx] =
      make_uint2(cluster, packed);
  DevCountToken(L.hist + f.hist_off, cluster, packed);
}

// Group g's section -- GroupHeader (global tree, default weighted-predictor header, no transforms), then the tokens of
// its channels -- written back to front so that it ends at bit `end_pos`; returns the position of its first bit.
JXLB_HD uint64_t DevEnclEmitGroup(const DevLPools& L, const DevLFrame& f, uint32_t g, const DevEncCode& code, uint32_t* words,
                                  uint64_t end_pos, bool group_header) {
  const uint32_t gx = g % f.xgroups, gy = g / f.xgroups;
  const uint32_t gw = f.xsize - gx * kEnclGroupDim < kEnclGroupDim ? f.xsize - gx * kEnclGroupDim : kEnclGroupDim;
  const uint32_t gh = f.ysize - gy * kEnclGroupDim < kEnclGroupDim ? f.ysize - gy * kEnclGroupDim : kEnclGroupDim;
  DevBackWriter w;
  w.Init(words, end_pos);
  DevRansPush(L.tokens + f.tok_off + static_cast<uint64_t>(g) * f.nch * kEnclGroupSamples, f.nch * gw * gh, code, w);
  if (group_header) w.Put(4, 0x3);
  w.Finish();
  return w.cursor;
}

#if defined(__CUDACC__)
// The same section written by a warp (DevRansPushWarp, jxlb_enc_dev.h); returns the position of the first bit.
__device__ __forceinline__ uint64_t DevEnclEmitGroupWarp(const DevLPools& L, const DevLFrame& f, uint32_t g, const DevEncCode& code,
                                                         uint32_t* words, uint64_t end_pos, bool group_header, uint32_t lane, uint4* stage) {
  const uint32_t gx = g % f.xgroups, gy = g / f.xgroups;
  const uint32_t gw = f.xsize - gx * kEnclGroupDim < kEnclGroupDim ? f.xsize - gx * kEnclGroupDim : kEnclGroupDim;
  const uint32_t gh = f.ysize - gy * kEnclGroupDim < kEnclGroupDim ? f.ysize - gy * kEnclGroupDim : kEnclGroupDim;
  uint64_t cursor = end_pos;
  DevRansPushWarp(L.tokens + f.tok_off + static_cast<uint64_t>(g) * f.nch * kEnclGroupSamples, f.nch * gw * gh, code, words, &cursor, lane, stage);
  if (group_header) DevWarpPut(words, &cursor, 4, 0x3, lane);
  return cursor;
}
#endif

}  // namespace jxlb

#endif  // JXLB_ENCL_DEV_H_
