// jxl_b200 host side of the lossless (Modular) encoder (kernels/jxlb_encl_dev.h): the fixed tree and its serialisation,
// image / frame headers, the global section, the assembly of the codestream from the emitted group sections.
//   headers          lib/jxl/image_metadata.cc:278-344, lib/jxl/frame_header.cc:206-440
//   MA tree tokens   lib/jxl/modular/encoding/enc_ma.cc (TokenizeTree), contexts lib/jxl/modular/encoding/ma_common.h:13-22
//   global section   lib/jxl/enc_modular.cc:1258-1330 (DC quantisation flag, tree, histograms, global GroupHeader)
//   group split      lib/jxl/enc_modular.cc:1400-1500: channels larger than the group size go to the per-group streams
#ifndef JXLB_ENCL_HOST_H_
#define JXLB_ENCL_HOST_H_

#include <functional>

#include "../kernels/jxlb_encl_const.h"
#include "jxlb_enc_host.h"

namespace jxlb {

static const int32_t kEnclCutoffValues[33] = {-500, -392, -255, -191, -127, -95, -63, -47, -31, -23, -15, -11, -7, -4, -3, -1, 0,
                                              1, 3, 5, 7, 11, 15, 23, 31, 47, 63, 95, 127, 191, 255, 392, 500};

// libjxl's fixed gradient tree (enc_encoding.cc:274-282 through MakeFixedTree, :180-203): a balanced tree over the
// cutoffs of property 9, Gradient predictor in every leaf; serialised breadth-first, leaves numbered in that order.
struct EnclTree {
  std::vector<std::pair<uint32_t, uint32_t>> tokens;  // (context, value) of the tree's own stream
  uint32_t num_leaves = 0;
  uint32_t leaf_of[34];  // number of cutoffs below the property value -> leaf
};

inline EnclTree BuildEnclTree() {
  struct N { int prop = -1; int32_t split = 0; int l = -1, r = -1; };
  std::vector<N> n;
  std::function<int(size_t, size_t)> fixed = [&](size_t begin, size_t end) -> int {
    const int id = static_cast<int>(n.size());
    n.push_back(N());
    if (begin >= end) return id;
    const size_t mid = (begin + end) / 2;
    n[id].prop = 9;
    n[id].split = kEnclCutoffValues[mid];
    const int l = fixed(mid + 1, end);
    const int r = fixed(begin, mid);
    n[id].l = l;
    n[id].r = r;
    return id;
  };
  const int root = fixed(0, 33);
  EnclTree t;
  std::vector<int> queue = {root};
  std::vector<int> leaf_id(n.size(), -1);
  for (size_t k = 0; k < queue.size(); k++) {
    const N& e = n[queue[k]];
    if (e.prop < 0) {
      leaf_id[queue[k]] = static_cast<int>(t.num_leaves++);
      t.tokens.push_back({1, 0});
      t.tokens.push_back({2, 5});  // Gradient
      t.tokens.push_back({3, 0});  // offset
      t.tokens.push_back({4, 0});  // multiplier: shift 0
      t.tokens.push_back({5, 0});  //             odd part 1 (coded as 0)
    } else {
      queue.push_back(e.l);
      queue.push_back(e.r);
      t.tokens.push_back({1, static_cast<uint32_t>(e.prop + 1)});
      t.tokens.push_back({0, PackSignedH(e.split)});
    }
  }
  for (int below = 0; below <= 33; below++) {
    // a value with `below` cutoffs under it: the cutoff itself counts as not below (the tree tests property > split)
    const int64_t v = below == 0 ? int64_t{kEnclCutoffValues[0]} : int64_t{kEnclCutoffValues[below - 1]} + 1;
    int pos = root;
    while (n[pos].prop >= 0) pos = v > n[pos].split ? n[pos].l : n[pos].r;
    t.leaf_of[below] = static_cast<uint32_t>(leaf_id[pos]);
  }
  return t;
}

struct EnclParams {
  uint32_t xsize = 0, ysize = 0;
  uint32_t nch = 3;   // 1 grey, 2 grey + alpha, 3 RGB, 4 RGBA
  uint32_t bits = 8;  // 8 or 16
  bool Alpha() const { return nch == 2 || nch == 4; }
  uint32_t NumColor() const { return nch < 3 ? 1 : 3; }
  // every channel fits one group: the whole image travels in the global stream, there are no group streams
  bool GlobalOnly() const { return xsize <= kEnclGroupDim && ysize <= kEnclGroupDim; }
  uint32_t XGroups() const { return (xsize + kEnclGroupDim - 1) / kEnclGroupDim; }
  uint32_t YGroups() const { return (ysize + kEnclGroupDim - 1) / kEnclGroupDim; }
};

inline void WriteEnclImageHeaders(BitWriter& w, const EnclParams& p) {
  w.Write(16, 0x0AFF);
  w.Write(1, 0);  // SizeHeader: not "small"
  WriteU32(w, p.ysize, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  w.Write(3, 0);
  WriteU32(w, p.xsize, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  // ImageMetadata
  w.Write(1, 0);  // not all_default
  w.Write(1, 0);  // no extra fields
  auto bit_depth = [&]() {
    w.Write(1, 0);  // integer samples
    WriteU32(w, p.bits, Val(8), Val(10), Val(12), BitsOffset(6, 1));
  };
  bit_depth();
  w.Write(1, p.bits <= 12 ? 1 : 0);  // modular_16_bit_buffer_sufficient
  WriteU32(w, p.Alpha() ? 1 : 0, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(12, 1));
  if (p.Alpha()) {
    if (p.bits == 8) {
      w.Write(1, 1);  // ExtraChannelInfo all_default: 8-bit alpha
    } else {
      w.Write(1, 0);
      WriteU32(w, kAlpha, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(6, 18));
      bit_depth();
      WriteU32(w, 0, Val(0), Val(3), Val(4), BitsOffset(3, 1));               // dim_shift
      WriteU32(w, 0, Val(0), Bits(4), BitsOffset(5, 16), BitsOffset(10, 48));  // name
      w.Write(1, 0);                                                          // not premultiplied
    }
  }
  w.Write(1, 0);  // xyb_encoded = false
  if (p.NumColor() == 3) {
    w.Write(1, 1);  // ColorEncoding all_default (sRGB)
  } else {
    w.Write(1, 0);
    w.Write(1, 0);  // no ICC
    WriteU32(w, kGray, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(6, 18));
    WriteU32(w, 1, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(6, 18));  // white point D65
    w.Write(1, 0);                                                        // no gamma
    WriteU32(w, kTFSRGB, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(6, 18));
    WriteU32(w, 1, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(6, 18));  // rendering intent: relative
  }
  WriteU64(w, 0);  // extensions
  w.Write(1, 1);   // CustomTransformData all_default
  w.ZeroPadToByte();
}

inline void WriteEnclFrameHeader(BitWriter& w, const EnclParams& p) {
  w.Write(1, 0);   // not all_default
  w.Write(2, 0);   // regular frame
  w.Write(1, 1);   // Modular
  WriteU64(w, 0);  // flags
  w.Write(1, 0);   // no YCbCr
  WriteU32(w, 1, Val(1), Val(2), Val(4), Val(8));  // upsampling
  if (p.Alpha()) WriteU32(w, 1, Val(1), Val(2), Val(4), Val(8));
  w.Write(2, kEnclGroupShift);  // group_size_shift (kernels/jxlb_encl_const.h)
  WriteU32(w, 1, Val(1), Val(2), Val(3), BitsOffset(3, 4));  // one pass
  w.Write(1, 0);                                             // no custom size or origin
  WriteU32(w, 0, Val(0), Val(1), Val(2), BitsOffset(2, 3));  // blend mode kReplace
  if (p.Alpha()) WriteU32(w, 0, Val(0), Val(1), Val(2), BitsOffset(2, 3));
  w.Write(1, 1);                                                          // is_last
  WriteU32(w, 0, Val(0), Bits(4), BitsOffset(5, 16), BitsOffset(10, 48));  // name
  w.Write(1, 0);   // LoopFilter not all_default
  w.Write(1, 0);   // no Gaborish
  w.Write(2, 0);   // no EPF
  WriteU64(w, 0);  // loop-filter extensions
  WriteU64(w, 0);  // frame-header extensions
}

// The global section up to (and including) the global GroupHeader: DC quantisation default, the tree, the code of the
// samples (`hist`: [34][256] counts from the device), the header with the RCT. `code` receives the encoder tables.
inline void WriteEnclGlobal(BitWriter& w, const EnclParams& p, const EnclTree& tree, const uint32_t* hist, EncCode* code) {
  w.Write(1, 1);  // default DC quantisation factors
  w.Write(1, 1);  // global MA tree
  std::vector<uint8_t> tree_clusters = {0, 1, 2, 3, 4, 5};
  WriteHostStream(w, 6, tree_clusters, 6, tree.tokens);
  std::vector<uint8_t> leaf_clusters(tree.num_leaves);
  for (size_t i = 0; i < leaf_clusters.size(); i++) leaf_clusters[i] = static_cast<uint8_t>(i % 250);
  WriteCodeHeader(w, leaf_clusters, tree.num_leaves, hist, code);
  // GroupHeader: global tree, default weighted-predictor header, the transforms
  w.Write(1, 1);
  w.Write(1, 1);
  const bool rct = p.NumColor() == 3;
  WriteU32(w, rct ? 1 : 0, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(8, 18));
  if (rct) {
    w.Write(2, kRCT);
    WriteU32(w, 0, Bits(3), BitsOffset(6, 8), BitsOffset(10, 72), BitsOffset(13, 1096));  // begin_c
    WriteU32(w, 6, Val(6), Bits(2), BitsOffset(4, 2), BitsOffset(6, 10));                  // YCoCg-R
  }
}

// `groups`: the emitted sections (one per group; for a GlobalOnly image the single token stream of the global section).
inline std::vector<uint8_t> AssembleEncl(const EnclParams& p, const BitWriter& global, const std::vector<EncSection>& groups) {
  std::vector<std::vector<uint8_t>> sections;
  if (p.GlobalOnly()) {
    BitWriter all;
    all.AppendBits(global.Bytes().data(), global.BitsWritten());
    all.AppendWordBits(groups[0].words, groups[0].first, groups[0].nbits);
    all.ZeroPadToByte();
    sections.push_back(all.Bytes());
  } else {
    BitWriter g = global;
    g.ZeroPadToByte();
    sections.push_back(g.Bytes());
    const uint32_t num_groups = p.XGroups() * p.YGroups();
    const uint32_t num_dc_groups = ((p.xsize + 8 * kEnclGroupDim - 1) / (8 * kEnclGroupDim)) * ((p.ysize + 8 * kEnclGroupDim - 1) / (8 * kEnclGroupDim));
    for (uint32_t i = 0; i < num_dc_groups; i++) sections.push_back({});  // no channel is shifted by 3 or more
    sections.push_back({});                                               // AC global: empty for Modular frames
    for (uint32_t i = 0; i < num_groups; i++) {
      BitWriter w;
      w.AppendWordBits(groups[i].words, groups[i].first, groups[i].nbits);
      w.ZeroPadToByte();
      sections.push_back(w.Bytes());
    }
  }
  BitWriter out;
  WriteEnclImageHeaders(out, p);
  WriteEnclFrameHeader(out, p);
  std::vector<size_t> sizes;
  for (const auto& s : sections) sizes.push_back(s.size());
  WriteToc(out, sizes);
  for (const auto& s : sections) out.AppendBytes(s.data(), s.size());
  return out.Bytes();
}

// Words reserved for a stream of `tokens` tokens: 12 bits of rANS state per symbol at most, the extra bits of a
// 17-bit residual under HybridUintConfig(4, 2, 0) (at most 15), the final state and the group header.
inline uint64_t EnclSectionWords(uint64_t tokens) { return (tokens * 27 + 36 + 31) / 32 + 2; }

}  // namespace jxlb

#endif  // JXLB_ENCL_HOST_H_
