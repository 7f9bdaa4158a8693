// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
//
// Modular sub-bitstream decoding: MA tree, properties, 14 predictors, the
// self-correcting weighted predictor, inverse RCT / palette / squeeze. Restates
//   lib/jxl/modular/encoding/dec_ma.cc:69-139 (tree),
//   lib/jxl/modular/encoding/encoding.cc:143-483 (channel decode; only the
//   generic path is restated, the fast paths there are bit-identical by design),
//   :530-660 (ModularDecode), lib/jxl/modular/encoding/context_predict.h,
//   lib/jxl/modular/transform/{rct,palette,squeeze,transform}.cc, palette.h, squeeze.h.
#ifndef JXLO_MODULAR_H_
#define JXLO_MODULAR_H_

#include <array>
#include <climits>
#include <cstdlib>
#include <vector>

#include "jxlo_entropy.h"

namespace jxlo {

constexpr int kNumStaticProps = 2;
constexpr int kNumNonrefProps = 16;  // 2 static + 13 + 1 WP
constexpr int kWPProp = 15;
constexpr int kNumPredictors = 14;

enum Predictor {
  kPredZero = 0, kPredLeft, kPredTop, kPredAverage0, kPredSelect, kPredGradient, kPredWeighted,
  kPredTopRight, kPredTopLeft, kPredLeftLeft, kPredAverage1, kPredAverage2, kPredAverage3, kPredAverage4
};

struct TreeNode {
  int32_t property;  // -1: leaf
  int32_t splitval;
  uint32_t lchild, rchild;  // leaf: lchild = context id
  uint32_t predictor;
  int64_t offset;
  uint32_t multiplier;
};
using Tree = std::vector<TreeNode>;

struct WPHeader {
  int32_t p1C = 16, p2C = 10, p3Ca = 7, p3Cb = 7, p3Cc = 7, p3Cd = 0, p3Ce = 0;
  uint32_t w[4] = {0xd, 0xc, 0xc, 0xc};
};

inline WPHeader ReadWPHeader(BitReader& br) {
  WPHeader h;
  if (br.ReadBool()) return h;
  h.p1C = br.Read(5);
  h.p2C = br.Read(5);
  h.p3Ca = br.Read(5);
  h.p3Cb = br.Read(5);
  h.p3Cc = br.Read(5);
  h.p3Cd = br.Read(5);
  h.p3Ce = br.Read(5);
  for (auto& w : h.w) w = br.Read(4);
  return h;
}

// lib/jxl/modular/encoding/dec_ma.cc:69-139
inline void ReadTree(BitReader& br, Tree* tree, size_t size_limit) {
  EntropyCode code;
  ReadEntropyCode(br, 6, &code);
  JXLO_CHECK(code.degenerate[code.ctx_map[1]] <= 0, "infinite tree");
  SymbolReader reader(&code, br);
  size_t leaf_id = 0, to_decode = 1;
  tree->clear();
  size_limit = std::min<size_t>(size_limit, size_t{1} << 22);
  while (to_decode > 0) {
    br.CheckInBounds();
    JXLO_CHECK(tree->size() <= size_limit, "tree too large");
    to_decode--;
    uint32_t prop1 = reader.ReadUint(1, br);
    JXLO_CHECK(prop1 <= 256, "bad tree property");
    int property = static_cast<int>(prop1) - 1;
    TreeNode n{};
    if (property == -1) {
      n.property = -1;
      n.predictor = reader.ReadUint(2, br);
      JXLO_CHECK(n.predictor < kNumPredictors, "bad predictor");
      n.offset = UnpackSigned(reader.ReadUint(3, br));
      uint32_t mul_log = reader.ReadUint(4, br);
      JXLO_CHECK(mul_log < 31, "bad multiplier log");
      uint32_t mul_bits = reader.ReadUint(5, br);
      JXLO_CHECK(mul_bits < (1u << (31 - mul_log)) - 1, "bad multiplier");
      n.multiplier = (mul_bits + 1) << mul_log;
      n.lchild = leaf_id++;
      tree->push_back(n);
      continue;
    }
    n.property = property;
    n.splitval = UnpackSigned(reader.ReadUint(0, br));
    n.lchild = tree->size() + to_decode + 1;
    n.rchild = tree->size() + to_decode + 2;
    n.multiplier = 1;
    tree->push_back(n);
    to_decode += 2;
  }
  JXLO_CHECK(reader.FinalStateOk(), "tree: bad ANS final state");
  // ValidateTree, lib/jxl/modular/encoding/dec_ma.cc:23-67: the range of every property must stay non-empty on the
  // way down (a split value outside the range its ancestors leave is an error), height at most 2048.
  int num_props = 0;
  for (const TreeNode& n : *tree) num_props = std::max(num_props, n.property + 1);
  std::vector<std::pair<int32_t, int32_t>> ranges(static_cast<size_t>(num_props) * tree->size(), {INT32_MIN, INT32_MAX});
  std::vector<int> height(tree->size(), 0);
  for (size_t i = 0; i < tree->size(); i++) {
    const TreeNode& n = (*tree)[i];
    JXLO_CHECK(height[i] <= 2048, "tree too tall");
    if (n.property == -1) continue;
    height[n.lchild] = height[n.rchild] = height[i] + 1;
    for (int q = 0; q < num_props; q++) {
      const auto cur = ranges[i * num_props + q];
      if (q == n.property) {
        JXLO_CHECK(!(cur.first > n.splitval || cur.second <= n.splitval), "invalid tree");
        ranges[n.lchild * num_props + q] = {n.splitval + 1, cur.second};
        ranges[n.rchild * num_props + q] = {cur.first, n.splitval};
      } else {
        ranges[n.lchild * num_props + q] = ranges[n.rchild * num_props + q] = cur;
      }
    }
  }
}

struct Channel {
  int w = 0, h = 0;
  int hshift = 0, vshift = 0;
  std::vector<int32_t> d;
  Channel() = default;
  Channel(int w_, int h_, int hs = 0, int vs = 0) : w(w_), h(h_), hshift(hs), vshift(vs), d(static_cast<size_t>(w_) * h_, 0) {}
  int32_t* Row(int y) { return d.data() + static_cast<size_t>(y) * w; }
  const int32_t* Row(int y) const { return d.data() + static_cast<size_t>(y) * w; }
  void Resize(int w_, int h_) {
    w = w_;
    h = h_;
    d.assign(static_cast<size_t>(w) * h, 0);
  }
};

struct SqueezeParams {
  bool horizontal = false, in_place = false;
  uint32_t begin_c = 0, num_c = 0;
};

enum TransformId { kRCT = 0, kPalette = 1, kSqueeze = 2 };

struct Transform {
  uint32_t id = kRCT;
  uint32_t begin_c = 0, rct_type = 6, num_c = 3, nb_colors = 256, nb_deltas = 0;
  uint32_t predictor = kPredZero;
  std::vector<SqueezeParams> squeezes;
};

inline Transform ReadTransform(BitReader& br) {
  Transform t;
  t.id = br.Read(2);
  JXLO_CHECK(t.id != 3, "invalid transform id");
  if (t.id == kRCT || t.id == kPalette)
    t.begin_c = ReadU32(br, Bits(3), BitsOffset(6, 8), BitsOffset(10, 72), BitsOffset(13, 1096));
  if (t.id == kRCT) {
    t.rct_type = ReadU32(br, Val(6), Bits(2), BitsOffset(4, 2), BitsOffset(6, 10));
    JXLO_CHECK(t.rct_type < 42, "bad rct type");
  }
  if (t.id == kPalette) {
    t.num_c = ReadU32(br, Val(1), Val(3), Val(4), BitsOffset(13, 1));
    t.nb_colors = ReadU32(br, BitsOffset(8, 0), BitsOffset(10, 256), BitsOffset(12, 1280), BitsOffset(16, 5376));
    t.nb_deltas = ReadU32(br, Val(0), BitsOffset(8, 1), BitsOffset(10, 257), BitsOffset(16, 1281));
    t.predictor = br.Read(4);
    JXLO_CHECK(t.predictor < kNumPredictors, "bad palette predictor");
  }
  if (t.id == kSqueeze) {
    uint32_t n = ReadU32(br, Val(0), BitsOffset(4, 1), BitsOffset(6, 9), BitsOffset(8, 41));
    t.squeezes.resize(n);
    for (auto& s : t.squeezes) {
      s.horizontal = br.ReadBool();
      s.in_place = br.ReadBool();
      s.begin_c = ReadU32(br, Bits(3), BitsOffset(6, 8), BitsOffset(10, 72), BitsOffset(13, 1096));
      s.num_c = ReadU32(br, Val(1), Val(2), Val(3), BitsOffset(4, 4));
    }
  }
  return t;
}

struct ModImage {
  std::vector<Channel> ch;
  std::vector<Transform> transforms;
  int w = 0, h = 0, bitdepth = 8;
  size_t nb_meta = 0;
};

struct GroupHeader {
  bool use_global_tree = false;
  WPHeader wp;
  std::vector<Transform> transforms;
};

inline GroupHeader ReadGroupHeader(BitReader& br) {
  GroupHeader g;
  g.use_global_tree = br.ReadBool();
  g.wp = ReadWPHeader(br);
  uint32_t n = ReadU32(br, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(8, 18));
  g.transforms.resize(n);
  for (auto& t : g.transforms) t = ReadTransform(br);
  return g;
}

// ---------------------------------------------------------------- predictors
inline int32_t ClampedGradient(int32_t n, int32_t w, int32_t l) {
  int32_t m = std::min(n, w), M = std::max(n, w);
  int32_t grad = static_cast<int32_t>(static_cast<uint32_t>(n) + static_cast<uint32_t>(w) - static_cast<uint32_t>(l));
  int32_t g = l < m ? M : grad;
  return l > M ? m : g;
}

struct Neighbors {
  int64_t left, top, topleft, topright, leftleft, toptop, toprightright;
};

// lib/jxl/modular/encoding/context_predict.h:496-504 (edge rules)
inline Neighbors LoadNeighbors(const int32_t* row, const int32_t* prev, const int32_t* prevprev, int x, int y, int w) {
  Neighbors n;
  n.left = x ? row[x - 1] : (y ? prev[x] : 0);
  n.top = y ? prev[x] : n.left;
  n.topleft = (x && y) ? prev[x - 1] : n.left;
  n.topright = (x + 1 < w && y) ? prev[x + 1] : n.top;
  n.leftleft = x > 1 ? row[x - 2] : n.left;
  n.toptop = y > 1 ? prevprev[x] : n.top;
  n.toprightright = (x + 2 < w && y) ? prev[x + 2] : n.topright;
  return n;
}

inline int64_t PredictOne(uint32_t p, const Neighbors& n, int64_t wp_pred) {
  switch (p) {
    case kPredZero: return 0;
    case kPredLeft: return n.left;
    case kPredTop: return n.top;
    case kPredSelect: {
      int64_t pp = n.left + n.top - n.topleft;
      int64_t pa = std::llabs(pp - n.left), pb = std::llabs(pp - n.top);
      return pa < pb ? n.left : n.top;
    }
    case kPredWeighted: return wp_pred;
    case kPredGradient: return ClampedGradient(n.left, n.top, n.topleft);
    case kPredTopLeft: return n.topleft;
    case kPredTopRight: return n.topright;
    case kPredLeftLeft: return n.leftleft;
    case kPredAverage0: return (n.left + n.top) / 2;
    case kPredAverage1: return (n.left + n.topleft) / 2;
    case kPredAverage2: return (n.topleft + n.top) / 2;
    case kPredAverage3: return (n.top + n.topright) / 2;
    case kPredAverage4:
      return (6 * n.top - 2 * n.toptop + 7 * n.left + n.leftleft + n.toprightright + 3 * n.topright + 8) / 16;
    default: return 0;
  }
}

// lib/jxl/modular/encoding/context_predict.h:63-214
struct WPState {
  static constexpr int64_t kExtraBits = 3;
  static constexpr int64_t kRound = ((1 << kExtraBits) >> 1) - 1;
  int64_t prediction[4] = {0, 0, 0, 0};
  int64_t pred = 0;
  std::vector<uint32_t> pred_errors[4];
  std::vector<int32_t> error;
  WPHeader hdr;

  WPState(const WPHeader& h, size_t xsize) : hdr(h) {
    for (auto& p : pred_errors) p.assign((xsize + 2) * 2, 0);
    error.assign((xsize + 2) * 2, 0);
  }

  static uint32_t DivLookup(uint32_t i) { return (1u << 24) / (i + 1); }

  static uint32_t ErrorWeight(uint64_t x, uint32_t maxweight) {
    int shift = static_cast<int>(FloorLog2(x + 1)) - 5;
    if (shift < 0) shift = 0;
    return 4 + ((static_cast<uint64_t>(maxweight) * DivLookup(x >> shift)) >> shift);
  }

  // Returns the prediction; *max_error receives the WP property value.
  int64_t Predict(size_t x, size_t y, size_t xsize, int64_t N, int64_t W, int64_t NE, int64_t NW, int64_t NN,
                  int32_t* max_error) {
    size_t cur_row = (y & 1) ? 0 : (xsize + 2);
    size_t prev_row = (y & 1) ? (xsize + 2) : 0;
    size_t pos_N = prev_row + x;
    size_t pos_NE = x < xsize - 1 ? pos_N + 1 : pos_N;
    size_t pos_NW = x > 0 ? pos_N - 1 : pos_N;
    uint32_t weights[4];
    for (int i = 0; i < 4; i++) {
      uint32_t e = pred_errors[i][pos_N] + pred_errors[i][pos_NE] + pred_errors[i][pos_NW];
      weights[i] = ErrorWeight(e, hdr.w[i]);
    }
    N *= 8; W *= 8; NE *= 8; NW *= 8; NN *= 8;
    int64_t teW = x == 0 ? 0 : error[cur_row + x - 1];
    int64_t teN = error[pos_N], teNW = error[pos_NW], teNE = error[pos_NE];
    int64_t sumWN = teN + teW;
    if (max_error) {
      int64_t p = teW;
      if (std::llabs(teN) > std::llabs(p)) p = teN;
      if (std::llabs(teNW) > std::llabs(p)) p = teNW;
      if (std::llabs(teNE) > std::llabs(p)) p = teNE;
      *max_error = static_cast<int32_t>(p);
    }
    prediction[0] = W + NE - N;
    prediction[1] = N - (((sumWN + teNE) * hdr.p1C) >> 5);
    prediction[2] = W - (((sumWN + teNW) * hdr.p2C) >> 5);
    prediction[3] = N - ((teNW * hdr.p3Ca + teN * hdr.p3Cb + teNE * hdr.p3Cc + (NN - N) * hdr.p3Cd +
                          (NW - W) * hdr.p3Ce) >> 5);
    // weighted average
    uint32_t wsum = weights[0] + weights[1] + weights[2] + weights[3];
    uint32_t log_weight = FloorLog2(wsum);
    wsum = 0;
    for (int i = 0; i < 4; i++) {
      weights[i] >>= log_weight - 4;
      wsum += weights[i];
    }
    int64_t sum = (wsum >> 1) - 1;
    for (int i = 0; i < 4; i++) sum += prediction[i] * weights[i];
    pred = (sum * DivLookup(wsum - 1)) >> 24;
    if (((teN ^ teW) | (teN ^ teNW)) > 0) return (pred + kRound) >> kExtraBits;
    int64_t mx = std::max(W, std::max(NE, N));
    int64_t mn = std::min(W, std::min(NE, N));
    pred = std::max(mn, std::min(mx, pred));
    return (pred + kRound) >> kExtraBits;
  }

  void Update(int64_t val, size_t x, size_t y, size_t xsize) {
    size_t cur_row = (y & 1) ? 0 : (xsize + 2);
    size_t prev_row = (y & 1) ? (xsize + 2) : 0;
    val *= 8;
    error[cur_row + x] = static_cast<int32_t>(pred - val);
    for (int i = 0; i < 4; i++) {
      int64_t err = (std::llabs(prediction[i] - val) + kRound) >> kExtraBits;
      pred_errors[i][cur_row + x] = static_cast<uint32_t>(err);
      pred_errors[i][prev_row + x + 1] += static_cast<uint32_t>(err);
    }
  }
};

// ---------------------------------------------------------------- channel decode
// Generic path of DecodeModularChannelMAANS (lib/jxl/modular/encoding/encoding.cc:143-483).
inline void DecodeChannel(BitReader& br, SymbolReader& reader, const EntropyCode& code, const Tree& tree,
                          const WPHeader& wp_header, int chan, uint32_t stream_id, ModImage* image) {
  Channel& ch = image->ch[chan];
  if (ch.w == 0 || ch.h == 0) return;
  // Which properties / WP does this tree need?
  int max_prop = 0;
  bool uses_wp = false;
  for (const TreeNode& n : tree) {
    if (n.property >= 0) {
      max_prop = std::max(max_prop, n.property + 1);
      if (n.property == kWPProp) uses_wp = true;
    } else if (n.predictor == kPredWeighted) {
      uses_wp = true;
    }
  }
  int num_props = std::max(max_prop, kNumNonrefProps);
  if (num_props > kNumNonrefProps) num_props = kNumNonrefProps + ((num_props - kNumNonrefProps + 3) / 4) * 4;
  std::vector<int32_t> props(num_props, 0);
  // reference channels: earlier channels with identical geometry (context_predict.h:382-413)
  std::vector<int> refs;
  for (int j = chan - 1; j >= 0 && static_cast<int>(refs.size()) * 4 < num_props - kNumNonrefProps; j--) {
    const Channel& r = image->ch[j];
    if (r.w != ch.w || r.h != ch.h || r.hshift != ch.hshift || r.vshift != ch.vshift) continue;
    refs.push_back(j);
  }
  WPState wp(wp_header, ch.w);
  const int w = ch.w;
  for (int y = 0; y < ch.h; y++) {
    int32_t* row = ch.Row(y);
    const int32_t* prev = y ? ch.Row(y - 1) : nullptr;
    const int32_t* prevprev = y > 1 ? ch.Row(y - 2) : nullptr;
    props[0] = chan;
    props[1] = static_cast<int32_t>(stream_id);
    props[2] = y;
    props[9] = 0;
    for (int x = 0; x < w; x++) {
      Neighbors n = LoadNeighbors(row, prev, prevprev, x, y, w);
      props[3] = x;
      props[4] = static_cast<int32_t>(n.top > 0 ? n.top : -n.top);
      props[5] = static_cast<int32_t>(n.left > 0 ? n.left : -n.left);
      props[6] = static_cast<int32_t>(n.top);
      props[7] = static_cast<int32_t>(n.left);
      props[8] = static_cast<int32_t>(n.left - props[9]);
      props[9] = static_cast<int32_t>(n.left + n.top - n.topleft);
      props[10] = static_cast<int32_t>(n.left - n.topleft);
      props[11] = static_cast<int32_t>(n.topleft - n.top);
      props[12] = static_cast<int32_t>(n.top - n.topright);
      props[13] = static_cast<int32_t>(n.top - n.toptop);
      props[14] = static_cast<int32_t>(n.left - n.leftleft);
      int64_t wp_pred = 0;
      if (uses_wp) wp_pred = wp.Predict(x, y, w, n.top, n.left, n.topright, n.topleft, n.toptop, &props[kWPProp]);
      int off = kNumNonrefProps;
      for (int j : refs) {
        const Channel& r = image->ch[j];
        const int32_t* rp = r.Row(y);
        const int32_t* rprev = r.Row(y ? y - 1 : 0);
        int64_t v = rp[x];
        int64_t vleft = x ? rp[x - 1] : 0;
        int64_t vtop = y ? rprev[x] : vleft;
        int64_t vtopleft = (x && y) ? rprev[x - 1] : vleft;
        int64_t vpred = ClampedGradient(vleft, vtop, vtopleft);
        props[off++] = static_cast<int32_t>(std::llabs(v));
        props[off++] = static_cast<int32_t>(v);
        props[off++] = static_cast<int32_t>(std::llabs(v - vpred));
        props[off++] = static_cast<int32_t>(v - vpred);
      }
      // tree walk: property > splitval ? left : right
      size_t pos = 0;
      while (tree[pos].property >= 0) {
        const TreeNode& t = tree[pos];
        pos = props[t.property] > t.splitval ? t.lchild : t.rchild;
      }
      const TreeNode& leaf = tree[pos];
      uint32_t u = reader.ReadUint(leaf.lchild, br);
      int64_t guess = leaf.offset + PredictOne(leaf.predictor, n, wp_pred);
      int64_t val = static_cast<int64_t>(UnpackSigned(u)) * static_cast<int64_t>(leaf.multiplier) + guess;
      row[x] = static_cast<int32_t>(val);
      if (uses_wp) wp.Update(row[x], x, y, w);
    }
  }
  (void)code;
}

// ---------------------------------------------------------------- transforms
inline bool EqualChannels(const ModImage& im, uint32_t c1, uint32_t c2) {
  JXLO_CHECK(c1 <= im.ch.size() && c2 < im.ch.size() && c2 >= c1, "transform: bad channel range");
  JXLO_CHECK(!(c1 < im.nb_meta && c2 >= im.nb_meta), "transform: mixes meta and non-meta channels");
  for (uint32_t c = c1 + 1; c <= c2; c++) {
    if (im.ch[c].w != im.ch[c1].w || im.ch[c].h != im.ch[c1].h || im.ch[c].hshift != im.ch[c1].hshift ||
        im.ch[c].vshift != im.ch[c1].vshift)
      return false;
  }
  return true;
}

constexpr size_t kMaxFirstPreviewSize = 8;

inline void DefaultSqueezeParams(std::vector<SqueezeParams>* params, const ModImage& image) {
  int nb = static_cast<int>(image.ch.size() - image.nb_meta);
  params->clear();
  size_t w = image.ch[image.nb_meta].w, h = image.ch[image.nb_meta].h;
  bool wide = w > h;
  if (nb > 2 && image.ch[image.nb_meta + 1].w == static_cast<int>(w) && image.ch[image.nb_meta + 1].h == static_cast<int>(h)) {
    SqueezeParams p;
    p.horizontal = true;
    p.in_place = false;
    p.begin_c = image.nb_meta + 1;
    p.num_c = 2;
    params->push_back(p);
    p.horizontal = false;
    params->push_back(p);
  }
  SqueezeParams p;
  p.begin_c = image.nb_meta;
  p.num_c = nb;
  p.in_place = true;
  if (!wide && h > kMaxFirstPreviewSize) {
    p.horizontal = false;
    params->push_back(p);
    h = (h + 1) / 2;
  }
  while (w > kMaxFirstPreviewSize || h > kMaxFirstPreviewSize) {
    if (w > kMaxFirstPreviewSize) {
      p.horizontal = true;
      params->push_back(p);
      w = (w + 1) / 2;
    }
    if (h > kMaxFirstPreviewSize) {
      p.horizontal = false;
      params->push_back(p);
      h = (h + 1) / 2;
    }
  }
}

inline void MetaApply(Transform& t, ModImage& im) {
  if (t.id == kRCT) {
    JXLO_CHECK(EqualChannels(im, t.begin_c, t.begin_c + 2), "rct: channels differ");
  } else if (t.id == kPalette) {
    uint32_t end_c = t.begin_c + t.num_c - 1;
    JXLO_CHECK(EqualChannels(im, t.begin_c, end_c), "palette: channels differ");
    size_t nb = t.num_c;
    if (t.begin_c >= im.nb_meta) {
      im.nb_meta++;
    } else {
      JXLO_CHECK(end_c < im.nb_meta, "palette: bad meta range");
      im.nb_meta += 2 - nb;
    }
    im.ch.erase(im.ch.begin() + t.begin_c + 1, im.ch.begin() + end_c + 1);
    Channel pch(t.nb_colors + t.nb_deltas, nb, -1, -1);
    im.ch.insert(im.ch.begin(), pch);
  } else {  // squeeze
    if (t.squeezes.empty()) DefaultSqueezeParams(&t.squeezes, im);
    for (const SqueezeParams& p : t.squeezes) {
      uint32_t beginc = p.begin_c, endc = p.begin_c + p.num_c - 1;
      JXLO_CHECK(beginc < im.ch.size() && endc < im.ch.size() && endc >= beginc, "squeeze: bad channel range");
      if (beginc < im.nb_meta) {
        JXLO_CHECK(endc < im.nb_meta && p.in_place, "squeeze: bad meta squeeze");
        im.nb_meta += p.num_c;
      }
      uint32_t offset = p.in_place ? endc + 1 : im.ch.size();
      for (uint32_t c = beginc; c <= endc; c++) {
        JXLO_CHECK(im.ch[c].hshift <= 30 && im.ch[c].vshift <= 30, "too many squeezes");
        int w = im.ch[c].w, h = im.ch[c].h;
        JXLO_CHECK(w && h, "squeezing empty channel");
        if (p.horizontal) {
          im.ch[c].w = (w + 1) / 2;
          if (im.ch[c].hshift >= 0) im.ch[c].hshift++;
          w = w - (w + 1) / 2;
        } else {
          im.ch[c].h = (h + 1) / 2;
          if (im.ch[c].vshift >= 0) im.ch[c].vshift++;
          h = h - (h + 1) / 2;
        }
        im.ch[c].Resize(im.ch[c].w, im.ch[c].h);
        Channel placeholder(w, h, im.ch[c].hshift, im.ch[c].vshift);
        im.ch.insert(im.ch.begin() + offset + (c - beginc), placeholder);
      }
    }
  }
}

inline void InvRCT(ModImage& im, uint32_t begin_c, uint32_t rct_type) {
  JXLO_CHECK(EqualChannels(im, begin_c, begin_c + 2), "rct: channels differ");
  if (rct_type == 0) return;
  int perm = rct_type / 7, custom = rct_type % 7;
  size_t m = begin_c;
  size_t n = im.ch[m].d.size();
  std::vector<int32_t> o0(n), o1(n), o2(n);
  const int32_t* a = im.ch[m].d.data();
  const int32_t* b = im.ch[m + 1].d.data();
  const int32_t* c = im.ch[m + 2].d.data();
  auto add = [](int32_t x, int32_t y) { return static_cast<int32_t>(static_cast<uint32_t>(x) + static_cast<uint32_t>(y)); };
  int second = custom >> 1, third = custom & 1;
  for (size_t i = 0; i < n; i++) {
    if (custom == 6) {
      int32_t tmp = add(a[i], -(c[i] >> 1));
      int32_t G = add(c[i], tmp);
      int32_t B = add(tmp, -(b[i] >> 1));
      int32_t R = add(B, b[i]);
      o0[i] = R; o1[i] = G; o2[i] = B;
    } else {
      int32_t F = a[i], S = b[i], T = c[i];
      if (third) T = add(T, F);
      if (second == 1) S = add(S, F);
      else if (second == 2) S = add(S, add(F, T) >> 1);
      o0[i] = F; o1[i] = S; o2[i] = T;
    }
  }
  im.ch[m + (perm % 3)].d = o0;
  im.ch[m + ((perm + 1 + perm / 3) % 3)].d = o1;
  im.ch[m + ((perm + 2 - perm / 3) % 3)].d = o2;
}

// lib/jxl/modular/transform/palette.h:54-130
inline int32_t PaletteValue(const Channel& pal, int index, int c, int bit_depth) {
  static const int16_t kDelta[72][3] = {
      {0, 0, 0}, {4, 4, 4}, {11, 0, 0}, {0, 0, -13}, {0, -12, 0}, {-10, -10, -10}, {-18, -18, -18}, {-27, -27, -27},
      {-18, -18, 0}, {0, 0, -32}, {-32, 0, 0}, {-37, -37, -37}, {0, -32, -32}, {24, 24, 45}, {50, 50, 50},
      {-45, -24, -24}, {-24, -45, -45}, {0, -24, -24}, {-34, -34, 0}, {-24, 0, -24}, {-45, -45, -24}, {64, 64, 64},
      {-32, 0, -32}, {0, -32, 0}, {-32, 0, 32}, {-24, -45, -24}, {45, 24, 45}, {24, -24, -45}, {-45, -24, 24},
      {80, 80, 80}, {64, 0, 0}, {0, 0, -64}, {0, -64, -64}, {-24, -24, 45}, {96, 96, 96}, {64, 64, 0}, {45, -24, -24},
      {34, -34, 0}, {112, 112, 112}, {24, -45, -45}, {45, 45, -24}, {0, -32, 32}, {24, -24, 45}, {0, 96, 96},
      {45, -24, 24}, {24, -45, -24}, {-24, -45, 24}, {0, -64, 0}, {96, 0, 0}, {128, 128, 128}, {64, 0, 64},
      {144, 144, 144}, {96, 96, 0}, {-36, -36, 36}, {45, -24, -45}, {45, -45, -24}, {0, 0, -96}, {0, 128, 128},
      {0, 96, 0}, {45, 24, -45}, {-128, 0, 0}, {24, -45, 24}, {-45, 24, -45}, {64, 0, -64}, {64, -64, -64},
      {96, 0, 96}, {45, -45, 24}, {24, 45, -45}, {64, 64, -64}, {128, 128, 0}, {0, 0, -128}, {-24, 45, -45}};
  const int palette_size = pal.w;
  if (index < 0) {
    if (c >= 3) return 0;
    index = -(index + 1);
    index %= 1 + 2 * 71;
    int32_t r = kDelta[(index + 1) >> 1][c] * ((index & 1) ? 1 : -1);
    if (bit_depth > 8) r *= 1 << (bit_depth - 8);
    return r;
  } else if (palette_size <= index && index < palette_size + 64) {
    if (c >= 3) return 0;
    index -= palette_size;
    index >>= c * 2;
    return static_cast<int32_t>((static_cast<uint64_t>(index % 4) * ((uint64_t{1} << bit_depth) - 1)) >> 2) +
           (1 << std::max(0, bit_depth - 3));
  } else if (palette_size + 64 <= index) {
    if (c >= 3) return 0;
    index -= palette_size + 64;
    if (c == 1) index /= 5;
    if (c == 2) index /= 25;
    return static_cast<int32_t>((static_cast<uint64_t>(index % 5) * ((uint64_t{1} << bit_depth) - 1)) >> 2);
  }
  return pal.Row(c)[index];
}

inline void InvPalette(ModImage& im, const Transform& t, const WPHeader& wp_header) {
  JXLO_CHECK(im.nb_meta >= 1, "palette without palette channel");
  int nb = im.ch[0].h;
  uint32_t c0 = t.begin_c + 1;
  JXLO_CHECK(c0 < im.ch.size(), "palette: channel out of range");
  JXLO_CHECK(nb >= 1, "palette: corrupted");
  int w = im.ch[c0].w, h = im.ch[c0].h;
  for (int i = 1; i < nb; i++) im.ch.insert(im.ch.begin() + c0 + 1, Channel(w, h, im.ch[c0].hshift, im.ch[c0].vshift));
  const Channel pal = im.ch[0];
  const int bit_depth = std::min(im.bitdepth, 24);
  if (w == 0) {
  } else if (t.nb_deltas == 0 && t.predictor == kPredZero) {
    for (int y = 0; y < h; y++) {
      for (int x = 0; x < w; x++) {
        int index = im.ch[c0].Row(y)[x];
        if (nb == 1) index = std::min(std::max(index, 0), pal.w - 1);
        for (int c = 0; c < nb; c++) im.ch[c0 + c].Row(y)[x] = PaletteValue(pal, index, c, bit_depth);
      }
    }
  } else {
    Channel indices = im.ch[c0];
    for (int c = 0; c < nb; c++) {
      Channel& ch = im.ch[c0 + c];
      WPState wp(wp_header, ch.w);
      for (int y = 0; y < ch.h; y++) {
        int32_t* row = ch.Row(y);
        const int32_t* prev = y ? ch.Row(y - 1) : nullptr;
        const int32_t* prevprev = y > 1 ? ch.Row(y - 2) : nullptr;
        for (int x = 0; x < ch.w; x++) {
          int index = indices.Row(y)[x];
          int64_t val = PaletteValue(pal, index, c, bit_depth);
          Neighbors n = LoadNeighbors(row, prev, prevprev, x, y, ch.w);
          int64_t wp_pred = 0;
          if (t.predictor == kPredWeighted) wp_pred = wp.Predict(x, y, ch.w, n.top, n.left, n.topright, n.topleft, n.toptop, nullptr);
          if (index < static_cast<int32_t>(t.nb_deltas)) val += PredictOne(t.predictor, n, wp_pred);
          row[x] = static_cast<int32_t>(val);
          if (t.predictor == kPredWeighted) wp.Update(row[x], x, y, ch.w);
        }
      }
    }
  }
  if (c0 >= im.nb_meta) {
    im.nb_meta--;
  } else {
    im.nb_meta -= 2 - nb;
  }
  im.ch.erase(im.ch.begin());
}

inline int64_t SmoothTendency(int64_t B, int64_t a, int64_t n) {
  int64_t diff = 0;
  if (B >= a && a >= n) {
    diff = (4 * B - 3 * n - a + 6) / 12;
    if (diff - (diff & 1) > 2 * (B - a)) diff = 2 * (B - a) + 1;
    if (diff + (diff & 1) > 2 * (a - n)) diff = 2 * (a - n);
  } else if (B <= a && a <= n) {
    diff = (4 * B - 3 * n - a - 6) / 12;
    if (diff + (diff & 1) < 2 * (B - a)) diff = 2 * (B - a) - 1;
    if (diff - (diff & 1) < 2 * (a - n)) diff = 2 * (a - n);
  }
  return diff;
}

inline void InvHSqueeze(ModImage& im, uint32_t c, uint32_t rc) {
  Channel& chin = im.ch[c];
  const Channel& res = im.ch[rc];
  JXLO_CHECK(chin.w == static_cast<int>(DivCeil(chin.w + res.w, 2)) && chin.h == res.h, "hsqueeze: bad dims");
  if (res.w == 0) {
    chin.hshift--;
    return;
  }
  Channel out(chin.w + res.w, chin.h, chin.hshift - 1, chin.vshift);
  if (res.h != 0) {
    for (int y = 0; y < chin.h; y++) {
      const int32_t* pr = res.Row(y);
      const int32_t* pa = chin.Row(y);
      int32_t* po = out.Row(y);
      for (int x = 0; x < res.w; x++) {
        int64_t avg = pa[x];
        int64_t next_avg = x + 1 < chin.w ? pa[x + 1] : avg;
        int64_t left = x ? po[(x << 1) - 1] : avg;
        int64_t diff = pr[x] + SmoothTendency(left, avg, next_avg);
        int64_t A = avg + diff / 2;
        po[x << 1] = static_cast<int32_t>(A);
        po[(x << 1) + 1] = static_cast<int32_t>(A - diff);
      }
      if (out.w & 1) po[out.w - 1] = pa[chin.w - 1];
    }
  }
  im.ch[c] = out;
}

inline void InvVSqueeze(ModImage& im, uint32_t c, uint32_t rc) {
  Channel& chin = im.ch[c];
  const Channel& res = im.ch[rc];
  JXLO_CHECK(chin.h == static_cast<int>(DivCeil(chin.h + res.h, 2)) && chin.w == res.w, "vsqueeze: bad dims");
  if (res.h == 0) {
    chin.vshift--;
    return;
  }
  Channel out(chin.w, chin.h + res.h, chin.hshift, chin.vshift - 1);
  if (res.w != 0) {
    for (int y = 0; y < res.h; y++) {
      const int32_t* pr = res.Row(y);
      const int32_t* pa = chin.Row(y);
      const int32_t* pn = chin.Row(y + 1 < chin.h ? y + 1 : y);
      int32_t* po = out.Row(y << 1);
      int32_t* pno = out.Row((y << 1) + 1);
      const int32_t* pp = y > 0 ? out.Row((y << 1) - 1) : pa;
      for (int x = 0; x < chin.w; x++) {
        int64_t avg = pa[x], next_avg = pn[x], top = pp[x];
        int64_t diff = pr[x] + SmoothTendency(top, avg, next_avg);
        int64_t o = avg + diff / 2;
        po[x] = static_cast<int32_t>(o);
        pno[x] = static_cast<int32_t>(o - diff);
      }
    }
    if (out.h & 1) {
      const int32_t* pa = chin.Row(chin.h - 1);
      int32_t* po = out.Row((chin.h - 1) << 1);
      for (int x = 0; x < chin.w; x++) po[x] = pa[x];
    }
  }
  im.ch[c] = out;
}

inline void InvSqueeze(ModImage& im, const std::vector<SqueezeParams>& params) {
  for (int i = static_cast<int>(params.size()) - 1; i >= 0; i--) {
    const SqueezeParams& p = params[i];
    uint32_t beginc = p.begin_c, endc = p.begin_c + p.num_c - 1;
    JXLO_CHECK(beginc < im.ch.size() && endc < im.ch.size(), "squeeze: bad range");
    uint32_t offset = p.in_place ? endc + 1 : im.ch.size() + beginc - endc - 1;
    if (beginc < im.nb_meta) {
      JXLO_CHECK(im.nb_meta > p.num_c, "squeeze: bad meta count");
      im.nb_meta -= p.num_c;
    }
    for (uint32_t c = beginc; c <= endc; c++) {
      uint32_t rc = offset + c - beginc;
      JXLO_CHECK(rc < im.ch.size(), "squeeze: residual out of range");
      JXLO_CHECK(im.ch[c].w >= im.ch[rc].w && im.ch[c].h >= im.ch[rc].h, "squeeze: corrupted");
      if (p.horizontal) InvHSqueeze(im, c, rc);
      else InvVSqueeze(im, c, rc);
    }
    im.ch.erase(im.ch.begin() + offset, im.ch.begin() + offset + (endc - beginc + 1));
  }
}

inline void UndoTransforms(ModImage& im, const WPHeader& wp) {
  while (!im.transforms.empty()) {
    Transform t = im.transforms.back();
    im.transforms.pop_back();
    if (t.id == kRCT) InvRCT(im, t.begin_c, t.rct_type);
    else if (t.id == kPalette) InvPalette(im, t, wp);
    else InvSqueeze(im, t.squeezes);
  }
}

struct ModularOptions {
  size_t max_chan_size = 0xFFFFFF;
  size_t group_dim = 0x1FFFFFFF;
};

// ModularDecode (lib/jxl/modular/encoding/encoding.cc:530-660). Returns the header
// it read so that the caller can undo transforms later.
inline GroupHeader ModularDecode(BitReader& br, ModImage& image, uint32_t stream_id, const ModularOptions& opt,
                                 const Tree* global_tree, const EntropyCode* global_code) {
  GroupHeader header;
  if (image.ch.empty()) return header;
  header = ReadGroupHeader(br);
  br.CheckInBounds();
  image.transforms = header.transforms;
  for (Transform& t : image.transforms) MetaApply(t, image);
  size_t nb_channels = image.ch.size();
  size_t num_chans = 0, distance_multiplier = 0;
  for (size_t i = 0; i < nb_channels; i++) {
    const Channel& c = image.ch[i];
    if (!c.w || !c.h) continue;
    if (i >= image.nb_meta && (static_cast<size_t>(c.w) > opt.max_chan_size || static_cast<size_t>(c.h) > opt.max_chan_size)) break;
    distance_multiplier = std::max<size_t>(distance_multiplier, c.w);
    num_chans++;
  }
  if (num_chans == 0) return header;
  Tree tree_storage;
  EntropyCode code_storage;
  const Tree* tree = global_tree;
  const EntropyCode* code = global_code;
  if (!header.use_global_tree) {
    uint64_t max_tree_size = 1024;
    for (size_t i = 0; i < nb_channels; i++) {
      const Channel& c = image.ch[i];
      if (i >= image.nb_meta && (static_cast<size_t>(c.w) > opt.max_chan_size || static_cast<size_t>(c.h) > opt.max_chan_size)) break;
      max_tree_size += static_cast<uint64_t>(c.w) * c.h;
    }
    max_tree_size = std::min<uint64_t>(1 << 20, max_tree_size);
    ReadTree(br, &tree_storage, max_tree_size);
    ReadEntropyCode(br, (tree_storage.size() + 1) / 2, &code_storage);
    tree = &tree_storage;
    code = &code_storage;
  } else {
    JXLO_CHECK(global_tree && global_code && !global_tree->empty(), "no global tree available");
  }
  SymbolReader reader(code, br, distance_multiplier);
  for (size_t i = 0; i < nb_channels; i++) {
    const Channel& c = image.ch[i];
    if (!c.w || !c.h) continue;
    if (i >= image.nb_meta && (static_cast<size_t>(c.w) > opt.max_chan_size || static_cast<size_t>(c.h) > opt.max_chan_size)) break;
    DecodeChannel(br, reader, *code, *tree, header.wp, i, stream_id, &image);
    br.CheckInBounds();
  }
  JXLO_CHECK(reader.FinalStateOk(), "modular: bad ANS final state");
  return header;
}

}  // namespace jxlo

#endif  // JXLO_MODULAR_H_
