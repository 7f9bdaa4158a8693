// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
//
// A plain lossless Modular *encoder*: the stream generator behind the decoder parity tests of the Modular rows that
// the reference's fixtures do not reach (palette, delta palette, squeeze, every predictor / property through fixed
// trees, prefix codes, LZ77) and the CPU statement of the lossless encode row that the CUDA encoder
// is checked against. It writes conforming codestreams the way libjxl does, without libjxl's tree learning:
//   headers              lib/jxl/enc_fields.cc, lib/jxl/image_metadata.cc:278-344, lib/jxl/frame_header.cc:206-440
//   forward RCT          lib/jxl/modular/transform/enc_rct.cc:17-68
//   forward palette      lib/jxl/modular/transform/enc_palette.cc:166-590 (exact colours; delta entries only where
//                        prediction + delta reproduces the pixel, so the result stays lossless)
//   forward squeeze      lib/jxl/modular/transform/enc_squeeze.cc:24-140 (default parameters, squeeze.cc:359-420)
//   channel tokens       lib/jxl/modular/encoding/enc_encoding.cc:296-520 (mirror of DecodeModularChannelMAANS)
//   stream / group split lib/jxl/enc_modular.cc:1258-1500 (global stream, DC-group and AC-group streams by shift)
#ifndef JXLO_ENC_MODULAR_H_
#define JXLO_ENC_MODULAR_H_

#include <array>
#include <map>

#include "jxlo_encode.h"

namespace jxlo {

struct ModularEncodeParams {
  uint32_t orientation = 1;  // ImageMetadata::orientation
  uint32_t splines = 0;      // > 0: this many seeded random splines (frame flag kSplines; three colour channels)
  uint32_t bits = 8;              // integer bits per sample, 1 .. 16
  uint32_t num_color = 3;         // 1 (grey) or 3
  bool alpha = false;             // one alpha extra channel of the same depth
  uint32_t group_size_shift = 1;  // group_dim = 128 << shift
  // 0: a single leaf; 1: fixed tree on the gradient property (9); 2: fixed weighted-predictor tree (property 15);
  // 3: a seeded random tree over properties 0 .. 15 with random predictors and offsets (decoder coverage)
  int tree = 0;
  uint32_t predictor = kPredGradient;  // leaves of trees 0 and 1
  uint32_t seed = 1;
  int rct = -1;                   // -1: none, else rct_type 0 .. 41 on the first three channels
  uint32_t palette_colors = 0;    // > 0: palette over the colour channels (and alpha) when they hold at most this many colours
  uint32_t palette_deltas = 0;    // delta entries in front of the colours (predicted entries, palette.cc:107-170)
  uint32_t palette_predictor = kPredZero;
  bool squeeze = false;           // Haar-like squeeze with the default parameters
  int entropy = 0;                // bit 0: prefix codes, bit 1: LZ77
  uint32_t lz77_min_symbol = 224;
};

// ---------------------------------------------------------------- forward transforms
inline void FwdRCT(ModImage& im, uint32_t begin_c, uint32_t rct_type) {
  if (rct_type == 0) return;
  const int perm = rct_type / 7, custom = rct_type % 7;
  const size_t m = begin_c;
  const size_t n = im.ch[m].d.size();
  const std::vector<int32_t> i0 = im.ch[m + (perm % 3)].d;
  const std::vector<int32_t> i1 = im.ch[m + ((perm + 1 + perm / 3) % 3)].d;
  const std::vector<int32_t> i2 = im.ch[m + ((perm + 2 - perm / 3) % 3)].d;
  auto sub = [](int32_t x, int32_t y) { return static_cast<int32_t>(static_cast<uint32_t>(x) - static_cast<uint32_t>(y)); };
  auto add = [](int32_t x, int32_t y) { return static_cast<int32_t>(static_cast<uint32_t>(x) + static_cast<uint32_t>(y)); };
  const int second = custom >> 1, third = custom & 1;
  for (size_t i = 0; i < n; i++) {
    int32_t a, b, c;
    if (custom == 6) {  // YCoCg-R
      const int32_t R = i0[i], G = i1[i], B = i2[i];
      b = sub(R, B);
      const int32_t tmp = add(B, b >> 1);
      c = sub(G, tmp);
      a = add(tmp, c >> 1);
    } else {
      a = i0[i];
      c = third ? sub(i2[i], i0[i]) : i2[i];
      b = second == 1 ? sub(i1[i], i0[i]) : (second == 2 ? sub(i1[i], add(i0[i], i2[i]) >> 1) : i1[i]);
    }
    im.ch[m].d[i] = a;
    im.ch[m + 1].d[i] = b;
    im.ch[m + 2].d[i] = c;
  }
}

// Forward palette over channels [begin_c, begin_c + num_c). Returns false (image untouched) when the channels hold
// more than `max_colors` distinct colours.
inline bool FwdPalette(ModImage& im, Transform* t, uint32_t max_colors, const WPHeader& wp_header) {
  const uint32_t c0 = t->begin_c, nb = t->num_c;
  const int w = im.ch[c0].w, h = im.ch[c0].h;
  const int bit_depth = std::min(im.bitdepth, 24);
  std::map<std::vector<int32_t>, int> color_index;
  std::vector<std::vector<int32_t>> colors;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      std::vector<int32_t> px(nb);
      for (uint32_t c = 0; c < nb; c++) px[c] = im.ch[c0 + c].Row(y)[x];
      if (color_index.emplace(px, static_cast<int>(colors.size())).second) {
        colors.push_back(px);
        if (colors.size() > max_colors) return false;
      }
    }
  // (sorted: neighbouring indices then hold similar colours, like libjxl's luma ordering)
  std::sort(colors.begin(), colors.end());
  for (size_t i = 0; i < colors.size(); i++) color_index[colors[i]] = static_cast<int>(i);
  const bool delta = t->nb_deltas > 0 || t->predictor != kPredZero;
  // delta entries: the most frequent (pixel - prediction) tuples
  std::vector<std::vector<int32_t>> deltas;
  std::vector<std::vector<int64_t>> preds;  // per pixel and channel, when delta
  if (delta) {
    preds.assign(static_cast<size_t>(w) * h, std::vector<int64_t>(nb, 0));
    for (uint32_t c = 0; c < nb; c++) {
      const Channel& ch = im.ch[c0 + c];
      WPState wp(wp_header, w);
      for (int y = 0; y < h; y++) {
        const int32_t* row = ch.Row(y);
        const int32_t* prev = y ? ch.Row(y - 1) : nullptr;
        const int32_t* prevprev = y > 1 ? ch.Row(y - 2) : nullptr;
        for (int x = 0; x < w; x++) {
          const Neighbors n = LoadNeighbors(row, prev, prevprev, x, y, w);
          int64_t wp_pred = 0;
          if (t->predictor == kPredWeighted) wp_pred = wp.Predict(x, y, w, n.top, n.left, n.topright, n.topleft, n.toptop, nullptr);
          preds[static_cast<size_t>(y) * w + x][c] = PredictOne(t->predictor, n, wp_pred);
          if (t->predictor == kPredWeighted) wp.Update(row[x], x, y, w);
        }
      }
    }
    std::map<std::vector<int32_t>, size_t> freq;
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        std::vector<int32_t> d(nb);
        for (uint32_t c = 0; c < nb; c++)
          d[c] = static_cast<int32_t>(im.ch[c0 + c].Row(y)[x] - preds[static_cast<size_t>(y) * w + x][c]);
        freq[d]++;
      }
    std::vector<std::pair<size_t, std::vector<int32_t>>> by_freq;
    for (auto& kv : freq) by_freq.push_back({kv.second, kv.first});
    std::stable_sort(by_freq.begin(), by_freq.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    for (size_t i = 0; i < by_freq.size() && deltas.size() < t->nb_deltas; i++) deltas.push_back(by_freq[i].second);
    while (deltas.size() < t->nb_deltas) deltas.push_back(std::vector<int32_t>(nb, 0));
  }
  t->nb_colors = static_cast<uint32_t>(colors.size());
  Channel pal(static_cast<int>(t->nb_colors + t->nb_deltas), static_cast<int>(nb), -1, -1);
  for (uint32_t c = 0; c < nb; c++) {
    for (uint32_t k = 0; k < t->nb_deltas; k++) pal.Row(c)[k] = deltas[k][c];
    for (uint32_t k = 0; k < t->nb_colors; k++) pal.Row(c)[t->nb_deltas + k] = colors[k][c];
  }
  std::map<std::vector<int32_t>, int> delta_index;
  for (size_t k = deltas.size(); k-- > 0;) delta_index[deltas[k]] = static_cast<int>(k);
  Channel index(w, h, im.ch[c0].hshift, im.ch[c0].vshift);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      std::vector<int32_t> px(nb);
      for (uint32_t c = 0; c < nb; c++) px[c] = im.ch[c0 + c].Row(y)[x];
      int idx = static_cast<int>(t->nb_deltas) + color_index[px];
      if (delta) {
        std::vector<int32_t> d(nb);
        for (uint32_t c = 0; c < nb; c++) d[c] = static_cast<int32_t>(px[c] - preds[static_cast<size_t>(y) * w + x][c]);
        auto it = delta_index.find(d);
        if (it != delta_index.end()) {
          idx = it->second;
        } else if (nb >= 3) {
          // the implicit delta palette behind negative indices (palette.h:54-84), when the tuple is one of its entries
          for (int k = 0; k < 143; k++) {
            bool same = true;
            for (uint32_t c = 0; c < nb && same; c++) same = PaletteValue(pal, -(k + 1), static_cast<int>(c), bit_depth) == d[c];
            if (same) {
              idx = -(k + 1);
              break;
            }
          }
        }
      }
      index.Row(y)[x] = idx;
    }
  // MetaApply's layout: palette channel first, the index channel replaces the first of the nb channels
  im.ch[c0] = index;
  im.ch.erase(im.ch.begin() + c0 + 1, im.ch.begin() + c0 + nb);
  im.ch.insert(im.ch.begin(), pal);
  im.nb_meta++;
  return true;
}

inline void FwdHSqueeze(const Channel& in, Channel* avg, Channel* res) {
  *avg = Channel((in.w + 1) / 2, in.h, in.hshift + 1, in.vshift);
  *res = Channel(in.w - (in.w + 1) / 2, in.h, in.hshift + 1, in.vshift);
  for (int y = 0; y < in.h; y++) {
    const int32_t* pi = in.Row(y);
    int32_t* pa = avg->Row(y);
    int32_t* pr = res->Row(y);
    for (int x = 0; x < res->w; x++) {
      const int64_t A = pi[2 * x], B = pi[2 * x + 1];
      pa[x] = static_cast<int32_t>(A - (A - B) / 2);
    }
    if (in.w & 1) pa[avg->w - 1] = pi[in.w - 1];
    for (int x = 0; x < res->w; x++) {
      const int64_t A = pi[2 * x], B = pi[2 * x + 1];
      const int64_t a = pa[x], next_avg = x + 1 < avg->w ? pa[x + 1] : a;
      const int64_t left = x ? pi[2 * x - 1] : a;
      pr[x] = static_cast<int32_t>((A - B) - SmoothTendency(left, a, next_avg));
    }
  }
}

inline void FwdVSqueeze(const Channel& in, Channel* avg, Channel* res) {
  *avg = Channel(in.w, (in.h + 1) / 2, in.hshift, in.vshift + 1);
  *res = Channel(in.w, in.h - (in.h + 1) / 2, in.hshift, in.vshift + 1);
  for (int y = 0; y < res->h; y++)
    for (int x = 0; x < in.w; x++) {
      const int64_t A = in.Row(2 * y)[x], B = in.Row(2 * y + 1)[x];
      avg->Row(y)[x] = static_cast<int32_t>(A - (A - B) / 2);
    }
  if (in.h & 1)
    for (int x = 0; x < in.w; x++) avg->Row(avg->h - 1)[x] = in.Row(in.h - 1)[x];
  for (int y = 0; y < res->h; y++)
    for (int x = 0; x < in.w; x++) {
      const int64_t A = in.Row(2 * y)[x], B = in.Row(2 * y + 1)[x];
      const int64_t a = avg->Row(y)[x], next_avg = y + 1 < avg->h ? avg->Row(y + 1)[x] : a;
      const int64_t top = y ? in.Row(2 * y - 1)[x] : a;
      res->Row(y)[x] = static_cast<int32_t>((A - B) - SmoothTendency(top, a, next_avg));
    }
}

inline void FwdSqueeze(ModImage& im, const std::vector<SqueezeParams>& params) {
  for (const SqueezeParams& p : params) {
    const uint32_t beginc = p.begin_c, endc = p.begin_c + p.num_c - 1;
    const uint32_t offset = p.in_place ? endc + 1 : static_cast<uint32_t>(im.ch.size());
    for (uint32_t c = beginc; c <= endc; c++) {
      Channel avg, res;
      if (p.horizontal) {
        FwdHSqueeze(im.ch[c], &avg, &res);
      } else {
        FwdVSqueeze(im.ch[c], &avg, &res);
      }
      im.ch[c] = avg;
      im.ch.insert(im.ch.begin() + offset + (c - beginc), res);
    }
  }
}

// ---------------------------------------------------------------- trees
inline GlobalTree SerialiseTree(const std::vector<EncTreeNode>& n, int root, const std::vector<int32_t>* offsets = nullptr) {
  GlobalTree g;
  std::vector<int> queue = {root};
  for (size_t k = 0; k < queue.size(); k++) {
    const EncTreeNode& e = n[queue[k]];
    TreeNode t{};
    if (e.property < 0) {
      const int32_t off = offsets ? (*offsets)[queue[k]] : 0;
      t.property = -1;
      t.predictor = e.predictor;
      t.multiplier = 1;
      t.offset = off;
      t.lchild = g.num_leaves++;
      g.tokens.push_back({1, 0});
      g.tokens.push_back({2, e.predictor});
      g.tokens.push_back({3, PackSigned(off)});
      g.tokens.push_back({4, 0});
      g.tokens.push_back({5, 0});
    } else {
      t.property = e.property;
      t.splitval = e.splitval;
      t.lchild = queue.size();
      t.rchild = queue.size() + 1;
      t.multiplier = 1;
      queue.push_back(e.left);
      queue.push_back(e.right);
      g.tokens.push_back({1, static_cast<uint32_t>(e.property + 1)});
      g.tokens.push_back({0, PackSigned(e.splitval)});
    }
    g.tree.push_back(t);
  }
  return g;
}

inline GlobalTree BuildModularTree(const ModularEncodeParams& p, int max_value) {
  std::vector<EncTreeNode> n;
  std::vector<int32_t> offsets;
  static const std::vector<int32_t> kCutoffs = {-500, -392, -255, -191, -127, -95, -63, -47, -31, -23, -15, -11, -7, -4, -3, -1, 0,
                                                1, 3, 5, 7, 11, 15, 23, 31, 47, 63, 95, 127, 191, 255, 392, 500};
  int root;
  if (p.tree == 0) {
    n.push_back(EncTreeNode());
    n.back().predictor = p.predictor;
    root = 0;
  } else if (p.tree == 1) {
    root = EncTreeFixed(&n, 9, kCutoffs, 0, kCutoffs.size(), p.predictor);
  } else if (p.tree == 2) {
    root = EncTreeFixed(&n, kWPProp, kCutoffs, 0, kCutoffs.size(), kPredWeighted);
  } else {
    std::mt19937 rng(p.seed * 7919u + 13u);
    auto rnd = [&](uint32_t m) { return static_cast<uint32_t>(rng() % m); };
    // recursive random tree: depth up to 5; splits on any of the 16 non-reference properties, inside the range the
    // ancestors leave for the property (ValidateTree, dec_ma.cc:23-67)
    using Ranges = std::array<std::pair<int64_t, int64_t>, 16>;  // [lo, hi) per property
    std::function<int(int, Ranges)> make = [&](int depth, Ranges ranges) -> int {
      const int id = static_cast<int>(n.size());
      n.push_back(EncTreeNode());
      offsets.push_back(0);
      int prop = -1;
      int64_t lo = 0, hi = 0;
      for (int attempt = 0; attempt < 8 && prop < 0 && depth < 5 && !(depth >= 2 && rnd(4) == 0); attempt++) {
        const int q = static_cast<int>(rnd(16));
        int64_t a, b;  // the values this generator likes for q: [a, b)
        if (q == 0) { a = 0; b = 3; }                         // channel
        else if (q == 1) { a = 0; b = 40; }                   // stream id
        else if (q == 2 || q == 3) { a = 0; b = 200; }        // y, x
        else if (q >= 4 && q <= 7) { a = 0; b = max_value; }  // |N|, |W|, N, W
        else { a = -32; b = 32; }                             // differences, WP error
        a = std::max(a, ranges[q].first);
        b = std::min(b, ranges[q].second - 1);  // (split value < upper end of the range)
        if (a >= b) continue;
        prop = q;
        lo = a;
        hi = b;
      }
      if (prop < 0) {
        n[id].predictor = rnd(kNumPredictors);
        offsets[id] = static_cast<int32_t>(rnd(7)) - 3;
        return id;
      }
      const int32_t split = static_cast<int32_t>(lo + rnd(static_cast<uint32_t>(hi - lo)));
      Ranges left = ranges, right = ranges;
      left[prop].first = split + 1;   // property > split
      right[prop].second = split;     // property <= split: [lo, split) in libjxl's bookkeeping
      const int l = make(depth + 1, left), r = make(depth + 1, right);
      n[id].property = prop;
      n[id].splitval = split;
      n[id].left = l;
      n[id].right = r;
      return id;
    };
    Ranges all;
    all.fill({INT32_MIN, INT32_MAX});
    root = make(0, all);
    return SerialiseTree(n, root, &offsets);
  }
  return SerialiseTree(n, root);
}

// Tokens of one Modular stream under `tree` (generic path: offsets, every predictor, the weighted predictor when the
// tree needs it). Channels with zero size are skipped like the decoder skips them.
inline void TokenizeChannels(const Tree& tree, uint32_t stream_id, const std::vector<const Channel*>& channels,
                             const std::vector<int>& chan_index, const WPHeader& wp_header, std::vector<Token>* toks) {
  bool uses_wp = false;
  for (const TreeNode& n : tree) uses_wp |= n.property == kWPProp || (n.property < 0 && n.predictor == kPredWeighted);
  for (size_t k = 0; k < channels.size(); k++) {
    const Channel& ch = *channels[k];
    if (ch.w == 0 || ch.h == 0) continue;
    std::vector<int32_t> props(kNumNonrefProps, 0);
    WPState wp(wp_header, ch.w);
    const int w = ch.w;
    for (int y = 0; y < ch.h; y++) {
      const int32_t* row = ch.Row(y);
      const int32_t* prev = y ? ch.Row(y - 1) : nullptr;
      const int32_t* prevprev = y > 1 ? ch.Row(y - 2) : nullptr;
      props[0] = chan_index[k];
      props[1] = static_cast<int32_t>(stream_id);
      props[2] = y;
      props[9] = 0;
      for (int x = 0; x < w; x++) {
        const Neighbors n = LoadNeighbors(row, prev, prevprev, x, y, w);
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
        size_t pos = 0;
        while (tree[pos].property >= 0) pos = props[tree[pos].property] > tree[pos].splitval ? tree[pos].lchild : tree[pos].rchild;
        const int64_t guess = tree[pos].offset + PredictOne(tree[pos].predictor, n, wp_pred);
        toks->push_back({tree[pos].lchild, PackSigned(static_cast<int32_t>(row[x] - guess))});
        if (uses_wp) wp.Update(row[x], x, y, w);
      }
    }
  }
}

inline void WriteTransform(BitWriter& w, const Transform& t) {
  w.Write(2, t.id);
  if (t.id == kRCT || t.id == kPalette) WriteU32(w, t.begin_c, Bits(3), BitsOffset(6, 8), BitsOffset(10, 72), BitsOffset(13, 1096));
  if (t.id == kRCT) WriteU32(w, t.rct_type, Val(6), Bits(2), BitsOffset(4, 2), BitsOffset(6, 10));
  if (t.id == kPalette) {
    WriteU32(w, t.num_c, Val(1), Val(3), Val(4), BitsOffset(13, 1));
    WriteU32(w, t.nb_colors, BitsOffset(8, 0), BitsOffset(10, 256), BitsOffset(12, 1280), BitsOffset(16, 5376));
    WriteU32(w, t.nb_deltas, Val(0), BitsOffset(8, 1), BitsOffset(10, 257), BitsOffset(16, 1281));
    w.Write(4, t.predictor);
  }
  if (t.id == kSqueeze) WriteU32(w, 0, Val(0), BitsOffset(4, 1), BitsOffset(6, 9), BitsOffset(8, 41));  // default parameters
}

inline void WriteGroupHeader(BitWriter& w, const std::vector<Transform>& transforms) {
  w.Write(1, 1);  // use_global_tree
  w.Write(1, 1);  // default weighted-predictor header
  WriteU32(w, static_cast<uint32_t>(transforms.size()), Val(0), Val(1), BitsOffset(4, 2), BitsOffset(8, 18));
  for (const Transform& t : transforms) WriteTransform(w, t);
}

inline void WriteModularImageHeaders(BitWriter& w, uint32_t xsize, uint32_t ysize, const ModularEncodeParams& p) {
  w.Write(16, 0x0AFF);
  w.Write(1, 0);  // SizeHeader: not "small"
  WriteU32(w, ysize, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  w.Write(3, 0);
  WriteU32(w, xsize, BitsOffset(9, 1), BitsOffset(13, 1), BitsOffset(18, 1), BitsOffset(30, 1));
  // ImageMetadata
  w.Write(1, 0);  // not all_default
  if (p.orientation == 1) {
    w.Write(1, 0);  // no extra fields
  } else {
    w.Write(1, 1);  // extra_fields: the orientation; no intrinsic size, preview, animation
    w.Write(3, p.orientation - 1);
    w.Write(3, 0);
  }
  auto bit_depth = [&]() {
    w.Write(1, 0);  // integer samples
    WriteU32(w, p.bits, Val(8), Val(10), Val(12), BitsOffset(6, 1));
  };
  bit_depth();
  w.Write(1, p.bits <= 12 ? 1 : 0);  // modular_16_bit_buffer_sufficient
  WriteU32(w, p.alpha ? 1 : 0, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(12, 1));
  if (p.alpha) {
    if (p.bits == 8) {
      w.Write(1, 1);  // ExtraChannelInfo all_default: 8-bit alpha
    } else {
      w.Write(1, 0);
      WriteU32(w, kAlpha, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(6, 18));  // Enum
      bit_depth();
      WriteU32(w, 0, Val(0), Val(3), Val(4), BitsOffset(3, 1));  // dim_shift
      WriteU32(w, 0, Val(0), Bits(4), BitsOffset(5, 16), BitsOffset(10, 48));  // name
      w.Write(1, 0);  // not premultiplied
    }
  }
  w.Write(1, 0);  // xyb_encoded = false
  if (p.num_color == 3) {
    w.Write(1, 1);  // ColorEncoding all_default (sRGB)
  } else {
    w.Write(1, 0);
    w.Write(1, 0);  // no ICC
    WriteU32(w, kGray, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(6, 18));
    WriteU32(w, 1, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(6, 18));  // white point D65
    w.Write(1, 0);  // no gamma
    WriteU32(w, kTFSRGB, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(6, 18));
    WriteU32(w, 1, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(6, 18));  // rendering intent: relative
  }
  if (p.orientation != 1) w.Write(1, 1);  // ToneMapping all_default (extra_fields)
  WriteU64(w, 0);  // extensions
  w.Write(1, 1);   // CustomTransformData all_default
  w.ZeroPadToByte();
}

inline void WriteModularFrameHeader(BitWriter& w, const ModularEncodeParams& p) {
  w.Write(1, 0);  // not all_default
  w.Write(2, kRegularFrame);
  w.Write(1, 1);  // Modular
  WriteU64(w, p.splines ? uint64_t{kFlagSplines} : uint64_t{0});  // flags
  w.Write(1, 0);  // no YCbCr
  WriteU32(w, 1, Val(1), Val(2), Val(4), Val(8));  // upsampling
  if (p.alpha) WriteU32(w, 1, Val(1), Val(2), Val(4), Val(8));
  w.Write(2, p.group_size_shift);
  WriteU32(w, 1, Val(1), Val(2), Val(3), BitsOffset(3, 4));  // one pass
  w.Write(1, 0);  // no custom size or origin
  WriteU32(w, 0, Val(0), Val(1), Val(2), BitsOffset(2, 3));  // blend mode kReplace
  if (p.alpha) WriteU32(w, 0, Val(0), Val(1), Val(2), BitsOffset(2, 3));
  w.Write(1, 1);  // is_last
  WriteU32(w, 0, Val(0), Bits(4), BitsOffset(5, 16), BitsOffset(10, 48));  // name
  w.Write(1, 0);  // LoopFilter not all_default
  w.Write(1, 0);  // no Gaborish
  w.Write(2, 0);  // no EPF
  WriteU64(w, 0);  // loop-filter extensions
  WriteU64(w, 0);  // frame-header extensions
}

// samples: interleaved, `channels` = num_color + alpha per pixel, uint16 each (values below 1 << bits).
inline std::vector<uint8_t> EncodeModular(const uint16_t* samples, uint32_t xsize, uint32_t ysize, const ModularEncodeParams& p) {
  JXLO_CHECK(xsize > 0 && ysize > 0 && p.bits >= 1 && p.bits <= 16, "bad image");
  JXLO_CHECK(p.num_color == 1 || p.num_color == 3, "bad channel count");
  const uint32_t nch = p.num_color + (p.alpha ? 1 : 0);
  FrameHeader fh;
  fh.is_modular = true;
  fh.group_size_shift = p.group_size_shift;
  fh.xsize = xsize;
  fh.ysize = ysize;
  fh.color_transform = kCTNone;
  const FrameDimensions dim = ToFrameDimensions(fh);
  ModImage full;
  full.w = xsize;
  full.h = ysize;
  full.bitdepth = p.bits;
  for (uint32_t c = 0; c < nch; c++) {
    full.ch.emplace_back(xsize, ysize);
    for (uint32_t y = 0; y < ysize; y++)
      for (uint32_t x = 0; x < xsize; x++) full.ch[c].Row(y)[x] = samples[(static_cast<size_t>(y) * xsize + x) * nch + c];
  }
  const WPHeader wp_header;
  std::vector<Transform> transforms;
  if (p.palette_colors > 0) {
    Transform t;
    t.id = kPalette;
    t.begin_c = 0;
    t.num_c = nch;
    t.nb_deltas = p.palette_deltas;
    t.predictor = p.palette_predictor;
    if (FwdPalette(full, &t, p.palette_colors, wp_header)) transforms.push_back(t);
  }
  if (p.rct >= 0 && full.ch.size() - full.nb_meta >= 3) {
    Transform t;
    t.id = kRCT;
    t.begin_c = static_cast<uint32_t>(full.nb_meta);
    t.rct_type = static_cast<uint32_t>(p.rct);
    FwdRCT(full, t.begin_c, t.rct_type);
    transforms.push_back(t);
  }
  if (p.squeeze) {
    Transform t;
    t.id = kSqueeze;
    std::vector<SqueezeParams> params;
    DefaultSqueezeParams(&params, full);
    FwdSqueeze(full, params);
    transforms.push_back(t);
  }
  int max_value = (1 << p.bits) - 1;
  const GlobalTree gtree = BuildModularTree(p, max_value);

  // which channels go where (ModularDecodeGlobal / ModularDecodeGroup, jxlo_frame.h)
  struct Stream {
    uint32_t id;
    std::vector<Channel> crops;          // storage for group crops
    std::vector<const Channel*> chans;
    std::vector<int> index;              // property 0 of each channel
    std::vector<Token> toks;
    uint32_t dist_mult = 0;
    bool present = false;                // a GroupHeader is written
  };
  auto finish_stream = [&](Stream* s) {
    for (const Channel* c : s->chans)
      if (c->w && c->h) s->dist_mult = std::max<uint32_t>(s->dist_mult, c->w);
    TokenizeChannels(gtree.tree, s->id, s->chans, s->index, wp_header, &s->toks);
  };
  Stream global;
  global.id = StreamGlobal();
  global.present = true;
  size_t first_group_chan = full.ch.size();
  for (size_t i = 0; i < full.ch.size(); i++) {
    const Channel& c = full.ch[i];
    if (i >= full.nb_meta && (static_cast<size_t>(c.w) > dim.group_dim || static_cast<size_t>(c.h) > dim.group_dim)) {
      first_group_chan = i;
      break;
    }
    global.chans.push_back(&c);
    global.index.push_back(static_cast<int>(i));
  }
  finish_stream(&global);
  auto group_stream = [&](uint32_t id, size_t x0, size_t y0, size_t xs, size_t ys, int min_shift, int max_shift) {
    Stream s;
    s.id = id;
    for (size_t c = first_group_chan; c < full.ch.size(); c++) {
      const Channel& fc = full.ch[c];
      const int shift = std::min(fc.hshift, fc.vshift);
      if (shift > max_shift || shift < min_shift) continue;
      const int rx = static_cast<int>(x0 >> fc.hshift), ry = static_cast<int>(y0 >> fc.vshift);
      int rw = static_cast<int>(xs >> fc.hshift), rh = static_cast<int>(ys >> fc.vshift);
      if (rx >= fc.w || ry >= fc.h) continue;
      rw = std::min(rw, fc.w - rx);
      rh = std::min(rh, fc.h - ry);
      if (rw <= 0 || rh <= 0) continue;
      Channel crop(rw, rh, fc.hshift, fc.vshift);
      for (int y = 0; y < rh; y++) std::memcpy(crop.Row(y), fc.Row(ry + y) + rx, sizeof(int32_t) * rw);
      s.crops.push_back(std::move(crop));
    }
    s.present = !s.crops.empty();
    return s;
  };
  std::vector<Stream> dc_streams, ac_streams;
  for (size_t g = 0; g < dim.num_dc_groups; g++) {
    const size_t gx = g % dim.xsize_dc_groups, gy = g / dim.xsize_dc_groups;
    dc_streams.push_back(group_stream(StreamModularDC(dim, g), gx * dim.dc_group_dim, gy * dim.dc_group_dim, dim.dc_group_dim,
                                      dim.dc_group_dim, 3, 1000));
  }
  for (size_t g = 0; g < dim.num_groups; g++) {
    const size_t gx = g % dim.xsize_groups, gy = g / dim.xsize_groups;
    ac_streams.push_back(group_stream(StreamModularAC(dim, g, 0), gx * dim.group_dim, gy * dim.group_dim, dim.group_dim,
                                      dim.group_dim, 0, 2));
  }
  for (auto* list : {&dc_streams, &ac_streams})
    for (Stream& s : *list) {
      for (size_t k = 0; k < s.crops.size(); k++) {
        s.chans.push_back(&s.crops[k]);
        s.index.push_back(static_cast<int>(k));
      }
      finish_stream(&s);
    }

  EntropyOptions eopt;
  eopt.use_prefix = (p.entropy & 1) != 0;
  eopt.lz77 = (p.entropy & 2) != 0;
  eopt.lz77_min_symbol = p.lz77_min_symbol;
  EntropyEncoder tree_code(6, {0, 1, 2, 3, 4, 5}, eopt);
  tree_code.Count(gtree.tokens);
  std::vector<uint8_t> leaf_clusters(gtree.num_leaves);
  // (at most 255 clusters besides the LZ77 distance cluster: fold the leaves of large trees)
  for (size_t i = 0; i < gtree.num_leaves; i++) leaf_clusters[i] = static_cast<uint8_t>(i % 250);
  EntropyEncoder code(gtree.num_leaves, leaf_clusters, eopt);
  code.Count(global.toks, global.dist_mult);
  for (auto* list : {&dc_streams, &ac_streams})
    for (Stream& s : *list) code.Count(s.toks, s.dist_mult);

  BitWriter dc_global;
  if (p.splines) WriteRandomSplines(dc_global, p.splines, xsize, ysize, p.seed, eopt);
  dc_global.Write(1, 1);  // default DC quantisation factors
  dc_global.Write(1, 1);  // global MA tree
  tree_code.WriteHeader(dc_global);
  tree_code.WriteTokens(dc_global, gtree.tokens);
  code.WriteHeader(dc_global);
  WriteGroupHeader(dc_global, transforms);
  if (!global.toks.empty()) code.WriteTokens(dc_global, global.toks, global.dist_mult);
  auto write_group = [&](BitWriter& w, const Stream& s) {
    if (!s.present) return;
    WriteGroupHeader(w, {});
    if (!s.toks.empty()) code.WriteTokens(w, s.toks, s.dist_mult);
  };
  std::vector<BitWriter> dc_groups(dim.num_dc_groups), ac_groups(dim.num_groups);
  for (size_t g = 0; g < dim.num_dc_groups; g++) write_group(dc_groups[g], dc_streams[g]);
  for (size_t g = 0; g < dim.num_groups; g++) write_group(ac_groups[g], ac_streams[g]);
  BitWriter ac_global;  // (empty for Modular frames)

  std::vector<std::vector<uint8_t>> sections;
  auto finish = [&](BitWriter& w) {
    w.ZeroPadToByte();
    sections.push_back(w.Bytes());
  };
  if (dim.num_groups == 1) {
    BitWriter all = dc_global;
    auto append_bits = [&](BitWriter& src) {
      const size_t n = src.BitsWritten();
      const std::vector<uint8_t>& b = src.Bytes();
      for (size_t i = 0; i < n; i++) all.Write(1, (b[i >> 3] >> (i & 7)) & 1);
    };
    append_bits(dc_groups[0]);
    append_bits(ac_global);
    append_bits(ac_groups[0]);
    finish(all);
  } else {
    finish(dc_global);
    for (auto& w : dc_groups) finish(w);
    finish(ac_global);
    for (auto& w : ac_groups) finish(w);
  }
  BitWriter out;
  WriteModularImageHeaders(out, xsize, ysize, p);
  WriteModularFrameHeader(out, p);
  WriteToc(out, sections);
  for (const auto& s : sections) out.Append(s);
  return out.Bytes();
}

}  // namespace jxlo

#endif  // JXLO_ENC_MODULAR_H_
