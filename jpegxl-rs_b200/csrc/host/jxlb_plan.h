// jxl_b200 host planner: turns one JPEG XL file into flat device descriptors.
//
// The host owns the bitstream parse (headers, TOC, histograms, MA trees, group
// headers); every per-sample loop is left to the CUDA kernels. The planner mirrors
// the control flow of
//   lib/jxl/dec_frame.cc:266-339, :478-555, :568-731 (section walk),
//   lib/jxl/dec_modular.cc:179-288 (DecodeGlobalInfo), :301-395 (DecodeGroup),
//   lib/jxl/modular/encoding/encoding.cc:530-660 (ModularDecode header part),
//   lib/jxl/modular/transform/*.cc (MetaApply + the order of the inverse transforms),
// but manipulates channel *descriptors* only: no sample is touched on the host.
#ifndef JXLB_PLAN_H_
#define JXLB_PLAN_H_

#include <algorithm>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "../kernels/jxlb_dev.h"
#include "../kernels/jxlb_vardct_desc.h"
#include "jxlb_headers.h"

namespace jxlb {

constexpr int kHostMaxProps = 16 + 4 * 8;

// ---------------------------------------------------------------- Modular fields
struct WPHeader {
  int32_t p1C = 16, p2C = 10, p3Ca = 7, p3Cb = 7, p3Cc = 7, p3Cd = 0, p3Ce = 0;
  uint32_t w[4] = {0xd, 0xc, 0xc, 0xc};
  void Pack(uint32_t out[3]) const {
    out[0] = p1C | (p2C << 8) | (p3Ca << 16) | (static_cast<uint32_t>(p3Cb) << 24);
    out[1] = p3Cc | (p3Cd << 8) | (p3Ce << 16);
    out[2] = w[0] | (w[1] << 8) | (w[2] << 16) | (w[3] << 24);
  }
};

inline WPHeader ReadWPHeader(BitReader& br) {  // lib/jxl/modular/encoding/context_predict.h:37-61
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

struct SqueezeParams {
  bool horizontal = false, in_place = false;
  uint32_t begin_c = 0, num_c = 0;
};

enum TransformId { kRCT = 0, kPalette = 1, kSqueeze = 2 };

struct Transform {
  uint32_t id = kRCT;
  uint32_t begin_c = 0, rct_type = 6, num_c = 3, nb_colors = 256, nb_deltas = 0, predictor = 0;
  std::vector<SqueezeParams> squeezes;
};

inline Transform ReadTransform(BitReader& br) {  // lib/jxl/modular/transform/transform.h:78-137
  Transform t;
  t.id = br.Read(2);
  JXLB_CHECK(t.id != 3, "invalid transform id");
  if (t.id == kRCT || t.id == kPalette)
    t.begin_c = ReadU32(br, Bits(3), BitsOffset(6, 8), BitsOffset(10, 72), BitsOffset(13, 1096));
  if (t.id == kRCT) {
    t.rct_type = ReadU32(br, Val(6), Bits(2), BitsOffset(4, 2), BitsOffset(6, 10));
    JXLB_CHECK(t.rct_type < 42, "bad rct type");
  }
  if (t.id == kPalette) {
    t.num_c = ReadU32(br, Val(1), Val(3), Val(4), BitsOffset(13, 1));
    t.nb_colors = ReadU32(br, BitsOffset(8, 0), BitsOffset(10, 256), BitsOffset(12, 1280), BitsOffset(16, 5376));
    t.nb_deltas = ReadU32(br, Val(0), BitsOffset(8, 1), BitsOffset(10, 257), BitsOffset(16, 1281));
    t.predictor = br.Read(4);
    JXLB_CHECK(t.predictor < 14, "bad palette predictor");
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

struct GroupHeader {  // lib/jxl/modular/encoding/encoding.h:32-54
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

// ---------------------------------------------------------------- the plan
// Everything one frame contributes to the batch; indices are frame-local and are
// relocated when the frame is merged into the batch.
// What a VarDCT frame adds to its FramePlan (offsets frame-local, relocated by MergeFrame).
struct VarDCTPlan {
  DevVFrame vf;
  std::vector<DevAcStream> ac_streams;
  std::vector<float> fpool;     // frame-specific dequantisation tables
  std::vector<uint16_t> opool;  // frame-specific coefficient orders
  std::vector<uint8_t> cpool;   // AC context maps
  std::vector<uint32_t> upool;  // block context map, order index, DC-group plane lists
  uint64_t farena_size = 0, barena_size = 0, uarena_size = 0, tok_size = 0;
  uint64_t pix_plane = 0;       // floats per padded pixel plane
  std::vector<DevPatch> patches;        // src offsets frame-local (farena)
  std::vector<DevRefFrame> ref_frames;  // reference-only frames decoded before this frame (dst frame-local)
};

// Reference slots of one codestream (lib/jxl/dec_cache.h reference_frames[4]) as far as patches need them:
// frames saved before the colour transform, as three float XYB planes in the frame's part of farena.
struct RefSlot {
  bool valid = false;
  uint32_t w = 0, h = 0;
  uint64_t off[3] = {0, 0, 0};
};

// A sub-stream that starts where an entropy-coded Modular stream ends (single-section frames, lib/jxl/dec_frame.cc:
// 597-677; raw quantisation tables inside the AC global section, lib/jxl/dec_modular.cc:765-812) can only be planned
// once the device has decoded that stream. The planner then stops with a *probe*: a plan that holds just that stream.
// The batch layer runs the probes of all such files in one launch of the Modular decode kernel, hands the end
// positions (and, for tables, the decoded samples) back through ProbeCtx::done and plans the file again.
struct ProbeResult {
  uint64_t end_bit = 0;            // first bit after the stream, relative to the codestream
  std::vector<int32_t> samples;    // the planes the planner asked for, back to back
};
struct ProbeCtx {
  std::vector<ProbeResult> done;   // answers of earlier rounds, in the order the planner asks
  size_t next = 0;
  bool pending = false;
  std::vector<uint32_t> want_planes;  // planes of the pending probe whose samples the planner needs
};
struct ProbePending {};  // thrown by the planner after it turned the plan into a probe

struct FramePlan {
  bool is_vardct = false;
  VarDCTPlan v;
  std::vector<DevAlias> alias;
  std::vector<uint32_t> prefix, cfg, refs;
  std::vector<uint16_t> lut;       // weighted-predictor LUTs of the channels that qualify
  std::vector<DevTreeNode> tree;
  std::vector<DevCode> codes;
  std::vector<DevChannel> chans;
  std::vector<DevStream> streams;  // bit_pos relative to the start of the file
  std::vector<DevPlane> planes;
  std::vector<DevOp> ops;
  std::vector<DevProgram> group_programs;
  std::vector<DevProgram> frame_levels;  // global transforms: one op per level, run after the group programs
  uint32_t chain_slots = 0;              // positions the AC decode kernel hands to chained Modular streams
  std::vector<float> spl_seg;            // splines of the frame (DevFrameOut::spl_*): segments ...
  std::vector<uint32_t> spl_idx;         // ... row offsets and index lists
  // group programs / frame levels from these indices on belong to the extra channels of the VarDCT frame: they run
  // after the second Modular launch (BatchPlan::late_*)
  size_t late_programs0 = static_cast<size_t>(-1), late_levels0 = static_cast<size_t>(-1);
  DevFrameOut out;
  uint64_t arena_size = 0;  // int32 elements
  uint32_t wp_slots = 0, wp_width = 0, lz77_slots = 0;
  // info for the API
  uint32_t xsize = 0, ysize = 0;
  ImageMetadata meta;
  uint64_t pixels = 0;
};

struct HChan {
  int w = 0, h = 0, hshift = 0, vshift = 0;
  uint32_t plane = kNoPlane;
};

struct HImage {
  std::vector<HChan> ch;
  size_t nb_meta = 0;
  int bitdepth = 8;
  std::vector<Transform> transforms;
};

struct HostTree {
  std::shared_ptr<std::vector<DevTreeNode>> nodes;  // full tree, leaves already mapped to clusters
  uint32_t code = 0;
  bool uses_wp = false;
  uint32_t num_props = 16;
  bool valid = false;
  bool lz77 = false;
};

class FramePlanner {
 public:
  explicit FramePlanner(FramePlan* plan) : p_(plan) {}

  uint32_t NewPlane(int w, int h, size_t extra = 0) {
    DevPlane pl;
    pl.off = p_->arena_size;
    pl.w = w;
    pl.h = h;
    p_->arena_size += static_cast<uint64_t>(w) * h + extra;
    p_->arena_size = (p_->arena_size + 3) & ~uint64_t{3};  // keep planes 16-byte aligned
    p_->planes.push_back(pl);
    return static_cast<uint32_t>(p_->planes.size() - 1);
  }

  uint32_t AddCode(const EntropyCode& c) {
    DevCode d{};
    d.alias_off = p_->alias.size();
    d.cfg_off = p_->cfg.size();
    d.prefix_off = p_->prefix.size();
    d.num_clusters = c.num_clusters;
    d.log_alpha_size = c.log_alpha_size;
    d.use_prefix = c.use_prefix;
    d.lz77_enabled = c.lz77_enabled;
    d.lz77_min_symbol = c.lz77_min_symbol;
    d.lz77_min_length = c.lz77_min_length;
    d.lz77_length_cfg = PackCfg(c.lz77_length_cfg.split_exponent, c.lz77_length_cfg.msb_in_token, c.lz77_length_cfg.lsb_in_token);
    d.lz77_dist_cluster = c.lz77_dist_cluster;
    for (const auto& h : c.cfg) p_->cfg.push_back(PackCfg(h.split_exponent, h.msb_in_token, h.lsb_in_token));
    if (c.use_prefix) {
      // per-cluster offsets (relative to prefix_off), then the tables
      size_t head = p_->prefix.size();
      p_->prefix.resize(head + c.num_clusters);
      for (uint32_t k = 0; k < c.num_clusters; k++) {
        p_->prefix[head + k] = static_cast<uint32_t>(p_->prefix.size() - head);
        p_->prefix.insert(p_->prefix.end(), c.prefix[k].t.begin(), c.prefix[k].t.end());
      }
    } else {
      for (const auto& a : c.alias) {
        DevAlias e;
        e.cutoff = a.cutoff;
        e.right_value = a.right_value;
        e.freq0 = a.freq0;
        e.offsets1 = a.offsets1;
        e.freq1_xor_freq0 = a.freq1_xor_freq0;
        p_->alias.push_back(e);
      }
    }
    p_->codes.push_back(d);
    return static_cast<uint32_t>(p_->codes.size() - 1);
  }

  // lib/jxl/modular/encoding/dec_ma.cc:69-139; leaves are stored with the cluster
  // (context map already applied).
  HostTree ReadTreeAndCode(BitReader& br, size_t size_limit) {
    EntropyCode tcode;
    ReadEntropyCode(br, 6, &tcode);
    JXLB_CHECK(tcode.degenerate[tcode.ctx_map[1]] <= 0, "infinite tree");
    SymbolReader reader(&tcode, br);
    struct Node { int32_t prop, splitval; uint32_t l, r, pred; int64_t off; uint32_t mul; };
    std::vector<Node> nodes;
    size_t leaf_id = 0, to_decode = 1;
    size_limit = std::min<size_t>(size_limit, size_t{1} << 22);
    HostTree ht;
    int max_prop = 0;
    while (to_decode > 0) {
      br.CheckInBounds();
      JXLB_CHECK(nodes.size() <= size_limit, "tree too large");
      to_decode--;
      uint32_t prop1 = reader.ReadUint(1, br);
      JXLB_CHECK(prop1 <= 256, "bad tree property");
      Node n{};
      n.prop = static_cast<int>(prop1) - 1;
      if (n.prop == -1) {
        n.pred = reader.ReadUint(2, br);
        JXLB_CHECK(n.pred < 14, "bad predictor");
        n.off = UnpackSigned(reader.ReadUint(3, br));
        uint32_t mul_log = reader.ReadUint(4, br);
        JXLB_CHECK(mul_log < 31, "bad multiplier log");
        uint32_t mul_bits = reader.ReadUint(5, br);
        JXLB_CHECK(mul_bits < (1u << (31 - mul_log)) - 1, "bad multiplier");
        n.mul = (mul_bits + 1) << mul_log;
        n.l = leaf_id++;
        if (n.pred == 6) ht.uses_wp = true;
      } else {
        n.splitval = UnpackSigned(reader.ReadUint(0, br));
        n.l = nodes.size() + to_decode + 1;
        n.r = nodes.size() + to_decode + 2;
        to_decode += 2;
        max_prop = std::max(max_prop, n.prop + 1);
        if (n.prop == 15) ht.uses_wp = true;
      }
      nodes.push_back(n);
    }
    JXLB_CHECK(reader.FinalStateOk(), "tree: bad ANS final state");
    JXLB_CHECK(max_prop <= kHostMaxProps, "MA tree references too many earlier channels");
    ValidateTree(nodes);
    EntropyCode code;
    ReadEntropyCode(br, (nodes.size() + 1) / 2, &code);
    ht.code = AddCode(code);
    ht.nodes = std::make_shared<std::vector<DevTreeNode>>();
    ht.num_props = std::max(max_prop, 16);
    if (ht.num_props > 16) ht.num_props = 16 + ((ht.num_props - 16 + 3) / 4) * 4;
    JXLB_CHECK(ht.num_props <= static_cast<uint32_t>(kHostMaxProps), "MA tree references too many earlier channels");
    for (const Node& n : nodes) {
      DevTreeNode d;
      if (n.prop < 0) {
        d.prop = -1;
        d.a = static_cast<int32_t>(code.ctx_map[n.l] | (n.pred << 16));
        d.b = static_cast<uint32_t>(static_cast<int32_t>(n.off));
        d.c = n.mul;
      } else {
        d.prop = n.prop;
        d.a = n.splitval;
        d.b = n.l;
        d.c = n.r;
      }
      ht.nodes->push_back(d);
    }
    ht.valid = true;
    ht.lz77 = code.lz77_enabled;
    return ht;
  }

  // Resolves the static properties (0 = channel, 1 = stream id) the way FilterTree does
  // (lib/jxl/modular/encoding/encoding.cc:36-138) and appends the remaining tree in
  // breadth-first order (children adjacent) to the tree pool. Returns its offset.
  uint32_t PruneTree(const std::vector<DevTreeNode>& t, int32_t chan, int32_t stream_id, bool* uses_wp) {
    const uint32_t off = p_->tree.size();
    auto resolve = [&](uint32_t i) {
      while (t[i].prop >= 0 && t[i].prop < 2) {
        const int32_t v = t[i].prop == 0 ? chan : stream_id;
        i = v > t[i].a ? t[i].b : t[i].c;
      }
      return i;
    };
    *uses_wp = false;
    std::vector<uint32_t> queue;  // source node of output node k
    queue.push_back(resolve(0));
    for (size_t k = 0; k < queue.size(); k++) {
      DevTreeNode n = t[queue[k]];
      if (n.prop < 0) {
        if ((static_cast<uint32_t>(n.a) >> 16) == 6) *uses_wp = true;
      } else {
        if (n.prop == 15) *uses_wp = true;
        const uint32_t l = resolve(n.b), r = resolve(n.c);
        n.b = queue.size();
        n.c = queue.size() + 1;
        queue.push_back(l);
        queue.push_back(r);
      }
      p_->tree.push_back(n);
    }
    return off;
  }

  // (JXLB200_NO_NW_LUT=1: the generic tree walk for every channel -- comparison runs and the emulation tests)
  static bool NoNwLut() { return std::getenv("JXLB200_NO_NW_LUT") != nullptr; }

  // If the pruned tree at `tree_off` only tests property 15 (weighted predictor max error) and all its leaves are
  // (Weighted, offset 0, multiplier 1), builds the property -> cluster table and marks the channel.
  void TryWpLut(DevChannel* dc, bool has_refs) {
    const DevTreeNode* t = p_->tree.data() + dc->tree_off;
    const size_t n = p_->tree.size() - dc->tree_off;
    int64_t lo = INT32_MAX, hi = INT32_MIN;
    for (size_t i = 0; i < n; i++) {
      if (t[i].prop >= 0) {
        if (t[i].prop != 15) return;
        lo = std::min<int64_t>(lo, t[i].a);
        hi = std::max<int64_t>(hi, t[i].a);
      } else if ((static_cast<uint32_t>(t[i].a) >> 16) != 6 || t[i].b != 0 || t[i].c != 1 ||
                 (static_cast<uint32_t>(t[i].a) & 0xFFFF) > 0xFFFF) {
        return;
      }
    }
    (void)has_refs;
    if (lo > hi) {  // a single leaf
      lo = 0;
      hi = -1;
    }
    const int64_t size = hi - lo + 2;  // values lo .. hi + 1; anything outside behaves like the nearest end
    if (size > 8192) return;
    dc->wp_lut = 1;
    dc->lut_off = p_->lut.size();
    dc->lut_lo = static_cast<int32_t>(lo);
    dc->lut_size = static_cast<uint32_t>(size);
    for (int64_t v = lo; v <= hi + 1; v++) {
      size_t pos = 0;
      while (t[pos].prop >= 0) pos = v > t[pos].a ? t[pos].b : t[pos].c;
      p_->lut.push_back(static_cast<uint16_t>(static_cast<uint32_t>(t[pos].a) & 0xFFFF));
    }
  }

  // If the pruned tree at `tree_off` only tests y (property 2), N (6) and W (7) with at most kNwMaxY / kNwThresholds
  // distinct split values each, and every leaf is (predictor other than Weighted, offset 0, multiplier 1, cluster < 256),
  // writes the (y, N, W) bucket table (layout: kernels/jxlb_modular_dev.h, kNwOff*) and marks the channel.
  // The gradient-property variant (nw_lut = 2): every inner node tests property 9, leaves as for TryNwLut.
  bool TryGradLut(DevChannel* dc) {
    const DevTreeNode* t = p_->tree.data() + dc->tree_off;
    const size_t n = p_->tree.size() - dc->tree_off;
    int64_t lo = INT32_MAX, hi = INT32_MIN;
    bool inner = false;
    for (size_t i = 0; i < n; i++) {
      if (t[i].prop >= 0) {
        if (t[i].prop != 9) return false;
        inner = true;
        lo = std::min<int64_t>(lo, t[i].a);
        hi = std::max<int64_t>(hi, t[i].a);
      } else {
        const uint32_t cluster = static_cast<uint32_t>(t[i].a) & 0xFFFF, predictor = static_cast<uint32_t>(t[i].a) >> 16;
        if (predictor == 6 || predictor > 13 || cluster > 0xFF || t[i].b != 0 || t[i].c != 1) return false;
      }
    }
    if (!inner || hi - lo + 2 > 8192) return false;
    dc->nw_lut = 2;
    dc->lut_off = p_->lut.size();
    dc->lut_lo = static_cast<int32_t>(lo);
    dc->lut_size = static_cast<uint32_t>(hi - lo + 2);
    for (int64_t v = lo; v <= hi + 1; v++) {
      size_t pos = 0;
      while (t[pos].prop >= 0) pos = v > t[pos].a ? t[pos].b : t[pos].c;
      const uint32_t a = static_cast<uint32_t>(t[pos].a);
      p_->lut.push_back(static_cast<uint16_t>((a & 0xFF) | ((a >> 16) << 8)));
    }
    return true;
  }

  void TryNwLut(DevChannel* dc) {
    if (TryGradLut(dc)) return;
    const DevTreeNode* t = p_->tree.data() + dc->tree_off;
    const size_t n = p_->tree.size() - dc->tree_off;
    std::vector<int32_t> thr[3];  // y, N, W
    for (size_t i = 0; i < n; i++) {
      if (t[i].prop >= 0) {
        const int k = t[i].prop == 2 ? 0 : (t[i].prop == 6 ? 1 : (t[i].prop == 7 ? 2 : -1));
        if (k < 0) return;
        thr[k].push_back(t[i].a);
      } else {
        const uint32_t cluster = static_cast<uint32_t>(t[i].a) & 0xFFFF, predictor = static_cast<uint32_t>(t[i].a) >> 16;
        if (predictor == 6 || predictor > 13 || cluster > 0xFF || t[i].b != 0 || t[i].c != 1) return;
      }
    }
    for (auto& v : thr) {
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
    }
    if (thr[0].size() > kNwMaxY || thr[1].size() > kNwThresholds || thr[2].size() > kNwThresholds) return;
    dc->nw_lut = 1;
    dc->lut_off = p_->lut.size();
    auto put32 = [&](int32_t v) {
      p_->lut.push_back(static_cast<uint16_t>(static_cast<uint32_t>(v) & 0xFFFF));
      p_->lut.push_back(static_cast<uint16_t>(static_cast<uint32_t>(v) >> 16));
    };
    p_->lut.push_back(static_cast<uint16_t>(thr[0].size()));
    p_->lut.push_back(0);
    const uint32_t cap[3] = {kNwMaxY, kNwThresholds, kNwThresholds};
    for (int k = 0; k < 3; k++)
      for (uint32_t i = 0; i < cap[k]; i++) put32(i < thr[k].size() ? thr[k][i] : INT32_MAX);
    // a value with b thresholds below it: anything in (thr[b - 1], thr[b]]
    auto rep = [&](int k, uint32_t b) -> int64_t {
      return b == 0 ? (thr[k].empty() ? 0 : static_cast<int64_t>(thr[k][0])) : static_cast<int64_t>(thr[k][b - 1]) + 1;
    };
    for (uint32_t by = 0; by <= thr[0].size(); by++)
      for (uint32_t bn = 0; bn <= kNwThresholds; bn++)
        for (uint32_t bw = 0; bw <= kNwThresholds; bw++) {
          // (buckets above the number of real thresholds cannot occur: the padding is INT32_MAX)
          const int64_t v[3] = {rep(0, by), rep(1, std::min<uint32_t>(bn, thr[1].size())), rep(2, std::min<uint32_t>(bw, thr[2].size()))};
          size_t pos = 0;
          while (t[pos].prop >= 0) {
            const int k = t[pos].prop == 2 ? 0 : (t[pos].prop == 6 ? 1 : 2);
            pos = v[k] > t[pos].a ? t[pos].b : t[pos].c;
          }
          const uint32_t a = static_cast<uint32_t>(t[pos].a);
          p_->lut.push_back(static_cast<uint16_t>((a & 0xFF) | ((a >> 16) << 8)));
        }
  }

  // The warp-cooperative kernel's view of a channel that TryWpLut / TryNwLut marked (DevChannel::coop*): the list of
  // clusters its lanes speculate on and the table re-expressed in lanes.
  void BuildCoopLut(DevChannel* dc) {
    if (dc->wp_lut == 0 && dc->nw_lut == 0) return;
    const bool nw = dc->nw_lut == 1;
    const size_t head = nw ? kNwOffTable : 0;
    const size_t count = nw ? (static_cast<size_t>(p_->lut[dc->lut_off]) + 1) * (kNwThresholds + 1) * (kNwThresholds + 1) : dc->lut_size;
    const uint32_t cmask = dc->wp_lut ? 0xFFFFu : 0xFFu;
    // visiting order: outwards from property value 0 (the frequent contexts of an error / gradient property), table
    // order for the (y, N, W) buckets
    std::vector<size_t> order;
    if (nw) {
      for (size_t i = 0; i < count; i++) order.push_back(i);
    } else {
      const int64_t i0 = std::min<int64_t>(std::max<int64_t>(-static_cast<int64_t>(dc->lut_lo), 0), static_cast<int64_t>(count) - 1);
      for (int64_t d = 0; d < static_cast<int64_t>(count); d++) {
        if (d == 0) {
          order.push_back(i0);
        } else {
          if (i0 + d < static_cast<int64_t>(count)) order.push_back(i0 + d);
          if (i0 - d >= 0) order.push_back(i0 - d);
        }
      }
    }
    std::vector<uint32_t> list;
    for (size_t i : order) {
      const uint32_t c = p_->lut[dc->lut_off + head + i] & cmask;
      if (c > 0xFF) return;  // (a miss carries the cluster in 8 bits)
      if (list.size() < 32 && std::find(list.begin(), list.end(), c) == list.end()) list.push_back(c);
    }
    bool miss = false;
    for (size_t i = 0; i < count; i++)
      miss = miss || std::find(list.begin(), list.end(), p_->lut[dc->lut_off + head + i] & cmask) == list.end();
    dc->coop = miss ? 2 : 1;
    dc->coop_list_off = p_->lut.size();
    p_->lut.push_back(static_cast<uint16_t>(list.size()));
    for (uint32_t c : list) p_->lut.push_back(static_cast<uint16_t>(c));
    dc->coop_lut_off = p_->lut.size();
    for (size_t i = 0; i < head; i++) p_->lut.push_back(p_->lut[dc->lut_off + i]);
    for (size_t i = 0; i < count; i++) {
      const uint32_t e = p_->lut[dc->lut_off + head + i];
      const uint32_t c = e & cmask, pred = dc->wp_lut ? 0u : (e >> 8);
      const auto it = std::find(list.begin(), list.end(), c);
      const uint32_t leaf = it != list.end() ? static_cast<uint32_t>(it - list.begin()) : (0x8000u | c);
      p_->lut.push_back(static_cast<uint16_t>(leaf | (pred << 8)));
    }
    // behind the table: the predictor of each row bucket of a (y, N, W) table / of the whole gradient table when all
    // its leaves agree, else 0xFF (DevCoopChannelRows: rows with one of Zero / Left / Gradient take DevCoopFastRow)
    if (!dc->wp_lut) {
      const size_t per = nw ? (kNwThresholds + 1) * (kNwThresholds + 1) : count;
      for (size_t b0 = 0; b0 < count; b0 += per) {
        uint32_t pr = p_->lut[dc->lut_off + head + b0] >> 8;
        for (size_t i = b0; i < b0 + per; i++)
          if ((p_->lut[dc->lut_off + head + i] >> 8) != pr) pr = 0xFF;
        p_->lut.push_back(static_cast<uint16_t>(pr));
      }
    }
  }

  // lib/jxl/modular/encoding/dec_ma.cc:23-67: property ranges must stay non-empty
  // on the way down, which also bounds the device-side walk.
  template <typename N>
  void ValidateTree(const std::vector<N>& nodes) {
    int num_props = 0;
    for (const N& n : nodes) num_props = std::max(num_props, n.prop + 1);
    if (num_props == 0) return;
    std::vector<std::pair<int32_t, int32_t>> ranges(static_cast<size_t>(num_props) * nodes.size(),
                                                    {INT32_MIN, INT32_MAX});
    std::vector<int> height(nodes.size(), 0);
    for (size_t i = 0; i < nodes.size(); i++) {
      JXLB_CHECK(height[i] <= 2048, "tree too tall");
      if (nodes[i].prop == -1) continue;
      JXLB_CHECK(nodes[i].l < nodes.size() && nodes[i].r < nodes.size(), "tree: child out of range");
      height[nodes[i].l] = height[nodes[i].r] = height[i] + 1;
      for (int q = 0; q < num_props; q++) {
        auto cur = ranges[i * num_props + q];
        if (q == nodes[i].prop) {
          int32_t val = nodes[i].splitval;
          JXLB_CHECK(!(cur.first > val || cur.second <= val), "invalid tree");
          ranges[nodes[i].l * num_props + q] = {val + 1, cur.second};
          ranges[nodes[i].r * num_props + q] = {cur.first, val};
        } else {
          ranges[nodes[i].l * num_props + q] = cur;
          ranges[nodes[i].r * num_props + q] = cur;
        }
      }
    }
  }

  // ---- channel descriptor transforms (no samples) ----
  static bool EqualChannels(const HImage& im, uint32_t c1, uint32_t c2) {
    JXLB_CHECK(c1 <= im.ch.size() && c2 < im.ch.size() && c2 >= c1, "transform: bad channel range");
    JXLB_CHECK(!(c1 < im.nb_meta && c2 >= im.nb_meta), "transform: mixes meta and non-meta channels");
    for (uint32_t c = c1 + 1; c <= c2; c++) {
      if (im.ch[c].w != im.ch[c1].w || im.ch[c].h != im.ch[c1].h || im.ch[c].hshift != im.ch[c1].hshift ||
          im.ch[c].vshift != im.ch[c1].vshift)
        return false;
    }
    return true;
  }

  static void DefaultSqueezeParams(std::vector<SqueezeParams>* params, const HImage& image) {
    int nb = static_cast<int>(image.ch.size() - image.nb_meta);
    params->clear();
    size_t w = image.ch[image.nb_meta].w, h = image.ch[image.nb_meta].h;
    bool wide = w > h;
    if (nb > 2 && image.ch[image.nb_meta + 1].w == static_cast<int>(w) && image.ch[image.nb_meta + 1].h == static_cast<int>(h)) {
      SqueezeParams q;
      q.horizontal = true;
      q.in_place = false;
      q.begin_c = image.nb_meta + 1;
      q.num_c = 2;
      params->push_back(q);
      q.horizontal = false;
      params->push_back(q);
    }
    SqueezeParams q;
    q.begin_c = image.nb_meta;
    q.num_c = nb;
    q.in_place = true;
    if (!wide && h > 8) {
      q.horizontal = false;
      params->push_back(q);
      h = (h + 1) / 2;
    }
    while (w > 8 || h > 8) {
      if (w > 8) {
        q.horizontal = true;
        params->push_back(q);
        w = (w + 1) / 2;
      }
      if (h > 8) {
        q.horizontal = false;
        params->push_back(q);
        h = (h + 1) / 2;
      }
    }
  }

  void MetaApply(Transform& t, HImage& im) {
    if (t.id == kRCT) {
      JXLB_CHECK(EqualChannels(im, t.begin_c, t.begin_c + 2), "rct: channels differ");
    } else if (t.id == kPalette) {
      uint32_t end_c = t.begin_c + t.num_c - 1;
      JXLB_CHECK(EqualChannels(im, t.begin_c, end_c), "palette: channels differ");
      size_t nb = t.num_c;
      if (t.begin_c >= im.nb_meta) {
        im.nb_meta++;
      } else {
        JXLB_CHECK(end_c < im.nb_meta, "palette: bad meta range");
        im.nb_meta += 2 - nb;
      }
      im.ch.erase(im.ch.begin() + t.begin_c + 1, im.ch.begin() + end_c + 1);
      HChan pch;
      pch.w = t.nb_colors + t.nb_deltas;
      pch.h = nb;
      pch.hshift = pch.vshift = -1;
      pch.plane = NewPlane(pch.w, pch.h);
      im.ch.insert(im.ch.begin(), pch);
    } else {
      if (t.squeezes.empty()) DefaultSqueezeParams(&t.squeezes, im);
      for (const SqueezeParams& q : t.squeezes) {
        uint32_t beginc = q.begin_c, endc = q.begin_c + q.num_c - 1;
        JXLB_CHECK(beginc < im.ch.size() && endc < im.ch.size() && endc >= beginc, "squeeze: bad channel range");
        if (beginc < im.nb_meta) {
          JXLB_CHECK(endc < im.nb_meta && q.in_place, "squeeze: bad meta squeeze");
          im.nb_meta += q.num_c;
        }
        uint32_t offset = q.in_place ? endc + 1 : im.ch.size();
        for (uint32_t c = beginc; c <= endc; c++) {
          JXLB_CHECK(im.ch[c].hshift <= 30 && im.ch[c].vshift <= 30, "too many squeezes");
          int w = im.ch[c].w, h = im.ch[c].h;
          JXLB_CHECK(w && h, "squeezing empty channel");
          if (q.horizontal) {
            im.ch[c].w = (w + 1) / 2;
            if (im.ch[c].hshift >= 0) im.ch[c].hshift++;
            w = w - (w + 1) / 2;
          } else {
            im.ch[c].h = (h + 1) / 2;
            if (im.ch[c].vshift >= 0) im.ch[c].vshift++;
            h = h - (h + 1) / 2;
          }
          im.ch[c].plane = NewPlane(im.ch[c].w, im.ch[c].h);
          HChan res;
          res.w = w;
          res.h = h;
          res.hshift = im.ch[c].hshift;
          res.vshift = im.ch[c].vshift;
          res.plane = NewPlane(w, h);
          im.ch.insert(im.ch.begin() + offset + (c - beginc), res);
        }
      }
    }
  }

  // Symbolic run of the inverse transforms: appends ops, updates descriptors.
  // Mirrors Transform::Inverse order (lib/jxl/modular/modular_image.cc undo_transforms).
  void EmitInverse(HImage& im, const WPHeader& wp, std::vector<DevOp>* ops) {
    while (!im.transforms.empty()) {
      Transform t = im.transforms.back();
      im.transforms.pop_back();
      if (t.id == kRCT) {
        JXLB_CHECK(EqualChannels(im, t.begin_c, t.begin_c + 2), "rct: channels differ");
        if (t.rct_type == 0) continue;
        DevOp op{};
        op.kind = kOpRCT;
        op.a = im.ch[t.begin_c].plane;
        op.b = im.ch[t.begin_c + 1].plane;
        op.c = im.ch[t.begin_c + 2].plane;
        op.p0 = t.rct_type;
        ops->push_back(op);
      } else if (t.id == kPalette) {
        JXLB_CHECK(im.nb_meta >= 1, "palette without palette channel");
        int nb = im.ch[0].h;
        uint32_t c0 = t.begin_c + 1;
        JXLB_CHECK(c0 < im.ch.size() && nb >= 1, "palette: corrupted");
        HChan idx = im.ch[c0];
        DevOp op{};
        op.kind = kOpPalette;
        op.a = idx.plane;
        op.b = im.ch[0].plane;
        op.p0 = nb;
        op.p1 = std::min(im.bitdepth, 24);
        op.p2 = t.nb_deltas;
        op.p3 = t.predictor;
        wp.Pack(op.wp_params);
        op.c = p_->planes.size();  // the nb - 1 new planes are consecutive
        for (int i = 1; i < nb; i++) {
          HChan nc = idx;
          nc.plane = NewPlane(idx.w, idx.h);
          im.ch.insert(im.ch.begin() + c0 + i, nc);
        }
        if (idx.w != 0 && !(t.nb_deltas == 0 && t.predictor == 0)) {
          // delta palette: keep a copy of the indices + room for the WP state
          size_t extra = static_cast<size_t>(nb) * 10 * (idx.w + 2);
          uint32_t scratch = NewPlane(idx.w, idx.h, extra);
          DevOp cp{};
          cp.kind = kOpCopy;
          cp.a = idx.plane;
          cp.b = scratch;
          ops->push_back(cp);
          op.pad = scratch;
        }
        if (idx.w != 0) ops->push_back(op);
        if (c0 >= im.nb_meta) {
          im.nb_meta--;
        } else {
          im.nb_meta -= 2 - nb;
        }
        im.ch.erase(im.ch.begin());
      } else {
        for (int i = static_cast<int>(t.squeezes.size()) - 1; i >= 0; i--) {
          const SqueezeParams& q = t.squeezes[i];
          uint32_t beginc = q.begin_c, endc = q.begin_c + q.num_c - 1;
          JXLB_CHECK(beginc < im.ch.size() && endc < im.ch.size(), "squeeze: bad range");
          uint32_t offset = q.in_place ? endc + 1 : im.ch.size() + beginc - endc - 1;
          if (beginc < im.nb_meta) {
            JXLB_CHECK(im.nb_meta > q.num_c, "squeeze: bad meta count");
            im.nb_meta -= q.num_c;
          }
          for (uint32_t c = beginc; c <= endc; c++) {
            uint32_t rc = offset + c - beginc;
            JXLB_CHECK(rc < im.ch.size(), "squeeze: residual out of range");
            HChan& a = im.ch[c];
            const HChan& r = im.ch[rc];
            JXLB_CHECK(a.w >= r.w && a.h >= r.h, "squeeze: corrupted");
            if (q.horizontal) {
              JXLB_CHECK(a.w == static_cast<int>(DivCeil(a.w + r.w, 2)) && a.h == r.h, "hsqueeze: bad dims");
              if (r.w == 0) {
                a.hshift--;
                continue;
              }
              HChan out = a;
              out.w = a.w + r.w;
              out.hshift = a.hshift - 1;
              out.plane = NewPlane(out.w, out.h);
              if (r.h != 0) {
                DevOp op{};
                op.kind = kOpHSqueeze;
                op.a = a.plane;
                op.b = r.plane;
                op.c = out.plane;
                ops->push_back(op);
              }
              a = out;
            } else {
              JXLB_CHECK(a.h == static_cast<int>(DivCeil(a.h + r.h, 2)) && a.w == r.w, "vsqueeze: bad dims");
              if (r.h == 0) {
                a.vshift--;
                continue;
              }
              HChan out = a;
              out.h = a.h + r.h;
              out.vshift = a.vshift - 1;
              out.plane = NewPlane(out.w, out.h);
              if (r.w != 0) {
                DevOp op{};
                op.kind = kOpVSqueeze;
                op.a = a.plane;
                op.b = r.plane;
                op.c = out.plane;
                ops->push_back(op);
              }
              a = out;
            }
          }
          im.ch.erase(im.ch.begin() + offset, im.ch.begin() + offset + (endc - beginc + 1));
        }
      }
    }
  }

  // Header part of ModularDecode + emission of the device stream. `file_bit_base`
  // is the bit offset of br's first byte inside the file.
  // What the device found when it decoded the stream planned last (`first_stream` = its index): the answer of an
  // earlier probe round, or -- after turning the plan into a probe for exactly that stream -- ProbePending.
  const ProbeResult& DeviceResult(ProbeCtx* pc, size_t first_stream, const std::vector<uint32_t>& planes) {
    JXLB_CHECK(pc != nullptr, "unsupported: sub-streams chained across host-parsed headers (no device probe available)");
    if (pc->next < pc->done.size()) return pc->done[pc->next++];
    JXLB_CHECK(p_->streams.size() == first_stream + 1, "internal: a probe is exactly one stream");
    p_->streams.erase(p_->streams.begin(), p_->streams.begin() + first_stream);
    p_->is_vardct = false;
    p_->group_programs.clear();
    p_->frame_levels.clear();
    p_->ops.clear();
    p_->out = DevFrameOut{};
    p_->out.vardct = 1;  // nothing to write
    pc->pending = true;
    pc->want_planes = planes;
    throw ProbePending{};
  }

  GroupHeader PlanStream(BitReader& br, uint64_t file_bit_base, HImage& image, uint32_t stream_id,
                         size_t max_chan_size, const HostTree& global) {
    GroupHeader header;
    if (image.ch.empty()) return header;
    header = ReadGroupHeader(br);
    br.CheckInBounds();
    image.transforms = header.transforms;
    for (Transform& t : image.transforms) MetaApply(t, image);
    const size_t nb_channels = image.ch.size();
    auto too_large = [&](size_t i) {
      return i >= image.nb_meta && (static_cast<size_t>(image.ch[i].w) > max_chan_size ||
                                    static_cast<size_t>(image.ch[i].h) > max_chan_size);
    };
    size_t num_chans = 0, distance_multiplier = 0;
    for (size_t i = 0; i < nb_channels; i++) {
      if (!image.ch[i].w || !image.ch[i].h) continue;
      if (too_large(i)) break;
      distance_multiplier = std::max<size_t>(distance_multiplier, image.ch[i].w);
      num_chans++;
    }
    if (num_chans == 0) return header;
    HostTree tree = global;
    if (!header.use_global_tree) {
      uint64_t max_tree_size = 1024;
      for (size_t i = 0; i < nb_channels; i++) {
        if (too_large(i)) break;
        max_tree_size += static_cast<uint64_t>(image.ch[i].w) * image.ch[i].h;
      }
      tree = ReadTreeAndCode(br, std::min<uint64_t>(1 << 20, max_tree_size));
    } else {
      JXLB_CHECK(global.valid, "no global tree available");
    }
    DevStream st{};
    st.bit_pos = file_bit_base + br.BitPos();
    st.bit_end = file_bit_base + br.Size() * 8;
    st.code = tree.code;
    st.tree_off = 0;
    st.stream_id = stream_id;
    st.chan_begin = p_->chans.size();
    st.dist_multiplier = distance_multiplier;
    header.wp.Pack(st.wp_params);
    st.uses_wp = 0;
    st.num_props = tree.num_props;
    st.lz77_slot = 0xFFFFFFFFu;
    uint32_t max_w = 0;
    for (size_t i = 0; i < nb_channels; i++) {
      const HChan& c = image.ch[i];
      if (!c.w || !c.h) continue;
      if (too_large(i)) break;
      DevChannel dc{};
      dc.plane = c.plane;
      dc.prop0 = i;
      dc.ref_off = p_->refs.size();
      const int want = (static_cast<int>(tree.num_props) - 16) / 4;
      for (int j = static_cast<int>(i) - 1; j >= 0 && static_cast<int>(dc.ref_count) < want; j--) {
        const HChan& r = image.ch[j];
        if (r.w != c.w || r.h != c.h || r.hshift != c.hshift || r.vshift != c.vshift) continue;
        p_->refs.push_back(r.plane);
        dc.ref_count++;
      }
      bool ch_wp = false;
      dc.tree_off = PruneTree(*tree.nodes, static_cast<int32_t>(i), static_cast<int32_t>(stream_id), &ch_wp);
      dc.uses_wp = ch_wp;
      if (ch_wp) st.uses_wp = 1;
      if (ch_wp) TryWpLut(&dc, dc.ref_count != 0);
      if (!ch_wp && !NoNwLut()) TryNwLut(&dc);
      BuildCoopLut(&dc);
      p_->chans.push_back(dc);
      max_w = std::max<uint32_t>(max_w, c.w);
    }
    st.chan_end = p_->chans.size();
    st.max_w = max_w;
    p_->wp_width = std::max(p_->wp_width, max_w);
    if (tree.lz77) st.lz77_slot = p_->lz77_slots++;
    p_->streams.push_back(st);
    return header;
  }

  // A stream whose position only the device knows (DevStream::chain_slot): the extra channels of one AC group of a
  // VarDCT frame, which follow the group's coefficients bit by bit (lib/jxl/dec_frame.cc:478-560). The device parses the
  // GroupHeader itself (the first channel's `preamble`) and refuses the stream unless it selects the global tree without
  // transforms -- what libjxl's encoder writes for such groups.
  void PlanChainedStream(const HImage& image, uint32_t stream_id, const HostTree& global, uint64_t bit_end, uint32_t chain_slot) {
    JXLB_CHECK(global.valid, "no global tree available");
    JXLB_CHECK(!global.lz77, "unsupported: LZ77 in the extra-channel streams of a VarDCT frame");
    DevStream st{};
    st.bit_pos = 0;
    st.bit_end = bit_end;
    st.chain_slot = chain_slot;
    st.code = global.code;
    st.stream_id = stream_id;
    st.chan_begin = p_->chans.size();
    WPHeader().Pack(st.wp_params);
    st.num_props = global.num_props;
    st.lz77_slot = 0xFFFFFFFFu;
    uint32_t max_w = 0;
    bool first = true;
    for (size_t i = 0; i < image.ch.size(); i++) {
      const HChan& c = image.ch[i];
      if (!c.w || !c.h) continue;
      DevChannel dc{};
      dc.plane = c.plane;
      dc.prop0 = i;
      dc.ref_off = p_->refs.size();
      const int want = (static_cast<int>(global.num_props) - 16) / 4;
      for (int j = static_cast<int>(i) - 1; j >= 0 && static_cast<int>(dc.ref_count) < want; j--) {
        const HChan& r = image.ch[j];
        if (r.w != c.w || r.h != c.h || r.hshift != c.hshift || r.vshift != c.vshift) continue;
        p_->refs.push_back(r.plane);
        dc.ref_count++;
      }
      bool ch_wp = false;
      dc.tree_off = PruneTree(*global.nodes, static_cast<int32_t>(i), static_cast<int32_t>(stream_id), &ch_wp);
      dc.uses_wp = ch_wp;
      if (ch_wp) st.uses_wp = 1;
      if (ch_wp) TryWpLut(&dc, dc.ref_count != 0);
      if (!ch_wp && !NoNwLut()) TryNwLut(&dc);
      BuildCoopLut(&dc);
      if (first) {
        dc.preamble = 1;  // (count_bits = 0: no channel of variable width)
        first = false;
      }
      p_->chans.push_back(dc);
      max_w = std::max<uint32_t>(max_w, c.w);
      st.dist_multiplier = std::max<uint32_t>(st.dist_multiplier, c.w);
    }
    st.chan_end = p_->chans.size();
    JXLB_CHECK(st.chan_end > st.chan_begin, "internal: chained stream without channels");
    st.max_w = max_w;
    p_->wp_width = std::max(p_->wp_width, max_w);
    p_->streams.push_back(st);
  }

 private:
  FramePlan* p_;
};

}  // namespace jxlb

#endif  // JXLB_PLAN_H_
