// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
// Render stages -- placeholder until the restatement lands.
#ifndef JXLO_RENDER_H_
#define JXLO_RENDER_H_
#include "jxlo_frame.h"
#include "jxlo_vardct.h"
namespace jxlo {
struct FeatureState {};
inline void ReadPatches(BitReader&, const FrameDimensions&, const ImageMetadata&, FeatureState*) {
  throw Error("jxlo: patches are not supported yet");
}
inline void RenderFrame(const FrameHeader& fh, const FrameDimensions&, const CodestreamState&, VarDCTState*,
                        const FeatureState&, std::vector<Plane>*, bool* is_xyb) {
  JXLO_CHECK(fh.color_transform == kCTNone, "colour transforms are not supported yet");
  JXLO_CHECK(!fh.lf.gab && fh.lf.epf_iters == 0, "loop filters are not supported yet");
  JXLO_CHECK(fh.upsampling == 1, "upsampling is not supported yet");
  *is_xyb = false;
}
}  // namespace jxlo
#endif
