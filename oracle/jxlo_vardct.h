// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
// VarDCT path -- placeholder until the restatement lands.
#ifndef JXLO_VARDCT_H_
#define JXLO_VARDCT_H_
#include "jxlo_frame.h"
namespace jxlo {
struct VarDCTState {
  float dc_quant[3];
  VarDCTState(const FrameHeader&, const FrameDimensions&, const ImageMetadata&) {
    throw Error("jxlo: VarDCT frames are not supported yet");
  }
};
inline void VarDCTReadGlobalDC(BitReader&, VarDCTState*) {}
inline void VarDCTReadDCGroup(BitReader&, VarDCTState*, ModularFrameState*, size_t) {}
inline void VarDCTReadACMetadata(BitReader&, VarDCTState*, ModularFrameState*, size_t) {}
inline void VarDCTFinalizeDC(VarDCTState*) {}
inline void VarDCTReadGlobalAC(BitReader&, VarDCTState*, ModularFrameState*) {}
inline void VarDCTReadACGroup(BitReader&, VarDCTState*, size_t, size_t) {}
inline void VarDCTToPixels(VarDCTState*, std::vector<Plane>*) {}
}  // namespace jxlo
#endif
