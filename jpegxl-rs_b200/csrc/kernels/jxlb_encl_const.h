// jxl_b200 lossless encoder: the constants shared by the device code (jxlb_encl_dev.h) and the host code
// (host/jxlb_encl_host.h).
#ifndef JXLB_ENCL_CONST_H_
#define JXLB_ENCL_CONST_H_

#include <stdint.h>

namespace jxlb {

// Group size of the lossless frames: FrameHeader::group_size_shift 0 = 128 x 128 (libjxl's
// JXL_ENC_FRAME_SETTING_MODULAR_GROUP_SIZE 0; its default is 1 = 256 x 256, lib/jxl/frame_header.h). A group is one rANS
// stream, i.e. one serial chain for the emission kernel: with 128 x 128 groups a 4K frame has 540 chains of 49 K tokens
// instead of 135 of 197 K (k_encl_emit: 50 -> 27 ms per 16 frames, no longer bound by the length of one chain) and a
// decoder finds four times the sections to decode in parallel, for the edge contexts of the extra group borders
// (7.7605 -> 7.7687 bpp on the 4K photograph).
constexpr uint32_t kEnclGroupShift = 0;
constexpr uint32_t kEnclGroupDim = 128u << kEnclGroupShift;
constexpr uint32_t kEnclGroupSamples = kEnclGroupDim * kEnclGroupDim;  // token slots of one channel of a group

}  // namespace jxlb

#endif  // JXLB_ENCL_CONST_H_
