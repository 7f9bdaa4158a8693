// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
// C entry points for ctypes (tests/, smoke(), bench.py cpu_baseline only).
#include <cstdio>
#include <cstring>

#include "jxlo_decode.h"

using namespace jxlo;

extern "C" {

struct JxloInfo {
  uint32_t xsize, ysize, bits_per_sample, exponent_bits, num_color_channels, num_extra_channels;
  uint32_t alpha_bits, xyb_encoded, orientation, num_frames;
};

static void SetErr(char* err, size_t n, const char* msg) {
  if (err && n) {
    std::snprintf(err, n, "%s", msg);
  }
}

void* jxlo_decode(const uint8_t* data, size_t size, char* err, size_t errlen) {
  try {
    DecodedImage* img = new DecodedImage(DecodeCodestream(data, size));
    return img;
  } catch (const std::exception& e) {
    SetErr(err, errlen, e.what());
    return nullptr;
  }
}

void jxlo_get_info(void* h, JxloInfo* info) {
  const DecodedImage* img = static_cast<const DecodedImage*>(h);
  info->xsize = img->xsize;
  info->ysize = img->ysize;
  info->bits_per_sample = img->meta.bit_depth.bits;
  info->exponent_bits = img->meta.bit_depth.exp_bits;
  info->num_color_channels = img->meta.color.IsGray() ? 1 : 3;
  info->num_extra_channels = img->meta.extra.size();
  int a = img->meta.AlphaIndex();
  info->alpha_bits = a >= 0 ? img->meta.extra[a].bit_depth.bits : 0;
  info->xyb_encoded = img->meta.xyb_encoded;
  info->orientation = img->meta.orientation;
  info->num_frames = img->frame_info.size();
}

const char* jxlo_frame_info(void* h, uint32_t i) {
  const DecodedImage* img = static_cast<const DecodedImage*>(h);
  return i < img->frame_info.size() ? img->frame_info[i].c_str() : "";
}

size_t jxlo_output_size(void* h, uint32_t num_channels, uint32_t data_type, size_t align) {
  const DecodedImage* img = static_cast<const DecodedImage*>(h);
  return OutputStride(img->xsize, num_channels, data_type, align) * img->ysize;
}

int jxlo_write_pixels(void* h, uint32_t num_channels, uint32_t data_type, uint32_t endianness, size_t align,
                      uint8_t* out, size_t out_size) {
  const DecodedImage* img = static_cast<const DecodedImage*>(h);
  if (out_size < jxlo_output_size(h, num_channels, data_type, align)) return 1;
  WritePixels(*img, num_channels, data_type, endianness, align, out);
  return 0;
}

void jxlo_free(void* h) { delete static_cast<DecodedImage*>(h); }

}  // extern "C"
