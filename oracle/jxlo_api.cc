// ORACLE (test infrastructure, not the product). See jxlo_bits.h.
// C entry points for ctypes (tests/, smoke(), bench.py cpu_baseline only).
#include <cstdio>
#include <cstring>

#include "jxlo_decode.h"
#include "jxlo_encode.h"
#include "jxlo_enc_modular.h"

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

// endianness: JxlEndianness, | 0x400 = undo the image's orientation (libjxl's default; rows are then `ysize` samples
// wide for the transposing orientations 5 - 8).
int jxlo_write_pixels(void* h, uint32_t num_channels, uint32_t data_type, uint32_t endianness, size_t align,
                      uint8_t* out, size_t out_size) {
  const DecodedImage* img = static_cast<const DecodedImage*>(h);
  const bool undo = (endianness & 0x400) != 0;
  endianness &= 0xFF;
  const bool transpose = undo && img->meta.orientation >= 5;
  const size_t need = transpose ? OutputStride(img->ysize, num_channels, data_type, align) * img->xsize
                                : jxlo_output_size(h, num_channels, data_type, align);
  if (out_size < need) return 1;
  WritePixels(*img, num_channels, data_type, endianness, align, out, undo);
  return 0;
}

void jxlo_free(void* h) { delete static_cast<DecodedImage*>(h); }

// ---- known-answer hooks for the transform restatement (tests/test_oracle_vardct.py) ----
// rows x cols pixels -> coefficients in the min x max layout (ComputeScaledDCT).
void jxlo_scaled_dct(int rows, int cols, const float* pixels, float* coeffs) {
  std::vector<float> scratch(static_cast<size_t>(rows) * cols + 3 * std::max(rows, cols) + 16);
  ScaledDCT(rows, cols, pixels, cols, coeffs, scratch.data());
}
void jxlo_scaled_idct(int rows, int cols, const float* coeffs, float* pixels) {
  std::vector<float> from(coeffs, coeffs + static_cast<size_t>(rows) * cols);
  std::vector<float> scratch(static_cast<size_t>(rows) * cols + 2 * std::max(rows, cols) + 16);
  ScaledIDCT(rows, cols, from.data(), pixels, cols, scratch.data());
}
// TransformToPixels for any AcStrategy; pixels is (8 * covered_y) x (8 * covered_x), dense.
void jxlo_transform_to_pixels(int strategy, const float* coeffs, float* pixels) {
  const size_t n = 64u * kCoveredX[strategy] * kCoveredY[strategy];
  std::vector<float> c(coeffs, coeffs + n), scratch(3 * n + 64);
  TransformToPixels(strategy, c.data(), pixels, 8 * kCoveredX[strategy], scratch.data());
}
// TransformFromPixels for any AcStrategy up to 64x64; pixels is (8 * covered_y) x (8 * covered_x), dense.
void jxlo_transform_from_pixels(int strategy, const float* pixels, float* coeffs) {
  const size_t n = 64u * kCoveredX[strategy] * kCoveredY[strategy];
  std::vector<float> scratch(5 * n + 1024);
  TransformFromPixelsRef(strategy, pixels, 8 * kCoveredX[strategy], coeffs, scratch.data());
}
// DCFromLowestFrequencies: coefficients of one varblock -> covered_y x covered_x DC values (dense).
void jxlo_dc_from_llf(int strategy, const float* coeffs, float* dc) {
  DCFromLowestFrequencies(strategy, coeffs, dc, kCoveredX[strategy]);
}
void jxlo_llf_from_dc(int strategy, const float* dc, size_t dc_stride, float* llf) {
  LowestFrequenciesFromDC(strategy, dc, dc_stride, llf);
}
void jxlo_natural_coeff_order(int strategy, uint32_t* order) { NaturalCoeffOrder(strategy, order); }
// Library dequantisation table of quant table `table` (3 * 64 * rx * ry floats); returns the count.
size_t jxlo_library_quant_table(int table, float* out, size_t cap) {
  std::vector<float> t = ComputeQuantTable(LibraryEncoding(table), table);
  if (out && cap >= t.size()) std::memcpy(out, t.data(), t.size() * sizeof(float));
  return t.size();
}
// VarDCT stream generator (oracle/jxlo_encode.h). Returns the codestream size, or 0 on error;
// call with out == NULL to get the size... the stream is kept until the next call on this thread.
static thread_local std::vector<uint8_t> g_encoded;
// The alpha plane (xsize * ysize uint8) of the NEXT jxlo_encode_vardct call on this thread (EncodeParams::alpha).
static thread_local const uint8_t* g_next_alpha = nullptr;
void jxlo_set_next_alpha(const uint8_t* alpha) { g_next_alpha = alpha; }

size_t jxlo_encode_vardct(const uint8_t* rgb, uint32_t xsize, uint32_t ysize, float distance, int strategy_mode,
                          uint32_t seed, int gab, uint32_t epf_iters, int dc_smoothing, int random_side_info,
                          uint32_t num_passes, int dc_tree, char* err, size_t errlen) {
  try {
    EncodeParams p;
    p.distance = distance;
    p.strategy_mode = strategy_mode;
    p.seed = seed;
    p.gab = (gab & 1) != 0;          // bit 0: Gaborish signalled; bit 1: no inverse Gaborish; bit 2: natural orders
    p.gab_inverse = (gab & 2) == 0;
    p.coeff_orders = (gab & 4) == 0;
    p.cfl = (gab & 8) == 0;          // bit 3: no chroma-from-luma fit
    p.adaptive_quant = (gab & 16) == 0;  // bit 4: constant quant field instead of libjxl's adaptive one
    p.prefix_codes = (gab & 32) != 0;    // bit 5: prefix codes instead of ANS in every stream
    p.upsampling = 1u << ((gab >> 6) & 3);  // bits 6-7: log2 of the frame upsampling
    p.orientation = 1 + ((gab >> 8) & 7);   // bits 8-10: the image's orientation - 1
    p.alpha = g_next_alpha;
    g_next_alpha = nullptr;
    p.splines = (gab >> 11) & 15;           // bits 11-14: number of random splines
    p.epf_iters = epf_iters;
    p.dc_smoothing = dc_smoothing != 0;
    p.random_side_info = random_side_info != 0;
    p.num_passes = num_passes;
    p.dc_tree = dc_tree;
    g_encoded = EncodeVarDCT(rgb, xsize, ysize, p);
    return g_encoded.size();
  } catch (const std::exception& e) {
    SetErr(err, errlen, e.what());
    return 0;
  }
}
// Lossless Modular stream generator (oracle/jxlo_enc_modular.h). `params`: 16 uint32 (see tests/jxlo.py encode_modular).
size_t jxlo_encode_modular(const uint16_t* samples, uint32_t xsize, uint32_t ysize, const uint32_t* params, char* err,
                           size_t errlen) {
  try {
    ModularEncodeParams p;
    p.bits = params[0];
    p.num_color = params[1];
    p.alpha = params[2] != 0;
    p.group_size_shift = params[3];
    p.tree = static_cast<int>(params[4]);
    p.predictor = params[5];
    p.seed = params[6];
    p.rct = static_cast<int>(params[7]) - 1;
    p.palette_colors = params[8];
    p.palette_deltas = params[9];
    p.palette_predictor = params[10];
    p.squeeze = params[11] != 0;
    p.entropy = static_cast<int>(params[12]);
    p.lz77_min_symbol = params[13] ? params[13] : 224;
    p.orientation = params[14] ? params[14] : 1;
    p.splines = params[15];
    g_encoded = EncodeModular(samples, xsize, ysize, p);
    return g_encoded.size();
  } catch (const std::exception& e) {
    SetErr(err, errlen, e.what());
    return 0;
  }
}
void jxlo_encoded_copy(uint8_t* out) { std::memcpy(out, g_encoded.data(), g_encoded.size()); }

float jxlo_fast_powf(float b, float e) { return FastPowf(b, e); }
float jxlo_srgb_from_linear(float v) { return SrgbFromLinear(v); }

}  // extern "C"
