// jxl_b200 host parser (product code; runs on the CPU in front of the CUDA kernels).
// Bit-level input for the host-side bitstream parse.
//
// Bit-level input. Follows the semantics of libjxl's BitReader
// (lib/jxl/dec_bit_reader.h:84-147): LSB-first inside little-endian bytes,
// reads past the end return zero bits and are remembered as an over-read.
#ifndef JXLB_BITS_H_
#define JXLB_BITS_H_

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

namespace jxlb {

struct Error : public std::runtime_error {
  explicit Error(const std::string& what) : std::runtime_error(what) {}
};

#define JXLB_CHECK(cond, msg)                                   \
  do {                                                          \
    if (!(cond)) throw ::jxlb::Error(std::string("jxl_b200: ") + msg); \
  } while (0)

class BitReader {
 public:
  BitReader() = default;
  BitReader(const uint8_t* data, size_t size) : data_(data), size_(size) {}

  // Up to 57 bits starting at the cursor, without consuming them.
  uint64_t Window() const {
    size_t byte = pos_ >> 3;
    uint64_t w = 0;
    if (byte + 8 <= size_) {
      std::memcpy(&w, data_ + byte, 8);
    } else if (byte < size_) {
      std::memcpy(&w, data_ + byte, size_ - byte);
    }
    return w >> (pos_ & 7);
  }
  uint32_t Peek(unsigned n) const {  // n <= 32
    return static_cast<uint32_t>(Window() & ((uint64_t{1} << n) - 1));
  }
  void Skip(size_t n) { pos_ += n; }
  void SeekTo(size_t bit) { pos_ = bit; }
  uint32_t Read(unsigned n) {
    uint32_t v = Peek(n);
    pos_ += n;
    return v;
  }
  uint64_t Read64(unsigned n) {  // n <= 56
    uint64_t v = Window() & ((uint64_t{1} << n) - 1);
    pos_ += n;
    return v;
  }
  bool ReadBool() { return Read(1) != 0; }

  // lib/jxl/dec_bit_reader.h:206-213: the padding must be zero.
  void AlignToByte() {
    unsigned rem = pos_ & 7;
    if (rem == 0) return;
    JXLB_CHECK(Read(8 - rem) == 0, "non-zero padding bits");
  }
  size_t BitPos() const { return pos_; }
  size_t BytePos() const { return (pos_ + 7) >> 3; }
  size_t Size() const { return size_; }
  const uint8_t* Data() const { return data_; }
  bool InBounds() const { return pos_ <= size_ * 8; }
  void CheckInBounds() const { JXLB_CHECK(InBounds(), "read past end of section"); }

 private:
  const uint8_t* data_ = nullptr;
  size_t size_ = 0;
  size_t pos_ = 0;
};

// ---- field coders (lib/jxl/fields.cc:499-640, lib/jxl/field_encodings.h) ----

// One of the four alternatives of a U32 field.
struct U32Dist {
  // bits == 0xFF marks a direct value.
  uint32_t bits;
  uint32_t offset;
};
inline constexpr U32Dist Val(uint32_t v) { return U32Dist{0xFF, v}; }
inline constexpr U32Dist Bits(uint32_t n) { return U32Dist{n, 0}; }
inline constexpr U32Dist BitsOffset(uint32_t n, uint32_t off) { return U32Dist{n, off}; }

inline uint32_t ReadU32(BitReader& br, U32Dist d0, U32Dist d1, U32Dist d2, U32Dist d3) {
  const U32Dist d[4] = {d0, d1, d2, d3};
  const U32Dist s = d[br.Read(2)];
  if (s.bits == 0xFF) return s.offset;
  return br.Read(s.bits) + s.offset;
}

// lib/jxl/fields.cc:549-575
inline uint64_t ReadU64(BitReader& br) {
  uint32_t sel = br.Read(2);
  if (sel == 0) return 0;
  if (sel == 1) return 1 + br.Read(4);
  if (sel == 2) return 17 + br.Read(8);
  uint64_t v = br.Read(12);
  unsigned shift = 12;
  while (br.Read(1)) {
    if (shift == 60) {
      v |= static_cast<uint64_t>(br.Read(4)) << shift;
      break;
    }
    v |= static_cast<uint64_t>(br.Read(8)) << shift;
    shift += 8;
  }
  return v;
}

// lib/jxl/fields.cc:605-640: IEEE binary16, inf/nan rejected.
inline float ReadF16(BitReader& br) {
  uint32_t h = br.Read(16);
  uint32_t sign = h >> 15, e = (h >> 10) & 31, m = h & 1023;
  JXLB_CHECK(e != 31, "F16 inf/nan");
  float v;
  if (e == 0) {
    v = (1.0f / 16384) * (m * (1.0f / 1024));
  } else {
    uint32_t b = ((e + 112) << 23) | (m << 13);
    std::memcpy(&v, &b, 4);
  }
  return sign ? -v : v;
}

// lib/jxl/fields.h:208-219
inline uint32_t ReadEnum(BitReader& br) {
  return ReadU32(br, Val(0), Val(1), BitsOffset(4, 2), BitsOffset(6, 18));
}

inline int32_t UnpackSigned(uint32_t u) {
  return static_cast<int32_t>((u >> 1) ^ (~(u & 1) + 1));
}

inline unsigned FloorLog2(uint64_t v) { return 63 - __builtin_clzll(v); }
inline unsigned CeilLog2(uint64_t v) { return v <= 1 ? 0 : FloorLog2(v - 1) + 1; }
inline size_t DivCeil(size_t a, size_t b) { return (a + b - 1) / b; }

// Extensions trailer shared by most bundles (lib/jxl/fields.cc:199-260):
// a U64 bitmask, one U64 bit-length per set bit, then that many payload bits.
inline void SkipExtensions(BitReader& br) {
  uint64_t ext = ReadU64(br);
  uint64_t total = 0;
  for (uint64_t rem = ext; rem != 0; rem &= rem - 1) total += ReadU64(br);
  br.Skip(total);
}

}  // namespace jxlb

#endif  // JXLB_BITS_H_
