// Host-side cost of planning a batch (what JxlB200DecoderPlanBatch does on the CPU): n copies of one file, t threads.
//   g++ -O2 -std=c++17 -pthread -I jpegxl-rs_b200/csrc/host -I jpegxl-rs_b200/csrc/kernels tools/host/plan_time.cc -o /tmp/plan_time
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iterator>
#include <vector>

#include "jxlb_batch.h"

int main(int argc, char** argv) {
  if (argc < 4) return 1;
  std::ifstream f(argv[1], std::ios::binary);
  std::vector<uint8_t> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const size_t n = std::atoi(argv[2]);
  const int threads = std::atoi(argv[3]);
  std::vector<const uint8_t*> files(n, data.data());
  std::vector<size_t> sizes(n, data.size());
  jxlb::PixelFormat fmt;
  fmt.num_channels = 3;
  for (int rep = 0; rep < 3; rep++) {
    jxlb::BatchPlan b;
    const auto t0 = std::chrono::steady_clock::now();
    jxlb::PlanBatch(files.data(), sizes.data(), n, fmt, threads, &b, nullptr);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::printf("%zu frames, %d threads: %.1f ms (%.2f ms per frame-thread)\n", n, threads, ms, ms * threads / n);
  }
  return 0;
}
