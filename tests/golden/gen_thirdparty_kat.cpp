// Golden-vector generator for the THIRD-PARTY pieces the reference's hot path depends on.
//
// The reference (thvasilo/cuda-sgd-sese-project) defines its per-epoch permutation and its
// initial weights through code that is NOT in its repository:
//   * glibc rand() (never seeded => srand(1))              call sites c_code/main.cu:90, :207
//   * libstdc++ std::random_shuffle (uses rand())          call sites c_code/main.cu:90, :207
//   * thrust::default_random_engine +
//     thrust::uniform_real_distribution<float>(0, 0.01)    call site  c_code/main.cu:389-395
//
// This program calls the REAL glibc / libstdc++ / Thrust (CCCL, host-only) implementations that
// ship in this image and prints known-answer values.  tests/golden/make_golden.py drives it (one
// fresh process per record so that the rand() stream always starts at seed 1) and commits the
// result as tests/golden/thirdparty_kat.json.  The oracle (oracle/sgd_oracle.c) and the product
// (csrc/host_rng.hpp) are both checked against that file.
//
// Build: g++ -O2 -std=c++17 -I/usr/local/cuda/include -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_CPP
//        -Wno-deprecated-declarations gen_thirdparty_kat.cpp -o gen_kat
#include <algorithm>
#include <cinttypes>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <thrust/random.h>

static void usage() {
  std::fprintf(stderr,
               "usage: gen_kat rand N | rand_at K | shuffle R EPOCHS [full] | weights C\n");
}

int main(int argc, char** argv) {
  if (argc < 3) { usage(); return 2; }
  const std::string mode = argv[1];
  if (mode == "rand") {  // first N outputs of unseeded rand()
    long n = std::atol(argv[2]);
    std::printf("[");
    for (long i = 0; i < n; ++i) std::printf("%s%d", i ? "," : "", rand());
    std::printf("]\n");
  } else if (mode == "rand_at") {  // output number K (0-based) and a wrap-around checksum of 0..K
    long k = std::atol(argv[2]);
    uint64_t sum = 0;
    int last = 0;
    for (long i = 0; i <= k; ++i) { last = rand(); sum += (uint64_t)(i + 1) * (uint64_t)last; }
    std::printf("{\"k\": %ld, \"value\": %d, \"checksum\": %" PRIu64 "}\n", k, last, sum);
  } else if (mode == "shuffle") {  // cumulative std::random_shuffle of identity(R), EPOCHS times
    if (argc < 4) { usage(); return 2; }
    long R = std::atol(argv[2]);
    int epochs = std::atoi(argv[3]);
    bool full = argc > 4 && std::strcmp(argv[4], "full") == 0;
    std::vector<int> ind(R);
    for (long i = 0; i < R; ++i) ind[i] = (int)i;
    std::printf("[");
    for (int e = 1; e <= epochs; ++e) {
      std::random_shuffle(ind.begin(), ind.end());
      uint64_t cs = 0;
      for (long i = 0; i < R; ++i) cs += (uint64_t)(i + 1) * (uint64_t)ind[i];
      std::printf("%s{\"R\": %ld, \"epoch\": %d, \"head\": [", e > 1 ? "," : "", R, e);
      long nh = full ? R : std::min<long>(R, 4);
      for (long i = 0; i < nh; ++i) std::printf("%s%d", i ? "," : "", ind[i]);
      std::printf("], \"last\": %d, \"checksum\": %" PRIu64 "}", ind[R - 1], cs);
    }
    std::printf("]\n");
  } else if (mode == "weights") {  // main.cu:389-395 with the real Thrust host RNG
    int C = std::atoi(argv[2]);
    thrust::default_random_engine rng;
    thrust::uniform_real_distribution<float> weight_dist(0.0, 0.01);
    std::printf("[");
    for (int i = 0; i < C; ++i) {
      float w = weight_dist(rng);
      uint32_t bits;
      std::memcpy(&bits, &w, 4);
      std::printf("%s\"0x%08x\"", i ? "," : "", bits);
    }
    std::printf("]\n");
  } else {
    usage();
    return 2;
  }
  return 0;
}
