// cuda_emu.hpp -- just enough of the CUDA execution model to run block-level kernels on the host: one OS
// thread per CUDA thread of a block, blocks one after the other, __syncthreads() = a barrier, atomics =
// GCC builtins, __shared__ = a static (the blocks of a launch run sequentially, so one copy is enough).
// Test infrastructure (tests/test_shuffle_emu.py): the GPU pool has no compute-sanitizer, so the device
// shuffle's kernels -- written with these primitives only -- are also executed here against the host loop.
#pragma once
#include <pthread.h>
#include <stdint.h>

#include <functional>
#include <thread>
#include <vector>

#define B200SGD_HOST_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

struct emu_dim3 {
  unsigned x = 1, y = 1, z = 1;
};
struct uint2 {
  uint32_t x, y;
};
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

inline thread_local emu_dim3 threadIdx, blockIdx;
inline emu_dim3 blockDim, gridDim;
inline pthread_barrier_t emu_barrier;

inline void __syncthreads() { pthread_barrier_wait(&emu_barrier); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned val) {
  __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
  return cmp;
}

// kernel<<<grid, block>>>(args...)  ->  emu_launch(grid, block, [&] { kernel(args...); });
inline void emu_launch(unsigned grid, unsigned block, const std::function<void()>& body) {
  gridDim.x = grid;
  blockDim.x = block;
  pthread_barrier_init(&emu_barrier, nullptr, block);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < block; ++t)
    th.emplace_back([&, t] {
      threadIdx.x = t;
      for (unsigned b = 0; b < grid; ++b) {
        blockIdx.x = b;
        body();
        pthread_barrier_wait(&emu_barrier);  // the next block reuses the "shared memory"
      }
    });
  for (auto& x : th) x.join();
  pthread_barrier_destroy(&emu_barrier);
}
