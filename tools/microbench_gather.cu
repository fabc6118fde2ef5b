// microbench_gather.cu -- what does one random 4-byte read cost in DRAM traffic on B200, by load flavour?
// (the shuffle's resolve pass moves ~110 bytes per random read under ncu.)  64 M reads at hashed
// indices of a 256 MB array, four in flight per thread.  Time: CUDA events; traffic: run under
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --csv
// usage: microbench_gather [l2fetch 0|32|64|128]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int V>
__device__ __forceinline__ int32_t load(const int32_t* p) {
  int32_t v;
  if (V == 0) v = *p;
  else if (V == 1) v = __ldg(p);
  else if (V == 2) v = __ldcg(p);
  else if (V == 3) v = __ldcs(p);
  else if (V == 4) v = __ldcv(p);
  else if (V == 5) asm volatile("ld.global.L2::64B.b32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (V == 6) asm volatile("ld.global.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (V == 7) asm volatile("ld.global.L1::evict_first.b32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (V == 8) asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (V == 9) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b32 %0, [%1];" : "=r"(v) : "l"(p));
  else v = *p;
  return v;
}

template <int V>
__global__ void __launch_bounds__(256) gather_kernel(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
    int32_t v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      v[u] = i < n ? load<V>(in + (mix((uint32_t)i) % (uint32_t)n)) : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * stride < n) out[i0 + u * stride] = v[u];
  }
}

// random 4-byte WRITES for comparison (the fill of a partial sector)
__global__ void __launch_bounds__(256) scatter_kernel(int32_t* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[mix((uint32_t)i) % (uint32_t)n] = (int32_t)i;
}

template <int V>
void run(const char* name, const int32_t* in, int64_t n, int32_t* out) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  gather_kernel<V><<<148 * 8, 256>>>(in, n, out);
  cudaEventRecord(a);
  gather_kernel<V><<<148 * 8, 256>>>(in, n, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  printf("{\"variant\": %d, \"load\": \"%s\", \"ms\": %.3f, \"greads_per_s\": %.1f, \"err\": \"%s\"}\n", V, name, ms, n / ms / 1e6,
         cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  const int l2f = argc > 1 ? atoi(argv[1]) : 0;
  if (l2f) printf("{\"set_l2_fetch\": %d, \"rc\": \"%s\"}\n", l2f, cudaGetErrorString(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)l2f)));
  size_t g = 0;
  cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
  printf("{\"l2_fetch_granularity\": %zu}\n", g);
  const int64_t n = 64ll << 20;
  int32_t *in, *out;
  cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 4);
  cudaMemset(in, 1, n * 4);
  run<0>("plain", in, n, out);
  run<1>("ld.global.nc", in, n, out);
  run<2>("ld.global.cg", in, n, out);
  run<3>("ld.global.cs", in, n, out);
  run<4>("ld.global.cv", in, n, out);
  run<5>("ld.global.L2::64B", in, n, out);
  run<6>("ld.global.L1::no_allocate", in, n, out);
  run<7>("ld.global.L1::evict_first", in, n, out);
  run<8>("ld.relaxed.gpu", in, n, out);
  run<9>("ld.global.nc.L1::no_allocate.L2::64B", in, n, out);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  scatter_kernel<<<148 * 8, 256>>>(out, n);
  cudaEventRecord(a);
  scatter_kernel<<<148 * 8, 256>>>(out, n);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  printf("{\"scatter_ms\": %.3f, \"gwrites_per_s\": %.1f}\n", ms, n / ms / 1e6);
  return 0;
}
