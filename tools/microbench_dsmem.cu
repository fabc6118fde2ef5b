// microbench_dsmem.cu -- signalling latencies between the CTAs of one thread-block cluster on B200
// (feeds DESIGN.md section 4; not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o microbench_dsmem tools/microbench_dsmem.cu
//
//   pingpong  : cluster of 2, one thread per CTA, round trip of a 4-byte message by
//               0 st.async + complete_tx on the receiver's mbarrier (receiver: mbarrier.try_wait)
//               1 st.shared::cluster of a tagged 8-byte word (receiver: polls its own shared memory)
//               2 cp.async.bulk shared::cta -> shared::cluster, 128 bytes, complete_tx (receiver: try_wait)
//               3 mbarrier.arrive.shared::cluster, no data (receiver: try_wait)
//   allgather : cluster of CS, 256 threads per CTA: every CTA delivers 1 KB to every CTA, every CTA
//               waits for all of them; per iteration, by
//               0 one 1 KB bulk copy per destination + mbarrier
//               1 tagged 8-byte words, st.shared::cluster, receivers poll shared memory
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      std::printf("{\"error\": \"%s\", \"at\": \"%s\"}\n", cudaGetErrorString(e_), #x);        \
      std::exit(1);                                                                            \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void st_async(uint32_t dst, uint32_t v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst), "r"(v), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void st_cluster_u2(uint32_t dst, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.u32 [%0], {%1,%2};" ::"r"(dst), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void bulk_s2c(uint32_t dst, uint32_t src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "r"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void remote_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint2 lds_volatile_u2(uint32_t a) {
  uint2 v;
  asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}

template <int MODE>
__global__ void pingpong_kernel(int iters, long long* cycles) {
  __shared__ __align__(128) uint64_t bar;
  __shared__ __align__(128) uint32_t word[64];
  const uint32_t me = ctarank(), peer = me ^ 1u;
  const uint32_t b = smem_u32(&bar), w = smem_u32(word);
  if (threadIdx.x == 0) {
    mbar_init(b, 1);
    word[0] = 0;
    word[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();
  if (threadIdx.x == 0) {
    const uint32_t rw = mapa(w, peer), rb = mapa(b, peer);
    auto send = [&](uint32_t tag) {
      if (MODE == 0) st_async(rw, tag, rb);
      if (MODE == 1) st_cluster_u2(rw, tag, tag);
      if (MODE == 2) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bulk_s2c(rw, w + 128, 128, rb);
      }
      if (MODE == 3) remote_arrive(rb);
    };
    auto recv = [&](int i) {
      if (MODE == 0) { mbar_expect(b, 4); mbar_wait(b, i & 1); }
      if (MODE == 1) { while (lds_volatile_u2(w).y != (uint32_t)(i + 1)) {} }
      if (MODE == 2) { mbar_expect(b, 128); mbar_wait(b, i & 1); }
      if (MODE == 3) { mbar_wait(b, i & 1); }
    };
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (me == 0) {
        send(i + 1);
        recv(i);
      } else {
        recv(i);
        send(i + 1);
      }
    }
    if (me == 0) *cycles = clock64() - t0;
  }
  cluster_sync();
}

// every CTA delivers 256 floats to every CTA of the cluster
template <int MODE>
__global__ void allgather_kernel(int iters, long long* cycles) {
  extern __shared__ __align__(128) unsigned char sm[];
  // [0,16): barriers (2)   [1024, 3072): send[2][256] floats   [4096, ...): recv
  const uint32_t s0 = smem_u32(sm);
  const uint32_t me = ctarank(), CS = nctarank();
  const int t = threadIdx.x;
  float* send = reinterpret_cast<float*>(sm + 1024);
  if (t == 0) {
    mbar_init(s0, 1);
    mbar_init(s0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect(s0, CS * 1024);
    mbar_expect(s0 + 8, CS * 1024);
  }
  for (int i = t; i < (int)(2 * CS * 256 * 2); i += blockDim.x) reinterpret_cast<uint32_t*>(sm + 4096)[i] = 0;
  __syncthreads();
  cluster_sync();
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int par = it & 1;
    const uint32_t tag = (uint32_t)it + 1u;
    const float v = acc + (float)t;
    if (MODE == 0) {
      if (t == 0 && it >= 1 && it + 1 < iters) mbar_expect(s0 + 8 * (par ^ 1), CS * 1024);
      send[par * 256 + t] = v;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      const int dst = 2 * (t >> 5) + (t & 31);
      if ((t & 31) < 2 && dst < (int)CS)
        bulk_s2c(mapa(s0 + 4096 + (par * 16 + me) * 1024, dst), smem_u32(send + par * 256), 1024, mapa(s0 + 8 * par, dst));
      if (t < 128) {
        mbar_wait(s0 + 8 * par, (it >> 1) & 1);
        const float* rb = reinterpret_cast<const float*>(sm + 4096) + par * 16 * 256 + t;
        float a0 = 0.f, a1 = 0.f;
        for (int i = 0; i < (int)CS; ++i) {
          a0 += rb[i * 256];
          a1 += rb[i * 256 + 128];
        }
        acc = a0 + a1;
      }
      __syncthreads();
    } else {
      // tagged words: value t of CTA me -> slot [par][me][t] (8 bytes) of every CTA
      const uint32_t slot = s0 + 4096 + ((par * 16 + me) * 256 + t) * 8;
      for (uint32_t dst = 0; dst < CS; ++dst) st_cluster_u2(mapa(slot, dst), __float_as_uint(v), tag);
      if (t < 128) {
        float a0 = 0.f, a1 = 0.f;
        for (uint32_t i = 0; i < CS; ++i) {
          const uint32_t a = s0 + 4096 + ((par * 16 + i) * 256 + t) * 8;
          uint2 x, y;
          do { x = lds_volatile_u2(a); } while (x.y != tag);
          do { y = lds_volatile_u2(a + 128 * 8); } while (y.y != tag);
          a0 += __uint_as_float(x.x);
          a1 += __uint_as_float(y.x);
        }
        acc = a0 + a1;
      }
      __syncthreads();
    }
  }
  if (t == 0 && me == 0) *cycles = clock64() - t0;
  if (acc == 12345.678f) std::printf("x");
  cluster_sync();
}

template <typename K>
static bool launch_cluster(K kern, int cs, int threads, size_t smem, int iters, long long* cyc) {
  cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(cs);
  lc.blockDim = dim3(threads);
  lc.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&lc, kern, iters, cyc);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    std::printf("{\"error\": \"%s\", \"cluster\": %d}\n", cudaGetErrorString(e), cs);
    cudaGetLastError();
    return false;
  }
  return true;
}

int main() {
  long long* cyc;
  CK(cudaMallocManaged(&cyc, sizeof(long long)));
  const int iters = 20000;
  const char* pp[] = {"st.async + complete_tx, try_wait", "st.shared::cluster tagged word, smem poll",
                      "cp.async.bulk 128 B + complete_tx, try_wait", "mbarrier.arrive remote, try_wait"};
  if (launch_cluster(pingpong_kernel<0>, 2, 32, 0, iters, cyc))
    std::printf("{\"test\": \"dsmem_pingpong\", \"method\": \"%s\", \"round_trip_cycles\": %.1f}\n", pp[0], (double)*cyc / iters);
  if (launch_cluster(pingpong_kernel<1>, 2, 32, 0, iters, cyc))
    std::printf("{\"test\": \"dsmem_pingpong\", \"method\": \"%s\", \"round_trip_cycles\": %.1f}\n", pp[1], (double)*cyc / iters);
  if (launch_cluster(pingpong_kernel<2>, 2, 32, 0, iters, cyc))
    std::printf("{\"test\": \"dsmem_pingpong\", \"method\": \"%s\", \"round_trip_cycles\": %.1f}\n", pp[2], (double)*cyc / iters);
  if (launch_cluster(pingpong_kernel<3>, 2, 32, 0, iters, cyc))
    std::printf("{\"test\": \"dsmem_pingpong\", \"method\": \"%s\", \"round_trip_cycles\": %.1f}\n", pp[3], (double)*cyc / iters);
  for (int cs : {2, 4, 8, 16}) {
    const size_t smem = 4096 + 2 * 16 * 256 * 8;
    if (launch_cluster(allgather_kernel<0>, cs, 256, smem, iters, cyc))
      std::printf("{\"test\": \"dsmem_allgather\", \"cluster\": %d, \"method\": \"1 KB bulk copy per destination + mbarrier\", \"cycles_per_iter\": %.1f}\n",
                  cs, (double)*cyc / iters);
    if (launch_cluster(allgather_kernel<1>, cs, 256, smem, iters, cyc))
      std::printf("{\"test\": \"dsmem_allgather\", \"cluster\": %d, \"method\": \"tagged 8-byte words, st.shared::cluster + smem poll\", \"cycles_per_iter\": %.1f}\n",
                  cs, (double)*cyc / iters);
  }
  return 0;
}
