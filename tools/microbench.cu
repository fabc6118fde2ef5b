// tools/microbench.cu -- inter-SM signalling latencies on B200 that the step-tail design rests on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/microbench.cu -o tools/microbench
// Run on the GPU box: ./tools/microbench
//
//   pingpong   : CTA 0 and CTA k bounce a tagged word through L2 (one-way latency = half a round)
//   fan        : the step tail in isolation: G CTAs publish a tagged float4, 1 owner gathers them,
//                publishes a result to G private mailboxes (or 1 shared word), all CTAs wait for it
//   gridbar    : red.release + ld.acquire counter barrier over G CTAs
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e)); \
      std::exit(1);                                                                        \
    }                                                                                      \
  } while (0)

enum Flavor { VOLATILE_SYS = 0, RELAXED_GPU = 1, RELEASE_ACQUIRE_GPU = 2 };

template <int F>
__device__ __forceinline__ void st_u32(unsigned* p, unsigned v) {
  if (F == VOLATILE_SYS) asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  if (F == RELAXED_GPU) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  if (F == RELEASE_ACQUIRE_GPU) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int F>
__device__ __forceinline__ unsigned ld_u32(const unsigned* p) {
  unsigned v;
  if (F == VOLATILE_SYS) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  if (F == RELAXED_GPU) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  if (F == RELEASE_ACQUIRE_GPU) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- ping-pong: block 0 <-> block `peer`, words 128 B apart ------------------------------------
template <int F>
__global__ void pingpong_kernel(unsigned* flags, int peer, int iters, long long* cycles) {
  if (threadIdx.x != 0) return;
  unsigned* a = flags;        // written by block 0
  unsigned* b = flags + 64;   // written by block peer
  if (blockIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
      st_u32<F>(a, (unsigned)i);
      while (ld_u32<F>(b) != (unsigned)i) {
      }
    }
    *cycles = clock64() - t0;
  } else if ((int)blockIdx.x == peer) {
    for (int i = 1; i <= iters; ++i) {
      while (ld_u32<F>(a) != (unsigned)i) {
      }
      st_u32<F>(b, (unsigned)i);
    }
  }
}

// ---- fan-in / fan-out with tagged 16-byte words ------------------------------------------------
__device__ __forceinline__ void st_u4(uint4* p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// every CTA: thread 0 publishes part[cta]; owner CTA (block 0), 32 lanes gather all G parts,
// then publish to mailboxes (private: G copies, 256 B apart; shared: 1 copy); everyone waits.
__global__ void fan_kernel(uint4* part, uint4* mbox, int private_boxes, int iters, int sleep_ns, long long* cycles) {
  const int G = gridDim.x, cta = blockIdx.x, lane = threadIdx.x;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    const unsigned tag = (unsigned)it;
    if (lane == 0) st_u4(part + (size_t)cta * 16, make_uint4(cta, tag, it, tag));
    if (cta == 0) {
      unsigned acc = 0;
      for (int src = lane; src < G; src += 32) {
        uint4 v;
        do { v = ld_u4(part + (size_t)src * 16); } while (v.y != tag || v.w != tag);
        acc += v.x;
      }
      for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
      if (private_boxes) {
        for (int dst = lane; dst < G; dst += 32) st_u4(mbox + (size_t)dst * 16, make_uint4(acc, tag, acc, tag));
      } else if (lane == 0) {
        st_u4(mbox, make_uint4(acc, tag, acc, tag));
      }
    }
    if (lane == 0) {
      const uint4* mine = mbox + (private_boxes ? (size_t)cta * 16 : 0);
      uint4 v;
      while (true) {
        v = ld_u4(mine);
        if (v.y == tag && v.w == tag) break;
        if (sleep_ns) __nanosleep(sleep_ns);
      }
    }
    __syncthreads();
  }
  if (cta == 0 && lane == 0) *cycles = clock64() - t0;
}

// fan2: closer to the real step tail.  Block = 256 threads.  Every CTA: thread 0 publishes a
// tagged word.  Owner CTAs (the first `owners` CTAs, one column each): thread t < G polls source t
// (all sources in parallel), block-reduces, thread t < G writes mailbox[t][owner].  Every CTA:
// `npoll` threads poll (variant A) their own word of the CTA's mailbox, or (variant B, sentinel)
// thread 0 polls the last word and then all words are read once and verified.
__global__ void fan2_kernel(uint4* part, uint4* mbox, int owners, int sentinel, int sleep_ns, int iters,
                            long long* cycles) {
  const int G = gridDim.x, cta = blockIdx.x, t = threadIdx.x;
  __shared__ unsigned red[8];
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    const unsigned tag = (unsigned)it;
    if (t < owners) st_u4(part + ((size_t)cta * 32 + t) * 2, make_uint4(cta, tag, it, tag));  // one word per owner column
    if (cta < owners) {
      unsigned acc = 0;
      if (t < G) {
        uint4 v;
        while (true) {
          v = ld_u4(part + ((size_t)t * 32 + cta) * 2);
          if (v.y == tag && v.w == tag) break;
          if (sleep_ns) __nanosleep(sleep_ns);
        }
        acc = v.x;
      }
      for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
      if ((t & 31) == 0) red[t >> 5] = acc;
      __syncthreads();
      acc = 0;
      for (int k = 0; k < 8; ++k) acc += red[k];
      __syncthreads();
      if (t < G) st_u4(mbox + ((size_t)t * 32 + cta) * 2, make_uint4(acc, tag, acc, tag));
    }
    // wait for all owner columns in my mailbox
    if (!sentinel) {
      if (t < owners) {
        uint4 v;
        while (true) {
          v = ld_u4(mbox + ((size_t)cta * 32 + t) * 2);
          if (v.y == tag && v.w == tag) break;
          if (sleep_ns) __nanosleep(sleep_ns);
        }
      }
    } else {
      if (t == 0) {
        uint4 v;
        while (true) {
          v = ld_u4(mbox + ((size_t)cta * 32 + (owners - 1)) * 2);
          if (v.y == tag && v.w == tag) break;
          if (sleep_ns) __nanosleep(sleep_ns);
        }
      }
      __syncthreads();
      if (t < owners) {
        uint4 v;
        while (true) {
          v = ld_u4(mbox + ((size_t)cta * 32 + t) * 2);
          if (v.y == tag && v.w == tag) break;
        }
      }
    }
    __syncthreads();
  }
  if (cta == 0 && t == 0) *cycles = clock64() - t0;
}

// push all-gather: ONE hop.  Every CTA pushes `words` tagged 16-byte words to the private inbox of
// every CTA (inbox[dst][src][w]); every CTA then polls its own inbox (G*words words).
__global__ void push_kernel(uint4* inbox, int words, int iters, long long* cycles) {
  const int G = gridDim.x, cta = blockIdx.x, t = threadIdx.x, T = blockDim.x;
  long long t0 = clock64();
  const int total = G * words;
  for (int it = 1; it <= iters; ++it) {
    const unsigned tag = (unsigned)it;
    uint4* buf = inbox + (size_t)(it & 1) * 64 * 64 * 64;  // double-buffered: nobody is 2 iterations ahead
    for (int i = t; i < total; i += T) {  // i = dst * words + w
      const int dst = i / words, w = i - dst * words;
      st_u4(buf + ((size_t)dst * G + cta) * words + w, make_uint4(cta, tag, w, tag));
    }
    unsigned acc = 0;
    for (int i = t; i < total; i += T) {  // i = src * words + w
      uint4 v;
      const uint4* src = buf + (size_t)cta * G * words + i;
      do { v = ld_u4(src); } while (v.y != tag || v.w != tag);
      acc += v.x;
    }
    if (acc == 0xffffffffu) cycles[1] = acc;
    __syncthreads();
  }
  if (cta == 0 && t == 0) *cycles = clock64() - t0;
}

// cluster all-gather through distributed shared memory: every CTA of a cluster writes `words`
// float4 into its slot of every peer's shared memory (st.shared::cluster), then one cluster barrier
// (arrive.release / wait.acquire); second variant adds the L2 exchange among cluster leaders
// (tagged words, all leaders gather all leaders) and a DSMEM broadcast of the result.
template <int CS>
__global__ void __cluster_dims__(CS, 1, 1) cluster_allgather_kernel(int words, int iters, int cross, uint4* lead_buf,
                                                                    long long* cycles) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float4 slots[];  // [2 parity][CS][words] + [words] result
  const int t = threadIdx.x, rank = cluster.block_rank();
  const int nclusters = gridDim.x / CS, cid = blockIdx.x / CS;
  float4* result = slots + 2 * CS * words;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    float4* base = slots + (size_t)(it & 1) * CS * words;
    // push my words to every peer (including myself)
    for (int i = t; i < CS * words; i += blockDim.x) {
      const int dst = i / words, w = i - dst * words;
      float4* remote = cluster.map_shared_rank(base + (size_t)rank * words + w, dst);
      *remote = make_float4((float)it, (float)rank, (float)w, 1.f);
    }
    cluster.sync();
    float4 acc = make_float4(0, 0, 0, 0);
    if (t < words)
      for (int r = 0; r < CS; ++r) {
        const float4 v = base[(size_t)r * words + t];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    if (cross) {
      const unsigned tag = (unsigned)it;
      uint4* buf = lead_buf + (size_t)(it & 1) * 64 * 64;
      if (rank == 0 && t < words)  // leader publishes the cluster sum
        st_u4(buf + (size_t)cid * 64 + t, make_uint4(__float_as_uint(acc.x), tag, __float_as_uint(acc.y), tag));
      if (rank == 0) {             // every leader gathers all leaders
        float s = 0.f;
        for (int i = t; i < nclusters * words; i += blockDim.x) {
          const int c = i / words, w = i - c * words;
          uint4 v;
          do { v = ld_u4(buf + (size_t)c * 64 + w); } while (v.y != tag || v.w != tag);
          s += __uint_as_float(v.x);
        }
        // broadcast to the cluster through DSMEM
        if (t < words)
          for (int dst = 0; dst < CS; ++dst) *cluster.map_shared_rank(result + t, dst) = make_float4(s, 0, 0, 0);
      }
      cluster.sync();
      if (t < words && result[t].x == -1.f) cycles[1] = 1;
    } else if (t < words && acc.x == -1.f) {
      cycles[1] = 1;
    }
  }
  if (blockIdx.x == 0 && t == 0) *cycles = clock64() - t0;
}

// ---- dependent-load latency and load overlap: one thread, lines written long ago by another SM
template <int MODE>  // 0: ld.volatile  1: ld.global.cg (weak)  2: ld.relaxed.gpu
__device__ __forceinline__ uint4 ld_mode(const uint4* p) {
  uint4 v;
  if (MODE == 0) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  if (MODE == 1) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  if (MODE == 2) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// NLOADS independent 16-byte loads (different 128-byte lines) issued back to back, then consumed;
// repeated `iters` times with a data dependence between repetitions.
template <int MODE, int NLOADS>
__global__ void loadlat_kernel(const uint4* buf, int iters, long long* cycles, unsigned* sink) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint4 v[NLOADS];
#pragma unroll
    for (int k = 0; k < NLOADS; ++k) v[k] = ld_mode<MODE>(buf + (size_t)((k * 8 + (acc & 1))));
#pragma unroll
    for (int k = 0; k < NLOADS; ++k) acc += v[k].x;
  }
  *cycles = clock64() - t0;
  *sink = acc;
}

// ---- store issue: NST back-to-back 16-byte stores to different lines by one thread; if a flavour
// is serialised (each store waits for the previous one's acknowledgement) the time per iteration
// grows with NST * latency instead of NST * issue cost.
template <int MODE, int NST>  // 0 st.volatile  1 st.relaxed.gpu  2 st.relaxed.sys  3 st.global.cg  4 st.global
__global__ void storeissue_kernel(uint4* buf, int iters, long long* cycles) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < NST; ++k) {
      uint4* p = buf + (size_t)k * 8 + (it & 1);
      const unsigned v = (unsigned)it;
      if (MODE == 0) asm volatile("st.volatile.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
      if (MODE == 1) asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
      if (MODE == 2) asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
      if (MODE == 3) asm volatile("st.global.cg.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
      if (MODE == 4) asm volatile("st.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
    }
  }
  *cycles = clock64() - t0;
}

// ---- counter grid barrier --------------------------------------------------------------------
__global__ void gridbar_kernel(unsigned* ctr, int iters, long long* cycles) {
  const unsigned G = gridDim.x;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      } while ((int)(v - (unsigned)it * G) < 0);
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = clock64() - t0;
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, dev));
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  std::printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d}\n", pr.name, pr.multiProcessorCount, khz);
  unsigned* flags;
  long long* cyc;
  CK(cudaMalloc(&flags, 1 << 20));
  CK(cudaMallocManaged(&cyc, 2 * sizeof(long long)));
  const int iters = 20000;

  auto coop = [&](const void* fn, int grid, int block, void** args) {
    CK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(block), args, 0, 0));
    CK(cudaDeviceSynchronize());
  };

  // ping-pong
  const char* names[] = {"volatile(sys)", "relaxed.gpu", "release/acquire.gpu"};
  for (int peer : {1, 2, 73, 74, 147}) {
    for (int f = 0; f < 3; ++f) {
      CK(cudaMemset(flags, 0, 1 << 20));
      int it = iters;
      void* args[] = {&flags, &peer, &it, &cyc};
      const void* fn = f == 0 ? (const void*)pingpong_kernel<0> : f == 1 ? (const void*)pingpong_kernel<1> : (const void*)pingpong_kernel<2>;
      coop(fn, 148, 32, args);
      std::printf("{\"test\": \"pingpong\", \"peer_block\": %d, \"flavor\": \"%s\", \"round_trip_cycles\": %.1f}\n", peer,
                  names[f], (double)*cyc / iters);
    }
  }
  // fan-in/fan-out
  uint4 *part, *mbox;
  CK(cudaMalloc(&part, 148 * 256 * 2));
  CK(cudaMalloc(&mbox, 148 * 256 * 2));
  for (int G : {2, 8, 32, 64, 148}) {
    for (int priv : {0, 1}) {
      for (int sleep_ns : {0, 100}) {
        CK(cudaMemset(part, 0, 148 * 256 * 2));
        CK(cudaMemset(mbox, 0, 148 * 256 * 2));
        int it = iters;
        void* args[] = {&part, &mbox, &priv, &it, &sleep_ns, &cyc};
        coop((const void*)fan_kernel, G, 32, args);
        std::printf("{\"test\": \"fan\", \"G\": %d, \"private_mailboxes\": %d, \"poll_sleep_ns\": %d, \"cycles_per_iter\": %.1f}\n", G, priv,
                    sleep_ns, (double)*cyc / iters);
      }
    }
  }
  // cluster / DSMEM all-gather
  {
    uint4* lb2;
    CK(cudaMalloc(&lb2, 2 * 64 * 64 * 16));
    const int words = 25;
    auto run_cluster = [&](auto kern, int cs, int grid, int cross) {
      CK(cudaMemset(lb2, 0, 2 * 64 * 64 * 16));
      int it = 5000, w = words;
      size_t smem = (size_t)(2 * cs * words + words) * sizeof(float4);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeCooperative;
      at[0].val.cooperative = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, kern, w, it, cross, lb2, cyc);
      if (e != cudaSuccess) { std::printf("{\"test\": \"cluster\", \"cs\": %d, \"error\": \"%s\"}\n", cs, cudaGetErrorString(e)); cudaGetLastError(); return; }
      CK(cudaDeviceSynchronize());
      std::printf("{\"test\": \"cluster_allgather\", \"cluster_size\": %d, \"grid\": %d, \"cross_cluster_L2\": %d, \"words16B\": %d, \"cycles_per_iter\": %.1f}\n",
                  cs, grid, cross, words, (double)*cyc / it);
    };
    run_cluster(cluster_allgather_kernel<2>, 2, 2, 0);
    run_cluster(cluster_allgather_kernel<4>, 4, 4, 0);
    run_cluster(cluster_allgather_kernel<8>, 8, 8, 0);
    run_cluster(cluster_allgather_kernel<8>, 8, 32, 0);
    run_cluster(cluster_allgather_kernel<8>, 8, 32, 1);
    run_cluster(cluster_allgather_kernel<8>, 8, 144, 0);
    run_cluster(cluster_allgather_kernel<8>, 8, 144, 1);
    run_cluster(cluster_allgather_kernel<4>, 4, 148, 1);
    run_cluster(cluster_allgather_kernel<2>, 2, 148, 1);
  }
  // push all-gather (one hop)
  {
    uint4* ib;
    CK(cudaMalloc(&ib, (size_t)2 * 64 * 64 * 64 * 16));
    for (int G : {2, 8, 16, 32, 48, 64}) {
      for (int words : {1, 50}) {
        CK(cudaMemset(ib, 0, (size_t)2 * 64 * 64 * 64 * 16));
        int it = 5000;
        void* args[] = {&ib, &words, &it, &cyc};
        coop((const void*)push_kernel, G, 512, args);
        std::printf("{\"test\": \"push_allgather\", \"G\": %d, \"words16B_per_cta\": %d, \"cycles_per_iter\": %.1f}\n", G, words,
                    (double)*cyc / it);
      }
    }
  }
  // fan2: parallel gather, several owners, mailbox polling variants
  {
    uint4 *p2, *m2;
    CK(cudaMalloc(&p2, 148 * 32 * 32));
    CK(cudaMalloc(&m2, 148 * 32 * 32));
    for (int G : {8, 32, 148}) {
      for (int owners : {1, 25}) {
        for (int sentinel : {0, 1}) {
          for (int sleep_ns : {0, 200}) {
            if (owners > G) continue;
            CK(cudaMemset(p2, 0, 148 * 32 * 32));
            CK(cudaMemset(m2, 0, 148 * 32 * 32));
            int it = 5000;
            void* args[] = {&p2, &m2, &owners, &sentinel, &sleep_ns, &it, &cyc};
            coop((const void*)fan2_kernel, G, 256, args);
            std::printf("{\"test\": \"fan2\", \"G\": %d, \"owners\": %d, \"sentinel\": %d, \"poll_sleep_ns\": %d, \"cycles_per_iter\": %.1f}\n",
                        G, owners, sentinel, sleep_ns, (double)*cyc / it);
          }
        }
      }
    }
  }
  // dependent load latency / overlap of several loads, by flavour
  {
    uint4* lb;
    CK(cudaMalloc(&lb, 1 << 16));
    CK(cudaMemset(lb, 0, 1 << 16));
    unsigned* sink;
    CK(cudaMalloc(&sink, 4));
    const char* mn[] = {"ld.volatile", "ld.global.cg", "ld.relaxed.gpu"};
    int it = 20000;
#define RUN_LL(MODE, N)                                                                              \
    loadlat_kernel<MODE, N><<<1, 32>>>(lb, it, cyc, sink);                                             \
    CK(cudaDeviceSynchronize());                                                                     \
    std::printf("{\"test\": \"loadlat\", \"flavor\": \"%s\", \"loads_in_flight\": %d, \"cycles_per_iter\": %.1f}\n", mn[MODE], N, (double)*cyc / it);
    RUN_LL(0, 1) RUN_LL(0, 2) RUN_LL(0, 4) RUN_LL(0, 8)
    RUN_LL(1, 1) RUN_LL(1, 2) RUN_LL(1, 4) RUN_LL(1, 8)
    RUN_LL(2, 1) RUN_LL(2, 2) RUN_LL(2, 4) RUN_LL(2, 8)
  }
  // store issue cost by flavour
  {
    uint4* sb;
    CK(cudaMalloc(&sb, 1 << 16));
    const char* mn[] = {"st.volatile", "st.relaxed.gpu", "st.relaxed.sys", "st.global.cg", "st.global"};
    int it = 20000;
#define RUN_ST(MODE, N)                                                                             \
    storeissue_kernel<MODE, N><<<1, 32>>>(sb, it, cyc);                                             \
    CK(cudaDeviceSynchronize());                                                                    \
    std::printf("{\"test\": \"storeissue\", \"flavor\": \"%s\", \"stores_per_iter\": %d, \"cycles_per_iter\": %.1f}\n", mn[MODE], N, (double)*cyc / it);
    RUN_ST(0, 1) RUN_ST(0, 4) RUN_ST(0, 16) RUN_ST(1, 1) RUN_ST(1, 4) RUN_ST(1, 16) RUN_ST(2, 4) RUN_ST(2, 16)
    RUN_ST(3, 1) RUN_ST(3, 4) RUN_ST(3, 16) RUN_ST(4, 4) RUN_ST(4, 16)
  }
  // counter barrier
  for (int G : {2, 8, 32, 64, 148}) {
    CK(cudaMemset(flags, 0, 1 << 20));
    int it = iters;
    void* args[] = {&flags, &it, &cyc};
    coop((const void*)gridbar_kernel, G, 128, args);
    std::printf("{\"test\": \"gridbar\", \"G\": %d, \"cycles_per_barrier\": %.1f}\n", G, (double)*cyc / iters);
  }
  return 0;
}
