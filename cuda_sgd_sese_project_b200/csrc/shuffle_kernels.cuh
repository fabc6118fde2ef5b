// shuffle_kernels.cuh -- the reference's per-epoch permutation, bit-exact, on the device.
//
// The reference shuffles its index vector on the HOST with std::random_shuffle driven by glibc
// rand() (main.cu:90, :207): for i = 1..n-1: j = rand() % (i+1); if (i != j) swap(v[i], v[j]).
// That loop is sequential and is the true epoch bottleneck at scale (23 ms at
// 4M rows on this box's host vs a few ms of step loop).  The SAME permutation is produced here in
// parallel.  The host only computes ~1000 start states of the rand() stream per epoch by jump-ahead
// (host_rng.hpp, ~1 ms) and uploads them (tens of KB); the stream itself is regenerated on the
// device inside K1.  Five small kernels:
//
//   Look at the values instead of the loop.  Position i is untouched until step i, so the value
//   in[i] sits at i until time i, when it moves to j_i (and whatever was at j_i moves to i).  A
//   value sitting at position p is thrown out the next time p is some later step's target ("hit"):
//   a hit of p at step h moves it to position h, where it again waits for the first hit of h ...
//   So with   hits(p) = { i > p : j_i = p }   sorted ascending, the trajectory of in[i] is
//       i --(time i)--> j_i --(successor of i in hits(j_i))--> h --> first(hits(h)) --> ...
//   until a position without a later hit is reached: its final place.  (A self swap j_i = i is a
//   no-op, exactly as the `if (i != j)` of libstdc++.)  Chains are short: first(hits(p)) is about
//   2p in expectation.
//
//   K1 targets : regenerate the stream chunk by chunk, j_i = rand() % (i+1), count hits per position
//   K2 scan    : exclusive prefix sum of the counts  -> CSR offsets
//   K3 fill    : scatter i into its target's list (atomic cursor that starts at the list's offset)
//   K4 order   : sort every (tiny) list ascending, record first(hits(p)) and every hit's successor nxt[i]
//   K5 resolve : one thread per value follows its trajectory (nxt[i], then first[..]) and writes out[final] = in[i]
// (~6.5 random 32-byte accesses per element in total: they, not the arithmetic, are the cost)
//
// tests/test_gpu_shuffle.py checks it against the host loop and the golden vectors of
// tests/golden/thirdparty_kat.json (real glibc + libstdc++ outputs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200sgd {

// K1 with the stream generated on the device: thread c regenerates chunk c of this epoch's rand()
// stream (values k in [c*L, (c+1)*L) of the epoch, consumed by steps i = k+1) from its 31-word
// start state (host_rng.hpp chunk_states) and turns every value into a swap target at once.
// x[k] = x[k-31] + x[k-3] (mod 2^32); rand() = x >> 1.  31 values per unrolled block keep the
// window in registers.  cnt must be zero on entry.
__global__ void __launch_bounds__(128) shuf_targets_stream_kernel(const uint32_t* __restrict__ states, int64_t L,
                                                                  int64_t nchunks, int64_t n,
                                                                  uint32_t* __restrict__ tgt, int32_t* __restrict__ cnt) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0) tgt[0] = 0;
  if (c >= nchunks) return;
  uint32_t r[31];
#pragma unroll
  for (int i = 0; i < 31; ++i) r[i] = states[c * 31 + i];
  int64_t k = c * L;                                         // index into the epoch's stream
  const int64_t kend = (k + L < n - 1) ? (k + L) : (n - 1);  // the epoch consumes n-1 values
  while (k < kend) {
#pragma unroll
    for (int i = 0; i < 31; ++i) {
      const uint32_t v = r[i] + (i < 3 ? r[28 + i] : r[i - 3]);  // r[i-3] already holds a NEW value
      r[i] = v;
      if (k < kend) {
        const int64_t step = k + 1;
        const uint32_t j = (v >> 1) % (uint32_t)(step + 1);  // rand() % ((i - first) + 1), stl_algo.h
        if ((int64_t)j != step) atomicAdd(cnt + j, 1);
        tgt[step] = j;
      }
      ++k;
    }
  }
}

// the reference's ind_vector before the first shuffle (main.cu:421-424), written on the device
__global__ void __launch_bounds__(256) shuf_iota_kernel(int32_t* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int32_t)i;
}

constexpr int kScanBlock = 1024;
constexpr int kScanPerThread = 4;
constexpr int kScanTile = kScanBlock * kScanPerThread;

__device__ __forceinline__ int32_t block_exclusive_scan(int32_t v, int32_t* warp_tot, int32_t* total) {
  // exclusive scan of one value per thread over a 1024-thread block
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t incl = v;
#pragma unroll
  for (int m = 1; m < 32; m <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, incl, m);
    if (lane >= m) incl += t;
  }
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  if (w == 0) {
    int32_t t = warp_tot[lane];
    int32_t ti = t;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const int32_t u = __shfl_up_sync(0xffffffffu, ti, m);
      if (lane >= m) ti += u;
    }
    warp_tot[lane] = ti - t;  // exclusive over warps
    if (lane == 31) *total = ti;
  }
  __syncthreads();
  const int32_t r = warp_tot[w] + incl - v;
  __syncthreads();
  return r;
}

// phase 1: per-tile totals
__global__ void __launch_bounds__(kScanBlock) scan_tile_sums_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                    int32_t* __restrict__ tile_sum) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t total;
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanPerThread;
  int32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k)
    if (base + k < n) s += in[base + k];
  block_exclusive_scan(s, warp_tot, &total);
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// phase 2: exclusive scan of the tile totals (single block, loops); also writes the grand total
__global__ void __launch_bounds__(kScanBlock) scan_tile_offsets_kernel(int32_t* __restrict__ tile_sum, int32_t ntiles,
                                                                       int32_t* __restrict__ grand_total) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t total;
  int32_t carry = 0;
  for (int base = 0; base < ntiles; base += kScanBlock) {
    const int i = base + threadIdx.x;
    const int32_t v = i < ntiles ? tile_sum[i] : 0;
    const int32_t ex = block_exclusive_scan(v, warp_tot, &total);
    if (i < ntiles) tile_sum[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

// phase 3: exclusive scan inside every tile plus the tile offset; out has n+1 entries
__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                const int32_t* __restrict__ tile_off,
                                                                const int32_t* __restrict__ grand_total,
                                                                int32_t* __restrict__ out) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t total;
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanPerThread;
  int32_t v[kScanPerThread];
  int32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  int32_t ex = block_exclusive_scan(s, warp_tot, &total) + tile_off[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanPerThread; ++k) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *grand_total;
}

// cursor must hold a copy of the CSR offsets (`start`) on entry: one random access (the atomic) finds the
// slot, instead of a read of start[j] plus an atomic on a zeroed counter.  Four elements per thread in
// flight: the kernel mostly runs beside the persistent step kernel, one block per SM.
__global__ void __launch_bounds__(256) shuf_fill_kernel(const uint32_t* __restrict__ tgt, int64_t n,
                                                        int32_t* __restrict__ cursor, int32_t* __restrict__ hits) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i0 < n; i0 += 4 * stride) {
    uint32_t j[4];
    int32_t slot[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      j[u] = i < n ? tgt[i] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      slot[u] = (i < n && (int64_t)j[u] != i) ? atomicAdd(cursor + j[u], 1) : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (slot[u] >= 0) hits[slot[u]] = (int32_t)(i0 + u * stride);
  }
}

// Sorts every (tiny) list, records first(hits(p)) and, for every hit i of p, its SUCCESSOR in the list:
// nxt[i] = the next later step that targets the same position, or -1 (i is the last one: the value that
// moved to p at time i stays there).  The resolve pass then reads nxt[i] in order instead of looking the
// list of its target up (start[j], hits[..]: two random reads per value less).
__global__ void __launch_bounds__(256) shuf_order_kernel(int64_t n, const int32_t* __restrict__ start,
                                                         int32_t* __restrict__ hits, int32_t* __restrict__ first,
                                                         int32_t* __restrict__ nxt) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
       p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t lo = start[p], hi = start[p + 1];
    for (int32_t a = lo + 1; a < hi; ++a) {  // insertion sort: lists have ~1 entry on average
      const int32_t key = hits[a];
      int32_t b = a - 1;
      while (b >= lo && hits[b] > key) {
        hits[b + 1] = hits[b];
        --b;
      }
      hits[b + 1] = key;
    }
    first[p] = hi > lo ? hits[lo] : -1;
    for (int32_t a = lo; a < hi; ++a) nxt[hits[a]] = (a + 1 < hi) ? hits[a + 1] : -1;
  }
}

__global__ void __launch_bounds__(256) shuf_resolve_kernel(const uint32_t* __restrict__ tgt, int64_t n,
                                                           const int32_t* __restrict__ nxt,
                                                           const int32_t* __restrict__ first,
                                                           const int32_t* __restrict__ in, int32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
    int32_t pos[4], val[4];
    bool live[4];  // still travelling
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      pos[u] = (int32_t)i;
      live[u] = i < n;
      val[u] = 0;
      if (i < n) {
        val[u] = in ? in[i] : (int32_t)i;  // in == nullptr: the position map itself (out[final] = i)
        const uint32_t j = tgt[i];
        if (i >= 1 && (int64_t)j != i) {
          // moved to j at time i; thrown out again by the next hit of j after time i, if any
          const int32_t nx = nxt[i];
          if (nx < 0) { pos[u] = (int32_t)j; live[u] = false; }
          else pos[u] = nx;
        }
      }
    }
    // from then on: the first hit of every position it lands on (all four chains advance together)
    bool any = true;
    while (any) {
      int32_t nx[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) nx[u] = live[u] ? first[pos[u]] : -1;
      any = false;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (nx[u] >= 0) { pos[u] = nx[u]; any = true; }
        else live[u] = false;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * stride < n) out[pos[u]] = val[u];
  }
}

// ------------------------------------------------------------------------------------------------
// Several GPUs (one process each): the ranks SHARE the permutation work instead of every rank
// recomputing it.  The loop above moves values around without looking at them, so epoch k's shuffle is
// a position map  src_k[p] = "the slot whose value ends up at p"  that depends on the rand() stream of
// epoch k only:  perm_k[p] = perm_{k-1}[src_k[p]].  Rank (k-1) mod N computes src_k with the five
// kernels above (in = nullptr) -- N epochs ahead, on its own stream -- into a buffer every rank has
// mapped (the exchange block, NVLink); every rank then applies it with one gather pass, reading the
// owner's buffer directly over NVLink.  Hand-shake: 32-bit sequence numbers written with system-scope
// release stores into the consumers' / the owner's exchange block and polled by one tiny wait kernel
// (a spinning kernel of one warp cannot keep the cooperative step kernel off the SMs).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) shuf_wait_flags_kernel(const uint32_t* flags, int first, int count, int stride,
                                                             uint32_t want, unsigned int* err, unsigned long long timeout_ns) {
  const int i = threadIdx.x;
  if (i >= count) return;
  const uint32_t* f = flags + first + i * stride;
  unsigned long long t0 = 0;
  unsigned int spins = 0;
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if ((int32_t)(v - want) >= 0) return;
    if ((++spins & 255u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > timeout_ns) { atomicCAS(err, 0u, 5u); return; }
    }
    __nanosleep(200);
  }
}

struct PeerFlagPtrs { uint32_t* p[8]; };

// thread r stores `value` into flag `idx` of target r (everything this stream did before is visible first)
__global__ void __launch_bounds__(32) shuf_set_flags_kernel(PeerFlagPtrs t, int ntargets, int idx, uint32_t value) {
  const int r = threadIdx.x;
  if (r >= ntargets || t.p[r] == nullptr) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(t.p[r] + idx), "r"(value) : "memory");
}

// out[p] = in[src[p]]: src is read once, in order (possibly from a peer GPU), `in` at random
__global__ void __launch_bounds__(256) shuf_apply_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ in,
                                                         int32_t* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
    int32_t s[4], v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) s[u] = (i + u * stride < n) ? __ldcs(src + i + u * stride) : 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // two sectors instead of the whole line per random read (shuffle2_kernels.cuh, ld_sparse)
      asm volatile("ld.global.L2::64B.b32 %0, [%1];" : "=r"(v[u]) : "l"(in + s[u]));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < n) out[i + u * stride] = v[u];
  }
}

}  // namespace b200sgd
