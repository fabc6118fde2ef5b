// shuffle2_kernels.cuh -- the reference's per-epoch permutation, bit-exact, on the device: the BACKWARD
// formulation (round 2).  Same contract as shuffle_kernels.cuh (std::random_shuffle driven by glibc
// rand(), main.cu:90 / :207:  for i = 1..n-1: j_i = rand() % (i+1); if (i != j_i) swap(v[i], v[j_i])),
// 1.5 random DRAM reads per element (0.5 on the chains, 1 for the final gather) instead of 6.5 random
// accesses; 12 instead of 30 GB of DRAM traffic per shuffle of 64M rows.
//
//   Ask every final position p where its value came from.  Walk the loop backwards from time n-1: the
//   content of p was last replaced by the LATEST step that targets p, h = last(hits(p)), hits(p) =
//   { i > p : j_i = p } -- and what step h put there is v[h], which no earlier step can have touched
//   (step i only touches positions <= i): src(p) = h, one look-up, for the half of all positions that
//   are ever hit.  A position nobody hits keeps what its OWN step p put there: the content of j_p just
//   before time p, i.e. what the latest step BEFORE p that targets j_p put there -- the predecessor of p
//   in hits(j_p) -- or, if p is the first hit of j_p, what j_p's own step put there ... :
//
//       last[p] = max hits(p)                          (or -1)
//       R[t]    = predecessor of t in hits(j_t)        (>= 0: the source), t itself if j_t = t (self swap,
//                 a no-op exactly as libstdc++'s `if (i != j)`), else ~j_t  (< 0: go on at step j_t)
//       src(p)  = last[p] >= 0 ? last[p] : follow r = R[p], R[~r], ... until r >= 0
//
//   The chain strictly decreases (j_t < t) and ends at R[0] = 0; it takes 0.5 random reads per element
//   on average.  out[p] = in[src(p)], or src itself (the position map several GPUs share, see
//   shuffle_kernels.cuh).
//
//   hits(p) are built by regrouping the pairs (j_i, i) by position in two steps, without a global atomic or
//   a random DRAM access per element (default pipeline, B200SGD_SHUFFLE=5; b200sgd_api.cu enqueue_perm_kernels):
//     K0  expand  : 64 sub-chunk start states of the rand() stream per host state (jump-ahead polynomials)
//     K1a stream<0,1>: regenerate the stream, count the hits per COARSE bucket (2^14..2^18 positions) in
//                   shared memory; the block's counts go into its column of a (bucket, block) matrix
//         scan    : exclusive scan of the matrix in memory order = one region per (bucket, block)
//     K1b stream<1,1>: regenerate again, scatter the pairs into the regions (shared-memory cursors)
//     K1c refine  : one block per coarse bucket: regroup its pairs by granule (256 positions)
//     K2  link    : runs of granules in shared memory: bin per position, sort the (tiny) lists, write
//                   last[] in order and the entries (i, R[i]) into the coarse range i >> 12
//     K2b place   : one block per range: entries put in place in shared memory, R[] written in order
//     K3  resolve : src(p), then out[p] = in[src(p)]
//   Other strategies (A/B, tests): 2 = pairs straight into granule regions through global cursors, R written
//   in place; 3 = coarse buckets through global cursors; 4 = as 5 with R written in place.
//
// The kernels use block-level primitives only (__syncthreads, shared / global atomicAdd): tests/emu/ runs
// the SAME source on the host (one OS thread per CUDA thread) against the host loop -- the GPU pool has no
// compute-sanitizer -- and tests/test_gpu_shuffle.py checks them on the device against the golden vectors
// of real glibc + libstdc++.
#pragma once
#ifndef B200SGD_HOST_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace b200sgd {

// A random 4-byte read.  A plain ld.global of a line that misses pulls (almost) the whole 128-byte line out
// of DRAM on B200 -- measured with tools/microbench_gather.cu: 3.45 L2 sectors and 92 bytes of DRAM traffic
// per random 4-byte load; .cs / L1::no_allocate / L1::evict_first: 4 sectors, 119 bytes;
// cudaLimitMaxL2FetchGranularity = 32: no change -- while the L2::64B qualifier fetches two sectors: 55 bytes.
__device__ __forceinline__ int32_t ld_sparse(const int32_t* p) {
#ifdef B200SGD_HOST_EMU
  return *p;
#else
  int32_t v;
  asm volatile("ld.global.L2::64B.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
#endif
}

constexpr int kGranLog = 8;                 // positions per granule = 256
constexpr int kGran = 1 << kGranLog;
constexpr int kLinkThreads = 256;
constexpr int kLinkGranules = 4;            // granules of one block job
constexpr int kLinkPos = kLinkGranules * kGran;  // 1024 positions: 4 per thread in the scan
// Hits staged per run.  One granule must fit: granule 0 expects 256 * (ln(n / 256) + 1) hits -- 3.4 k at
// 64 M rows, 4.3 k at 2^31 (sigma ~ 65) -- anything larger raises the context's error word.  27.5 KB of
// shared memory and <= 40 registers: a block fits on an SM BESIDE a CTA of the persistent step kernel.
constexpr int kLinkCap = 5632;

// K0: more chunks than the host pays for.  The host hands over one start state per chunk of L values
// (jump-ahead, ~0.6 us each: 8192 at most); thread (c, s) derives the state of sub-chunk s of chunk c,
// s * (L / S) values further on, with the same polynomial trick (host_rng.hpp apply()): subjump[s-1] =
// z^(s L / S) mod p, x[m + jump] = sum_j a_j x[m + j].  S x as many threads then regenerate the stream
// (the atomics' round trips are what the stream kernels wait for).
__global__ void __launch_bounds__(128) shuf2_expand_states_kernel(const uint32_t* __restrict__ states, int64_t nchunks,
                                                                  const uint32_t* __restrict__ subjump, int S,
                                                                  uint32_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks * S) return;
  const int64_t c = t / S;
  const int s = (int)(t - c * S);
  uint32_t ext[61];
  for (int i = 0; i < 31; ++i) ext[i] = states[c * 31 + i];
  if (s == 0) {
    for (int i = 0; i < 31; ++i) out[t * 31 + i] = ext[i];
    return;
  }
  for (int i = 31; i < 61; ++i) ext[i] = ext[i - 31] + ext[i - 3];
  const uint32_t* a = subjump + (s - 1) * 31;
  for (int i = 0; i < 31; ++i) {
    uint32_t acc = 0;
    for (int j = 0; j < 31; ++j) acc += a[j] * ext[i + j];
    out[t * 31 + i] = acc;
  }
}

constexpr int kCoarseMax = 4096;  // coarse buckets at most: their counters fit 16 KB of shared memory

// K1a (MODE 0) / K1b (MODE 1): thread c regenerates chunk c of this epoch's rand() stream (values k in
// [c*L, (c+1)*L), consumed by steps i = k+1) from its 31-word start state (host_rng.hpp chunk_states):
// x[k] = x[k-31] + x[k-3] (mod 2^32), rand() = x >> 1, j = rand() % (i+1) (stl_algo.h).
//   MODE 0: bins[j >> shift] += 1 for every hit
//   MODE 1: bins[j >> shift] (the offset of the bin's region) hands out the slot of the pair (j, i); self
//           swaps and step 0 get their R entry here (R != nullptr).  The atomics of a third of a round (11
//           values) are issued back to back so that their round trips overlap; 88 registers at most, so
//           that a block of 128 fits beside a CTA of the step kernel.
//   BLOCKBINS = false: bins are global counters / cursors -- one L2 atomic per element, and the L2 does
//           ~30 G of them per second whatever they hit (measured: 64 M in 2.7 ms; 8 MB of DRAM traffic).
//   BLOCKBINS = true : the bins (<= kCoarseMax coarse buckets) live in shared memory.  MODE 0 leaves the
//           block's counts in column `block` of the matrix bins[bucket * vblocks + block]; the exclusive
//           scan of that matrix in memory order gives every (bucket, block) its own region, MODE 1
//           starts its shared cursors there: no global atomic at all, and a deterministic layout.
template <int MODE, bool BLOCKBINS>
__global__ void __launch_bounds__(128) shuf2_stream_kernel(const uint32_t* __restrict__ states, int64_t L,
                                                           int64_t nchunks, int64_t n, int shift, int nbins, int vblocks,
                                                           int32_t* __restrict__ gbins, uint2* __restrict__ pairs,
                                                           int32_t* __restrict__ R) {
  __shared__ int32_t s_bins[BLOCKBINS ? kCoarseMax : 1];
  int32_t* bins = BLOCKBINS ? s_bins : gbins;
  // `vblocks` virtual blocks of 128 chunks (= the columns of the count matrix), taken in turns by the grid:
  // the map of several GPUs runs with one block per SM so that it never keeps the next epoch's cooperative
  // step kernel from starting
  for (int vb = blockIdx.x; vb < vblocks; vb += gridDim.x) {
    if (BLOCKBINS) {
      __syncthreads();  // the previous virtual block is done with the bins
      for (int b = threadIdx.x; b < nbins; b += blockDim.x)
        s_bins[b] = MODE == 0 ? 0 : gbins[(int64_t)b * vblocks + vb];
      __syncthreads();
    }
    const int64_t c = (int64_t)vb * blockDim.x + threadIdx.x;
    if (MODE == 1 && c == 0 && R != nullptr) R[0] = 0;
    uint32_t r[31];
    uint32_t k = 0, kend = 0;  // 32-bit step arithmetic (n < 2^31): k = index into the epoch's stream
    if (c < nchunks) {
#pragma unroll
      for (int i = 0; i < 31; ++i) r[i] = states[c * 31 + i];
      k = (uint32_t)(c * L);
      kend = (uint32_t)((c * L + L < n - 1) ? (c * L + L) : (n - 1));  // the epoch consumes n-1 values
    }
    while (k < kend) {
#pragma unroll
      for (int h = 0; h < 3; ++h) {  // values 0..10, 11..21 and 22..30 of the round
        constexpr int kHalf = 11;
        uint32_t jj[kHalf];
        int32_t slot[kHalf];
#pragma unroll
        for (int u = 0; u < kHalf; ++u) {
          const int i = h * kHalf + u;
          if (i < 31) {
            const uint32_t v = r[i] + (i < 3 ? r[28 + i] : r[i - 3]);  // r[i-3] already holds a NEW value
            r[i] = v;
            jj[u] = (v >> 1) % (k + (uint32_t)i + 2u);  // step = k + i + 1; j = rand() % (step + 1)
          }
        }
#pragma unroll
        for (int u = 0; u < kHalf; ++u) {
          const int i = h * kHalf + u;
          slot[u] = -1;
          if (i < 31 && k + (uint32_t)i < kend && jj[u] != k + (uint32_t)i + 1u) {
            if (MODE == 0) atomicAdd(bins + (jj[u] >> shift), 1);
            else slot[u] = atomicAdd(bins + (jj[u] >> shift), 1);
          }
        }
        if (MODE == 1) {
#pragma unroll
          for (int u = 0; u < kHalf; ++u) {
            const int i = h * kHalf + u;
            if (i < 31 && k + (uint32_t)i < kend) {
              const uint32_t step = k + (uint32_t)i + 1u;
              if (slot[u] >= 0) pairs[slot[u]] = make_uint2(jj[u], step);
              else if (R != nullptr) R[step] = (int32_t)step;  // self swap: the value stays
            }
          }
        }
      }
      k += 31;
    }
    if (BLOCKBINS && MODE == 0) {
      __syncthreads();
      for (int b = threadIdx.x; b < nbins; b += blockDim.x) gbins[(int64_t)b * vblocks + vb] = s_bins[b];
    }
  }
}

// exclusive scan of s[0 .. 1024) by a block of 256 threads (thread t owns entries 4t .. 4t+3); barriers only
__device__ __forceinline__ void block_scan_1024(int32_t* s, int32_t* s_part) {
  const int tid = threadIdx.x;
  int32_t v[4], sum = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[e] = s[4 * tid + e];
    sum += v[e];
  }
  s_part[tid] = sum;
  __syncthreads();
  for (int off = 1; off < kLinkThreads; off <<= 1) {
    const int32_t t = tid >= off ? s_part[tid - off] : 0;
    __syncthreads();
    s_part[tid] += t;
    __syncthreads();
  }
  int32_t ex = s_part[tid] - sum;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    s[4 * tid + e] = ex;
    ex += v[e];
  }
  __syncthreads();
}

// K1c refine (two-level scatter).  With 256 k granules the scatter of K1b has 256 k open write fronts --
// more than stay in L2 beside the row stream of the step kernel, so every 8-byte pair costs a sector
// fill and a write-back.  Instead K1b scatters into COARSE buckets of 2^cshift positions (<= 8192 of
// them, <= 1024 granules each: their write fronts stay in L2 and complete sector by sector), and this
// kernel, one block per coarse bucket, regroups a bucket by granule: count, scan, scatter -- the
// bucket's region is small and written by one block.  It also produces the granules' offsets gstart[].
// Bucket b's region starts at cstart[b * cstride] (cstride = the stream kernels' grid when the offsets
// are the scanned (bucket, block) matrix).
__global__ void __launch_bounds__(kLinkThreads) shuf2_refine_kernel(int64_t ngran, int64_t ncoarse, int cshift,
                                                                    const int32_t* __restrict__ cstart, int64_t cstride,
                                                                    const uint2* __restrict__ pairs1,
                                                                    uint2* __restrict__ pairs2,
                                                                    int32_t* __restrict__ gstart) {
  __shared__ int32_t s_h[1024];
  __shared__ int32_t s_part[kLinkThreads];
  const int tid = threadIdx.x;
  const int gshift = cshift - kGranLog;  // granules per coarse bucket = 2^gshift <= 1024
  for (int64_t b = blockIdx.x; b < ncoarse; b += gridDim.x) {
    const int32_t base = cstart[b * cstride], m = cstart[(b + 1) * cstride] - base;
    const int64_t g0 = b << gshift;
    for (int k = tid; k < 1024; k += kLinkThreads) s_h[k] = 0;
    __syncthreads();
    for (int h = tid; h < m; h += kLinkThreads) atomicAdd(&s_h[(int)((int64_t)(pairs1[base + h].x >> kGranLog) - g0)], 1);
    __syncthreads();
    block_scan_1024(s_h, s_part);
    for (int k = tid; k < (1 << gshift); k += kLinkThreads)
      if (g0 + k < ngran) gstart[g0 + k] = base + s_h[k];
    if (b == ncoarse - 1 && tid == 0) gstart[ngran] = base + m;
    __syncthreads();
    for (int h = tid; h < m; h += kLinkThreads) {
      const uint2 q = pairs1[base + h];
      pairs2[base + atomicAdd(&s_h[(int)((int64_t)(q.x >> kGranLog) - g0)], 1)] = q;
    }
    __syncthreads();
  }
}

// rcur[k] = k << rshift: the cursors of the coarse ranges of R (every step owns at most one entry)
__global__ void __launch_bounds__(256) shuf2_range_cursor_kernel(int32_t* __restrict__ rcur, int64_t nranges, int rshift) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nranges; k += (int64_t)gridDim.x * blockDim.x)
    rcur[k] = (int32_t)(k << rshift);
}

constexpr int kRangeLog = 12;  // steps per coarse range of R: 16 KB of shared memory in shuf2_place_kernel

// K2 link.  Job = kLinkGranules consecutive granules, cut into runs of as many granules as fit the
// staging capacity (low positions collect ~ln(n/p) hits each, high ones almost none).  Per run: count
// the hits per position, exclusive scan, bin, sort every list, then last[] (in order) and R[]:
//   RSCATTER = false: R[i] written in place -- one random 4-byte write (sector fill + write-back) per element
//   RSCATTER = true : the entries (i, R[i]) go to the coarse range i >> 12 of `rpairs` (cursor rcur),
//                     shuf2_place_kernel puts them in place: sequential traffic only
template <bool RSCATTER>
__global__ void __launch_bounds__(kLinkThreads) shuf2_link_kernel(int64_t n, int64_t ngran,
                                                                  const int32_t* __restrict__ gstart,
                                                                  const uint2* __restrict__ pairs,
                                                                  int32_t* __restrict__ last, int32_t* __restrict__ R,
                                                                  int32_t* __restrict__ rcur, uint2* __restrict__ rpairs,
                                                                  unsigned int* __restrict__ err) {
  __shared__ int32_t s_cnt[kLinkPos];
  __shared__ int32_t s_part[kLinkThreads];
  __shared__ int32_t s_list[kLinkCap];
  __shared__ int32_t s_gs[kLinkGranules + 1];
  const int tid = threadIdx.x;
  const int64_t njobs = (ngran + kLinkGranules - 1) / kLinkGranules;
  for (int64_t job = blockIdx.x; job < njobs; job += gridDim.x) {
    const int64_t gb = job * kLinkGranules;
    const int kn = (int)((ngran - gb < kLinkGranules) ? (ngran - gb) : kLinkGranules);
    if (tid <= kLinkGranules) s_gs[tid] = gstart[gb + (tid < kn ? tid : kn)];
    __syncthreads();
    int k0 = 0;
    while (k0 < kn) {  // every thread takes the same decisions: the barriers below are uniform
      int k1 = k0 + 1;
      while (k1 < kn && s_gs[k1 + 1] - s_gs[k0] <= kLinkCap) ++k1;
      const int32_t base = s_gs[k0], m = s_gs[k1] - base;
      const int64_t pos0 = (gb + k0) << kGranLog;
      const int64_t pend = ((gb + k1) << kGranLog) < n ? ((gb + k1) << kGranLog) : n;
      const int P = (int)(pend - pos0);  // <= kLinkPos
      if (m > kLinkCap) {                // a single granule beyond the staging capacity (see kLinkCap)
        if (tid == 0) atomicCAS(err, 0u, 6u);
        k0 = k1;
        continue;
      }
      for (int p = tid; p < kLinkPos; p += kLinkThreads) s_cnt[p] = 0;
      __syncthreads();
      // the first 4 pairs of a thread stay in registers for the binning pass; longer runs are re-read (L2)
      uint2 pr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int h = tid + u * kLinkThreads;
        pr[u] = make_uint2(0u, 0u);
        if (h < m) {
          pr[u] = pairs[base + h];
          atomicAdd(&s_cnt[(int)((int64_t)pr[u].x - pos0)], 1);
        }
      }
      for (int h = tid + 4 * kLinkThreads; h < m; h += kLinkThreads) {
        const uint2 q = pairs[base + h];
        atomicAdd(&s_cnt[(int)((int64_t)q.x - pos0)], 1);
      }
      __syncthreads();
      block_scan_1024(s_cnt, s_part);  // s_cnt[p]: the list's cursor; after the binning pass: the list's END
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int h = tid + u * kLinkThreads;
        if (h < m) s_list[atomicAdd(&s_cnt[(int)((int64_t)pr[u].x - pos0)], 1)] = (int32_t)pr[u].y;
      }
      for (int h = tid + 4 * kLinkThreads; h < m; h += kLinkThreads) {
        const uint2 q = pairs[base + h];
        s_list[atomicAdd(&s_cnt[(int)((int64_t)q.x - pos0)], 1)] = (int32_t)q.y;
      }
      __syncthreads();
      for (int p = tid; p < P; p += kLinkThreads) {
        const int32_t lo = p ? s_cnt[p - 1] : 0, hi = s_cnt[p];
        for (int32_t a = lo + 1; a < hi; ++a) {  // insertion sort: ~1 entry on average, ~ln(n/p) at the low end
          const int32_t key = s_list[a];
          int32_t b = a - 1;
          while (b >= lo && s_list[b] > key) {
            s_list[b + 1] = s_list[b];
            --b;
          }
          s_list[b + 1] = key;
        }
        last[pos0 + p] = hi > lo ? s_list[hi - 1] : -1;
        int32_t prev = ~(int32_t)(pos0 + p);  // the first hit of p goes on at p's own step
        for (int32_t a = lo; a < hi; ++a) {
          const int32_t i = s_list[a];
          if (RSCATTER) rpairs[atomicAdd(rcur + (i >> kRangeLog), 1)] = make_uint2((uint32_t)i, (uint32_t)prev);
          else R[i] = prev;
          prev = i;
        }
      }
      __syncthreads();
      k0 = k1;
    }
    __syncthreads();  // s_gs is rewritten by the next job
  }
}

// K2b place: one block per coarse range of 4096 steps: the range's entries (at most one per step, in
// arrival order) are put in place in shared memory and written out in order.  Steps without an entry are
// step 0 and the self swaps (j_i = i): R[i] = i.
__global__ void __launch_bounds__(256) shuf2_place_kernel(int64_t n, int64_t nranges, const int32_t* __restrict__ rcur,
                                                          const uint2* __restrict__ rpairs, int32_t* __restrict__ R) {
  __shared__ int32_t s_r[1 << kRangeLog];
  constexpr int32_t kUnset = INT32_MIN;  // not a value of R: R >= 0 or R = ~j with j <= 2^31 - 2
  const int tid = threadIdx.x;
  for (int64_t rg = blockIdx.x; rg < nranges; rg += gridDim.x) {
    const int64_t i0 = rg << kRangeLog;
    const int32_t m = rcur[rg] - (int32_t)i0;
    for (int k = tid; k < (1 << kRangeLog); k += 256) s_r[k] = kUnset;
    __syncthreads();
    for (int h = tid; h < m; h += 256) {
      const uint2 e = rpairs[i0 + h];
      s_r[e.x & ((1u << kRangeLog) - 1)] = (int32_t)e.y;
    }
    __syncthreads();
    for (int k = tid; k < (1 << kRangeLog); k += 256)
      if (i0 + k < n) R[i0 + k] = s_r[k] == kUnset ? (int32_t)(i0 + k) : s_r[k];
    __syncthreads();
  }
}

// K3 resolve: src(p) as above, four positions per thread in flight; out[p] = in[src(p)], or the position
// map itself (in == nullptr).
__global__ void __launch_bounds__(256) shuf2_resolve_kernel(int64_t n, const int32_t* __restrict__ last,
                                                            const int32_t* __restrict__ R,
                                                            const int32_t* __restrict__ in, int32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
    int32_t s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      s[u] = i < n ? last[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n && s[u] < 0) s[u] = R[i];
    }
    bool any = true;
    while (any) {  // the four chains advance together
      int32_t nx[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) nx[u] = s[u] < 0 ? ld_sparse(R + ~s[u]) : s[u];
      any = false;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s[u] = nx[u];
        any = any || nx[u] < 0;
      }
    }
    int32_t val[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      val[u] = (in != nullptr && i < n) ? ld_sparse(in + s[u]) : s[u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n) out[i] = val[u];
    }
  }
}

}  // namespace b200sgd
