// sgd_small_kernel.cuh -- the epoch kernel for SMALL steps of narrow rows (BASELINE config 2:
// 4M x 100, B = 1000 -- a step moves 0.45 MB, 70 ns of HBM time, and is pure cross-SM latency).
//
// Same algorithm, same record layout, same tagged-word dataflow reduction as sgd_epoch_kernel
// (sgd_kernels.cuh; reference: cuda_iteration / cublas_iteration, main.cu:56-163 / :168-301), but
// built for the shortest dependent chain instead of for generality, because at this size every
// warp executes its few hundred instructions of a step exactly once and nothing hides their latency
// or their instruction fetches (the generic kernel is ~9500 SASS instructions, this one ~1000):
//
//   * column mapping of the gradient: phase A -- 8 lanes per row, 4 rows per warp, 8 warps = one
//     32-row ring tile -- leaves the residuals in shared memory; after one CTA barrier thread c
//     accumulates  g[c] = sum_r res[r] * x[r][c]  over the rows of the tile straight from the staged
//     records.  The CTA partial is then already one float per thread: no cross-lane folding of the
//     row groups, no cross-warp scratch, no second pass.
//   * one float per tagged 8-byte word {value, tag}: thread c publishes its column with ONE store;
//     an owner warp gathers a column with one load per lane and a 5-level butterfly on a scalar.
//   * split duties: warps 0-3 (the column threads) publish and then poll the CTA's mailbox for the
//     new weights; warps 4-7 own the CTA's columns: they start polling for partials right after
//     phase A and never touch the gradient.
//   * multi-GPU (MULTI): every CTA stores its column partials into the owners' gather arrays of
//     EVERY rank over NVLink (peer-mapped memory); the owners of all ranks add the nranks * G
//     partials in the same fixed order, so the replicas stay bit-identical.
//
// Restrictions (the host falls back to sgd_epoch_kernel otherwise): C + 1 <= 128, tile::gather4
// available, persistent engine, dataflow reduction, no residual output, grid > 1.
#pragma once
#include "sgd_kernels.cuh"

namespace b200sgd {

constexpr int kSmallNW = 8;              // consumer warps
constexpr int kSmallCT = 32 * kSmallNW;  // consumer threads
constexpr int kSmallTileRows = 32;       // rows per ring tile = one producer round (8 x gather4)
constexpr int kSmallColWarps = 4;        // warps 0..3: one gradient column per thread (<= 128 columns)
constexpr size_t kSmallFixedSmem = 2048; // barriers 1024, weights 512, residuals 256, pad

__device__ __forceinline__ void ll_store_f1(uint2* dst, float v, uint32_t tag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(dst), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_volatile_u2(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
// blocks until the word carries `tag` (watchdog as in ll_wait_f4)
__device__ __forceinline__ float ll_wait_f1(const uint2* src, uint32_t tag, unsigned int* err) {
  uint2 v = ld_volatile_u2(src);
  if (v.y != tag) {
    SpinGuard guard;
    do {
      if (guard.expired(err, 3u)) break;
      v = ld_volatile_u2(src);
    } while (v.y != tag);
  }
  return __uint_as_float(v.x);
}
// lane q of a Q-lane group adds sources q, q+Q, ... (ascending, up to four loads in flight)
__device__ __forceinline__ float ll_gather_f1(const uint2* base, int nsrc, int q, int Q, uint32_t tag,
                                              unsigned int* err) {
  float acc = 0.f;
  int i = q;
  for (; i + 3 * Q < nsrc; i += 4 * Q) {
    const uint2 u0 = ld_volatile_u2(base + i), u1 = ld_volatile_u2(base + i + Q);
    const uint2 u2 = ld_volatile_u2(base + i + 2 * Q), u3 = ld_volatile_u2(base + i + 3 * Q);
    acc += (u0.y == tag) ? __uint_as_float(u0.x) : ll_wait_f1(base + i, tag, err);
    acc += (u1.y == tag) ? __uint_as_float(u1.x) : ll_wait_f1(base + i + Q, tag, err);
    acc += (u2.y == tag) ? __uint_as_float(u2.x) : ll_wait_f1(base + i + 2 * Q, tag, err);
    acc += (u3.y == tag) ? __uint_as_float(u3.x) : ll_wait_f1(base + i + 3 * Q, tag, err);
  }
  for (; i < nsrc; i += Q) acc += ll_wait_f1(base + i, tag, err);
  return acc;
}

// EpochParams fields used: rec_map, ld, C, C4, perm, step_off (MULTI), B, nb, first_step, a, w, g_last,
// err, ntiles, tile_log2, slot_bytes, ll_part, ll_w, tag_base, rank, nranks, ll_xpart_local/peer, loss.
// Gather arrays as uint2 [2 parities][4*C4 columns][S sources]; S = G padded to 4 (single GPU, the
// bytes of ll_part) or nranks*G (MULTI, the bytes of ll_xpart).  Mailboxes: ll_w as uint2
// [G][4*C4p], one private copy of the weights per CTA.
// PROF (B200SGD_PHASE_PROFILE=1 only): lane 0 of warp 0 (a column warp) and of warp 4 (an owner warp) sum the
// cycles between marks into p.dbg: [cta][8] for the column warp, [grid + cta][8] for the owner warp.
template <bool MULTI, bool PROF = false>
__global__ void __launch_bounds__(32 * (kSmallNW + 1), 1) sgd_small_epoch_kernel(const __grid_constant__ EpochParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty_bar = full_bar + kMaxTiles;
  float* w_s = reinterpret_cast<float*>(smem_raw + 1024);  // [128], zero beyond C
  float* r_s = reinterpret_cast<float*>(smem_raw + 1536);  // [2][32] residuals of a tile
  unsigned char* ring = smem_raw + kSmallFixedSmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const uint32_t tile_mask = (uint32_t)p.ntiles - 1u;
  const uint32_t tile_bytes = (uint32_t)kSmallTileRows * p.slot_bytes;
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), ring0 = smem_u32(ring);

  if (tid == 0) {
    for (int i = 0; i < p.ntiles; ++i) {
      mbar_init(full0 + 8u * i, 1);
      mbar_init(empty0 + 8u * i, kSmallNW);
    }
    fence_barrier_init();
  }
  for (int c = tid; c < 128; c += blockDim.x) w_s[c] = (c < p.C) ? p.w[c] : 0.0f;
  __syncthreads();

  // rows of step s that this CTA handles: [off + lo, off + lo + n) of the (local) row list
  auto step_slice = [&](int s, int64_t* off, int* n) {
    const int64_t sa = (int64_t)p.first_step + s;
    int64_t rows = p.B, o = sa * p.B;
    if (MULTI) {
      o = p.step_off[sa];
      rows = p.step_off[sa + 1] - o;
    }
    const int64_t lo = slice_lo(rows, cta, G);
    *off = o + lo;
    *n = (int)(slice_lo(rows, cta + 1, G) - lo);
  };

  if (warp == 0) {
    // ===================================== PRODUCER =====================================
    // round = up to 32 rows of one step = one ring tile: the lane-0 flow control, then one
    // tile::gather4 per four rows.  The indices of the next round are in flight meanwhile; the
    // producer runs ahead of the consumers by the whole ring, across step boundaries.
    int s = 0, j0 = 0, n = 0;
    int64_t off = 0;
    uint32_t tq = 0;
    auto settle = [&]() {  // to the next step in which this CTA has rows
      while (s < p.nb) {
        step_slice(s, &off, &n);
        if (n > 0) break;
        ++s;
      }
    };
    settle();
    int cnt = (s < p.nb) ? min(kSmallTileRows, n - j0) : 0;
    int32_t idx = 0;
    if (lane < cnt) idx = __ldg(p.perm + off + j0 + lane);
    while (cnt > 0) {
      const int cur_cnt = cnt;
      const int32_t cur_idx = idx;
      const uint32_t cur_tq = tq;
      j0 += cur_cnt;
      ++tq;
      if (j0 >= n) {
        j0 = 0;
        ++s;
        settle();
      }
      cnt = (s < p.nb) ? min(kSmallTileRows, n - j0) : 0;
      idx = 0;
      if (lane < cnt) idx = __ldg(p.perm + off + j0 + lane);

      const uint32_t tile = cur_tq & tile_mask;
      const uint32_t use = cur_tq >> p.tile_log2;
      const uint32_t fb = full0 + 8u * tile;
      const int groups = (cur_cnt + 3) >> 2;
      if (lane == 0) {
        mbar_wait_relaxed(empty0 + 8u * tile, (use & 1u) ^ 1u);
        mbar_arrive_expect_tx(fb, (uint32_t)groups * 4u * p.slot_bytes);
      }
      const int b4 = (lane << 2) & 31;
      const int32_t r0 = __shfl_sync(0xffffffffu, cur_idx, b4);
      int32_t r1 = __shfl_sync(0xffffffffu, cur_idx, (b4 + 1) & 31);
      int32_t r2 = __shfl_sync(0xffffffffu, cur_idx, (b4 + 2) & 31);
      int32_t r3 = __shfl_sync(0xffffffffu, cur_idx, (b4 + 3) & 31);
      if (4 * lane + 1 >= cur_cnt) r1 = r0;  // a short last group repeats its first row into slots nobody reads
      if (4 * lane + 2 >= cur_cnt) r2 = r0;
      if (4 * lane + 3 >= cur_cnt) r3 = r0;
      if (lane < groups)
        tma_gather4_g2s(ring0 + tile * tile_bytes + (uint32_t)(4 * lane) * p.slot_bytes, &p.rec_map, r0, r1, r2, r3, fb);
      __syncwarp();
    }
    return;
  }

  // ======================================= CONSUMERS =======================================
  const int ctid = tid - 32, cw = warp - 1;
  const int rg = lane >> 3, sl = lane & 7;  // phase A: row group of the warp, sub-lane within the row
  const int C = p.C, ld = (int)p.ld, ld4 = ld >> 2;
  const float Bf = (float)p.B, a = p.a;
  const int nsrc = MULTI ? p.nranks * G : G;         // partials per column
  const int S = MULTI ? nsrc : ((G + 3) & ~3);       // their stride in the gather array
  const int cpad = 4 * p.C4;                         // columns per parity
  const int mail_stride = 4 * ((p.C4 + 3) & ~3);
  uint2* const gat = reinterpret_cast<uint2*>(MULTI ? p.ll_xpart_local : p.ll_part);
  const uint2* const my_mail = reinterpret_cast<const uint2*>(p.ll_w) + (int64_t)cta * mail_stride;
  // columns this CTA owns, and how the owner warps (4..7) share them: Q lanes per column
  const int f0 = (int)slice_lo(C, cta, G);
  const int nf = (int)slice_lo(C, cta + 1, G) - f0;
  int Q = 1;
  while (Q < nsrc && Q < 32) Q <<= 1;
  const int cpw = 32 / Q;  // columns one warp gathers at a time
  const int oq = lane % Q, ocol = lane / Q;

  double loss_abs = 0.0, loss_sq = 0.0;  // running training loss (lane sl == 0 of every row group)
  unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long t_prev = 0;
  const bool prof = PROF && lane == 0 && (cw == 0 || cw == kSmallColWarps);
  int cur_s = 0;
  auto mark = [&](int i) {  // absolute SM clock at the marks of the middle step (slot 7: end of the step before)
    if (PROF && prof) {
      const long long t = clock64();
      if (cur_s == p.nb / 2) ph[i] = (unsigned long long)t;
      if (cur_s == p.nb / 2 - 1 && i == 6) ph[7] = (unsigned long long)t;
    }
  };
  if (PROF && prof) t_prev = clock64();
  uint32_t tq = 0;
  int64_t off;
  int n;
  step_slice(0, &off, &n);

  for (int s = 0; s < p.nb; ++s) {
    if (PROF) cur_s = s;
    const uint32_t tag = p.tag_base + (uint32_t)s + 1u;
    const int par = (p.first_step + s) & 1;
    int64_t off_next = 0;
    int n_next = 0;
    if (MULTI && s + 1 < p.nb) step_slice(s + 1, &off_next, &n_next);  // loads in flight during the step
    float4 wv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wv[k] = reinterpret_cast<const float4*>(w_s)[sl + 8 * k];
    float acc0 = 0.f, acc1 = 0.f;  // gradient column ctid (warps 0..3), even / odd rows
    const int nt = (n + kSmallTileRows - 1) / kSmallTileRows;
    for (int ti = 0; ti < nt; ++ti, ++tq) {
      const uint32_t tile = tq & tile_mask;
      const uint32_t use = tq >> p.tile_log2;
      const int cnt = min(kSmallTileRows, n - ti * kSmallTileRows);
      const unsigned char* tbase = ring + (size_t)tile * tile_bytes;
      float* rr = r_s + (tq & 1u) * 32;
      // ---- phase A: residual of row cw*4 + rg  (r = x.w - y, sgd_thrust.cu:43-48 / :352-364)
      const int r = cw * 4 + rg;
      const bool live = r < cnt;
      const float4* row = reinterpret_cast<const float4*>(tbase + (size_t)r * p.slot_bytes);
      mbar_wait(full0 + 8u * tile, use & 1u);
      mark(0);  // rows landed
      float4 x[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        x[k] = (live && sl + 8 * k < ld4) ? row[sl + 8 * k] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float y = live ? reinterpret_cast<const float*>(row)[C] : 0.f;  // the label rides in the record
      float d[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // four independent chains; w_s is zero for the label and the padding
        d[k] = x[k].x * wv[k].x;
        d[k] = fmaf(x[k].y, wv[k].y, d[k]);
        d[k] = fmaf(x[k].z, wv[k].z, d[k]);
        d[k] = fmaf(x[k].w, wv[k].w, d[k]);
      }
      float dot = (d[0] + d[1]) + (d[2] + d[3]);
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      const float res = dot - y;  // rows beyond the tile: 0
      if (sl == 0) {
        rr[r] = res;
        loss_abs += (double)fabsf(res);
        loss_sq += (double)res * (double)res;
      }
      if (cw >= kSmallColWarps) {  // my rows are in registers and I do not come back to this tile
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8u * tile);
      }
      mark(1);  // phase A done
      consumer_bar<kSmallCT>();
      mark(2);  // residuals visible
      // ---- phase B: thread c adds res[r] * x[r][c] over the rows of the tile, from the staged records
      if (cw < kSmallColWarps) {
        if (ctid < ld) {
          const float* col = reinterpret_cast<const float*>(tbase) + ctid;
          int q = 0;
#pragma unroll 4
          for (; q + 1 < cnt; q += 2) {
            acc0 = fmaf(rr[q], col[(size_t)q * ld], acc0);
            acc1 = fmaf(rr[q + 1], col[(size_t)(q + 1) * ld], acc1);
          }
          if (q < cnt) acc0 = fmaf(rr[q], col[(size_t)q * ld], acc0);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8u * tile);
      }
      mark(3);  // phase B done (column warps)
    }

    // ---- step tail: no barrier, no atomics; every float in an 8-byte word {value, tag}
    if (cw < kSmallColWarps) {
      if (ctid < C) {
        const float v = acc0 + acc1;
        if (MULTI) {
          const int64_t at = ((int64_t)par * cpad + ctid) * S + (int64_t)p.rank * G + cta;
          for (int pr = 0; pr < p.nranks; ++pr) ll_store_f1(reinterpret_cast<uint2*>(p.ll_xpart_peer[pr]) + at, v, tag);
        } else {
          ll_store_f1(gat + ((int64_t)par * cpad + ctid) * S + cta, v, tag);
        }
        mark(4);  // published
        w_s[ctid] = ll_wait_f1(my_mail + ctid, tag, p.err);  // the owners' answer
      }
      mark(5);  // new weights arrived
    } else {
      // owners: gather the nsrc partials of a column in a fixed order (lane q: sources q, q+Q, ..;
      // butterfly), apply  w = a * (g / B) + w  (sgd_thrust.cu:269, main.cu:147-150 / :287-290) and
      // store the new weight into the mailbox of every CTA
      for (int base = (cw - kSmallColWarps) * cpw; base < nf; base += (kSmallNW - kSmallColWarps) * cpw) {
        const int col = base + ocol;
        const bool active = col < nf;
        const int c = f0 + (active ? col : 0);
        const float w_old = w_s[c];  // read before this step's answer can overwrite it
        float tot = 0.f;
        if (active) tot = ll_gather_f1(gat + ((int64_t)par * cpad + c) * S, nsrc, oq, Q, tag, p.err);
        for (int m = 1; m < Q; m <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, m);
        mark(4);  // (owner) partials gathered
        if (active) {
          const float gs = tot / Bf;
          const float wn = fmaf(a, gs, w_old);
          uint2* dst = reinterpret_cast<uint2*>(p.ll_w) + c;
          for (int k = oq; k < G; k += Q) ll_store_f1(dst + (int64_t)k * mail_stride, wn, tag);
          if (oq == 0 && s == p.nb - 1) p.g_last[c] = gs;
        }
        mark(5);  // (owner) answer stored
      }
    }
    consumer_bar<kSmallCT>();  // w_s complete
    mark(6);
    if (MULTI) {
      off = off_next;
      n = n_next;
    }
  }

  if (PROF && prof && p.dbg != nullptr)
    for (int i = 0; i < 8; ++i) p.dbg[((cw == 0 ? 0 : (int64_t)G) + cta) * 8 + i] = ph[i];
  // running loss of the launch: per-CTA sums of |r| and r^2 (lane 0 of every warp after the fold)
  loss_abs += __shfl_xor_sync(0xffffffffu, loss_abs, 8);
  loss_abs += __shfl_xor_sync(0xffffffffu, loss_abs, 16);
  loss_sq += __shfl_xor_sync(0xffffffffu, loss_sq, 8);
  loss_sq += __shfl_xor_sync(0xffffffffu, loss_sq, 16);
  if (p.loss != nullptr) {
    double* ls = reinterpret_cast<double*>(r_s);
    if (lane == 0) {
      ls[2 * cw] = loss_abs;
      ls[2 * cw + 1] = loss_sq;
    }
    consumer_bar<kSmallCT>();
    if (ctid == 0) {
      double sa = 0.0, sb = 0.0;
      for (int wq = 0; wq < kSmallNW; ++wq) {
        sa += ls[2 * wq];
        sb += ls[2 * wq + 1];
      }
      p.loss[2 * cta] += sa;
      p.loss[2 * cta + 1] += sb;
    }
  }
  if (cta == 0)  // every CTA holds the same bits
    for (int c = ctid; c < C; c += kSmallCT) p.w[c] = w_s[c];
}

}  // namespace b200sgd
