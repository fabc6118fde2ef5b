// sgd_cluster_kernel.cuh -- the epoch kernel for small and mid-size steps of narrow rows (BASELINE
// config 2: 4M x 100, B = 1000: a step moves 0.45 MB = 70 ns of HBM time; what it costs is the cross-SM
// dependency "step k+1 needs the weights of step k, which need every row of step k").
//
// Same algorithm, record layout and summation discipline as sgd_epoch_kernel (sgd_kernels.cuh;
// reference: cuda_iteration / cublas_iteration, main.cu:56-163 / :168-301), but the CTAs that share a
// step form ONE thread-block cluster (16 CTAs, or 8) and the per-step gradient exchange never leaves
// the SMs:
//
//   * Through L2 (the generic kernel) a hop is "store, then poll": ~800 cycles SM to SM, quantised by
//     the poll round trip and stretched by the slowest of 32 publishers -- measured 2400 + 2000
//     cycles for the two hops of a 6500-cycle step.  Here every warp sends its 16 column sums from
//     registers straight into the shared memory of every CTA of the cluster (st.async, four columns
//     per lane, one instruction per eight destination CTAs), each store completing its bytes on the
//     RECEIVER's mbarrier: the receiver sleeps on a local mbarrier instead of polling L2, and there
//     is no second hop -- every CTA adds the same CS partials in the same order and updates its own
//     copy of the weights (bit-identical in all CTAs, no atomics, reruns reproduce).
//   * Column mapping of the gradient: phase A (8 lanes per row, 4 rows per warp, 8 warps = one 32-row
//     ring tile) leaves the residuals in shared memory; after one CTA barrier warp w adds
//     res[r] * x[r][c] for its columns 16w .. 16w+15 straight from the staged records, one half warp
//     over the even rows, the other over the odd rows; one shuffle folds the halves.  A warp therefore
//     ends the step with 16 finished column sums -- exactly what it sends.
//   * The rows of the next step are loaded into registers while the exchange is in flight.
//   * Producer warps: TMA tile::gather4 (four rows per instruction) into the ring, running ahead of
//     the consumers across step boundaries, as in the generic kernel.
//   * Mid-size batches and several GPUs (MULTI): several clusters ("groups") take part in a step and
//     meet in an inbox of tagged words in global memory (L2 on one GPU, NVLink between GPUs).
//
// Restrictions (the host falls back to sgd_epoch_kernel otherwise): C + 1 <= 128, tile::gather4
// available, persistent engine, dataflow reduction, no residual output, 48 .. 10240 rows per rank and step.
#pragma once
#include "sgd_kernels.cuh"

namespace b200sgd {

constexpr int kClNP = 4;              // producer warps (a round's indices come from HBM: ~1600 cycles per round and warp)
constexpr int kClNW = 8;              // consumer warps
constexpr int kClCT = 32 * kClNW;     // consumer threads
constexpr int kClTileRows = 32;       // rows per ring tile = one producer round (8 x gather4)
constexpr int kClMaxCS = 16;          // CTAs per cluster (16 needs the non-portable opt-in)
constexpr int kClMaxTiles = 16;
constexpr int kClMaxChunk = 4;        // ring tiles whose phase A runs as one block
// shared memory map (bytes)
constexpr uint32_t kClOffFull = 0;      // uint64 full[16]
constexpr uint32_t kClOffEmpty = 128;   // uint64 empty[16]
constexpr uint32_t kClOffRecvBar = 256; // uint64 recv[2]
constexpr uint32_t kClOffW = 512;       // float w_s[128]
constexpr uint32_t kClOffR = 1024;      // float r_s[2][kClMaxChunk][32]
constexpr int kClSlot = 144;            // floats per source slot: 128 columns + 16, so that consecutive sources sit 16 banks apart
constexpr uint32_t kClOffRecv = 2048;   // float recv[2][kClMaxCS][kClSlot]
constexpr uint32_t kClOffRing = kClOffRecv + 2 * kClMaxCS * kClSlot * 4;  // 20480, 128-byte aligned
constexpr uint32_t kClRingSlack = 512;  // the last row's unpredicated 128-float read may run past the ring
constexpr size_t kClFixedSmem = kClOffRing + kClRingSlack;

// tagged 8-byte words {value, tag} through global memory (cross-GPU hop): see ll_store_f4
__device__ __forceinline__ void ll_store_f1(uint2* dst, float v, uint32_t tag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(dst), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_volatile_u2(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
// two fp32 FMAs in one instruction (SASS FFMA2; each half rounds exactly like fmaf)
__device__ __forceinline__ float2 ffma2(float2 a, float s, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%4};\n\tmov.b64 rc, {%5,%6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0,%1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(s), "f"(c.x), "f"(c.y));
  return d;
}
// EpochParams fields used: rec_map, ld, C, perm, B, nb, first_step, a, w, g_last, ntiles, tile_log2,
// slot_bytes, loss, err.  gridDim.x == cluster size (one cluster).
//   CHUNK = ring tiles per phase-A block (1, 2 or 4: the host takes the tiles of a step, capped at 4)
//   PROF  = (B200SGD_PHASE_PROFILE=1) lane 0 of consumer warp 0 stores the SM clock at the marks of the
//           middle step into p.dbg[cta][0..7], the clock at the end of the step before into p.dbg[grid+cta][0]
//   MULTI = more than one cluster takes part in a step: several clusters on this GPU (gridDim.x = NG * CS,
//           mid-size batches) and / or several GPUs (nranks > 1, one cluster each, rows of a step from the
//           bucketed row list step_off).  "Group" q = rank * NG + cluster.  After the cluster exchange the
//           first CTA of a cluster stores the cluster's column sums as tagged 8-byte words into slot
//           [parity][q] of the inbox of every rank (own rank: L2; peers: NVLink), every CTA polls its own
//           rank's inbox for the other groups and adds the groups' sums in group order (identical bits
//           in every CTA of every GPU).  Inbox: ll_inbox as uint2 [2][nranks * NG][4*C4p].
template <int CHUNK, bool PROF = false, bool MULTI = false>
__global__ void __launch_bounds__(32 * (kClNW + kClNP), 1) sgd_cluster_epoch_kernel(const __grid_constant__ EpochParams p) {
  static_assert(CHUNK >= 1 && CHUNK <= kClMaxChunk, "r_s holds kClMaxChunk tiles of residuals");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* w_s = reinterpret_cast<float*>(smem_raw + kClOffW);
  float* r_s = reinterpret_cast<float*>(smem_raw + kClOffR);
  const float* recv_s = reinterpret_cast<const float*>(smem_raw + kClOffRecv);
  unsigned char* ring = smem_raw + kClOffRing;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = (int)cluster_ctarank(), CS = (int)cluster_nctarank();
  const int gcta = (int)blockIdx.x, G = (int)gridDim.x;  // CTA within the rank; the rank's rows are cut over all G CTAs
  const int NG = G / CS, grp = gcta / CS;                // clusters of this rank, mine
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t full0 = smem0 + kClOffFull, empty0 = smem0 + kClOffEmpty, recvbar0 = smem0 + kClOffRecvBar;
  const uint32_t ring0 = smem0 + kClOffRing;
  const uint32_t tile_mask = (uint32_t)p.ntiles - 1u;
  const uint32_t tile_bytes = (uint32_t)kClTileRows * p.slot_bytes;

  if (tid == 0) {
    for (int i = 0; i < p.ntiles; ++i) {
      mbar_init(full0 + 8u * i, 1);
      mbar_init(empty0 + 8u * i, kClNW);
    }
    mbar_init(recvbar0, 1);
    mbar_init(recvbar0 + 8u, 1);
    fence_barrier_init();
    // arm the receive barriers of steps 0 and 1 before anybody can send (later ones: one step ahead)
    const uint32_t expect0 = cluster_nctarank() * (((uint32_t)p.C * 4u + 15u) & ~15u);
    mbar_arrive_expect_tx(recvbar0, expect0);
    mbar_arrive_expect_tx(recvbar0 + 8u, expect0);
  }
  for (int c = tid; c < 128; c += blockDim.x) w_s[c] = (c < p.C) ? p.w[c] : 0.0f;
  // phase B reads whole tiles (rows beyond a short slice are multiplied by a zero residual): the
  // ring must never hold bit patterns of NaN or Inf
  for (uint32_t i = tid; i < ((uint32_t)p.ntiles * tile_bytes + kClRingSlack) / 16u; i += blockDim.x)
    reinterpret_cast<float4*>(ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async_smem();  // the TMA writes that follow go through the async proxy
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers exist before anybody sends

  // my rows of step s: [off, off + n) of the row list (single GPU: the same slice of every batch)
  auto step_rows = [&](int s, int64_t* off, int* n) {
    const int64_t sa = (int64_t)p.first_step + s;
    int64_t rows = p.B, o = sa * p.B;
    if (MULTI && p.step_off != nullptr) {
      o = p.step_off[sa];
      rows = p.step_off[sa + 1] - o;
    }
    const int64_t lo = slice_lo(rows, gcta, G);
    *off = o + lo;
    *n = (int)(slice_lo(rows, gcta + 1, G) - lo);
  };

  if (warp < kClNP) {
    // ===================================== PRODUCERS =====================================
    // round = up to 32 rows of one step = one ring tile; producer warp w issues rounds w, w + NP, ...:
    // lane 0 does the flow control, then one tile::gather4 per four rows.  The indices of a warp's next
    // round are in flight while it issues the current one (they come from HBM, ~1600 cycles: one warp
    // alone sustains only one round per that time).
    int s = 0, j0 = 0, n = 0;
    int64_t off = 0;
    uint32_t tq = 0;
    auto settle = [&]() {  // to the next step in which this CTA has rows
      while (s < p.nb) {
        step_rows(s, &off, &n);
        if (n > 0) break;
        ++s;
      }
    };
    auto advance = [&]() {  // to the next round of the CTA's sequence
      if (s >= p.nb) return;
      j0 += min(kClTileRows, n - j0);
      ++tq;
      if (j0 >= n) {
        j0 = 0;
        ++s;
        settle();
      }
    };
    auto round_cnt = [&]() -> int { return (s < p.nb) ? min(kClTileRows, n - j0) : 0; };
    settle();
    for (int k = 0; k < warp; ++k) advance();
    int cnt = round_cnt();
    int32_t idx = 0;
    if (lane < cnt) idx = __ldg(p.perm + off + j0 + lane);
    while (cnt > 0) {
      const int cur_cnt = cnt;
      const int32_t cur_idx = idx;
      const uint32_t cur_tq = tq;
      for (int k = 0; k < kClNP; ++k) advance();
      cnt = round_cnt();
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
  } else {
    // ======================================= CONSUMERS =======================================
    const int ctid = tid - 32 * kClNP, cw = warp - kClNP;
    const int rg = lane >> 3, sl = lane & 7;  // phase A: row group of the warp, sub-lane within the row
    const int hp = lane >> 4;                 // sum phase: which half of the sources this half warp adds
    const int cls = lane >> 3;                // phase B: my rows of a tile are cls, cls + 4, cls + 8, ...
    const int colB = 16 * cw + 2 * (lane & 7);  // ... and my two gradient columns (16 columns per warp)
    const int C = p.C, ld = (int)p.ld;
    const float Bf = (float)p.B, a = p.a;
    const uint32_t expect = (uint32_t)CS * (((uint32_t)C * 4u + 15u) & ~15u);  // bytes every CTA receives per step
    const bool armer = ctid == kClCT - 1;  // arms the receive barriers (kept off warp 0's path)
    const int rowA = cw * 4 + rg;             // my row of every tile in phase A
    const int rslot = (rowA & 3) * 8 + (rowA >> 2);  // residuals are stored by row class (row mod 4)
    int n;                                    // my rows of the current step
    {
      int64_t off_unused;
      if (p.nb > 0) step_rows(0, &off_unused, &n); else n = 0;
    }
    const int cpad = 4 * ((p.C4 + 3) & ~3);   // MULTI: columns per group in the inbox

    // running training loss: fp32 per lane (a lane sees <= 2 * nb * tiles residuals per launch), double from there on
    float loss_abs = 0.f, loss_sq = 0.f;
    uint32_t tq = 0, chunk_seq = 0;
    int cur_s = 0;
    auto mark = [&](int i) {
      if (PROF && ctid == 0 && p.dbg != nullptr) {
        const long long t = clock64();
        if (cur_s == p.nb / 2) p.dbg[(int64_t)gcta * 8 + i] = (unsigned long long)t;
        if (cur_s == p.nb / 2 - 1 && i == 7) p.dbg[((int64_t)G + gcta) * 8] = (unsigned long long)t;
      }
    };

    // rows of a phase-A block: wait for the tiles, fetch my row of each into registers
    const unsigned char* tb[CHUNK];
    float4 x[CHUNK][4];
    float y[CHUNK];
    bool live[CHUNK];
    auto load_block = [&](int t0, int nch, int n) {
      // all probes of the block first, back to back (a probe takes ~100 cycles even when the tile has
      // landed long ago), the retry loop only for tiles that are late
      bool landed[CHUNK];
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        const uint32_t q = tq + (uint32_t)u;
        const uint32_t tile = q & tile_mask;
        tb[u] = ring + (size_t)tile * tile_bytes;
        landed[u] = u >= nch || mbar_try_wait(full0 + 8u * tile, (q >> p.tile_log2) & 1u);
      }
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        const uint32_t q = tq + (uint32_t)u;
        if (!landed[u]) mbar_wait_guarded(full0 + 8u * (q & tile_mask), (q >> p.tile_log2) & 1u, p.err);
      }
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        // no predicates on the loads: the ring (plus the slack after it) only ever holds finite values,
        // columns beyond C meet zero weights, and rows beyond the slice get their residual forced to 0
        live[u] = u < nch && rowA < n - (t0 + u) * kClTileRows;
        const float4* row = reinterpret_cast<const float4*>(tb[u] + (size_t)rowA * p.slot_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) x[u][k] = row[sl + 8 * k];
        y[u] = reinterpret_cast<const float*>(row)[C];  // the label rides in the record
      }
    };
    if (p.nb > 0) load_block(0, min(CHUNK, (n + kClTileRows - 1) / kClTileRows), n);

    for (int s = 0; s < p.nb; ++s) {
      if (PROF) cur_s = s;
      const int par = s & 1;
      const uint32_t rbar = recvbar0 + 8u * par;
      const int nt = (n + kClTileRows - 1) / kClTileRows;
      int n_next = n;  // rows of the next step (MULTI: from the row list, loads in flight during the step)
      if (MULTI && p.step_off != nullptr && s + 1 < p.nb) {
        int64_t off_unused;
        step_rows(s + 1, &off_unused, &n_next);
      }
      // arm the barrier of step s+1: its previous phase (step s-1) is complete and every local waiter
      // has left it (end-of-step barrier); nobody can send step s+1 before it has my step s
      if (armer && s >= 1 && s + 1 < p.nb) mbar_arrive_expect_tx(recvbar0 + 8u * (par ^ 1), expect);
      float4 wv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) wv[k] = reinterpret_cast<const float4*>(w_s)[sl + 8 * k];
      float2 acc[4];  // columns colB, colB + 1 over the rows of my class
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = make_float2(0.f, 0.f);
      // the tiles of the step, CHUNK at a time: phase A of all of them as one straight-line block
      // (their latencies overlap), one CTA barrier, phase B of all of them.  The rows of the first
      // block are already in registers (fetched while the previous step's exchange was in flight).
      for (int t0 = 0; t0 < nt; t0 += CHUNK, ++chunk_seq) {
        const int nch = min(CHUNK, nt - t0);
        float* rr = r_s + (chunk_seq & 1u) * (kClMaxChunk * 32);
        if (t0 > 0) load_block(t0, nch, n);
        mark(0);  // rows in registers
        // ---- phase A: residual of row rowA of every tile  (r = x.w - y, sgd_thrust.cu:43-48 / :352-364)
        float dot[CHUNK];
#pragma unroll
        for (int u = 0; u < CHUNK; ++u) {
          float d[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // independent chains; w_s is zero for the label and the padding
            d[k] = x[u][k].x * wv[k].x;
            d[k] = fmaf(x[u][k].y, wv[k].y, d[k]);
            d[k] = fmaf(x[u][k].z, wv[k].z, d[k]);
            d[k] = fmaf(x[u][k].w, wv[k].w, d[k]);
          }
          dot[u] = (d[0] + d[1]) + (d[2] + d[3]);
        }
#pragma unroll
        for (int m = 1; m <= 4; m <<= 1) {
#pragma unroll
          for (int u = 0; u < CHUNK; ++u) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], m);
        }
#pragma unroll
        for (int u = 0; u < CHUNK; ++u) {
          const float res = live[u] ? dot[u] - y[u] : 0.f;  // rows beyond the step's slice: 0
          if (sl == 0) rr[u * 32 + rslot] = res;
          loss_abs += fabsf(res);  // every lane of the row group carries the same residual
          loss_sq = fmaf(res, res, loss_sq);
        }
        mark(1);  // phase A done
        consumer_bar<kClCT>();
        mark(2);  // residuals visible
        // ---- phase B: res[r] * x[r][colB, colB+1] over the 8 rows of my class of every tile, two columns
        // per lane and instruction (FFMA2).  No row bound: rows beyond the slice have res = 0 and finite
        // (stale or zero-initialised) records.  (Adjacent rows are ld = 16 mod 32 floats apart for C = 100:
        // the two row classes of a half warp hit disjoint banks.)
        if (colB < ld) {
#pragma unroll
          for (int u = 0; u < CHUNK; ++u) {
            if (u < nch) {
              const float* col = reinterpret_cast<const float*>(tb[u]) + (size_t)cls * ld + colB;
              const float4* r4 = reinterpret_cast<const float4*>(rr + u * 32 + 8 * cls);
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const float4 rv = r4[j];  // rows 4*(4j..4j+3) + cls
                acc[0] = ffma2(*reinterpret_cast<const float2*>(col + (size_t)(16 * j + 0) * ld), rv.x, acc[0]);
                acc[1] = ffma2(*reinterpret_cast<const float2*>(col + (size_t)(16 * j + 4) * ld), rv.y, acc[1]);
                acc[2] = ffma2(*reinterpret_cast<const float2*>(col + (size_t)(16 * j + 8) * ld), rv.z, acc[2]);
                acc[3] = ffma2(*reinterpret_cast<const float2*>(col + (size_t)(16 * j + 12) * ld), rv.w, acc[3]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int u = 0; u < CHUNK; ++u)
            if (u < nch) mbar_arrive(empty0 + 8u * ((tq + (uint32_t)u) & tile_mask));
        }
        tq += (uint32_t)nch;
      }

      // ---- exchange: fold the four row classes (two shuffle levels on two columns); the warp's 16 column
      // sums then go to slot [par][cta][16*cw ..] of EVERY CTA straight from registers: lane l carries
      // the four columns 4*(l & 3) .. +3 to CTA (l >> 2) (+ 8 in the second pass), so ONE st.async.v4
      // serves 8 destination CTAs, each store completing its 16 bytes on its destination's barrier.
      // No staging, no proxy fence, no CTA barrier.  (Issuing costs a warp ~60 cycles per remote-store
      // instruction or bulk copy whatever its size: 16 of them per warp were 1000 cycles per step.)
      {
        float2 v = make_float2((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y));
        v.x += __shfl_xor_sync(0xffffffffu, v.x, 8);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, 8);
        v.x += __shfl_xor_sync(0xffffffffu, v.x, 16);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, 16);
        const int q4 = 4 * (lane & 3);  // my four columns live in lanes q4/2 and q4/2 + 1
        float4 f;
        f.x = __shfl_sync(0xffffffffu, v.x, (q4 >> 1) + 0);
        f.y = __shfl_sync(0xffffffffu, v.y, (q4 >> 1) + 0);
        f.z = __shfl_sync(0xffffffffu, v.x, (q4 >> 1) + 1);
        f.w = __shfl_sync(0xffffffffu, v.y, (q4 >> 1) + 1);
        if (16 * cw + q4 < C) {  // whole 16-byte units up to C (the sum of these is `expect`)
          const uint32_t slot = smem0 + kClOffRecv + (uint32_t)(((par * kClMaxCS + cta) * kClSlot + 16 * cw + q4) * 4);
#pragma unroll
          for (int pass = 0; pass < kClMaxCS / 8; ++pass) {
            const uint32_t dst = (uint32_t)(8 * pass + (lane >> 2));
            if ((int)dst < CS) st_async_f4(map_to_cta(slot, dst), f, map_to_cta(rbar, dst));
          }
        }
        mark(3);  // sent
      }
      // the rows of the next step do not depend on the weights: fetch them while the exchange is in flight
      if (s + 1 < p.nb) load_block(0, min(CHUNK, (n_next + kClTileRows - 1) / kClTileRows), n_next);
      mark(4);
      // ---- every CTA adds the CS partials of a column in the same order and applies
      //      w = a * (g / B) + w   (sgd_thrust.cu:269, main.cu:147-150 / :287-290)
      // (warp cw owns columns 16*cw ..: its half warps add one half of the sources each, one shuffle folds)
      {
        mbar_wait_guarded(rbar, (uint32_t)(s >> 1) & 1u, p.err);
        mark(5);  // all partials arrived
        const int c = 16 * cw + (lane & 15);
        // half warp hp adds the sources hp, hp + 2, ... (16 banks apart from the other half's)
        const float* rb = recv_s + ((size_t)par * kClMaxCS + (size_t)hp) * kClSlot + c;
        float t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i += 2) {
          t0 += rb[(2 * i + 0) * kClSlot];
          t1 += rb[(2 * i + 2) * kClSlot];
        }
        if (CS > 8) {
#pragma unroll
          for (int i = 4; i < 8; i += 2) {
            t0 += rb[(2 * i + 0) * kClSlot];
            t1 += rb[(2 * i + 2) * kClSlot];
          }
        }
        float tot = t0 + t1;
        tot += __shfl_xor_sync(0xffffffffu, tot, 16);  // this GPU's rows
        if (lane < 16 && c < C) {
          if (MULTI) {
            // second hop through global memory (L2 on this GPU, NVLink to the peers): the cluster's first
            // CTA publishes the cluster's sum, everybody polls the own rank's inbox for the other
            // groups and adds the groups' sums in group order
            const uint32_t tag = p.tag_base + (uint32_t)s + 1u;
            const int TG = p.nranks * NG, me = p.rank * NG + grp;
            const int64_t at = ((int64_t)par * TG + me) * cpad + c;
            if (cta == 0)
              for (int r = 0; r < p.nranks; ++r)
                if (NG > 1 || r != p.rank) ll_store_f1(reinterpret_cast<uint2*>(p.ll_inbox_peer[r]) + at, tot, tag);
            const uint2* inbox = reinterpret_cast<const uint2*>(p.ll_inbox_local) + (int64_t)par * TG * cpad + c;
            float all = 0.f;
            for (int q0 = 0; q0 < TG; q0 += 8) {
              // poll all missing groups of the batch together until the last one has arrived: after the
              // last arrival one L2 round trip is left, not one per group
              float vq[8];
              uint32_t pending = 0;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                vq[k] = tot;
                if (q0 + k < TG && q0 + k != me) pending |= 1u << k;
              }
              SpinGuard guard;
              while (pending != 0u) {
                uint2 u[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                  if (pending & (1u << k)) u[k] = ld_volatile_u2(inbox + (int64_t)(q0 + k) * cpad);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                  if ((pending & (1u << k)) && u[k].y == tag) {
                    vq[k] = __uint_as_float(u[k].x);
                    pending &= ~(1u << k);
                  }
                if (pending != 0u && guard.expired(p.err, 3u)) break;
              }
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (q0 + k < TG) all += vq[k];
            }
            tot = all;
          }
          const float gs = tot / Bf;
          w_s[c] = fmaf(a, gs, w_s[c]);
          if (gcta == 0 && s == p.nb - 1) p.g_last[c] = gs;
        }
      }
      mark(6);  // weights updated
      consumer_bar<kClCT>();  // w_s complete
      mark(7);
      n = n_next;
    }

    // running loss of the launch: per-CTA sums of |r| and r^2 (lanes 0, 8, 16, 24 hold the four row groups)
    double la = (double)loss_abs, lq = (double)loss_sq;
    la += __shfl_xor_sync(0xffffffffu, la, 8);
    la += __shfl_xor_sync(0xffffffffu, la, 16);
    lq += __shfl_xor_sync(0xffffffffu, lq, 8);
    lq += __shfl_xor_sync(0xffffffffu, lq, 16);
    if (p.loss != nullptr) {
      double* ls = reinterpret_cast<double*>(r_s);  // free after the last end-of-step barrier
      if (lane == 0) {
        ls[2 * cw] = la;
        ls[2 * cw + 1] = lq;
      }
      consumer_bar<kClCT>();
      if (ctid == 0) {
        double sa = 0.0, sb = 0.0;
        for (int wq = 0; wq < kClNW; ++wq) {
          sa += ls[2 * wq];
          sb += ls[2 * wq + 1];
        }
        p.loss[2 * gcta] += sa;
        p.loss[2 * gcta + 1] += sb;
      }
    }
    if (gcta == 0)  // every CTA holds the same bits
      for (int c = ctid; c < C; c += kClCT) p.w[c] = w_s[c];
  }
  cluster_sync_all();  // nobody leaves while a peer may still write into its shared memory
}

}  // namespace b200sgd
