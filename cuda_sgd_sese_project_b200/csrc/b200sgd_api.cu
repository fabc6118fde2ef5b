// b200sgd_api.cu -- the engine behind include/b200sgd.h: context, data residency in HBM, the
// per-epoch permutation pipeline, kernel dispatch, timing, multi-GPU peer plumbing.
//
// Reference correspondence (all under /root/reference/c_code):
//   sgd_create / sgd_load_rows   <-> main.cu:381-433  (device vectors, weight init, index vector)
//   sgd_shuffle                  <-> main.cu:90-91, :207-211 (+ removes sgd_thrust.cu:374-402)
//   sgd_epoch                    <-> main.cu:94-160, :227-300 (the batch loop)
//   sgd_run                      <-> main.cu:456-508
//   sgd_mean_abs_error           <-> sgd_thrust.cu:79-135
// No Thrust, no cuBLAS, no CPU fallback: every compute entry point needs an sm_100 device.
#include "b200sgd.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "host_io.hpp"
#include "host_rng.hpp"
#include "sgd_kernels.cuh"
#include "sgd_cluster_kernel.cuh"
#include "shuffle_kernels.cuh"
#include "shuffle2_kernels.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return fail(SGD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + \
                                    __FILE__ + ":" + std::to_string(__LINE__) + ")");       \
  } while (0)
#define SGD_TRY(expr)             \
  do {                            \
    int _rc = (expr);             \
    if (_rc != SGD_OK) return _rc; \
  } while (0)

using clk = std::chrono::steady_clock;
double ms_since(clk::time_point t0) {
  return std::chrono::duration<double, std::milli>(clk::now() - t0).count();
}

constexpr size_t kMaxDynSmem = 232448;  // 227 KB: sm_100 per-CTA opt-in maximum
constexpr int kGraphChunk = 256;        // step-kernel nodes per instantiated graph
constexpr int kLossSlots = 64;          // epochs whose running loss can be in flight uncollected

struct KernelChoice {
  int KC = 0, U = 0, NW = 0;
  int cw = 0;  // COLSPLIT: consumer warps along the columns (0: row / column / narrow mapping)
  int npmax = b200sgd::kMaxProducerWarps;  // producer warps the variant is compiled for
  bool wide = false, narrow8 = false;
  const void* fn = nullptr;
  const void* fn_lean = nullptr;  // same mapping with the run-time switches folded away, or nullptr
  const void* fn_clt = nullptr;   // same mapping with the CLUSTER step tail (launched as thread-block clusters)
  size_t fixed_smem = 0;
  int tile_rows() const { return cw > 0 ? (NW / cw) * U : NW * U; }
};

template <int KC, int U, int NW, bool WIDE = false, bool NARROW8 = false, bool WITH_LEAN = false, int CW = 0,
          int NPMAX = b200sgd::kMaxProducerWarps>
KernelChoice make_choice() {
  KernelChoice k;
  k.KC = KC; k.U = U; k.NW = NW;
  k.cw = CW;
  k.npmax = NPMAX;
  k.wide = WIDE;
  k.narrow8 = NARROW8;
  k.fn = (const void*)b200sgd::sgd_epoch_kernel<KC, U, NW, WIDE, NARROW8, false, false, CW, NPMAX>;
  if constexpr (WITH_LEAN) k.fn_lean = (const void*)b200sgd::sgd_epoch_kernel<KC, U, NW, WIDE, NARROW8, true, false, CW, NPMAX>;
  k.fn_clt = (const void*)b200sgd::sgd_epoch_kernel<KC, U, NW, WIDE, NARROW8, false, true, CW, NPMAX>;
  k.fixed_smem = b200sgd::EpochSmem<KC, NW, WIDE, U, CW>::kFixed;
  return k;
}

int variant_env() {  // B200SGD_VARIANT: A/B switch for tuning runs
  const char* e = std::getenv("B200SGD_VARIANT");
  return e ? std::atoi(e) : 0;
}

// Kernel variant by row width and, for narrow rows, by how many rows one CTA handles per step (few
// rows -> no row unrolling, so the latency-critical step tail is not slowed down by warps that only
// execute control code).  Rows of 129 .. 8192 features run the COLSPLIT mapping (sgd_kernels.cuh):
// CW warps along the columns, KC float4 columns per lane, U rows per warp and tile in registers.
// B200SGD_VARIANT=3 / 9 select the round-1 column / row mappings where they exist (A/B runs).
bool choose_kernel(int C4, int64_t rows_per_cta, KernelChoice* out) {
  const int v = variant_env();
  if (C4 <= 32) {
    if (rows_per_cta <= 8) *out = make_choice<1, 1, 8>();
    else if (rows_per_cta <= 16) *out = make_choice<1, 2, 8>();
    else if (rows_per_cta >= 128 && v == 4) *out = make_choice<1, 8, 8, false, true, true>();  // A/B: 2 x 4 rows per warp, tiles of 64 (no gain: 13.1 vs 12.8 us at B = 100000)
    else *out = make_choice<1, 4, 8, false, true, true>();  // 8 lanes per row, 4 rows per warp, tiles of 32
  } else if (C4 <= 64) {
    if (v == 8) *out = make_choice<1, 8, 8, false, false, false, 2>();   // COLSPLIT 2 x 4 warps (A/B)
    else *out = make_choice<2, 4, 8>();                                  // rows: 4 per warp in flight, tiles of 32
  } else if (C4 <= 128) {
    if (v == 8) *out = make_choice<1, 8, 8, false, false, false, 4>();   // COLSPLIT 4 x 2 warps, tiles of 16 rows (A/B)
    else if (v == 7) *out = make_choice<4, 1, 8>();                      // tiles of 8 rows (A/B)
    else *out = make_choice<4, 2, 8>();                                  // rows: 2 per warp in flight, tiles of 16
  } else if (C4 <= 256) {
    if (v == 8) *out = make_choice<1, 8, 8, false, false, false, 8>();   // COLSPLIT 8 x 1 warps (A/B)
    else *out = make_choice<8, 1, 8>();                                  // tiles of 8 rows
  } else if (C4 <= 512) {
    if (v == 3) *out = make_choice<2, 1, 8, true>();                     // round-1 column mapping (A/B)
    else if (v == 7) *out = make_choice<2, 4, 8, false, false, false, 8>();  // COLSPLIT, tiles of 4 rows (A/B)
    else *out = make_choice<2, 8, 8, false, false, false, 8>();          // COLSPLIT, tiles of 8 rows
  } else if (C4 <= 1024) {
    if (v == 3) *out = make_choice<4, 1, 8, true>();
    else *out = make_choice<4, 2, 8, false, false, false, 8>();          // tiles of 2 rows
  } else if (C4 <= 2048) {
    if (v == 3) *out = make_choice<8, 1, 8, true>();
    else *out = make_choice<8, 1, 8, false, false, false, 8>();          // one row per tile
  } else if (C4 <= 4096) {
    *out = make_choice<16, 1, 8, true>();  // column mapping: C <= 16384
  } else {
    return false;
  }
  return true;
}

}  // namespace

struct sgd_ctx {
  sgd_config cfg{};
  int64_t R = 0;  // rows of the whole data set
  int32_t C = 0, B = 0, nb = 0;
  int64_t ld = 0;  // floats per record
  int32_t C4 = 0;
  int64_t rows_per_rank = 0, row_begin = 0, row_end = 0, local_rows = 0;
  int device = 0, num_sms = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr, shuf_stream = nullptr;
  cudaEvent_t ev_shuf[2] = {nullptr, nullptr};  // device shuffle into d_perm[i] finished
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_perm[2] = {nullptr, nullptr};  // upload of d_perm[i] finished
  cudaEvent_t ev_done[3] = {nullptr, nullptr, nullptr};  // last reader of d_perm[i] done; [2]: shuffle scratch free

  // device state
  float* d_rec = nullptr;
  float* d_w = nullptr;
  float* d_wtrue = nullptr;
  int32_t* d_perm[2] = {nullptr, nullptr};  // global permutation, double-buffered across epochs
  // nranks > 1: bucketed local rows and nb+1 step offsets, one set per permutation buffer (the
  // schedule of epoch e+1 is built on the shuffle stream while epoch e still reads the other set)
  int32_t* d_perm_local[2] = {nullptr, nullptr};
  int64_t* d_step_off[2] = {nullptr, nullptr};
  int64_t* d_counts = nullptr;
  bool bucketed[2] = {false, false};        // the set belongs to the permutation in d_perm[i]
  float4* d_partials = nullptr;
  float4* d_wstage = nullptr;
  uint4* d_ll_part = nullptr;  // DATAFLOW reduction: tagged partials [grid][C4][2]
  uint4* d_ll_w = nullptr;     //                     tagged new weights [w_copies][C4][2]
  int w_copies = 1;
  uint32_t step_tag = 0;       // tags handed out so far (persistent engine)
  float* d_glast = nullptr;
  float* d_rout = nullptr;
  int64_t rout_len = 0;
  unsigned int* d_bar = nullptr;  // [0] barrier counter, [1] watchdog error word
  b200sgd::StepCtl* d_ctl = nullptr;
  unsigned long long* d_dbg = nullptr;  // phase profile, only with B200SGD_PHASE_PROFILE=1
  double* d_loss = nullptr;             // [kLossSlots][grid][2] running-loss sums, one slot per epoch
  int loss_pending = 0, loss_next = 0;  // epochs launched but not collected; next slot
  std::vector<double> curve_abs, curve_sq;  // per-epoch mean |r| and 0.5 * mean r^2 (this rank's rows)
  std::vector<int64_t> curve_rows;
  double* d_mae = nullptr;
  int mae_blocks = 0;

  // device shuffle (shuffle_kernels.cuh): raw stream / targets, hit lists in CSR form
  bool device_shuffle = false;
  uint32_t* d_tgt = nullptr;
  uint32_t* d_states = nullptr;            // per-chunk start states of the rand() stream
  uint32_t* h_states[2] = {nullptr, nullptr};
  int64_t chunk_len = 0, nchunks = 0;
  b200sgd::GlibcRand::Poly jump_chunk{};
  int32_t *d_cnt = nullptr, *d_start = nullptr, *d_hits = nullptr, *d_first = nullptr, *d_tilesum = nullptr;
  int32_t scan_tiles = 0;
  // B200SGD_SHUFFLE: 1 = the forward kernels (shuffle_kernels.cuh); backward formulation (shuffle2_kernels.cuh):
  //   2 = pairs scattered by granule (global cursors), R written in place
  //   3 = pairs by coarse bucket (global cursors) + refine, R by coarse range + place
  //   4 = pairs by coarse bucket (per-block counts, scanned (bucket, block) matrix: no global atomics) + refine, R in place
  //   5 = as 4, R by coarse range + place
  int shuf_ver = 5;
  static constexpr int kSubChunks = 64;  // device-derived sub-chunk states per host chunk state
  int cshift = 14;                      // log2 positions per coarse bucket (<= kCoarseMax buckets)
  int stream_grid = 1;                  // blocks of the stream kernels = columns of the count matrix
  bool shuf_confine_default = false;    // one GPU: permutation kernels on the SMs the step grid leaves idle
  int32_t *d_mcnt = nullptr, *d_moff = nullptr;  // (bucket, block) count matrix and its exclusive scan
  int64_t ngran = 0, ncoarse = 0, nranges = 0;
  uint2 *d_pairs1 = nullptr, *d_pairs2 = nullptr;  // coarse-bucketed pairs (later: the entries of R by range) / pairs by granule
  int32_t *d_last = nullptr, *d_R = nullptr, *d_gstart = nullptr, *d_ccnt = nullptr, *d_cstart = nullptr, *d_rcur = nullptr;
  uint32_t *d_states_x = nullptr, *d_subjump = nullptr;
  // Look-ahead: the rand() stream is never seeded, so the NEXT permutation is known in advance.  sgd_create
  // and the tail of sgd_run / sgd_iteration enqueue it (into d_perm[cur ^ 1], on the shuffle stream, beside
  // the data upload / the last epoch's step kernel); the next shuffle request just takes it.
  bool lookahead = true;        // B200SGD_NO_LOOKAHEAD=1 switches it off
  bool ahead = false;           // d_perm[cur ^ 1] is (being) filled with the permutation of shuffle shuf_no
  bool ahead_shared = false;    // ... by the shared (multi-rank) path
  int64_t reapply_k = 0;        // sgd_set_permutation replaced the input of shared shuffle k: apply its map again
  b200sgd::GlibcRand rng_saved; // the generator before the look-ahead (unshared path: rolled back by sgd_set_permutation)
  bool host_ind_stale = false;  // h_ind lags behind the device permutation
  // several ranks: the permutation work is shared (shuffle_kernels.cuh): rank (k-1) % nranks computes
  // the position map of shuffle k ahead of time, everybody applies it with one gather pass
  bool shared_shuffle = false;     // buffers exist (device shuffle, nranks > 1); used once the peers are connected
  int64_t shuf_no = 0;             // shuffles performed so far (the same number on every rank)
  int64_t own_next = 0;            // next shuffle this rank computes the map of
  int64_t shared_first = 0;        // number of the first shuffle done the shared way (earlier ones: every rank on its own)
  b200sgd::GlibcRand rng_own;      // stands at the start of shuffle own_next's stream
  size_t flag_off = 0, src_off[2] = {0, 0};  // byte offsets of the flags / the two map buffers in the exchange block
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev_own_up[2] = {nullptr, nullptr};  // upload out of h_states_own[i] finished
  uint32_t* h_states_own[2] = {nullptr, nullptr};

  // host permutation state
  std::vector<int32_t> h_ind;                // the reference's ind_vector (main.cu:421-424)
  int32_t* h_stage[2] = {nullptr, nullptr};  // pinned upload staging
  b200sgd::GlibcRand rng;
  int cur = 0;  // which d_perm / h_stage holds the active permutation
  bool perm_valid = false, data_loaded = false, have_true_w = false;
  float last_shuffle_ms = 0.f;
  int last_shuffle_launches = 0;

  // kernel configuration
  KernelChoice kc{};
  int grid = 0, reduce_mode = 0, ntiles = 0, tile_log2 = 0, engine = SGD_ENGINE_PERSISTENT;
  CUtensorMap rec_map{};
  bool use_gather4 = false;  // narrow rows: TMA tile::gather4, four rows per instruction
  bool have_rec_map = false;  // the tensor map of the records exists
  // sgd_cluster_epoch_kernel (small steps of narrow rows): the grid is one cluster of cluster_cs CTAs
  int cluster_cs = 0, cluster_ng = 1, cluster_ntiles = 0, cluster_tile_log2 = 0;  // CTAs per cluster, clusters
  const void* cluster_fn = nullptr;
  bool cluster_coop = true;        // pass the cooperative attribute with several clusters (dropped if refused)
  uint2* d_cl_inbox = nullptr;     // one GPU, several clusters: tagged cluster sums [2][clusters][4*C4p]
  size_t cluster_smem = 0;
  // CLUSTER step tail of the generic kernel (sgd_kernels.cuh, CLT): clt_ng clusters of clt_cs CTAs
  bool clt = false;
  int clt_cs = 0, clt_ng = 0, clt_slice = 0;
  size_t clt_extra = 0;          // shared memory of the exchange (before the ring)
  uint4* d_clt_inbox = nullptr;  // one GPU: tagged slice sums of the clusters [2][clusters][C4p][2]
  int inflight_tiles = 0;        // producers keep at most this many ring tiles in flight (0: the whole ring)
  int nprod = 1;  // producer warps
  int tile_rows = 0;
  size_t smem_bytes = 0;
  cudaGraphExec_t graph_exec = nullptr;
  int graph_nodes = 0;

  // multi-GPU
  void* d_xchg_block = nullptr;  // inbox of tagged slice sums [2][nranks][C4][2] uint4 | direct partials
  size_t xchg_bytes = 0, xpart_off = 0;
  bool direct = false;           // narrow rows: partials pushed straight to every rank's owners
  void* peer_block[8] = {nullptr};
  bool peer_open[8] = {false};
  bool peers_ready = false;
};

namespace {

int check_ctx(const sgd_ctx* c) {
  if (!c) return fail(SGD_ERR_ARG, "null context");
  return SGD_OK;
}

int bind(const sgd_ctx* c) {
  CUDA_TRY(cudaSetDevice(c->device));
  return SGD_OK;
}

// main.cu:343-344: num_batches = (int)floor(R / (float)B).  The float division rounds up for some
// R > 2^24 (R = 99,999,999, B = 1000 gives 100,000 batches = R + 1 rows), where the reference reads
// past its index vector; the integer quotient is the same value wherever the reference is well
// defined, so the result is clamped to it.
void batch_geometry(int64_t R, int32_t batch_arg, int32_t* B, int32_t* nb) {
  *B = batch_arg == 0 ? (int32_t)R : batch_arg;
  *nb = 0;
  if (*B > 0) {
    const int64_t f = (int64_t)std::floor((float)R / (float)(*B));
    *nb = (int32_t)std::min<int64_t>(f, R / *B);
  }
}

void free_all(sgd_ctx* c) {
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->shuf_stream) cudaStreamSynchronize(c->shuf_stream);
  for (int i = 0; i < 8; ++i)
    if (c->peer_open[i] && c->peer_block[i]) cudaIpcCloseMemHandle(c->peer_block[i]);
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  cudaFree(c->d_rec); cudaFree(c->d_w); cudaFree(c->d_wtrue);
  cudaFree(c->d_perm[0]); cudaFree(c->d_perm[1]); cudaFree(c->d_perm_local[0]); cudaFree(c->d_perm_local[1]);
  cudaFree(c->d_step_off[0]); cudaFree(c->d_step_off[1]); cudaFree(c->d_counts); cudaFree(c->d_partials); cudaFree(c->d_wstage);
  cudaFree(c->d_glast); cudaFree(c->d_rout); cudaFree(c->d_bar); cudaFree(c->d_ctl);
  cudaFree(c->d_ll_part); cudaFree(c->d_ll_w); cudaFree(c->d_dbg); cudaFree(c->d_loss);
  cudaFree(c->d_tgt); cudaFree(c->d_cnt); cudaFree(c->d_start); cudaFree(c->d_hits); cudaFree(c->d_first);
  cudaFree(c->d_tilesum); cudaFree(c->d_states);
  cudaFree(c->d_pairs1); cudaFree(c->d_pairs2); cudaFree(c->d_last); cudaFree(c->d_R); cudaFree(c->d_gstart);
  cudaFree(c->d_mcnt); cudaFree(c->d_moff);
  cudaFree(c->d_ccnt); cudaFree(c->d_cstart); cudaFree(c->d_rcur); cudaFree(c->d_states_x); cudaFree(c->d_subjump);
  if (c->own_stream) cudaStreamSynchronize(c->own_stream);
  if (c->h_states[0]) cudaFreeHost(c->h_states[0]);
  if (c->h_states[1]) cudaFreeHost(c->h_states[1]);
  for (int i = 0; i < 2; ++i) {
    if (c->h_states_own[i]) cudaFreeHost(c->h_states_own[i]);
    if (c->ev_own_up[i]) cudaEventDestroy(c->ev_own_up[i]);
  }
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  cudaFree(c->d_mae); cudaFree(c->d_xchg_block); cudaFree(c->d_cl_inbox); cudaFree(c->d_clt_inbox);
  if (c->h_stage[0]) cudaFreeHost(c->h_stage[0]);
  if (c->h_stage[1]) cudaFreeHost(c->h_stage[1]);
  cudaEvent_t evs[] = {c->ev0, c->ev1, c->ev_perm[0], c->ev_perm[1], c->ev_done[0], c->ev_done[1], c->ev_done[2]};
  for (cudaEvent_t e : evs)
    if (e) cudaEventDestroy(e);
  if (c->ev_shuf[0]) cudaEventDestroy(c->ev_shuf[0]);
  if (c->ev_shuf[1]) cudaEventDestroy(c->ev_shuf[1]);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->shuf_stream) cudaStreamDestroy(c->shuf_stream);
}

// Decide grid size, reduction mode and ring depth for this shape (DESIGN.md "launch geometry").
int configure_launch(sgd_ctx* c) {
  if (c->C4 > 4096) return fail(SGD_ERR_ARG, "cols > 16384 is not supported by this build of the step kernel");
  // rows this rank sees per step
  const int64_t rows_step = std::max<int64_t>(1, (int64_t)c->B / std::max(1, c->cfg.nranks));
  int grid = c->cfg.grid;
  if (grid <= 0) {
    // Small steps are bound by the step tail (publish -> gather -> publish -> poll, ~1.7 us at 32
    // CTAs, ~2.7 us at 148, measured), which grows with the number of CTAs, so they get
    // ~24 rows per CTA; once a step moves a few MB the row stream dominates and every SM is used.
    const int64_t step_bytes = rows_step * c->ld * 4;
    if (step_bytes < (2ll << 20)) {
      // power of two: measured 3.3 us/step at 32 or 64 CTAs vs 4.8-5.8 us at 41, 48 or 100 on
      // the same shape (some owners' polling then runs 2.5x late; cause not pinned down)
      grid = 1;
      while (grid * 2 <= 64 && (int64_t)grid * 2 * 24 <= rows_step) grid *= 2;
    } else if (step_bytes < (400ll << 20) && c->num_sms >= 128) {
      // 128 CTAs beat 148 until a step streams several hundred MB (C = 512, B = 8192: 7.3 vs 8.8 us;
      // C = 2048, B = 4096: 13.6 vs 14.8 us; C = 2048, B = 65536: 94.6 vs 93.3 us): the gather tree
      // of the tail is fully populated at a power of two
      grid = 128;
    } else {
      grid = c->num_sms;
    }
  }
  // Small steps of narrow rows: the whole grid is one thread-block cluster that exchanges the gradient
  // through distributed shared memory (sgd_cluster_kernel.cuh); multi-GPU adds one NVLink hop.  B200SGD_CLUSTER=0
  // switches it off, 8 / 16 pick the cluster size (A/B runs).
  c->cluster_cs = 0;
  c->cluster_ng = 1;
  {
    const char* e = std::getenv("B200SGD_CLUSTER");
    // 16 CTAs per cluster; 8 for very small steps (100 rows: 0.85 us per step against 0.97: half the messages)
    const int want = e ? std::atoi(e) : (rows_step <= 256 ? 8 : b200sgd::kClMaxCS);
    const char* em = std::getenv("B200SGD_CLUSTER_MAX_ROWS");  // largest step (rows per rank) the cluster kernel takes
    // measured at C = 100 (us per step, cluster / generic kernel): 4096 rows 2.79 / 3.87, 8192 3.16 / 4.4,
    // 10000 3.23 / 4.6, 16384 5.1 / 4.9, 30000 6.6 / 6.2, 100000 15.2 / 12.7: beyond ~12k rows the generic
    // kernel wins (it streams on 128 SMs, 7 clusters are 112, and the per-SM TMA row rate is the limit)
    const int64_t max_rows = em ? std::atoll(em) : 10240;
    if ((want == 8 || want == 16) && c->ld <= 128 && c->cfg.engine != SGD_ENGINE_GRAPH &&
        (c->cfg.reduce == SGD_REDUCE_AUTO || c->cfg.reduce == SGD_REDUCE_DATAFLOW) &&
        !(c->cfg.flags & SGD_FLAG_RESIDUALS) && rows_step >= 48 && rows_step <= max_rows) {
      // clusters: the L2 hop between clusters costs ~1 us per step, a ring tile of 32 rows ~0.2-0.3 us
      // (2000 rows: 1.93 us on one cluster, 2.61 on two; 3000 rows: 3.54 / 2.86), so one cluster serves
      // steps of up to 2048 rows (one four-tile block per CTA); beyond, enough clusters for <= 64 rows
      // per CTA and step, as far as they fit on the GPU at once (configure_cluster: 7 clusters of 16
      // CTAs on a B200).  Several GPUs: one cluster per GPU.
      int ng = 1;
      if (c->cfg.nranks == 1 && rows_step > 2048)
        ng = (int)std::min<int64_t>(c->num_sms / want, (rows_step + 64 * want - 1) / (64 * want));
      if (const char* en = std::getenv("B200SGD_CLUSTERS")) ng = std::max(1, std::atoi(en));
      if (c->cfg.grid > 0 && c->cfg.grid % want == 0) ng = c->cfg.grid / want;  // explicit grid: that many clusters
      if (c->cfg.nranks > 1) ng = 1;
      if ((c->cfg.grid <= 0 || c->cfg.grid == ng * want) && ng * want <= c->num_sms) {
        c->cluster_cs = want;
        c->cluster_ng = ng;
        grid = ng * want;
      }
    }
  }
  grid = std::max(1, std::min(grid, c->num_sms));
  const size_t slot_bytes = (size_t)c->ld * 4;
  // Kernel variant, producer warps and ring for `g` CTAs and `extra` bytes of exchange buffers in
  // front of the ring.  The ring takes every tile that fits (any count >= 2, not only powers of two).
  auto geometry = [&](int g, size_t extra) -> int {
    if (!choose_kernel(c->C4, (rows_step + g - 1) / g, &c->kc)) return fail(SGD_ERR_ARG, "unsupported row width");
    // producer warps: one TMA bulk copy per row is issued lane by lane (~30-40 cycles each), so
    // short rows need several issuing warps to keep up with HBM; long rows and tiny batches do not
    const int64_t rows_cta = (rows_step + g - 1) / g;
    int np = 1;
    if (slot_bytes <= 1024) np = rows_cta >= 64 ? 4 : 2;
    else if (slot_bytes <= 4096) np = rows_cta >= 32 ? 2 : 1;
    if (const char* e = std::getenv("B200SGD_PRODUCER_WARPS")) np = std::atoi(e);
    c->nprod = std::max(1, std::min(np, c->kc.npmax));
    int tile_rows = c->kc.tile_rows();
    const size_t avail = kMaxDynSmem - c->kc.fixed_smem - extra;
    if (c->kc.wide) {
      // column mapping: as many rows per tile as two tiles allow (<= 32, a multiple of NW if possible)
      tile_rows = (int)std::min<size_t>(32, avail / (2 * slot_bytes));
      if (tile_rows >= c->kc.NW) tile_rows -= tile_rows % c->kc.NW;
      if (tile_rows < 1) return fail(SGD_ERR_ARG, "row too wide for the shared-memory ring");
    }
    c->tile_rows = tile_rows;
    const size_t tile_bytes = (size_t)tile_rows * slot_bytes;
    if (2 * tile_bytes > avail) return fail(SGD_ERR_ARG, "row too wide for the shared-memory ring");
    int ntiles = (int)std::min<size_t>(b200sgd::kMaxTiles, avail / tile_bytes);
    if (const char* e = std::getenv("B200SGD_RING_TILES")) ntiles = std::max(2, std::min(ntiles, std::atoi(e)));
    c->nprod = std::min(c->nprod, ntiles);
    c->ntiles = ntiles;
    c->tile_log2 = 0;
    c->smem_bytes = c->kc.fixed_smem + extra + (size_t)ntiles * tile_bytes;
    // Experiment switch: cap the bytes in flight per CTA.  Measured (profiles/r2_tail_matrix.md): no
    // effect on the step tail down to 64 KB per SM, bandwidth lost below that -> off by default.
    int64_t kb = 0;
    if (const char* e = std::getenv("B200SGD_INFLIGHT_KB")) kb = std::atoll(e);
    int infl = kb > 0 ? (int)std::max<int64_t>(1, (kb * 1024 + (int64_t)tile_bytes / 2) / (int64_t)tile_bytes) : 0;
    if (infl >= ntiles) infl = 0;  // the ring itself is the bound
    c->inflight_tiles = infl;
    return SGD_OK;
  };
  // CLUSTER step tail (sgd_kernels.cuh, CLT) for everything the cluster kernel above does not take and
  // that runs on at least two clusters' worth of CTAs: the grid becomes clt_ng clusters of clt_cs CTAs.
  c->clt = false;
  if (c->cluster_cs == 0 && c->cfg.engine != SGD_ENGINE_GRAPH &&
      (c->cfg.reduce == SGD_REDUCE_AUTO || c->cfg.reduce == SGD_REDUCE_CLUSTER)) {
    // CTAs per cluster: 16 (7 clusters = 112 SMs on a B200) while the step tail matters -- fewer clusters
    // meet in L2: C = 2048, B = 4096: 8.9 us per step against 9.9 with 8 --, 8 (15 clusters = 120 SMs)
    // once a step streams >= 64 MB and the number of SMs counts (C = 512, B = 65536: 0.82 vs 0.79 of peak)
    int cs = rows_step * c->ld * 4 >= (64ll << 20) ? 8 : 16;
    if (cs == 16 && c->cfg.grid <= 0 && grid < 32) cs = 8;  // small grids: two clusters of 8 rather than none
    if (c->cfg.grid > 0 && c->cfg.grid % 16 != 0) cs = 8;
    if (const char* e = std::getenv("B200SGD_CLT")) cs = std::atoi(e);  // 0 = off; 2, 4, 8, 16 = CTAs per cluster
    const bool cs_ok = cs == 2 || cs == 4 || cs == 8 || cs == 16;
    const bool grid_ok = c->cfg.grid <= 0 ? grid >= 2 * cs : (c->cfg.grid % cs == 0);
    // (narrow rows on one GPU keep the LEAN 8-lanes-per-row kernel: B = 100000 at C = 100 runs 12.3 us per
    // step there against 13.1 with the cluster tail, whose instantiation keeps its run-time switches)
    const bool narrow_single = c->C4 <= 32 && c->cfg.nranks == 1 && c->cfg.reduce != SGD_REDUCE_CLUSTER && !std::getenv("B200SGD_CLT");
    if (cs_ok && grid_ok && c->C4 <= 2048 && !narrow_single) {
      const int slice = (c->C4 + cs - 1) / cs;
      int g = c->cfg.grid > 0 ? c->cfg.grid : std::max(cs, grid / cs * cs);
      for (int pass = 0; pass < 3; ++pass) {
        // exchange region: 4 barriers | reduce-scatter buffers [2][cs][slice] | cross-cluster partials [consumer threads]
        choose_kernel(c->C4, (rows_step + g - 1) / g, &c->kc);
        const size_t extra = (128 + (size_t)32 * cs * slice + (size_t)16 * 32 * c->kc.NW + 127) / 128 * 128;
        if (geometry(g, extra) != SGD_OK) break;
        const void* fn = c->kc.fn_clt;
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
        if (cs > 8) CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(g);
        lc.blockDim = dim3(32 * (c->kc.NW + c->nprod));
        lc.dynamicSmemBytes = c->smem_bytes;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        lc.attrs = at;
        lc.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, fn, &lc) != cudaSuccess || nclusters < 1) {
          cudaGetLastError();
          break;
        }
        if (std::getenv("B200SGD_DEBUG"))
          std::fprintf(stderr, "[b200sgd] cluster tail: %d CTAs per cluster, at most %d clusters resident, %d CTAs wanted\n", cs,
                       nclusters, g);
        int want = g / cs;
        if (c->cfg.grid <= 0 && grid >= c->num_sms - 20) want = nclusters;  // streaming steps: every cluster the GPU holds
        if (const char* en = std::getenv("B200SGD_CLUSTERS")) want = std::max(1, std::atoi(en));
        const int ng = std::min(want, nclusters);
        if (c->cfg.grid > 0 && ng * cs != c->cfg.grid) break;  // an explicit grid that does not fit
        if (ng * cs == g) {
          c->clt = true;
          c->clt_cs = cs;
          c->clt_ng = ng;
          c->clt_slice = slice;
          c->clt_extra = extra;
          grid = g;
          break;
        }
        g = ng * cs;  // re-derive the per-CTA geometry for the grid that fits
      }
    }
  }
  if (!c->clt) SGD_TRY(geometry(grid, 0));
  CUDA_TRY(cudaFuncSetAttribute(c->kc.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
  if (c->kc.fn_lean)
    CUDA_TRY(cudaFuncSetAttribute(c->kc.fn_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, c->clt ? c->kc.fn_clt : c->kc.fn, 32 * (c->kc.NW + c->nprod),
                                                         c->smem_bytes));
  if (occ < 1) return fail(SGD_ERR_CUDA, "step kernel does not fit on an SM");
  c->grid = grid;
  int mode = c->cfg.reduce;
  if (c->cfg.nranks > 1) mode = SGD_REDUCE_DATAFLOW;  // the fused NVLink exchange lives there
  if (mode < SGD_REDUCE_ALLGATHER || mode > SGD_REDUCE_DATAFLOW_1HOP) mode = SGD_REDUCE_DATAFLOW;
  if (c->clt) mode = SGD_REDUCE_CLUSTER;
  c->reduce_mode = mode;
  c->engine = (c->cfg.engine == SGD_ENGINE_GRAPH) ? SGD_ENGINE_GRAPH : SGD_ENGINE_PERSISTENT;
  if (c->engine == SGD_ENGINE_GRAPH && c->cfg.nranks > 1)
    return fail(SGD_ERR_ARG, "the graph engine is single-GPU; use the persistent engine with nranks > 1");
  return SGD_OK;
}

// Ring geometry, attributes and schedulability of the cluster kernel; falls back to the generic
// kernel (cluster_cs = 0) when the tensor map or a free GPC for the cluster is missing.
int configure_cluster(sgd_ctx* c) {
  if (c->cluster_cs == 0) return SGD_OK;
  const int cs = c->cluster_cs;
  const int grid0 = c->grid;  // what the generic kernel's buffers and geometry were derived for
  c->cluster_cs = 0;
  if (!c->have_rec_map) return SGD_OK;
  const size_t tile_bytes = (size_t)b200sgd::kClTileRows * c->ld * 4;
  int ntiles = 2, lg = 1;
  while (ntiles * 2 <= b200sgd::kClMaxTiles && b200sgd::kClFixedSmem + (size_t)(ntiles * 2) * tile_bytes <= kMaxDynSmem) {
    ntiles *= 2;
    ++lg;
  }
  c->cluster_ntiles = ntiles;
  c->cluster_tile_log2 = lg;
  c->cluster_smem = b200sgd::kClFixedSmem + (size_t)ntiles * tile_bytes;
  const bool prof = c->d_dbg != nullptr;
  const int64_t rows_rank = std::max<int64_t>(1, (int64_t)c->B / c->cfg.nranks);
  for (;;) {  // until the wanted number of clusters is resident at once
    // instantiation: tiles of one CTA per step -> phase-A block size; more than one group -> MULTI
    using namespace b200sgd;
    const bool multi = c->cfg.nranks > 1 || c->cluster_ng > 1;
    const int64_t ctas = (int64_t)cs * c->cluster_ng;
    const int tiles = (int)(((rows_rank + ctas - 1) / ctas + kClTileRows - 1) / kClTileRows);
    const void* fn = nullptr;
    if (multi) {
      fn = tiles <= 1 ? (const void*)sgd_cluster_epoch_kernel<1, false, true>
         : tiles <= 2 ? (const void*)sgd_cluster_epoch_kernel<2, false, true>
                      : (const void*)sgd_cluster_epoch_kernel<4, false, true>;
    } else if (prof) {
      fn = tiles <= 1 ? (const void*)sgd_cluster_epoch_kernel<1, true> : tiles <= 2 ? (const void*)sgd_cluster_epoch_kernel<2, true>
                                                                                    : (const void*)sgd_cluster_epoch_kernel<4, true>;
    } else {
      fn = tiles <= 1 ? (const void*)sgd_cluster_epoch_kernel<1> : tiles <= 2 ? (const void*)sgd_cluster_epoch_kernel<2>
                                                                              : (const void*)sgd_cluster_epoch_kernel<4>;
    }
    c->cluster_fn = fn;
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->cluster_smem));
    if (cs > 8) CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(cs * c->cluster_ng);
    lc.blockDim = dim3(32 * (kClNW + kClNP));
    lc.dynamicSmemBytes = c->cluster_smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    int nclusters = 0;
    // every cluster of the launch must be resident at once: they wait for one another's sums
    if (cudaOccupancyMaxActiveClusters(&nclusters, fn, &lc) != cudaSuccess || nclusters < 1) {
      cudaGetLastError();
      c->grid = grid0;  // generic kernel with the geometry configure_launch derived
      return SGD_OK;
    }
    if (std::getenv("B200SGD_DEBUG"))
      std::fprintf(stderr, "[b200sgd] cluster size %d: at most %d clusters resident, %d wanted\n", cs, nclusters, c->cluster_ng);
    if (nclusters >= c->cluster_ng) break;
    if (c->cfg.grid > 0) {  // an explicit grid that does not fit: generic kernel
      c->grid = grid0;
      return SGD_OK;
    }
    c->cluster_ng = nclusters;           // as many as fit (a B200 keeps 7 clusters of 16 CTAs, 15 of 8)
    c->grid = cs * c->cluster_ng;
  }
  if (c->cluster_ng > 1 && c->cfg.nranks == 1) {  // the clusters of this GPU meet in a local inbox of tagged sums
    const size_t words = 2 * (size_t)c->cluster_ng * 4 * (((size_t)c->C4 + 3) & ~(size_t)3);
    CUDA_TRY(cudaMalloc(&c->d_cl_inbox, sizeof(uint2) * words));
    CUDA_TRY(cudaMemset(c->d_cl_inbox, 0, sizeof(uint2) * words));
  }
  c->cluster_cs = cs;
  return SGD_OK;
}

// 2-D tensor map of the records for tile::gather4 (rows of at most 256 floats).  The driver entry
// point is resolved at run time, so the library does not link libcuda.
int make_record_tensor_map(sgd_ctx* c) {
  c->use_gather4 = false;
  c->have_rec_map = false;
  if (c->ld > 256 || c->local_rows < 1 || std::getenv("B200SGD_NO_GATHER4")) return SGD_OK;
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return SGD_OK;  // per-row bulk copies still work
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)c->ld, (cuuint64_t)c->local_rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)c->ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)c->ld, 1};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;
  if (const char* e = std::getenv("B200SGD_L2_PROMO")) {  // experiment: 1 = 64 B, 2 = 128 B, 3 = 256 B
    const int v = std::atoi(e);
    promo = v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
          : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
  }
  const CUresult r = ((encode_fn)fn)(&c->rec_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c->d_rec, gdim, gstride, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  c->have_rec_map = (r == CUDA_SUCCESS);
  // the generic kernel issues gather4 when its ring tile is a whole number of four-row groups
  c->use_gather4 = c->have_rec_map && c->kc.tile_rows() % 4 == 0 && !c->kc.wide;
  return SGD_OK;
}

void fill_params(sgd_ctx* c, b200sgd::EpochParams* p, int32_t epoch, int buf) {
  std::memset(p, 0, sizeof(*p));
  p->rec_map = c->rec_map;
  p->use_gather4 = c->use_gather4 ? 1 : 0;
  p->rec = c->d_rec;
  p->ld = c->ld;
  p->C = c->C;
  p->C4 = c->C4;
  p->perm = c->cfg.nranks > 1 ? c->d_perm_local[buf] : c->d_perm[buf];
  p->step_off = c->cfg.nranks > 1 ? c->d_step_off[buf] : nullptr;
  p->B = c->B;
  p->nb = c->nb;
  p->first_step = 0;
  // main.cu:144 / :282: float a = -(learning_rate / std::pow(epoch, 0.25));
  p->a = (float)(-((double)c->cfg.learning_rate / std::pow((double)epoch, 0.25)));
  p->w = c->d_w;
  p->partials = c->d_partials;
  p->w_stage = c->d_wstage;
  p->g_last = c->d_glast;
  p->r_out = c->d_rout;
  p->bar = c->d_bar;
  p->err = c->d_bar + 1;
  p->reduce_mode = c->reduce_mode;
  p->ntiles = c->ntiles;
  p->tile_log2 = c->tile_log2;
  p->tile_rows = c->tile_rows;
  p->slot_bytes = (uint32_t)(c->ld * 4);
  p->nprod = c->nprod;
  {
    const char* e = std::getenv("B200SGD_PREFETCH_STEPS");
    p->prefetch_steps = e ? std::atoi(e) : 0;
  }
  p->ll_part = c->d_ll_part;
  p->ll_w = c->d_ll_w;
  p->w_copies = c->w_copies;
  p->tag_base = c->step_tag;
  p->rank = c->cfg.rank;
  p->nranks = c->cfg.nranks;
  p->ctl = nullptr;
  p->dbg = c->d_dbg;
  p->inflight_tiles = c->inflight_tiles;
  if (const char* e = std::getenv("B200SGD_DBGSKIP")) p->dbg_skip = std::atoi(e);
  if (c->clt) {
    p->cl_cs = c->clt_cs;
    p->cl_ng = c->clt_ng;
    p->cl_slice = c->clt_slice;
    p->cl_extra = (uint32_t)c->clt_extra;
    if (c->cfg.nranks == 1) {
      p->cl_inbox_local = c->d_clt_inbox;
      p->cl_inbox_peer[0] = c->d_clt_inbox;
    } else {
      p->cl_inbox_local = (uint4*)c->d_xchg_block;
      for (int r = 0; r < c->cfg.nranks; ++r) p->cl_inbox_peer[r] = (uint4*)c->peer_block[r];
    }
  }
}

int upload_perm(sgd_ctx* c, int buf) {
  // h_stage[buf] -> d_perm[buf] on the copy stream, once the last reader of d_perm[buf] is done;
  // the compute stream waits on ev_perm[buf] before it uses the buffer
  c->bucketed[buf] = false;
  CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_done[buf], 0));
  CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_shuf[buf], 0));  // a discarded look-ahead may still be writing it
  CUDA_TRY(cudaMemcpyAsync(c->d_perm[buf], c->h_stage[buf], sizeof(int32_t) * (size_t)c->R,
                           cudaMemcpyHostToDevice, c->copy_stream));
  CUDA_TRY(cudaEventRecord(c->ev_perm[buf], c->copy_stream));
  return SGD_OK;
}

// nranks > 1: keep, per step, the rows this rank owns.  Runs on `st`: the compute stream, or
// the shuffle stream right behind the shared shuffle (then it is off the critical path between epochs).
int bucket_perm(sgd_ctx* c, int buf, int* launches, cudaStream_t st = nullptr) {
  if (c->cfg.nranks == 1 || c->nb == 0 || c->bucketed[buf]) return SGD_OK;
  if (!st) st = c->stream;
  const int blocks = std::min(c->nb, 4 * c->num_sms);
  b200sgd::sgd_bucket_count_kernel<<<blocks, 256, 0, st>>>(
      c->d_perm[buf], c->B, c->nb, c->row_begin, c->row_end, c->d_counts);
  b200sgd::sgd_scan_kernel<<<1, 1024, 0, st>>>(c->d_counts, c->nb, c->d_step_off[buf]);
  b200sgd::sgd_bucket_write_kernel<<<blocks, 256, 0, st>>>(
      c->d_perm[buf], c->B, c->nb, c->row_begin, c->row_end, c->d_step_off[buf], c->d_perm_local[buf]);
  *launches += 3;
  CUDA_TRY(cudaGetLastError());
  c->bucketed[buf] = true;
  return SGD_OK;
}

// GRAPH engine: capture kGraphChunk parameter-free step kernels once; replay per epoch.
int build_graph(sgd_ctx* c) {
  if (c->graph_exec) return SGD_OK;
  const int nodes = std::max(1, std::min(c->nb, kGraphChunk));
  b200sgd::EpochParams p;
  fill_params(c, &p, 1, 0);
  p.nb = 1;
  p.ctl = c->d_ctl;
  void* args[] = {&p};
  cudaGraph_t graph = nullptr;
  CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < nodes && e == cudaSuccess; ++i)
    e = cudaLaunchKernel(c->kc.fn, dim3(c->grid), dim3(32 * (c->kc.NW + c->nprod)), args, c->smem_bytes, c->stream);
  cudaError_t e2 = cudaStreamEndCapture(c->stream, &graph);
  if (e != cudaSuccess || e2 != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    return fail(SGD_ERR_CUDA, std::string("graph capture failed: ") +
                                  cudaGetErrorString(e != cudaSuccess ? e : e2));
  }
  cudaError_t e3 = cudaGraphInstantiate(&c->graph_exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e3 != cudaSuccess) return fail(SGD_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e3));
  c->graph_nodes = nodes;
  return SGD_OK;
}

// Enqueues the step loop of one epoch on the compute stream (no host sync).
int launch_epoch(sgd_ctx* c, int32_t epoch, int buf, int* launches) {
  if (c->nb == 0) return SGD_OK;
  b200sgd::EpochParams p;
  fill_params(c, &p, epoch, buf);
  if (c->cfg.nranks > 1) {
    if (!c->peers_ready) return fail(SGD_ERR_STATE, "nranks > 1 but sgd_peer_ready was not called");
    p.ll_inbox_local = (uint4*)c->d_xchg_block;
    for (int r = 0; r < c->cfg.nranks; ++r) p.ll_inbox_peer[r] = (uint4*)c->peer_block[r];
    p.direct = c->direct ? 1 : 0;
    if (c->direct) {
      p.ll_xpart_local = (uint4*)((char*)c->d_xchg_block + c->xpart_off);
      for (int r = 0; r < c->cfg.nranks; ++r) p.ll_xpart_peer[r] = (uint4*)((char*)c->peer_block[r] + c->xpart_off);
    }
  }
  // running-loss slot of this epoch (collected after the next host sync)
  if (c->loss_pending >= kLossSlots) return fail(SGD_ERR_STATE, "too many epochs in flight without a host sync");
  double* loss_slot = c->d_loss + (size_t)c->loss_next * 2 * c->grid;
  CUDA_TRY(cudaMemsetAsync(loss_slot, 0, sizeof(double) * 2 * (size_t)c->grid, c->stream));
  p.loss = loss_slot;
  c->loss_next = (c->loss_next + 1) % kLossSlots;
  c->loss_pending += 1;
  if (c->engine == SGD_ENGINE_GRAPH) {
    SGD_TRY(build_graph(c));
    // per-epoch state of the parameter-free nodes: {perm, a, step = 0, nb}; bar_base keeps counting
    b200sgd::StepCtl head{};
    head.perm = p.perm;
    head.a = p.a;
    head.step = 0;
    head.nb_total = c->nb;
    head.loss = loss_slot;
    CUDA_TRY(cudaMemcpyAsync(c->d_ctl, &head, offsetof(b200sgd::StepCtl, bar_base), cudaMemcpyHostToDevice, c->stream));
    const int replays = (c->nb + c->graph_nodes - 1) / c->graph_nodes;
    for (int i = 0; i < replays; ++i) CUDA_TRY(cudaGraphLaunch(c->graph_exec, c->stream));
    *launches += c->nb;
    return SGD_OK;
  }
  CUDA_TRY(cudaMemsetAsync(c->d_bar, 0, sizeof(unsigned int), c->stream));
  c->step_tag += (uint32_t)c->nb;  // every rank advances identically: tags agree across GPUs
  void* args[] = {&p};
  if (c->cluster_cs > 0) {  // narrow rows, small and mid-size steps: clusters, gradient exchange in shared memory
    p.ntiles = c->cluster_ntiles;
    p.tile_log2 = c->cluster_tile_log2;
    if (c->cfg.nranks == 1 && c->cluster_ng > 1) {  // the clusters of this GPU meet in a local inbox
      p.ll_inbox_local = (uint4*)c->d_cl_inbox;
      p.ll_inbox_peer[0] = (uint4*)c->d_cl_inbox;
    }
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(c->cluster_cs * c->cluster_ng);
    lc.blockDim = dim3(32 * (b200sgd::kClNW + b200sgd::kClNP));
    lc.dynamicSmemBytes = c->cluster_smem;
    lc.stream = c->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c->cluster_cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;  // several clusters wait for one another: co-residency
    at[1].val.cooperative = 1;
    lc.attrs = at;
    lc.numAttrs = (c->cluster_ng > 1 && c->cluster_coop) ? 2 : 1;
    cudaError_t e = cudaLaunchKernelExC(&lc, c->cluster_fn, args);
    if (e != cudaSuccess && lc.numAttrs == 2) {  // cooperative + cluster refused: the occupancy check stands in
      cudaGetLastError();
      c->cluster_coop = false;
      lc.numAttrs = 1;
      e = cudaLaunchKernelExC(&lc, c->cluster_fn, args);
    }
    CUDA_TRY(e);
    *launches += 1;
    return SGD_OK;
  }
  if (c->clt) {  // generic kernel, cluster step tail: clusters wait for one another, so cooperative as well
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(c->grid);
    lc.blockDim = dim3(32 * (c->kc.NW + c->nprod));
    lc.dynamicSmemBytes = c->smem_bytes;
    lc.stream = c->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c->clt_cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    lc.attrs = at;
    // (B200SGD_NO_COOP=1: ncu cannot replay a launch that is both cooperative and clustered; without the
    // attribute co-residency rests on the occupancy check of configure_launch, which holds on an idle GPU)
    lc.numAttrs = (c->clt_ng > 1 && c->cluster_coop && !std::getenv("B200SGD_NO_COOP")) ? 2 : 1;
    cudaError_t e = cudaLaunchKernelExC(&lc, c->kc.fn_clt, args);
    if (e != cudaSuccess && lc.numAttrs == 2) {  // cooperative + cluster refused: the occupancy check stands in
      cudaGetLastError();
      c->cluster_coop = false;
      lc.numAttrs = 1;
      e = cudaLaunchKernelExC(&lc, c->kc.fn_clt, args);
    }
    CUDA_TRY(e);
    *launches += 1;
    return SGD_OK;
  }
  // the instantiation without run-time switches, when none of them is in use (see LEAN)
  const bool lean = c->kc.fn_lean != nullptr && p.ctl == nullptr && p.step_off == nullptr && p.prefetch_steps == 0 &&
                    p.use_gather4 != 0 && p.r_out == nullptr && p.dbg == nullptr && p.nranks == 1 &&
                    p.reduce_mode == SGD_REDUCE_DATAFLOW && std::getenv("B200SGD_NO_LEAN") == nullptr;
  CUDA_TRY(cudaLaunchCooperativeKernel(lean ? c->kc.fn_lean : c->kc.fn, dim3(c->grid),
                                       dim3(32 * (c->kc.NW + c->nprod)), args, c->smem_bytes, c->stream));
  *launches += 1;
  return SGD_OK;
}

size_t shuf_smem_bytes();

// Host part of one device shuffle: the start state of every chunk of the n-1 values of the rand()
// stream this epoch consumes (main.cu:90 / :207 consume exactly n-1 per std::random_shuffle), by
// jump-ahead; then the host generator is moved past the epoch.  ~1 us per chunk.
void prepare_stream_states(sgd_ctx* c, int buf) {
  c->rng.chunk_states(c->jump_chunk, c->nchunks, c->h_states[buf]);
  c->rng.skip((uint64_t)(c->R - 1));
}

// The kernels of one device shuffle on stream `st`: d_states (uploaded) -> out[p] = in[src(p)], or the
// position map src itself (in == nullptr).  `blocks`: grid of the grid-stride kernels.
int enqueue_perm_kernels(sgd_ctx* c, cudaStream_t st, int blocks, size_t smem, const int32_t* in, int32_t* out, int* launches) {
  const int64_t n = c->R;
  if (c->shuf_ver == 1) {  // forward formulation (shuffle_kernels.cuh)
    CUDA_TRY(cudaMemsetAsync(c->d_cnt, 0, sizeof(int32_t) * ((size_t)n + 1), st));
    b200sgd::shuf_targets_stream_kernel<<<(int)((c->nchunks + 127) / 128), 128, smem, st>>>(c->d_states, c->chunk_len, c->nchunks, n,
                                                                                     c->d_tgt, c->d_cnt);
    b200sgd::scan_tile_sums_kernel<<<c->scan_tiles, b200sgd::kScanBlock, 0, st>>>(c->d_cnt, n, c->d_tilesum);
    b200sgd::scan_tile_offsets_kernel<<<1, b200sgd::kScanBlock, 0, st>>>(c->d_tilesum, c->scan_tiles, c->d_tilesum + c->scan_tiles);
    b200sgd::scan_apply_kernel<<<c->scan_tiles, b200sgd::kScanBlock, 0, st>>>(c->d_cnt, n, c->d_tilesum, c->d_tilesum + c->scan_tiles,
                                                                          c->d_start);
    // the fill cursors start at the lists' offsets; afterwards the array is free and takes the successors
    CUDA_TRY(cudaMemcpyAsync(c->d_cnt, c->d_start, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, st));
    b200sgd::shuf_fill_kernel<<<blocks, 256, smem, st>>>(c->d_tgt, n, c->d_cnt, c->d_hits);
    b200sgd::shuf_order_kernel<<<blocks, 256, smem, st>>>(n, c->d_start, c->d_hits, c->d_first, c->d_cnt);
    b200sgd::shuf_resolve_kernel<<<blocks, 256, smem, st>>>(c->d_tgt, n, c->d_cnt, c->d_first, in, out);
    CUDA_TRY(cudaGetLastError());
    *launches += 7;
    return SGD_OK;
  }
  // backward formulation (shuffle2_kernels.cuh)
  constexpr int S = sgd_ctx::kSubChunks;
  const int64_t Lx = c->chunk_len / S, nx = (n - 1 + Lx - 1) / Lx;  // sub-chunks in use (<= nchunks * S)
  const int sgrid = c->stream_grid;            // virtual blocks of the stream kernels (columns of the count matrix)
  const int pgrid = std::min(sgrid, blocks);   // blocks launched
  unsigned int* err = c->d_bar + 1;
  // Confinement (smem != 0: B200SGD_SHUF_SMEM_KB, or B200SGD_SHUF_CONFINE=1 on one GPU): every kernel asks for
  // so much dynamic shared memory that static + dynamic = `smem` bytes -- more than an SM has left beside a CTA
  // of the step kernel -- so the block scheduler keeps the permutation work on the SMs the step grid leaves
  // idle (28 of 148 with 15 clusters of 8) and off the consumers' shared-memory pipe.
  auto dyn = [&](size_t static_bytes) -> size_t { return smem > static_bytes ? smem - static_bytes : 0; };
  const size_t dyn_stream = dyn(sizeof(int32_t) * b200sgd::kCoarseMax), dyn_plain = dyn(0);
  const size_t dyn_link = dyn(sizeof(int32_t) * (b200sgd::kLinkPos + b200sgd::kLinkThreads + b200sgd::kLinkCap + 8));
  const size_t dyn_refine = dyn(sizeof(int32_t) * (1024 + b200sgd::kLinkThreads)), dyn_place = dyn(sizeof(int32_t) << b200sgd::kRangeLog);
  b200sgd::shuf2_expand_states_kernel<<<(int)((c->nchunks * S + 127) / 128), 128, 0, st>>>(c->d_states, c->nchunks, c->d_subjump, S,
                                                                                     c->d_states_x);
  auto scan = [&](const int32_t* in, int64_t cells, int32_t* out) {  // out[0 .. cells]: exclusive scan + total
    const int tiles = (int)((cells + b200sgd::kScanTile - 1) / b200sgd::kScanTile);
    b200sgd::scan_tile_sums_kernel<<<tiles, b200sgd::kScanBlock, 0, st>>>(in, cells, c->d_tilesum);
    b200sgd::scan_tile_offsets_kernel<<<1, b200sgd::kScanBlock, 0, st>>>(c->d_tilesum, tiles, c->d_tilesum + tiles);
    b200sgd::scan_apply_kernel<<<tiles, b200sgd::kScanBlock, 0, st>>>(in, cells, c->d_tilesum, c->d_tilesum + tiles, out);
    *launches += 3;
  };
  const bool rscatter = c->shuf_ver == 3 || c->shuf_ver == 5;
  int32_t* R_stream = rscatter ? nullptr : c->d_R;  // self swaps: written by the stream kernel, or by the place kernel
  if (c->shuf_ver == 2) {
    CUDA_TRY(cudaMemsetAsync(c->d_ccnt, 0, sizeof(int32_t) * ((size_t)c->ngran + 1), st));
    b200sgd::shuf2_stream_kernel<0, false><<<pgrid, 128, dyn_plain, st>>>(c->d_states_x, Lx, nx, n, b200sgd::kGranLog, 0, sgrid, c->d_ccnt, nullptr, nullptr);
    scan(c->d_ccnt, c->ngran, c->d_gstart);
    // the counters become the scatter's cursors
    CUDA_TRY(cudaMemcpyAsync(c->d_ccnt, c->d_gstart, sizeof(int32_t) * ((size_t)c->ngran + 1), cudaMemcpyDeviceToDevice, st));
    b200sgd::shuf2_stream_kernel<1, false><<<pgrid, 128, dyn_plain, st>>>(c->d_states_x, Lx, nx, n, b200sgd::kGranLog, 0, sgrid, c->d_ccnt, c->d_pairs2, R_stream);
    *launches += 5;
  } else {
    const int32_t* coff = c->d_cstart;
    int64_t cstride = 1;
    if (c->shuf_ver == 3) {
      CUDA_TRY(cudaMemsetAsync(c->d_ccnt, 0, sizeof(int32_t) * ((size_t)c->ncoarse + 1), st));
      b200sgd::shuf2_stream_kernel<0, false><<<pgrid, 128, dyn_plain, st>>>(c->d_states_x, Lx, nx, n, c->cshift, 0, sgrid, c->d_ccnt, nullptr, nullptr);
      scan(c->d_ccnt, c->ncoarse, c->d_cstart);
      CUDA_TRY(cudaMemcpyAsync(c->d_ccnt, c->d_cstart, sizeof(int32_t) * ((size_t)c->ncoarse + 1), cudaMemcpyDeviceToDevice, st));
      b200sgd::shuf2_stream_kernel<1, false><<<pgrid, 128, dyn_plain, st>>>(c->d_states_x, Lx, nx, n, c->cshift, 0, sgrid, c->d_ccnt, c->d_pairs1, R_stream);
      *launches += 2;
    } else {
      const int64_t cells = c->ncoarse * (int64_t)sgrid;
      b200sgd::shuf2_stream_kernel<0, true><<<pgrid, 128, dyn_stream, st>>>(c->d_states_x, Lx, nx, n, c->cshift, (int)c->ncoarse, sgrid, c->d_mcnt, nullptr, nullptr);
      scan(c->d_mcnt, cells, c->d_moff);
      b200sgd::shuf2_stream_kernel<1, true><<<pgrid, 128, dyn_stream, st>>>(c->d_states_x, Lx, nx, n, c->cshift, (int)c->ncoarse, sgrid, c->d_moff, c->d_pairs1, R_stream);
      coff = c->d_moff;
      cstride = sgrid;
    }
    b200sgd::shuf2_refine_kernel<<<(int)std::min<int64_t>(c->ncoarse, blocks), b200sgd::kLinkThreads, dyn_refine, st>>>(
        c->ngran, c->ncoarse, c->cshift, coff, cstride, c->d_pairs1, c->d_pairs2, c->d_gstart);
    *launches += 4;
  }
  if (!rscatter) {
    b200sgd::shuf2_link_kernel<false><<<blocks, b200sgd::kLinkThreads, dyn_link, st>>>(n, c->ngran, c->d_gstart, c->d_pairs2, c->d_last, c->d_R,
                                                                               nullptr, nullptr, err);
    *launches += 1;
  } else {
    b200sgd::shuf2_range_cursor_kernel<<<(int)std::min<int64_t>((c->nranges + 255) / 256, blocks), 256, 0, st>>>(c->d_rcur, c->nranges,
                                                                                                         b200sgd::kRangeLog);
    // the coarse pairs are dead: their array takes the entries of R, by range
    b200sgd::shuf2_link_kernel<true><<<blocks, b200sgd::kLinkThreads, dyn_link, st>>>(n, c->ngran, c->d_gstart, c->d_pairs2, c->d_last, nullptr,
                                                                              c->d_rcur, c->d_pairs1, err);
    b200sgd::shuf2_place_kernel<<<(int)std::min<int64_t>(c->nranges, blocks), 256, dyn_place, st>>>(n, c->nranges, c->d_rcur, c->d_pairs1, c->d_R);
    *launches += 3;
  }
  b200sgd::shuf2_resolve_kernel<<<blocks, 256, dyn_plain, st>>>(n, c->d_last, c->d_R, in, out);
  CUDA_TRY(cudaGetLastError());
  *launches += 2;
  return SGD_OK;
}

// Device part: upload the chunk states (copy stream), then K1..K5 on the SHUFFLE stream turn
// d_perm[src] into d_perm[dst] -- concurrently with the epoch kernel that is still reading
// d_perm[src] on the compute stream.  No host sync.
int enqueue_device_shuffle(sgd_ctx* c, int stage_buf, int src, int dst, int* launches) {
  const int64_t n = c->R;
  c->bucketed[dst] = false;
  CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_done[2], 0));  // previous shuffle done with d_states
  CUDA_TRY(cudaMemcpyAsync(c->d_states, c->h_states[stage_buf], sizeof(uint32_t) * 31 * (size_t)c->nchunks,
                           cudaMemcpyHostToDevice, c->copy_stream));
  CUDA_TRY(cudaEventRecord(c->ev_perm[stage_buf], c->copy_stream));
  CUDA_TRY(cudaStreamWaitEvent(c->shuf_stream, c->ev_perm[stage_buf], 0));
  CUDA_TRY(cudaStreamWaitEvent(c->shuf_stream, c->ev_done[dst], 0));  // last epoch that read d_perm[dst]
  CUDA_TRY(cudaStreamWaitEvent(c->shuf_stream, c->ev_perm[src], 0));  // d_perm[src] uploaded (sgd_set_permutation)
  int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->num_sms * 16);
  size_t smem = shuf_smem_bytes();
  {  // B200SGD_SHUF_CONFINE=1|0: keep the permutation kernels off the step kernel's SMs (see enqueue_perm_kernels)
    const char* e = std::getenv("B200SGD_SHUF_CONFINE");
    const bool want = e ? e[0] == '1' : c->shuf_confine_default;
    const size_t sm_total = 233472, step_cta = c->smem_bytes + 1024;   // bytes of an SM; a step CTA with its reserve
    const size_t room = sm_total > step_cta ? sm_total - step_cta : 0;  // what a second block could get, its 1 KB reserve included
    const int idle_sms = c->num_sms - c->grid;
    if (want && smem == 0 && c->shuf_ver >= 2 && idle_sms >= 16 && room < ((size_t)46 << 10)) {
      smem = room;  // static + dynamic = room, plus the block's own 1 KB reserve: does not fit
      blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)idle_sms * 6);
    }
  }
  SGD_TRY(enqueue_perm_kernels(c, c->shuf_stream, blocks, smem, c->d_perm[src], c->d_perm[dst], launches));
  CUDA_TRY(cudaEventRecord(c->ev_done[2], c->shuf_stream));
  CUDA_TRY(cudaEventRecord(c->ev_shuf[dst], c->shuf_stream));
  return SGD_OK;
}

// ---- shared permutation work (nranks > 1, peers connected): see shuffle_kernels.cuh -----------------
constexpr int kFlagReady = 0;   // [owner rank][buffer]: number of the shuffle whose map stands in the owner's buffer
constexpr int kFlagDone = 16;   // [buffer][consumer rank] (in the OWNER's block): last shuffle the consumer applied from it

bool use_shared_shuffle(const sgd_ctx* c) { return c->shared_shuffle && c->peers_ready; }

// Dynamic shared memory the permutation kernels ASK for without using it (B200SGD_SHUF_SMEM_KB): a block
// that wants 40 KB does not fit on an SM beside a step CTA (199-227 KB), so the block scheduler keeps the
// permutation work on the SMs the step kernel leaves idle (28 of 148 with 15 clusters of 8, 36 with 7 of 16).
size_t shuf_smem_bytes() {
  static const size_t v = [] {
    const char* e = std::getenv("B200SGD_SHUF_SMEM_KB");
    return e ? (size_t)std::min(48, std::max(0, std::atoi(e))) * 1024 : (size_t)0;
  }();
  return v;
}

unsigned long long watchdog_ns(const sgd_ctx* c) {
  unsigned long long ns = c->cfg.nranks > 1 ? 30000000000ull : 4000000000ull;
  if (const char* e = std::getenv("B200SGD_WATCHDOG_MS")) ns = 1000000ull * (unsigned long long)std::max(1ll, std::atoll(e));
  return ns;
}

// Rank (k-1) % N: the position map of shuffle k -> own buffer ((k-1) / N) & 1, on own_stream.
int enqueue_own_map(sgd_ctx* c, int64_t k, int* launches) {
  const int N = c->cfg.nranks, me = c->cfg.rank;
  const int b = (int)(((k - 1) / N) & 1);
  const int64_t n = c->R;
  // the host's part: chunk start states of shuffle k's stream (rng_own stands at its start)
  CUDA_TRY(cudaEventSynchronize(c->ev_own_up[b]));  // staging buffer free again
  c->rng_own.chunk_states(c->jump_chunk, c->nchunks, c->h_states_own[b]);
  if (n > 1) c->rng_own.skip((uint64_t)N * (uint64_t)(n - 1));  // to my next shuffle
  cudaStream_t st = c->own_stream;
  CUDA_TRY(cudaStreamWaitEvent(st, c->ev_done[2], 0));  // a shuffle of the unshared kind may still use the scratch arrays
  uint32_t* flags = (uint32_t*)((char*)c->d_xchg_block + c->flag_off);
  if (k - 2 * (int64_t)N >= c->shared_first) {  // every rank has applied the map that stood in this buffer before (shuffle k - 2N)
    b200sgd::shuf_wait_flags_kernel<<<1, 32, 0, st>>>(flags, kFlagDone + b * 8, N, 1, (uint32_t)(k - 2 * N), c->d_bar + 1,
                                                    watchdog_ns(c));
    *launches += 1;
  }
  CUDA_TRY(cudaMemcpyAsync(c->d_states, c->h_states_own[b], sizeof(uint32_t) * 31 * (size_t)c->nchunks, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaEventRecord(c->ev_own_up[b], st));
  // At most 8 / N blocks per SM: a map is needed every N-th epoch only, and a GPU filled with long-running
  // map blocks keeps the next epoch's cooperative step kernel from starting until they drain (one
  // 256-thread block fits an SM beside a step CTA; more do not) -- measured at N = 8 as ~2.8 ms per epoch
  int bps = std::max(1, 8 / N);
  if (const char* e = std::getenv("B200SGD_MAP_BLOCKS_PER_SM")) bps = std::max(1, std::atoi(e));
  int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)c->num_sms * bps);
  // Confinement (B200SGD_MAP_CONFINE=1|0; default: on from 4 ranks with the backward kernels): keep the map off
  // the step kernel's SMs altogether -- its kernels ASK for 40 KB of shared memory (static + dynamic), so their
  // blocks do not fit beside a step CTA (199-227 KB) and run on the SMs the step grid leaves idle (36 of 148
  // with 7 clusters of 16), five per SM.  A map then takes ~5x as long (its link / refine / resolve kernels are
  // latency-bound and get 180 instead of 1184+ resident blocks: ~32 ms at 64M rows, from the one-GPU run
  // of B200SGD_SHUF_CONFINE), but a rank needs one map per N epochs only, and the step loop keeps its SMs to
  // itself.  History: with the forward kernels (13 ms alone, ~2x that confined on 28-36 SMs) the epoch waited
  // for the map: N = 2: 28.3 ms per epoch against 23.9 co-located, N = 8: 11.6 against 10.2, N = 4: 16.3
  // against 16.7.  With the backward kernels co-located maps cost the latency-bound step loop 15.4 instead
  // of 10.4 us per step at N = 4 and 9.2 instead of 7.9 at N = 8 (profiles/r3_scale_c5_n4.json, _n8.json),
  // while N x (epoch time) is 44 / 68 ms at N = 4 / 8: room for a confined map.  At N = 2 (2 x 15.8 ms) it
  // is not: co-located.  The default from 4 ranks follows from those numbers; it could not be re-measured
  // on the final build (GPU budget of the round spent) -- B200SGD_MAP_CONFINE=0 restores co-located maps.
  size_t map_smem = shuf_smem_bytes();
  const int idle_sms = c->num_sms - c->grid;
  const char* force = std::getenv("B200SGD_MAP_CONFINE");
  const bool confine = force ? force[0] == '1' : (N >= 4 && c->shuf_ver >= 2);
  if (idle_sms >= 16 && confine && map_smem == 0) {
    map_smem = (size_t)40 << 10;
    blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)idle_sms * 5);
  }
  int32_t* map = (int32_t*)((char*)c->d_xchg_block + c->src_off[b]);
  SGD_TRY(enqueue_perm_kernels(c, st, blocks, map_smem, nullptr, map, launches));
  b200sgd::PeerFlagPtrs t{};
  for (int r = 0; r < N; ++r) t.p[r] = (uint32_t*)((char*)c->peer_block[r] + c->flag_off);
  b200sgd::shuf_set_flags_kernel<<<1, 32, 0, st>>>(t, N, kFlagReady + me * 2 + b, (uint32_t)k);
  CUDA_TRY(cudaGetLastError());
  *launches += 1;
  return SGD_OK;
}

// Every rank: d_perm[dst] = d_perm[src] o (position map of shuffle k), read from its owner's buffer over
// NVLink once the owner has published it; then this rank's bucketed row lists.
int apply_shared_map(sgd_ctx* c, int64_t k, int src, int dst, int* launches) {
  const int N = c->cfg.nranks, me = c->cfg.rank;
  const int owner = (int)((k - 1) % N), b = (int)(((k - 1) / N) & 1);
  cudaStream_t st = c->shuf_stream;
  CUDA_TRY(cudaStreamWaitEvent(st, c->ev_done[dst], 0));  // last epoch that read d_perm[dst]
  CUDA_TRY(cudaStreamWaitEvent(st, c->ev_perm[src], 0));  // d_perm[src] uploaded (sgd_set_permutation)
  uint32_t* flags = (uint32_t*)((char*)c->d_xchg_block + c->flag_off);
  b200sgd::shuf_wait_flags_kernel<<<1, 32, 0, st>>>(flags, kFlagReady + owner * 2 + b, 1, 1, (uint32_t)k, c->d_bar + 1, watchdog_ns(c));
  const int64_t n = c->R;
  const int blocks = (int)std::min<int64_t>((n + 1023) / 1024, (int64_t)c->num_sms * 8);
  const int32_t* map = (const int32_t*)((const char*)c->peer_block[owner] + c->src_off[b]);
  b200sgd::shuf_apply_kernel<<<blocks, 256, 0, st>>>(map, c->d_perm[src], c->d_perm[dst], n);
  b200sgd::PeerFlagPtrs t{};
  t.p[0] = (uint32_t*)((char*)c->peer_block[owner] + c->flag_off);
  b200sgd::shuf_set_flags_kernel<<<1, 32, 0, st>>>(t, 1, kFlagDone + b * 8 + me, (uint32_t)k);
  CUDA_TRY(cudaGetLastError());
  // the schedule of the new permutation right behind it (d_counts: only this stream uses it in this mode)
  c->bucketed[dst] = false;
  SGD_TRY(bucket_perm(c, dst, launches, st));
  CUDA_TRY(cudaEventRecord(c->ev_shuf[dst], st));
  *launches += 3;
  return SGD_OK;
}

// Every rank: shuffle number k = shuf_no + 1 turns d_perm[src] into d_perm[dst] on the shuffle stream.
int enqueue_shared_shuffle(sgd_ctx* c, int src, int dst, int* launches) {
  const int N = c->cfg.nranks;
  const int64_t k = c->shuf_no + 1;
  if (c->shared_first == 0) {  // first shared shuffle: skip the maps of shuffles that were done the unshared way
    c->shared_first = k;
    while (c->own_next < k) {
      c->own_next += N;
      if (c->R > 1) c->rng_own.skip((uint64_t)N * (uint64_t)(c->R - 1));
    }
  }
  // Shuffles come in blocks of N (one map per rank).  My map of the NEXT block is started when the
  // current block begins -- on every rank at the same shuffle: computing a map slows the step loop of
  // the GPU it runs on by ~40 % (random 32-byte accesses against the row stream), and a step is as slow
  // as the slowest rank, so the ranks had better be slow together for one map's time and undisturbed
  // for the rest of the block than take turns at being the slow one all the time.
  static const bool stagger = std::getenv("B200SGD_MAP_STAGGER") != nullptr;  // A/B: every rank N - 1 shuffles ahead of its own use
  while (stagger ? c->own_next <= k + N - 1 : (c->own_next - 1) / N <= (k - 1) / N + 1) {
    SGD_TRY(enqueue_own_map(c, c->own_next, launches));
    c->own_next += N;
  }
  if (c->R > 1) c->rng.skip((uint64_t)(c->R - 1));  // the context's own generator stays in step (host-side fallbacks)
  SGD_TRY(apply_shared_map(c, k, src, dst, launches));
  c->shuf_no = k;
  return SGD_OK;
}

// Pinned staging of host-side permutations: only the host shuffle and sgd_set_permutation need it, and
// page-locking 8 bytes per row costs milliseconds, so it is allocated on first use.
int ensure_host_stage(sgd_ctx* c) {
  for (int i = 0; i < 2; ++i)
    if (!c->h_stage[i]) CUDA_TRY(cudaMallocHost(&c->h_stage[i], sizeof(int32_t) * (size_t)c->R));
  return SGD_OK;
}

// One device shuffle d_perm[src] -> d_perm[dst], enqueued (no host sync): the shared way with several
// connected ranks, else all of it on this GPU.
int enqueue_next_shuffle(sgd_ctx* c, int src, int dst, int* launches) {
  if (use_shared_shuffle(c)) return enqueue_shared_shuffle(c, src, dst, launches);
  CUDA_TRY(cudaEventSynchronize(c->ev_perm[dst]));  // staging buffer free again
  c->rng_saved = c->rng;
  prepare_stream_states(c, dst);
  SGD_TRY(enqueue_device_shuffle(c, dst, src, dst, launches));
  c->shuf_no += 1;
  return SGD_OK;
}

// A shuffle request: take the look-ahead if there is one, else enqueue the shuffle now.  Flips c->cur.
int take_or_enqueue_shuffle(sgd_ctx* c, int* launches) {
  const int buf = c->cur ^ 1;
  if (c->ahead) {
    c->ahead = false;  // d_perm[buf] is already being produced on the shuffle stream
  } else if (c->reapply_k) {
    SGD_TRY(apply_shared_map(c, c->reapply_k, c->cur, buf, launches));
    c->reapply_k = 0;
  } else {
    SGD_TRY(enqueue_next_shuffle(c, c->cur, buf, launches));
  }
  c->cur = buf;
  c->perm_valid = true;
  c->host_ind_stale = true;
  return SGD_OK;
}

// Enqueue the NEXT permutation into d_perm[cur ^ 1] without making it current.
int enqueue_lookahead(sgd_ctx* c, int* launches) {
  if (!c->device_shuffle || !c->lookahead || c->ahead || c->reapply_k) return SGD_OK;
  c->ahead_shared = use_shared_shuffle(c);
  SGD_TRY(enqueue_next_shuffle(c, c->cur, c->cur ^ 1, launches));
  c->ahead = true;
  return SGD_OK;
}

int sync_host_ind(sgd_ctx* c) {  // bring h_ind up to date after device shuffles
  if (!c->host_ind_stale) return SGD_OK;
  c->h_ind.resize((size_t)c->R);
  CUDA_TRY(cudaStreamSynchronize(c->shuf_stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpy(c->h_ind.data(), c->d_perm[c->cur], sizeof(int32_t) * (size_t)c->R, cudaMemcpyDeviceToHost));
  c->host_ind_stale = false;
  return SGD_OK;
}

// After a host sync: fold the running-loss sums of the epochs launched since the last collection
// into the per-epoch curve (mean |r| and 0.5 * mean r^2 over the rows this rank processed).
int collect_losses(sgd_ctx* c) {
  if (c->loss_pending == 0) return SGD_OK;
  std::vector<double> h((size_t)2 * c->grid);
  const int64_t rows_epoch = c->cfg.nranks > 1 ? -1 : (int64_t)c->nb * c->B;
  for (int k = c->loss_pending; k >= 1; --k) {
    const int slot = ((c->loss_next - k) % kLossSlots + kLossSlots) % kLossSlots;
    CUDA_TRY(cudaMemcpy(h.data(), c->d_loss + (size_t)slot * 2 * c->grid, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
    double a = 0.0, b = 0.0;
    for (int g = 0; g < c->grid; ++g) {
      a += h[2 * (size_t)g];
      b += h[2 * (size_t)g + 1];
    }
    c->curve_abs.push_back(a);
    c->curve_sq.push_back(b);
    c->curve_rows.push_back(rows_epoch);
  }
  c->loss_pending = 0;
  return SGD_OK;
}

int check_watchdog(sgd_ctx* c) {
  unsigned int err = 0;
  CUDA_TRY(cudaMemcpy(&err, c->d_bar + 1, sizeof(err), cudaMemcpyDeviceToHost));
  if (err != 0) {
    cudaMemset(c->d_bar + 1, 0, sizeof(unsigned int));
    return fail(c->cfg.nranks > 1 ? SGD_ERR_COMM : SGD_ERR_CUDA,
                err == 5   ? "shared shuffle timed out (a peer rank did not deliver / apply its permutation map)"
                : err == 6 ? "device shuffle: a granule of 256 positions collected more hits than the link kernel stages (B200SGD_SHUFFLE=1 selects the forward kernels)"
                : err == 3 ? "dataflow reduction timed out (a CTA or a peer rank did not publish)"
                : err == 4 ? "cluster exchange timed out (rows or partial sums did not arrive)"
                           : "grid barrier timed out (CTAs not co-resident?)");
  }
  return SGD_OK;
}

void fill_timings(const sgd_ctx* c, sgd_timings* t, float shuffle_ms, float loop_ms, int launches, int epochs) {
  if (!t) return;
  t->shuffle_ms = shuffle_ms;
  t->steploop_ms = loop_ms;
  t->epoch_ms = shuffle_ms + loop_ms;
  t->steps = c->nb * epochs;
  t->launches = launches;
  t->samples = (int64_t)c->nb * c->B * epochs;
  // DESIGN.md / SURVEY 8d: per step 4*B*C (features, read once) + 4*B (labels) + 4*B (indices)
  // + 4*C (w read) + 4*C (w write); summed over ranks this is what one rank sees / nranks
  const int64_t rows = (int64_t)c->nb * c->B / std::max(1, c->cfg.nranks);
  t->bytes = ((int64_t)4 * rows * c->C + 8 * rows + (int64_t)c->nb * 8 * c->C) * epochs;
}

}  // namespace

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

int sgd_abi_version(void) { return B200SGD_ABI_VERSION; }

const char* sgd_last_error(void) { return g_last_error.c_str(); }

int sgd_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, i) == cudaSuccess && pr.major == 10) ++ok;
  }
  return ok;
}

int sgd_read_csv(const char* path, int64_t rows, int32_t cols, float* X, float* y) {
  if (!path || !X || !y || rows < 0 || cols < 0) return fail(SGD_ERR_ARG, "sgd_read_csv: bad argument");
  std::string err;
  const int rc = b200sgd::read_csv(path, rows, cols, X, y, &err);
  if (rc == -2) return fail(SGD_ERR_IO, err);
  if (rc == -3) return fail(SGD_ERR_PARSE, err);
  return SGD_OK;
}

int sgd_write_csv(const char* path, int64_t rows, int32_t cols, const float* X, const float* y) {
  if (!path || !X || !y || rows < 0 || cols < 0) return fail(SGD_ERR_ARG, "sgd_write_csv: bad argument");
  if (b200sgd::write_csv(path, rows, cols, X, y) != 0) return fail(SGD_ERR_IO, std::string("cannot write ") + path);
  return SGD_OK;
}

int sgd_get_filename_from_path(const char* path, char* out, size_t out_cap) {
  if (!path || !out || out_cap == 0) return fail(SGD_ERR_ARG, "sgd_get_filename_from_path: bad argument");
  const std::string s = b200sgd::filename_from_path(path);
  if (s.size() + 1 > out_cap) return fail(SGD_ERR_ARG, "output buffer too small");
  std::memcpy(out, s.c_str(), s.size() + 1);
  return SGD_OK;
}

int sgd_write_experiment_output(const char* path, float shuffle_time, float total_gradient_time,
                                float gpu_time, float transfer_time, float error) {
  if (!path) return fail(SGD_ERR_ARG, "null path");
  if (b200sgd::write_experiment_json(path, shuffle_time, total_gradient_time, gpu_time, transfer_time, error) != 0)
    return fail(SGD_ERR_IO, std::string("cannot write ") + path);
  return SGD_OK;
}

int sgd_init_weights(float* w, int32_t cols) {
  if (!w || cols < 0) return fail(SGD_ERR_ARG, "sgd_init_weights: bad argument");
  b200sgd::thrust_minstd_weights(w, cols);
  return SGD_OK;
}

int sgd_batch_geometry(int64_t rows, int32_t batch_arg, int32_t* batch_out, int32_t* steps_out) {
  if (rows <= 0 || rows > 2147483647ll || batch_arg < 0) return fail(SGD_ERR_ARG, "bad rows / batch");
  int32_t B, nb;
  batch_geometry(rows, batch_arg, &B, &nb);
  if (batch_out) *batch_out = B;
  if (steps_out) *steps_out = nb;
  return SGD_OK;
}

int sgd_host_shuffle(int32_t* ind, int64_t rows, uint32_t seed, int32_t skip_epochs, int32_t epochs) {
  if (!ind || rows < 0 || skip_epochs < 0 || epochs < 0) return fail(SGD_ERR_ARG, "sgd_host_shuffle: bad argument");
  b200sgd::GlibcRand g(seed);
  // every pass consumes rows-1 values of the stream (stl_algo.h random_shuffle)
  if (skip_epochs > 0 && rows > 1) {
    std::vector<uint32_t> sink(1 << 16);
    uint64_t left = (uint64_t)skip_epochs * (uint64_t)(rows - 1);
    while (left) {
      const size_t n = (size_t)std::min<uint64_t>(left, sink.size());
      g.raw(sink.data(), n);
      left -= n;
    }
  }
  for (int e = 0; e < epochs; ++e) b200sgd::exact_shuffle(g, ind, rows);
  return SGD_OK;
}

int sgd_shard_range(int64_t rows, int32_t rank, int32_t nranks, int64_t* begin, int64_t* end) {
  if (rows <= 0 || nranks < 1 || rank < 0 || rank >= nranks) return fail(SGD_ERR_ARG, "bad shard arguments");
  const int64_t per = (rows + nranks - 1) / nranks;
  const int64_t b = std::min<int64_t>(rows, per * rank);
  if (begin) *begin = b;
  if (end) *end = std::min<int64_t>(rows, b + per);
  return SGD_OK;
}

int sgd_host_bucket(const int32_t* perm, int64_t rows, int32_t batch_arg, int32_t rank, int32_t nranks,
                    int32_t* local_perm, int64_t local_cap, int64_t* step_off) {
  if (!perm || !local_perm || !step_off) return fail(SGD_ERR_ARG, "null argument");
  int64_t rb = 0, re = 0;
  SGD_TRY(sgd_shard_range(rows, rank, nranks, &rb, &re));
  int32_t B, nb;
  batch_geometry(rows, batch_arg, &B, &nb);
  int64_t n = 0;
  for (int32_t s = 0; s < nb; ++s) {
    step_off[s] = n;
    const int32_t* p = perm + (int64_t)s * B;
    for (int32_t i = 0; i < B; ++i) {
      const int64_t r = p[i];
      if (r >= rb && r < re) {
        if (n >= local_cap) return fail(SGD_ERR_ARG, "local_perm too small");
        local_perm[n++] = (int32_t)(r - rb);
      }
    }
  }
  step_off[nb] = n;
  return SGD_OK;
}

int sgd_create(const sgd_config* cfg, const float* X, const float* y, sgd_ctx** out) {
  if (!cfg || !out) return fail(SGD_ERR_ARG, "sgd_create: null argument");
  *out = nullptr;
  if (cfg->rows <= 0 || cfg->rows > 2147483647ll) return fail(SGD_ERR_ARG, "rows must be in [1, 2^31-1]");
  if (cfg->cols <= 0) return fail(SGD_ERR_ARG, "cols must be positive");
  if (cfg->batch < 0) return fail(SGD_ERR_ARG, "batch must not be negative");  // batch > rows: 0 steps, as main.cu:344
  if (cfg->nranks < 1 || cfg->nranks > 8 || cfg->rank < 0 || cfg->rank >= cfg->nranks)
    return fail(SGD_ERR_ARG, "need 1 <= nranks <= 8 and 0 <= rank < nranks");
  if ((X == nullptr) != (y == nullptr)) return fail(SGD_ERR_ARG, "X and y must both be given or both be NULL");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(SGD_ERR_CUDA, "no CUDA device: libb200sgd has no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(SGD_ERR_ARG, "device ordinal out of range");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(SGD_ERR_CUDA, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                  "; this library contains sm_100a code only");

  sgd_ctx* c = new (std::nothrow) sgd_ctx();
  if (!c) return fail(SGD_ERR_NOMEM, "out of host memory");
  c->cfg = *cfg;
  if (c->cfg.shuffle_seed == 0) c->cfg.shuffle_seed = 1;
  c->R = cfg->rows;
  c->C = cfg->cols;
  batch_geometry(c->R, cfg->batch, &c->B, &c->nb);
  c->ld = ((int64_t)c->C + 1 + 15) / 16 * 16;  // 64-byte records: whole HBM bursts, see DESIGN.md
  if (const char* e = std::getenv("B200SGD_LD_ALIGN")) {  // experiment: record pitch in floats (a multiple of 16)
    const int64_t al = std::max<int64_t>(16, std::atoll(e) / 16 * 16);
    c->ld = ((int64_t)c->C + 1 + al - 1) / al * al;
  }
  c->C4 = (c->C + 3) / 4;
  c->rows_per_rank = (c->R + cfg->nranks - 1) / cfg->nranks;
  c->row_begin = std::min<int64_t>(c->R, c->rows_per_rank * cfg->rank);
  c->row_end = std::min<int64_t>(c->R, c->row_begin + c->rows_per_rank);
  c->local_rows = c->row_end - c->row_begin;
  c->device = cfg->device;
  c->num_sms = prop.multiProcessorCount;
  c->rng.reseed(c->cfg.shuffle_seed);

  auto bail = [&](int rc) {
    const std::string keep = g_last_error;
    free_all(c);
    delete c;
    g_last_error = keep;
    return rc;
  };
  const bool dbg_time = std::getenv("B200SGD_DEBUG") != nullptr;
  const auto t_setup = clk::now();
  auto stamp = [&](const char* what) {
    if (dbg_time) std::fprintf(stderr, "[b200sgd] sgd_create %-28s %8.2f ms\n", what, ms_since(t_setup));
  };
  auto setup = [&]() -> int {
    SGD_TRY(bind(c));
    int prio_lo = 0, prio_hi = 0;  // the step kernel's CTAs go first when an SM frees up
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->shuf_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_shuf[0], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_shuf[1], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    for (int i = 0; i < 2; ++i) {
      CUDA_TRY(cudaEventCreateWithFlags(&c->ev_perm[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_done[2], cudaEventDisableTiming));
    {  // watchdog of the in-kernel waits (sgd_kernels.cuh, SpinGuard)
      unsigned long long ns = c->cfg.nranks > 1 ? 30000000000ull : 4000000000ull;
      if (const char* e = std::getenv("B200SGD_WATCHDOG_MS")) ns = 1000000ull * (unsigned long long)std::max(1ll, std::atoll(e));
      CUDA_TRY(cudaMemcpyToSymbol(b200sgd::c_spin_timeout_ns, &ns, sizeof(ns)));
    }
    // B200SGD_L2_FETCH=32|64|128: granularity in which the L2 fetches from DRAM (device-wide limit).  The
    // permutation kernels and the row gather of narrow records are random accesses of 4 - 448 bytes; ncu
    // shows the shuffle's resolve pass moving ~117 bytes of DRAM traffic per 4-byte random read.
    if (const char* e = std::getenv("B200SGD_L2_FETCH")) {
      const int g = std::atoi(e);
      if (g == 32 || g == 64 || g == 128) {
        if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g) != cudaSuccess) cudaGetLastError();
      }
    }
    stamp("streams, events");
    SGD_TRY(configure_launch(c));
    stamp("configure_launch");
    const size_t rec_bytes = sizeof(float) * (size_t)std::max<int64_t>(1, c->local_rows) * (size_t)c->ld;
    if (cudaMalloc(&c->d_rec, rec_bytes) != cudaSuccess) {
      cudaGetLastError();
      return fail(SGD_ERR_NOMEM, "cannot allocate " + std::to_string(rec_bytes >> 20) + " MiB of HBM for the data");
    }
    stamp("record array");
    SGD_TRY(make_record_tensor_map(c));
    CUDA_TRY(cudaMalloc(&c->d_w, sizeof(float) * c->C));
    CUDA_TRY(cudaMalloc(&c->d_wtrue, sizeof(float) * c->C));
    for (int i = 0; i < 2; ++i) {
      CUDA_TRY(cudaMalloc(&c->d_perm[i], sizeof(int32_t) * (size_t)c->R));
    }
    if (c->cfg.nranks > 1) {
      for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaMalloc(&c->d_perm_local[i], sizeof(int32_t) * (size_t)std::max<int64_t>(1, c->local_rows)));
        CUDA_TRY(cudaMalloc(&c->d_step_off[i], sizeof(int64_t) * ((size_t)c->nb + 1)));
      }
      CUDA_TRY(cudaMalloc(&c->d_counts, sizeof(int64_t) * ((size_t)c->nb + 1)));
    }
    CUDA_TRY(cudaMalloc(&c->d_partials, sizeof(float4) * 2 * (size_t)c->grid * c->C4));
    CUDA_TRY(cudaMemset(c->d_partials, 0, sizeof(float4) * 2 * (size_t)c->grid * c->C4));
    CUDA_TRY(cudaMalloc(&c->d_wstage, sizeof(float4) * c->C4));
    const size_t Gp = ((size_t)c->grid + 3) & ~(size_t)3, C4p = ((size_t)c->C4 + 3) & ~(size_t)3;
    CUDA_TRY(cudaMalloc(&c->d_ll_part, sizeof(uint4) * 2 * 2 * Gp * c->C4));
    CUDA_TRY(cudaMemset(c->d_ll_part, 0, sizeof(uint4) * 2 * 2 * Gp * c->C4));
    // private weight mailboxes per CTA while they stay small (latency-bound shapes); one shared
    // copy for wide rows, where steps are long and the extra L2 traffic would not pay
    c->w_copies = ((size_t)c->grid * c->C4 * 32 <= (8u << 20)) ? c->grid : 1;
    CUDA_TRY(cudaMalloc(&c->d_ll_w, sizeof(uint4) * 2 * (size_t)c->w_copies * C4p));
    CUDA_TRY(cudaMemset(c->d_ll_w, 0, sizeof(uint4) * 2 * (size_t)c->w_copies * C4p));
    CUDA_TRY(cudaMalloc(&c->d_glast, sizeof(float4) * c->C4));
    CUDA_TRY(cudaMemset(c->d_glast, 0, sizeof(float4) * c->C4));
    CUDA_TRY(cudaMalloc(&c->d_bar, sizeof(unsigned int) * 2));
    CUDA_TRY(cudaMemset(c->d_bar, 0, sizeof(unsigned int) * 2));
    if (const char* e = std::getenv("B200SGD_PHASE_PROFILE"); e && e[0] == '1') {
      CUDA_TRY(cudaMalloc(&c->d_dbg, sizeof(unsigned long long) * 16 * (size_t)c->grid));
      CUDA_TRY(cudaMemset(c->d_dbg, 0, sizeof(unsigned long long) * 16 * (size_t)c->grid));
    }
    CUDA_TRY(cudaMalloc(&c->d_loss, sizeof(double) * 2 * (size_t)c->grid * kLossSlots));
    CUDA_TRY(cudaMalloc(&c->d_ctl, sizeof(b200sgd::StepCtl)));
    CUDA_TRY(cudaMemset(c->d_ctl, 0, sizeof(b200sgd::StepCtl)));
    if (c->cfg.flags & SGD_FLAG_RESIDUALS) {
      c->rout_len = c->cfg.nranks > 1 ? c->local_rows : (int64_t)c->nb * c->B;
      CUDA_TRY(cudaMalloc(&c->d_rout, sizeof(float) * (size_t)std::max<int64_t>(1, c->rout_len)));
      CUDA_TRY(cudaMemset(c->d_rout, 0, sizeof(float) * (size_t)std::max<int64_t>(1, c->rout_len)));
    }
    stamp("step buffers");
    // bit-exact permutation on the device for anything but tiny data sets
    c->device_shuffle = !(c->cfg.flags & SGD_FLAG_HOST_SHUFFLE) && c->R >= 32768;
    if (c->device_shuffle) {
      const size_t n = (size_t)c->R;
      c->scan_tiles = (int32_t)((n + b200sgd::kScanTile - 1) / b200sgd::kScanTile);
      if (const char* e = std::getenv("B200SGD_SHUFFLE")) c->shuf_ver = std::min(5, std::max(1, std::atoi(e)));
      // the refine kernel bins at most 1024 granules per coarse bucket: beyond 2^30 rows scatter by granule directly
      if (c->shuf_ver >= 3 && c->R > (1ll << 30)) c->shuf_ver = 2;
      CUDA_TRY(cudaMalloc(&c->d_tilesum, sizeof(int32_t) * ((size_t)c->scan_tiles + 1)));
      if (c->shuf_ver == 1) {
        CUDA_TRY(cudaMalloc(&c->d_tgt, sizeof(uint32_t) * n));
        CUDA_TRY(cudaMalloc(&c->d_cnt, sizeof(int32_t) * (n + 1)));
        CUDA_TRY(cudaMalloc(&c->d_start, sizeof(int32_t) * (n + 1)));
        CUDA_TRY(cudaMalloc(&c->d_hits, sizeof(int32_t) * n));
        CUDA_TRY(cudaMalloc(&c->d_first, sizeof(int32_t) * n));
      } else {
        c->ngran = (int64_t)((n + b200sgd::kGran - 1) >> b200sgd::kGranLog);
        // coarse buckets of >= 2^14 positions, at most 2048 of them (measured at 64M rows: 2^14 / 2^15 / 2^16
        // positions -> 28.7 / 28.5 / 28.4 ms per epoch of config 5 with the shuffle beside the step kernel; fewer,
        // larger buckets mean fewer open write fronts for the scatter but more work per refine block)
        c->cshift = 14;
        while (c->cshift < 18 && (((int64_t)n + (1ll << c->cshift) - 1) >> c->cshift) > 2048) ++c->cshift;
        if (const char* e = std::getenv("B200SGD_COARSE_SHIFT")) c->cshift = std::min(18, std::max(14, std::atoi(e)));
        while (c->cshift < 18 && (((int64_t)n + (1ll << c->cshift) - 1) >> c->cshift) > b200sgd::kCoarseMax) ++c->cshift;
        c->ncoarse = ((int64_t)n + (1ll << c->cshift) - 1) >> c->cshift;
        c->nranges = ((int64_t)n + (1ll << b200sgd::kRangeLog) - 1) >> b200sgd::kRangeLog;
        CUDA_TRY(cudaMalloc(&c->d_pairs2, sizeof(uint2) * n));
        CUDA_TRY(cudaMalloc(&c->d_last, sizeof(int32_t) * n));
        CUDA_TRY(cudaMalloc(&c->d_R, sizeof(int32_t) * n));
        CUDA_TRY(cudaMalloc(&c->d_gstart, sizeof(int32_t) * ((size_t)c->ngran + 1)));
        // version 2 counts per granule, version 3 per coarse bucket: one counter array serves both
        CUDA_TRY(cudaMalloc(&c->d_ccnt, sizeof(int32_t) * ((size_t)c->ngran + 1)));
        CUDA_TRY(cudaMalloc(&c->d_cstart, sizeof(int32_t) * ((size_t)c->ncoarse + 1)));
        if (c->shuf_ver >= 3) {
          CUDA_TRY(cudaMalloc(&c->d_pairs1, sizeof(uint2) * ((size_t)c->nranges << b200sgd::kRangeLog)));
          CUDA_TRY(cudaMalloc(&c->d_rcur, sizeof(int32_t) * (size_t)c->nranges));
        }
      }
      // the epoch's rand() stream is regenerated on the device from up to 8192 chunk start states
      // (one thread per chunk: at 64M rows 1024 chunks took 10.4 ms, 8192 take ~1.3 ms; the host
      // pays ~0.6 us per state, hidden behind the running epoch)
      // (backward kernels: 1024 host states at most, 64 sub-chunk states each derived on the device --
      // the host's share of a shuffle drops from ~5 ms to ~0.6 ms at 64M rows)
      int64_t L = 1024;
      while (L * (c->shuf_ver == 1 ? 8192 : 1024) < c->R) L *= 2;
      c->chunk_len = L;
      c->nchunks = (c->R - 1 + L - 1) / L;
      c->jump_chunk = b200sgd::GlibcRand::poly_z_pow((uint64_t)L);
      CUDA_TRY(cudaMalloc(&c->d_states, sizeof(uint32_t) * 31 * (size_t)c->nchunks));
      for (int i = 0; i < 2; ++i) CUDA_TRY(cudaMallocHost(&c->h_states[i], sizeof(uint32_t) * 31 * (size_t)c->nchunks));
      if (c->shuf_ver != 1) {  // sub-chunk states are derived on the device (shuf2_expand_states_kernel)
        constexpr int S = sgd_ctx::kSubChunks;
        CUDA_TRY(cudaMalloc(&c->d_states_x, sizeof(uint32_t) * 31 * (size_t)c->nchunks * S));
        std::vector<uint32_t> sj((size_t)31 * (S - 1));
        for (int k = 1; k < S; ++k) {
          const b200sgd::GlibcRand::Poly pz = b200sgd::GlibcRand::poly_z_pow((uint64_t)(L / S) * (uint64_t)k);
          std::memcpy(sj.data() + (size_t)31 * (k - 1), pz.a, sizeof(pz.a));
        }
        CUDA_TRY(cudaMalloc(&c->d_subjump, sizeof(uint32_t) * sj.size()));
        CUDA_TRY(cudaMemcpy(c->d_subjump, sj.data(), sizeof(uint32_t) * sj.size(), cudaMemcpyHostToDevice));
        const int64_t Lx = L / S, nx = (c->R - 1 + Lx - 1) / Lx;
        c->stream_grid = (int)std::max<int64_t>(1, (nx + 127) / 128);
        if (c->shuf_ver >= 4) {
          const size_t cells = (size_t)c->ncoarse * (size_t)c->stream_grid + 1;
          CUDA_TRY(cudaMalloc(&c->d_mcnt, sizeof(int32_t) * cells));
          CUDA_TRY(cudaMalloc(&c->d_moff, sizeof(int32_t) * cells));
          const size_t tiles = (cells + b200sgd::kScanTile - 1) / b200sgd::kScanTile;
          if (tiles > (size_t)c->scan_tiles) {  // small data sets: the matrix has more cells than rows
            cudaFree(c->d_tilesum);
            c->d_tilesum = nullptr;
            CUDA_TRY(cudaMalloc(&c->d_tilesum, sizeof(int32_t) * (tiles + 1)));
          }
        }
      }
    }
    stamp("shuffle buffers");
    c->mae_blocks = c->num_sms * 8;
    if (sizeof(float) * (size_t)c->ld > (48u << 10))  // weights in shared memory: opt in above 48 KB
      CUDA_TRY(cudaFuncSetAttribute((const void*)b200sgd::sgd_mae_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(sizeof(float) * (size_t)c->ld)));
    CUDA_TRY(cudaMalloc(&c->d_mae, sizeof(double) * c->mae_blocks));
    if (c->cfg.nranks > 1) {
      c->xchg_bytes = sizeof(uint4) * 2 * 2 * (size_t)c->cfg.nranks * (((size_t)c->C4 + 3) & ~(size_t)3);
      // direct push of the partials when the NVLink traffic per step stays small (narrow rows)
      const size_t push_bytes = (size_t)(c->cfg.nranks - 1) * c->grid * c->C4 * 32;
      c->direct = !c->clt && (c->grid % 4 == 0) && push_bytes <= (1u << 20) && !std::getenv("B200SGD_NO_DIRECT");
      c->xpart_off = (c->xchg_bytes + 255) / 256 * 256;
      if (c->direct) c->xchg_bytes = c->xpart_off + sizeof(uint4) * 2 * 2 * (size_t)c->C4 * c->cfg.nranks * c->grid;
      // cluster tail: tagged slice sums of every cluster of every rank, [2][nranks * clusters][C4p] float4
      if (c->clt) c->xchg_bytes = sizeof(uint4) * 2 * 2 * (size_t)c->cfg.nranks * c->clt_ng * (((size_t)c->C4 + 3) & ~(size_t)3);
      if (c->device_shuffle && !std::getenv("B200SGD_NO_SHARED_SHUFFLE")) {
        // shared permutation work: sequence-number flags + two position-map buffers every peer can read
        c->shared_shuffle = true;
        c->flag_off = (c->xchg_bytes + 4095) / 4096 * 4096;
        c->src_off[0] = c->flag_off + 4096;
        c->src_off[1] = c->src_off[0] + (sizeof(int32_t) * (size_t)c->R + 4095) / 4096 * 4096;
        c->xchg_bytes = c->src_off[1] + sizeof(int32_t) * (size_t)c->R;
        int plo = 0, phi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&plo, &phi));
        CUDA_TRY(cudaStreamCreateWithPriority(&c->own_stream, cudaStreamNonBlocking, plo));
        for (int i = 0; i < 2; ++i) {
          CUDA_TRY(cudaEventCreateWithFlags(&c->ev_own_up[i], cudaEventDisableTiming));
          CUDA_TRY(cudaMallocHost(&c->h_states_own[i], sizeof(uint32_t) * 31 * (size_t)c->nchunks));
        }
        c->own_next = c->cfg.rank + 1;
        c->rng_own = c->rng;
        if (c->cfg.rank > 0 && c->R > 1) c->rng_own.skip((uint64_t)c->cfg.rank * (uint64_t)(c->R - 1));
      }
      CUDA_TRY(cudaMalloc(&c->d_xchg_block, c->xchg_bytes));
      CUDA_TRY(cudaMemset(c->d_xchg_block, 0, std::min<size_t>(c->xchg_bytes, c->shared_shuffle ? c->src_off[0] : c->xchg_bytes)));
      c->peer_block[c->cfg.rank] = c->d_xchg_block;
    }
    if (c->clt && c->cfg.nranks == 1 && c->clt_ng > 1) {
      const size_t bytes = sizeof(uint4) * 2 * 2 * (size_t)c->clt_ng * (((size_t)c->C4 + 3) & ~(size_t)3);
      CUDA_TRY(cudaMalloc(&c->d_clt_inbox, bytes));
      CUDA_TRY(cudaMemset(c->d_clt_inbox, 0, bytes));
    }
    SGD_TRY(configure_cluster(c));
    stamp("cluster configuration");
    // main.cu:389-395: weights ~ U(0, 0.01) from the Thrust minstd stream, drawn on the host
    std::vector<float> w0(c->C);
    b200sgd::thrust_minstd_weights(w0.data(), c->C);
    CUDA_TRY(cudaMemcpy(c->d_w, w0.data(), sizeof(float) * c->C, cudaMemcpyHostToDevice));
    // main.cu:421-426: identity index vector; the host copy is materialised when somebody needs it
    b200sgd::shuf_iota_kernel<<<std::max(1, std::min<int>(c->num_sms * 8, (int)((c->R + 255) / 256))), 256, 0, c->stream>>>(
        c->d_perm[0], c->R);
    CUDA_TRY(cudaGetLastError());
    c->host_ind_stale = true;
    CUDA_TRY(cudaDeviceSynchronize());
    stamp("weights, identity, sync");
    // the first permutation does not depend on the data: it is produced beside the upload of the rows
    c->lookahead = std::getenv("B200SGD_NO_LOOKAHEAD") == nullptr;
    int la = 0;
    SGD_TRY(enqueue_lookahead(c, &la));
    return SGD_OK;
  };
  int rc = setup();
  if (rc != SGD_OK) return bail(rc);
  if (X) {
    rc = sgd_load_rows(c, 0, c->R, X, y);
    if (rc != SGD_OK) return bail(rc);
  }
  *out = c;
  return SGD_OK;
}

void sgd_destroy(sgd_ctx* ctx) {
  if (!ctx) return;
  free_all(ctx);
  delete ctx;
}

int sgd_load_rows(sgd_ctx* c, int64_t row_begin, int64_t nrows, const float* X, const float* y) {
  SGD_TRY(check_ctx(c));
  if (!X || !y || row_begin < 0 || nrows < 0 || row_begin + nrows > c->R)
    return fail(SGD_ERR_ARG, "sgd_load_rows: bad range");
  SGD_TRY(bind(c));
  // intersect with this rank's shard
  const int64_t lo = std::max(row_begin, c->row_begin), hi = std::min(row_begin + nrows, c->row_end);
  if (hi > lo) {
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(hi - lo, (64ll << 20) / (4ll * (c->C + 1))));
    float *dX = nullptr, *dy = nullptr;
    CUDA_TRY(cudaMalloc(&dX, sizeof(float) * (size_t)chunk * c->C));
    if (cudaMalloc(&dy, sizeof(float) * (size_t)chunk) != cudaSuccess) {
      cudaFree(dX);
      return fail(SGD_ERR_NOMEM, "staging allocation failed");
    }
    int rc = SGD_OK;
    for (int64_t r0 = lo; r0 < hi && rc == SGD_OK; r0 += chunk) {
      const int64_t n = std::min(chunk, hi - r0);
      const float* hx = X + (r0 - row_begin) * c->C;
      const float* hy = y + (r0 - row_begin);
      cudaError_t e = cudaMemcpyAsync(dX, hx, sizeof(float) * (size_t)n * c->C, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(dy, hy, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess) {
        const int blocks = (int)std::min<int64_t>((n * c->ld + 255) / 256, (int64_t)c->num_sms * 16);
        b200sgd::sgd_pack_kernel<<<blocks, 256, 0, c->stream>>>(dX, dy, n, c->C, c->ld,
                                                                c->d_rec + (r0 - c->row_begin) * c->ld);
        e = cudaGetLastError();
      }
      if (e != cudaSuccess) rc = fail(SGD_ERR_CUDA, std::string("upload failed: ") + cudaGetErrorString(e));
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(dX);
    cudaFree(dy);
    if (rc != SGD_OK) return rc;
    if (e != cudaSuccess) return fail(SGD_ERR_CUDA, std::string("upload failed: ") + cudaGetErrorString(e));
  }
  c->data_loaded = true;
  return SGD_OK;
}

// ---- CSV -> HBM with the parse and the upload overlapped, optional raw cache -----------------------
namespace {
struct RawCacheHeader {  // <csv>.f32: header, then rows * (cols + 1) floats (`x..., label` per row)
  char magic[8];
  uint32_t version, cols;
  int64_t rows, csv_size, csv_mtime_ns;
  char pad[24];
};
static_assert(sizeof(RawCacheHeader) == 64, "raw cache header layout");
const char kRawMagic[8] = {'B', '2', 'S', 'G', 'D', 'F', '3', '2'};
}  // namespace

int sgd_load_csv(sgd_ctx* c, const char* path, int32_t flags, sgd_load_stats* stats) {
  SGD_TRY(check_ctx(c));
  if (!path) return fail(SGD_ERR_ARG, "sgd_load_csv: null path");
  SGD_TRY(bind(c));
  const auto t_all = clk::now();
  const int64_t R = c->R, W = (int64_t)c->C + 1;
  sgd_load_stats st{};
  // pieces of ~4 MB of text; a pipeline stage is as many pieces as fit the staging buffer (~96 MB of
  // floats), parsed by all host threads (each takes the next unparsed piece)
  b200sgd::CsvFile csv;
  std::string err;
  if (csv.open(path, (size_t)4 << 20, &err) != 0) return fail(SGD_ERR_IO, err);
  const std::string cache_path = std::string(path) + ".f32";
  bool from_cache = false;
  FILE* fc = nullptr;
  if (flags & SGD_CSV_CACHE) {
    fc = std::fopen(cache_path.c_str(), "rb");
    RawCacheHeader h{};
    if (fc && std::fread(&h, sizeof(h), 1, fc) == 1 && std::memcmp(h.magic, kRawMagic, 8) == 0 && h.version == 1 &&
        h.cols == (uint32_t)c->C && h.rows == R && h.csv_size == csv.size() && h.csv_mtime_ns == csv.mtime_ns()) {
      from_cache = true;
    } else if (fc) {
      std::fclose(fc);
      fc = nullptr;
    }
  }
  // staging: two pinned buffers of `stage_rows` records each
  const int64_t stage_rows = std::max<int64_t>(1, std::min<int64_t>(R, ((int64_t)96 << 20) / (4 * W)));
  float* h_stage[2] = {nullptr, nullptr};
  float* d_stage[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int rc = SGD_OK;
  auto cleanup = [&] {
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 2; ++i) {
      if (h_stage[i]) cudaFreeHost(h_stage[i]);
      if (d_stage[i]) cudaFree(d_stage[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
    }
    if (fc) std::fclose(fc);
  };
  for (int i = 0; i < 2 && rc == SGD_OK; ++i) {
    if (cudaMallocHost(&h_stage[i], sizeof(float) * (size_t)(stage_rows * W)) != cudaSuccess ||
        cudaMalloc(&d_stage[i], sizeof(float) * (size_t)(stage_rows * W)) != cudaSuccess ||
        cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      rc = fail(SGD_ERR_NOMEM, "sgd_load_csv: staging allocation failed");
    }
  }
  if (rc != SGD_OK) { cleanup(); return rc; }
  // upload of stage buffer b holding file rows [r0, r0 + n): the part inside this rank's shard
  auto upload = [&](int b, int64_t r0, int64_t n) -> int {
    const int64_t lo = std::max(r0, c->row_begin), hi = std::min(r0 + n, c->row_end);
    if (hi > lo) {
      const float* hs = h_stage[b] + (lo - r0) * W;
      CUDA_TRY(cudaMemcpyAsync(d_stage[b], hs, sizeof(float) * (size_t)((hi - lo) * W), cudaMemcpyHostToDevice, c->stream));
      const int blocks = (int)std::min<int64_t>(((hi - lo) * c->ld + 255) / 256, (int64_t)c->num_sms * 16);
      b200sgd::sgd_pack_rec_kernel<<<blocks, 256, 0, c->stream>>>(d_stage[b], hi - lo, c->C, c->ld,
                                                                 c->d_rec + (lo - c->row_begin) * c->ld);
      CUDA_TRY(cudaGetLastError());
      st.bytes_uploaded += (hi - lo) * W * 4;
    }
    CUDA_TRY(cudaEventRecord(ev[b], c->stream));
    return SGD_OK;
  };
  FILE* fw = nullptr;  // cache being written (only when the whole file passes through this rank)
  std::string tmp_path;
  if ((flags & SGD_CSV_CACHE) && !from_cache) {
    tmp_path = cache_path + ".tmp";
    fw = std::fopen(tmp_path.c_str(), "wb");
    if (fw) {
      RawCacheHeader h{};
      std::memcpy(h.magic, kRawMagic, 8);
      h.version = 1;
      h.cols = (uint32_t)c->C;
      h.rows = R;
      h.csv_size = csv.size();
      h.csv_mtime_ns = csv.mtime_ns();
      if (std::fwrite(&h, sizeof(h), 1, fw) != 1) { std::fclose(fw); fw = nullptr; }
    }
  }
  int64_t rows_seen = 0;
  int stage = 0;
  if (from_cache) {
    st.from_cache = 1;
    for (int64_t r0 = 0; r0 < R && rc == SGD_OK; r0 += stage_rows, stage ^= 1) {
      const int64_t n = std::min(stage_rows, R - r0);
      if (cudaEventSynchronize(ev[stage]) != cudaSuccess) { rc = fail(SGD_ERR_CUDA, "sgd_load_csv: upload failed"); break; }
      const bool mine = r0 < c->row_end && r0 + n > c->row_begin;
      if (mine) {
        if (std::fread(h_stage[stage], sizeof(float) * (size_t)W, (size_t)n, fc) != (size_t)n) {
          rc = fail(SGD_ERR_IO, "raw cache " + cache_path + " is truncated");
          break;
        }
        rc = upload(stage, r0, n);
      } else if (std::fseek(fc, (long)(sizeof(float) * (size_t)(W * n)), SEEK_CUR) != 0) {
        rc = fail(SGD_ERR_IO, "raw cache seek failed");
      }
    }
    rows_seen = R;
  } else {
    // a stage = consecutive pieces holding at most stage_rows rows, parsed by one thread per piece
    const auto& pcs = csv.pieces();
    size_t k = 0;
    while (k < pcs.size() && rc == SGD_OK) {
      size_t k1 = k;
      int64_t n = 0;
      while (k1 < pcs.size() && (k1 == k || n + pcs[k1].rows <= stage_rows)) n += pcs[k1++].rows;
      const int64_t r0 = pcs[k].row0;
      if (n > stage_rows) {  // cannot happen for well-formed input (>= 2 bytes of text per value); a file of empty lines could
        rc = fail(SGD_ERR_PARSE, "a 4 MB piece of the CSV holds more rows than the staging buffer");
        break;
      }
      if (cudaEventSynchronize(ev[stage]) != cudaSuccess) { rc = fail(SGD_ERR_CUDA, "sgd_load_csv: upload failed"); break; }
      const bool mine = (r0 < c->row_end && r0 + n > c->row_begin) || fw != nullptr;
      if (mine) {
        const auto tp = clk::now();
        const int64_t nz = std::min<int64_t>(n, std::max<int64_t>(0, R - r0));
        std::memset(h_stage[stage], 0, sizeof(float) * (size_t)(nz * W));  // thrust::host_vector zero-fills
        std::atomic<int> prc{0};
        std::atomic<size_t> next{k};
        const unsigned hw = std::thread::hardware_concurrency();
        const size_t T = std::min<size_t>(k1 - k, hw ? hw : 4);
        std::vector<std::thread> th;
        for (size_t t = 0; t < T; ++t)
          th.emplace_back([&] {
            for (size_t q = next.fetch_add(1); q < k1; q = next.fetch_add(1)) {
              const int r = csv.parse_piece(q, R, c->C, h_stage[stage], r0);
              if (r != 0) prc.store(r);
            }
          });
        for (auto& t : th) t.join();
        st.parse_seconds += ms_since(tp) * 1e-3;
        if (prc.load() != 0) {
          rc = fail(SGD_ERR_PARSE, std::string("CSV has more rows/columns than announced or an unparsable cell: ") + path);
          break;
        }
        if (fw && nz > 0 && std::fwrite(h_stage[stage], sizeof(float) * (size_t)W, (size_t)nz, fw) != (size_t)nz) {
          std::fclose(fw);
          fw = nullptr;
          std::remove(tmp_path.c_str());
        }
        if (nz > 0) rc = upload(stage, r0, nz);
      }
      rows_seen = std::max(rows_seen, std::min(R, r0 + n));
      stage ^= 1;
      k = k1;
    }
  }
  if (rc == SGD_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(SGD_ERR_CUDA, "sgd_load_csv: upload failed");
  if (rc == SGD_OK && rows_seen < R) {
    // fewer lines than announced: the remaining rows are zeros (the reference's zero-filled host vectors)
    const int64_t lo = std::max(rows_seen, c->row_begin), hi = c->row_end;
    if (hi > lo) {
      if (cudaMemsetAsync(c->d_rec + (lo - c->row_begin) * c->ld, 0, sizeof(float) * (size_t)((hi - lo) * c->ld), c->stream) != cudaSuccess ||
          cudaStreamSynchronize(c->stream) != cudaSuccess)
        rc = fail(SGD_ERR_CUDA, "sgd_load_csv: zero fill failed");
    }
    if (fw) {  // the cache holds exactly `rows` records
      std::vector<float> z((size_t)W, 0.f);
      for (int64_t r = rows_seen; r < R && fw; ++r)
        if (std::fwrite(z.data(), sizeof(float) * (size_t)W, 1, fw) != 1) { std::fclose(fw); fw = nullptr; std::remove(tmp_path.c_str()); }
    }
  }
  if (fw) {
    const bool ok = std::fclose(fw) == 0 && rc == SGD_OK;
    fw = nullptr;
    if (ok) std::rename(tmp_path.c_str(), cache_path.c_str());
    else std::remove(tmp_path.c_str());
    st.cache_written = ok ? 1 : 0;
  }
  cleanup();
  if (rc != SGD_OK) return rc;
  c->data_loaded = true;
  st.rows = R;
  st.total_seconds = ms_since(t_all) * 1e-3;
  if (stats) *stats = st;
  return SGD_OK;
}

int sgd_load_synthetic(sgd_ctx* c, uint64_t seed, int32_t round_fp16) {
  SGD_TRY(check_ctx(c));
  SGD_TRY(bind(c));
  b200sgd::sgd_true_weights_kernel<<<(c->C + 255) / 256, 256, 0, c->stream>>>(seed, c->C, c->d_wtrue);
  if (c->local_rows > 0) {
    const int blocks = (int)std::min<int64_t>((c->local_rows + 7) / 8, (int64_t)c->num_sms * 16);
    b200sgd::sgd_synth_kernel<<<blocks, 256, 0, c->stream>>>(seed, c->row_begin, c->local_rows, c->C, c->ld,
                                                             c->d_wtrue, round_fp16, c->d_rec);
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->data_loaded = true;
  c->have_true_w = true;
  return SGD_OK;
}

int sgd_get_rows(sgd_ctx* c, int64_t row_begin, int64_t nrows, float* X, float* y) {
  SGD_TRY(check_ctx(c));
  if (!X || !y || nrows < 0 || row_begin < c->row_begin || row_begin + nrows > c->row_end)
    return fail(SGD_ERR_ARG, "sgd_get_rows: range outside this rank's shard");
  if (!c->data_loaded) return fail(SGD_ERR_STATE, "no data loaded");
  if (nrows == 0) return SGD_OK;
  SGD_TRY(bind(c));
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(nrows, (64ll << 20) / (4ll * (c->C + 1))));
  float *dX = nullptr, *dy = nullptr;
  CUDA_TRY(cudaMalloc(&dX, sizeof(float) * (size_t)chunk * c->C));
  if (cudaMalloc(&dy, sizeof(float) * (size_t)chunk) != cudaSuccess) {
    cudaFree(dX);
    return fail(SGD_ERR_NOMEM, "staging allocation failed");
  }
  cudaError_t e = cudaSuccess;
  for (int64_t r0 = row_begin; r0 < row_begin + nrows && e == cudaSuccess; r0 += chunk) {
    const int64_t n = std::min(chunk, row_begin + nrows - r0);
    const int blocks = (int)std::min<int64_t>((n * (c->C + 1) + 255) / 256, (int64_t)c->num_sms * 16);
    b200sgd::sgd_unpack_kernel<<<blocks, 256, 0, c->stream>>>(c->d_rec + (r0 - c->row_begin) * c->ld, n, c->C,
                                                              c->ld, dX, dy);
    e = cudaGetLastError();
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(X + (r0 - row_begin) * c->C, dX, sizeof(float) * (size_t)n * c->C, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(y + (r0 - row_begin), dy, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  }
  cudaFree(dX);
  cudaFree(dy);
  if (e != cudaSuccess) return fail(SGD_ERR_CUDA, std::string("sgd_get_rows: ") + cudaGetErrorString(e));
  return SGD_OK;
}

int sgd_get_true_weights(sgd_ctx* c, float* w_true) {
  SGD_TRY(check_ctx(c));
  if (!w_true) return fail(SGD_ERR_ARG, "null output");
  if (!c->have_true_w) return fail(SGD_ERR_STATE, "sgd_load_synthetic was not called");
  SGD_TRY(bind(c));
  CUDA_TRY(cudaMemcpy(w_true, c->d_wtrue, sizeof(float) * c->C, cudaMemcpyDeviceToHost));
  return SGD_OK;
}

int sgd_get_weights(sgd_ctx* c, float* w) {
  SGD_TRY(check_ctx(c));
  if (!w) return fail(SGD_ERR_ARG, "null output");
  SGD_TRY(bind(c));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpy(w, c->d_w, sizeof(float) * c->C, cudaMemcpyDeviceToHost));
  return SGD_OK;
}

int sgd_set_weights(sgd_ctx* c, const float* w) {
  SGD_TRY(check_ctx(c));
  if (!w) return fail(SGD_ERR_ARG, "null input");
  SGD_TRY(bind(c));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpy(c->d_w, w, sizeof(float) * c->C, cudaMemcpyHostToDevice));
  return SGD_OK;
}

int sgd_get_config(const sgd_ctx* c, sgd_config* out) {
  SGD_TRY(check_ctx(c));
  if (!out) return fail(SGD_ERR_ARG, "null output");
  *out = c->cfg;
  out->grid = c->grid;
  out->reduce = c->reduce_mode;
  out->engine = c->engine;
  out->batch = c->B;
  out->reserved[0] = c->nb;
  out->reserved[1] = c->ntiles * c->tile_rows;
  out->reserved[2] = (int32_t)c->smem_bytes;
  // KC: negative = column mapping; 0 = cluster kernel (then U = CTAs of the cluster)
  out->reserved[3] = c->cluster_cs > 0 ? 0 : c->kc.wide ? -c->kc.KC : c->kc.KC;
  out->reserved[4] = c->cluster_cs > 0 ? c->cluster_cs : c->kc.narrow8 ? -c->kc.U : c->kc.U;  // negative: 8 lanes per row
  out->reserved[5] = c->kc.NW | (c->kc.cw << 8);  // bits 8..: COLSPLIT warps along the columns
  out->reserved[6] = (int32_t)c->ld;
  return SGD_OK;
}

int64_t sgd_local_rows(const sgd_ctx* c) { return c ? c->local_rows : 0; }
int64_t sgd_local_row_begin(const sgd_ctx* c) { return c ? c->row_begin : 0; }

// ---- permutation --------------------------------------------------------------------------------
int sgd_shuffle(sgd_ctx* c) {
  SGD_TRY(check_ctx(c));
  SGD_TRY(bind(c));
  const auto t0 = clk::now();
  if (c->device_shuffle) {
    int launches = 0;
    SGD_TRY(take_or_enqueue_shuffle(c, &launches));
    CUDA_TRY(cudaStreamSynchronize(c->shuf_stream));
    SGD_TRY(check_watchdog(c));
    c->last_shuffle_ms = (float)ms_since(t0);
    c->last_shuffle_launches = launches;
    return SGD_OK;
  }
  SGD_TRY(sync_host_ind(c));
  SGD_TRY(ensure_host_stage(c));
  // main.cu:90 / :207: std::random_shuffle(ind_vector.begin(), ind_vector.end()), cumulative
  b200sgd::exact_shuffle(c->rng, c->h_ind.data(), c->R);
  const int buf = c->cur ^ 1;
  CUDA_TRY(cudaEventSynchronize(c->ev_perm[buf]));  // earlier upload out of this staging buffer
  std::memcpy(c->h_stage[buf], c->h_ind.data(), sizeof(int32_t) * (size_t)c->R);
  SGD_TRY(upload_perm(c, buf));  // main.cu:91 / :211: batch_indices = ind_vector
  CUDA_TRY(cudaEventSynchronize(c->ev_perm[buf]));
  c->cur = buf;
  c->perm_valid = true;
  c->last_shuffle_ms = (float)ms_since(t0);
  return SGD_OK;
}

int sgd_set_permutation(sgd_ctx* c, const int32_t* perm) {
  SGD_TRY(check_ctx(c));
  if (!perm) return fail(SGD_ERR_ARG, "null permutation");
  {  // a permutation of 0..R-1: the bucketed row lists of nranks > 1 are sized for exactly that
    std::vector<uint64_t> seen(((size_t)c->R + 63) / 64, 0);
    for (int64_t i = 0; i < c->R; ++i) {
      const int64_t v = perm[i];
      if (v < 0 || v >= c->R) return fail(SGD_ERR_ARG, "permutation entry out of range");
      uint64_t& wd = seen[(size_t)v >> 6];
      const uint64_t bit = 1ull << (v & 63);
      if (wd & bit) return fail(SGD_ERR_ARG, "not a permutation: row " + std::to_string(v) + " appears twice");
      wd |= bit;
    }
  }
  SGD_TRY(bind(c));
  if (c->ahead) {  // the look-ahead shuffled the OLD permutation
    if (c->ahead_shared) {
      c->reapply_k = c->shuf_no;  // its position map stays valid: the next shuffle request applies it to the new one
    } else {
      c->rng = c->rng_saved;  // the next shuffle draws the same rand() values again
      c->shuf_no -= 1;
    }
    c->ahead = false;
  }
  SGD_TRY(ensure_host_stage(c));
  c->h_ind.resize((size_t)c->R);
  std::memcpy(c->h_ind.data(), perm, sizeof(int32_t) * (size_t)c->R);
  c->host_ind_stale = false;
  const int buf = c->cur ^ 1;
  CUDA_TRY(cudaEventSynchronize(c->ev_perm[buf]));
  std::memcpy(c->h_stage[buf], perm, sizeof(int32_t) * (size_t)c->R);
  SGD_TRY(upload_perm(c, buf));
  CUDA_TRY(cudaEventSynchronize(c->ev_perm[buf]));
  c->cur = buf;
  c->perm_valid = true;
  c->last_shuffle_ms = 0.f;
  return SGD_OK;
}

int sgd_get_permutation(sgd_ctx* c, int32_t* perm) {
  SGD_TRY(check_ctx(c));
  if (!perm) return fail(SGD_ERR_ARG, "null output");
  SGD_TRY(bind(c));
  if (c->perm_valid) {
    // read it back from the DEVICE copy the step kernel consumes
    CUDA_TRY(cudaStreamSynchronize(c->shuf_stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(perm, c->d_perm[c->cur], sizeof(int32_t) * (size_t)c->R, cudaMemcpyDeviceToHost));
  } else {
    SGD_TRY(sync_host_ind(c));  // before the first shuffle: the identity
    std::memcpy(perm, c->h_ind.data(), sizeof(int32_t) * (size_t)c->R);
  }
  return SGD_OK;
}

// ---- the step loop ------------------------------------------------------------------------------
namespace {
int epoch_impl(sgd_ctx* c, int32_t epoch, sgd_timings* out, bool lookahead) {
  SGD_TRY(check_ctx(c));
  if (epoch < 1) return fail(SGD_ERR_ARG, "epoch is 1-based");
  if (!c->data_loaded) return fail(SGD_ERR_STATE, "no data loaded");
  if (!c->perm_valid) return fail(SGD_ERR_STATE, "no permutation: call sgd_shuffle or sgd_set_permutation first");
  SGD_TRY(bind(c));
  int launches = 0;
  CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_perm[c->cur], 0));
  CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_shuf[c->cur], 0));
  SGD_TRY(bucket_perm(c, c->cur, &launches));
  CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  SGD_TRY(launch_epoch(c, epoch, c->cur, &launches));
  CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  CUDA_TRY(cudaEventRecord(c->ev_done[c->cur], c->stream));
  if (lookahead) SGD_TRY(enqueue_lookahead(c, &launches));  // the next iteration's permutation, beside this epoch
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  SGD_TRY(check_watchdog(c));
  SGD_TRY(collect_losses(c));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  fill_timings(c, out, c->last_shuffle_ms, ms, launches + c->last_shuffle_launches, 1);
  c->last_shuffle_ms = 0.f;
  c->last_shuffle_launches = 0;
  return SGD_OK;
}
}  // namespace

// sgd_epoch runs the step kernel on its own (what the sweeps and `kernel_alone` time); sgd_iteration, the
// drop-in for one cuda_iteration / cublas_iteration call, also starts the next call's permutation.
int sgd_epoch(sgd_ctx* c, int32_t epoch, sgd_timings* out) { return epoch_impl(c, epoch, out, false); }

int sgd_iteration(sgd_ctx* c, int32_t epoch, sgd_timings* out) {
  SGD_TRY(sgd_shuffle(c));
  return epoch_impl(c, epoch, out, true);
}

int sgd_run(sgd_ctx* c, int32_t first_epoch, int32_t num_epochs, sgd_timings* total) {
  SGD_TRY(check_ctx(c));
  if (first_epoch < 1 || num_epochs < 0) return fail(SGD_ERR_ARG, "bad epoch range");
  if (!c->data_loaded) return fail(SGD_ERR_STATE, "no data loaded");
  SGD_TRY(bind(c));
  if (num_epochs == 0) {
    fill_timings(c, total, 0.f, 0.f, 0, 0);
    return SGD_OK;
  }
  const auto t0 = clk::now();
  std::vector<cudaEvent_t> evs(2 * (size_t)num_epochs, nullptr);
  auto cleanup = [&] {
    for (cudaEvent_t e : evs)
      if (e) cudaEventDestroy(e);
  };
  for (auto& e : evs)
    if (cudaEventCreate(&e) != cudaSuccess) {
      cleanup();
      return fail(SGD_ERR_CUDA, "cudaEventCreate failed");
    }
  int launches = 0;
  int rc = SGD_OK;
  auto stage_next = [&]() -> int {  // host part of the shuffle runs while the GPU runs the previous epoch
    if (c->device_shuffle) return take_or_enqueue_shuffle(c, &launches);
    SGD_TRY(sync_host_ind(c));
    SGD_TRY(ensure_host_stage(c));
    b200sgd::exact_shuffle(c->rng, c->h_ind.data(), c->R);
    const int buf = c->cur ^ 1;
    CUDA_TRY(cudaEventSynchronize(c->ev_perm[buf]));
    std::memcpy(c->h_stage[buf], c->h_ind.data(), sizeof(int32_t) * (size_t)c->R);
    SGD_TRY(upload_perm(c, buf));
    c->cur = buf;
    c->perm_valid = true;
    return SGD_OK;
  };
  rc = stage_next();
  for (int e = 0; e < num_epochs && rc == SGD_OK; ++e) {
    const int buf = c->cur;
    auto enqueue = [&]() -> int {
      CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_perm[buf], 0));
      CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_shuf[buf], 0));
      SGD_TRY(bucket_perm(c, buf, &launches));
      CUDA_TRY(cudaEventRecord(evs[2 * e], c->stream));
      SGD_TRY(launch_epoch(c, first_epoch + e, buf, &launches));
      CUDA_TRY(cudaEventRecord(evs[2 * e + 1], c->stream));
      CUDA_TRY(cudaEventRecord(c->ev_done[buf], c->stream));
      return SGD_OK;
    };
    rc = enqueue();
    if (rc == SGD_OK && e + 1 < num_epochs) rc = stage_next();  // overlaps with the kernel above
    else if (rc == SGD_OK) rc = enqueue_lookahead(c, &launches);  // the next call's first permutation, beside the last epoch
    if (rc == SGD_OK && c->loss_pending >= kLossSlots - 1) {      // very long runs: drain the loss slots
      if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(SGD_ERR_CUDA, "epoch loop failed");
      if (rc == SGD_OK) rc = collect_losses(c);
    }
  }
  cudaError_t se = cudaStreamSynchronize(c->stream);
  if (se == cudaSuccess) se = cudaStreamSynchronize(c->shuf_stream);
  if (rc == SGD_OK && se != cudaSuccess) rc = fail(SGD_ERR_CUDA, std::string("epoch loop: ") + cudaGetErrorString(se));
  if (rc == SGD_OK) rc = check_watchdog(c);
  if (rc == SGD_OK) rc = collect_losses(c);
  float loop_ms = 0.f;
  if (rc == SGD_OK) {
    for (int e = 0; e < num_epochs; ++e) {
      float ms = 0.f;
      if (c->nb > 0 && cudaEventElapsedTime(&ms, evs[2 * e], evs[2 * e + 1]) == cudaSuccess) loop_ms += ms;
    }
  }
  cleanup();
  if (rc != SGD_OK) return rc;
  const float wall = (float)ms_since(t0);
  fill_timings(c, total, std::max(0.f, wall - loop_ms), loop_ms, launches, num_epochs);
  if (total) total->epoch_ms = wall;
  return SGD_OK;
}

int sgd_get_residuals(sgd_ctx* c, float* r, int64_t count) {
  SGD_TRY(check_ctx(c));
  if (!c->d_rout) return fail(SGD_ERR_STATE, "context was created without SGD_FLAG_RESIDUALS");
  if (!r || count < 0 || count > c->rout_len) return fail(SGD_ERR_ARG, "bad residual count");
  SGD_TRY(bind(c));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpy(r, c->d_rout, sizeof(float) * (size_t)count, cudaMemcpyDeviceToHost));
  return SGD_OK;
}

int sgd_get_last_gradient(sgd_ctx* c, float* g) {
  SGD_TRY(check_ctx(c));
  if (!g) return fail(SGD_ERR_ARG, "null output");
  SGD_TRY(bind(c));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  std::vector<float> tmp((size_t)c->C4 * 4);
  CUDA_TRY(cudaMemcpy(tmp.data(), c->d_glast, sizeof(float) * tmp.size(), cudaMemcpyDeviceToHost));
  std::memcpy(g, tmp.data(), sizeof(float) * c->C);
  return SGD_OK;
}

int sgd_mean_abs_error(sgd_ctx* c, float* mae_out, double* sum_out) {
  SGD_TRY(check_ctx(c));
  if (!c->data_loaded) return fail(SGD_ERR_STATE, "no data loaded");
  SGD_TRY(bind(c));
  double total = 0.0;
  if (c->local_rows > 0) {
    // lanes per row: the power of two that leaves a lane about four float4 of the record
    int L = 1;
    while (L < 32 && (int64_t)L * 4 < c->ld / 4) L *= 2;
    const int64_t rows_per_block = (int64_t)8 * (32 / L) * 2;  // 8 warps, 32/L groups, two rows per pass
    const int blocks = (int)std::min<int64_t>(c->mae_blocks, (c->local_rows + rows_per_block - 1) / rows_per_block);
    b200sgd::sgd_mae_kernel<<<blocks, 256, sizeof(float) * (size_t)c->ld, c->stream>>>(
        c->d_rec, c->ld, c->C, c->local_rows, c->d_w, L, c->d_mae);
    CUDA_TRY(cudaGetLastError());
    std::vector<double> h((size_t)blocks);
    CUDA_TRY(cudaMemcpyAsync(h.data(), c->d_mae, sizeof(double) * blocks, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (double v : h) total += v;
  }
  if (sum_out) *sum_out = total;
  // sgd_thrust.cu:134: return error_sum / R
  if (mae_out) *mae_out = (float)(total / (double)(c->cfg.nranks > 1 ? std::max<int64_t>(1, c->local_rows) : c->R));
  return SGD_OK;
}

int sgd_get_loss_curve(sgd_ctx* c, double* sum_abs, double* sum_sq, int32_t max_epochs) {
  SGD_TRY(check_ctx(c));
  if (max_epochs < 0) return fail(SGD_ERR_ARG, "bad count");
  const int n = (int)std::min<size_t>((size_t)max_epochs, c->curve_abs.size());
  const size_t first = c->curve_abs.size() - (size_t)n;  // the most recent epochs
  for (int i = 0; i < n; ++i) {
    if (sum_abs) sum_abs[i] = c->curve_abs[first + i];
    if (sum_sq) sum_sq[i] = c->curve_sq[first + i];
  }
  return n;
}

int sgd_get_local_schedule(sgd_ctx* c, int32_t* local_perm, int64_t local_cap, int64_t* step_off) {
  SGD_TRY(check_ctx(c));
  if (!local_perm || !step_off) return fail(SGD_ERR_ARG, "null output");
  if (c->cfg.nranks < 2) return fail(SGD_ERR_STATE, "single-rank contexts have no bucketed schedule");
  SGD_TRY(bind(c));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpy(step_off, c->d_step_off[c->cur], sizeof(int64_t) * ((size_t)c->nb + 1), cudaMemcpyDeviceToHost));
  const int64_t n = step_off[c->nb];
  if (n > local_cap) return fail(SGD_ERR_ARG, "local_perm too small");
  CUDA_TRY(cudaMemcpy(local_perm, c->d_perm_local[c->cur], sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
  return SGD_OK;
}

int sgd_debug_phase_cycles(sgd_ctx* c, uint64_t* out, int32_t max_ctas) {
  SGD_TRY(check_ctx(c));
  if (!out || max_ctas < 0) return fail(SGD_ERR_ARG, "bad output");
  if (!c->d_dbg) return fail(SGD_ERR_STATE, "set B200SGD_PHASE_PROFILE=1 before sgd_create");
  SGD_TRY(bind(c));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  const int n = std::min<int>(max_ctas, 2 * c->grid);  // [grid..2*grid): wall-clock snapshot of one step
  CUDA_TRY(cudaMemcpy(out, c->d_dbg, sizeof(uint64_t) * 8 * (size_t)n, cudaMemcpyDeviceToHost));
  return n;
}

// ---- peers --------------------------------------------------------------------------------------
int sgd_peer_export(sgd_ctx* c, void* handle_out) {
  SGD_TRY(check_ctx(c));
  if (!handle_out) return fail(SGD_ERR_ARG, "null handle");
  if (c->cfg.nranks < 2) return fail(SGD_ERR_STATE, "single-rank context has no exchange buffer");
  SGD_TRY(bind(c));
  static_assert(sizeof(cudaIpcMemHandle_t) + 16 <= SGD_PEER_HANDLE_BYTES, "handle too small");
  std::memset(handle_out, 0, SGD_PEER_HANDLE_BYTES);
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, c->d_xchg_block));
  std::memcpy(handle_out, &h, sizeof(h));
  const uint64_t meta[2] = {(uint64_t)c->xchg_bytes, (uint64_t)c->grid};
  std::memcpy((char*)handle_out + sizeof(h), meta, sizeof(meta));
  return SGD_OK;
}

int sgd_peer_import(sgd_ctx* c, int32_t peer_rank, const void* handle) {
  SGD_TRY(check_ctx(c));
  if (!handle || peer_rank < 0 || peer_rank >= c->cfg.nranks) return fail(SGD_ERR_ARG, "bad peer");
  if (peer_rank == c->cfg.rank) return SGD_OK;
  SGD_TRY(bind(c));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  uint64_t meta[2];
  std::memcpy(meta, (const char*)handle + sizeof(h), sizeof(meta));
  if (meta[0] != c->xchg_bytes || meta[1] != (uint64_t)c->grid)
    return fail(SGD_ERR_COMM, "peer was configured with a different shape / grid");
  void* ptr = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(SGD_ERR_COMM, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  }
  c->peer_block[peer_rank] = ptr;
  c->peer_open[peer_rank] = true;
  return SGD_OK;
}

int sgd_peer_ready(sgd_ctx* c) {
  SGD_TRY(check_ctx(c));
  for (int r = 0; r < c->cfg.nranks; ++r)
    if (!c->peer_block[r]) return fail(SGD_ERR_STATE, "peer " + std::to_string(r) + " was not imported");
  c->peers_ready = true;
  return SGD_OK;
}

}  // extern "C"
