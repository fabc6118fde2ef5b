/*
 * include/b200sgd.h -- C-ABI of libb200sgd.so, the B200-native mini-batch SGD engine for linear
 * regression that drops in for the hot path of thvasilo/cuda-sgd-sese-project.
 *
 * The reference has no plugin / FFI layer: its boundary is the free functions c_code/main.cu calls
 * and the process contract of ./main.  Every entry point below names the reference interface it
 * replaces (file:line under /root/reference/c_code).  Plain pointers and sizes only; no C++ or
 * torch types cross this boundary.  All functions return 0 (SGD_OK) or a negative sgd_status;
 * sgd_last_error() gives the text of the last failure on the calling thread.  Nothing throws.
 *
 * A context is single-owner (not thread-safe), bound to one CUDA device, and owns every device
 * allocation, stream, event and peer mapping it creates.  Host arrays passed in stay caller-owned.
 *
 * There is NO CPU fallback anywhere behind this header: every compute entry point fails with
 * SGD_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef B200SGD_H_
#define B200SGD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SGD_ABI_VERSION 1

typedef enum {
  SGD_OK = 0,
  SGD_ERR_ARG = -1,     /* bad argument / unsupported shape                         */
  SGD_ERR_IO = -2,      /* file could not be opened / written                       */
  SGD_ERR_PARSE = -3,   /* CSV has more rows / columns than announced               */
  SGD_ERR_CUDA = -4,    /* CUDA runtime / driver error, or no sm_100 device         */
  SGD_ERR_STATE = -5,   /* call sequence error (e.g. epoch before data was loaded)  */
  SGD_ERR_NOMEM = -6,   /* host or device allocation failed                         */
  SGD_ERR_COMM = -7     /* peer / collective set-up failed                          */
} sgd_status;

/* Which of the reference's three code paths the caller asked for (main.cu:345, typedefs.cuh:36).
 * All three are served by the same fused kernel; the selector only changes the report label and
 * the JSON suffix ("-cublas.json" / "-cuda.json", main.cu:545,557). */
typedef enum { SGD_PATH_CUBLAS = 0, SGD_PATH_PLAIN_CUDA = 1, SGD_PATH_LVL1 = 2 } sgd_codepath;

/* How the sequential step chain is executed. */
typedef enum {
  SGD_ENGINE_AUTO = 0,
  SGD_ENGINE_PERSISTENT = 1, /* one cooperative launch per epoch, steps chained inside (default) */
  SGD_ENGINE_GRAPH = 2       /* one fused kernel per step, steps captured in a CUDA graph       */
} sgd_engine;

/* How the per-CTA partial gradients become the new weights (DESIGN.md "grid reduction").  AUTO (and
 * DATAFLOW) lets narrow rows with small and mid-size steps run on the cluster kernel (DESIGN.md 3.2:
 * the CTAs of a thread-block cluster exchange their partials through distributed shared memory);
 * the other values pin the generic kernel to one scheme (A/B measurements). */
typedef enum {
  SGD_REDUCE_AUTO = 0,
  SGD_REDUCE_ALLGATHER = 1, /* every CTA sums all partials itself: 1 grid barrier, G^2*C traffic */
  SGD_REDUCE_SCATTER = 2,   /* column-sliced reduce + broadcast: 2 grid barriers, 2*G*C traffic  */
  SGD_REDUCE_DATAFLOW = 3,  /* column-sliced, tagged words instead of barriers: 2 one-way hops   */
  SGD_REDUCE_DATAFLOW_1HOP = 4, /* every CTA gathers all tagged partials itself: 1 hop, small grids */
  SGD_REDUCE_CLUSTER = 5    /* reduce-scatter and all-gather inside thread-block clusters (distributed
                               shared memory), one tagged hop between clusters / GPUs: what AUTO picks
                               for everything the cluster kernel does not take                        */
} sgd_reduce;

typedef struct sgd_config {
  int64_t rows;        /* R: number of data points of the WHOLE data set       (main.cu:341)   */
  int32_t cols;        /* C: number of features                                (main.cu:342)   */
  int32_t batch;       /* batch size argument, 0 = whole data set              (main.cu:343)   */
  float learning_rate; /*                                                      (main.cu:338)   */
  int32_t codepath;    /* sgd_codepath                                         (main.cu:345)   */
  int32_t device;      /* CUDA device ordinal                                                  */
  int32_t rank;        /* data-parallel rank of this context (0 for single GPU)               */
  int32_t nranks;      /* number of row shards / GPUs (1 for single GPU)                       */
  int32_t engine;      /* sgd_engine                                                           */
  int32_t reduce;      /* sgd_reduce                                                           */
  int32_t grid;        /* CTAs of the persistent kernel, 0 = auto (cluster kernel: a multiple of 8 / 16) */
  uint32_t shuffle_seed; /* seed of the rand() stream; 0 or 1 = the reference's unseeded rand() */
  int32_t flags;       /* SGD_FLAG_*                                                           */
  int32_t reserved[7];
} sgd_config;

#define SGD_FLAG_HOST_SHUFFLE 0x1 /* force the host Fisher-Yates even where the device one exists */
#define SGD_FLAG_RESIDUALS 0x2    /* keep the residuals of every step (sgd_get_residuals)         */

/* True per-epoch times in ms.  The reference's CudaTimings / CuBLASTimings (main.cu:33-49) are
 * accumulated wrongly (main.cu:154-159, :294-299); these are the real values. */
typedef struct sgd_timings {
  float shuffle_ms;   /* permutation generation + upload not hidden behind the previous epoch   */
  float steploop_ms;  /* first step launched -> last weight update complete (device time)       */
  float epoch_ms;     /* shuffle_ms + steploop_ms: what the reference's gpu_time adds per epoch */
  int32_t steps;      /* mini-batch steps executed = floor(R/(float)B)             (main.cu:344) */
  int32_t launches;   /* kernels of this library launched for the epoch                         */
  int64_t samples;    /* steps * B                                                              */
  int64_t bytes;      /* algorithmic HBM bytes of the step loop (DESIGN.md formula)             */
} sgd_timings;

typedef struct sgd_ctx sgd_ctx;

/* ---- library ------------------------------------------------------------------------------- */
int sgd_abi_version(void);
const char* sgd_last_error(void);
/* number of usable sm_100 devices (0 on a CPU-only host; never an error) */
int sgd_device_count(void);

/* ---- host-side pieces of the reference's L1 layer ------------------------------------------ */
/* replaces bool read_csv(std::string, thrust_host_float&, thrust_host_float&, int R, int C)
 * (sgd_io.cuh:91, sgd_io.cu:38-77).  X is R*C row-major, y is R, both caller-allocated.
 * SGD_ERR_IO when the file cannot be opened (reference: returns false, main.cu:366-370);
 * SGD_ERR_PARSE where the reference would write outside the buffers (sgd_io.cu:57-68). */
int sgd_read_csv(const char* path, int64_t rows, int32_t cols, float* X, float* y);
/* the inverse of read_csv: the file format of data_set_creation/generate_data.py:52-59 (csv.writer
 * rows `f1,...,fC,label`), every value as the shortest text that parses back to the same float;
 * written with all host cores (GB-size inputs of the benchmark's reference arm). */
int sgd_write_csv(const char* path, int64_t rows, int32_t cols, const float* X, const float* y);
/* replaces std::string get_filename_from_path(std::string) (sgd_io.cu:107-113) */
int sgd_get_filename_from_path(const char* path, char* out, size_t out_cap);
/* replaces ExperimentOutput + write_experiment_output + write_json (sgd_io.cuh:28-60,
 * sgd_io.cu:82-105): the five-key JSON, byte-compatible with jsoncpp's StyledWriter. */
int sgd_write_experiment_output(const char* path, float shuffle_time, float total_gradient_time,
                                float gpu_time, float transfer_time, float error);
/* replaces the weight initialisation loop main.cu:389-395 (Thrust minstd, seed 1, U(0,0.01)) */
int sgd_init_weights(float* w, int32_t cols);
/* replaces main.cu:343-344: effective batch size and number of batches (tail rows dropped) */
int sgd_batch_geometry(int64_t rows, int32_t batch_arg, int32_t* batch_out, int32_t* steps_out);
/* replaces the host side of main.cu:90 / :207: `epochs` cumulative std::random_shuffle passes over
 * ind[0..rows) driven by glibc rand() seeded with `seed` (0/1 = unseeded).  ind is caller-owned
 * and must already hold the previous state (identity before the first epoch, main.cu:421-424). */
int sgd_host_shuffle(int32_t* ind, int64_t rows, uint32_t seed, int32_t skip_epochs,
                     int32_t epochs);

/* ---- row sharding of the data-parallel path (new functionality, BASELINE.json north_star) ------
 * rank k owns rows [k*ceil(R/N), min(R, (k+1)*ceil(R/N))). */
int sgd_shard_range(int64_t rows, int32_t rank, int32_t nranks, int64_t* begin, int64_t* end);
/* host statement of the per-epoch bucketing the device does (sgd_bucket_* kernels): from the GLOBAL
 * permutation keep, per step, the rows `rank` owns, batch order preserved, as local row numbers;
 * step_off gets steps+1 offsets into local_perm. */
int sgd_host_bucket(const int32_t* perm, int64_t rows, int32_t batch_arg, int32_t rank, int32_t nranks,
                    int32_t* local_perm, int64_t local_cap, int64_t* step_off);

/* ---- context: replaces the set-up half of main() (main.cu:381-433) --------------------------
 * Allocates device storage for this rank's row shard [rank*ceil(R/nranks), ...), initialises the
 * weights (main.cu:389-395) and the identity index vector (main.cu:421-424).  X/y may be NULL
 * (load later); otherwise they are the host arrays read_csv filled (WHOLE data set, the shard is
 * cut out here) and are uploaded as main.cu:381-384 does. */
int sgd_create(const sgd_config* cfg, const float* X, const float* y, sgd_ctx** out);
void sgd_destroy(sgd_ctx* ctx);
/* upload rows [row_begin, row_begin+nrows) of the whole data set (host, row-major, C per row) */
int sgd_load_rows(sgd_ctx* ctx, int64_t row_begin, int64_t nrows, const float* X, const float* y);
/* read_csv + upload in one call (main.cu:365 + :381-384), the way the reference's dominant wall time
 * ("the most time consuming part", report/sese_report_tvas.tex:737-740) is best spent: the file is
 * mapped and cut into pieces of whole lines, all host cores parse piece k+1 into pinned memory while
 * piece k is copied to the device and packed into records, and no full host copy of the data ever
 * exists.  Same cell semantics and errors as sgd_read_csv.  With SGD_CSV_CACHE the parsed rows are
 * also written to `<path>.f32` (64-byte header: magic, rows, cols, size and mtime of the CSV; then
 * rows * (cols + 1) floats); a later call finds the cache, checks the header against the CSV and
 * streams the floats instead of parsing text.  stats may be NULL. */
#define SGD_CSV_CACHE 0x1
typedef struct sgd_load_stats {
  int64_t rows;           /* rows of the data set                                   */
  int64_t bytes_uploaded; /* host -> device bytes of this rank                      */
  double parse_seconds;   /* wall time inside the text parser (0 from the cache)    */
  double total_seconds;   /* whole call                                             */
  int32_t from_cache;     /* 1: the raw cache was used                              */
  int32_t cache_written;  /* 1: this call wrote the raw cache                       */
} sgd_load_stats;
int sgd_load_csv(sgd_ctx* ctx, const char* path, int32_t flags, sgd_load_stats* stats);
/* fill this rank's shard on the device with the data_set_creation/generate_data.py distribution
 * (X~N(0,1), w_true~100*U(0,1), y = X w_true), counter-based so the values do not depend on the
 * number of ranks.  round_fp16 != 0 rounds through float16 as generate_data.py:37 does. */
int sgd_load_synthetic(sgd_ctx* ctx, uint64_t seed, int32_t round_fp16);
/* read rows back (any rows of THIS rank's shard, global numbering) -- for parity tests */
int sgd_get_rows(sgd_ctx* ctx, int64_t row_begin, int64_t nrows, float* X, float* y);
int sgd_get_true_weights(sgd_ctx* ctx, float* w_true); /* of sgd_load_synthetic */

int sgd_get_weights(sgd_ctx* ctx, float* w);           /* what main.cu:560 would print */
int sgd_set_weights(sgd_ctx* ctx, const float* w);
int sgd_get_config(const sgd_ctx* ctx, sgd_config* out);
int64_t sgd_local_rows(const sgd_ctx* ctx);
int64_t sgd_local_row_begin(const sgd_ctx* ctx);

/* ---- the hot path ---------------------------------------------------------------------------
 * sgd_shuffle: replaces `std::random_shuffle(ind_vector); batch_indices = ind_vector;`
 *   (main.cu:90-91, :207-211) and removes permute_data_and_labels (sgd_thrust.cu:374-402): the
 *   permutation stays an index vector, rows are gathered inside the step kernel.
 * sgd_epoch: replaces the batch loop of cuda_iteration (main.cu:94-160) and cublas_iteration
 *   (main.cu:227-300) -- calculate_gradients / calculate_loss_derivative_cublas /
 *   scale_matrix_rows_by_vector / calculate_scaled_col_sums / calculate_row_sums / the two
 *   thrust::transform updates -- by the fused sm_100a step kernel.  `epoch` is 1-based and only
 *   enters through the step size a = -(lr / pow(epoch, 0.25)) (main.cu:144, :282).
 * sgd_iteration = sgd_shuffle + sgd_epoch: the drop-in for one call of cuda_iteration /
 *   cublas_iteration (main.cu:56, :168).
 * sgd_run: replaces the epoch loop main.cu:456-508 with the next permutation produced while the
 *   current epoch runs.
 * Look-ahead: the reference never seeds rand(), so the permutation sequence is fixed; sgd_create, the last
 *   epoch of sgd_run and sgd_iteration already enqueue the NEXT permutation on the device (beside the
 *   upload / the step kernel) and the next shuffle request takes it.  Not observable through this
 *   interface: sgd_get_permutation returns the current one, sgd_set_permutation discards the look-ahead
 *   and the next shuffle is drawn from the same rand() values.  sgd_epoch alone never starts one. */
int sgd_shuffle(sgd_ctx* ctx);
int sgd_set_permutation(sgd_ctx* ctx, const int32_t* perm); /* testing.cu:80-84 style, R entries */
int sgd_get_permutation(sgd_ctx* ctx, int32_t* perm);       /* R entries, global row numbers     */
int sgd_epoch(sgd_ctx* ctx, int32_t epoch, sgd_timings* out);
int sgd_iteration(sgd_ctx* ctx, int32_t epoch, sgd_timings* out);
int sgd_run(sgd_ctx* ctx, int32_t first_epoch, int32_t num_epochs, sgd_timings* total);
/* residuals r = X_b w - y_b of the last epoch's steps (batch order), needs SGD_FLAG_RESIDUALS:
 * what calculate_loss_derivative_cublas (sgd_thrust.cu:323-371) leaves in loss_derivative_d */
int sgd_get_residuals(sgd_ctx* ctx, float* r, int64_t count);
/* gradient g = X_b^T r / B of the last executed step: what calculate_scaled_col_sums
 * (sgd_thrust.cu:276-291) leaves in col_sums / main.cu:133 in row_sums */
int sgd_get_last_gradient(sgd_ctx* ctx, float* g);
/* replaces calculate_mean_abs_error_cublas (sgd_thrust.cu:79-135, call main.cu:525-530).
 * sum_out (optional) receives the SUM of absolute errors over this rank's rows: with nranks > 1
 * the caller adds the ranks' sums and divides by R.  mae_out (optional) is sum / R for nranks == 1
 * (the reference's value, sgd_thrust.cu:134) and sum / local rows otherwise. */
int sgd_mean_abs_error(sgd_ctx* ctx, float* mae_out, double* sum_out);

/* Running training loss per epoch (extension: the reference's squared_errors kernel,
 * sgd_thrust.cu:56-74, and its `errors` buffer, main.cu:406-407, are dead code).  For each of the
 * most recent `max_epochs` executed epochs: sum_abs = sum over the rows this rank processed of |r|,
 * sum_sq = sum of r*r, r = x.w - y with the weights the step used.  Divide by steps*B for the mean
 * absolute error / twice the half-MSE (nranks > 1: add the ranks first).  Returns the count. */
int sgd_get_loss_curve(sgd_ctx* ctx, double* sum_abs, double* sum_sq, int32_t max_epochs);

/* nranks > 1: the bucketed schedule of the last executed epoch, as the device built it */
int sgd_get_local_schedule(sgd_ctx* ctx, int32_t* local_perm, int64_t local_cap, int64_t* step_off);

/* diagnostics: when the context was created with B200SGD_PHASE_PROFILE=1 in the environment, the
 * epoch kernel sums, per CTA, the SM cycles of the phases of a step (rows | CTA reduce | publish
 * partial | gather+update (1-hop) | owner work (2-hop) | wait for weights | CTA barrier | -) of
 * the LAST launch.  out gets 8 values per CTA; returns the number of CTAs written or < 0.  (The
 * cluster kernel stores the SM clock at eight marks of its middle step instead) */
int sgd_debug_phase_cycles(sgd_ctx* ctx, uint64_t* out, int32_t max_ctas);

/* ---- multi-GPU (new functionality, BASELINE.json north_star; one process per GPU) -----------
 * The d-float gradient exchange runs inside the step kernel over NVLink peer memory.  Each rank
 * exports an opaque handle of its exchange buffer, the launcher (torch.distributed) all-gathers
 * the handles, every rank imports its peers'.  Handles are SGD_PEER_HANDLE_BYTES bytes. */
#define SGD_PEER_HANDLE_BYTES 128
int sgd_peer_export(sgd_ctx* ctx, void* handle_out);
int sgd_peer_import(sgd_ctx* ctx, int32_t peer_rank, const void* handle);
int sgd_peer_ready(sgd_ctx* ctx); /* after all imports; enables the fused exchange */

#ifdef __cplusplus
}
#endif
#endif /* B200SGD_H_ */
