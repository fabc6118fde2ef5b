/*
 * oracle/sgd_oracle.h -- CPU restatement of the mini-batch SGD hot path of
 * thvasilo/cuda-sgd-sese-project (reference files cited per function in sgd_oracle.c).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product library
 * (libb200sgd.so) never links, loads or calls anything in this directory.
 *
 * Parity status: PINNED.
 *   - every known-answer fixture of the reference's own test code (c_code/testing.cu +
 *     c_code/test-data/) is checked in tests/test_oracle.py;
 *   - the third-party pieces the reference leans on (glibc rand, libstdc++ random_shuffle, Thrust
 *     minstd + uniform_real_distribution) are pinned against outputs of the REAL implementations
 *     run in this image (tests/golden/thirdparty_kat.json, generator committed beside it);
 *   - the unmodified reference binary (oracle/_ref/main, built by oracle/build_ref.sh from the
 *     sources where they lie) is run on the GPU box by tests/test_reference_binary.py and its
 *     MAE / weights are compared with this oracle.
 */
#ifndef SGD_ORACLE_H_
#define SGD_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- glibc rand() (TYPE_3 additive feedback, seed 1) : call sites main.cu:90, main.cu:207 ---- */
typedef struct {
  uint32_t st[34];
  uint64_t n;
} oracle_rand_t;
void oracle_rand_init(oracle_rand_t* g, uint32_t seed);
int32_t oracle_rand_next(oracle_rand_t* g);

/* ---- libstdc++ std::random_shuffle(first,last) : main.cu:90 / main.cu:207 ---- */
void oracle_random_shuffle(oracle_rand_t* g, int32_t* ind, int64_t n);

/* ---- weight init, main.cu:389-395 (thrust::default_random_engine + uniform_real_distribution) */
void oracle_init_weights(float* w, int32_t C);

/* ---- main.cu:343-344 : batch size and number of batches ---- */
int32_t oracle_effective_batch(int32_t R, int32_t B_arg);
int32_t oracle_num_batches(int32_t R, int32_t B);

/* ---- main.cu:144 / main.cu:282 : a = -(lr / pow(epoch, 0.25)) narrowed to float ---- */
float oracle_step_scale(float lr, int32_t epoch);

/* ---- sgd_thrust.cu:323-364 (cuBLAS path) / :19-51 (plain path): r = X_b w - y_b ---- */
void oracle_residuals(const float* X, const float* y, const float* w, const int32_t* rows,
                      int32_t nrows, int32_t C, float* r_out);

/* ---- sgd_thrust.cu:301-315 + :254-291 : g[c] = (sum_i r[i] * X[rows[i],c]) / (float)B ---- */
void oracle_gradient(const float* X, const float* r, const int32_t* rows, int32_t nrows, int32_t C,
                     float denom, float* g_out);

/* ---- one epoch (main.cu:56-163 plain path == main.cu:168-301 cuBLAS path, same maths).
 *      `ind` must already hold this epoch's permutation.  acc64 != 0 uses double accumulators
 *      (used to bound fp32 summation-order effects).  threads > 1 = OpenMP CPU baseline. ---- */
void oracle_epoch(const float* X, const float* y, int64_t R, int32_t C, int32_t B,
                  const int32_t* ind, float* w, float lr, int32_t epoch, int acc64, int threads);

/* Multi-rank emulation of the row-sharded data-parallel step (new functionality, north_star):
 * every step's batch is split by owner (owner = row / ceil(R/nranks)), each rank sums its own
 * partial gradient, partials are added in rank order.  Equals oracle_epoch up to fp32 order. */
void oracle_epoch_sharded(const float* X, const float* y, int64_t R, int32_t C, int32_t B,
                          const int32_t* ind, float* w, float lr, int32_t epoch, int32_t nranks);

/* ---- sgd_thrust.cu:79-135 : mean absolute error over the data in ORIGINAL order ---- */
float oracle_mean_abs_error(const float* X, const float* y, const float* w, int64_t R, int32_t C,
                            int threads);

/* ---- whole run, main.cu:338-530 minus I/O: returns the MAE, leaves final weights in w_out and
 *      (if non-NULL) the last permutation in ind_out. ---- */
float oracle_run(const float* X, const float* y, int64_t R, int32_t C, int32_t B_arg, float lr,
                 int32_t epochs, int acc64, int threads, float* w_out, int32_t* ind_out);

/* ---- sgd_io.cu:38-77 : read_csv ("f1,...,fC,label" lines, no header).  Returns 0 on success,
 *      -1 if the file cannot be opened (reference returns false).  Unlike the reference it
 *      refuses to write outside the caller's buffers (returns -2). ---- */
int oracle_read_csv(const char* path, int64_t R, int32_t C, float* X, float* y);

/* ---- the test ops of testing.cu that are not on the hot path but pin the fixtures ---- */
void oracle_permute_rows(const float* X, const float* y, const int32_t* order, int64_t R, int32_t C,
                         float* Xs, float* ys); /* sgd_thrust.cu:374-402 */
void oracle_scale_rows(const float* X, const float* v, int64_t R, int32_t C,
                       float* out); /* sgd_thrust.cu:301-315 */
void oracle_col_sums(const float* G, int64_t R, int32_t C, float scale,
                     float* out); /* sgd_thrust.cu:216-291 (scale=1 => calculate_column_sums) */

int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif /* SGD_ORACLE_H_ */
