/*
 * oracle/sgd_oracle.c -- CPU restatement of the reference's mini-batch SGD path.
 *
 * TEST INFRASTRUCTURE ONLY (see sgd_oracle.h).  Plain C99 + optional OpenMP.  Every function cites
 * the file:line of /root/reference/c_code it restates.  Nothing here is copied from the reference:
 * the reference is Thrust/cuBLAS/CUDA code, this is a scalar CPU description of the same maths.
 *
 * threads == 1  : the scalar oracle used by the parity tests (deterministic summation order:
 *                 ascending batch position, which is the order of sgd_thrust.cu:266-268).
 * threads  > 1  : the "OpenMP transliteration of the plain-CUDA codepath" CPU baseline
 *                 (sgd_thrust.cu:140-193: one worker per batch row; :198-213: reduction).
 */
#include "sgd_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * glibc rand(): main.cu:90 and main.cu:207 call std::random_shuffle without a generator, which
 * libstdc++ implements with rand(); the program never calls srand(), so the stream is srand(1).
 * glibc's default generator (TYPE_3) is an additive lagged-Fibonacci generator with lags 31 and 3
 * over uint32, warmed up from a Lehmer LCG (16807 mod 2^31-1) and with the first 310 outputs
 * discarded; rand() returns the top 31 bits.  (glibc 2.39 stdlib/random_r.c; pinned by
 * tests/golden/thirdparty_kat.json which was produced by the real rand().)
 * ---------------------------------------------------------------------------------------------- */
void oracle_rand_init(oracle_rand_t* g, uint32_t seed) {
  if (seed == 0) seed = 1;
  int32_t word = (int32_t)seed;
  g->st[0] = (uint32_t)word;
  for (int i = 1; i < 31; ++i) {
    /* 16807 * word mod (2^31 - 1) without overflow (Schrage) */
    int64_t hi = word / 127773;
    int64_t lo = word % 127773;
    int64_t t = 16807 * lo - 2836 * hi;
    if (t < 0) t += 2147483647;
    word = (int32_t)t;
    g->st[i] = (uint32_t)word;
  }
  for (int i = 31; i < 34; ++i) g->st[i] = g->st[i - 31];
  g->n = 34;
  for (int i = 0; i < 310; ++i) {
    uint32_t v = g->st[(g->n - 31) % 34] + g->st[(g->n - 3) % 34];
    g->st[g->n % 34] = v;
    g->n++;
  }
}

int32_t oracle_rand_next(oracle_rand_t* g) {
  uint32_t v = g->st[(g->n - 31) % 34] + g->st[(g->n - 3) % 34];
  g->st[g->n % 34] = v;
  g->n++;
  return (int32_t)(v >> 1);
}

/* ------------------------------------------------------------------------------------------------
 * std::random_shuffle(first, last), libstdc++ 13 bits/stl_algo.h (the two-iterator overload):
 *   for (i = first + 1; i != last; ++i) { j = first + rand() % ((i - first) + 1); if (i != j) swap }
 * Applied cumulatively: the index vector is initialised once (main.cu:421-424) and never reset.
 * ---------------------------------------------------------------------------------------------- */
void oracle_random_shuffle(oracle_rand_t* g, int32_t* ind, int64_t n) {
  for (int64_t i = 1; i < n; ++i) {
    int64_t j = (int64_t)(oracle_rand_next(g) % (int32_t)(i + 1));
    if (i != j) {
      int32_t t = ind[i];
      ind[i] = ind[j];
      ind[j] = t;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * main.cu:389-395: thrust::default_random_engine (= minstd_rand: x <- 48271 x mod 2^31-1, seed 1)
 * through thrust::uniform_real_distribution<float>(0.0, 0.01):
 *   u = float(x - min) / (1.0f + float(max - min)), min = 1, max = 2^31 - 2;  w = u * (b - a) + a
 * ---------------------------------------------------------------------------------------------- */
void oracle_init_weights(float* w, int32_t C) {
  uint64_t x = 1;
  const float denom = 1.0f + (float)2147483645u;
  const float a = 0.0f, b = 0.01f;
  for (int32_t c = 0; c < C; ++c) {
    x = (48271ull * x) % 2147483647ull;
    float u = (float)(uint32_t)(x - 1);
    u /= denom;
    w[c] = (u * (b - a)) + a;
  }
}

/* main.cu:343 : batchsize 0 means "the whole data set" */
int32_t oracle_effective_batch(int32_t R, int32_t B_arg) { return B_arg == 0 ? R : B_arg; }

/* main.cu:344 : (int)std::floor(R / (float)batchsize) -- a FLOAT division */
int32_t oracle_num_batches(int32_t R, int32_t B) { return (int32_t)floorf((float)R / (float)B); }

/* main.cu:144 / main.cu:282 : std::pow(int, double) is evaluated in double, lr is promoted,
 * the negated quotient is narrowed to float. */
float oracle_step_scale(float lr, int32_t epoch) {
  return (float)(-((double)lr / pow((double)epoch, 0.25)));
}

/* sgd_thrust.cu:43-48 (plain path) / :340-364 (Scopy + Sgemv with alpha=1, beta=-1):
 * r[i] = sum_c X[rows[i],c] * w[c] - y[rows[i]], fp32, ascending c. */
void oracle_residuals(const float* X, const float* y, const float* w, const int32_t* rows,
                      int32_t nrows, int32_t C, float* r_out) {
  for (int32_t i = 0; i < nrows; ++i) {
    const int64_t row = rows ? rows[i] : i;
    const float* x = X + row * (int64_t)C;
    float pred = 0.0f;
    for (int32_t c = 0; c < C; ++c) pred += x[c] * w[c];
    r_out[i] = pred - y[row];
  }
}

/* sgd_thrust.cu:301-315 (G[i,c] = X_b[i,c] * r[i], rounded to fp32) followed by
 * sgd_thrust.cu:262-269 (sum over i ascending, then / scaling_factor). */
void oracle_gradient(const float* X, const float* r, const int32_t* rows, int32_t nrows, int32_t C,
                     float denom, float* g_out) {
  for (int32_t c = 0; c < C; ++c) g_out[c] = 0.0f;
  for (int32_t i = 0; i < nrows; ++i) {
    const int64_t row = rows ? rows[i] : i;
    const float* x = X + row * (int64_t)C;
    for (int32_t c = 0; c < C; ++c) {
      volatile float prod = x[c] * r[i]; /* the product is stored as fp32 before it is summed */
      g_out[c] += prod;
    }
  }
  for (int32_t c = 0; c < C; ++c) g_out[c] = g_out[c] / denom;
}

static void step_serial(const float* X, const float* y, int32_t C, int32_t B, const int32_t* rows,
                        float* w, float a, int acc64, float* r, float* g, double* g64) {
  if (!acc64) {
    oracle_residuals(X, y, w, rows, B, C, r);
    oracle_gradient(X, r, rows, B, C, (float)B, g);
  } else {
    for (int32_t c = 0; c < C; ++c) g64[c] = 0.0;
    for (int32_t i = 0; i < B; ++i) {
      const float* x = X + (int64_t)rows[i] * C;
      double pred = 0.0;
      for (int32_t c = 0; c < C; ++c) pred += (double)x[c] * (double)w[c];
      double res = pred - (double)y[rows[i]];
      for (int32_t c = 0; c < C; ++c) g64[c] += res * (double)x[c];
    }
    for (int32_t c = 0; c < C; ++c) g[c] = (float)(g64[c] / (double)B);
  }
  /* main.cu:147-150 / :287-290 : w = a * g + w */
  for (int32_t c = 0; c < C; ++c) w[c] = a * g[c] + w[c];
}

/* main.cu:94-160 (plain) / main.cu:227-300 (cuBLAS): the batch loop of one epoch. */
void oracle_epoch(const float* X, const float* y, int64_t R, int32_t C, int32_t B,
                  const int32_t* ind, float* w, float lr, int32_t epoch, int acc64, int threads) {
  const int32_t nb = oracle_num_batches((int32_t)R, B);
  const float a = oracle_step_scale(lr, epoch);
  float* g = (float*)malloc(sizeof(float) * (size_t)C);
  if (threads <= 1) {
    float* r = (float*)malloc(sizeof(float) * (size_t)B);
    double* g64 = acc64 ? (double*)malloc(sizeof(double) * (size_t)C) : NULL;
    for (int32_t b = 0; b < nb; ++b)
      step_serial(X, y, C, B, ind + (int64_t)b * B, w, a, acc64, r, g, g64);
    free(r);
    free(g64);
    free(g);
    return;
  }
#ifdef _OPENMP
  /* CPU baseline: one worker per batch row (sgd_thrust.cu:152-158), per-thread partial gradient
   * (what K1's transposed gradient matrix + reduce_by_key, sgd_thrust.cu:198-213, amount to). */
  float* part = (float*)calloc((size_t)threads * (size_t)C, sizeof(float));
#pragma omp parallel num_threads(threads)
  {
    const int tid = omp_get_thread_num();
    const int nt = omp_get_num_threads();
    float* mine = part + (size_t)tid * C;
    for (int32_t b = 0; b < nb; ++b) {
      const int32_t* rows = ind + (int64_t)b * B;
      for (int32_t c = 0; c < C; ++c) mine[c] = 0.0f;
#pragma omp for schedule(static) nowait
      for (int32_t i = 0; i < B; ++i) {
        const float* x = X + (int64_t)rows[i] * C;
        float pred = 0.0f;
        for (int32_t c = 0; c < C; ++c) pred += x[c] * w[c];
        const float res = pred - y[rows[i]];
        for (int32_t c = 0; c < C; ++c) mine[c] += res * x[c];
      }
#pragma omp barrier
#pragma omp for schedule(static)
      for (int32_t c = 0; c < C; ++c) {
        float s = 0.0f;
        for (int t = 0; t < nt; ++t) s += part[(size_t)t * C + c];
        w[c] = a * (s / (float)B) + w[c];
      }
      /* implicit barrier of the omp for: w is complete before the next batch reads it */
    }
  }
  free(part);
#else
  (void)threads;
  float* r = (float*)malloc(sizeof(float) * (size_t)B);
  for (int32_t b = 0; b < nb; ++b) step_serial(X, y, C, B, ind + (int64_t)b * B, w, a, 0, r, g, NULL);
  free(r);
#endif
  free(g);
}

/* New functionality (BASELINE.json north_star): row-sharded data parallel step.
 * rank k owns rows [k*rpr, (k+1)*rpr), rpr = ceil(R/nranks).  Per step every rank sums
 * r_i * x_i over the batch rows it owns (batch order kept), partials are added in rank order,
 * then the same update as the single-rank path. */
void oracle_epoch_sharded(const float* X, const float* y, int64_t R, int32_t C, int32_t B,
                          const int32_t* ind, float* w, float lr, int32_t epoch, int32_t nranks) {
  const int32_t nb = oracle_num_batches((int32_t)R, B);
  const float a = oracle_step_scale(lr, epoch);
  const int64_t rpr = (R + nranks - 1) / nranks;
  float* part = (float*)malloc(sizeof(float) * (size_t)C * (size_t)nranks);
  for (int32_t b = 0; b < nb; ++b) {
    const int32_t* rows = ind + (int64_t)b * B;
    memset(part, 0, sizeof(float) * (size_t)C * (size_t)nranks);
    for (int32_t i = 0; i < B; ++i) {
      const int64_t row = rows[i];
      const float* x = X + row * C;
      float* mine = part + (size_t)(row / rpr) * C;
      float pred = 0.0f;
      for (int32_t c = 0; c < C; ++c) pred += x[c] * w[c];
      const float res = pred - y[row];
      for (int32_t c = 0; c < C; ++c) mine[c] += res * x[c];
    }
    for (int32_t c = 0; c < C; ++c) {
      float s = 0.0f;
      for (int32_t k = 0; k < nranks; ++k) s += part[(size_t)k * C + c];
      w[c] = a * (s / (float)B) + w[c];
    }
  }
  free(part);
}

/* sgd_thrust.cu:79-135 : Scopy(labels) ; Sgemv(alpha=1, beta=-1) ; Sasum ; / R.
 * The residual is fp32 as in the reference; the absolute values are summed in double because
 * cublasSasum's internal order is unspecified and a sequential fp32 sum of R ~ 1e7 terms would be
 * the LESS faithful stand-in (the parity tolerance for the MAE is 1e-4 relative). */
float oracle_mean_abs_error(const float* X, const float* y, const float* w, int64_t R, int32_t C,
                            int threads) {
  double total = 0.0;
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : total) num_threads(threads > 0 ? threads : 1) schedule(static)
#else
  (void)threads;
#endif
  for (int64_t i = 0; i < R; ++i) {
    const float* x = X + i * C;
    float pred = 0.0f;
    for (int32_t c = 0; c < C; ++c) pred += x[c] * w[c];
    total += fabs((double)(pred - y[i]));
  }
  return (float)(total / (double)R);
}

/* main.cu:338-530 without the I/O. */
float oracle_run(const float* X, const float* y, int64_t R, int32_t C, int32_t B_arg, float lr,
                 int32_t epochs, int acc64, int threads, float* w_out, int32_t* ind_out) {
  const int32_t B = oracle_effective_batch((int32_t)R, B_arg);
  float* w = (float*)malloc(sizeof(float) * (size_t)C);
  int32_t* ind = (int32_t*)malloc(sizeof(int32_t) * (size_t)R);
  oracle_init_weights(w, C);                     /* main.cu:389-395 */
  for (int64_t i = 0; i < R; ++i) ind[i] = (int32_t)i; /* main.cu:421-424 */
  oracle_rand_t g;
  oracle_rand_init(&g, 1);
  for (int32_t e = 1; e <= epochs; ++e) {        /* main.cu:456 */
    oracle_random_shuffle(&g, ind, R);           /* main.cu:90 / :207 */
    oracle_epoch(X, y, R, C, B, ind, w, lr, e, acc64, threads);
  }
  const float mae = oracle_mean_abs_error(X, y, w, R, C, threads); /* main.cu:525-530 */
  if (w_out) memcpy(w_out, w, sizeof(float) * (size_t)C);
  if (ind_out) memcpy(ind_out, ind, sizeof(int32_t) * (size_t)R);
  free(w);
  free(ind);
  return mae;
}

/* sgd_io.cu:38-77.  Line by line, cells split on ',', each cell through the same conversion
 * std::stof performs (strtof); column C is the label, columns < C are features. */
int oracle_read_csv(const char* path, int64_t R, int32_t C, float* X, float* y) {
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  char* line = NULL;
  size_t cap = 0;
  int64_t row = 0;
  int rc = 0;
  while (getline(&line, &cap, f) != -1) {
    char* p = line;
    int32_t col = 0;
    while (*p && *p != '\n' && *p != '\r') {
      char* end = p;
      float v = strtof(p, &end);
      if (end == p) break; /* std::stof would throw here */
      if (row >= R || col > C) { rc = -2; goto done; }
      if (col != C) X[row * C + col] = v; else y[row] = v;
      ++col;
      p = end;
      while (*p && *p != ',' && *p != '\n' && *p != '\r') ++p;
      if (*p == ',') ++p;
    }
    ++row;
  }
done:
  free(line);
  fclose(f);
  return rc;
}

/* sgd_thrust.cu:374-402 with copy_idx_func (sgd_thrust.cuh:122-134): Xs[i,:] = X[order[i],:] */
void oracle_permute_rows(const float* X, const float* y, const int32_t* order, int64_t R, int32_t C,
                         float* Xs, float* ys) {
  for (int64_t i = 0; i < R; ++i) {
    memcpy(Xs + i * C, X + (int64_t)order[i] * C, sizeof(float) * (size_t)C);
    ys[i] = y[order[i]];
  }
}

/* sgd_thrust.cu:301-315 */
void oracle_scale_rows(const float* X, const float* v, int64_t R, int32_t C, float* out) {
  for (int64_t i = 0; i < R; ++i)
    for (int32_t c = 0; c < C; ++c) out[i * C + c] = X[i * C + c] * v[i];
}

/* sgd_thrust.cu:216-251 (scale 1) and :254-291 (sum then divide by the scaling factor) */
void oracle_col_sums(const float* G, int64_t R, int32_t C, float scale, float* out) {
  for (int32_t c = 0; c < C; ++c) {
    float s = 0.0f;
    for (int64_t i = 0; i < R; ++i) s += G[i * C + c];
    out[c] = s / scale;
  }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
