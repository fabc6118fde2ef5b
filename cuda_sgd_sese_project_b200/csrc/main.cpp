// main.cpp -- the reference's command line (c_code/main.cu:303-571) over libb200sgd's C-ABI.
//
//   ./main [learning_rate] [iterations] [data_csv_file] [num_rows] [num_features] [batchsize] [cudaRun]
//
// Same seven positionals, same meaning (main.cu:303-314, :338-345): batchsize 0 = whole data set,
// cudaRun 1 = "plain CUDA" label / -cuda.json, 0 = "cuBLAS" label / -cublas.json.  All selectors run
// the same fused sm_100a kernel.  Same stdout lines (main.cu:356,373,386,517-523,536-553,562) and
// the same five-key JSON named <csv-basename>-{cublas,cuda}.json in the working directory, so
// experiments/plot_parametrized.py keeps working.  Deviations: the timing
// fields hold TRUE values (the reference re-adds its accumulators every batch, main.cu:154-159,
// :294-299) and an unreadable CSV exits with 2 instead of 0 (main.cu:366-370).
//
// Optional trailing arguments (not in the reference): --synthetic=SEED  --fp16  --engine=graph
// --grid=N  --reduce=allgather|scatter  --dump-weights=FILE  --device=N  --per-epoch
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "b200sgd.h"

namespace {
double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
bool starts_with(const char* s, const char* p) { return std::strncmp(s, p, std::strlen(p)) == 0; }
}  // namespace

int main(int argc, char** argv) {
  int npos = 0;
  std::vector<const char*> pos, opts;
  for (int i = 1; i < argc; ++i) {
    if (starts_with(argv[i], "--")) opts.push_back(argv[i]);
    else pos.push_back(argv[i]);
  }
  npos = (int)pos.size();
  if (npos != 7) {  // main.cu:332-336
    std::cout << "usage: ./sgd_thrust.o [learning_rate] "
                 "[iterations] [data_csv_file] [num_rows] [num_features] [batchsize] [cudaRun]"
              << std::endl;
    return 1;
  }
  const float learning_rate = (float)atof(pos[0]);  // main.cu:338
  const int MAX_EPOCHS = atoi(pos[1]);
  const std::string filepath = pos[2];
  const int R = atoi(pos[3]);
  const int C = atoi(pos[4]);
  const int batch_arg = atoi(pos[5]);
  const bool cudaRun = atoi(pos[6]) == 1;  // main.cu:345

  bool synthetic = false, fp16 = false, per_epoch = false, use_cache = false;
  unsigned long long seed = 42;
  std::string dump_weights;
  sgd_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.rows = R;
  cfg.cols = C;
  cfg.batch = batch_arg;
  cfg.learning_rate = learning_rate;
  cfg.codepath = cudaRun ? SGD_PATH_PLAIN_CUDA : SGD_PATH_CUBLAS;
  cfg.nranks = 1;
  for (const char* o : opts) {
    if (starts_with(o, "--synthetic=")) { synthetic = true; seed = strtoull(o + 12, nullptr, 10); }
    else if (!std::strcmp(o, "--fp16")) fp16 = true;
    else if (!std::strcmp(o, "--per-epoch")) per_epoch = true;
    else if (!std::strcmp(o, "--cache")) use_cache = true;
    else if (!std::strcmp(o, "--engine=graph")) cfg.engine = SGD_ENGINE_GRAPH;
    else if (!std::strcmp(o, "--engine=persistent")) cfg.engine = SGD_ENGINE_PERSISTENT;
    else if (starts_with(o, "--grid=")) cfg.grid = atoi(o + 7);
    else if (!std::strcmp(o, "--reduce=allgather")) cfg.reduce = SGD_REDUCE_ALLGATHER;
    else if (!std::strcmp(o, "--reduce=scatter")) cfg.reduce = SGD_REDUCE_SCATTER;
    else if (starts_with(o, "--dump-weights=")) dump_weights = o + 15;
    else if (starts_with(o, "--device=")) cfg.device = atoi(o + 9);
    else { std::cerr << "unknown option " << o << std::endl; return 1; }
  }

  const double t_mem0 = now_ms();
  std::cout << "Data size: ~" << float(R) * C * sizeof(float) / (1024 * 1024) << "MB" << std::endl;  // main.cu:356

  // main.cu:365 (read_csv) + :381-384 (device vectors): one overlapped pass here -- all host cores parse
  // the next piece of the file into pinned memory while the previous one is copied and packed on the
  // device (sgd_load_csv); `--cache` keeps the parsed rows in <csv>.f32 for the next run
  if (!synthetic) {  // main.cu:366-370: an unreadable file ends the run before anything is allocated
    FILE* f = std::fopen(filepath.c_str(), "rb");
    if (!f) {
      std::cout << "Warning: Could not open file: " << filepath << std::endl;
      std::cout << "Could not read file, exiting..." << std::endl;  // main.cu:368
      return 2;
    }
    std::fclose(f);
  }
  const double t_copy0 = now_ms();
  sgd_ctx* ctx = nullptr;
  int rc = sgd_create(&cfg, nullptr, nullptr, &ctx);
  if (rc != SGD_OK) {
    std::cerr << "b200sgd: " << sgd_last_error() << std::endl;
    return 3;
  }
  double csv_time = 0.0;
  if (!synthetic) {
    sgd_load_stats st;
    std::memset(&st, 0, sizeof(st));
    rc = sgd_load_csv(ctx, filepath.c_str(), use_cache ? SGD_CSV_CACHE : 0, &st);
    if (rc == SGD_ERR_IO || rc == SGD_ERR_PARSE) {
      std::cout << "Warning: " << sgd_last_error() << std::endl;
      std::cout << "Could not read file, exiting..." << std::endl;  // main.cu:368
      sgd_destroy(ctx);
      return 2;
    }
    csv_time = st.parse_seconds * 1e3;
  } else {
    rc = sgd_load_synthetic(ctx, seed, fp16 ? 1 : 0);
  }
  printf("%s time = %f ms\n", "csv read", csv_time);  // main.cu:373 (the parser's share of the overlapped load)
  if (rc != SGD_OK) {
    std::cerr << "b200sgd: " << sgd_last_error() << std::endl;
    sgd_destroy(ctx);
    return 3;
  }
  printf("%s time = %f ms\n", "GPU copy", now_ms() - t_copy0);  // main.cu:386
  const float transfer_time = (float)(now_ms() - t_mem0);        // main.cu:436-441

  if (cudaRun) {
#ifdef LVL1
    std::cout << "NOTE: Using cuBLAS dynamic parallelism!" << std::endl;  // main.cu:451-455 (label only)
#endif
  }

  // main.cu:456-508: the epoch loop.  gpu_time spans the whole loop including the per-epoch
  // shuffles, exactly as the reference's event pair main.cu:444-445 / :511-514 does.
  sgd_timings tot;
  std::memset(&tot, 0, sizeof(tot));
  const double t_gpu0 = now_ms();
  if (per_epoch) {
    for (int epoch = 1; epoch <= MAX_EPOCHS && rc == SGD_OK; ++epoch) {
      sgd_timings t;
      rc = sgd_iteration(ctx, epoch, &t);  // one cuda_iteration / cublas_iteration call
      tot.shuffle_ms += t.shuffle_ms;
      tot.steploop_ms += t.steploop_ms;
      tot.steps += t.steps;
      tot.samples += t.samples;
      tot.bytes += t.bytes;
      tot.launches += t.launches;
    }
  } else {
    rc = sgd_run(ctx, 1, MAX_EPOCHS, &tot);
  }
  if (rc != SGD_OK) {
    std::cerr << "b200sgd: " << sgd_last_error() << std::endl;
    sgd_destroy(ctx);
    return 3;
  }
  const float gpu_time = (float)(now_ms() - t_gpu0);

  if (!cudaRun) printf("\nCUBLAS execution: \n");  // main.cu:516-520
  else printf("\nPlain CUDA execution: \n");
  printf("Data transfer time = %f ms\n", transfer_time);
  printf("GPU time = %f ms\n", gpu_time);

  float avg_abs_error = 0.f;  // main.cu:525-530
  rc = sgd_mean_abs_error(ctx, &avg_abs_error, nullptr);
  if (rc != SGD_OK) {
    std::cerr << "b200sgd: " << sgd_last_error() << std::endl;
    sgd_destroy(ctx);
    return 3;
  }

  char base[4096];
  sgd_get_filename_from_path(filepath.c_str(), base, sizeof(base));
  const std::string filename = base;
  printf("Total gradient time = %f ms\n", tot.steploop_ms);  // main.cu:536 / :547
  printf("Shuffle time = %f ms\n", tot.shuffle_ms);
  if (!cudaRun) {
    printf("\n CUBLAS-specific timings: \n");  // one fused kernel: no separate phases to report
    printf("Derivative time = %f ms\n", 0.0);
    printf("Matrix scale time = %f ms\n", 0.0);
    printf("Col sum/scale time = %f ms\n", 0.0);
    printf("saxpy time = %f ms\n", 0.0);
  } else {
    printf("\n Plain-CUDA-specific timings: \n");
    printf("gradient matrix time = %f ms\n", 0.0);
    printf("gradient vec. scale time = %f ms\n", 0.0);
    printf("row sum time = %f ms\n", 0.0);
    printf("saxpy time = %f ms\n", 0.0);
  }
  const std::string json = filename + (cudaRun ? "-cuda.json" : "-cublas.json");  // main.cu:545 / :557
  if (sgd_write_experiment_output(json.c_str(), tot.shuffle_ms, tot.steploop_ms, gpu_time, transfer_time,
                                  avg_abs_error) != SGD_OK)
    std::cerr << "b200sgd: " << sgd_last_error() << std::endl;

  // additions of this engine (the reference prints nothing comparable)
  if (tot.steploop_ms > 0.f) {
    printf("\n B200 fused-step engine: \n");
    printf("steps = %d, kernel launches = %d\n", tot.steps, tot.launches);
    printf("samples/s (epoch loop incl. shuffle) = %.6g\n", (double)tot.samples / (gpu_time * 1e-3));
    printf("samples/s (step loop only) = %.6g\n", (double)tot.samples / (tot.steploop_ms * 1e-3));
    printf("algorithmic HBM GB/s (step loop) = %.6g\n", (double)tot.bytes / (tot.steploop_ms * 1e-3) * 1e-9);
  }

  if (!dump_weights.empty()) {  // what the commented-out print_vector at main.cu:560 would show
    std::vector<float> w((size_t)C);
    if (sgd_get_weights(ctx, w.data()) == SGD_OK) {
      FILE* f = std::fopen(dump_weights.c_str(), "w");
      if (f) {
        for (int i = 0; i < C; ++i) std::fprintf(f, "%.9g\n", w[i]);
        std::fclose(f);
      }
    }
  }

  std::cout << "Mean absolute error : " << avg_abs_error << std::endl;  // main.cu:562
  sgd_destroy(ctx);
  return 0;
}
