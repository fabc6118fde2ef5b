// host_io.hpp -- see host_io.cpp
#pragma once
#include <cstdint>
#include <string>

namespace b200sgd {

// 0 ok, -2 cannot open (reference: read_csv returns false), -3 data does not fit R x (C+1)
int read_csv(const char* path, int64_t R, int32_t C, float* X, float* y, std::string* err);
// `f1,...,fC,label` lines, shortest round-trip text; 0 ok, -2 cannot write
int write_csv(const char* path, int64_t R, int32_t C, const float* X, const float* y);
std::string filename_from_path(const std::string& filepath);
std::string experiment_json(float shuffle_time, float total_gradient_time, float gpu_time,
                            float transfer_time, float error);
int write_experiment_json(const char* path, float shuffle_time, float total_gradient_time,
                          float gpu_time, float transfer_time, float error);

}  // namespace b200sgd
