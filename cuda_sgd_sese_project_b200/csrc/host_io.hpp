// host_io.hpp -- see host_io.cpp
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace b200sgd {

// 0 ok, -2 cannot open (reference: read_csv returns false), -3 data does not fit R x (C+1)
int read_csv(const char* path, int64_t R, int32_t C, float* X, float* y, std::string* err);
// A CSV mapped into memory and cut into pieces of whole lines with known first rows (sgd_load_csv
// parses piece k+1 while piece k is on its way to the GPU).
class CsvFile {
 public:
  struct Piece { size_t begin = 0, end = 0; int64_t row0 = 0, rows = 0; };
  ~CsvFile();
  int open(const char* path, size_t piece_bytes, std::string* err);  // 0 ok, -2 cannot open
  const std::vector<Piece>& pieces() const { return pieces_; }
  int64_t size() const { return size_; }
  int64_t mtime_ns() const { return mtime_ns_; }
  // rows of piece k as records `x_0..x_{C-1}, label` (C + 1 floats each) into rec, whose first record is
  // row rec_row0 of the file; 0 ok, -3 more rows / columns than announced or an unparsable cell
  int parse_piece(size_t k, int64_t R, int32_t C, float* rec, int64_t rec_row0) const;
 private:
  const char* base_ = nullptr;
  size_t len_ = 0;
  int64_t size_ = 0, mtime_ns_ = 0;
  std::vector<Piece> pieces_;
};

// `f1,...,fC,label` lines, shortest round-trip text; 0 ok, -2 cannot write
int write_csv(const char* path, int64_t R, int32_t C, const float* X, const float* y);
std::string filename_from_path(const std::string& filepath);
std::string experiment_json(float shuffle_time, float total_gradient_time, float gpu_time,
                            float transfer_time, float error);
int write_experiment_json(const char* path, float shuffle_time, float total_gradient_time,
                          float gpu_time, float transfer_time, float error);

}  // namespace b200sgd
