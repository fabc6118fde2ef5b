// host_io.cpp -- host-side L1 layer of the reference behind the C-ABI:
//   read_csv                 (sgd_io.cu:38-77)   -> b200sgd::read_csv   (parallel, bounds-checked)
//   get_filename_from_path   (sgd_io.cu:107-113) -> b200sgd::filename_from_path
//   ExperimentOutput/write_json (sgd_io.cuh:28-60, sgd_io.cu:82-105) -> b200sgd::write_experiment_json
// No jsoncpp: the five-key object is emitted directly in the byte layout jsoncpp 0.10.5's
// StyledWriter produces (3-space indent, `"key" : value`, keys sorted, "%.17g", trailing newline).
#include "host_io.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace b200sgd {

namespace {

// One cell -> float with std::stof's semantics (strtof: leading blanks, optional sign, longest
// valid prefix, trailing junk ignored).  from_chars is the fast path when it consumes the WHOLE
// cell; anything else ("+2", " .5", "0x1p3", "inf", "1.5abc") goes through strtof on a NUL-terminated
// copy of the whole cell, so that the value agrees with stof.  Where stof would throw
// (std::invalid_argument: no number; std::out_of_range: strtof sets ERANGE) the cell is rejected.
inline bool parse_cell(const char* b, const char* e, float* out) {
  while (b < e && (*b == ' ' || *b == '\t')) ++b;
  if (b == e) return false;
  if (*b != '+') {
    auto res = std::from_chars(b, e, *out);
    if (res.ec == std::errc() && res.ptr == e) return true;
  }
  char tmp[64];
  std::string big;
  const size_t n = (size_t)(e - b);
  const char* z;
  if (n < sizeof(tmp)) {
    std::memcpy(tmp, b, n);
    tmp[n] = 0;
    z = tmp;
  } else {
    big.assign(b, n);
    z = big.c_str();
  }
  char* end = nullptr;
  errno = 0;
  const float v = std::strtof(z, &end);
  if (end == z) return false;       // std::stof would throw std::invalid_argument
  if (errno == ERANGE) return false;  // std::stof would throw std::out_of_range
  *out = v;
  return true;
}

// Parses the lines in [b, e) (whole lines) starting at row `row`.  Returns 0 or a negative code.
// Row r goes to X[r * px .. + C) and y[r * py] (px = C, py = 1: the reference's two arrays; px = py = C + 1
// with y = X + C: records `x..., label`, the layout of the raw cache and of the upload staging).
int parse_lines(const char* b, const char* e, int64_t row, int64_t R, int32_t C, float* X, float* y, int64_t px, int64_t py) {
  while (b < e) {
    const char* eol = (const char*)std::memchr(b, '\n', (size_t)(e - b));
    if (!eol) eol = e;
    const char* le = eol;
    if (le > b && le[-1] == '\r') --le;
    if (le > b) {
      if (row >= R) return -3;
      int32_t col = 0;
      const char* p = b;
      while (true) {
        const char* q = (const char*)std::memchr(p, ',', (size_t)(le - p));
        const char* ce = q ? q : le;
        float v;
        if (!parse_cell(p, ce, &v)) return -3;
        if (col > C) return -3;  // the reference would write into the next row here
        if (col != C) X[row * px + col] = v; else y[row * py] = v;
        ++col;
        if (!q) break;
        p = q + 1;
        if (p == le) break;  // trailing comma: getline yields no further cell
      }
    }  // a blank line only bumps the row counter, exactly as the reference's getline loop does
    ++row;
    b = eol + 1;
  }
  return 0;
}

}  // namespace

int read_csv(const char* path, int64_t R, int32_t C, float* X, float* y, std::string* err) {
  int fd = ::open(path, O_RDONLY);
  if (fd < 0) {
    if (err) *err = std::string("Could not open file: ") + path;
    return -2;
  }
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); if (err) *err = "fstat failed"; return -2; }
  const size_t len = (size_t)st.st_size;
  std::memset(X, 0, sizeof(float) * (size_t)R * (size_t)C);  // thrust::host_vector zero-fills
  std::memset(y, 0, sizeof(float) * (size_t)R);
  if (len == 0) { ::close(fd); return 0; }
  const char* base = (const char*)mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (base == MAP_FAILED) { if (err) *err = "mmap failed"; return -2; }
  madvise((void*)base, len, MADV_SEQUENTIAL);

  unsigned hw = std::thread::hardware_concurrency();
  size_t nchunks = len < (size_t)(4u << 20) ? 1 : (hw ? hw : 4);
  if (nchunks > 64) nchunks = 64;
  // chunk boundaries snapped to line starts
  std::vector<size_t> cut(nchunks + 1, len);
  cut[0] = 0;
  for (size_t k = 1; k < nchunks; ++k) {
    size_t pos = len / nchunks * k;
    if (pos < cut[k - 1]) pos = cut[k - 1];
    const char* nl = (const char*)std::memchr(base + pos, '\n', len - pos);
    cut[k] = nl ? (size_t)(nl - base) + 1 : len;
  }
  // rows (= newline count) per chunk -> first row of every chunk
  std::vector<int64_t> rows_in(nchunks, 0);
  {
    std::vector<std::thread> th;
    for (size_t k = 0; k < nchunks; ++k)
      th.emplace_back([&, k] {
        int64_t n = 0;
        const char* p = base + cut[k];
        const char* e = base + cut[k + 1];
        while (p < e) {
          const char* nl = (const char*)std::memchr(p, '\n', (size_t)(e - p));
          if (!nl) { ++n; break; }  // last line without EOL
          ++n;
          p = nl + 1;
        }
        rows_in[k] = n;
      });
    for (auto& t : th) t.join();
  }
  std::vector<int64_t> row0(nchunks, 0);
  for (size_t k = 1; k < nchunks; ++k) row0[k] = row0[k - 1] + rows_in[k - 1];
  std::atomic<int> rc{0};
  {
    std::vector<std::thread> th;
    for (size_t k = 0; k < nchunks; ++k)
      th.emplace_back([&, k] {
        int r = parse_lines(base + cut[k], base + cut[k + 1], row0[k], R, C, X, y, C, 1);
        if (r != 0) rc.store(r);
      });
    for (auto& t : th) t.join();
  }
  munmap((void*)base, len);
  if (rc.load() != 0 && err)
    *err = std::string("CSV has more rows/columns than announced or an unparsable cell: ") + path;
  return rc.load();
}

// ---- chunked access for the overlapped load (sgd_load_csv): the file is mapped once and cut into
// pieces of whole lines whose first row is known, so that any piece can be parsed on its own ----
CsvFile::~CsvFile() {
  if (base_) munmap((void*)base_, len_);
}

int CsvFile::open(const char* path, size_t piece_bytes, std::string* err) {
  int fd = ::open(path, O_RDONLY);
  if (fd < 0) {
    if (err) *err = std::string("Could not open file: ") + path;
    return -2;
  }
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); if (err) *err = "fstat failed"; return -2; }
  len_ = (size_t)st.st_size;
  size_ = (int64_t)st.st_size;
  mtime_ns_ = (int64_t)st.st_mtim.tv_sec * 1000000000ll + (int64_t)st.st_mtim.tv_nsec;
  if (len_ == 0) { ::close(fd); return 0; }
  base_ = (const char*)mmap(nullptr, len_, PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (base_ == MAP_FAILED) { base_ = nullptr; if (err) *err = "mmap failed"; return -2; }
  madvise((void*)base_, len_, MADV_SEQUENTIAL);
  const size_t np = std::max<size_t>(1, (len_ + piece_bytes - 1) / piece_bytes);
  std::vector<size_t> cut(np + 1, len_);
  cut[0] = 0;
  for (size_t k = 1; k < np; ++k) {
    size_t pos = std::max(cut[k - 1], len_ / np * k);
    const char* nl = (const char*)std::memchr(base_ + pos, '\n', len_ - pos);
    cut[k] = nl ? (size_t)(nl - base_) + 1 : len_;
  }
  pieces_.assign(np, Piece{});
  unsigned hw = std::thread::hardware_concurrency();
  const size_t T = std::min<size_t>(np, hw ? hw : 4);
  std::atomic<size_t> next{0};
  std::vector<std::thread> th;
  for (size_t t = 0; t < T; ++t)
    th.emplace_back([&] {
      for (size_t k = next.fetch_add(1); k < np; k = next.fetch_add(1)) {
        int64_t n = 0;
        const char* p = base_ + cut[k];
        const char* e = base_ + cut[k + 1];
        while (p < e) {
          const char* nl = (const char*)std::memchr(p, '\n', (size_t)(e - p));
          ++n;
          if (!nl) break;  // last line without EOL
          p = nl + 1;
        }
        pieces_[k].begin = cut[k];
        pieces_[k].end = cut[k + 1];
        pieces_[k].rows = n;
      }
    });
  for (auto& t : th) t.join();
  int64_t r = 0;
  for (auto& pc : pieces_) {
    pc.row0 = r;
    r += pc.rows;
  }
  return 0;
}

int CsvFile::parse_piece(size_t k, int64_t R, int32_t C, float* rec, int64_t rec_row0) const {
  const Piece& pc = pieces_[k];
  float* X = rec - rec_row0 * ((int64_t)C + 1);  // so that row r lands at rec + (r - rec_row0) * (C + 1)
  return parse_lines(base_ + pc.begin, base_ + pc.end, pc.row0, R, C, X, X + C, (int64_t)C + 1, (int64_t)C + 1);
}

// The reference's data_set_creation scripts write `f1,...,fC,label` lines with csv.writer; this is the
// same format with the shortest text that parses back to the same float (std::to_chars), written by
// all cores: the benchmark's reference arm and the ingestion measurements need GB-size files.
int write_csv(const char* path, int64_t R, int32_t C, const float* X, const float* y) {
  FILE* f = std::fopen(path, "wb");
  if (!f) return -2;
  unsigned hw = std::thread::hardware_concurrency();
  const size_t T = hw ? (hw > 32 ? 32 : hw) : 4;
  const int64_t rows_per_task = std::max<int64_t>(1, (int64_t)(4 << 20) / (12 * ((int64_t)C + 1)));  // ~4 MB of text
  std::vector<std::string> out(T);
  int rc = 0;
  for (int64_t r0 = 0; r0 < R && rc == 0; r0 += rows_per_task * (int64_t)T) {
    std::vector<std::thread> th;
    for (size_t k = 0; k < T; ++k)
      th.emplace_back([&, k] {
        std::string& s = out[k];
        s.clear();
        const int64_t b = r0 + (int64_t)k * rows_per_task, e = std::min(R, b + rows_per_task);
        char buf[32];
        for (int64_t r = b; r < e; ++r) {
          for (int32_t c = 0; c <= C; ++c) {
            const float v = c < C ? X[r * (int64_t)C + c] : y[r];
            auto res = std::to_chars(buf, buf + sizeof(buf), v);
            s.append(buf, (size_t)(res.ptr - buf));
            s.push_back(c < C ? ',' : '\n');
          }
        }
      });
    for (auto& t : th) t.join();
    for (size_t k = 0; k < T; ++k)
      if (!out[k].empty() && std::fwrite(out[k].data(), 1, out[k].size(), f) != out[k].size()) rc = -2;
  }
  if (std::fclose(f) != 0) rc = -2;
  return rc;
}

std::string filename_from_path(const std::string& filepath) {
  // sgd_io.cu:107-113: strip up to the last '/', then from the last '.'
  auto pos = filepath.find_last_of('/');
  std::string name = filepath.substr(pos + 1);  // npos + 1 == 0
  pos = name.find_last_of('.');
  return name.substr(0, pos);
}

static std::string json_number(float f) {
  // Json::Value(float) promotes to double; valueToString(double) prints "%.17g"
  const double v = (double)f;
  char buf[40];
  if (std::isfinite(v)) std::snprintf(buf, sizeof(buf), "%.17g", v);
  else if (v != v) std::snprintf(buf, sizeof(buf), "null");
  else std::snprintf(buf, sizeof(buf), v < 0 ? "-1e+9999" : "1e+9999");
  return buf;
}

std::string experiment_json(float shuffle_time, float total_gradient_time, float gpu_time,
                            float transfer_time, float error) {
  // keys in jsoncpp's (sorted) order
  std::string s = "{\n";
  s += "   \"error\" : " + json_number(error) + ",\n";
  s += "   \"gpu_time\" : " + json_number(gpu_time) + ",\n";
  s += "   \"shuffle_time\" : " + json_number(shuffle_time) + ",\n";
  s += "   \"total_gradient_time\" : " + json_number(total_gradient_time) + ",\n";
  s += "   \"transfer_time\" : " + json_number(transfer_time) + "\n";
  s += "}\n";
  return s;
}

int write_experiment_json(const char* path, float shuffle_time, float total_gradient_time,
                          float gpu_time, float transfer_time, float error) {
  FILE* f = std::fopen(path, "w");
  if (!f) return -2;
  const std::string s = experiment_json(shuffle_time, total_gradient_time, gpu_time, transfer_time, error);
  const size_t n = std::fwrite(s.data(), 1, s.size(), f);
  std::fclose(f);
  return n == s.size() ? 0 : -2;
}

}  // namespace b200sgd
