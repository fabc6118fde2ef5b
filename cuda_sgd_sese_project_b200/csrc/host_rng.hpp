// host_rng.hpp -- the two host-side random streams the reference's results depend on, written for
// throughput (the per-epoch shuffle is the reference's real bottleneck at scale).
//
//  * GlibcRand      : the stream std::random_shuffle consumes at main.cu:90 / main.cu:207.  The
//                     reference never calls srand(), so it is glibc's default generator with seed
//                     1: an additive lagged-Fibonacci sequence o[k] = o[k-31] + o[k-3] (mod 2^32),
//                     rand() = o[k] >> 1.  Generated here in bulk, lock-free, ~1 ns per value
//                     (glibc's rand() takes a lock per call).
//  * exact_shuffle  : libstdc++'s std::random_shuffle(first,last):
//                         for i in 1..n-1: j = rand() % (i+1); if (i != j) swap(v[i], v[j])
//                     with the swap targets computed a block ahead and prefetched.
//  * thrust_minstd_weights : main.cu:389-395.
//
// Product code: must stay independent of oracle/ (which restates the same streams for checking).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace b200sgd {

class GlibcRand {
 public:
  explicit GlibcRand(uint32_t seed = 1) { reseed(seed); }

  void reseed(uint32_t seed) {
    if (seed == 0) seed = 1;  // glibc maps 0 to 1
    int32_t word = (int32_t)seed;
    uint32_t r[34];
    r[0] = (uint32_t)word;
    for (int i = 1; i < 31; ++i) {
      // word = 16807 * word mod (2^31 - 1), Schrage's decomposition as glibc does it
      const int64_t hi = word / 127773, lo = word % 127773;
      int64_t t = 16807 * lo - 2836 * hi;
      if (t < 0) t += 2147483647;
      word = (int32_t)t;
      r[i] = (uint32_t)word;
    }
    for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
    // keep the last 31 values; hist_[k] is o[n-31+k]
    std::memcpy(hist_, r + 3, sizeof(hist_));
    uint32_t sink[310];
    raw(sink, 310);  // glibc discards the first 310 outputs
  }

  // next n raw 32-bit sums (rand() value = raw >> 1)
  void raw(uint32_t* out, size_t n) {
    // first 31 outputs depend on history, the rest on earlier outputs of this call
    size_t head = n < 31 ? n : 31;
    for (size_t k = 0; k < head; ++k) {
      const uint32_t a = hist_[k];  // o[k-31]
      const uint32_t b = k >= 3 ? out[k - 3] : hist_[28 + k];
      out[k] = a + b;
    }
    for (size_t k = 31; k < n; ++k) out[k] = out[k - 31] + out[k - 3];
    // new history = last 31 outputs
    if (n >= 31) {
      std::memcpy(hist_, out + n - 31, sizeof(hist_));
    } else {
      uint32_t tmp[31];
      std::memcpy(tmp, hist_ + n, (31 - n) * sizeof(uint32_t));
      std::memcpy(tmp + (31 - n), out, n * sizeof(uint32_t));
      std::memcpy(hist_, tmp, sizeof(hist_));
    }
  }

  int32_t next() {
    uint32_t v;
    raw(&v, 1);
    return (int32_t)(v >> 1);
  }

  // ---- jump-ahead ------------------------------------------------------------------------------
  // The stream is a linear recurrence x[k] = x[k-31] + x[k-3] over Z/2^32, characteristic
  // polynomial p(z) = z^31 - z^28 - 1.  If z^L = sum_j a_j z^j (mod p) then x[m+L] = sum_j a_j
  // x[m+j] for every m, so the generator can be moved L steps ahead with 31 dot products instead
  // of L additions.  Used to hand the device one start state per chunk of the stream (the stream
  // itself is then generated on the GPU, shuffle_kernels.cuh) and to keep this generator in step.
  struct Poly {
    uint32_t a[31];
  };
  static Poly poly_mul(const Poly& f, const Poly& g) {
    uint32_t prod[61] = {0};
    for (int i = 0; i < 31; ++i)
      for (int j = 0; j < 31; ++j) prod[i + j] += f.a[i] * g.a[j];
    for (int d = 60; d >= 31; --d) {  // z^d = z^(d-3) + z^(d-31)
      const uint32_t c = prod[d];
      prod[d - 3] += c;
      prod[d - 31] += c;
    }
    Poly r;
    for (int i = 0; i < 31; ++i) r.a[i] = prod[i];
    return r;
  }
  static Poly poly_z_pow(uint64_t n) {  // z^n mod p
    Poly result{}, base{};
    result.a[0] = 1;
    base.a[1] = 1;
    while (n) {
      if (n & 1) result = poly_mul(result, base);
      base = poly_mul(base, base);
      n >>= 1;
    }
    return result;
  }
  // state = the last 31 outputs; returns the state `jump` (= the polynomial's exponent) steps later
  static void apply(const Poly& jump, const uint32_t* state, uint32_t* out) {
    uint32_t ext[61];
    for (int i = 0; i < 31; ++i) ext[i] = state[i];
    for (int i = 31; i < 61; ++i) ext[i] = ext[i - 31] + ext[i - 3];
    for (int i = 0; i < 31; ++i) {
      uint32_t acc = 0;
      for (int j = 0; j < 31; ++j) acc += jump.a[j] * ext[i + j];
      out[i] = acc;
    }
  }
  void skip(uint64_t count) {
    if (count == 0) return;
    const Poly pz = poly_z_pow(count);
    uint32_t nxt[31];
    apply(pz, hist_, nxt);
    std::memcpy(hist_, nxt, sizeof(hist_));
  }
  // start states (31 words each) of `nchunks` consecutive chunks of `chunk` values, beginning at the
  // current position; the generator itself does not move
  void chunk_states(const Poly& jump_chunk, int64_t nchunks, uint32_t* out) const {
    if (nchunks <= 0) return;
    std::memcpy(out, hist_, sizeof(hist_));
    for (int64_t c = 1; c < nchunks; ++c) apply(jump_chunk, out + (c - 1) * 31, out + c * 31);
  }
  const uint32_t* state() const { return hist_; }

 private:
  uint32_t hist_[31];
};

// One std::random_shuffle pass over ind[0..n) consuming exactly n-1 values of `g`.
inline void exact_shuffle(GlibcRand& g, int32_t* ind, int64_t n) {
  constexpr int64_t BLK = 1024;
  uint32_t raw[BLK];
  uint32_t tgt[2][BLK];
  if (n < 2) return;
  // software pipeline: targets of block b+1 are computed and prefetched before block b is swapped
  auto prepare = [&](int64_t i0, int64_t cnt, uint32_t* t) {
    g.raw(raw, (size_t)cnt);
    for (int64_t k = 0; k < cnt; ++k) {
      const uint32_t j = (raw[k] >> 1) % (uint32_t)(i0 + k + 1);
      t[k] = j;
      __builtin_prefetch(ind + j, 1, 1);
    }
  };
  int64_t i0 = 1;
  int cur = 0;
  int64_t cnt = (n - i0) < BLK ? (n - i0) : BLK;
  prepare(i0, cnt, tgt[cur]);
  while (cnt > 0) {
    const int64_t ni0 = i0 + cnt;
    const int64_t ncnt = (n - ni0) < BLK ? (n - ni0) : BLK;
    if (ncnt > 0) prepare(ni0, ncnt, tgt[cur ^ 1]);
    const uint32_t* t = tgt[cur];
    for (int64_t k = 0; k < cnt; ++k) {
      const int64_t i = i0 + k;
      const int64_t j = t[k];
      const int32_t vi = ind[i], vj = ind[j];  // i == j swaps with itself: no-op, as `if (i != j)`
      ind[i] = vj;
      ind[j] = vi;
    }
    i0 = ni0;
    cnt = ncnt;
    cur ^= 1;
  }
}

// main.cu:389-395: thrust::default_random_engine (minstd_rand, a = 48271, m = 2^31-1, seed 1)
// through thrust::uniform_real_distribution<float>(0.0, 0.01).
inline void thrust_minstd_weights(float* w, int32_t cols) {
  uint64_t x = 1;
  const float denom = 1.0f + (float)2147483645u;  // 1 + float(max - min)
  const float lo = 0.0f, hi = 0.01f;
  for (int32_t c = 0; c < cols; ++c) {
    x = (48271ull * x) % 2147483647ull;
    float u = (float)(uint32_t)(x - 1u);  // urng() - min
    u /= denom;
    w[c] = (u * (hi - lo)) + lo;
  }
}

}  // namespace b200sgd
