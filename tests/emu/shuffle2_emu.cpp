// shuffle2_emu.cpp -- runs the kernels of csrc/shuffle2_kernels.cuh on the host (cuda_emu.hpp) through the
// same sequence of launches as b200sgd_api.cu and compares every epoch's permutation (and the position
// map of the multi-GPU mode) with the host loop of host_rng.hpp (exact_shuffle, itself pinned by the
// real glibc/libstdc++ golden vectors in tests/test_host.py).
// usage: shuffle2_emu <rows> <epochs> <chunk_len> <grid of the link/resolve kernels> <variant 2|3> [coarse shift]
//   variant 2: scatter by granule (global cursors), R written in place
//   variant 3: coarse buckets (global cursors) + refine, R by coarse ranges + place
//   variant 4: coarse buckets (per-block counts in shared memory, scanned (bucket, block) matrix) + refine, R in place
//   variant 5: as 4, R by coarse ranges + place
#include "cuda_emu.hpp"

#include <cstdio>
#include <cstdlib>
#include <numeric>

#include "../../cuda_sgd_sese_project_b200/csrc/host_rng.hpp"
#include "../../cuda_sgd_sese_project_b200/csrc/shuffle2_kernels.cuh"

using namespace b200sgd;

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? std::atoll(argv[1]) : 100000;
  const int epochs = argc > 2 ? std::atoi(argv[2]) : 2;
  const int64_t L = argc > 3 ? std::atoll(argv[3]) : 1024;
  const unsigned grid = argc > 4 ? (unsigned)std::atoi(argv[4]) : 7;
  const int variant = argc > 5 ? std::atoi(argv[5]) : 2;
  int cshift = 14;
  while (cshift < 18 && ((n + (1ll << cshift) - 1) >> cshift) > kCoarseMax) ++cshift;
  if (argc > 6) cshift = std::atoi(argv[6]);
  const int64_t ncoarse = (n + (1ll << cshift) - 1) >> cshift;
  const int64_t nranges = (n + (1ll << kRangeLog) - 1) >> kRangeLog;
  const int S = 64;  // sub-chunks per host chunk (shuf2_expand_states_kernel)
  const int64_t Lh = L * S, nchunks_h = (n - 1 + Lh - 1) / Lh;
  const int64_t nchunks = (n - 1 + L - 1) / L;
  const int64_t ngran = (n + kGran - 1) / kGran;
  GlibcRand g_dev, g_host;
  const GlibcRand::Poly jump = GlibcRand::poly_z_pow((uint64_t)Lh);
  std::vector<uint32_t> states_h((size_t)31 * std::max<int64_t>(1, nchunks_h)), subjump((size_t)31 * (S - 1));
  for (int s = 1; s < S; ++s) {
    const GlibcRand::Poly pz = GlibcRand::poly_z_pow((uint64_t)(L * s));
    for (int j = 0; j < 31; ++j) subjump[(size_t)(s - 1) * 31 + j] = pz.a[j];
  }
  std::vector<uint32_t> states((size_t)31 * std::max<int64_t>(1, nchunks_h * S));
  std::vector<int32_t> gcnt((size_t)ngran + 1), gstart((size_t)ngran + 1), cursor((size_t)ngran + 1);
  std::vector<uint2> pairs((size_t)n), pairs1((size_t)(nranges << kRangeLog));
  std::vector<int32_t> mcnt, moff;
  std::vector<int32_t> ccnt((size_t)ncoarse + 1), cstart((size_t)ncoarse + 1), ccur((size_t)ncoarse + 1), rcur((size_t)nranges);
  std::vector<int32_t> last((size_t)n, -7), R((size_t)n, INT32_MIN), perm((size_t)n), out((size_t)n), map((size_t)n);
  std::vector<int32_t> ref((size_t)n);
  std::iota(perm.begin(), perm.end(), 0);
  std::iota(ref.begin(), ref.end(), 0);
  unsigned err = 0;
  for (int e = 1; e <= epochs; ++e) {
    g_dev.chunk_states(jump, nchunks_h, states_h.data());
    emu_launch((unsigned)((nchunks_h * S + 127) / 128), 128, [&] {
      shuf2_expand_states_kernel(states_h.data(), nchunks_h, subjump.data(), S, states.data());
    });
    g_dev.skip((uint64_t)(n - 1));
    std::fill(gcnt.begin(), gcnt.end(), 0);
    std::fill(last.begin(), last.end(), -7);
    std::fill(R.begin(), R.end(), INT32_MIN);
    const unsigned sblocks = (unsigned)((nchunks + 127) / 128);
    const unsigned pblocks = std::max(1u, std::min(sblocks, grid / 2 + 1));  // the grid takes the virtual blocks in turns
    if (variant == 2) {
      emu_launch(pblocks, 128, [&] {
        shuf2_stream_kernel<0, false>(states.data(), L, nchunks, n, kGranLog, 0, (int)sblocks, gcnt.data(), nullptr, nullptr);
      });
      int32_t acc = 0;  // the device uses the three scan kernels of shuffle_kernels.cuh (warp shuffles: not emulated)
      for (int64_t g = 0; g <= ngran; ++g) {
        gstart[(size_t)g] = acc;
        if (g < ngran) acc += gcnt[(size_t)g];
      }
      cursor = gstart;
      emu_launch(pblocks, 128, [&] {
        shuf2_stream_kernel<1, false>(states.data(), L, nchunks, n, kGranLog, 0, (int)sblocks, cursor.data(), pairs.data(), R.data());
      });
      emu_launch(grid, kLinkThreads, [&] {
        shuf2_link_kernel<false>(n, ngran, gstart.data(), pairs.data(), last.data(), R.data(), nullptr, nullptr, &err);
      });
    } else {
      std::fill(ccnt.begin(), ccnt.end(), 0);
      std::fill(gstart.begin(), gstart.end(), -12345);
      const bool rscatter = variant == 3 || variant == 5;
      int64_t cstride = 1;
      const int32_t* coff = cstart.data();
      if (variant == 3) {
        emu_launch(pblocks, 128, [&] {
          shuf2_stream_kernel<0, false>(states.data(), L, nchunks, n, cshift, 0, (int)sblocks, ccnt.data(), nullptr, nullptr);
        });
        int32_t acc = 0;
        for (int64_t b = 0; b <= ncoarse; ++b) {
          cstart[(size_t)b] = acc;
          if (b < ncoarse) acc += ccnt[(size_t)b];
        }
        ccur = cstart;
        emu_launch(pblocks, 128, [&] {
          shuf2_stream_kernel<1, false>(states.data(), L, nchunks, n, cshift, 0, (int)sblocks, ccur.data(), pairs1.data(), nullptr);
        });
      } else {
        mcnt.assign((size_t)ncoarse * sblocks + 1, -1);
        emu_launch(pblocks, 128, [&] {
          shuf2_stream_kernel<0, true>(states.data(), L, nchunks, n, cshift, (int)ncoarse, (int)sblocks, mcnt.data(), nullptr, nullptr);
        });
        moff.resize(mcnt.size());
        int32_t acc = 0;
        for (size_t q = 0; q + 1 < mcnt.size(); ++q) {
          moff[q] = acc;
          acc += mcnt[q];
        }
        moff.back() = acc;
        mcnt = moff;  // MODE 1 reads its cursors' start values from the scanned matrix
        emu_launch(pblocks, 128, [&] {
          shuf2_stream_kernel<1, true>(states.data(), L, nchunks, n, cshift, (int)ncoarse, (int)sblocks, mcnt.data(), pairs1.data(),
                                       rscatter ? nullptr : R.data());
        });
        cstride = sblocks;
        coff = moff.data();
      }
      emu_launch(grid, kLinkThreads, [&] {
        shuf2_refine_kernel(ngran, ncoarse, cshift, coff, cstride, pairs1.data(), pairs.data(), gstart.data());
      });
      if (!rscatter) {
        emu_launch(grid, kLinkThreads, [&] {
          shuf2_link_kernel<false>(n, ngran, gstart.data(), pairs.data(), last.data(), R.data(), nullptr, nullptr, &err);
        });
      } else {
      emu_launch(3, 256, [&] { shuf2_range_cursor_kernel(rcur.data(), nranges, kRangeLog); });
      emu_launch(grid, kLinkThreads, [&] {  // the entries of R go where the coarse pairs were
        shuf2_link_kernel<true>(n, ngran, gstart.data(), pairs.data(), last.data(), nullptr, rcur.data(), pairs1.data(), &err);
      });
      emu_launch(grid + 2, 256, [&] { shuf2_place_kernel(n, nranges, rcur.data(), pairs1.data(), R.data()); });
      }
    }
    emu_launch(grid, 256, [&] { shuf2_resolve_kernel(n, last.data(), R.data(), perm.data(), out.data()); });
    emu_launch(grid + 1, 256, [&] { shuf2_resolve_kernel(n, last.data(), R.data(), nullptr, map.data()); });
    if (err) {
      std::printf("FAIL: error word %u\n", err);
      return 1;
    }
    exact_shuffle(g_host, ref.data(), n);
    int64_t bad = 0, bad_map = 0, unset = 0;
    for (int64_t p = 0; p < n; ++p) {
      bad += out[(size_t)p] != ref[(size_t)p];
      bad_map += perm[(size_t)map[(size_t)p]] != ref[(size_t)p];
      unset += last[(size_t)p] == -7 || R[(size_t)p] == INT32_MIN;
    }
    std::printf("variant %d rows %lld epoch %d chunk %lld hits %d: %lld wrong, %lld wrong via the map, %lld unset\n", variant, (long long)n, e,
                (long long)L, gstart[(size_t)ngran], (long long)bad, (long long)bad_map, (long long)unset);
    if (bad || bad_map || unset) return 1;
    perm.swap(out);
  }
  std::printf("OK\n");
  return 0;
}
