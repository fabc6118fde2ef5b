"""The kernels of the device shuffle (csrc/shuffle2_kernels.cuh) are written with block-level primitives
only, so the SAME source also compiles for the host (tests/emu/cuda_emu.hpp: one OS thread per CUDA
thread, __syncthreads = a barrier).  This runs them through the launch sequence of b200sgd_api.cu and
compares every epoch's permutation and position map with the host loop (exact_shuffle, pinned by the real
glibc/libstdc++ golden vectors in test_host.py).  No GPU needed; the GPU pool has no compute-sanitizer."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emu") / "shuffle2_emu")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-Wno-unknown-pragmas",
                    os.path.join(HERE, "emu", "shuffle2_emu.cpp"), "-o", exe], check=True)
    return exe


# rows, epochs, sub-chunk length, grid of the grid-stride kernels, variant (B200SGD_SHUFFLE 2..5)
@pytest.mark.parametrize("args", [
    (100000, 2, 128, 7, 4), (100000, 2, 128, 7, 5), (100000, 2, 128, 7, 3), (100000, 2, 1024, 7, 2),
    (32769, 2, 128, 1, 4), (300001, 2, 256, 40, 4), (36865, 2, 128, 3, 5), (257, 2, 128, 2, 4),
    (1025, 2, 8, 2, 4), (524289, 1, 128, 64, 4),
])
def test_shuffle_kernels_on_the_host_match_the_host_loop(emu, args):
    out = subprocess.run([emu] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


def test_coarse_bucket_shifts_of_very_large_data_sets(emu):
    # 2^31 rows would use 2^18 positions per coarse bucket (1024 granules): force the shift on a small case
    for shift, variant in ((16, 4), (18, 4), (9, 5), (18, 3)):
        out = subprocess.run([emu, "600000", "1", "128", "9", str(variant), str(shift)], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
