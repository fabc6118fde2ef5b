"""The device permutation must equal the reference's host loop (std::random_shuffle + glibc rand(),
main.cu:90 / :207) bit for bit, cumulatively over epochs -- in every formulation B200SGD_SHUFFLE selects: the backward one
(csrc/shuffle2_kernels.cuh; 5 = the default, 2 / 3 / 4 = its other scatter strategies) and the forward
one of round 1 (csrc/shuffle_kernels.cuh, = 1)."""
import numpy as np
import pytest

import cuda_sgd_sese_project_b200 as pkg
from cuda_sgd_sese_project_b200 import _lib
from oracle import loader as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[4, 5, 3, 2, 1], ids=["backward-blockbins", "backward-blockbins-rscatter", "backward-2level", "backward-direct", "forward"])
def shuffle_version(request, monkeypatch):
    monkeypatch.setenv("B200SGD_SHUFFLE", str(request.param))  # read by sgd_create
    return request.param


def _checksum(ind):
    i = np.arange(1, ind.shape[0] + 1, dtype=np.uint64)
    return int((i * ind.astype(np.uint64)).sum(dtype=np.uint64))


@pytest.mark.parametrize("R", [32768, 32769, 40000, 65536, 100003, 1 << 20])
def test_device_shuffle_equals_oracle_over_epochs(R, shuffle_version):
    g = orc.Rand()
    ind = np.arange(R, dtype=np.int32)
    with pkg.SGDEngine(R, 3, 1000, 0.1) as eng:
        eng.load_synthetic(seed=1)
        for epoch in range(3):
            g.shuffle(ind)                  # cumulative, never reset (main.cu:421-424)
            eng.shuffle()
            got = eng.permutation()
            assert np.array_equal(got, ind), f"epoch {epoch + 1}: first mismatch at {np.nonzero(got != ind)[0][:5]}"


def test_device_and_host_paths_agree_and_interleave(shuffle_version):
    R = 50000
    with pkg.SGDEngine(R, 2, 500, 0.1) as dev, pkg.SGDEngine(R, 2, 500, 0.1, ) as dev2:
        dev.load_synthetic(seed=1)
        dev2.load_synthetic(seed=1)
        # same engine type twice must give identical streams
        for _ in range(2):
            dev.shuffle(); dev2.shuffle()
            assert np.array_equal(dev.permutation(), dev2.permutation())
    ref = np.arange(R, dtype=np.int32)
    pkg.host_shuffle(ref, epochs=2)
    with pkg.SGDEngine(R, 2, 500, 0.1) as dev:
        dev.load_synthetic(seed=1)
        dev.shuffle(); dev.shuffle()
        assert np.array_equal(dev.permutation(), ref)


@pytest.mark.parametrize("R", [4000000, 16777216, 67108864])   # configs 2/3, 4 and 5
def test_device_shuffle_golden_kat(kat, R, shuffle_version):
    # golden values from the REAL glibc / libstdc++ (tests/golden/thirdparty_kat.json)
    if shuffle_version != 5 and R > 4000000:   # 5 = the default strategy
        pytest.skip("the non-default formulations are checked at 4M rows")
    recs = [r for r in kat["shuffle_kat"] if r["R"] == R]
    with pkg.SGDEngine(R, 1, 1000, 0.1) as eng:
        eng.load_synthetic(seed=1)
        for rec in recs:
            eng.shuffle()
            p = eng.permutation()
            assert p[:4].tolist() == rec["head"] and int(p[-1]) == rec["last"]
            assert _checksum(p) == rec["checksum"]


def test_set_permutation_then_device_shuffle_continues_from_it(shuffle_version):
    R = 40000
    start = np.arange(R, dtype=np.int32)[::-1].copy()
    g = orc.Rand()
    want = start.copy()
    g.shuffle(want)
    with pkg.SGDEngine(R, 2, 500, 0.1) as eng:
        eng.load_synthetic(seed=1)
        eng.set_permutation(start)
        eng.shuffle()
        assert np.array_equal(eng.permutation(), want)


def test_odd_sizes_around_the_granule_and_range_boundaries(shuffle_version):
    # 256 positions per granule, 1024 per link job, 4096 steps per range of R, 8192 per coarse bucket
    for R in (32768 + 255, 32768 + 257, 36864, 36865, 40960 + 1023, 49152 + 1):
        g = orc.Rand()
        ind = np.arange(R, dtype=np.int32)
        with pkg.SGDEngine(R, 1, 1000, 0.1) as eng:
            eng.load_synthetic(seed=1)
            for _ in range(2):
                g.shuffle(ind)
                eng.shuffle()
                assert np.array_equal(eng.permutation(), ind), R


def test_lookahead_is_transparent():
    """sgd_create, sgd_run and sgd_iteration enqueue the NEXT permutation ahead of time (the rand() stream is
    never seeded); shuffle requests, sgd_get_permutation and sgd_set_permutation must not see a difference."""
    R = 50000
    g = orc.Rand()
    ind = np.arange(R, dtype=np.int32)
    with pkg.SGDEngine(R, 4, 1000, 0.01) as eng:
        eng.load_synthetic(seed=1)
        eng.run(1, 2)                                   # shuffles 1, 2; 3 is now ahead
        g.shuffle(ind); g.shuffle(ind)
        assert np.array_equal(eng.permutation(), ind)
        assert np.array_equal(eng.permutation(), ind)   # reading does not consume the look-ahead
        eng.shuffle()                                   # takes shuffle 3
        g.shuffle(ind)
        assert np.array_equal(eng.permutation(), ind)
        eng.run(4, 1)                                   # shuffle 4; 5 ahead
        g.shuffle(ind)
        assert np.array_equal(eng.permutation(), ind)
        start = ind[::-1].copy()
        eng.set_permutation(start)                      # discards the look-ahead: shuffle 5 must be drawn again
        assert np.array_equal(eng.permutation(), start)
        eng.shuffle()
        g.shuffle(start)
        assert np.array_equal(eng.permutation(), start)
        eng.iteration(6)                                # shuffle 6 + epoch; 7 ahead
        g.shuffle(start)
        assert np.array_equal(eng.permutation(), start)
        eng.iteration(7)
        g.shuffle(start)
        assert np.array_equal(eng.permutation(), start)


def test_lookahead_off_gives_the_same_permutations(monkeypatch):
    R = 40000
    monkeypatch.setenv("B200SGD_NO_LOOKAHEAD", "1")
    g = orc.Rand()
    ind = np.arange(R, dtype=np.int32)
    with pkg.SGDEngine(R, 4, 1000, 0.01) as eng:
        eng.load_synthetic(seed=1)
        eng.run(1, 2)
        g.shuffle(ind); g.shuffle(ind)
        assert np.array_equal(eng.permutation(), ind)
