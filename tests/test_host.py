"""Host-side logic of the product (libb200sgd.so through the Python mirror), checked against the
third-party known answers, the oracle and golden byte strings.  CPU only."""
import json
import os

import numpy as np
import pytest

import cuda_sgd_sese_project_b200 as pkg
from oracle import loader as orc


def _checksum(ind):
    i = np.arange(1, ind.shape[0] + 1, dtype=np.uint64)
    return int((i * ind.astype(np.uint64)).sum(dtype=np.uint64))


@pytest.mark.parametrize("n", [1, 2, 5, 10, 33])
def test_host_shuffle_small_exact(kat, n):
    ind = np.arange(n, dtype=np.int32)
    for e, rec in enumerate(kat["shuffle_full"][str(n)]):
        pkg.host_shuffle(ind, seed=1, skip_epochs=e, epochs=1)  # permutation indices are bit-exact
        assert ind.tolist() == rec["head"]


@pytest.mark.parametrize("n", [1000, 10000, 4000000])
def test_host_shuffle_kat(kat, n):
    recs = [r for r in kat["shuffle_kat"] if r["R"] == n]
    ind = np.arange(n, dtype=np.int32)
    for e, rec in enumerate(recs):
        pkg.host_shuffle(ind, seed=1, skip_epochs=e, epochs=1)
        assert ind[:4].tolist() == rec["head"] and int(ind[-1]) == rec["last"]
        assert _checksum(ind) == rec["checksum"]


@pytest.mark.slow
@pytest.mark.parametrize("n", [16777216])
def test_host_shuffle_kat_large(kat, n):
    recs = [r for r in kat["shuffle_kat"] if r["R"] == n]
    ind = np.arange(n, dtype=np.int32)
    pkg.host_shuffle(ind, seed=1, epochs=2)
    assert ind[:4].tolist() == recs[1]["head"] and _checksum(ind) == recs[1]["checksum"]


def test_host_shuffle_equals_oracle_for_other_seeds_and_sizes():
    for seed, n in [(1, 1025), (7, 31), (12345, 2049), (99, 100003)]:
        a = np.arange(n, dtype=np.int32)
        pkg.host_shuffle(a, seed=seed, epochs=3)
        g = orc.Rand(seed)
        b = np.arange(n, dtype=np.int32)
        for _ in range(3):
            g.shuffle(b)
        assert np.array_equal(a, b), (seed, n)


def test_init_weights(kat):
    want = np.array([int(h, 16) for h in kat["weights_hex"]], dtype=np.uint32)
    for C in (1, 100, 2048):
        assert np.array_equal(pkg.init_weights(C).view(np.uint32), want[:C])


def test_batch_geometry_matches_oracle():
    for R, B in [(40, 0), (10000, 1000), (10001, 1000), (999, 999), (4000000, 1000), (67108864, 65536), (16777217, 4096), (1000, 7)]:
        assert pkg.batch_geometry(R, B) == orc.num_batches(R, B)


def test_read_csv(tmp_path, fixtures):
    p = tmp_path / "t.csv"
    p.write_text("1,2,3,0.1\n4,5,6,0.2\n7,8,9,0.3")
    ok, X, y = pkg.read_csv(str(p), 3, 3)
    assert ok and np.array_equal(X, fixtures["example_X"]) and np.array_equal(y, fixtures["example_y"])
    p.write_text("-0.53906,-6.0791e-1, .8374 ,-142.38\r\n1e3,+2,3.5,-0\n")
    ok, X, y = pkg.read_csv(str(p), 2, 3)
    rc, Xo, yo = orc.read_csv(str(p), 2, 3)
    assert ok and rc == 0 and np.array_equal(X, Xo) and np.array_equal(y, yo)
    ok, _, _ = pkg.read_csv(str(tmp_path / "missing.csv"), 2, 3)
    assert not ok  # reference: read_csv returns false
    p.write_text("1,2\n3,4\n5,6\n")
    with pytest.raises(pkg.SgdError) as ei:
        pkg.read_csv(str(p), 2, 1)
    assert ei.value.code == pkg._lib.SGD_ERR_PARSE
    p.write_text("1,2,3\n")
    with pytest.raises(pkg.SgdError):
        pkg.read_csv(str(p), 1, 1)   # more columns than announced
    p.write_text("1,,3\n")
    with pytest.raises(pkg.SgdError):
        pkg.read_csv(str(p), 1, 2)   # std::stof("") throws in the reference


def test_read_csv_large_parallel_equals_oracle(tmp_path):
    rng = np.random.default_rng(0)
    R, C = 60000, 12   # > 4 MB of text -> exercises the multi-threaded chunking
    a = rng.standard_normal((R, C + 1)).astype(np.float32)
    p = tmp_path / "big.csv"
    with open(p, "w") as f:
        for row in a:
            f.write(",".join("%.9g" % v for v in row) + "\n")
    assert os.path.getsize(p) > (4 << 20)
    ok, X, y = pkg.read_csv(str(p), R, C)
    rc, Xo, yo = orc.read_csv(str(p), R, C)
    assert ok and rc == 0
    assert np.array_equal(X, Xo) and np.array_equal(y, yo)
    assert np.array_equal(X, a[:, :C]) and np.array_equal(y, a[:, C])


def test_write_csv_round_trips_and_special_cells(tmp_path):
    rng = np.random.default_rng(3)
    X = rng.standard_normal((1234, 7)).astype(np.float32)
    X[0, 0], X[1, 1], X[2, 2] = 0.0, -0.0, 1e-30
    X[3, 3], X[4, 4] = np.float32(3.4e38), np.float32(1.17549435e-38)
    y = (100 * rng.standard_normal(1234)).astype(np.float32)
    p = tmp_path / "rt.csv"
    pkg.write_csv(str(p), X, y)
    ok, X2, y2 = pkg.read_csv(str(p), 1234, 7)
    assert ok and np.array_equal(X2.view(np.uint32), X.view(np.uint32)) and np.array_equal(y2, y)
    rc, Xo, yo = orc.read_csv(str(p), 1234, 7)     # the oracle's getline / stof restatement reads the same file
    assert rc == 0 and np.array_equal(Xo, X) and np.array_equal(yo, y)
    # spellings std::stof accepts and from_chars alone would not: hex floats, a leading '+',
    # trailing junk, cells longer than 63 characters; out-of-range values are rejected (stof throws)
    q = tmp_path / "odd.csv"
    q.write_text("0x1p3,+2.5,1.5abc," + "0." + "0" * 30 + "1" + "0" * 40 + ",7\n")
    ok, X3, y3 = pkg.read_csv(str(q), 1, 4)
    assert ok and X3[0].tolist() == [8.0, 2.5, 1.5, float(np.float32(1e-31))] and y3[0] == 7.0
    q.write_text("1e-81,1\n")          # underflow: strtof sets ERANGE, std::stof throws out_of_range
    with pytest.raises(pkg.SgdError):
        pkg.read_csv(str(q), 1, 1)
    q.write_text("1e50,1\n")
    with pytest.raises(pkg.SgdError):
        pkg.read_csv(str(q), 1, 1)


def test_batch_geometry_is_clamped_where_the_float_division_rounds_up():
    # main.cu:344 computes floor(R / (float)B) in float: for these R it exceeds R // B
    for R, B in ((99_999_999, 1000), (33_554_431, 4096), (2**31 - 1, 65536)):
        f = int(np.floor(np.float32(R) / np.float32(B)))
        assert f * B > R                                   # the reference would index past its vector
        b, nb = pkg.batch_geometry(R, B)
        assert (b, nb) == (B, R // B) and nb * B <= R
    for R, B in ((10_000, 1000), (4_000_000, 1000), (67_108_864, 65536), (16_777_216, 4096), (1000, 3)):
        assert pkg.batch_geometry(R, B) == orc.num_batches(R, B)


def test_get_filename_from_path():
    # sgd_io.cuh docstring example: "/path/to/file.csv" -> "file"
    assert pkg.get_filename_from_path("/path/to/file.csv") == "file"
    assert pkg.get_filename_from_path("data/5xy.csv") == "5xy"
    assert pkg.get_filename_from_path("plain") == "plain"
    assert pkg.get_filename_from_path("a/b.c/test_number_of_features-1000.csv") == "test_number_of_features-1000"


def test_experiment_json_is_byte_compatible_with_jsoncpp(tmp_path):
    # bytes of experiments/results/100b-1200b-cublas/test_number_of_samples_scalable-1000-cublas.json
    golden = ('{\n   "error" : 0.016775146126747131,\n   "gpu_time" : 102.58531188964844,\n'
              '   "shuffle_time" : 421.97488403320312,\n   "total_gradient_time" : 292.31671142578125,\n'
              '   "transfer_time" : 795.40350341796875\n}\n')
    vals = json.loads(golden)
    p = tmp_path / "x-cublas.json"
    pkg.write_experiment_output(str(p), np.float32(vals["shuffle_time"]), np.float32(vals["total_gradient_time"]),
                                np.float32(vals["gpu_time"]), np.float32(vals["transfer_time"]), np.float32(vals["error"]))
    assert p.read_text() == golden
    assert sorted(json.loads(p.read_text())) == ["error", "gpu_time", "shuffle_time", "total_gradient_time", "transfer_time"]
    pkg.write_experiment_output(str(p), 0.0, 8.0, float("nan"), float("inf"), -1.5)
    assert p.read_text() == ('{\n   "error" : -1.5,\n   "gpu_time" : null,\n   "shuffle_time" : 0,\n'
                             '   "total_gradient_time" : 8,\n   "transfer_time" : 1e+9999\n}\n')


def test_shuffle_is_a_position_map_independent_of_the_values():
    """What the ranks' shared permutation work rests on (csrc/shuffle_kernels.cuh): std::random_shuffle moves
    values without looking at them, so epoch k's pass is a position map src_k that depends on epoch k's
    rand() stream only, and perm_k[p] = perm_{k-1}[src_k[p]].  Checked on the host: the map of every epoch
    comes from shuffling the IDENTITY with that epoch's part of the stream (skip_epochs = k - 1)."""
    R, E = 20011, 5
    perm = np.arange(R, dtype=np.int32)
    want = np.arange(R, dtype=np.int32)
    for k in range(1, E + 1):
        src = np.arange(R, dtype=np.int32)
        pkg.host_shuffle(src, skip_epochs=k - 1, epochs=1)       # the map of shuffle k alone
        perm = perm[src]                                         # applied by a gather
        pkg.host_shuffle(want, skip_epochs=k - 1, epochs=1)      # the reference's cumulative loop
        assert np.array_equal(perm, want), k
    g = orc.Rand()
    ind = np.arange(R, dtype=np.int32)
    for _ in range(E):
        g.shuffle(ind)
    assert np.array_equal(ind, want)
