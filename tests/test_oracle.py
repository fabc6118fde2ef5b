"""Pins the CPU oracle (oracle/sgd_oracle.c):
  1. against the REAL third-party implementations that define the reference's permutation and
     initial weights (tests/golden/thirdparty_kat.json, see make_golden.py);
  2. against every known-answer fixture of the reference's own test code (c_code/testing.cu);
  3. against an independent float64 numpy restatement of the whole run.
All CPU, no GPU, a few seconds."""
import os

import numpy as np
import pytest

from oracle import loader as orc
from tests.helpers import make_regression, numpy_run, rel_l2


# ---------------------------------------------------------------- third-party pieces
def test_rand_stream_matches_glibc(kat):
    g = orc.Rand()
    got = [g.next() for _ in range(len(kat["rand_first"]))]
    assert got == kat["rand_first"]


def test_rand_far_positions(kat):
    g = orc.Rand()
    want = {r["k"]: r for r in kat["rand_at"] if r["k"] <= 999999}
    s = 0
    for i in range(max(want) + 1):
        v = g.next()
        s = (s + (i + 1) * v) % (1 << 64)
        if i in want:
            assert v == want[i]["value"]
            assert s == want[i]["checksum"]


@pytest.mark.parametrize("n", [1, 2, 5, 10, 33])
def test_random_shuffle_small_exact(kat, n):
    g = orc.Rand()
    ind = np.arange(n, dtype=np.int32)
    for rec in kat["shuffle_full"][str(n)]:
        g.shuffle(ind)  # cumulative: never reset (main.cu:421-424)
        assert ind.tolist() == rec["head"], rec


def _checksum(ind):
    i = np.arange(1, ind.shape[0] + 1, dtype=np.uint64)
    return int((i * ind.astype(np.uint64)).sum(dtype=np.uint64))


@pytest.mark.parametrize("n", [1000, 10000, 4000000])
def test_random_shuffle_kat(kat, n):
    recs = [r for r in kat["shuffle_kat"] if r["R"] == n]
    assert len(recs) == 2
    g = orc.Rand()
    ind = np.arange(n, dtype=np.int32)
    for rec in recs:
        g.shuffle(ind)
        assert ind[:4].tolist() == rec["head"]
        assert int(ind[-1]) == rec["last"]
        assert _checksum(ind) == rec["checksum"]
        assert np.array_equal(np.sort(ind), np.arange(n, dtype=np.int32))


def test_weight_init_matches_thrust(kat):
    want = np.array([int(h, 16) for h in kat["weights_hex"]], dtype=np.uint32).view(np.float32)
    for C in (1, 6, 100, 2048):
        got = orc.init_weights(C)
        assert np.array_equal(got.view(np.uint32), want[:C].view(np.uint32))
    assert np.all((want >= 0) & (want < 0.01))


# ---------------------------------------------------------------- scalar pieces
def test_batch_arithmetic():
    # main.cu:343-344: B=0 -> whole data; nb = floor(R/(float)B), tail dropped
    assert orc.num_batches(40, 0) == (40, 1)
    assert orc.num_batches(10000, 1000) == (1000, 10)
    assert orc.num_batches(10001, 1000) == (1000, 10)
    assert orc.num_batches(999, 1000) == (1000, 0)
    assert orc.num_batches(4000000, 1000) == (1000, 4000)
    assert orc.num_batches(67108864, 65536) == (65536, 1024)


def test_step_scale():
    # main.cu:144 / :282: exponent 0.25 (NOT the report's sqrt), double math narrowed to float
    assert orc.step_scale(0.1, 1) == pytest.approx(-0.1, rel=1e-7)
    for e in (2, 3, 16, 20):
        want = np.float32(-(float(np.float32(0.1)) / (e ** 0.25)))
        assert orc.step_scale(0.1, e) == float(want)
    assert orc.step_scale(0.1, 16) == pytest.approx(-0.05, rel=1e-6)


# ---------------------------------------------------------------- the reference's own fixtures
def test_fixture_abs_error(fixtures):
    # testing.cu:15-55, "Mean absolute error (8.0 expected)" at testing.cu:54
    X, y = fixtures["gemv_test_X"], fixtures["gemv_test_y"]
    assert X.shape == (5, 4)
    assert orc.mean_abs_error(X, y, np.ones(4, np.float32)) == 8.0


def test_fixture_gemv(fixtures):
    # testing.cu:225-275: loss derivative with w = 1 on gemv_test: row i sums to 4(i+1), label 4
    X, y = fixtures["gemv_test_X"], fixtures["gemv_test_y"]
    r = orc.residuals(X, y, np.ones(4, np.float32))
    assert r.tolist() == [0.0, 4.0, 8.0, 12.0, 16.0]


def test_fixture_permutation(fixtures):
    # testing.cu:57-105, order vector at testing.cu:80-84
    X, y = fixtures["permutation_test_X"], fixtures["permutation_test_y"]
    Xs, ys = orc.permute_rows(X, y, [4, 2, 0, 1, 3])
    assert ys.tolist() == [5.0, 3.0, 1.0, 2.0, 4.0]
    assert Xs[:, 0].tolist() == [5.0, 3.0, 1.0, 2.0, 4.0]
    assert np.all(Xs == Xs[:, :1])


def test_fixture_col_sums(fixtures):
    # testing.cu:107-138 and :140-188 (scale 2.0)
    X = fixtures["col_sum_test_X"]
    assert orc.col_sums(X).tolist() == [5.0, 10.0, 15.0, 20.0]
    assert orc.col_sums(X, 2.0).tolist() == [2.5, 5.0, 7.5, 10.0]


def test_fixture_matrix_scale(fixtures):
    # testing.cu:190-223: ones scaled by the "labels" 2..6
    X, v = fixtures["matrix_scale_test_X"], fixtures["matrix_scale_test_y"]
    out = orc.scale_rows(X, v)
    assert np.array_equal(out, np.repeat(np.arange(2, 7, dtype=np.float32)[:, None], 4, axis=1))


def test_gradient_is_colmean_of_scaled_rows(fixtures):
    # cuBLAS path: scale rows by r (sgd_thrust.cu:301-315) then column mean (:254-291)
    X, y = fixtures["enb2012_X"], fixtures["enb2012_y"]
    w = orc.init_weights(8)
    rows = np.arange(100, 200, dtype=np.int32)
    r = orc.residuals(X, y, w, rows)
    g = orc.gradient(X, r, 100.0, rows)
    g2 = orc.col_sums(orc.scale_rows(X[rows], r), 100.0)
    assert np.array_equal(g, g2)


def test_read_csv(tmp_path, fixtures):
    # sgd_io.cu:38-77 semantics: no header, last column is the label, stof per cell
    p = tmp_path / "t.csv"
    p.write_text("1,2,3,0.1\n4,5,6,0.2\n7,8,9,0.3")  # c_code/test-data/example.csv shape, no EOL
    rc, X, y = orc.read_csv(str(p), 3, 3)
    assert rc == 0
    assert np.array_equal(X, fixtures["example_X"]) and np.array_equal(y, fixtures["example_y"])
    p.write_text("-0.53906,-6.0791e-1, .8374 ,-142.38\r\n1e3,+2,3.5,-0\n")
    rc, X, y = orc.read_csv(str(p), 2, 3)
    assert rc == 0
    assert np.allclose(X, [[-0.53906, -0.60791, 0.8374], [1000, 2, 3.5]]) and y.tolist() == [np.float32(-142.38), 0.0]
    assert orc.read_csv(str(tmp_path / "missing.csv"), 2, 3)[0] == -1
    p.write_text("1,2\n3,4\n5,6\n")
    assert orc.read_csv(str(p), 2, 1)[0] == -2  # reference would write out of bounds


# ---------------------------------------------------------------- whole-run cross-checks
def _perms(R, epochs):
    g = orc.Rand()
    ind = np.arange(R, dtype=np.int32)
    out = []
    for _ in range(epochs):
        g.shuffle(ind)
        out.append(ind.copy())
    return out


@pytest.mark.parametrize("name,C,B,lr,epochs", [
    ("xy5", 1, 0, 1e-5, 10),                 # the docstring example main.cu:311
    ("correct_prediction", 3, 100, 0.1, 10),
    ("correct_prediction", 3, 7, 0.05, 3),   # ragged: 1000 = 142*7 + 6, tail rows dropped
])
def test_run_matches_numpy_restatement(fixtures, name, C, B, lr, epochs):
    X, y = fixtures[name + "_X"], fixtures[name + "_y"]
    R = X.shape[0]
    mae32, w32, ind = orc.run(X, y, B, lr, epochs)
    mae64, w64, _ = orc.run(X, y, B, lr, epochs, acc64=True)
    perms = _perms(R, epochs)
    assert np.array_equal(ind, perms[-1])
    mae_np, w_np = numpy_run(X, y, B, lr, epochs, perms, orc.init_weights(C))
    assert rel_l2(w32, w_np) < 1e-5
    assert rel_l2(w64, w_np) < 1e-6
    assert abs(mae32 - mae_np) <= 1e-4 * abs(mae_np)


def test_run_published_sanity(fixtures):
    # SURVEY Appendix B indicative values for these fixtures
    mae, w, _ = orc.run(fixtures["xy5_X"], fixtures["xy5_y"], 0, 1e-5, 10)
    assert w[0] == pytest.approx(1.04520, rel=1e-4) and mae == pytest.approx(198.1447, rel=1e-4)
    mae, w, _ = orc.run(fixtures["correct_prediction_X"], fixtures["correct_prediction_y"], 100, 0.1, 10)
    assert np.allclose(w, [25.9529, 98.2672, 81.7486], rtol=1e-4)
    assert mae == pytest.approx(0.1208, rel=2e-3)
    # converges towards the generating weights (correct_prediction_weights.csv)
    assert rel_l2(w, fixtures["correct_prediction_true_w"]) < 2e-3


def test_c1_config_converges():
    # BASELINE config 1: 10k x 100, B=1000, lr 0.1, 20 epochs (report tex:536-538), fp16-rounded
    X, y, w_true = make_regression(10000, 100, seed=42, fp16_round=True)
    mae, w, _ = orc.run(X, y, 1000, 0.1, 20)
    assert rel_l2(w, w_true) < 5e-4
    mae64, w64, _ = orc.run(X, y, 1000, 0.1, 20, acc64=True)
    assert rel_l2(w, w64) < 1e-5      # fp32 vs fp64 accumulators (SURVEY 8c tolerance sanity)


def test_openmp_baseline_equals_scalar_oracle_within_tolerance():
    X, y, _ = make_regression(6000, 37, seed=3)
    mae1, w1, ind1 = orc.run(X, y, 250, 0.05, 3, threads=1)
    maeN, wN, indN = orc.run(X, y, 250, 0.05, 3, threads=max(2, min(8, orc.max_threads())))
    assert np.array_equal(ind1, indN)
    assert rel_l2(wN, w1) < 1e-5 and abs(maeN - mae1) <= 1e-4 * mae1


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_sharded_epoch_equals_single_rank(nranks):
    X, y, _ = make_regression(5000, 64, seed=5)
    perms = _perms(5000, 2)
    w1 = orc.init_weights(64); wN = w1.copy()
    for e in (1, 2):
        orc.epoch(X, y, 500, perms[e - 1], w1, 0.1, e)
        orc.epoch_sharded(X, y, 500, perms[e - 1], wN, 0.1, e, nranks)
    assert rel_l2(wN, w1) < 1e-5
