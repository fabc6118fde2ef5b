"""Parity of the CUDA path (libb200sgd.so, through the C-ABI) with the CPU oracle.  Needs a B200.

Bars (BASELINE.json north_star): permutation indices bit-exact; weights rel-L2 <= 1e-4 and MAE
within 1e-4 relative of the oracle (fp32 arithmetic, different summation order)."""
import os
import subprocess

import numpy as np
import pytest

import cuda_sgd_sese_project_b200 as pkg
from cuda_sgd_sese_project_b200 import _lib
from oracle import loader as orc
from tests.helpers import make_regression, rel_l2

pytestmark = pytest.mark.gpu

W_TOL = 1e-4     # relative L2 on the weights (north_star)
MAE_TOL = 1e-4   # relative on the mean absolute error


def oracle_epochs(X, y, B_arg, lr, epochs, seed=1):
    """Oracle run that also returns the permutation of every epoch."""
    R, C = X.shape
    B, _ = orc.num_batches(R, B_arg)
    g = orc.Rand(seed)
    ind = np.arange(R, dtype=np.int32)
    w = orc.init_weights(C)
    perms = []
    for e in range(1, epochs + 1):
        g.shuffle(ind)
        perms.append(ind.copy())
        orc.epoch(X, y, B, ind, w, lr, e)
    return w, perms, orc.mean_abs_error(X, y, w)


def check_run(X, y, B_arg, lr, epochs, **kw):
    R, C = X.shape
    w_o, perms, mae_o = oracle_epochs(X, y, B_arg, lr, epochs)
    with pkg.SGDEngine(R, C, B_arg, lr, data=X, labels=y, **kw) as eng:
        for e in range(1, epochs + 1):
            eng.iteration(e)                       # one cuda_iteration / cublas_iteration
            assert np.array_equal(eng.permutation(), perms[e - 1]), f"permutation differs in epoch {e}"
        w = eng.weights
        mae = eng.calculate_mean_abs_error()
    assert np.all(np.isfinite(w))
    assert rel_l2(w, w_o) <= W_TOL, (rel_l2(w, w_o), kw)
    # relative to the oracle's MAE, with a floor at fp32 resolution of the labels (a converged,
    # noise-free fit has an MAE that is pure rounding noise)
    assert abs(mae - mae_o) <= MAE_TOL * abs(mae_o) + 1e-6 * float(np.mean(np.abs(y))), (mae, mae_o)
    return w, mae


# ------------------------------------------------------------------ the reference's own fixtures
def test_fixture_abs_error(fixtures):
    # testing.cu:15-55 "Mean absolute error (8.0 expected)"
    X, y = fixtures["gemv_test_X"], fixtures["gemv_test_y"]
    with pkg.SGDEngine(5, 4, 0, 0.0, data=X, labels=y) as eng:
        eng.weights = np.ones(4, np.float32)
        assert eng.calculate_mean_abs_error() == 8.0


def test_fixture_gemv_residuals_and_gradient(fixtures):
    # testing.cu:225-275: loss derivative with w = 1 -> [0,4,8,12,16]; then the scaled column sums
    X, y = fixtures["gemv_test_X"], fixtures["gemv_test_y"]
    with pkg.SGDEngine(5, 4, 0, 0.0, data=X, labels=y, keep_residuals=True) as eng:
        eng.weights = np.ones(4, np.float32)
        eng.set_permutation(np.arange(5))
        eng.epoch(1)
        r = eng.residuals()
        assert r.tolist() == [0.0, 4.0, 8.0, 12.0, 16.0]
        g = eng.last_gradient()
        assert np.allclose(g, orc.gradient(X, r, 5.0), rtol=1e-6)
        assert np.array_equal(eng.weights, np.ones(4, np.float32))   # lr = 0: a = -0


def test_fixture_permutation(fixtures):
    # testing.cu:57-105 with the order vector of testing.cu:80-84: the gather follows `order`
    X, y = fixtures["permutation_test_X"], fixtures["permutation_test_y"]
    with pkg.SGDEngine(5, 4, 0, 0.0, data=X, labels=y, keep_residuals=True) as eng:
        eng.weights = np.zeros(4, np.float32)
        eng.set_permutation([4, 2, 0, 1, 3])
        assert eng.permutation().tolist() == [4, 2, 0, 1, 3]
        eng.epoch(1)
        assert (-eng.residuals()).tolist() == [5.0, 3.0, 1.0, 2.0, 4.0]   # permuted labels
        eng.weights = np.array([1, 0, 0, 0], np.float32)
        eng.epoch(1)
        assert eng.residuals().tolist() == [0.0] * 5                        # permuted rows match labels


def test_fixture_col_sums_and_matrix_scale(fixtures):
    # testing.cu:107-188 (column sums [5,10,15,20], scaled by 2 -> [2.5,5,7.5,10]) and :190-223
    # (rows scaled by 2..6): g = X^T r / B with chosen r reproduces both expectations
    X = fixtures["col_sum_test_X"]
    # labels = -1 and w = 0 give r = +1 for every row -> g = column sums / B
    with pkg.SGDEngine(5, 4, 0, 0.0, data=X, labels=-np.ones(5, np.float32)) as eng:
        eng.weights = np.zeros(4, np.float32)
        eng.set_permutation(np.arange(5))
        eng.epoch(1)
        assert (eng.last_gradient() * 5).tolist() == [5.0, 10.0, 15.0, 20.0]
    Xs, v = fixtures["matrix_scale_test_X"], fixtures["matrix_scale_test_y"]
    with pkg.SGDEngine(5, 4, 0, 0.0, data=Xs, labels=-v) as eng:   # r_i = v_i = 2..6
        eng.weights = np.zeros(4, np.float32)
        eng.set_permutation(np.arange(5))
        eng.epoch(1)
        assert (eng.last_gradient() * 5).tolist() == [20.0] * 4      # sum of the scaled rows


@pytest.mark.parametrize("name,C,B,lr,epochs", [
    ("xy5", 1, 0, 1e-5, 10),                 # main.cu:311 docstring example
    ("correct_prediction", 3, 100, 0.1, 10),
    ("correct_prediction", 3, 7, 0.05, 3),   # ragged tail (1000 = 142*7 + 6)
    ("enb2012", 8, 64, 1e-6, 4),
])
def test_end_to_end_fixtures(fixtures, name, C, B, lr, epochs):
    check_run(fixtures[name + "_X"], fixtures[name + "_y"], B, lr, epochs)


# ------------------------------------------------------------------ BASELINE config 1
def test_config_c1():
    X, y, w_true = make_regression(10000, 100, seed=42, fp16_round=True)
    w, mae = check_run(X, y, 1000, 0.1, 20)
    assert rel_l2(w, w_true) < 5e-4


def test_config_c1_run_pipeline_equals_iteration_loop():
    X, y, _ = make_regression(10000, 100, seed=42, fp16_round=True)
    w_o, perms, mae_o = oracle_epochs(X, y, 1000, 0.01, 10)
    with pkg.SGDEngine(10000, 100, 1000, 0.01, data=X, labels=y) as eng:
        t = eng.run(1, 10)                                   # overlapped shuffle pipeline
        assert t["steps"] == 100 and t["samples"] == 100000 and t["launches"] == 10
        assert np.array_equal(eng.permutation(), perms[-1])
        assert rel_l2(eng.weights, w_o) <= W_TOL
        assert abs(eng.calculate_mean_abs_error() - mae_o) <= MAE_TOL * mae_o


# ------------------------------------------------------------------ shapes and edge cases
@pytest.mark.parametrize("C", [1, 2, 3, 4, 5, 31, 32, 33, 100, 127, 128, 129, 255, 256, 300, 512, 513, 1000, 1024, 2048])
def test_feature_counts(C):
    R = 1500 if C <= 512 else 600
    X, y, _ = make_regression(R, C, seed=C)
    check_run(X, y, 250, 0.02, 2)


@pytest.mark.parametrize("C,B", [(2049, 250), (3000, 64), (4096, 600), (5000, 100), (8192, 33), (10000, 200), (16384, 50)])
def test_wide_feature_counts_column_mapping(C, B):
    # the reference's feature-scaling experiments go up to 10000 features (results/2000f-10000f-*)
    X, y, _ = make_regression(600, C, seed=C)
    check_run(X, y, B, 0.002, 2)


def test_more_than_16384_features_is_rejected():
    with pytest.raises(pkg.SgdError) as ei:
        pkg.SGDEngine(100, 16385, 10, 0.1)
    assert ei.value.code == _lib.SGD_ERR_ARG


def test_column_mapping_equals_row_mapping_at_2048(monkeypatch):
    X, y, _ = make_regression(3000, 2048, seed=5)
    w_row, _ = check_run(X, y, 500, 0.01, 2)
    monkeypatch.setenv("B200SGD_VARIANT", "3")
    w_col, _ = check_run(X, y, 500, 0.01, 2)
    assert rel_l2(w_col, w_row) < 1e-5


@pytest.mark.parametrize("R,B", [(1, 0), (2, 1), (7, 7), (1000, 1), (1000, 3), (1000, 999), (1000, 1000),
                                 (1000, 0), (4097, 64), (5000, 4096), (300, 301 - 1)])
def test_batch_shapes(R, B):
    X, y, _ = make_regression(R, 17, seed=R + B)
    check_run(X, y, B, 0.01, 2)


def test_batch_larger_than_rows_runs_zero_steps():
    # main.cu:344: floor(R / (float)B) = 0 batches; the reference still shuffles and reports the
    # error of the initial weights
    X, y, _ = make_regression(10, 3)
    with pkg.SGDEngine(10, 3, 11, 0.1, data=X, labels=y) as eng:
        assert eng.num_batches == 0
        t = eng.iteration(1)
        assert t["steps"] == 0 and t["samples"] == 0
        assert np.array_equal(eng.weights, orc.init_weights(3))
        perm = np.arange(10, dtype=np.int32)
        orc.Rand().shuffle(perm)
        assert np.array_equal(eng.permutation(), perm)
        assert eng.calculate_mean_abs_error() == pytest.approx(orc.mean_abs_error(X, y, orc.init_weights(3)), rel=1e-5)
    with pytest.raises(pkg.SgdError) as ei:
        pkg.SGDEngine(10, 3, -1, 0.1, data=X, labels=y)
    assert ei.value.code == _lib.SGD_ERR_ARG


def test_set_permutation_rejects_duplicates():
    with pkg.SGDEngine(100, 3, 10, 0.1) as eng:
        bad = np.arange(100, dtype=np.int32)
        bad[7] = 8
        with pytest.raises(pkg.SgdError) as ei:
            eng.set_permutation(bad)
        assert ei.value.code == _lib.SGD_ERR_ARG


@pytest.mark.parametrize("reduce", [_lib.SGD_REDUCE_ALLGATHER, _lib.SGD_REDUCE_SCATTER, _lib.SGD_REDUCE_DATAFLOW, _lib.SGD_REDUCE_DATAFLOW_1HOP])
@pytest.mark.parametrize("grid", [1, 2, 7, 37, 148])
def test_reduction_modes_and_grids(reduce, grid):
    X, y, _ = make_regression(6000, 100, seed=9)
    check_run(X, y, 500, 0.05, 2, reduce=reduce, grid=grid)


@pytest.mark.parametrize("C,B", [(100, 1000), (512, 256), (2048, 96)])
def test_graph_engine(C, B, monkeypatch):
    monkeypatch.setenv("B200SGD_CLUSTER", "0")   # the persistent engine on the same (generic) kernel
    X, y, _ = make_regression(3000, C, seed=C + 1)
    w_p, _ = check_run(X, y, B, 0.02, 2, engine=_lib.SGD_ENGINE_PERSISTENT, reduce=_lib.SGD_REDUCE_DATAFLOW)
    w_g, _ = check_run(X, y, B, 0.02, 2, engine=_lib.SGD_ENGINE_GRAPH, reduce=_lib.SGD_REDUCE_DATAFLOW)
    assert np.array_equal(w_p, w_g)     # same kernel, same order of operations


# ------------------------------------------------------------------ the cluster kernel (small steps, narrow rows)
@pytest.mark.parametrize("cluster", [8, 16])
@pytest.mark.parametrize("R,C,B", [
    (3000, 1, 48), (3000, 3, 100), (4000, 31, 333), (5000, 64, 777), (20000, 100, 1000),
    (9000, 127, 2048),      # 128 / 256 rows per CTA and step: four-tile blocks, two blocks per step at 8 CTAs
    (2500, 100, 1200),      # 2 steps, tail rows dropped
    (5000, 15, 49),         # fewer rows than lanes in most CTAs
])
def test_cluster_kernel_shapes(R, C, B, cluster, monkeypatch):
    monkeypatch.setenv("B200SGD_CLUSTER", str(cluster))
    X, y, _ = make_regression(R, C, seed=R + C + B)
    with pkg.SGDEngine(R, C, B, 0.02) as probe:
        assert probe.kernel_shape[0] == 0 and probe.cluster_size == cluster and probe.grid % cluster == 0
    check_run(X, y, B, 0.02, 3)


@pytest.mark.parametrize("R,C,B,clusters", [
    (20000, 100, 4096, 0),      # the launch heuristic picks the number of clusters (4 x 16 CTAs)
    (20000, 100, 1000, 2),      # forced: two clusters on a small step (some tiles nearly empty)
    (30000, 64, 10000, 8),      # 8 x 16 CTAs, 78 rows per CTA and step
    (9000, 7, 3000, 3),         # an odd number of clusters
    (150000, 100, 50000, 0),    # 390 rows per CTA: many four-tile blocks per step
])
def test_cluster_kernel_several_clusters(R, C, B, clusters, monkeypatch):
    # mid-size batches: every cluster sums its rows in shared memory, the clusters meet in an L2 inbox
    if clusters:
        monkeypatch.setenv("B200SGD_CLUSTERS", str(clusters))
    monkeypatch.setenv("B200SGD_CLUSTER_MAX_ROWS", "100000")   # beyond the default hand-over to sgd_epoch_kernel
    X, y, _ = make_regression(R, C, seed=R + C + B)
    with pkg.SGDEngine(R, C, B, 0.02) as probe:
        assert probe.kernel_shape[0] == 0 and probe.grid > probe.cluster_size
        if clusters and clusters <= 7:
            assert probe.grid == clusters * probe.cluster_size
    check_run(X, y, B, 0.02, 3)


def test_cluster_kernel_agrees_with_the_generic_kernel(monkeypatch):
    X, y, _ = make_regression(20000, 100, seed=77)
    w_c, mae_c = check_run(X, y, 1000, 0.05, 3)
    monkeypatch.setenv("B200SGD_CLUSTER", "0")
    with pkg.SGDEngine(20000, 100, 1000, 0.05) as probe:
        assert probe.kernel_shape[0] != 0
    w_g, mae_g = check_run(X, y, 1000, 0.05, 3)
    assert rel_l2(w_c, w_g) < 1e-5 and abs(mae_c - mae_g) <= 1e-5 * abs(mae_g) + 1e-6 * float(np.mean(np.abs(y)))


@pytest.mark.parametrize("R,C,B", [(60000, 100, 30000), (40000, 31, 20000), (70000, 127, 33333), (50000, 1, 25000)])
@pytest.mark.parametrize("variant,u", [("0", -4), ("4", -8)])
def test_narrow_rows_streaming_steps(R, C, B, variant, u, monkeypatch):
    # steps beyond the cluster kernel's reach: 8 lanes per row; one (default) or two rows per lane group
    monkeypatch.setenv("B200SGD_VARIANT", variant)
    X, y, _ = make_regression(R, C, seed=R + C)
    with pkg.SGDEngine(R, C, B, 0.02) as probe:
        assert probe.kernel_shape[0] == 1 and probe.kernel_shape[1] == u, probe.kernel_shape
    check_run(X, y, B, 0.02, 3)


def test_cluster_kernel_is_not_used_where_it_does_not_apply():
    for kw in (dict(C=128, B=1000), dict(C=100, B=24), dict(C=100, B=16384)):
        with pkg.SGDEngine(20000, kw["C"], kw["B"], 0.1) as eng:
            assert eng.kernel_shape[0] != 0, kw
    with pkg.SGDEngine(20000, 100, 1000, 0.1, keep_residuals=True) as eng:
        assert eng.kernel_shape[0] != 0
    with pkg.SGDEngine(20000, 100, 1000, 0.1, engine=_lib.SGD_ENGINE_GRAPH) as eng:
        assert eng.kernel_shape[0] != 0


# ------------------------------------------------------------------ the cluster step tail of the generic kernel
@pytest.mark.parametrize("R,C,B,cs,clusters", [
    (6000, 512, 2048, 8, 0),      # BASELINE config 5 row width: 16 float4 columns per CTA slice
    (3000, 2048, 600, 8, 0),      # config 4 row width, ~5 rows per CTA and step
    (3000, 2048, 600, 16, 0),     # 16 CTAs per cluster (non-portable size)
    (4000, 300, 1000, 4, 3),      # three clusters of four
    (4000, 300, 1000, 2, 5),      # pairs
    (5000, 129, 777, 8, 2),       # 33 float4 columns: ragged slices (5 per CTA, the last CTA gets none... 33 = 6*5 + 3)
    (9000, 100, 4500, 8, 4),      # narrow rows (8 lanes per row) beyond the cluster kernel's reach
    (1200, 5000, 200, 8, 0),      # column mapping (C > 2048)
    (600, 8192, 33, 4, 2),        # slices of 512 float4 columns: two passes of 256 threads
    (2000, 1000, 2000, 8, 1),     # ONE cluster: nothing goes through L2
])
def test_cluster_tail_shapes(R, C, B, cs, clusters, monkeypatch):
    monkeypatch.setenv("B200SGD_CLUSTER", "0")     # keep narrow rows off sgd_cluster_epoch_kernel
    monkeypatch.setenv("B200SGD_CLT", str(cs))
    if clusters:
        monkeypatch.setenv("B200SGD_CLUSTERS", str(clusters))
    X, y, _ = make_regression(R, C, seed=R + C + B)
    grid = clusters * cs if clusters else 0
    with pkg.SGDEngine(R, C, B, 0.01, grid=grid) as probe:
        assert probe.reduce == _lib.SGD_REDUCE_CLUSTER and probe.grid % cs == 0, (probe.reduce, probe.grid)
        if clusters:
            assert probe.grid == clusters * cs
    check_run(X, y, B, 0.01, 3, grid=grid)


def test_cluster_tail_agrees_with_the_dataflow_reduction(monkeypatch):
    X, y, _ = make_regression(8000, 512, seed=3)
    w_c, mae_c = check_run(X, y, 2000, 0.02, 3)
    with pkg.SGDEngine(8000, 512, 2000, 0.02) as probe:
        assert probe.reduce == _lib.SGD_REDUCE_CLUSTER
    w_d, mae_d = check_run(X, y, 2000, 0.02, 3, reduce=_lib.SGD_REDUCE_DATAFLOW)
    assert rel_l2(w_c, w_d) < 1e-5


@pytest.mark.parametrize("kb", [0, 16, 48, 1000])
def test_producers_in_flight_cap(kb, monkeypatch):
    # the cap only changes WHEN rows are requested, never the result
    monkeypatch.setenv("B200SGD_INFLIGHT_KB", str(kb))
    X, y, _ = make_regression(6000, 512, seed=8)
    w, _ = check_run(X, y, 3000, 0.02, 2)
    monkeypatch.setenv("B200SGD_INFLIGHT_KB", "64")
    w2, _ = check_run(X, y, 3000, 0.02, 2)
    assert np.array_equal(w, w2)


def test_graph_engine_many_steps_reuses_graph():
    X, y, _ = make_regression(9000, 20, seed=4)
    check_run(X, y, 10, 0.005, 2, engine=_lib.SGD_ENGINE_GRAPH)   # 900 steps > one graph chunk


def test_results_are_bit_reproducible():
    X, y, _ = make_regression(20000, 100, seed=11)
    runs = []
    for _ in range(2):
        with pkg.SGDEngine(20000, 100, 1000, 0.1, data=X, labels=y) as eng:
            eng.run(1, 3)
            runs.append(eng.weights)
    assert np.array_equal(runs[0], runs[1])


def test_linearity_of_the_gradient():
    # size-independent property: g(X, y1 + y2, w=0) = g(X, y1, 0) + g(X, y2, 0)
    X, y1, _ = make_regression(4096, 64, seed=1)
    _, y2, _ = make_regression(4096, 64, seed=2)
    gs = []
    for yy in (y1, y2, y1 + y2):
        with pkg.SGDEngine(4096, 64, 0, 0.0, data=X, labels=yy) as eng:
            eng.weights = np.zeros(64, np.float32)
            eng.set_permutation(np.arange(4096))
            eng.epoch(1)
            gs.append(eng.last_gradient().astype(np.float64))
    assert rel_l2(gs[0] + gs[1], gs[2]) < 1e-5


def test_running_loss_curve_matches_oracle_residuals():
    # extension: per-epoch sums of |r| and r^2 over the rows each step used
    X, y, _ = make_regression(6000, 100, seed=21)
    R, C, B, lr, epochs = 6000, 100, 500, 0.05, 3
    g = orc.Rand()
    ind = np.arange(R, dtype=np.int32)
    w = orc.init_weights(C)
    want_abs, want_sq = [], []
    for e in range(1, epochs + 1):
        g.shuffle(ind)
        sa = ss = 0.0
        a = orc.step_scale(lr, e)
        for b in range(R // B):
            rows = ind[b * B:(b + 1) * B]
            r = orc.residuals(X, y, w, rows).astype(np.float64)
            sa += np.abs(r).sum(); ss += (r * r).sum()
            gvec = orc.gradient(X, r.astype(np.float32), float(B), rows)
            w = (np.float32(a) * gvec + w).astype(np.float32)
        want_abs.append(sa); want_sq.append(ss)
    for engine in (_lib.SGD_ENGINE_PERSISTENT, _lib.SGD_ENGINE_GRAPH):
        with pkg.SGDEngine(R, C, B, lr, data=X, labels=y, engine=engine) as eng:
            eng.run(1, epochs)
            got_abs, got_sq = eng.loss_curve()
        assert got_abs.shape == (epochs,)
        assert np.allclose(got_abs, want_abs, rtol=2e-4) and np.allclose(got_sq, want_sq, rtol=4e-4)
        assert got_abs[0] > got_abs[-1]          # the loss goes down


# ------------------------------------------------------------------ synthetic generator
def test_synthetic_generator_properties_and_parity():
    R, C = 50000, 100
    with pkg.SGDEngine(R, C, 1000, 0.1) as eng:
        eng.load_synthetic(seed=42)
        X, y = eng.get_rows(0, R)
        wt = eng.true_weights()
        assert abs(X.mean()) < 0.01 and abs(X.std() - 1.0) < 0.01
        assert wt.min() >= 0 and wt.max() <= 100 and 30 < wt.mean() < 70
        assert rel_l2(y, X.astype(np.float64) @ wt.astype(np.float64)) < 1e-5
        w_o, perms, mae_o = oracle_epochs(X, y, 1000, 0.1, 3)
        eng.run(1, 3)
        assert np.array_equal(eng.permutation(), perms[-1])
        assert rel_l2(eng.weights, w_o) <= W_TOL
        assert abs(eng.calculate_mean_abs_error() - mae_o) <= MAE_TOL * mae_o + 1e-6 * float(np.mean(np.abs(y)))
    # a shard of a 4-rank layout generates the same rows (values do not depend on the rank count)
    with pkg.SGDEngine(R, C, 1000, 0.1, rank=2, nranks=4) as sh:
        sh.load_synthetic(seed=42)
        assert sh.local_row_begin == 25000 and sh.local_rows == 12500
        Xs, ys = sh.get_rows(25000, 12500)
        assert np.array_equal(Xs, X[25000:37500]) and np.array_equal(ys, y[25000:37500])


def test_synthetic_fp16_rounding_matches_generate_data():
    with pkg.SGDEngine(4000, 10, 100, 0.1) as eng:
        eng.load_synthetic(seed=7, round_fp16=True)     # generate_data.py:37
        X, y = eng.get_rows(0, 4000)
        assert np.array_equal(X, X.astype(np.float16).astype(np.float32))
        assert np.array_equal(y, y.astype(np.float16).astype(np.float32))


# ------------------------------------------------------------------ BASELINE config 2 at full size
@pytest.mark.slow
def test_config_c2_full_size_properties(kat):
    """4M x 100, B=1000 (BASELINE configs[1]): golden permutation checksum, convergence to the
    generating weights, and full-size parity with the OpenMP oracle on the same device-generated data."""
    R, C, B = 4000000, 100, 1000
    with pkg.SGDEngine(R, C, B, 0.1) as eng:
        eng.load_synthetic(seed=42)
        t = eng.run(1, 2)
        assert t["steps"] == 8000
        perm = eng.permutation()
        rec = [r for r in kat["shuffle_kat"] if r["R"] == R and r["epoch"] == 2][0]
        i = np.arange(1, R + 1, dtype=np.uint64)
        assert perm[:4].tolist() == rec["head"] and int(perm[-1]) == rec["last"]
        assert int((i * perm.astype(np.uint64)).sum(dtype=np.uint64)) == rec["checksum"]
        w = eng.weights
        mae = eng.calculate_mean_abs_error()
        assert rel_l2(w, eng.true_weights()) < 1e-3
        X, y = eng.get_rows(0, R)
    mae_o, w_o, ind_o = orc.run(X, y, B, 0.1, 2, threads=orc.max_threads())
    assert np.array_equal(perm, ind_o)
    assert rel_l2(w, w_o) <= W_TOL
    # the fit has converged to fp32 resolution of the labels (|y| ~ 450): the MAE is rounding noise
    assert abs(mae - mae_o) <= MAE_TOL * mae_o + 1e-6 * float(np.mean(np.abs(y)))


# ------------------------------------------------------------------ BASELINE configs 4 and 5 at full size
@pytest.mark.slow
@pytest.mark.parametrize("R,C,B,lr,tol", [
    (16777216, 2048, 4096, 0.1, 2e-3),      # config 4: 138.5 GB of records on one GPU, 4096 steps
    (16777216, 2048, 65536, 0.1, 0.9),      # config 4 at the streaming batch size: 256 steps (far from converged)
    (8388608, 512, 8192, 0.1, 2e-3),        # config 5's per-GPU shape at 8 GPUs (1024 steps of 8192 rows)
    (67108864, 512, 65536, 0.1, 0.6),       # config 5 on one GPU: 141.7 GB, 1024 steps of 65536 rows
])
def test_configs_4_and_5_full_size_properties(kat, R, C, B, lr, tol):
    """One epoch at full size: the golden permutation checksum (real glibc + libstdc++), the step count,
    finite weights that moved towards the generating ones, a mean absolute error consistent with them
    (an independent pass over the data), and the running loss of the epoch (summed inside the step
    kernel) equal to what the residual statistics imply."""
    with pkg.SGDEngine(R, C, B, lr) as eng:
        eng.load_synthetic(seed=42)
        w0 = eng.weights
        wt = eng.true_weights()
        mae0 = eng.calculate_mean_abs_error()
        t = eng.run(1, 1)
        assert t["steps"] == R // B and t["samples"] == (R // B) * B
        rec = [r for r in kat["shuffle_kat"] if r["R"] == R and r["epoch"] == 1]
        if rec:
            perm = eng.permutation()
            i = np.arange(1, R + 1, dtype=np.uint64)
            assert perm[:4].tolist() == rec[0]["head"] and int(perm[-1]) == rec[0]["last"]
            assert int((i * perm.astype(np.uint64)).sum(dtype=np.uint64)) == rec[0]["checksum"]
            del perm, i
        w = eng.weights
        assert np.all(np.isfinite(w))
        e0, e1 = rel_l2(w0, wt), rel_l2(w, wt)
        assert e1 < tol and e1 < 0.9 * e0, (e0, e1)
        mae = eng.calculate_mean_abs_error()
        assert mae < mae0 and np.isfinite(mae)
        # |X (w - w*)| for X ~ N(0, 1): the residual is N(0, |w - w*|^2), so its mean absolute value is
        # sqrt(2 / pi) |w - w*| (labels are y = X w* up to fp32 rounding of a C-term sum)
        want = np.sqrt(2 / np.pi) * float(np.linalg.norm(w.astype(np.float64) - wt))
        noise = 4e-7 * float(np.linalg.norm(wt)) * np.sqrt(C)
        assert abs(mae - want) <= 0.02 * want + noise, (mae, want)
        la, lq = eng.loss_curve()
        assert la.shape == (1,) and 0 < la[0] / t["samples"] < mae0 * 1.01      # running mean |r| of the epoch
