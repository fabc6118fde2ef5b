#!/usr/bin/env python
"""Tiny end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck): row mapping,
column mapping, device shuffle, MAE, both engines.  Checks parity with the oracle as well."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cuda_sgd_sese_project_b200 as pkg
from oracle import loader as orc
from tests.helpers import make_regression, rel_l2

for (R, C, B, grid, engine) in [(600, 100, 50, 4, 1), (600, 100, 50, 4, 2), (300, 2304, 30, 3, 1), (40000, 8, 1000, 8, 1)]:
    X, y, _ = make_regression(R, C, seed=1)
    mae_o, w_o, ind_o = orc.run(X, y, B, 0.01, 2)
    with pkg.SGDEngine(R, C, B, 0.01, data=X, labels=y, grid=grid, engine=engine) as eng:
        eng.run(1, 2)
        assert np.array_equal(eng.permutation(), ind_o)
        assert rel_l2(eng.weights, w_o) < 1e-4
        mae = eng.calculate_mean_abs_error()
    print("ok", R, C, B, grid, engine, mae, mae_o, flush=True)
