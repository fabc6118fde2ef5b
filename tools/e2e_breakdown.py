#!/usr/bin/env python
"""Host-side timing of the end-to-end call sequence of bench.py's e2e leg (one GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import cuda_sgd_sese_project_b200 as pkg

R, C, B, K = 4_000_000, 100, 1000, 10
Xh = torch.empty((R, C), dtype=torch.float32, pin_memory=True)
yh = torch.empty((R,), dtype=torch.float32, pin_memory=True)
with pkg.SGDEngine(R, C, B, 0.1) as g:
    g.load_synthetic(seed=42)
    X, y = g.get_rows(0, R)
    Xh.numpy()[:] = X
    yh.numpy()[:] = y
    del X, y
for it in range(3):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    e = pkg.SGDEngine(R, C, B, 0.1); t.append(time.perf_counter())
    e.load_rows_ptr(0, R, Xh.data_ptr(), yh.data_ptr()); t.append(time.perf_counter())
    r = e.run(1, K); t.append(time.perf_counter())
    w = e.weights; t.append(time.perf_counter())
    s = e.abs_error_sum(); t.append(time.perf_counter())
    e.close(); t.append(time.perf_counter())
    names = ["create", "load_rows", "run(10 epochs)", "weights", "mae", "close"]
    print({n: round((t[i + 1] - t[i]) * 1e3, 2) for i, n in enumerate(names)}, "total ms", round((t[-1] - t[0]) * 1e3, 1), flush=True)
