#!/usr/bin/env python
"""One-off parity check of a (R, C, B, grid, reduce) case against the CPU oracle on the GPU box.

    python tools/check_case.py R C B [grid] [reduce] [epochs]
Test infrastructure (uses oracle/); honours the B200SGD_* experiment switches.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import cuda_sgd_sese_project_b200 as pkg  # noqa: E402
from oracle import loader as orc  # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:]]
    R, C, B = a[0], a[1], a[2]
    grid = a[3] if len(a) > 3 else 0
    reduce = a[4] if len(a) > 4 else 0
    epochs = a[5] if len(a) > 5 else 3
    rng = np.random.default_rng(7)
    X = rng.standard_normal((R, C)).astype(np.float32)
    w_true = (100.0 * rng.random(C)).astype(np.float32)
    y = (X @ w_true).astype(np.float32)
    mae_o, w_o, ind_o = orc.run(X, y, B, 0.05, epochs)
    with pkg.SGDEngine(R, C, B, 0.05, data=X, labels=y, grid=grid, reduce=reduce) as eng:
        eng.run(1, epochs)
        w = eng.weights
        perm = eng.permutation()
        mae = eng.calculate_mean_abs_error()
        shape = eng.kernel_shape
        g = eng.grid
    rel = float(np.linalg.norm(w.astype(np.float64) - w_o) / np.linalg.norm(w_o))
    ok = bool(np.array_equal(perm, ind_o)) and rel <= 1e-4 and abs(mae - mae_o) <= 1e-4 * mae_o + 1e-6 * float(np.abs(y).mean())
    print(json.dumps({"case": a, "kernel": shape, "grid": g, "rel_l2_w": rel, "mae": mae, "mae_oracle": mae_o,
                      "perm_exact": bool(np.array_equal(perm, ind_o)), "ok": ok,
                      "variant": os.environ.get("B200SGD_VARIANT")}), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
