"""Shared helpers for the parity tests: synthetic data in the reference's distribution and an
INDEPENDENT numpy restatement of Appendix A of SURVEY.md (vectorised, so its summation order
differs from both the C oracle and the GPU kernels)."""
from __future__ import annotations

import numpy as np


def make_regression(R: int, C: int, seed: int = 42, fp16_round: bool = False, noise: float = 0.0):
    """Same distribution as data_set_creation/generate_data.py:22-39 (sklearn make_regression with
    all features informative, bias 0): X ~ N(0,1), w_true ~ 100*U(0,1), y = X w_true (+ noise).
    generate_data.py:37 stores the table as float16 -> fp16_round=True; generate_scalable_data.py
    :30-32 keeps fp32."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((R, C))
    w_true = 100.0 * rng.random(C)
    y = X @ w_true
    if noise:
        y = y + noise * rng.standard_normal(R)
    if fp16_round:
        X = X.astype(np.float16)
        y = y.astype(np.float16)
    return X.astype(np.float32), y.astype(np.float32), w_true


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def numpy_run(X, y, B_arg, lr, epochs, perms, w0):
    """Appendix A in float64 numpy given the per-epoch permutations and initial weights."""
    R, C = X.shape
    B = R if B_arg == 0 else B_arg
    nb = int(np.floor(np.float32(R) / np.float32(B)))
    w = np.asarray(w0, dtype=np.float64).copy()
    X64 = X.astype(np.float64)
    y64 = y.astype(np.float64)
    for e in range(1, epochs + 1):
        ind = perms[e - 1]
        a = float(np.float32(-(float(np.float32(lr)) / (float(e) ** 0.25))))
        for b in range(nb):
            rows = ind[b * B:(b + 1) * B]
            Xb = X64[rows]
            r = Xb @ w - y64[rows]
            g = (Xb.T @ r) / B
            w = a * g + w
    mae = float(np.mean(np.abs(X64 @ w - y64)))
    return mae, w
