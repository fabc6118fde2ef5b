"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, from __graft_entry__.smoke() and from the
cpu_baseline / --impl reference legs of bench.py -- never from the product package
(cuda_sgd_sese_project_b200), which must fail loudly when its CUDA library is missing instead of
falling back to anything in here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class RandState(C.Structure):
    _fields_ = [("st", C.c_uint32 * 34), ("n", C.c_uint64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sgd_oracle.c")
    hdr = os.path.join(_HERE, "sgd_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    L.oracle_rand_init.argtypes = [C.POINTER(RandState), C.c_uint32]
    L.oracle_rand_next.argtypes = [C.POINTER(RandState)]
    L.oracle_rand_next.restype = C.c_int32
    L.oracle_random_shuffle.argtypes = [C.POINTER(RandState), _i32p, C.c_int64]
    L.oracle_init_weights.argtypes = [_f32p, C.c_int32]
    L.oracle_effective_batch.argtypes = [C.c_int32, C.c_int32]
    L.oracle_effective_batch.restype = C.c_int32
    L.oracle_num_batches.argtypes = [C.c_int32, C.c_int32]
    L.oracle_num_batches.restype = C.c_int32
    L.oracle_step_scale.argtypes = [C.c_float, C.c_int32]
    L.oracle_step_scale.restype = C.c_float
    L.oracle_residuals.argtypes = [_f32p, _f32p, _f32p, C.c_void_p, C.c_int32, C.c_int32, _f32p]
    L.oracle_gradient.argtypes = [_f32p, _f32p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, _f32p]
    L.oracle_epoch.argtypes = [_f32p, _f32p, C.c_int64, C.c_int32, C.c_int32, _i32p, _f32p,
                               C.c_float, C.c_int32, C.c_int, C.c_int]
    L.oracle_epoch_sharded.argtypes = [_f32p, _f32p, C.c_int64, C.c_int32, C.c_int32, _i32p, _f32p,
                                       C.c_float, C.c_int32, C.c_int32]
    L.oracle_mean_abs_error.argtypes = [_f32p, _f32p, _f32p, C.c_int64, C.c_int32, C.c_int]
    L.oracle_mean_abs_error.restype = C.c_float
    L.oracle_run.argtypes = [_f32p, _f32p, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_int32,
                             C.c_int, C.c_int, _f32p, C.c_void_p]
    L.oracle_run.restype = C.c_float
    L.oracle_read_csv.argtypes = [C.c_char_p, C.c_int64, C.c_int32, _f32p, _f32p]
    L.oracle_read_csv.restype = C.c_int
    L.oracle_permute_rows.argtypes = [_f32p, _f32p, _i32p, C.c_int64, C.c_int32, _f32p, _f32p]
    L.oracle_scale_rows.argtypes = [_f32p, _f32p, C.c_int64, C.c_int32, _f32p]
    L.oracle_col_sums.argtypes = [_f32p, C.c_int64, C.c_int32, C.c_float, _f32p]
    L.oracle_max_threads.restype = C.c_int
    _lib = L
    return L


# ---- thin numpy-level helpers (what the tests call) -------------------------------------------

def _c32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class Rand:
    """glibc rand() stream, seed 1 unless told otherwise."""

    def __init__(self, seed: int = 1):
        self.state = RandState()
        lib().oracle_rand_init(C.byref(self.state), seed)

    def next(self) -> int:
        return int(lib().oracle_rand_next(C.byref(self.state)))

    def shuffle(self, ind: np.ndarray) -> None:
        assert ind.dtype == np.int32 and ind.flags.c_contiguous
        lib().oracle_random_shuffle(C.byref(self.state), ind, ind.shape[0])


def init_weights(Cn: int) -> np.ndarray:
    w = np.empty(Cn, dtype=np.float32)
    lib().oracle_init_weights(w, Cn)
    return w


def num_batches(R: int, B_arg: int):
    B = lib().oracle_effective_batch(R, B_arg)
    return B, lib().oracle_num_batches(R, B)


def step_scale(lr: float, epoch: int) -> float:
    return float(lib().oracle_step_scale(lr, epoch))


def residuals(X, y, w, rows=None) -> np.ndarray:
    X = _c32(X); y = _c32(y); w = _c32(w)
    Cn = X.shape[1]
    if rows is None:
        n, ptr = X.shape[0], None
    else:
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        n, ptr = rows.shape[0], rows.ctypes.data
    out = np.empty(n, dtype=np.float32)
    lib().oracle_residuals(X, y, w, ptr, n, Cn, out)
    return out


def gradient(X, r, denom, rows=None) -> np.ndarray:
    X = _c32(X); r = _c32(r)
    Cn = X.shape[1]
    if rows is None:
        n, ptr = X.shape[0], None
    else:
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        n, ptr = rows.shape[0], rows.ctypes.data
    out = np.empty(Cn, dtype=np.float32)
    lib().oracle_gradient(X, r, ptr, n, Cn, float(denom), out)
    return out


def epoch(X, y, B, ind, w, lr, epoch_no, acc64=False, threads=1) -> None:
    """In-place update of w over one epoch with permutation `ind`."""
    X = _c32(X); y = _c32(y)
    assert w.dtype == np.float32 and w.flags.c_contiguous
    R, Cn = X.shape
    lib().oracle_epoch(X, y, R, Cn, B, np.ascontiguousarray(ind, dtype=np.int32), w, lr, epoch_no,
                       int(acc64), threads)


def epoch_sharded(X, y, B, ind, w, lr, epoch_no, nranks) -> None:
    X = _c32(X); y = _c32(y)
    R, Cn = X.shape
    lib().oracle_epoch_sharded(X, y, R, Cn, B, np.ascontiguousarray(ind, dtype=np.int32), w, lr,
                               epoch_no, nranks)


def mean_abs_error(X, y, w, threads=1) -> float:
    X = _c32(X); y = _c32(y); w = _c32(w)
    R, Cn = X.shape
    return float(lib().oracle_mean_abs_error(X, y, w, R, Cn, threads))


def run(X, y, B_arg, lr, epochs, acc64=False, threads=1):
    """Whole reference run (main.cu:338-530 minus I/O) -> (mae, weights, last permutation)."""
    X = _c32(X); y = _c32(y)
    R, Cn = X.shape
    w = np.empty(Cn, dtype=np.float32)
    ind = np.empty(R, dtype=np.int32)
    mae = lib().oracle_run(X, y, R, Cn, B_arg, lr, epochs, int(acc64), threads, w, ind.ctypes.data)
    return float(mae), w, ind


def read_csv(path: str, R: int, Cn: int):
    X = np.zeros((R, Cn), dtype=np.float32)
    y = np.zeros(R, dtype=np.float32)
    rc = lib().oracle_read_csv(os.fsencode(path), R, Cn, X, y)
    return rc, X, y


def permute_rows(X, y, order):
    X = _c32(X); y = _c32(y)
    R, Cn = X.shape
    Xs = np.empty_like(X); ys = np.empty_like(y)
    lib().oracle_permute_rows(X, y, np.ascontiguousarray(order, dtype=np.int32), R, Cn, Xs, ys)
    return Xs, ys


def scale_rows(X, v):
    X = _c32(X); v = _c32(v)
    out = np.empty_like(X)
    lib().oracle_scale_rows(X, v, X.shape[0], X.shape[1], out)
    return out


def col_sums(G, scale=1.0):
    G = _c32(G)
    out = np.empty(G.shape[1], dtype=np.float32)
    lib().oracle_col_sums(G, G.shape[0], G.shape[1], float(scale), out)
    return out


def max_threads() -> int:
    return int(lib().oracle_max_threads())
