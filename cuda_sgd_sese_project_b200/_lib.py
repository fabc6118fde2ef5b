"""ctypes binding of libb200sgd.so (include/b200sgd.h).  No torch types cross this boundary.

The library is built in-tree by `__graft_entry__.build()` / `csrc/build.sh`.  If it is missing this
module raises: there is no Python, numpy or CPU fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200sgd.so")

SGD_OK = 0
SGD_ERR_ARG, SGD_ERR_IO, SGD_ERR_PARSE, SGD_ERR_CUDA, SGD_ERR_STATE, SGD_ERR_NOMEM, SGD_ERR_COMM = (
    -1, -2, -3, -4, -5, -6, -7)
SGD_PATH_CUBLAS, SGD_PATH_PLAIN_CUDA, SGD_PATH_LVL1 = 0, 1, 2
SGD_ENGINE_AUTO, SGD_ENGINE_PERSISTENT, SGD_ENGINE_GRAPH = 0, 1, 2
SGD_REDUCE_AUTO, SGD_REDUCE_ALLGATHER, SGD_REDUCE_SCATTER, SGD_REDUCE_DATAFLOW, SGD_REDUCE_DATAFLOW_1HOP = 0, 1, 2, 3, 4
SGD_REDUCE_CLUSTER = 5
SGD_FLAG_HOST_SHUFFLE, SGD_FLAG_RESIDUALS = 0x1, 0x2
SGD_PEER_HANDLE_BYTES = 128

# every symbol include/b200sgd.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = [
    "sgd_abi_version", "sgd_last_error", "sgd_device_count", "sgd_read_csv", "sgd_write_csv",
    "sgd_get_filename_from_path", "sgd_write_experiment_output", "sgd_init_weights",
    "sgd_batch_geometry", "sgd_host_shuffle", "sgd_shard_range", "sgd_host_bucket", "sgd_get_local_schedule", "sgd_get_loss_curve",
    "sgd_create", "sgd_destroy", "sgd_load_rows", "sgd_load_csv",
    "sgd_load_synthetic", "sgd_get_rows", "sgd_get_true_weights", "sgd_get_weights",
    "sgd_set_weights", "sgd_get_config", "sgd_local_rows", "sgd_local_row_begin", "sgd_shuffle",
    "sgd_set_permutation", "sgd_get_permutation", "sgd_epoch", "sgd_iteration", "sgd_run",
    "sgd_get_residuals", "sgd_get_last_gradient", "sgd_mean_abs_error", "sgd_debug_phase_cycles",
    "sgd_peer_export",
    "sgd_peer_import", "sgd_peer_ready",
]


class SgdConfig(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("cols", C.c_int32), ("batch", C.c_int32),
        ("learning_rate", C.c_float), ("codepath", C.c_int32), ("device", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32), ("engine", C.c_int32), ("reduce", C.c_int32),
        ("grid", C.c_int32), ("shuffle_seed", C.c_uint32), ("flags", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class SgdTimings(C.Structure):
    _fields_ = [
        ("shuffle_ms", C.c_float), ("steploop_ms", C.c_float), ("epoch_ms", C.c_float),
        ("steps", C.c_int32), ("launches", C.c_int32), ("samples", C.c_int64), ("bytes", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SgdLoadStats(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("bytes_uploaded", C.c_int64), ("parse_seconds", C.c_double),
        ("total_seconds", C.c_double), ("from_cache", C.c_int32), ("cache_written", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


SGD_CSV_CACHE = 0x1


class SgdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200sgd error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Loads libb200sgd.so or raises -- never substitutes anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(cuda_sgd_sese_project_b200 has no fallback implementation)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    sig = {
        "sgd_abi_version": ([], C.c_int),
        "sgd_last_error": ([], C.c_char_p),
        "sgd_device_count": ([], C.c_int),
        "sgd_read_csv": ([C.c_char_p, i64, i32, vp, vp], C.c_int),
        "sgd_write_csv": ([C.c_char_p, i64, i32, vp, vp], C.c_int),
        "sgd_get_filename_from_path": ([C.c_char_p, C.c_char_p, C.c_size_t], C.c_int),
        "sgd_write_experiment_output": ([C.c_char_p, f32, f32, f32, f32, f32], C.c_int),
        "sgd_init_weights": ([vp, i32], C.c_int),
        "sgd_batch_geometry": ([i64, i32, C.POINTER(i32), C.POINTER(i32)], C.c_int),
        "sgd_host_shuffle": ([vp, i64, C.c_uint32, i32, i32], C.c_int),
        "sgd_shard_range": ([i64, i32, i32, C.POINTER(i64), C.POINTER(i64)], C.c_int),
        "sgd_host_bucket": ([vp, i64, i32, i32, i32, vp, i64, vp], C.c_int),
        "sgd_get_local_schedule": ([vp, vp, i64, vp], C.c_int),
        "sgd_get_loss_curve": ([vp, vp, vp, i32], C.c_int),
        "sgd_create": ([C.POINTER(SgdConfig), vp, vp, C.POINTER(vp)], C.c_int),
        "sgd_destroy": ([vp], None),
        "sgd_load_rows": ([vp, i64, i64, vp, vp], C.c_int),
        "sgd_load_csv": ([vp, C.c_char_p, i32, C.POINTER(SgdLoadStats)], C.c_int),
        "sgd_load_synthetic": ([vp, C.c_uint64, i32], C.c_int),
        "sgd_get_rows": ([vp, i64, i64, vp, vp], C.c_int),
        "sgd_get_true_weights": ([vp, vp], C.c_int),
        "sgd_get_weights": ([vp, vp], C.c_int),
        "sgd_set_weights": ([vp, vp], C.c_int),
        "sgd_get_config": ([vp, C.POINTER(SgdConfig)], C.c_int),
        "sgd_local_rows": ([vp], i64),
        "sgd_local_row_begin": ([vp], i64),
        "sgd_shuffle": ([vp], C.c_int),
        "sgd_set_permutation": ([vp, vp], C.c_int),
        "sgd_get_permutation": ([vp, vp], C.c_int),
        "sgd_epoch": ([vp, i32, C.POINTER(SgdTimings)], C.c_int),
        "sgd_iteration": ([vp, i32, C.POINTER(SgdTimings)], C.c_int),
        "sgd_run": ([vp, i32, i32, C.POINTER(SgdTimings)], C.c_int),
        "sgd_get_residuals": ([vp, vp, i64], C.c_int),
        "sgd_get_last_gradient": ([vp, vp], C.c_int),
        "sgd_mean_abs_error": ([vp, C.POINTER(f32), C.POINTER(C.c_double)], C.c_int),
        "sgd_debug_phase_cycles": ([vp, vp, i32], C.c_int),
        "sgd_peer_export": ([vp, vp], C.c_int),
        "sgd_peer_import": ([vp, i32, vp], C.c_int),
        "sgd_peer_ready": ([vp], C.c_int),
    }
    assert sorted(sig) == sorted(EXPORTS)
    for name, (argtypes, restype) in sig.items():
        fn = getattr(L, name)  # AttributeError if the library does not export it
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != SGD_OK:
        raise SgdError(rc, (load().sgd_last_error() or b"").decode(errors="replace"))
