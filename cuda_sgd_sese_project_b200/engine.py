"""Host-side mirror of the reference's interface for the SGD hot path, over the C-ABI.

Names, argument meaning and error behaviour follow /root/reference/c_code (main.cu, sgd_io.cuh,
sgd_thrust.cuh) so that the parity tests read like the reference's own (manual) tests:

    read_csv, get_filename_from_path, write_experiment_output   <-> sgd_io.cuh:28-91
    SGDEngine.cuda_iteration / cublas_iteration                  <-> main.cu:56, :168
    SGDEngine.calculate_mean_abs_error                           <-> sgd_thrust.cu:79-135
    run_main(argv)                                               <-> main.cu:315-571

All arrays are numpy (host buffers); PyTorch is not needed here at all.  Compute happens in
libb200sgd.so only -- there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------- L1 layer
def read_csv(filename: str, R: int, C_: int):
    """sgd_io.cu:38-77.  Returns (ok, data[R,C], labels[R]); ok False <=> file could not be opened
    (the reference returns false); raises SgdError(PARSE) where the reference would overrun."""
    X = np.zeros((R, C_), dtype=np.float32)
    y = np.zeros(R, dtype=np.float32)
    rc = L.load().sgd_read_csv(os.fsencode(filename), R, C_, _ptr(X), _ptr(y))
    if rc == L.SGD_ERR_IO:
        return False, X, y
    L.check(rc)
    return True, X, y


def write_csv(filename: str, data, labels) -> None:
    """`f1,...,fC,label` lines (the format of data_set_creation/generate_data.py:52-59), multi-threaded."""
    X = _f32(data)
    y = _f32(labels).reshape(-1)
    X = X.reshape(y.shape[0], -1)
    L.check(L.load().sgd_write_csv(os.fsencode(filename), X.shape[0], X.shape[1], _ptr(X), _ptr(y)))


def get_filename_from_path(filepath: str) -> str:
    """sgd_io.cu:107-113."""
    buf = C.create_string_buffer(4096)
    L.check(L.load().sgd_get_filename_from_path(os.fsencode(filepath), buf, len(buf)))
    return buf.value.decode()


def write_experiment_output(filepath: str, shuffle_time: float, total_gradient_time: float,
                            gpu_time: float, transfer_time: float, error: float) -> None:
    """ExperimentOutput + write_experiment_output, sgd_io.cuh:28-60 / sgd_io.cu:93-105."""
    L.check(L.load().sgd_write_experiment_output(os.fsencode(filepath), shuffle_time,
                                                 total_gradient_time, gpu_time, transfer_time, error))


def init_weights(C_: int) -> np.ndarray:
    """main.cu:389-395."""
    w = np.empty(C_, dtype=np.float32)
    L.check(L.load().sgd_init_weights(_ptr(w), C_))
    return w


def batch_geometry(R: int, batch_arg: int):
    """main.cu:343-344 -> (batchsize, num_batches)."""
    b, n = C.c_int32(), C.c_int32()
    L.check(L.load().sgd_batch_geometry(R, batch_arg, C.byref(b), C.byref(n)))
    return b.value, n.value


def host_shuffle(ind: np.ndarray, seed: int = 1, skip_epochs: int = 0, epochs: int = 1) -> np.ndarray:
    """`epochs` cumulative std::random_shuffle passes (main.cu:90 / :207) over ind, in place."""
    assert ind.dtype == np.int32 and ind.flags.c_contiguous
    L.check(L.load().sgd_host_shuffle(_ptr(ind), ind.shape[0], seed, skip_epochs, epochs))
    return ind


def shard_range(R: int, rank: int, nranks: int):
    """Rows [begin, end) owned by `rank` of `nranks`."""
    b, e = C.c_int64(), C.c_int64()
    L.check(L.load().sgd_shard_range(R, rank, nranks, C.byref(b), C.byref(e)))
    return b.value, e.value


def host_bucket(perm: np.ndarray, batch_arg: int, rank: int, nranks: int):
    """Host statement of the per-epoch bucketing: (local_perm, step_off) of `rank`."""
    perm = np.ascontiguousarray(perm, dtype=np.int32)
    R = perm.shape[0]
    _, nb = batch_geometry(R, batch_arg)
    b, e = shard_range(R, rank, nranks)
    local = np.empty(max(1, e - b), dtype=np.int32)
    off = np.empty(nb + 1, dtype=np.int64)
    L.check(L.load().sgd_host_bucket(_ptr(perm), R, batch_arg, rank, nranks, _ptr(local), local.shape[0], _ptr(off)))
    return local[: off[nb]].copy(), off


def device_count() -> int:
    return int(L.load().sgd_device_count())


# ---------------------------------------------------------------------------------- engine
class SGDEngine:
    """Owns one libb200sgd context = the state main() keeps between main.cu:381 and :530."""

    def __init__(self, R: int, C_: int, batchsize: int, learning_rate: float, *, data=None,
                 labels=None, cudaRun: int = 0, device: int = 0, rank: int = 0, nranks: int = 1,
                 engine: int = L.SGD_ENGINE_AUTO, reduce: int = L.SGD_REDUCE_AUTO, grid: int = 0,
                 shuffle_seed: int = 1, keep_residuals: bool = False):
        lib = L.load()
        cfg = L.SgdConfig()
        cfg.rows, cfg.cols, cfg.batch = R, C_, batchsize
        cfg.learning_rate = learning_rate
        cfg.codepath = cudaRun
        cfg.device, cfg.rank, cfg.nranks = device, rank, nranks
        cfg.engine, cfg.reduce, cfg.grid = engine, reduce, grid
        cfg.shuffle_seed = shuffle_seed
        cfg.flags = L.SGD_FLAG_RESIDUALS if keep_residuals else 0
        self._h = C.c_void_p()
        self.R, self.C = R, C_
        X = y = None
        if data is not None:
            X = _f32(data).reshape(R, C_)
            y = _f32(labels).reshape(R)
        L.check(lib.sgd_create(C.byref(cfg), _ptr(X) if X is not None else None,
                               _ptr(y) if y is not None else None, C.byref(self._h)))
        out = L.SgdConfig()
        L.check(lib.sgd_get_config(self._h, C.byref(out)))
        self.batchsize = out.batch
        self.num_batches = out.reserved[0]
        self.grid, self.reduce, self.engine = out.grid, out.reduce, out.engine
        self.nslots, self.smem_bytes = out.reserved[1], out.reserved[2]
        self.kernel_shape = (out.reserved[3], out.reserved[4], out.reserved[5] & 0xff)  # KC, U, NW; KC = 0: cluster kernel (U = cluster size)
        self.colsplit = out.reserved[5] >> 8   # COLSPLIT mapping: consumer warps along the columns (0: other mappings)
        self.cluster_size = out.reserved[4] if out.reserved[3] == 0 else 0             # CTAs of the one cluster, or 0
        self.ld = out.reserved[6]
        self.rank, self.nranks = rank, nranks
        self.local_rows = int(lib.sgd_local_rows(self._h))
        self.local_row_begin = int(lib.sgd_local_row_begin(self._h))

    # -- lifetime
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            L.load().sgd_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data
    def load_rows(self, row_begin: int, data, labels) -> None:
        X = _f32(data)
        y = _f32(labels)
        n = y.shape[0]
        assert X.size == n * self.C
        L.check(L.load().sgd_load_rows(self._h, row_begin, n, _ptr(X), _ptr(y)))

    def load_rows_ptr(self, row_begin: int, nrows: int, data_ptr: int, labels_ptr: int) -> None:
        """Raw host pointers (e.g. pinned torch tensors' data_ptr())."""
        L.check(L.load().sgd_load_rows(self._h, row_begin, nrows, C.c_void_p(data_ptr), C.c_void_p(labels_ptr)))

    def load_csv(self, filename: str, cache: bool = False) -> dict:
        """read_csv + upload in one overlapped pass (main.cu:365 + :381-384); optional raw cache `<csv>.f32`."""
        st = L.SgdLoadStats()
        L.check(L.load().sgd_load_csv(self._h, os.fsencode(filename), L.SGD_CSV_CACHE if cache else 0, C.byref(st)))
        return st.as_dict()

    def load_synthetic(self, seed: int = 42, round_fp16: bool = False) -> None:
        L.check(L.load().sgd_load_synthetic(self._h, seed, int(round_fp16)))

    def get_rows(self, row_begin: int, nrows: int):
        X = np.empty((nrows, self.C), dtype=np.float32)
        y = np.empty(nrows, dtype=np.float32)
        L.check(L.load().sgd_get_rows(self._h, row_begin, nrows, _ptr(X), _ptr(y)))
        return X, y

    def true_weights(self) -> np.ndarray:
        w = np.empty(self.C, dtype=np.float32)
        L.check(L.load().sgd_get_true_weights(self._h, _ptr(w)))
        return w

    @property
    def weights(self) -> np.ndarray:
        w = np.empty(self.C, dtype=np.float32)
        L.check(L.load().sgd_get_weights(self._h, _ptr(w)))
        return w

    @weights.setter
    def weights(self, w) -> None:
        w = _f32(w)
        assert w.shape == (self.C,)
        L.check(L.load().sgd_set_weights(self._h, _ptr(w)))

    # -- permutation
    def shuffle(self) -> None:
        """std::random_shuffle(ind_vector); batch_indices = ind_vector  (main.cu:90-91 / :207-211)."""
        L.check(L.load().sgd_shuffle(self._h))

    def set_permutation(self, order) -> None:
        order = np.ascontiguousarray(order, dtype=np.int32)
        assert order.shape == (self.R,)
        L.check(L.load().sgd_set_permutation(self._h, _ptr(order)))

    def permutation(self) -> np.ndarray:
        p = np.empty(self.R, dtype=np.int32)
        L.check(L.load().sgd_get_permutation(self._h, _ptr(p)))
        return p

    # -- hot path
    def epoch(self, epoch: int) -> dict:
        """The batch loop only (main.cu:94-160 / :227-300) with the current permutation."""
        t = L.SgdTimings()
        L.check(L.load().sgd_epoch(self._h, epoch, C.byref(t)))
        return t.as_dict()

    def iteration(self, epoch: int) -> dict:
        t = L.SgdTimings()
        L.check(L.load().sgd_iteration(self._h, epoch, C.byref(t)))
        return t.as_dict()

    # the reference's two entry points; both are the same fused kernel here
    def cuda_iteration(self, epoch: int) -> dict:      # main.cu:56
        return self.iteration(epoch)

    def cublas_iteration(self, epoch: int) -> dict:    # main.cu:168
        return self.iteration(epoch)

    def run(self, first_epoch: int, num_epochs: int) -> dict:
        """main.cu:456-508 with the next permutation produced while the current epoch runs."""
        t = L.SgdTimings()
        L.check(L.load().sgd_run(self._h, first_epoch, num_epochs, C.byref(t)))
        return t.as_dict()

    def residuals(self, count: int | None = None) -> np.ndarray:
        n = self.num_batches * self.batchsize if count is None else count
        r = np.empty(n, dtype=np.float32)
        L.check(L.load().sgd_get_residuals(self._h, _ptr(r), n))
        return r

    def last_gradient(self) -> np.ndarray:
        g = np.empty(self.C, dtype=np.float32)
        L.check(L.load().sgd_get_last_gradient(self._h, _ptr(g)))
        return g

    def calculate_mean_abs_error(self) -> float:
        """sgd_thrust.cu:79-135 (single rank)."""
        mae = C.c_float()
        L.check(L.load().sgd_mean_abs_error(self._h, C.byref(mae), None))
        return float(mae.value)

    def abs_error_sum(self) -> float:
        s = C.c_double()
        L.check(L.load().sgd_mean_abs_error(self._h, None, C.byref(s)))
        return float(s.value)

    def loss_curve(self, max_epochs: int = 1 << 20):
        """(sum |r|, sum r^2) per executed epoch (most recent `max_epochs`), this rank's rows."""
        a = np.zeros(min(max_epochs, 1 << 20), dtype=np.float64)
        b = np.zeros_like(a)
        n = L.load().sgd_get_loss_curve(self._h, _ptr(a), _ptr(b), a.shape[0])
        if n < 0:
            L.check(n)
        return a[:n].copy(), b[:n].copy()

    def local_schedule(self):
        """nranks > 1: (local_perm, step_off) of the last executed epoch as built on the device."""
        local = np.empty(max(1, self.local_rows), dtype=np.int32)
        off = np.empty(self.num_batches + 1, dtype=np.int64)
        L.check(L.load().sgd_get_local_schedule(self._h, _ptr(local), local.shape[0], _ptr(off)))
        return local[: off[self.num_batches]].copy(), off

    @property
    def kernel_name(self) -> str:
        kc, u, nw = self.kernel_shape
        if kc == 0 and u > 0:
            return "sgd_cluster_epoch_kernel (%d cluster%s of %d CTAs, %d consumer warps)" % (
                self.grid // u, "" if self.grid == u else "s", u, nw)
        return "sgd_epoch_kernel<KC=%d,U=%d,NW=%d%s%s%s%s>" % (
            abs(kc), abs(u), nw, ",WIDE" if kc < 0 else "", ",NARROW8" if u < 0 else "",
            ",CW=%d" % self.colsplit if self.colsplit else "", ",CLT" if self.reduce == L.SGD_REDUCE_CLUSTER else "")

    def debug_phase_cycles(self) -> np.ndarray:
        """[grid, 8] cycle sums per phase of the last launch (needs B200SGD_PHASE_PROFILE=1)."""
        out = np.zeros((2 * self.grid, 8), dtype=np.uint64)
        n = L.load().sgd_debug_phase_cycles(self._h, _ptr(out), 2 * self.grid)
        if n < 0:
            L.check(n)
        return out[:self.grid]

    def debug_step_timeline(self) -> np.ndarray:
        """[grid, 8] globaltimer (ns) at the phase marks of the middle step of the last launch."""
        out = np.zeros((2 * self.grid, 8), dtype=np.uint64)
        n = L.load().sgd_debug_phase_cycles(self._h, _ptr(out), 2 * self.grid)
        if n < 0:
            L.check(n)
        return out[self.grid:]

    # -- peers (multi-GPU)
    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(L.SGD_PEER_HANDLE_BYTES)
        L.check(L.load().sgd_peer_export(self._h, buf))
        return buf.raw

    def peer_import(self, peer_rank: int, handle: bytes) -> None:
        assert len(handle) == L.SGD_PEER_HANDLE_BYTES
        L.check(L.load().sgd_peer_import(self._h, peer_rank, C.c_char_p(handle)))

    def peer_ready(self) -> None:
        L.check(L.load().sgd_peer_ready(self._h))
