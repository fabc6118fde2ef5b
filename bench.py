#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native SGD engine (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1|c4|c5]

A bench "step" is ONE SGD EPOCH of the workload (BASELINE.json's metric is quoted per epoch): the
exact per-epoch permutation + floor(R/B) fused mini-batch steps.  At N = 1 the workload is BASELINE
configs[1] (c2: 4,000,000 x 100 fp32, batch 1000, ~1.6 GB).  At N > 1 it is weak-scaled: every GPU
keeps 4,000,000 rows and the global batch is 1000*N, so each GPU processes ~1000 rows per step and
the d-float gradient is exchanged over NVLink inside the step kernel.

One JSON line on stdout (rank 0):
  value      samples/s over K epochs with the data resident in HBM, INCLUDING whatever part of the
             bit-exact permutation is not hidden behind the previous epoch (what the reference's
             gpu_time covers, main.cu:444-514); wall clock between barriers + synchronize, max over ranks
  steploop   samples/s and GB/s of the step loop alone (CUDA events around the epoch kernels)
  roofline   the dominant kernel (sgd_epoch_kernel): algorithmic bytes / CUDA-event time vs the
             measured HBM copy peak (MEASURED_PEAKS.json)
  e2e        same metric through the C-ABI from HOST buffers: sgd_create (H2D of the data set from
             pinned memory) + K epochs (H2D of each permutation) + weights/MAE read-back; one warm
             pass, then the median of three timed passes (all three are listed)
  roofline   (c2: the dominant kernel is sgd_cluster_epoch_kernel, see config.kernel)
  cpu_baseline  OpenMP port of the reference's plain-CUDA path (oracle/, kind "port") on the host cores

--impl reference times the UNMODIFIED reference binary (oracle/_ref/main, its cuBLAS code path,
cudaRun = 0) on the same B200 on a bounded sample of the workload (same C and B, fewer rows; its
per-step cost does not depend on R).  The reference is a CUDA program and has no CPU path; if the
binary is missing or fails, the OpenMP port is timed instead and the line says so.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (rows per GPU, cols, batch per GPU, lr)
    "c1": (10_000, 100, 1000, 0.1),
    "c2": (4_000_000, 100, 1000, 0.1),
    # the report's own ~1.6 GB case (sese_report_tvas.tex:690-706, source of the README's "~5x"): 200k x 2k
    "c2p": (200_000, 2000, 1000, 0.02),
    "c4": (16_777_216, 2048, 4096, 0.1),     # 137 GB: one GPU holds it
    "c4b": (16_777_216, 2048, 65536, 0.1),
    "c5": (8_388_608, 512, 8192, 0.1),       # x8 GPUs = 64M x 512, batch 65536
}
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._t is not None:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def profile_traffic(workload: str):
    """dram bytes per launch of the epoch kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


# ------------------------------------------------------------------------------- CPU baseline
def cpu_port_baseline(X, y, B, lr, epochs):
    """OpenMP port of the reference's plain-CUDA path (oracle/), all host cores, whole run incl.
    the exact shuffle -> samples/s."""
    from oracle import loader as orc
    threads = orc.max_threads()
    R = X.shape[0]
    _, nb = orc.num_batches(R, B)
    w = orc.init_weights(X.shape[1])
    ind = np.arange(R, dtype=np.int32)
    g = orc.Rand()
    g.shuffle(ind)
    orc.epoch(X[: min(R, 50 * B)], y[: min(R, 50 * B)], B, np.arange(min(R, 50 * B), dtype=np.int32), w.copy(), lr, 1,
              threads=threads)  # warm the thread pool
    t0 = time.perf_counter()
    for e in range(1, epochs + 1):
        g.shuffle(ind)
        orc.epoch(X, y, B, ind, w, lr, e, threads=threads)
    dt = time.perf_counter() - t0
    return (epochs * nb * B) / dt, threads, dt


# ------------------------------------------------------------------------------- reference arm
def reference_arm(args, cols, batch, lr):
    import torch
    ref = os.path.join(ROOT, "oracle", "_ref", "main")
    K, W = args.steps, args.warmup
    sample_rows = args.ref_rows
    rng = np.random.default_rng(42)
    X = rng.standard_normal((sample_rows, cols)).astype(np.float32)
    w_true = (100.0 * rng.random(cols)).astype(np.float32)
    y = (X @ w_true).astype(np.float32)
    nb = int(np.floor(np.float32(sample_rows) / np.float32(batch)))
    line = {"impl": "reference", "metric": "samples_per_sec_per_sgd_epoch", "unit": "samples/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: sample of {sample_rows} x {cols} fp32 rows, batch {batch}, "
                                   f"lr {lr} (same C and B as the workload; the reference's per-step cost "
                                   "does not depend on R)", "l2": "not flushed: reference run as shipped"}}
    ok = False
    if os.path.exists(ref) and torch.cuda.is_available():
        with tempfile.TemporaryDirectory() as tmp:
            csv = os.path.join(tmp, "bench_sample.csv")
            with open(csv, "w") as f:
                for row, lab in zip(X, y):
                    f.write(",".join("%.9g" % v for v in row) + ",%.9g\n" % lab)

            def run(epochs, cuda_run="0"):
                r = subprocess.run([ref, repr(lr), str(epochs), csv, str(sample_rows), str(cols), str(batch), cuda_run],
                                   capture_output=True, text=True, cwd=tmp, timeout=1500)
                if r.returncode != 0:
                    raise RuntimeError(r.stderr[-500:])
                return float(re.search(r"GPU time = (\S+) ms", r.stdout).group(1))
            def timed_ms(cuda_run):
                """gpu_time of the K timed epochs: gpu_time(W + K epochs) - gpu_time(W epochs) of two runs.
                Every run pays CUDA / cuBLAS start-up inside its gpu_time with a jitter of hundreds of ms, so
                for a small K the difference can come out near zero or negative; then the average epoch of
                the long run (start-up included: an upper bound of the epoch time) stands in."""
                run(1, cuda_run)  # discarded: the first start of the binary on a fresh box pays one-time library loads
                t_w = run(W, cuda_run) if W > 0 else 0.0
                t_all = run(W + K, cuda_run)
                avg = t_all * K / (W + K)
                diff = t_all - t_w
                ok_diff = 0.25 * avg <= diff <= t_all
                return (diff if ok_diff else avg), {"gpu_time_ms_all": t_all, "gpu_time_ms_warmup": t_w,
                                                    "method": "difference of two runs" if ok_diff else
                                                              "average epoch of the long run (difference implausible)"}
            try:
                ms, how = timed_ms("0")
                value = K * nb * batch / (ms * 1e-3)
                line["reference_timing"] = how
                # for the record: the reference's other code path (cudaRun = 1, plain CUDA kernel +
                # Thrust), which on a B200 is the faster of the two (no cublasCreate per step)
                try:
                    ms1, how1 = timed_ms("1")
                    line["reference_plain_cuda_path"] = {"value": K * nb * batch / (ms1 * 1e-3), "unit": "samples/s",
                                                         "ms_per_step": ms1 / K, "timing": how1}
                except Exception as ex:
                    line["reference_plain_cuda_path"] = {"error": str(ex)[:120]}
                line.update({"value": value, "ms_per_step": ms / K,
                             "cpu_baseline": {"value": value, "unit": "samples/s", "cores": 1, "kind": "reference",
                                              "sample": f"unmodified reference binary (cuBLAS code path, cudaRun=0) on the "
                                                        f"B200 -- the reference has no CPU path; {sample_rows} rows, "
                                                        f"gpu_time({W + K} epochs) - gpu_time({W} epochs)"}})
                ok = True
            except Exception as ex:  # fall through to the port
                line["reference_binary_error"] = str(ex)[:200]
    if not ok:
        v, threads, dt = cpu_port_baseline(X, y, batch, lr, max(1, K))
        line.update({"value": v, "ms_per_step": dt * 1e3 / max(1, K),
                     "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                                      "sample": f"OpenMP port of the reference's plain-CUDA path, {sample_rows} rows, {max(1, K)} epochs"}})
    line["e2e"] = {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    line["gpu_launches"] = 0
    return line


# ------------------------------------------------------------------------------- our arm
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10, help="timed epochs")
    ap.add_argument("--warmup", type=int, default=3, help="untimed warm-up epochs (>= 3)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU")
    ap.add_argument("--cols", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0, help="override batch per GPU")
    ap.add_argument("--engine", default="persistent", choices=["persistent", "graph"])
    ap.add_argument("--grid", type=int, default=0)
    ap.add_argument("--reduce", default="auto", choices=["auto", "allgather", "scatter", "dataflow", "dataflow1"])
    ap.add_argument("--ref-rows", type=int, default=100_000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-epochs", type=int, default=100, help="epochs of the bounded cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    rows_pg, cols, batch_pg, lr = WORKLOADS[args.workload]
    rows_pg = args.rows or rows_pg
    cols = args.cols or cols
    batch_pg = args.batch or batch_pg

    if args.impl == "reference":
        if rank != 0:
            return 0
        print(json.dumps(reference_arm(args, cols, batch_pg, lr)), flush=True)
        return 0

    import torch
    import torch.distributed as dist

    import cuda_sgd_sese_project_b200 as pkg
    from cuda_sgd_sese_project_b200 import _lib

    if not torch.cuda.is_available() or pkg.device_count() < 1:
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    N = world
    if N > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert N == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    R, B = rows_pg * N, batch_pg * N
    K, W = args.steps, max(args.warmup, 3)
    eng = pkg.SGDEngine(R, cols, B, lr, device=local_rank, rank=rank, nranks=N,
                        engine=_lib.SGD_ENGINE_GRAPH if args.engine == "graph" else _lib.SGD_ENGINE_PERSISTENT,
                        reduce={"auto": 0, "allgather": 1, "scatter": 2, "dataflow": 3, "dataflow1": 4}[args.reduce], grid=args.grid)
    eng.load_synthetic(seed=42)
    from cuda_sgd_sese_project_b200.dist import connect_peers
    if N > 1:
        connect_peers(eng)

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: K epochs, data resident, permutation pipeline included
    eng.run(1, W)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    t = eng.run(W + 1, K)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    times = torch.tensor([wall, t["steploop_ms"] * 1e-3], dtype=torch.float64, device="cuda")
    if N > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    wall, loop_s = float(times[0]), float(times[1])
    samples = t["samples"]                      # K * steps * B (whole job)
    bytes_rank = t["bytes"]                     # algorithmic bytes this rank moved in K epochs
    mae_sum = torch.tensor([eng.abs_error_sum()], dtype=torch.float64, device="cuda")
    if N > 1:
        dist.all_reduce(mae_sum)
    mae = float(mae_sum[0]) / R
    w_final = eng.weights
    w_err = float(np.linalg.norm(w_final - eng.true_weights()) / np.linalg.norm(eng.true_weights()))

    # ---- e2e: through the C-ABI from pinned HOST buffers, at N GPUs: every rank holds its row
    # shard in pinned host memory; the timed region has sgd_create, the H2D of the shard, K epochs
    # (permutation state upload inside), the weights and MAE read-back
    e2e = None
    cpu = None
    if not args.no_e2e and rows_pg * (cols + 1) * 4 <= (24 << 30):
        rb, nloc = eng.local_row_begin, eng.local_rows
        Xh = torch.empty((nloc, cols), dtype=torch.float32, pin_memory=True)
        yh = torch.empty((nloc,), dtype=torch.float32, pin_memory=True)
        Xn, yn = eng.get_rows(rb, nloc)
        Xh.numpy()[:] = Xn
        yh.numpy()[:] = yn
        del Xn, yn
        eng.close()
        kw = dict(device=local_rank, rank=rank, nranks=N,
                  engine=_lib.SGD_ENGINE_GRAPH if args.engine == "graph" else _lib.SGD_ENGINE_PERSISTENT,
                  reduce={"auto": 0, "allgather": 1, "scatter": 2, "dataflow": 3, "dataflow1": 4}[args.reduce], grid=args.grid)
        passes = []
        for it in range(4):  # first pass warms allocator / page tables; three timed passes, the median is reported
            barrier()
            t0 = time.perf_counter()
            e2 = pkg.SGDEngine(R, cols, B, lr, **kw)
            e2.load_rows_ptr(rb, nloc, Xh.data_ptr(), yh.data_ptr())   # H2D from pinned memory
            if N > 1:
                connect_peers(e2)
            te = e2.run(1, K)
            w_e2e = e2.weights                                          # D2H
            s_e2e = torch.tensor([e2.abs_error_sum()], dtype=torch.float64, device="cuda")  # D2H
            if N > 1:
                dist.all_reduce(s_e2e)
            barrier()
            if it > 0:
                passes.append(time.perf_counter() - t0)
            e2.close()
        dts = torch.tensor(passes, dtype=torch.float64, device="cuda")
        if N > 1:
            dist.all_reduce(dts, op=dist.ReduceOp.MAX)   # every pass: the slowest rank
        passes = sorted(float(v) for v in dts)
        dt = passes[len(passes) // 2]
        e2e = {"value": te["samples"] / dt, "unit": "samples/s",
               "h2d_bytes_per_step": int((4 * R * (cols + 1) + K * 31 * 4 * 8192) / K),
               "d2h_bytes_per_step": int(N * (4 * cols + 8 * 148 * 8) / K),
               "seconds": dt, "seconds_all_passes": passes, "note": "median of three passes of: sgd_create + upload of every rank's row shard from pinned host memory + "
                                      f"{K} epochs + weights and MAE read-back, amortised over the {K} epochs; max over ranks",
               "mae": float(s_e2e[0]) / R}
        # ---- cpu_baseline (rank 0, N = 1): OpenMP port on the host cores, bounded sample
        if N == 1 and not args.no_cpu:
            sample = min(R, 4_000_000)
            cpu_epochs = args.cpu_epochs
            v, threads, dtc = cpu_port_baseline(Xh.numpy()[:sample], yh.numpy()[:sample], B, lr, cpu_epochs)
            cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                   "sample": f"first {sample} rows of the workload (same C = {cols}, B = {B}), {cpu_epochs} epochs incl. the "
                             f"exact shuffle, OpenMP port of the reference's plain-CUDA path, {dtc:.1f} s of CPU work"}
    else:
        eng.close()

    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = bytes_rank / loop_s * 1e-9                      # GB/s per GPU, step loop
        steps_epoch = t["steps"] // K
        line = {
            "metric": "samples_per_sec_per_sgd_epoch", "value": samples / wall, "unit": "samples/s",
            "n_gpus": N, "steps": K, "warmup": W, "ms_per_step": wall * 1e3 / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {R} x {cols} fp32 rows ({4 * R * (cols + 1) / 1e9:.2f} GB, "
                                   f"{rows_pg} rows per GPU), batch {B}, lr {lr}, {steps_epoch} steps per epoch; "
                                   "one bench step = one SGD epoch incl. its bit-exact permutation",
                       "l2": "inputs larger than L2 (per-GPU data %.1f GB vs 126 MB)" % (4 * rows_pg * (cols + 1) / 1e9),
                       "engine": args.engine, "grid": eng.grid,
                       "reduce": ("cluster all-gather in shared memory" + (" + L2 hop between clusters" if eng.grid > eng.cluster_size else "")) if eng.cluster_size else
                                 {1: "allgather", 2: "scatter", 3: "dataflow-2hop", 4: "dataflow-1hop"}[eng.reduce],
                       "kernel": eng.kernel_name,
                       "ring_slots": eng.nslots, "smem_bytes": eng.smem_bytes,
                       "parallelism": f"dp{N} row-sharded, fused NVLink gradient exchange" if N > 1 else "single GPU"},
            "steploop": {"samples_per_s": samples / loop_s, "ms_per_epoch": loop_s * 1e3 / K,
                         "us_per_sgd_step": loop_s * 1e6 / t["steps"], "hbm_gbs_per_gpu": achieved},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": profile_traffic(args.workload),
                         "peak_source": peak_src, "kernel": eng.kernel_name.split("<")[0].split(" (")[0] + " (1 launch per epoch)",
                         "algorithmic_bytes_per_launch": bytes_rank // K},
            "clocks": clocks,
            "gpu_launches": t["launches"],
            "quality": {"mae": mae, "weights_rel_err_vs_generating_weights": w_err},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if N > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
