#!/usr/bin/env python
"""Step-loop latency / bandwidth sweeps on one GPU (not the bench contract; feeds DESIGN.md).

    python tools/sweep.py grid      # c2 shape, grid x reduce-mode sweep -> us per SGD step
    python tools/sweep.py batch     # BASELINE config 3: batch 10 .. 100k on 4M x 100
    python tools/sweep.py wide      # C = 512 / 2048 at large batches -> GB/s vs the HBM peak
Prints one JSON object per configuration.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import cuda_sgd_sese_project_b200 as pkg  # noqa: E402

PEAK = 6542.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def run(R, C, B, lr=0.01, epochs=3, warm=1, shuffle=True, **kw):
    with pkg.SGDEngine(R, C, B, lr, **kw) as eng:
        eng.load_synthetic(seed=42)
        if shuffle:
            eng.shuffle()
        else:
            eng.set_permutation(np.arange(R, dtype=np.int32))
        for e in range(warm):
            eng.epoch(1 + e)
        best = None
        for e in range(epochs):
            t = eng.epoch(1 + warm + e)
            if best is None or t["steploop_ms"] < best["steploop_ms"]:
                best = t
        w = eng.weights
        phases = None
        if os.environ.get("B200SGD_PHASE_PROFILE") == "1":
            d = eng.debug_phase_cycles().astype(np.float64) / max(1, best["steps"])
            names = ["rows", "cta_reduce", "publish", "gather", "owner_publish", "wait_w", "bar2", "wait_rows"]
            own = d[:, 3] > 50   # CTAs that own columns (2-hop) / all CTAs (1-hop)
            phases = {n: [round(float(d[:, i].mean())), round(float(d[:, i].max()))] for i, n in enumerate(names) if d[:, i].any()}
            if own.any():
                phases["owners_only"] = {n: round(float(d[own, i].mean())) for i, n in enumerate(names) if d[:, i].any()}
                phases["n_owners"] = int(own.sum())
                phases["per_owner_gather"] = {int(i): round(float(d[i, 3])) for i in np.nonzero(own)[0]}
                phases["per_owner_wait_w"] = {int(i): round(float(d[i, 5])) for i in np.nonzero(own)[0]}
        out = {"R": R, "C": C, "B": B, "grid": eng.grid, "reduce": eng.reduce, "engine": eng.engine,
               "kernel": eng.kernel_shape, "slots": eng.nslots, "steps": best["steps"],
               "ms_per_epoch": round(best["steploop_ms"], 4),
               "us_per_step": round(best["steploop_ms"] * 1e3 / max(1, best["steps"]), 3),
               "gbs": round(best["bytes"] / (best["steploop_ms"] * 1e-3) * 1e-9, 1),
               "frac_of_peak": round(best["bytes"] / (best["steploop_ms"] * 1e-3) * 1e-9 / PEAK, 4),
               "msamples_per_s": round(best["samples"] / (best["steploop_ms"] * 1e-3) * 1e-6, 1),
               "finite": bool(np.all(np.isfinite(w)))}
        if phases and out["kernel"][0] == 0:   # cluster kernel: SM clock at the marks of one step
            col = eng.debug_phase_cycles().astype(np.int64)
            prev = eng.debug_step_timeline().astype(np.int64)[:, 0:1]
            cn = ["step_top", "phaseA", "barA", "phaseB+sent", "next_rows_loaded", "all_arrived", "updated", "end_bar"]
            cr = col - prev
            out["step_clock_min_med_max"] = {n: [int(cr[:, i].min()), int(np.median(cr[:, i])), int(cr[:, i].max())] for i, n in enumerate(cn)}
        elif phases:
            out["phase_cycles_mean_max"] = phases
            tl = eng.debug_step_timeline().astype(np.int64)
            t0 = tl[:, 0][tl[:, 0] > 0].min()
            rel = np.where(tl > 0, tl - t0, -1)
            own = d[:, 3] > 50
            def stats(col, mask=None):
                v = rel[:, col] if mask is None else rel[mask, col]
                v = v[v >= 0]
                return [int(v.min()), int(np.median(v)), int(v.max())] if v.size else None
            out["timeline_ns_min_med_max"] = {"rows_done": stats(0), "cta_reduced": stats(1), "published": stats(2),
                                              "gathered(owners)": stats(3, own), "owner_published": stats(4, own),
                                              "weights_arrived": stats(5), "step_end": stats(6)}
        out.update({k: v for k, v in kw.items()})
        print(json.dumps(out), flush=True)
        return out


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "grid"
    if what == "grid":
        for reduce in (4, 3, 1):
            for grid in (8, 16, 24, 32, 48, 64, 100, 148):
                if reduce in (1, 4) and grid > 64:
                    continue
                run(1_000_000, 100, 1000, reduce=reduce, grid=grid)
        for engine in (2,):
            run(1_000_000, 100, 1000, engine=engine, grid=148)
    elif what == "small":   # small grids for the latency-bound c2 shape: fewer sources per gather
        for variant in ("0", "6"):
            os.environ["B200SGD_VARIANT"] = variant
            for reduce in (4, 3):
                for grid in (2, 4, 8, 16, 32):
                    r = run(1_000_000, 100, 1000, reduce=reduce, grid=grid)
        os.environ.pop("B200SGD_VARIANT", None)
    elif what == "batch":
        for B in (10, 100, 1000, 4096, 10000, 100000):
            run(4_000_000, 100, B, lr=0.001, epochs=2 if B >= 100 else 1, warm=1 if B >= 100 else 0)
    elif what == "wide":
        run(2_097_152, 512, 8192, epochs=2)
        run(2_097_152, 512, 65536, epochs=2)
        run(1_048_576, 2048, 4096, epochs=2)
        run(1_048_576, 2048, 65536, epochs=2)
        run(4_000_000, 100, 100000, epochs=2)
        run(4_000_000, 100, 100000, epochs=2, shuffle=False)   # sequential rows: gather cost isolated
    elif what == "one":   # python tools/sweep.py one R C B [grid] [reduce]  (ncu target)
        a = [int(v) for v in sys.argv[2:]]
        kw = {}
        if len(a) > 3:
            kw["grid"] = a[3]
        if len(a) > 4:
            kw["reduce"] = a[4]
        run(a[0], a[1], a[2], epochs=1, warm=1, **kw)
    elif what == "shuffle":
        for n in (4_000_000, 16_777_216, 33_554_432):
            ind = np.arange(n, dtype=np.int32)
            t0 = time.perf_counter()
            pkg.host_shuffle(ind, epochs=1)
            print(json.dumps({"host_shuffle_rows": n, "ms": round((time.perf_counter() - t0) * 1e3, 2)}), flush=True)


if __name__ == "__main__":
    main()
