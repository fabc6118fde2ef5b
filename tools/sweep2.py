#!/usr/bin/env python
"""Round-2 step-tail experiments on one GPU: shape x (reduction, cluster size, bytes in flight).

    python tools/sweep2.py tail   [--quick]      # the matrix behind DESIGN.md "loaded latency"
    python tools/sweep2.py one R C B [key=value ...]   (env keys: CLT, INFLIGHT_KB, CLUSTERS, RING_TILES, PROFILE)

Every configuration is one fresh engine (environment switches are read at sgd_create).  One JSON
object per line; the per-phase marks are SM-clock sums per step (B200SGD_PHASE_PROFILE=1).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import cuda_sgd_sese_project_b200 as pkg  # noqa: E402

PEAK = 6542.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

ENV_KEYS = {"CLT": "B200SGD_CLT", "INFLIGHT_KB": "B200SGD_INFLIGHT_KB", "CLUSTERS": "B200SGD_CLUSTERS",
            "RING_TILES": "B200SGD_RING_TILES", "PROFILE": "B200SGD_PHASE_PROFILE", "CLUSTER": "B200SGD_CLUSTER",
            "PRODUCER_WARPS": "B200SGD_PRODUCER_WARPS", "VARIANT": "B200SGD_VARIANT", "DBGSKIP": "B200SGD_DBGSKIP"}


def run(R, C, B, lr=0.01, epochs=2, warm=1, reduce=0, grid=0, seq=0, **env):
    saved = {}
    for k, v in env.items():
        name = ENV_KEYS[k]
        saved[name] = os.environ.get(name)
        os.environ[name] = str(v)
    try:
        with pkg.SGDEngine(R, C, B, lr, reduce=reduce, grid=grid) as eng:
            eng.load_synthetic(seed=42)
            if seq:   # rows in storage order: what the random gather itself costs
                eng.set_permutation(np.arange(R, dtype=np.int32))
            else:
                eng.shuffle()
            for e in range(warm):
                eng.epoch(1 + e)
            best = None
            for e in range(epochs):
                t = eng.epoch(1 + warm + e)
                if best is None or t["steploop_ms"] < best["steploop_ms"]:
                    best = t
            w = eng.weights
            out = {"R": R, "C": C, "B": B, "grid": eng.grid, "reduce": eng.reduce, "kernel": eng.kernel_shape,
                   "ring_rows": eng.nslots, "smem": eng.smem_bytes,
                   "us_per_step": round(best["steploop_ms"] * 1e3 / max(1, best["steps"]), 3),
                   "gbs": round(best["bytes"] / (best["steploop_ms"] * 1e-3) * 1e-9, 1),
                   "frac": round(best["bytes"] / (best["steploop_ms"] * 1e-3) * 1e-9 / PEAK, 4),
                   "finite": bool(np.all(np.isfinite(w))),
                   "w_err": float(np.linalg.norm(w - eng.true_weights()) / np.linalg.norm(eng.true_weights()))}
            if os.environ.get("B200SGD_PHASE_PROFILE") == "1":
                d = eng.debug_phase_cycles().astype(np.float64) / max(1, best["steps"])
                out["marks_mean"] = [round(float(d[:, i].mean())) for i in range(8)]
                out["marks_max"] = [round(float(d[:, i].max())) for i in range(8)]
            out.update(env)
            if seq:
                out["rows_in_storage_order"] = True
            print(json.dumps(out), flush=True)
            return out
    except Exception as ex:  # keep the sweep going
        print(json.dumps({"R": R, "C": C, "B": B, "error": str(ex)[:200], **env}), flush=True)
        return None
    finally:
        for name, v in saved.items():
            if v is None:
                os.environ.pop(name, None)
            else:
                os.environ[name] = v


SHAPES = [
    (1_048_576, 2048, 4096),    # config 4 row width, B = 4096
    (2_097_152, 512, 8192),     # config 5 per-GPU shape at 8 GPUs
    (2_097_152, 512, 65536),    # config 5 on one GPU
    (1_048_576, 2048, 65536),
    (4_000_000, 100, 100000),
    (4_000_000, 100, 16384),
    (200_000, 2000, 1000),      # the report's own 1.6 GB case
]


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "tail"
    if what == "tail":
        quick = "--quick" in sys.argv
        for (R, C, B) in SHAPES[:3] if quick else SHAPES:
            run(R, C, B, reduce=3, CLUSTER=0)                 # round-1 tail (two hops through L2)
            for cs in (8, 16):
                run(R, C, B, CLT=cs, CLUSTER=0)
                if C > 128:
                    run(R, C, B, CLT=cs, CLUSTER=0, VARIANT=8)    # COLSPLIT mapping
        # where the time goes: marks of the cluster tail
        for (R, C, B) in SHAPES[:2]:
            for cs in (8, 16):
                run(R, C, B, CLT=cs, CLUSTER=0, PROFILE=1)
                run(R, C, B, CLT=cs, CLUSTER=0, PROFILE=1, VARIANT=8)
    elif what == "inflight":   # the round-2 negative result: bytes in flight vs step time
        for (R, C, B) in SHAPES[:5]:
            for kb in (0, 32, 64, 96, 128):
                run(R, C, B, CLT=8, INFLIGHT_KB=kb, CLUSTER=0)
    elif what == "one":
        a = [int(v) for v in sys.argv[2:5]]
        kv = dict(x.split("=") for x in sys.argv[5:])
        kw = {k: int(v) for k, v in kv.items() if k in ("reduce", "grid", "epochs", "warm", "seq")}
        env = {k: v for k, v in kv.items() if k in ENV_KEYS}
        run(a[0], a[1], a[2], **kw, **env)


if __name__ == "__main__":
    main()
