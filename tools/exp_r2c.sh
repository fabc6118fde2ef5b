#!/usr/bin/env bash
mkdir -p gpurun_out
for cs in 8 16; do
python tools/sweep2.py one 1048576 2048 4096 CLT=$cs CLUSTER=0 PROFILE=1
python tools/sweep2.py one 2097152 512 8192 CLT=$cs CLUSTER=0 PROFILE=1
done > gpurun_out/r2c.jsonl 2>gpurun_out/r2c.err
python tools/sweep2.py one 2097152 512 8192 CLT=8 CLUSTER=0 PROFILE=1 RING_TILES=2 >> gpurun_out/r2c.jsonl 2>>gpurun_out/r2c.err
python tools/sweep2.py one 2097152 512 65536 CLT=8 CLUSTER=0 PROFILE=1 >> gpurun_out/r2c.jsonl 2>>gpurun_out/r2c.err
python tools/sweep2.py one 1048576 2048 65536 CLT=8 CLUSTER=0 PROFILE=1 >> gpurun_out/r2c.jsonl 2>>gpurun_out/r2c.err
python - <<'PY' >> gpurun_out/r2c.jsonl 2>>gpurun_out/r2c.err
import time, json, sys
sys.path.insert(0, '.')
import numpy as np
import cuda_sgd_sese_project_b200 as pkg
for (R, C) in ((4_000_000, 100), (2_097_152, 512), (1_048_576, 2048)):
    with pkg.SGDEngine(R, C, 1000, 0.1) as eng:
        eng.load_synthetic(seed=1)
        eng.calculate_mean_abs_error()
        t0 = time.perf_counter()
        for _ in range(10):
            m = eng.calculate_mean_abs_error()
        dt = (time.perf_counter() - t0) / 10
        print(json.dumps({"mae_R": R, "C": C, "ld": eng.ld, "ms": dt * 1e3, "gbs": R * eng.ld * 4 / dt * 1e-9, "mae": m}))
PY
cat gpurun_out/r2c.jsonl
