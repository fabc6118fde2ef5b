#!/usr/bin/env python
"""Wall time of the bit-exact device shuffle (sgd_shuffle, synchronous) by data set size."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_sgd_sese_project_b200 as pkg
for R in [int(a) for a in sys.argv[1:]] or [4_000_000, 16_777_216, 67_108_864]:
    with pkg.SGDEngine(R, 1, 1000, 0.1) as eng:
        eng.load_synthetic(seed=1)
        eng.shuffle()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); eng.shuffle(); ts.append((time.perf_counter() - t0) * 1e3)
        print(json.dumps({"rows": R, "device_shuffle_ms_best_of_3": round(min(ts), 3), "all": [round(t, 3) for t in ts]}), flush=True)
