#!/usr/bin/env python
"""Wall time of the bit-exact device shuffle (sgd_shuffle, synchronous) by data set size and formulation
(B200SGD_SHUFFLE, see b200sgd_api.cu: 1 = forward / round 1, 2..5 = backward with different scatters).
usage: shuffle_time.py [rows ...] [v=4,5,3,2,1]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_sgd_sese_project_b200 as pkg
vers = [4, 5, 3, 2, 1]
sizes = []
for a in sys.argv[1:]:
    if a.startswith("v="):
        vers = [int(v) for v in a[2:].split(",")]
    else:
        sizes.append(int(a))
for R in sizes or [4_000_000, 16_777_216, 67_108_864]:
    for ver in vers:
        os.environ["B200SGD_SHUFFLE"] = str(ver)
        with pkg.SGDEngine(R, 1, 1000, 0.1) as eng:
            eng.load_synthetic(seed=1)
            eng.shuffle()
            ts = []
            for _ in range(4):
                t0 = time.perf_counter(); eng.shuffle(); ts.append((time.perf_counter() - t0) * 1e3)
            p = eng.permutation()
            print(json.dumps({"rows": R, "shuffle_version": ver, "device_shuffle_ms_best_of_4": round(min(ts), 3),
                              "all": [round(t, 3) for t in ts], "head": p[:4].tolist()}), flush=True)
