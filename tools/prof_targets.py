#!/usr/bin/env python
"""Small, single-purpose launches for ncu (one shape per call; see tools/exp_ncu.sh):
    python tools/prof_targets.py epoch R C B      # shuffle, 1 warm-up epoch, 1 epoch (the capture target is the last launch)
    python tools/prof_targets.py mae R C          # two MAE passes
    python tools/prof_targets.py shuffle R        # two device shuffles"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_sgd_sese_project_b200 as pkg

what = sys.argv[1]
a = [int(v) for v in sys.argv[2:]]
if what == "epoch":
    with pkg.SGDEngine(a[0], a[1], a[2], 0.01) as eng:
        eng.load_synthetic(seed=42)
        eng.shuffle()
        eng.epoch(1)
        t = eng.epoch(2)
        print(eng.kernel_name, eng.grid, t["steploop_ms"], t["bytes"])
elif what == "mae":
    with pkg.SGDEngine(a[0], a[1], 1000, 0.01) as eng:
        eng.load_synthetic(seed=42)
        print(eng.calculate_mean_abs_error(), eng.calculate_mean_abs_error())
elif what == "shuffle":
    with pkg.SGDEngine(a[0], 1, 1000, 0.01) as eng:
        eng.load_synthetic(seed=42)
        eng.shuffle()
        eng.shuffle()
        print("ok")
