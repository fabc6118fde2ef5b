#!/usr/bin/env python
"""Step-loop timing on several GPUs without the concurrent permutation work (torchrun):
    python -m torch.distributed.run --nproc-per-node N tools/sweep_multi.py R C B [PROFILE]
shuffle, then epochs one by one (CUDA events around the epoch kernel, max over ranks); rank 0 prints
one JSON line, with the per-phase marks of the cluster tail when B200SGD_PHASE_PROFILE=1."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import cuda_sgd_sese_project_b200 as pkg
from cuda_sgd_sese_project_b200.dist import connect_peers

rank, N, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
R, C, B = (int(v) for v in sys.argv[1:4])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
eng = pkg.SGDEngine(R, C, B, 0.01, device=lr_, rank=rank, nranks=N)
eng.load_synthetic(seed=42)
connect_peers(eng)
best = None
for e in range(1, 4):
    eng.shuffle()
    dist.barrier()
    t = eng.epoch(e)
    ms = torch.tensor([t["steploop_ms"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if e > 1 and (best is None or float(ms[0]) < best):
        best = float(ms[0])
out = {"N": N, "R": R, "C": C, "B": B, "rows_per_gpu_step": B / N, "grid": eng.grid, "kernel": eng.kernel_name,
       "us_per_step": round(best * 1e3 / t["steps"], 3),
       "frac": round(t["bytes"] / (best * 1e-3) * 1e-9 / 6542.4, 4)}
if os.environ.get("B200SGD_PHASE_PROFILE") == "1":
    d = eng.debug_phase_cycles().astype(np.float64) / max(1, t["steps"])
    out["marks_mean"] = [round(float(d[:, i].mean())) for i in range(8)]
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
eng.close()
dist.destroy_process_group()
