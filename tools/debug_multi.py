#!/usr/bin/env python
"""Debug helper: 2-rank run of a tiny case, prints loss curve / gradient / weight change per rank."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch, torch.distributed as dist
import cuda_sgd_sese_project_b200 as pkg
from cuda_sgd_sese_project_b200.dist import connect_peers
from tests.helpers import make_regression

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
R, C, B = 4000, 100, 1000
X, y, _ = make_regression(R, C, seed=11)
eng = pkg.SGDEngine(R, C, B, 0.05, device=rank, rank=rank, nranks=world)
b, e = pkg.shard_range(R, rank, world)
eng.load_rows(b, X[b:e], y[b:e])
connect_peers(eng)
w0 = eng.weights.copy()
eng.iteration(1)
w1 = eng.weights
la, ls = eng.loss_curve()
g = eng.last_gradient()
local, off = eng.local_schedule()
print(rank, "kernel", eng.kernel_name, "grid", eng.grid, "loss", la, ls, "|g|", float(np.abs(g).sum()), "|dw|", float(np.abs(w1 - w0).sum()),
      "off", off[:5], "local[:4]", local[:4], flush=True)
dist.barrier()
eng.close()
dist.destroy_process_group()
