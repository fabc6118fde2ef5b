#!/usr/bin/env bash
# round 2, session 2, call 11 (4 GPUs): strong-scaling bench line of config 5 at N = 4 (no e2e / cpu legs)
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --no-cpu --no-e2e > gpurun_out/r3k_bench_n4.json 2> gpurun_out/r3k_bench_n4.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r3k_bench_n4.json").read().strip().splitlines()[-1])
    print(round(d["value"] / 1e6, 1), round(d["ms_per_step"], 2), round(d["steploop"]["us_per_sgd_step"], 2), round(d["roofline"]["frac"], 3), round(d["roofline"]["kernel_alone"]["us_per_sgd_step"], 2), d["quality"]["replicas_bit_identical"], d["quality"]["permutation_golden_check"][1]["match"])
except Exception as e:
    print("parse failed", e)
PY
