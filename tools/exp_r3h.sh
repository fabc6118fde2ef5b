#!/usr/bin/env bash
# round 2, session 2, call 8 (8 GPUs): the strong-scaling bench line of config 5 at N = 8 with the backward shuffle
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --no-cpu > gpurun_out/r3h_bench_n8.json 2> gpurun_out/r3h_bench_n8.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r3h_bench_n8.json").read().strip().splitlines()[-1])
    print(round(d["value"] / 1e6, 1), round(d["ms_per_step"], 2), round(d["steploop"]["us_per_sgd_step"], 2), round(d["roofline"]["frac"], 3), round(d["roofline"]["kernel_alone"]["us_per_sgd_step"], 2), (d.get("e2e") or {}).get("value"), json.dumps(d["quality"])[:1200])
except Exception as e:
    print("parse failed", e)
PY
tail -3 gpurun_out/r3h_bench_n8.err | cut -c1-300
