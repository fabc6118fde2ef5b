#!/usr/bin/env bash
# round 2, session 2, call 6 (2 GPUs): multi-rank tests (shared permutation maps from the backward kernels,
# look-ahead, set_permutation after a shared look-ahead) and the strong-scaling bench line at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r3f_pytest_multi.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r3f_pytest_multi.log
tail -4 gpurun_out/r3f_pytest_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu > gpurun_out/r3f_bench_n2.json 2> gpurun_out/r3f_bench_n2.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r3f_bench_n2.json").read().strip().splitlines()[-1])
    print(round(d["value"] / 1e6, 1), round(d["ms_per_step"], 2), round(d["steploop"]["us_per_sgd_step"], 2), round(d["roofline"]["frac"], 3), round(d["roofline"]["kernel_alone"]["us_per_sgd_step"], 2), json.dumps(d["quality"])[:900])
except Exception as e:
    print("parse failed", e)
PY
