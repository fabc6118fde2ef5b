#!/usr/bin/env bash
# round 2, session 2, call 9: permutation kernels confined to the SMs the step grid leaves idle (one GPU) -- A/B
mkdir -p gpurun_out
run_bench() { # tag env...
  local tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-sweep --no-e2e > gpurun_out/r3i_bench_$tag.json 2> gpurun_out/r3i_bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r3i_bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["value"] / 1e6, 1), round(d["ms_per_step"], 2), round(d["steploop"]["us_per_sgd_step"], 2), round(d["roofline"]["frac"], 3), round(d["roofline"]["kernel_alone"]["us_per_sgd_step"], 2), d["quality"]["permutation_golden_check"][1]["match"])
except Exception as e:
    print("$tag parse failed", e)
PY
}
run_bench confine B200SGD_SHUF_CONFINE=1
run_bench colocated B200SGD_SHUF_CONFINE=0
B200SGD_SHUF_CONFINE=1 timeout 100 python tools/shuffle_time.py 67108864 v=5 | cut -c1-150
