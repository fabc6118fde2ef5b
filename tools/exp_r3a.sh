#!/usr/bin/env bash
# round 2, session 2, call 1: the backward shuffle on the device -- parity (all three formulations),
# wall times, bench A/B against the forward kernels on the same box, ncu launch list at 64M rows
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shuffle.py -q -x > gpurun_out/r3a_pytest_shuffle.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r3a_pytest_shuffle.log
tail -5 gpurun_out/r3a_pytest_shuffle.log
timeout 200 python tools/shuffle_time.py > gpurun_out/r3a_shuffle_time.jsonl 2> gpurun_out/r3a_shuffle_time.err
echo "shuffle_time rc=$?"; cat gpurun_out/r3a_shuffle_time.jsonl | cut -c1-200
for v in 3 1 2; do
  B200SGD_SHUFFLE=$v timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-sweep --no-e2e > gpurun_out/r3a_bench_v$v.json 2> gpurun_out/r3a_bench_v$v.err
  echo "bench v$v rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r3a_bench_v$v.json").read().strip().splitlines()[-1])
    print("v$v", d["value"], d["ms_per_step"], d["steploop"]["us_per_sgd_step"], d["roofline"]["frac"], d["quality"]["permutation_golden_check"][0]["match"], d["quality"]["permutation_golden_check"][1]["match"])
except Exception as e:
    print("v$v parse failed", e)
PY
done
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/r3a_launches_shuffle64m.csv python tools/prof_targets.py shuffle 67108864 > gpurun_out/r3a_ncu_shuffle.log 2>&1
echo "ncu rc=$?"
