#!/usr/bin/env bash
# round 2, session 2, call 2: scatter strategies of the backward shuffle (B200SGD_SHUFFLE=2..5), L2 fetch granularity
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shuffle.py -q -x > gpurun_out/r3b_pytest_shuffle.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r3b_pytest_shuffle.log
tail -3 gpurun_out/r3b_pytest_shuffle.log
timeout 200 python tools/shuffle_time.py 4000000 67108864 > gpurun_out/r3b_shuffle_time.jsonl 2> gpurun_out/r3b_shuffle_time.err
B200SGD_L2_FETCH=32 timeout 200 python tools/shuffle_time.py 67108864 v=4,2,1 > gpurun_out/r3b_shuffle_time_l2f32.jsonl 2>> gpurun_out/r3b_shuffle_time.err
echo "shuffle_time rc=$?"; cat gpurun_out/r3b_shuffle_time.jsonl gpurun_out/r3b_shuffle_time_l2f32.jsonl | cut -c1-140
run_bench() { # tag env...
  local tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-sweep --no-e2e > gpurun_out/r3b_bench_$tag.json 2> gpurun_out/r3b_bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r3b_bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["value"] / 1e6, 1), round(d["ms_per_step"], 2), round(d["steploop"]["us_per_sgd_step"], 2), round(d["roofline"]["frac"], 3), round(d["roofline"]["kernel_alone"]["us_per_sgd_step"], 2), d["quality"]["permutation_golden_check"][1]["match"])
except Exception as e:
    print("$tag parse failed", e)
PY
}
run_bench v4 B200SGD_SHUFFLE=4
run_bench v5 B200SGD_SHUFFLE=5
run_bench v2 B200SGD_SHUFFLE=2
run_bench v4_l2f32 B200SGD_SHUFFLE=4 B200SGD_L2_FETCH=32
run_bench v2_l2f32 B200SGD_SHUFFLE=2 B200SGD_L2_FETCH=32
run_bench v1_l2f32 B200SGD_SHUFFLE=1 B200SGD_L2_FETCH=32
for v in 4 2; do
B200SGD_SHUFFLE=$v timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/r3b_launches_shuffle64m_v$v.csv python tools/prof_targets.py shuffle 67108864 > gpurun_out/r3b_ncu_shuffle_v$v.log 2>&1
echo "ncu v$v rc=$?"
done
B200SGD_SHUFFLE=4 B200SGD_L2_FETCH=32 timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/r3b_launches_shuffle64m_v4_l2f32.csv python tools/prof_targets.py shuffle 67108864 > gpurun_out/r3b_ncu_shuffle_v4l.log 2>&1
