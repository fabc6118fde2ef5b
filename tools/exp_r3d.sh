#!/usr/bin/env bash
# round 2, session 2, call 4: L2::64B loads in resolve/apply, 64 sub-chunk states per host state, coarse shift A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shuffle.py -q -x > gpurun_out/r3d_pytest_shuffle.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r3d_pytest_shuffle.log
tail -3 gpurun_out/r3d_pytest_shuffle.log
timeout 200 python tools/shuffle_time.py 4000000 67108864 v=5,4,2,1 > gpurun_out/r3d_shuffle_time.jsonl 2> gpurun_out/r3d_shuffle_time.err
B200SGD_COARSE_SHIFT=16 timeout 200 python tools/shuffle_time.py 67108864 v=5,4 > gpurun_out/r3d_shuffle_time_cs16.jsonl 2>> gpurun_out/r3d_shuffle_time.err
echo "shuffle_time rc=$?"; cat gpurun_out/r3d_shuffle_time.jsonl gpurun_out/r3d_shuffle_time_cs16.jsonl | cut -c1-140
run_bench() { # tag env...
  local tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-sweep --no-e2e > gpurun_out/r3d_bench_$tag.json 2> gpurun_out/r3d_bench_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r3d_bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["value"] / 1e6, 1), round(d["ms_per_step"], 2), round(d["steploop"]["us_per_sgd_step"], 2), round(d["roofline"]["frac"], 3), round(d["roofline"]["kernel_alone"]["us_per_sgd_step"], 2), d["quality"]["permutation_golden_check"][1]["match"])
except Exception as e:
    print("$tag parse failed", e)
PY
}
run_bench v5 B200SGD_SHUFFLE=5
run_bench v4 B200SGD_SHUFFLE=4
run_bench v5_cs15 B200SGD_SHUFFLE=5 B200SGD_COARSE_SHIFT=15
run_bench v5_cs16 B200SGD_SHUFFLE=5 B200SGD_COARSE_SHIFT=16
run_bench v4_cs16 B200SGD_SHUFFLE=4 B200SGD_COARSE_SHIFT=16
run_bench v2 B200SGD_SHUFFLE=2
for t in "5 14" "5 16"; do set -- $t
B200SGD_SHUFFLE=$1 B200SGD_COARSE_SHIFT=$2 timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/r3d_launches_shuffle64m_v$1_cs$2.csv python tools/prof_targets.py shuffle 67108864 > gpurun_out/r3d_ncu_shuffle_v$1_cs$2.log 2>&1
echo "ncu v$1 cs$2 rc=$?"
done
