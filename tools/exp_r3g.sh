#!/usr/bin/env bash
# round 2, session 2, call 7: ncu evidence for the backward shuffle (full captures of its four main kernels at 64M
# rows, launch list of the default bench run) and the default bench line of this build
mkdir -p gpurun_out /tmp/ncu
timeout 400 python bench.py > gpurun_out/r3g_bench_n1.json 2> gpurun_out/r3g_bench_n1.err
echo "bench rc=$?"
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name regex skip
  local name=$1; local rx=$2; local skip=$3
  $NCU -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/r3_$name python tools/prof_targets.py shuffle 67108864 > gpurun_out/ncu_${name}.log 2>&1
  echo "$name rc=$?"
  if [ -f /tmp/ncu/r3_$name.ncu-rep ]; then
    ncu -i /tmp/ncu/r3_$name.ncu-rep --page raw --csv > gpurun_out/r3_${name}_raw.csv 2>/dev/null
    ncu -i /tmp/ncu/r3_$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r3_${name}_source.csv.gz
  fi
}
python tools/prof_targets.py shuffle 67108864 > gpurun_out/ncu_shuffle_plain.log 2>&1 && {
cap shuf2_resolve shuf2_resolve_kernel 1
cap shuf2_link shuf2_link_kernel 1
cap shuf2_stream_scatter "shuf2_stream_kernel<1" 1
cap shuf2_refine shuf2_refine_kernel 1
}
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/r3g_launches_shuffle64m.csv python tools/prof_targets.py shuffle 67108864 > gpurun_out/r3g_ncu_shuffle.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --no-e2e > gpurun_out/r3g_bench_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3g_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --no-e2e > gpurun_out/r3g_ncu_bench.log 2>&1
echo "bench list rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r3g_bench_n1.json").read().strip().splitlines()[-1])
print(round(d["value"] / 1e6, 1), round(d["ms_per_step"], 2), round(d["steploop"]["us_per_sgd_step"], 2), round(d["roofline"]["frac"], 3), round(d["roofline"]["kernel_alone"]["us_per_sgd_step"], 2), (d.get("e2e") or {}).get("value"), d["gpu_launches"])
PY
du -sh gpurun_out
