#!/usr/bin/env bash
# one gpurun call: ncu captures of the round-2 kernels (each target first runs plain, then under ncu);
# the reports stay on the box (19 MB each with sources), their raw / source pages come back as CSV
mkdir -p gpurun_out /tmp/ncu
export B200SGD_NO_COOP=1
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name regex skip args...
  local name=$1; local rx=$2; local skip=$3; shift 3
  python tools/prof_targets.py "$@" > gpurun_out/ncu_${name}_plain.log 2>&1 && \
  $NCU -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/r2_$name python tools/prof_targets.py "$@" > gpurun_out/ncu_${name}.log 2>&1
  echo "$name rc=$?"; tail -4 gpurun_out/ncu_${name}.log | cut -c1-300
  if [ -f /tmp/ncu/r2_$name.ncu-rep ]; then
    ncu -i /tmp/ncu/r2_$name.ncu-rep --page raw --csv > gpurun_out/r2_${name}_raw.csv 2>/dev/null
    ncu -i /tmp/ncu/r2_$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2_${name}_source.csv.gz
  fi
}
cap epoch_c2048_b4096 sgd_epoch_kernel 1 epoch 1048576 2048 4096
cap epoch_c512_b8192 sgd_epoch_kernel 1 epoch 2097152 512 8192
cap epoch_c512_b65536 sgd_epoch_kernel 1 epoch 4194304 512 65536
cap epoch_c2048_b65536 sgd_epoch_kernel 1 epoch 1048576 2048 65536
python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --no-e2e > gpurun_out/ncu_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --no-e2e > gpurun_out/ncu_bench.log 2>&1
echo "bench list rc=$?"; tail -3 gpurun_out/ncu_bench.log | cut -c1-300
du -sh gpurun_out
