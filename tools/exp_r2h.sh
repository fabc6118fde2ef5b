#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=memory.total,memory.used --format=csv > gpurun_out/r2h_smi.txt; free -g >> gpurun_out/r2h_smi.txt; nproc >> gpurun_out/r2h_smi.txt
( time python bench.py --steps 3 --warmup 3 --rows 4194304 --e2e-rows 1048576 --cpu-seconds 3 ) > gpurun_out/r2h_small.json 2> gpurun_out/r2h_small.err
echo "small rc=$?"
( time python bench.py ) > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"
tail -5 gpurun_out/r2h_bench.err
