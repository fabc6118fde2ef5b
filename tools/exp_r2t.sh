#!/usr/bin/env bash
mkdir -p gpurun_out
for kb in 0 40; do
B200SGD_SHUF_SMEM_KB=$kb timeout 300 python bench.py --no-cpu --no-e2e --no-sweep > gpurun_out/r2t_bench_smem$kb.json 2> gpurun_out/r2t_smem$kb.err
echo "smem $kb rc=$?"
done
B200SGD_SHUF_SMEM_KB=40 python tools/shuffle_time.py 67108864 > gpurun_out/r2t_shuffle40.jsonl 2>&1; cat gpurun_out/r2t_shuffle40.jsonl
