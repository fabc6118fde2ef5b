#!/usr/bin/env bash
mkdir -p gpurun_out
{
for cs in 2 4 8 16; do
python tools/sweep2.py one 4194304 512 65536 CLT=$cs CLUSTER=0
python tools/sweep2.py one 2097152 512 8192 CLT=$cs CLUSTER=0
python tools/sweep2.py one 1048576 2048 4096 CLT=$cs CLUSTER=0
python tools/sweep2.py one 4000000 100 100000 CLT=$cs CLUSTER=0
done
} > gpurun_out/r2y_sweep.jsonl 2> gpurun_out/r2y_sweep.err
cut -c1-330 gpurun_out/r2y_sweep.jsonl
