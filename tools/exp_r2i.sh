#!/usr/bin/env bash
mkdir -p gpurun_out
{
for R in 2097152 8388608 16777216 33554432 67108864; do
python tools/sweep2.py one $R 512 65536 CLUSTER=0
done
python tools/sweep2.py one 67108864 512 65536 CLUSTER=0 PROFILE=1
python tools/sweep2.py one 8388608 512 8192 CLUSTER=0
python tools/sweep2.py one 16777216 2048 4096 CLUSTER=0 PROFILE=1
} > gpurun_out/r2i.jsonl 2> gpurun_out/r2i.err
