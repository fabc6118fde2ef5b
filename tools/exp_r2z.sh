#!/usr/bin/env bash
mkdir -p gpurun_out
{
python tools/sweep2.py one 4194304 512 65536 CLUSTER=0 seq=1
python tools/sweep2.py one 4194304 512 65536 CLUSTER=0
python tools/sweep2.py one 4000000 100 100000 seq=1
python tools/sweep2.py one 4000000 100 100000
python tools/sweep2.py one 2097152 512 8192 CLUSTER=0 seq=1
python tools/sweep2.py one 1048576 2048 65536 CLUSTER=0 seq=1
python tools/sweep2.py one 4194304 256 65536 CLUSTER=0
python tools/sweep2.py one 4194304 1024 65536 CLUSTER=0
} > gpurun_out/r2z_sweep.jsonl 2> gpurun_out/r2z_sweep.err
