#!/usr/bin/env bash
mkdir -p gpurun_out
{
for sk in 0 1 2 4 3 7; do
python tools/sweep2.py one 2097152 512 8192 CLT=8 CLUSTER=0 PROFILE=1 DBGSKIP=$sk
done
python tools/sweep2.py one 2097152 512 8192 CLT=8 CLUSTER=0 PROFILE=1 VARIANT=7
python tools/sweep2.py one 2097152 512 8192 CLT=8 CLUSTER=0 PROFILE=1 VARIANT=8
python tools/sweep2.py one 2097152 512 65536 CLT=8 CLUSTER=0 PROFILE=1
python tools/sweep2.py one 2097152 512 65536 CLT=8 CLUSTER=0 PROFILE=1 VARIANT=8
python tools/sweep2.py one 1048576 2048 4096 CLT=8 CLUSTER=0 PROFILE=1
python tools/sweep2.py one 1048576 2048 4096 CLT=16 CLUSTER=0 PROFILE=1
python tools/sweep2.py one 1048576 2048 4096 CLT=8 CLUSTER=0 PROFILE=1 VARIANT=7
python tools/sweep2.py one 1048576 2048 65536 CLT=8 CLUSTER=0 PROFILE=1
} > gpurun_out/r2g.jsonl 2> gpurun_out/r2g.err
