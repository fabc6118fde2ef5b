#!/usr/bin/env bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
{
B200SGD_PHASE_PROFILE=1 $T tools/sweep_multi.py 16777216 512 16384
B200SGD_PHASE_PROFILE=1 $T tools/sweep_multi.py 16777216 512 131072
$T tools/sweep_multi.py 16777216 512 16384
$T tools/sweep_multi.py 16777216 512 131072
$T tools/sweep_multi.py 4194304 2048 8192
} > gpurun_out/r2m.jsonl 2> gpurun_out/r2m.err
timeout 600 $T bench.py --gpus 2 --no-cpu --no-e2e > gpurun_out/r2m_bench_n2.json 2>> gpurun_out/r2m.err
grep -h "^{" gpurun_out/r2m.jsonl
