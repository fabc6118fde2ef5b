#!/usr/bin/env bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538"
timeout 300 $T bench.py --gpus 8 --no-cpu --no-e2e > gpurun_out/r2n_c5_n8_simul.json 2> gpurun_out/r2n.err
B200SGD_MAP_STAGGER=1 timeout 300 $T bench.py --gpus 8 --no-cpu --no-e2e > gpurun_out/r2n_c5_n8_stagger.json 2>> gpurun_out/r2n.err
B200SGD_PHASE_PROFILE=1 timeout 300 $T tools/sweep_multi.py 67108864 512 65536 > gpurun_out/r2n_prof.jsonl 2>> gpurun_out/r2n.err
grep -h "^{" gpurun_out/r2n_prof.jsonl
