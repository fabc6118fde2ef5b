#!/usr/bin/env bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2u_pytest.log 2>&1; tail -2 gpurun_out/r2u_pytest.log
timeout 300 $T bench.py --gpus 2 --no-cpu --no-e2e > gpurun_out/r2u_n2_confined.json 2> gpurun_out/r2u.err; echo rc=$?
B200SGD_MAP_COLOCATE=1 timeout 300 $T bench.py --gpus 2 --no-cpu --no-e2e > gpurun_out/r2u_n2_colocate.json 2>> gpurun_out/r2u.err; echo rc=$?
