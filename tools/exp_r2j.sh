#!/usr/bin/env bash
# 2 GPUs: multi-GPU parity tests, then the bench at N = 2 (c5 strong) 
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -3 gpurun_out/r2k_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err
echo "bench rc=$?"
tail -3 gpurun_out/r2k_bench_n2.err
