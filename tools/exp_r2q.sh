#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -4 gpurun_out/r2q_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err
echo "bench rc=$?"
