#!/usr/bin/env bash
# round 2, GPU call A: parity of the new step tail + the tail / in-flight matrix
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -3 gpurun_out/r2a_pytest.log
B200SGD_DEBUG=1 timeout 120 python tools/sweep2.py one 1048576 2048 4096 CLT=8 > gpurun_out/r2a_dbg.log 2>&1
timeout 1500 python tools/sweep2.py tail > gpurun_out/r2a_tail.jsonl 2> gpurun_out/r2a_tail.err
echo "sweep rc=$?"
tail -5 gpurun_out/r2a_tail.jsonl
