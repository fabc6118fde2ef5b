#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -3 gpurun_out/r2b_pytest.log
timeout 900 python tools/sweep2.py tail > gpurun_out/r2b_tail.jsonl 2> gpurun_out/r2b_tail.err
echo "sweep rc=$?"
