#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2x_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
tail -3 gpurun_out/r2x_pytest.log
{
python tools/sweep2.py one 4000000 100 100000
python tools/sweep2.py one 4000000 100 100000 VARIANT=4
python tools/sweep2.py one 4000000 100 30000
python tools/sweep2.py one 4000000 100 30000 VARIANT=4
python tools/sweep2.py one 4000000 100 16384
python tools/sweep2.py one 4000000 100 100000 PROFILE=1
} > gpurun_out/r2x_sweep.jsonl 2> gpurun_out/r2x_sweep.err
cat gpurun_out/r2x_sweep.jsonl | cut -c1-400
