#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -3 gpurun_out/r2p_pytest.log
python tools/shuffle_time.py 4000000 16777216 67108864 > gpurun_out/r2p_shuffle.jsonl 2>&1
cat gpurun_out/r2p_shuffle.jsonl
timeout 600 python bench.py --no-sweep --no-cpu > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
echo "bench rc=$?"
