#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
tail -3 gpurun_out/r2s_pytest.log
{
python tools/sweep2.py one 1048576 2048 4096 CLUSTER=0
python tools/sweep2.py one 2097152 512 8192 CLUSTER=0
python tools/sweep2.py one 4194304 512 65536 CLUSTER=0
python tools/sweep2.py one 1048576 2048 65536 CLUSTER=0
python tools/sweep2.py one 200000 2000 1000 CLUSTER=0
python tools/sweep2.py one 4000000 100 100000
B200SGD_NO_GATHER4=1 python tools/sweep2.py one 4000000 100 100000 PRODUCER_WARPS=4
B200SGD_NO_GATHER4=1 python tools/sweep2.py one 4000000 100 100000 PRODUCER_WARPS=2
python tools/sweep2.py one 4000000 100 100000 grid=148 reduce=3
python tools/sweep2.py one 4000000 100 30000
} > gpurun_out/r2s_sweep.jsonl 2> gpurun_out/r2s_sweep.err
cat gpurun_out/r2s_sweep.jsonl | cut -c1-330
( time python bench.py --impl reference ) > gpurun_out/r2s_ref.json 2> gpurun_out/r2s_ref.err
echo "ref rc=$?"; tail -4 gpurun_out/r2s_ref.err
( time python bench.py ) > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?"; tail -4 gpurun_out/r2s_bench.err
