#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -3 gpurun_out/r2o_pytest.log
df -h /tmp | tail -1
timeout 900 python tools/ingest_bench.py > gpurun_out/r2o_ingest.json 2> gpurun_out/r2o_ingest.err
echo "ingest rc=$?"; cat gpurun_out/r2o_ingest.json; tail -3 gpurun_out/r2o_ingest.err
