#!/usr/bin/env bash
# round 2, session 2, call 10: final validation of the build -- whole GPU suite, smoke, the bench launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r3j_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r3j_pytest.log
tail -4 gpurun_out/r3j_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3j_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r3j_smoke.log
B200SGD_NO_COOP=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3j_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --no-e2e > gpurun_out/r3j_ncu_bench.log 2>&1
echo "bench list rc=$?"
