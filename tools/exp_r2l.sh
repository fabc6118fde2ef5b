#!/usr/bin/env bash
# 8 GPUs: multi-GPU parity tests, then strong-scaling runs of c5 and c4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -3 gpurun_out/r2r_pytest.log
run() { # N workload extra...
  local n=$1; local w=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --workload $w --no-cpu "$@" > gpurun_out/r2r_${w}_n$n.json 2> gpurun_out/r2r_${w}_n$n.err
  echo "$w n=$n rc=$?"
}
run 8 c5
run 4 c5 --no-e2e
run 8 c4 --no-e2e
run 4 c4 --no-e2e
run 2 c4 --no-e2e
run 8 c4b --no-e2e
