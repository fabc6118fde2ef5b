#!/usr/bin/env bash
mkdir -p gpurun_out
run() { # N tag env...
  local n=$1; local tag=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --no-cpu --no-e2e > gpurun_out/r2v_n${n}_$tag.json 2>> gpurun_out/r2v.err
  echo "n=$n $tag rc=$?"
}
run 8 confined X=1
run 8 colocate B200SGD_MAP_COLOCATE=1
run 4 confined X=1
run 4 colocate B200SGD_MAP_COLOCATE=1
