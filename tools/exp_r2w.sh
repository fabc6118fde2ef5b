#!/usr/bin/env bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2w_smoke.log
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv -k regex:sgd_epoch_kernel"
run() { # tag env...
  local tag=$1; shift
  env "$@" python tools/sweep2.py one 4000000 100 100000 epochs=1 warm=1 > gpurun_out/r2w_${tag}.json 2>&1 && \
  env "$@" ncu $M --log-file gpurun_out/r2w_${tag}_ncu.csv python tools/sweep2.py one 4000000 100 100000 epochs=1 warm=1 > /dev/null 2>&1
  echo "$tag rc=$?"; cat gpurun_out/r2w_${tag}.json | cut -c1-260
}
run pitch448_none X=1
run pitch512_none B200SGD_LD_ALIGN=128
run pitch512_l2_128 B200SGD_LD_ALIGN=128 B200SGD_L2_PROMO=2
run pitch448_l2_128 B200SGD_L2_PROMO=2
run pitch512_l2_256 B200SGD_LD_ALIGN=128 B200SGD_L2_PROMO=3
grep -h "sgd_epoch_kernel" gpurun_out/r2w_*_ncu.csv | cut -d, -f5,13- | head -40
