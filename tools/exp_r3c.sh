#!/usr/bin/env bash
mkdir -p gpurun_out
./tools/microbench_gather > gpurun_out/r3c_gather.jsonl 2>&1
./tools/microbench_gather 32 > gpurun_out/r3c_gather_l2f32.jsonl 2>&1
cat gpurun_out/r3c_gather.jsonl gpurun_out/r3c_gather_l2f32.jsonl
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum --clock-control none --csv --log-file gpurun_out/r3c_gather_ncu.csv ./tools/microbench_gather > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r3c_gather_ncu_l2f32.csv ./tools/microbench_gather 32 > /dev/null 2>&1
echo done
