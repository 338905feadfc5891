#!/bin/bash
# ncu: full capture of the dominant kernel, launch list of a bench run; then the big configs at full size
mkdir -p gpurun_out/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bwt_sort -c 1 -o gpurun_out/bwt_r2_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --set h2d_overlap=0 > gpurun_out/ncu/ncu_full.log 2>&1
tail -2 gpurun_out/ncu/ncu_full.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu/ncu_launch.log | cut -c1-300
( time timeout 1200 python tools/run_configs.py big ) > gpurun_out/configs_r2_big.txt 2>&1
cut -c1-400 gpurun_out/configs_r2_big.txt
