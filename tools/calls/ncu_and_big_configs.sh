#!/bin/bash
mkdir -p gpurun_out/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bwt_sort -c 1 -o gpurun_out/bwt_r2_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --set h2d_overlap=0 > gpurun_out/ncu/ncu_full.log 2>&1
tail -2 gpurun_out/ncu/ncu_full.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:mtf_apply -c 1 -o gpurun_out/mtf_apply_r2 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --set h2d_overlap=0 --set mtf_overlap=0 > gpurun_out/ncu/ncu_mtf.log 2>&1
( time timeout 2400 python tools/run_configs.py big ) > gpurun_out/configs_r2_big.txt 2>&1
cat gpurun_out/configs_r2_big.txt | cut -c1-500
