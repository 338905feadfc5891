#!/bin/bash
# full GPU test suite + bench both arms + ncu launch list + ncu full capture of the sort + configs + sanitizer
mkdir -p gpurun_out/c11
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/c11/pytest.log 2>&1
tail -6 gpurun_out/c11/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/c11/bench.err
tail -c 1200 gpurun_out/bench_r2.json; tail -2 gpurun_out/c11/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c11/ncu_launch.log 2>&1
tail -2 gpurun_out/c11/ncu_launch.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bwt_sort -c 1 -o gpurun_out/bwt_r2_full -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/c11/ncu_full.log 2>&1
tail -2 gpurun_out/c11/ncu_full.log | cut -c1-300
ls -la gpurun_out/bwt_r2_full.ncu-rep
( time timeout 1500 python tools/run_configs.py small ) > gpurun_out/configs_r2_small.txt 2>&1
tail -5 gpurun_out/configs_r2_small.txt | cut -c1-400
