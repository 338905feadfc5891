#!/bin/bash
# final single-GPU evidence: full test suite, sanitizer, bench (both arms), ncu launch list
mkdir -p gpurun_out/ev1
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/ev1/pytest.log 2>&1
grep -E "passed|failed" gpurun_out/ev1/pytest.log | tail -2
bash tools/calls/sanitize.sh
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2.json 2> gpurun_out/ev1/bench.err
tail -c 900 gpurun_out/bench_r2.json; tail -2 gpurun_out/ev1/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2>> gpurun_out/ev1/bench.err
head -c 300 gpurun_out/bench_r2_reference.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ev1/ncu_launch.log 2>&1
