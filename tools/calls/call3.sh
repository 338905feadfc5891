#!/bin/bash
mkdir -p gpurun_out/c3
( time timeout 900 python -m pytest tests/test_bwt_gpu.py -x -q ) > gpurun_out/c3/pytest_bwt.log 2>&1
tail -15 gpurun_out/c3/pytest_bwt.log
for k in text source binary mixed; do timeout 300 python tools/bwt_perf.py $k 296 9 0 2>&1 | tail -2; done > gpurun_out/c3/perf.log 2>&1
cat gpurun_out/c3/perf.log
timeout 300 python tools/bwt_perf.py mixed 600 9 0,8 >> gpurun_out/c3/perf.log 2>&1; tail -2 gpurun_out/c3/perf.log
( time timeout 900 python -m pytest tests/test_encode_gpu.py tests/test_sharded_gpu.py tests/test_fuzz_gpu.py -x -q ) > gpurun_out/c3/pytest_enc.log 2>&1
tail -5 gpurun_out/c3/pytest_enc.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c3/bench1.json 2> gpurun_out/c3/bench1.err
tail -c 2500 gpurun_out/c3/bench1.json; tail -3 gpurun_out/c3/bench1.err
