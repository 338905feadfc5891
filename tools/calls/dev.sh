#!/bin/bash
# development call: sort parity on both kernels, then sort timings (mixed / text / periodic)
out=gpurun_out/${1:-dev}
mkdir -p $out
( time timeout 600 python -m pytest tests/test_bwt_gpu.py -q -x ) > $out/pytest.log 2>&1
tail -4 $out/pytest.log
{
timeout 120 python tools/bwt_perf.py mixed 600 9 0
timeout 120 python tools/bwt_perf.py text 296 9 0
timeout 120 python tools/bwt_perf.py text 137 9 8
timeout 120 python tools/bwt_perf.py mixed 137 9 8
timeout 120 python tools/bwt_perf.py ab 75 9 -1,0
timeout 120 python tools/bwt_perf.py period1000 75 9 -1,0
} > $out/perf.txt 2>&1
cat $out/perf.txt
