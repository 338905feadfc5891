#!/bin/bash
mkdir -p gpurun_out/c20
( time timeout 900 python -m pytest tests/test_mtf_huff_gpu.py tests/test_encode_gpu.py tests/test_fuzz_gpu.py -x -q ) > gpurun_out/c20/pytest.log 2>&1
grep -E "passed|failed" gpurun_out/c20/pytest.log | tail -2
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'value',l['value'],'e2e',l['e2e']['value'],'stages',l['stage_ms'], l['parity_check']['sha256'][:8])
"; }
run
run --set mtf_overlap=0
run --workload random-1GiB-L9
run --workload random-1GiB-L9 --set mtf_overlap=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mtf_ -c 12 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --set h2d_overlap=0 --set mtf_overlap=0 2>&1 | grep -E "mtf_|gpu__time" | paste - - | awk '{print $1, $(NF)}' | head -12
