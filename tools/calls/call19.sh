#!/bin/bash
mkdir -p gpurun_out/c19
( time timeout 900 python -m pytest tests/test_bwt_gpu.py -x -q ) > gpurun_out/c19/pytest.log 2>&1
tail -3 gpurun_out/c19/pytest.log | head -1
for k in text mixed; do timeout 300 python tools/bwt_perf.py $k 296 9 0 2>&1 | tail -1; done
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'value',l['value'],'e2e',l['e2e']['value'],'stages',l['stage_ms'], l['parity_check']['sha256'][:8])
"; }
run
