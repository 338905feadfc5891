#!/bin/bash
mkdir -p gpurun_out/c17
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/c17/pytest.log 2>&1
grep -E "passed|failed|Error" gpurun_out/c17/pytest.log | tail -3
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'value',l['value'],'e2e',l['e2e']['value'],'stages',l['stage_ms'], l['parity_check'])
"; }
run
run --set post_groups=0
run --set post_groups=2
run --set post_groups=6
run --set post_groups=8
