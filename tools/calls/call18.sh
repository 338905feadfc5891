#!/bin/bash
mkdir -p gpurun_out/c18
( time timeout 900 python -m pytest tests/test_mtf_huff_gpu.py tests/test_cli.py -x -q ) > gpurun_out/c18/pytest.log 2>&1
tail -12 gpurun_out/c18/pytest.log
run() { timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'value',l['value'],'e2e',l['e2e']['value'],'stages',l['stage_ms'], l['parity_check'])
"; }
run --set huff_literal=1
