#!/bin/bash
run() { timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'value',l['value'],'e2e',l['e2e']['value'], 'e2e_ms', l['e2e']['ms_per_step'])
"; }
run --set piece_blocks_per_sm_x16=3
run --set piece_blocks_per_sm_x16=4
run --set piece_blocks_per_sm_x16=5
run --set piece_blocks_per_sm_x16=6
run --set piece_blocks_per_sm_x16=7
run --set piece_blocks_per_sm_x16=6 --set h2d_pieces=2
