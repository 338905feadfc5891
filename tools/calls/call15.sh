#!/bin/bash
mkdir -p gpurun_out/c15
run() { timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'value',l['value'],'e2e',l['e2e']['value'],'bwt',l['stage_ms']['bwt_ms'],'mtf',l['stage_ms']['mtf_ms'],'total',l['stage_ms']['total_ms'], 'frac', l['roofline']['frac'])
"; }
run
run --set bwt_ctas_per_sm=1
run --set mtf_overlap=50
run --set mtf_overlap=85
run --set mtf_overlap=85 --set mtf_groups=4
run --set mtf_overlap=95 --set mtf_groups=4
run --set h2d_pieces=2
run --set h2d_pieces=4
run --set piece_blocks_per_sm_x16=12
run --set piece_blocks_per_sm_x16=24
