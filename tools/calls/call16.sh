#!/bin/bash
mkdir -p gpurun_out/c16
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/c16/pytest.log 2>&1
tail -4 gpurun_out/c16/pytest.log
for k in text mixed; do timeout 300 python tools/bwt_perf.py $k 296 9 0 2>&1 | tail -1; done | tee gpurun_out/c16/perf.log
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'value',l['value'],'e2e',l['e2e']['value'],'stages',l['stage_ms'], 'frac', l['roofline']['frac'], l['parity_check'])
"; }
run
run --workload mixed-256MiB-L9 --set bwt_cluster_below=400
run --workload mixed-256MiB-L9 --set bwt_cluster_below=400 --set mtf_overlap=0
