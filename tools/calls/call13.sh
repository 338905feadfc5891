#!/bin/bash
mkdir -p gpurun_out/c13
( time timeout 900 python -m pytest tests/test_bwt_gpu.py tests/test_fuzz_gpu.py -x -q ) > gpurun_out/c13/pytest_bwt.log 2>&1
tail -4 gpurun_out/c13/pytest_bwt.log
for k in ab period1000 binary text mixed; do timeout 300 python tools/bwt_perf.py $k 296 9 0 2>&1 | tail -1; done | tee gpurun_out/c13/perf.log
timeout 300 python tools/bwt_perf.py ab 75 9 0,8 2>&1 | tail -2 | tee -a gpurun_out/c13/perf.log
timeout 300 python tools/bwt_perf.py period1000 75 9 0,8 2>&1 | tail -2 | tee -a gpurun_out/c13/perf.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c13/bench1.json 2> gpurun_out/c13/bench1.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/c13/bench1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','stage_ms','parity_check')}, l['e2e'])
PY
tail -3 gpurun_out/c13/bench1.err
