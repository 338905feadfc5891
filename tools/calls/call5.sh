#!/bin/bash
mkdir -p gpurun_out/c5
( time timeout 900 python -m pytest tests/test_bwt_gpu.py -x -q ) > gpurun_out/c5/pytest_bwt.log 2>&1
tail -4 gpurun_out/c5/pytest_bwt.log
for tok in 0 8 16 24 32 48 96; do timeout 300 python tools/bwt_perf.py text 296 9 0 bwt_tokens=$tok 2>&1 | tail -1 | sed "s/^/tok=$tok /"; done > gpurun_out/c5/perf_tok.log 2>&1
cat gpurun_out/c5/perf_tok.log
for tok in 0 16 24 48; do timeout 300 python tools/bwt_perf.py mixed 600 9 0 bwt_tokens=$tok 2>&1 | tail -1 | sed "s/^/tok=$tok /"; done > gpurun_out/c5/perf_tok_mixed.log 2>&1
cat gpurun_out/c5/perf_tok_mixed.log
timeout 300 python tools/bwt_perf.py text 16 9 0 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c5/bench1.json 2> gpurun_out/c5/bench1.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/c5/bench1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','stage_ms','parity_check','bwt')}, l['e2e']['value'], l['roofline']['frac'])
PY
tail -3 gpurun_out/c5/bench1.err
