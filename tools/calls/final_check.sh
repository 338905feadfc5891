#!/bin/bash
mkdir -p gpurun_out/final
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/final/pytest.log 2>&1
grep -E "passed|failed" gpurun_out/final/pytest.log | tail -2
timeout 600 python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/final/bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','stage_ms','parity_check','gpu_launches','clocks')}, l['e2e'], l['roofline']['frac'], l['cpu_baseline']['value'])
PY
