#!/bin/bash
mkdir -p gpurun_out/c9
( time timeout 900 python -m pytest tests/test_verify_gpu.py tests/test_sharded_gpu.py tests/test_multigpu.py -x -q ) > gpurun_out/c9/pytest.log 2>&1
tail -12 gpurun_out/c9/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-replicas > gpurun_out/c9/bench2.json 2> gpurun_out/c9/bench2.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/c9/bench2.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','stage_ms','parity_check')}, l['e2e'])
PY
tail -3 gpurun_out/c9/bench2.err
