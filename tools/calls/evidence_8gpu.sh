#!/bin/bash
# final 8-GPU evidence: multi-GPU tests, strong scaling 1 GiB (N=8, N=1 on the same box), 8 GiB (N=8, N=1)
mkdir -p gpurun_out/ev8
nproc > gpurun_out/ev8/host.txt
( time timeout 600 python -m pytest tests/test_multigpu.py tests/test_sharded_gpu.py -q -v ) > gpurun_out/ev8/pytest_multigpu.log 2>&1
grep -E "passed|failed" gpurun_out/ev8/pytest_multigpu.log | tail -1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-replicas > gpurun_out/ev8/bench_n8.json 2> gpurun_out/ev8/bench_n8.err
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ev8/bench_n1.json 2> gpurun_out/ev8/bench_n1.err
timeout 900 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 3 --warmup 2 --workload mixed-8GiB-L9 --no-replicas > gpurun_out/ev8/bench8g_n8.json 2> gpurun_out/ev8/bench8g_n8.err
timeout 900 python bench.py --gpus 1 --steps 2 --warmup 1 --workload mixed-8GiB-L9 --no-cpu-baseline > gpurun_out/ev8/bench8g_n1.json 2> gpurun_out/ev8/bench8g_n1.err
python - <<'PY'
import json
for f in ("bench_n1","bench_n8","bench8g_n1","bench8g_n8"):
    try:
        l=json.loads(open(f'gpurun_out/ev8/{f}.json').read().strip().splitlines()[-1])
        print(f, "value",l['value'],"e2e",l['e2e']['value'],"ms",l['ms_per_step'],"stages",l['stage_ms'],"parity",l.get('parity_check'))
    except Exception as e:
        print(f, "ERR", e); print(open(f'gpurun_out/ev8/{f}.err').read()[-800:])
PY
