#!/bin/bash
# 8-GPU call: multi-GPU tests, the strong-scaling bench at N=8 (1 GiB and 8 GiB), N=1 at 8 GiB on the same box
mkdir -p gpurun_out/c8
nvidia-smi -L > gpurun_out/c8/gpus.txt; nproc >> gpurun_out/c8/gpus.txt
( time timeout 900 python -m pytest tests/test_multigpu.py tests/test_sharded_gpu.py -q -v ) > gpurun_out/c8/pytest_multigpu.log 2>&1
tail -25 gpurun_out/c8/pytest_multigpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/c8/bench_n8.json 2> gpurun_out/c8/bench_n8.err
tail -c 2800 gpurun_out/c8/bench_n8.json; tail -3 gpurun_out/c8/bench_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 --no-replicas > gpurun_out/c8/bench_n4.json 2> gpurun_out/c8/bench_n4.err
tail -c 600 gpurun_out/c8/bench_n4.json | head -c 600; echo
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c8/bench_n1.json 2> gpurun_out/c8/bench_n1.err
head -c 300 gpurun_out/c8/bench_n1.json; echo
timeout 900 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 3 --warmup 1 --workload mixed-8GiB-L9 --no-replicas > gpurun_out/c8/bench8g_n8.json 2> gpurun_out/c8/bench8g_n8.err
tail -c 2500 gpurun_out/c8/bench8g_n8.json; tail -3 gpurun_out/c8/bench8g_n8.err
timeout 900 python bench.py --gpus 1 --steps 2 --warmup 1 --workload mixed-8GiB-L9 --no-cpu-baseline > gpurun_out/c8/bench8g_n1.json 2> gpurun_out/c8/bench8g_n1.err
tail -c 2500 gpurun_out/c8/bench8g_n1.json; tail -3 gpurun_out/c8/bench8g_n1.err
