#!/bin/bash
# GPU call 1 (2 GPUs): full GPU test suite, 1-GPU bench both arms, 2-GPU torchrun bench
mkdir -p gpurun_out/c1
nvidia-smi -L > gpurun_out/c1/gpus.txt 2>&1
nproc >> gpurun_out/c1/gpus.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1/pytest.log 2>&1
tail -5 gpurun_out/c1/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c1/bench1.json 2> gpurun_out/c1/bench1.err
tail -c 3000 gpurun_out/c1/bench1.json; tail -3 gpurun_out/c1/bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c1/bench2.json 2> gpurun_out/c1/bench2.err
tail -c 3000 gpurun_out/c1/bench2.json; tail -5 gpurun_out/c1/bench2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c1/ref.json 2> gpurun_out/c1/ref.err
tail -c 1500 gpurun_out/c1/ref.json
