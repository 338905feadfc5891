#!/bin/bash
mkdir -p gpurun_out/c6
( time timeout 900 python -m pytest tests/test_bwt_gpu.py -x -q ) > gpurun_out/c6/pytest_bwt.log 2>&1
tail -4 gpurun_out/c6/pytest_bwt.log
for tok in 0 8 12 16 24 32; do timeout 300 python tools/bwt_perf.py text 296 9 0 bwt_tokens=$tok 2>&1 | tail -1 | sed "s/^/tok=$tok /"; done > gpurun_out/c6/perf_tok.log 2>&1
cat gpurun_out/c6/perf_tok.log
for tok in 0 12 16 24; do timeout 300 python tools/bwt_perf.py mixed 600 9 0 bwt_tokens=$tok 2>&1 | tail -1 | sed "s/^/tok=$tok /"; done > gpurun_out/c6/perf_tok_mixed.log 2>&1
cat gpurun_out/c6/perf_tok_mixed.log
timeout 300 python tools/bwt_perf.py text 16 9 0 2>&1 | tail -1
