#!/bin/bash
mkdir -p gpurun_out/c4
for nb in 16 28 56 112 148 296; do timeout 300 python tools/bwt_perf.py text $nb 9 0 2>&1 | tail -1; done > gpurun_out/c4/perf_nb.log 2>&1
cat gpurun_out/c4/perf_nb.log
