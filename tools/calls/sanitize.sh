#!/bin/bash
# compute-sanitizer over the path (memcheck, racecheck, synccheck, initcheck); summaries -> gpurun_out/san/
mkdir -p gpurun_out/san
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for part in encode bwt; do
    ( time timeout 1500 $CS --tool $tool --print-limit 20 python tools/sanitize_target.py $part 200 ) > gpurun_out/san/${tool}_${part}.log 2>&1
    echo "== $tool $part: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|real" gpurun_out/san/${tool}_${part}.log | tail -4
  done
done
