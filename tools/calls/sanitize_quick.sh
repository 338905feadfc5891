#!/bin/bash
# compute-sanitizer, short form: memcheck and synccheck over both targets, racecheck over the BWT batch
# (racecheck over the encode target takes ~8 min: tools/calls/sanitize.sh runs all six)
mkdir -p gpurun_out/san
CS=/usr/local/cuda/bin/compute-sanitizer
run() {
  ( time timeout 900 $CS --tool $1 --print-limit 20 python tools/sanitize_target.py $2 200 ) > gpurun_out/san/$1_$2.log 2>&1
  echo "== $1 $2: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|real" gpurun_out/san/$1_$2.log | tail -4
}
run memcheck encode
run memcheck bwt
run synccheck encode
run synccheck bwt
run racecheck bwt
