#!/bin/bash
# step time against the MTF-beside-the-sort knobs (1 GiB mixed, level 9, input resident)
out=gpurun_out/${1:-sweep}
mkdir -p $out
for cfg in "70 2" "80 2" "85 3" "90 3" "80 4" "95 4"; do
  set -- $cfg
  timeout 200 python bench.py --no-cpu-baseline --steps 4 --warmup 2 --set mtf_overlap=$1 --set mtf_groups=$2 > $out/b_$1_$2.json 2> $out/b_$1_$2.err
  python - <<PY
import json
l = json.loads(open("$out/b_$1_$2.json").read().strip().splitlines()[-1])
print("mtf_overlap=$1 mtf_groups=$2", l["ms_per_step"], l["e2e"]["ms_per_step"], l["stage_ms"])
PY
done
