#!/bin/bash
# development call: MTF / Huffman parity, then the bench line without the CPU legs (mixed and random)
out=gpurun_out/${1:-devm}
mkdir -p $out
( time timeout 600 python -m pytest tests/test_mtf_huff_gpu.py -q -x ) > $out/pytest.log 2>&1
tail -4 $out/pytest.log
timeout 300 python bench.py --no-cpu-baseline > $out/bench.json 2> $out/bench.err
timeout 300 python bench.py --no-cpu-baseline --workload random-1GiB-L9 > $out/bench_random.json 2> $out/bench_random.err
timeout 300 python bench.py --no-cpu-baseline --set mtf_overlap=0 --steps 3 --warmup 2 > $out/bench_serial.json 2> $out/bench_serial.err
python - <<PY
import json
for f in ("bench.json", "bench_random.json", "bench_serial.json"):
    try:
        l = json.loads(open("$out/" + f).read().strip().splitlines()[-1])
        print(f, l["value"], l["ms_per_step"], l["e2e"]["value"], l["stage_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
