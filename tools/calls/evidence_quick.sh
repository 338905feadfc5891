#!/bin/bash
# full GPU test suite, smoke, bench (both arms)
out=gpurun_out/${1:-evq}
mkdir -p $out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
( time timeout 1500 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
grep -E "passed|failed" $out/pytest.log | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err
tail -c 600 $out/bench.json; tail -2 $out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err
head -c 300 $out/bench_reference.json; echo
