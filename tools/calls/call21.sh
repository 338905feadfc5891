#!/bin/bash
mkdir -p gpurun_out/c21
for wl in mixed-1GiB-L9 random-1GiB-L9; do
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:mtf_ -c 4 --csv --log-file gpurun_out/c21/mtf_$wl.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --set h2d_overlap=0 --set mtf_overlap=0 --workload $wl > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/c21/mtf_$wl.csv')) if len(r)>5]
h=next(r for r in rows if 'Kernel Name' in r); i0=rows.index(h)+1
ki,mi,vi=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value')
for r in rows[i0:]:
    print('$wl', r[ki].split('(')[0], r[mi], r[vi])
PY
done
