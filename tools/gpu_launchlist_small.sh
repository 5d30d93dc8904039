#!/bin/bash
mkdir -p gpurun_out
P="python tools/prof_small.py ${1:-64}"
timeout 300 $P > gpurun_out/prof_small_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_small.csv $P > gpurun_out/ncu_launches_small.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches_small.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    agg.setdefault(row['Kernel Name'][:60],[]).append(float(row['Metric Value'].replace(',','')))
for k,v in agg.items(): print(f"{k:62s} n={len(v):3d} mean_us={sum(v)/len(v)/1e3:9.2f} min_us={min(v)/1e3:9.2f}")
PY
