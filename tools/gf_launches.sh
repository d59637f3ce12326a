#!/bin/bash
# per-kernel times of one gf_probe.py run (ncu launch list): prints the median per kernel name
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/_l.csv python tools/gf_probe.py > /dev/null 2>&1
python - <<'PY'
import csv, statistics
from collections import defaultdict
lines=[l for l in open('gpurun_out/_l.csv') if not l.startswith('==')]
d=defaultdict(list)
for r in csv.DictReader(lines):
    if r.get('Metric Name')=='gpu__time_duration.sum' and 'rf::' in r['Kernel Name']:
        d[r['Kernel Name'].split('(')[0]].append(float(r['Metric Value'])/1e3)
for k,v in d.items(): print("%-40s n=%2d median %.1f us min %.1f" % (k[-40:], len(v), statistics.median(v), min(v)))
PY
