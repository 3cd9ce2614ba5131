#!/bin/bash
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,lts__t_bytes.sum --clock-control none -k regex:wgrad_tc -c 120 --csv --log-file gpurun_out/r02k_wgrad_ncu.csv python bench.py --workload train --steps 1 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r02k_wgrad_ncu.csv')))
hdr = None
out = {}
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    key = (d['ID'], d['Kernel Name'][:40], d['Grid Size'] if 'Grid Size' in d else '')
    out.setdefault(key, {})[d['Metric Name']] = d['Metric Value']
n = 0
for k, v in out.items():
    print(k, v)
    n += 1
    if n > 40: break
PY
