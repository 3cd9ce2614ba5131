#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"plane_x_kernel|plane_t_kernel|plane_xs_kernel" --csv --log-file gpurun_out/r02q_probe.csv python tools/fold_probe.py 2072 > gpurun_out/r02q_probe.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r02q_probe.csv')))
hdr = None; out = {}
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    out.setdefault(int(d['ID']), {'k': d['Kernel Name'][:60]})[d['Metric Name']] = d['Metric Value']
for i in sorted(out)[:60]:
    print(i, out[i])
PY
