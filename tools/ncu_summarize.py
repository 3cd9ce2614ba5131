"""Turns the raw ncu outputs of tools/ncu_round.sh into the small summaries kept under profiles/:
   python tools/ncu_summarize.py <tag>   (reads gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_*.ncu-rep)
   -> profiles/<tag>_launches_summary.csv (per kernel: launches, total us, share of the step)
   -> profiles/<tag>_ncu_summary.json     (key metrics of every full capture)
   -> profiles/<tag>_<name>_ncu_full.csv  (raw page of every capture)"""
import csv
import re
import glob
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__cluster_size', 'smsp__inst_executed.sum']


def launches(tag):
    path = os.path.join(ROOT, 'gpurun_out', f'{tag}_launches.csv')
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('=='))]
    hdr = rows[0]
    ki, mi, vi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
    agg = {}
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
            continue
        name = r[ki].replace('void ', '').replace('unnamed>::', '').replace('nsc::', '')
        name = re.sub(r'\([A-Za-z_ ,*:<>0-9]*\)$', '', name)      # drop the parameter list
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', '')) / 1e3     # ns -> us
    tot = sum(a[1] for a in agg.values())
    out = os.path.join(ROOT, 'profiles', f'{tag}_launches_summary.csv')
    with open(out, 'w') as f:
        f.write('kernel,launches,total_us,share\n')
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{n},{us:.1f},{us / tot:.4f}\n')
    print(open(out).read())


def captures(tag):
    summ = {}
    for rep in sorted(glob.glob(os.path.join(ROOT, 'gpurun_out', f'{tag}_*.ncu-rep'))):
        name = os.path.basename(rep)[len(tag) + 1:-len('.ncu-rep')]
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        open(os.path.join(ROOT, 'profiles', f'{tag}_{name}_ncu_full.csv'), 'w').write(raw)
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        summ[name] = {'kernel': d.get('Kernel Name'), 'metrics': {k: d[k] for k in KEYS if k in d}, 'units': {k: u[k] for k in KEYS if k in u}}
    json.dump(summ, open(os.path.join(ROOT, 'profiles', f'{tag}_ncu_summary.json'), 'w'), indent=1)
    print(json.dumps({k: v['metrics'] for k, v in summ.items()}, indent=1))


if __name__ == '__main__':
    tag = sys.argv[1]
    launches(tag)
    captures(tag)
