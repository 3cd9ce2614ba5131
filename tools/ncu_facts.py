"""Turns the raw ncu outputs of tools/ncu_r02.sh into the small, committed evidence files bench.py and the docs cite:
   python tools/ncu_facts.py [tag]      (reads gpurun_out/<tag>_launches.csv, <tag>_step_records.json, <tag>_*.ncu-rep)
   -> profiles/<tag>_launches.csv           every launch of ONE step: layer name (from the library's own launch records, matched by
                                            order), ncu kernel name, device time, DRAM bytes read / written
   -> profiles/<tag>_launches_summary.csv   per layer name: launches, time, share of the step, DRAM bytes, bytes per frame
   -> profiles/<tag>_ncu_summary.json       key metrics of every `--set full` capture (+ raw pages <tag>_<name>_ncu_full.csv)
   -> profiles/<tag>_ncu_facts.json         what bench.py's roofline record reads: per layer dram bytes per launch, tensor-pipe
                                            activity, and the whole step's DRAM bytes per frame"""
import csv
import glob
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__cluster_size', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active']


def short(name):
    name = name.replace('void ', '').replace('(anonymous namespace)::', '').replace('unnamed>::', '').replace('nsc::', '')
    return re.sub(r'\(.*\)$', '', name)


def num(v):
    return float(v.replace(',', '')) if v not in ('', 'n/a') else 0.0


def to_bytes(v, unit):
    return num(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def to_us(v, unit):
    return num(v) * {'ns': 1e-3, 'us': 1, 'usecond': 1, 'msecond': 1e3, 'ms': 1e3, 'second': 1e6, 'nsecond': 1e-3}.get(unit, 1e-3)


def launches(tag):
    path = os.path.join(ROOT, 'gpurun_out', f'{tag}_launches.csv')
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('=='))]
    hdr = rows[0]
    ii, ki, mi, ui, vi = hdr.index('ID'), hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value')
    per = {}
    order = []
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        k = r[ii]
        if k not in per:
            per[k] = {'kernel': short(r[ki]), 'us': 0.0, 'rd': 0.0, 'wr': 0.0, 'ours': 'nsc::' in r[ki]}
            order.append(k)
        if r[mi] == 'gpu__time_duration.sum':
            per[k]['us'] = to_us(r[vi], r[ui])
        elif r[mi] == 'dram__bytes_read.sum':
            per[k]['rd'] = to_bytes(r[vi], r[ui])
        elif r[mi] == 'dram__bytes_write.sum':
            per[k]['wr'] = to_bytes(r[vi], r[ui])
    seq = [per[k] for k in order]
    rec = json.load(open(os.path.join(ROOT, 'gpurun_out', f'{tag}_step_records.json')))
    frames = rec['frames']
    names = [r['name'] for r in rec['records']]
    ours = seq if len(seq) == len(names) else [s for s in seq if s['ours']]     # (a step launches nothing but the library's kernels)
    matched = len(ours) == len(names)
    if matched:
        for s, r in zip(ours, rec['records']):
            s['layer'] = r['name']
            s['alg_bytes'] = r['bytes']
            s['flops'] = r['flops']
    for s in seq:
        s.setdefault('layer', s['kernel'])
    with open(os.path.join(ROOT, 'profiles', f'{tag}_launches.csv'), 'w') as f:
        f.write('layer,kernel,us,dram_read_bytes,dram_write_bytes\n')
        for s in seq:
            f.write(f'"{s["layer"]}","{s["kernel"]}",{s["us"]:.2f},{s["rd"]:.0f},{s["wr"]:.0f}\n')
    agg = {}
    for s in seq:
        a = agg.setdefault(s['layer'], {'n': 0, 'us': 0.0, 'dram': 0.0, 'alg': 0.0, 'kernel': s['kernel']})
        a['n'] += 1; a['us'] += s['us']; a['dram'] += s['rd'] + s['wr']; a['alg'] += s.get('alg_bytes', 0.0)
    tot_us = sum(a['us'] for a in agg.values())
    tot_dram = sum(a['dram'] for a in agg.values())
    with open(os.path.join(ROOT, 'profiles', f'{tag}_launches_summary.csv'), 'w') as f:
        f.write(f'# one device-resident cq2 step of {frames} frames under ncu (serialised, cold caches: compare SHARES); '
                f'layer names matched to ncu rows by launch order: {matched}\n')
        f.write('layer,kernel,launches,total_us,share_of_step,dram_bytes,dram_bytes_per_frame,dram_vs_algorithmic\n')
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
            f.write(f'"{k}","{a["kernel"]}",{a["n"]},{a["us"]:.1f},{a["us"] / tot_us:.4f},{a["dram"]:.0f},{a["dram"] / frames:.0f},'
                    f'{(a["dram"] / a["alg"]) if a["alg"] else float("nan"):.3f}\n')
        f.write(f'"TOTAL","",{len(seq)},{tot_us:.1f},1.0,{tot_dram:.0f},{tot_dram / frames:.0f},\n')
    print(open(os.path.join(ROOT, 'profiles', f'{tag}_launches_summary.csv')).read())
    kernels = {}
    for k, a in agg.items():
        if a['alg']:
            kernels[k] = {'dram_bytes_per_launch': a['dram'] / a['n'], 'algorithmic_bytes_per_launch': a['alg'] / a['n'],
                          'launches_in_step': a['n'], 'ncu_us_per_launch': a['us'] / a['n'], 'share_of_step_ncu': a['us'] / tot_us,
                          'source': f'profiles/{tag}_launches.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one step of {frames} frames)'}
    return {'kernels': kernels, 'whole_step': {'dram_bytes_per_frame': tot_dram / frames, 'frames': frames, 'launches': len(seq),
                                               'source': f'profiles/{tag}_launches_summary.csv'}}


def captures(tag, facts):
    summ = {}
    for rep in sorted(glob.glob(os.path.join(ROOT, 'gpurun_out', f'{tag}_*.ncu-rep'))):
        name = os.path.basename(rep)[len(tag) + 1:-len('.ncu-rep')]
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        open(os.path.join(ROOT, 'profiles', f'{tag}_{name}_ncu_full.csv'), 'w').write(raw)
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            key = f"{name}#{d.get('ID')}"
            summ[key] = {'kernel': short(d.get('Kernel Name', '')), 'grid': d.get('launch__grid_size'),
                         'metrics': {k: d[k] for k in KEYS if k in d}, 'units': {k: u[k] for k in KEYS if k in u}}
    json.dump(summ, open(os.path.join(ROOT, 'profiles', f'{tag}_ncu_summary.json'), 'w'), indent=1)
    print(json.dumps({k: v['metrics'] for k, v in summ.items()}, indent=1))
    return summ


if __name__ == '__main__':
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
    facts = launches(tag)
    summ = captures(tag, facts)
    # tensor-pipe activity of the captured kernels, attached to the layers they are (first launches of a step: see tools/ncu_r02.sh)
    which = {'xs_20to100': 'pX2_k9d1s1_c20to100', 'x_down': 'pX2_k9d1s2_c100to100', 'x_fold': 'pF2_k9d1_c20to20', 'lpc_analyze': 'lpc_analyze'}
    for key, v in summ.items():
        cap = key.split('#')[0]
        layer = which.get(cap)
        if cap == 't_conv1_conv2':
            layer = 'pT2_k9d1_c100to20' if '9, 1>' in v['kernel'] or 'ILi20ELi9ELi1' in v['kernel'] else 'pT2_k9d1_c20to20'
        if layer and layer in facts['kernels']:
            t = v['metrics'].get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')
            facts['kernels'][layer]['sm__pipe_tensor_cycles_active_pct'] = float(t) if t else None
            facts['kernels'][layer]['full_capture'] = f'profiles/{tag}_{cap}_ncu_full.csv'
    json.dump(facts, open(os.path.join(ROOT, 'profiles', f'{tag}_ncu_facts.json'), 'w'), indent=1)
    print(json.dumps(facts['whole_step'], indent=1))
