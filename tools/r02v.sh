#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r02v_bench_default.json 2> gpurun_out/r02v_bench_default.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02v_bench_default.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'unprepared', d['unprepared'])
print('roofline', {k: d['roofline'][k] for k in ('kernel', 'frac', 'achieved', 'launch_ms', 'share_of_step')})
for k, v in d['sub_records'].items():
    if 'rows' in v: print(k, [(r['codecs'], r['frames_per_gpu'], round(r['value']), round(r['e2e'])) for r in v['rows']])
    else: print(k, v.get('value'), v.get('ms_per_step', v.get('ms_per_call')), v.get('e2e', ''), v.get('error', ''))
PY
tail -3 gpurun_out/r02v_bench_default.err
