#!/bin/bash
# GPU tests + headline bench (+ optional env variants given as arguments, one quoted string each)
set -u
mkdir -p gpurun_out
L=gpurun_out/round_check.log
: > $L
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu.log)" >> $L
echo "== default" >> $L
timeout 300 python bench.py --frames 33152 --steps 3 --warmup 3 --no-cpu-baseline >> $L 2>&1
for cfg in "$@"; do
  echo "== $cfg" >> $L
  env $cfg timeout 300 python bench.py --frames 33152 --steps 3 --warmup 3 --no-cpu-baseline >> $L 2>&1
done
python - <<'PY'
import json
for line in open('gpurun_out/round_check.log'):
    if line.startswith('=='): print(line.strip()); continue
    if line.startswith('{'):
        d = json.loads(line); print(round(d['value']), round(d['ms_per_step'], 2), round(d['e2e']['value']))
        kb = d.get('kernel_breakdown', {})
        for k, v in sorted(kb.items(), key=lambda kv: -kv[1]['ms'])[:14]: print('   ', k, v['ms'], v.get('launches'))
    elif 'rc=' in line or 'Error' in line or 'error' in line: print(line.strip()[:300])
PY
