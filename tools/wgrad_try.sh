#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/wgrad_try.log
: > $L
timeout 900 python -m pytest tests/test_gpu_training.py -x -q >> $L 2>&1
echo "pytest rc=$?" >> $L
for cfg in "X=1"; do
  echo "== $cfg" >> $L
  env $cfg timeout 300 python bench.py --workload train --steps 5 --warmup 3 >> $L 2>&1
done
python - <<'PY'
import json
for line in open('gpurun_out/wgrad_try.log'):
    if line.startswith('=='): print(line.strip()); continue
    if line.startswith('{'):
        d = json.loads(line); print(round(d['value']), round(d['ms_per_step'], 2))
        kb = d.get('kernel_breakdown', {})
        for k, v in sorted(kb.items(), key=lambda kv: -kv[1]['ms'])[:8]: print('   ', k, v['ms'], v.get('launches'), v.get('tflops'))
    elif 'Error' in line or 'error' in line or 'rc=' in line or 'passed' in line or 'failed' in line: print(line.strip()[:300])
PY
