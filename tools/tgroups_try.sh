#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/tgroups_try.log
: > $L
timeout 400 python -m pytest tests/test_gpu_plane.py -x -q -k "taps_in_n or full_size or plane_path" >> $L 2>&1
echo "pytest rc=$?" >> $L
for cfg in "NSC_PLANE_TGROUPS=5" "NSC_PLANE_TGROUPS=3"; do
  echo "== $cfg" >> $L
  env $cfg timeout 200 python bench.py --frames 33152 --steps 3 --warmup 3 --no-cpu-baseline >> $L 2>&1
done
python - <<'PY'
import json
for line in open('gpurun_out/tgroups_try.log'):
    if line.startswith('=='): print(line.strip()); continue
    if line.startswith('{'):
        d = json.loads(line); print(round(d['value']), round(d['ms_per_step'], 2), round(d['e2e']['value']))
        kb = d.get('kernel_breakdown', {})
        for k, v in sorted(kb.items(), key=lambda kv: -kv[1]['ms'])[:9]: print('   ', k, v['ms'], v.get('launches'))
    elif 'passed' in line or 'failed' in line or 'rc=' in line or 'Error' in line or 'error' in line or 'assert' in line: print(line.strip()[:300])
PY
