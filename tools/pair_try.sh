#!/bin/bash
# CTA-pair (cta_group::2) mode: probe of the leader's waits, parity of the pair layers, then a bench line per setting.
set -u
mkdir -p gpurun_out
L=gpurun_out/pair_try.log
: > $L
timeout 300 python tools/pair_probe.py >> $L 2>&1
timeout 600 python -m pytest tests/test_gpu_plane.py -x -q -k "cta_pair or narrow_layers_on or tap_shift or multicast" >> $L 2>&1
echo "pytest rc=$?" >> $L
for cfg in "NSC_PLANE_PAIR=0" "NSC_PLANE_PAIR=1"; do
  echo "== $cfg" >> $L
  env $cfg timeout 300 python bench.py --frames 33152 --steps 3 --warmup 3 --no-cpu-baseline >> $L 2>&1
done
python - <<'PY'
import json
for line in open('gpurun_out/pair_try.log'):
    if line.startswith('=='): print(line.strip()); continue
    if line.startswith('{'):
        d = json.loads(line); print(round(d['value']), round(d['ms_per_step'], 2), round(d['e2e']['value']))
        kb = d.get('kernel_breakdown', {})
        for k, v in sorted(kb.items(), key=lambda kv: -kv[1]['ms'])[:14]: print('   ', k, v['ms'], v.get('launches'))
    elif 'per-tile' in line or 'passed' in line or 'failed' in line or 'rc=' in line or 'Error' in line or 'error' in line: print(line.strip()[:300])
PY
