#!/bin/bash
# Frames-per-pass sweep of the plane path (NSC_PLANE_CHUNK): per-launch fill/drain overhead vs workspace size.
#   tools/chunk_sweep.sh  ->  gpurun_out/chunk_sweep.log  (one bench line per chunk size)
set -u
mkdir -p gpurun_out
: > gpurun_out/chunk_sweep.log
for c in 2072 4144 8288 16576; do
  echo "== NSC_PLANE_CHUNK=$c" >> gpurun_out/chunk_sweep.log
  NSC_PLANE_CHUNK=$c timeout 300 python bench.py --frames 33152 --steps 3 --warmup 3 --no-cpu-baseline >> gpurun_out/chunk_sweep.log 2>&1
done
python - <<'PY'
import json
for line in open('gpurun_out/chunk_sweep.log'):
    if line.startswith('=='): print(line.strip()); continue
    if line.startswith('{'):
        d = json.loads(line); print(round(d['value']), d['ms_per_step'], d['e2e']['value'])
PY
