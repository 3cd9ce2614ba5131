#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plane.py tests/test_gpu_block.py -x -q -m gpu -k "tap_shift or pair or staged or fused or folded or plane_path" > gpurun_out/r02x_plane.log 2>&1
tail -2 gpurun_out/r02x_plane.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02x_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k, v in list(d['kernel_breakdown'].items())[:12]:
    print(k, v['ms'], v['launches'])
PY
