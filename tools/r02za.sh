#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plane.py -x -q -m gpu -k "tap_shift or plane_path or pair_knob" > gpurun_out/r02za_plane.log 2>&1
tail -2 gpurun_out/r02za_plane.log
for ps in 1 0; do
NSC_PLANE_PSPLIT=$ps timeout 400 python bench.py --steps 5 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r02za_bench_$ps.json 2> gpurun_out/r02za_bench.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02za_bench_$ps.json'))
print('psplit=$ps', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])
for k, v in list(d['kernel_breakdown'].items())[:6]:
    print('   ', k, v['ms'], v['launches'])
PY
done
