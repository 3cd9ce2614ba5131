#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plane.py tests/test_gpu_block.py -x -q -m gpu -k "taps_in_n or fused or folded or plane_path or pair" > gpurun_out/r02z_plane.log 2>&1
tail -2 gpurun_out/r02z_plane.log
for pair in 1 0; do
NSC_PLANE_PAIR=$pair timeout 400 python bench.py --steps 5 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r02z_bench_$pair.json 2> gpurun_out/r02z_bench.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02z_bench_$pair.json'))
print('pair=$pair', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])
for k, v in list(d['kernel_breakdown'].items())[:8]:
    print('   ', k, v['ms'], v['launches'])
PY
done
