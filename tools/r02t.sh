#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plane.py tests/test_gpu_block.py -x -q -m gpu > gpurun_out/r02t_plane.log 2>&1
tail -3 gpurun_out/r02t_plane.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "codec or cascade or cq or graph or prepared" > gpurun_out/r02t_codec.log 2>&1
tail -3 gpurun_out/r02t_codec.log
for pdl in 1 0; do
NSC_PLANE_PDL=$pdl timeout 400 python bench.py --steps 5 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r02t_bench_pdl$pdl.json 2> gpurun_out/r02t_bench.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02t_bench_pdl$pdl.json'))
print('pdl=$pdl', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
PY
done
