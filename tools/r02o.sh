#!/bin/bash
# round-2 session 3: folded narrow conv (48 -> 48 k5 on pairs of positions)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_block.py -x -q -m gpu -k "folded" > gpurun_out/r02o_block.log 2>&1
tail -15 gpurun_out/r02o_block.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plane.py -x -q -m gpu -k "codec or cascade or cq or plane_path or folded" > gpurun_out/r02o_codec.log 2>&1
tail -8 gpurun_out/r02o_codec.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02o_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k, v in d['kernel_breakdown'].items():
    print(k, v['ms'], v['launches'], v.get('gbs'))
PY
