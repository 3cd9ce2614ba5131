#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plane.py tests/test_gpu_parity_depth.py tests/test_gpu_pipeline.py tests/test_gpu_training.py -x -q > gpurun_out/r02c_fold_pytest.log 2>&1
tail -12 gpurun_out/r02c_fold_pytest.log
for k in "NSC_PLANE_CHUNK=2072" "NSC_PLANE_CHUNK=1036" "NSC_PLANE_CHUNK=592" "NSC_PLANE_FOLD=0"; do
  echo "=== $k" >> gpurun_out/r02c_knobs.log
  env $k python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])
kb=d.get('kernel_breakdown',{})
print({k:round(v['ms'],2) for k,v in kb.items() if v['ms']>0.3})" >> gpurun_out/r02c_knobs.log 2>&1
done
cat gpurun_out/r02c_knobs.log
