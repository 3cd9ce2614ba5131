#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -x -q > gpurun_out/r02j_pytest.log 2>&1
tail -25 gpurun_out/r02j_pytest.log | cut -c1-300
for k in "NSC_WGRAD_TC=1" "NSC_WGRAD_TC_DIRECT=1"; do
  echo "=== $k"
  env $k timeout 300 python bench.py --workload train --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['launches_per_step'])
print({k:(round(v['ms'],3), v['launches']) for k,v in d['kernel_breakdown'].items() if v['ms']>0.25})"
done
