#!/bin/bash
# session baseline: GPU tests, default bench line with sub-records, two knob probes
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02b_pytest.log
tail -3 gpurun_out/r02b_pytest.log
python bench.py > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err
tail -c 600 gpurun_out/r02b_bench_default.json
for k in "NSC_PLANE_NARROW=X" "NSC_PLANE_CHUNK=4144"; do
  echo "=== $k" >> gpurun_out/r02b_knobs.log
  env $k python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
kb=d.get('kernel_breakdown',{})
print({k:round(v['ms'],2) for k,v in kb.items() if v['ms']>1})" >> gpurun_out/r02b_knobs.log 2>&1
done
cat gpurun_out/r02b_knobs.log
