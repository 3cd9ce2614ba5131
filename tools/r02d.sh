#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plane.py tests/test_gpu_parity.py -x -q -k "prepared or folded or cascade or cq_ or codec_" > gpurun_out/r02d_pytest.log 2>&1
tail -12 gpurun_out/r02d_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
python - <<'PY'
import json
for line in open('gpurun_out/r02d_bench.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])
        print({k:round(v['ms'],2) for k,v in d['kernel_breakdown'].items() if v['ms']>0.3})
        for k,v in d['sub_records'].items():
            print(k, json.dumps(v)[:700])
PY
