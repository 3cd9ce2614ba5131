#!/bin/bash
# round-2 session 3, call 1: batched corpus path, LPC front end hoisted, new lpc_analyze kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_framing.py -x -q -m gpu -k "lpc or cq or quantizer or pipeline or utterance or framing or batch or code or silence or smoke" > gpurun_out/r02n_pytest.log 2>&1
tail -5 gpurun_out/r02n_pytest.log
timeout 300 python bench.py --workload corpus --steps 2 --warmup 1 > gpurun_out/r02n_corpus.json 2> gpurun_out/r02n_corpus.err
tail -c 600 gpurun_out/r02n_corpus.json
timeout 300 python tools/corpus_phases.py > gpurun_out/r02n_phases.log 2>&1
tail -3 gpurun_out/r02n_phases.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02n_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k, v in d['kernel_breakdown'].items():
    if not k.startswith('p'): print(k, v)
PY
