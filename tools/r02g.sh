#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_depth.py tests/test_gpu_plane.py -x -q -k "gln or codec_forward or plane_path_matches or golden or reference_run" > gpurun_out/r02g_pytest.log 2>&1
tail -8 gpurun_out/r02g_pytest.log
timeout 300 python - > gpurun_out/r02g_bench.log 2>&1 <<'PY'
import sys, json, argparse
sys.path.insert(0, '.')
import torch, bench
from nsc_b200 import _lib
lib = _lib.load()
args = argparse.Namespace(precision='tc_f16x3')
for rt, st, fr in (('gln', (2, 2), 4144), ('gln', (2,), 4144)):
    r = bench.measure_variant(args, 1, 0, 'cuda:0', lib, rt, st, frames=fr)
    print(json.dumps(r))
PY
cat gpurun_out/r02g_bench.log | cut -c1-1000
