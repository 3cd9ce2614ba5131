#!/bin/bash
# compute-sanitizer memcheck over the kernels this round added or changed: the fused gate-pair layer (k15, gate product in the
# epilogue; plain and de-interleaved / interleaving form), the depthwise half of the separable up-conv, K = 1 tap-shift layers, the
# head epilogues with the folded quantiser / cascade accumulation, two stride-2 stages (128-position levels, 25-channel images), the
# fused block kernel with the deferred frame publish, and the 'gln' training backward.  Small batches: the sanitizer slows kernels
# 10-100x.   -> gpurun_out/r02_sanitizer_memcheck.log
set -u
mkdir -p gpurun_out
cat > /tmp/san_child.py <<'PY'
import sys, os
root = os.environ.get('GRAFT_REPO_ROOT', '/root/repo')
sys.path.insert(0, root); sys.path.insert(0, root + '/tests')
import numpy as np, torch
from nsc_b200 import codec
from util import ar_frames
import test_gpu_parity as tp, test_gpu_block as tb, test_gpu_training as tt
for rt, st in (('gln', (2,)), ('bottleneck', (2, 2))):
    for B in (3, 4):
        oc, gc = tp._make_pair(rt, st, seed=3, precision='tc_f16x3')
        tp._check_codec(oc, gc, ar_frames(B, 512, seed=31, std=0.3), False)
        print('codec', rt, st, B, 'ok')
tp.test_cascade_vs_oracle()
tp.test_cq_feedforward_vs_oracle_and_golden()
print('cascade / cq with folded heads ok')
tb.test_fused_block_vs_oracle(2, 512, 100, 1, False)
tb.test_fused_block_equals_three_launches(127, 512, 100, 2)
print('fused block ok')
tt.test_backward_matches_autograd_gln('fp32')
print('gln training backward ok')
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san_child.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -8 gpurun_out/r02_sanitizer_memcheck.log
