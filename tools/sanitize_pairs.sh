#!/bin/bash
# compute-sanitizer memcheck over the kernels this round changed: the paired up-sampling conv (cta_group::2), the taps-in-N kernel
# with two tiles' worth of spill slots, and the training backward (pipelined weight gradient, head correlation, sub-pixel data
# gradient).  Small batches: the sanitizer slows kernels by 10-100x.   -> gpurun_out/sanitizer_memcheck_r01w.log
set -u
mkdir -p gpurun_out
cat > /tmp/san_child.py <<'PY'
import sys, os
root = os.environ.get('GRAFT_REPO_ROOT', '/root/repo')
sys.path.insert(0, root); sys.path.insert(0, root + '/tests')
import test_gpu_plane as t
for layer, Bs in ((t.X_LAYERS[5], (2, 150)), (t.X_LAYERS[4], (3,)), (t.T_LAYERS[0], (5, 149)), (t.T_LAYERS[4], (5, 149)), (t.X_LAYERS[0], (3,))):
    for B in Bs:
        e = t._run(B=B, precision=1, seed=B, **layer)
        print(layer, B, e)
        assert e < 2e-5
import test_gpu_training as tt
tt.test_backward_matches_autograd_two_codecs(-20.0, 1.0, 'tc_f16x3')
print('training backward ok')
PY
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san_child.py > gpurun_out/sanitizer_memcheck_r01w.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_r01w.log
tail -5 gpurun_out/sanitizer_memcheck_r01w.log
