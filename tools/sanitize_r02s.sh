#!/bin/bash
# compute-sanitizer memcheck over the kernels of round 2's last session: the folded narrow conv (fold-output epilogue of the taps-in-N
# and Toeplitz kernels, the 48 -> 48 k5 / k9 kernel with three issuing threads and the staged unfolding epilogue, paired tiles for
# dilation 2), the product-split stride-2 conv, the taps-in-N kernel as CTA pairs, programmatic dependent launch, the rewritten LPC
# analysis kernel (both batch shapes), the batch framing entry points and the batched corpus path.  Small batches: the sanitizer slows
# kernels 10-100x.   -> gpurun_out/r02s_sanitizer_memcheck.log
set -u
mkdir -p gpurun_out
cat > /tmp/san_child.py <<'PY'
import sys, os
root = os.environ.get('GRAFT_REPO_ROOT', '/root/repo')
sys.path.insert(0, root); sys.path.insert(0, root + '/tests')
import numpy as np, torch
from util import ar_frames
import test_gpu_parity as tp, test_gpu_block as tb, test_gpu_pipeline as tpl, test_gpu_plane as tpn
from nsc_b200 import nn_core_operator as nn, lpc_utilities as lu
from oracle import ref_nn
for (L, wide, dil) in ((512, 100, 1), (512, 100, 2), (256, 100, 1), (256, 100, 2), (512, 50, 2)):
    for B in (2, 5):
        ps = ref_nn.ParamStream(seed=wide + dil + B)
        x = np.random.RandomState(B).randn(B, L, wide).astype(np.float32)
        ref = ref_nn.the_bottleneck(torch.from_numpy(x), wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, ps=ps).numpy()
        params = [tuple(tb.cu(p) for p in t) for t in ps.params]
        got = nn.the_bottleneck(tb.cu(x), wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, params=params, fused='folded')
        torch.cuda.synchronize()
        assert nn.last_engine == 'tc_folded' and tp.rel_err(got.cpu().numpy(), ref) < 5e-5
    print('folded block', L, wide, dil, 'ok')
for st in ((2,), (2, 2)):
    for B in (3, 4):
        oc, gc = tp._make_pair('bottleneck', st, seed=3, precision='tc_f16x3')
        tp._check_codec(oc, gc, ar_frames(B, 512, seed=31, std=0.3), False)
    print('codec (folded narrow convs, split stride-2 conv, dependent launches)', st, 'ok')
tp.test_cq_feedforward_vs_oracle_and_golden()
print('cq ok')
win = tp.cu(ar_frames(300, 1024, seed=5))
a = lu.lpc_analysis_windows(win, 16)
b = torch.cat([lu.lpc_analysis_windows(win[i:i + 37], 16) for i in range(0, 300, 37)])
assert torch.equal(a, b)
print('lpc analysis ok')
tpl.test_batch_framing_entry_points_match_the_single_utterance_ones()
tpl.test_batched_corpus_path_is_bit_identical_to_the_per_utterance_path()
print('batch framing / corpus ok')
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san_child.py > gpurun_out/r02s_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02s_sanitizer_memcheck.log
tail -12 gpurun_out/r02s_sanitizer_memcheck.log
cat > /tmp/san_child2.py <<'PY'
import sys, os
root = os.environ.get('GRAFT_REPO_ROOT', '/root/repo')
sys.path.insert(0, root); sys.path.insert(0, root + '/tests')
import test_gpu_plane as t
for layer in (t.T_LAYERS[0], t.T_LAYERS[1]):
    for B in (2, 6):
        e = t._run(B=B, precision=1, seed=B, **layer)
        assert e < 2e-5, (layer, B, e)
print('taps-in-N CTA pairs ok')
PY
NSC_PLANE_PAIR=2 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san_child2.py >> gpurun_out/r02s_sanitizer_memcheck.log 2>&1
echo "memcheck (NSC_PLANE_PAIR=2) rc=$?" >> gpurun_out/r02s_sanitizer_memcheck.log
tail -4 gpurun_out/r02s_sanitizer_memcheck.log
