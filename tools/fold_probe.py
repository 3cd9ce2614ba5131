"""One bottleneck block with the folded narrow conv (operator surface, fused='folded') at a full pass: for ncu captures of the
48 -> 48 k5 kernel (plane_x_kernel) and event timings of the three launches."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from oracle import ref_nn
from nsc_b200 import nn_core_operator as nn

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2072
dev = 'cuda:0'
for (L, wide, dil) in ((512, 100, 1), (512, 100, 2), (256, 100, 1)):
    ps = ref_nn.ParamStream(seed=1)
    ref_nn.the_bottleneck(torch.zeros(1, 128, wide), wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, ps=ps)
    params = [tuple(torch.from_numpy(np.ascontiguousarray(p)).to(dev) for p in t) for t in ps.params]
    x = torch.randn(B, L, wide, device=dev)
    for mode in ('folded', False):
        for _ in range(2):
            y = nn.the_bottleneck(x, wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, params=params, fused=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = nn.the_bottleneck(x, wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, params=params, fused=mode)
        e1.record(); torch.cuda.synchronize()
        print(L, wide, dil, mode, 'ms per block call (incl. fp32 edges):', e0.elapsed_time(e1) / 5)

if os.environ.get('NSC_FOLD_STATS'):
    import ctypes as C
    from nsc_b200 import _lib
    lib = _lib.load()
    ps = ref_nn.ParamStream(seed=1)
    ref_nn.the_bottleneck(torch.zeros(1, 128, 100), wide_layer=100, narrow_layer=20, dilation_rate=1, is_last_flat=False, ps=ps)
    params = [tuple(torch.from_numpy(np.ascontiguousarray(p)).to(dev) for p in t) for t in ps.params]
    x = torch.randn(B, 512, 100, device=dev)
    y = nn.the_bottleneck(x, wide_layer=100, narrow_layer=20, dilation_rate=1, is_last_flat=False, params=params, fused='folded')
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (160 * 8))()
    lib.nsc_debug_block_stats.argtypes = [C.POINTER(C.c_ulonglong), C.c_int32]
    n = lib.nsc_debug_block_stats(buf, 160)
    a = np.array(list(buf)).reshape(160, 8)[:148]
    print('per CTA (mean over CTAs with tiles): tiles', a[:, 4][a[:, 4] > 0].mean(), 'issue-section cycles/tile', (a[:, 5] / np.maximum(a[:, 4], 1))[a[:, 4] > 0].mean(),
          'wait a_full/tile', (a[:, 6] / np.maximum(a[:, 4], 1))[a[:, 4] > 0].mean(), 'wait acc_empty/tile', (a[:, 7] / np.maximum(a[:, 4], 1))[a[:, 4] > 0].mean())
