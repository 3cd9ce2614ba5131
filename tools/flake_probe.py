"""Which side of test_fused_block_equals_three_launches[700-512-100-1] is unstable?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import test_gpu_block as tb
from nsc_b200 import nn_core_operator as nn
B, L, wide, dil = 700, 512, 100, 1
ps = tb._params(wide, dil, seed=3)
params = [tuple(tb.cu(p) for p in t) for t in ps.params]
x = tb.cu(np.random.RandomState(B).randn(B, L, wide).astype(np.float32))
ref = None
stats = {'fused_unstable': 0, 'plain_unstable': 0, 'fused_ne_plain': 0}
outs = {'fused': [], 'plain': []}
for rep in range(12):
    a = nn.the_bottleneck(x, wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, params=params)
    b = nn.the_bottleneck(x, wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, params=params, fused=False)
    torch.cuda.synchronize()
    outs['fused'].append(a.clone()); outs['plain'].append(b.clone())
f0, p0 = outs['fused'][0], outs['plain'][0]
for k in range(12):
    fa, pb = outs['fused'][k], outs['plain'][k]
    df = (fa != f0); dp = (pb != p0); dx = (fa != pb)
    print(k, 'fused!=fused0', int(df.sum()), 'plain!=plain0', int(dp.sum()), 'fused!=plain', int(dx.sum()),
          'frames', sorted(set(torch.nonzero(dx.reshape(B, -1).any(1)).flatten().tolist()))[:8],
          'max|d|', float((fa - pb).abs().max()))
