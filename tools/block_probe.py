"""One small fused-block launch (for compute-sanitizer / debugging): python tools/block_probe.py [B] [L] [wide] [dil]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from nsc_b200 import nn_core_operator as nn
from oracle import ref_nn
B, L, wide, dil = [int(v) for v in (sys.argv[1:5] + ['2', '512', '100', '1'][len(sys.argv) - 1:])]
ps = ref_nn.ParamStream(seed=1)
x = np.random.RandomState(8).randn(B, L, wide).astype(np.float32)
ref = ref_nn.the_bottleneck(torch.from_numpy(x[:2]), wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, ps=ps).numpy()
params = [tuple(torch.from_numpy(p).cuda() for p in t) for t in ps.params]
got = nn.the_bottleneck(torch.from_numpy(x).cuda(), wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, params=params)
torch.cuda.synchronize()
print(nn.last_engine, float(np.abs(got[:2].cpu().numpy() - ref).max() / np.abs(ref).max()))
