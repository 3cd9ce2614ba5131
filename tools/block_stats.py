"""Per-role counters of the fused block kernel (run with NSC_BLOCK_STATS=1): python tools/block_stats.py [B] [L] [wide] [dil]
Prints, per role, CTAs, work units, mean epilogue-loop time, time waited for a free ring slot / a ready frame / own stores."""
import ctypes as C, os, sys
os.environ.setdefault('NSC_BLOCK_STATS', '1')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from nsc_b200 import nn_core_operator as nn, _lib
from oracle import ref_nn
B, L, wide, dil = [int(v) for v in (sys.argv[1:5] + ['2072', '512', '100', '1'][len(sys.argv) - 1:])]
ps = ref_nn.ParamStream(seed=1)
ref_nn.the_bottleneck(torch.zeros(1, 128, wide), wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, ps=ps)
params = [tuple(torch.from_numpy(p).cuda() for p in t) for t in ps.params]
x = torch.randn(B, L, wide, device='cuda')
lib = _lib.load()
for rep in range(3):
    lib.nsc_profile_begin(64)
    y = nn.the_bottleneck(x, wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, params=params)
    n = C.c_int32(0); names = C.create_string_buffer(64 * 32); ms = (C.c_float * 64)(); fl = (C.c_double * 64)(); by = (C.c_double * 64)()
    lib.nsc_profile_end(C.byref(n), names, ms, fl, by, 64)
    t = {names.raw[i * 32:(i + 1) * 32].split(b'\0')[0].decode(): ms[i] for i in range(n.value)}
buf = (C.c_ulonglong * (160 * 8))()
nc = lib.nsc_debug_block_stats(buf, 160)
a = np.array(buf[:], dtype=np.float64).reshape(160, 8)
clk = torch.cuda.clock_rate() * 1e3 if hasattr(torch.cuda, 'clock_rate') else 1.9e9
print(nn.last_engine, {k: round(v, 3) for k, v in t.items() if k.startswith('pB')}, 'split', os.environ.get('NSC_BLOCK_SPLIT'), 'ring', os.environ.get('NSC_BLOCK_RING'))
for role in (1, 2, 3):
    r = a[a[:, 0] == role]
    if len(r) == 0: continue
    us = lambda v: round(float(v) / 1.9e3, 1)     # cycles -> us at ~1.9 GHz
    print(f"role {role}: ctas {len(r)} units/cta {r[:,4].mean():.1f} loop {us(r[:,1].mean())} us (max {us(r[:,1].max())}) "
          f"wait_slot {us(r[:,2].mean())} wait_frame {us(r[:,3].mean())} wait_store {us(r[:,5].mean())} "
          f"issuer: wait_input {us(r[:,6].mean())} wait_acc {us(r[:,7].mean())}")
