#!/bin/bash
# fused block kernel after the deferred frame publish: role split sweep (per-role counters) for the three block shapes
out=gpurun_out/${1:-r02b_block_sweep}.log
: > $out
run() { echo "=== $*" >> $out; env "$@" timeout 120 python tools/block_stats.py $SHAPE >> $out 2>&1; }
SHAPE="2072 512 100 1"
echo "##### 100-channel block, L 512" >> $out
for s in 50,36 44,32 40,30 36,28 32,26 40,36 46,40; do run NSC_BLOCK_SPLIT=$s NSC_BLOCK_RING=128; done
run NSC_BLOCK_SPLIT=40,30 NSC_BLOCK_RING=256
run NSC_BLOCK_SPLIT=40,30 NSC_BLOCK_RING=64
SHAPE="2072 256 100 2"
echo "##### 100-channel block, L 256" >> $out
for s in 50,36 40,30 36,28 44,36; do run NSC_BLOCK_SPLIT=$s NSC_BLOCK_RING=128; done
SHAPE="2072 512 50 2"
echo "##### 50-channel block, L 512" >> $out
for s in 52,44 44,40 40,36 48,48 36,36; do run NSC_BLOCK_SPLIT=$s NSC_BLOCK_RING=128; done
echo "##### unfused reference" >> $out
for sh in "2072 512 100 1" "2072 256 100 2" "2072 512 50 2"; do
python - $sh >> $out 2>&1 <<'PY'
import ctypes as C, os, sys
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from nsc_b200 import nn_core_operator as nn, _lib
from oracle import ref_nn
B, L, wide, dil = [int(v) for v in sys.argv[1:5]]
ps = ref_nn.ParamStream(seed=1)
ref_nn.the_bottleneck(torch.zeros(1, 128, wide), wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, ps=ps)
params = [tuple(torch.from_numpy(p).cuda() for p in t) for t in ps.params]
x = torch.randn(B, L, wide, device='cuda')
lib = _lib.load()
for rep in range(3):
    lib.nsc_profile_begin(64)
    y = nn.the_bottleneck(x, wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, params=params, fused=False)
    n = C.c_int32(0); names = C.create_string_buffer(64 * 32); ms = (C.c_float * 64)(); fl = (C.c_double * 64)(); by = (C.c_double * 64)()
    lib.nsc_profile_end(C.byref(n), names, ms, fl, by, 64)
    t = {names.raw[i * 32:(i + 1) * 32].split(b'\0')[0].decode(): ms[i] for i in range(n.value)}
print('unfused', sys.argv[1:5], {k: round(v, 3) for k, v in t.items() if k.startswith('p') and not k.startswith('plane')}, 'sum', round(sum(v for k, v in t.items() if k.startswith('p') and not k.startswith('plane')), 3))
PY
done
cat $out
