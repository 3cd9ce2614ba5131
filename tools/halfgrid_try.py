"""Do persistent grids on HALF the SMs reach more than half the throughput?  If so, two streams side by side (each kernel on its own
74 SMs) overlap the HBM-bound layers of one half-batch with the tensor/epilogue-bound layers of the other.  Run under gpurun."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + '/tests')
from util import ar_frames
from nsc_b200 import codec, lpc_utilities as lu
dev = 'cuda'
cfg = codec.CodecConfig(resnet_type='bottleneck')
gcs = [codec.NeuralCodec(cfg, device=dev, seed=5), codec.NeuralCodec(cfg, device=dev, seed=6)]
cmA = codec.CMRL(gcs, res_scalar=1.0); cmB = codec.CMRL(gcs, res_scalar=1.0)
B = 16576
win = torch.from_numpy(np.tile(ar_frames(4144, 1024, seed=1), (4, 1))).to(dev); x = win[:, 256:768].contiguous()
lsf = lu.lpc_analysis_windows(win, 16, dtype=torch.float32)
def single():
    return cmA.feedforward_lpc(x, lsf, False, 1.0)
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
h = B // 2
xa, xb, la, lb = x[:h].contiguous(), x[h:].contiguous(), lsf[:h].contiguous(), lsf[h:].contiguous()
def dual():
    cur = torch.cuda.current_stream()
    sA.wait_stream(cur); sB.wait_stream(cur)
    with torch.cuda.stream(sA): ra = cmA.feedforward_lpc(xa, la, False, 1.0)
    with torch.cuda.stream(sB): rb = cmB.feedforward_lpc(xb, lb, False, 1.0)
    cur.wait_stream(sA); cur.wait_stream(sB)
    return ra, rb
def timeit(fn, n=4):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rep in range(2):
    t1 = timeit(single); t2 = timeit(dual)
    print(f'single stream {t1:.2f} ms ({B*0.03/t1*1e3:.0f}x)   two streams {t2:.2f} ms ({B*0.03/t2*1e3:.0f}x)')
r1 = single(); ra, rb = dual(); torch.cuda.synchronize()
print('identical', torch.equal(torch.cat([ra['synthesized'], rb['synthesized']]), r1['synthesized']))
'''
for env in ({}, {'NSC_SMS': '74'}, {'NSC_SMS': '100'}, {'NSC_SMS': '48'}):
    print('==', env, flush=True)
    r = subprocess.run([sys.executable, '-c', CHILD % {'root': ROOT}], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    if r.returncode:
        print('FAILED', r.stderr[-3000:])
