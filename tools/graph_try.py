import sys, torch, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from util import ar_frames
from nsc_b200 import codec, lpc_utilities as lu
dev='cuda'
cfg=codec.CodecConfig(resnet_type='bottleneck')
cm=codec.CMRL([codec.NeuralCodec(cfg,device=dev,seed=5),codec.NeuralCodec(cfg,device=dev,seed=6)],res_scalar=1.0)
for B in (1,128,1024):
    win=torch.from_numpy(ar_frames(B,1024,seed=1)).to(dev); x=win[:,256:768].contiguous()
    def step():
        lsf=lu.lpc_analysis_windows(win,16,dtype=torch.float32)
        return cm.feedforward_lpc(x,lsf,False,1.0)
    for _ in range(3): r0=step()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): step()
    e1.record(); torch.cuda.synchronize()
    t_plain=e0.elapsed_time(e1)/20
    g=torch.cuda.CUDAGraph()
    s=torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): step()
    torch.cuda.current_stream().wait_stream(s)
    try:
        with torch.cuda.graph(g):
            r=step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20): g.replay()
        e1.record(); torch.cuda.synchronize()
        t_graph=e0.elapsed_time(e1)/20
        ok=torch.equal(r['synthesized'],r0['synthesized']) and all(torch.equal(a,b) for a,b in zip(r['idx'],r0['idx']))
        print(f"B={B}: plain {t_plain*1e3:.0f} us, graph {t_graph*1e3:.0f} us, identical={ok}, xRT plain {B*0.03/(t_plain*1e-3):.0f} graph {B*0.03/(t_graph*1e-3):.0f}")
    except Exception as ex:
        print('graph capture failed:', repr(ex)[:300])
