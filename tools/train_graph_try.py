"""Training step: eager launches vs the forward + backward of CQTrainer.loss_and_grads captured ONCE in a CUDA graph (Adam stays
outside: its bias correction takes the step count as a host scalar)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from nsc_b200 import codec, lpc_utilities as lu
from nsc_b200.training import CQTrainer
dev = 'cuda:0'
B = 128
for rt in ('bottleneck', 'gln'):
    cfg = codec.CodecConfig(resnet_type=rt)
    cm = codec.CMRL([codec.NeuralCodec(cfg, device=dev, seed=5), codec.NeuralCodec(cfg, device=dev, seed=6)], res_scalar=1.0)
    tr = CQTrainer.finetuning_lpc(cm, (60.0, 10.0, 10.0, 0.0), lr=2e-6)
    x_np, win_np = bench.synth_audio(B, seed=4321)
    x = torch.from_numpy(x_np).to(dev) * 0.3
    lsf = lu.lpc_analysis_windows(torch.from_numpy(win_np).to(dev), 16, dtype=torch.float32)

    def timeit(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    t_eager = timeit(lambda: tr.step(x, lsf))
    t_fb = timeit(lambda: tr.loss_and_grads(x, lsf))
    # capture
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): tr.loss_and_grads(x, lsf)
    torch.cuda.current_stream().wait_stream(s)
    ref = tr.loss_and_grads(x, lsf)
    gref = [t.clone() for t in tr.grads]
    try:
        with torch.cuda.graph(g):
            out = tr.loss_and_grads(x, lsf)
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        same = all(torch.equal(a, b) for a, b in zip(gref, tr.grads)) and torch.equal(ref['loss_vector'], out['loss_vector'])
        t_graph = timeit(lambda: g.replay())
        t_graph_adam = timeit(lambda: (g.replay(), tr.apply_adam()))
        print(f"{rt}: eager step {t_eager:.2f} ms (fwd+bwd {t_fb:.2f}); graph fwd+bwd {t_graph:.2f} ms, + eager adam {t_graph_adam:.2f} ms; identical={same}")
    except Exception as ex:
        print(rt, 'capture failed:', repr(ex)[:400])
