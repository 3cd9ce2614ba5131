"""Where a corpus step goes (bench.py --workload corpus): per-phase device time of nsc_b200.pipeline.code_utterances, 360 x 10 s."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from nsc_b200 import codec, pipeline, bitstream, _lib
dev = 'cuda:0'
cfg = codec.CodecConfig(resnet_type='bottleneck')
cm = codec.CMRL([codec.NeuralCodec(cfg, device=dev, seed=5), codec.NeuralCodec(cfg, device=dev, seed=6)], res_scalar=1.0)
n_utt, T = 360, 160000
x_np, _ = bench.synth_audio(64, seed=4321)
base = np.tile(x_np.reshape(-1), -(-T * 8 // x_np.size))
host = [torch.from_numpy(np.ascontiguousarray(base[(i % 8) * 4000:(i % 8) * 4000 + T])).pin_memory() for i in range(n_utt)]
lib = _lib.load()

def phase(name, fn, acc):
    torch.cuda.synchronize(); t0 = time.perf_counter(); l0 = lib.nsc_launch_count()
    r = fn()
    torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    acc[name + '_launches'] = lib.nsc_launch_count() - l0
    return r

for rep in range(3):
    acc = {}
    sigs = phase('h2d', lambda: [h.to(dev, non_blocking=True) for h in host], acc)
    ana = phase('analysis', lambda: pipeline.analysis_batch(torch.stack(sigs)), acc)
    frames, lsf = ana['frames'].reshape(-1, 512), ana['lsf'].reshape(-1, 16)
    r = phase('cq_forward', lambda: cm.feedforward_lpc(frames, lsf, False, 1.0), acc)
    rec = phase('pack', lambda: bitstream.pack_frames(r['lsf_idx'], r['idx'], [32, 32], cm.n_lsf_bins), acc)
    n = ana['n_used']
    syn = phase('synthesis', lambda: [pipeline.synthesis_batch(r['synthesized'].reshape(n_utt, n, 512), ana['n_seg'], ana['n_seg2'], ana['std']),
                                      pipeline.synthesis_batch(r['decoded'].reshape(n_utt, n, 512), ana['n_seg'], ana['n_seg2'], 1.0, de_emphasis=False)], acc)
    phase('d2h', lambda: [syn[0].cpu(), rec.cpu()], acc)
    print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in acc.items()}, 'frames', frames.shape[0])
