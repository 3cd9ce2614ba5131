"""Generates tests/golden/*.npz with the ORACLE (oracle/ is the only thing executed here).

The reference ships no golden vectors and cannot run in this environment, so these
fixtures pin the oracle's outputs on seeded inputs: a later change to oracle/ or to the CUDA path that moves a
number shows up as a diff against a committed file.  Weights are NOT stored -- they are regenerated from the
seed (numpy RandomState streams are stable) -- only inputs that are cheap and all outputs.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import ref_codec, ref_loss, ref_lpc, ref_nn  # noqa: E402
from util import ar_frames, quantizer_edge_codes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
LSF_BINS = np.load(os.path.join(HERE, 'lsf_bins_f64.npy')).astype(np.float32)


def golden_quantizer():
    bins32 = np.linspace(-1, 1, 32).astype(np.float32)
    rng = np.random.RandomState(7)
    x32 = np.concatenate([quantizer_edge_codes(bins32), rng.uniform(-1.1, 1.1, 512).astype(np.float32)])
    x32 = x32[: (len(x32) // 64) * 64].reshape(-1, 64, 1)
    soft, code = ref_nn.scalar_softmax_quantization(torch.from_numpy(x32), -300.0, bins32, 1.0, False, 64, 32)
    idx = ref_nn.quantizer_indices(torch.from_numpy(x32), -300.0, bins32)
    xl = np.concatenate([quantizer_edge_codes(LSF_BINS), rng.uniform(0, np.pi, 256).astype(np.float32)])
    xl = xl[: (len(xl) // 16) * 16].reshape(-1, 16, 1)
    idx_l = ref_nn.quantizer_indices(torch.from_numpy(xl), -300.0, LSF_BINS)
    _, code_l = ref_nn.scalar_softmax_quantization(torch.from_numpy(xl), -300.0, LSF_BINS, 1.0, False, 16, 256)
    np.savez_compressed(os.path.join(HERE, 'quantizer.npz'), x32=x32, idx32=idx.numpy().astype(np.uint8),
                        code32=code.numpy(), xl=xl, idx_l=idx_l.numpy().astype(np.uint8), code_l=code_l.numpy())


def golden_lpc():
    x = ar_frames(6, 1024, seed=11)
    lsf = ref_lpc.lpc_analysis_windows(x, 16)
    frames = ar_frames(6, 512, seed=12)
    lsf32 = lsf.astype(np.float32)
    poly = ref_lpc.lsf2poly_after_quan(lsf32, 16)
    res = ref_lpc.lpc_analysis_get_residual(frames[:, :, None], poly)
    syn = ref_lpc.lpc_synthesizer_tr(poly, res)
    lsf_tr = ref_lpc.lpc_analysis_at_train(frames[:, :, None], 16)
    np.savez_compressed(os.path.join(HERE, 'lpc.npz'), windows=x, lsf=lsf, frames=frames, poly=poly, res=res, syn=syn,
                        lsf_train=lsf_tr)


def golden_losses():
    a = ar_frames(5, 512, seed=21, std=0.3)
    b = (a + 0.05 * np.random.RandomState(22).randn(*a.shape)).astype(np.float32)
    t = ref_loss.mse_loss(torch.from_numpy(b), torch.from_numpy(a)).numpy()
    f = ref_loss.mfcc_loss(torch.from_numpy(b), torch.from_numpy(a)).numpy()
    np.savez_compressed(os.path.join(HERE, 'losses.npz'), ori=a, dec=b, time_loss=t, freq_loss=f)


def golden_codec():
    out = {}
    for name, rt, st in [('bn2', 'bottleneck', (2,)), ('gln2', 'gln', (2,)), ('bn4', 'bottleneck', (2, 2))]:
        oc = ref_codec.OracleCodec(ref_codec.OracleCodecCfg(resnet_type=rt, strides=st), seed=3)
        x = ar_frames(2, 512, seed=31, std=0.3)
        r = oc.forward(torch.from_numpy(x)[:, :, None], False, 1.0)
        out[name + '_x'] = x
        out[name + '_floating'] = r['floating_code'].numpy()[:, :, 0]
        out[name + '_code'] = r['code'].numpy()[:, :, 0]
        out[name + '_out'] = r['out'].numpy()
    np.savez_compressed(os.path.join(HERE, 'codec.npz'), **out)


def golden_cq():
    cfg = ref_codec.OracleCodecCfg()
    codecs = [ref_codec.OracleCodec(cfg, seed=5), ref_codec.OracleCodec(cfg, seed=6)]
    x = ar_frames(3, 512, seed=41)
    win = ar_frames(3, 1024, seed=42)
    lsf = ref_lpc.lpc_analysis_windows(win, 16).astype(np.float32)
    r = ref_codec.cq_feedforward(codecs, -300.0, LSF_BINS, torch.from_numpy(x)[:, :, None],
                                 torch.from_numpy(lsf)[:, :, None], False, 1.0, res_scalar=1.0)
    np.savez_compressed(os.path.join(HERE, 'cq.npz'), x=x, lsf=lsf, poly=r['poly'], res_x=r['res_x'].numpy()[:, :, 0],
                        decoded=r['decoded'].numpy(), synthesized=r['synthesized'],
                        floating0=r['per'][0]['floating_code'].numpy()[:, :, 0],
                        floating1=r['per'][1]['floating_code'].numpy()[:, :, 0],
                        time_loss=r['time_loss'].numpy(), freq_loss=r['freq_loss'].numpy(),
                        ent=np.array([float(e) for e in r['ent']]), ent_lpc=float(r['ent_lpc']))


if __name__ == '__main__':
    torch.manual_seed(0)
    torch.set_num_threads(1)
    golden_quantizer()
    golden_lpc()
    golden_losses()
    golden_codec()
    golden_cq()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
