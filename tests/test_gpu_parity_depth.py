"""Parity in depth (round-2 review items): per-FRAME errors over an 80 dB range of input levels, a direct oracle comparison of the
headline workload (cq2) at more than one engine pass (one pass + 300 frames), the end-to-end hard-code agreement rate against the
float32 oracle AND against float64 truth, and gated_bottleneck_decoder."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_codec, ref_lpc, ref_nn
from util import ar_frames, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _pair(seed, precision='tc_f16x3', rt='bottleneck', st=(2,)):
    from nsc_b200 import codec
    oc = ref_codec.OracleCodec(ref_codec.OracleCodecCfg(resnet_type=rt, strides=st), seed=seed)
    cfg = codec.CodecConfig(resnet_type=rt, the_strides=st, precision=precision)
    return oc, codec.NeuralCodec(cfg, torch.from_numpy(codec.pack_params_numpy(cfg, oc.conv_params, oc.alpha, oc.bins)).to(DEV))


def _frame_err(a, b):
    """per-frame max error relative to that frame's own peak"""
    a = np.asarray(a, np.float64).reshape(len(a), -1)
    b = np.asarray(b, np.float64).reshape(len(b), -1)
    return np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-300)


# the levels the review names: unit scale, the pure-time-domain normalisation 1/33.46 (constants.py:16), -60 dB, -80 dB, +60 dB
LEVELS = [1.0, 1.0 / 33.461480140686035, 1e-3, 1e-4, 1e3]


@pytest.mark.parametrize('precision', ['fp32', 'tc_f16x3'])
def test_encoder_per_frame_error_over_80_db(precision):
    """ENCODER (28 conv layers + tanh head, no quantiser decision inside): per-frame error of the floating code against float64
    truth, frames of ONE batch spanning 1e-4 ... 1e3 (a quiet frame cannot hide behind a loud one).  The float32 oracle's own
    distance from truth is the yardstick.  MEASURED (B200): at levels 1e-4 ... 1 the tensor-core path sits at 3-5e-6, the same as
    the float32 oracle (biases dominate quiet frames; nothing collapses when the fp16 lo half goes subnormal).  At level 1e3 the
    pre-tanh sums cancel from ~1e3 down to O(1), which exposes the mantissa: fp16 hi + fp16 lo carries 22 bits against float32's 24,
    so the error is ~4x the float32 oracle's (3.5e-4 vs 8e-5 of the code peak).  Hence: within the 1e-4 bar for |x| <= 64 -- the
    domain stated in include/nsc_b200.h (the reference normalises to unit variance or to 1/33.46, constants.py:16) -- and at most
    5x the float32 error anywhere."""
    oc, gc = _pair(seed=3, precision=precision)
    base = ar_frames(len(LEVELS) * 3, 512, seed=123, std=1.0)
    lv = np.repeat(np.array(LEVELS, np.float32), 3)
    x = (base * lv[:, None]).astype(np.float32)
    oc.ps._cursor = 0
    truth = oc.encoder(torch.from_numpy(x).double()[:, :, None])[:, :, 0].numpy()
    oc.ps._cursor = 0
    ref32 = oc.encoder(torch.from_numpy(x)[:, :, None])[:, :, 0].numpy()
    got = gc.encode(cu(x))['floating_code'].cpu().numpy()
    e_gpu, e_ref = _frame_err(got, truth), _frame_err(ref32, truth)
    print(f"{precision}: per-frame error vs float64 truth by level "
          + ", ".join(f"{l:g}: gpu {e_gpu[3 * i:3 * i + 3].max():.1e} / f32 oracle {e_ref[3 * i:3 * i + 3].max():.1e}" for i, l in enumerate(LEVELS)))
    supported = lv <= 64.0
    assert (e_gpu[supported] <= e_ref[supported] + 1e-4).all(), (e_gpu, e_ref)
    assert (e_gpu[supported] < 2e-5).all(), e_gpu                     # in fact float32-level on the whole supported range
    assert (e_gpu <= 5.0 * e_ref + 2e-5).all(), (e_gpu, e_ref)        # 22 vs 24 mantissa bits, nothing worse, at any level


def test_decoder_per_frame_error_on_identical_codes():
    """DECODER: identical hard codes in, per-frame waveform error vs the float32 oracle (the codes fix the level: bins in [-1, 1])."""
    oc, gc = _pair(seed=4)
    x = ar_frames(12, 512, seed=9, std=0.3)
    enc = gc.encode(cu(x))
    code = enc['code'].cpu().numpy()
    oc.forward(torch.from_numpy(x)[:1, :, None], False, 1.0)
    n_enc = 1 + 1 + 1 + 2 * 2 * 3
    oc.ps._cursor = n_enc
    out_o = oc.decoder(torch.from_numpy(code)[:, :, None])[:, :, 0].numpy()
    out_g = gc.decode_indices(enc['idx']).cpu().numpy()
    assert _frame_err(out_g, out_o).max() < 1e-4


def test_cq2_direct_oracle_comparison_beyond_one_pass():
    """The headline workload at one engine pass + 300 frames (4,144 + 300: a ragged second pass) against the oracle
    run on the SAME frames: LSF codes bit-exact, LPC polynomial / residual within 1e-4, soft-path decoded and synthesized audio
    within 1e-4, hard codes agreeing except at quantiser boundaries, hard-path audio within 1e-4 on frames whose codes all agree."""
    from nsc_b200 import codec, lpc_utilities as lu
    pairs = [_pair(seed=5), _pair(seed=6)]
    cm = codec.CMRL([p[1] for p in pairs], res_scalar=1.0)
    B = cm.pass_frames() + 300
    win = ar_frames(B, 1024, seed=77, std=1.0)
    x = np.ascontiguousarray(win[:, 256:768])
    bins = np.load(os.path.join(GOLD, 'lsf_bins_f64.npy')).astype(np.float32)
    lsf_g = lu.lpc_analysis_windows(cu(win), 16, dtype=torch.float32)
    lsf_o = ref_lpc.lpc_analysis_windows(win, 16).astype(np.float32)
    assert rel_err(lsf_g.cpu().numpy(), lsf_o) < 1e-6
    lsf = lsf_g.cpu().numpy()                      # both sides continue from the same float32 LSFs
    soft_g = cm.feedforward_lpc(cu(x), cu(lsf), True, 1.0)
    hard_g = cm.feedforward_lpc(cu(x), cu(lsf), False, 1.0)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        soft_o = ref_codec.cq_feedforward([p[0] for p in pairs], -300.0, bins, torch.from_numpy(x)[:, :, None],
                                          torch.from_numpy(lsf)[:, :, None], True, 1.0)
        hard_o = ref_codec.cq_feedforward([p[0] for p in pairs], -300.0, bins, torch.from_numpy(x)[:, :, None],
                                          torch.from_numpy(lsf)[:, :, None], False, 1.0)
    idx_o = ref_nn.quantizer_indices(torch.from_numpy(lsf)[:, :, None], -300.0, bins).numpy().astype(np.uint8)
    assert np.array_equal(hard_g['lsf_idx'].cpu().numpy(), idx_o)
    assert rel_err(hard_g['res_x'].cpu().numpy(), hard_o['res_x'].numpy()[:, :, 0]) < 1e-4
    assert rel_err(soft_g['decoded'].cpu().numpy(), soft_o['decoded'].numpy()) < 1e-4
    assert rel_err(soft_g['synthesized'].cpu().numpy(), soft_o['synthesized']) < 1e-4
    # hard path: code agreement, then audio on the frames where every code of both codecs agrees
    same = np.ones(B, bool)
    agree = []
    for k in range(2):
        # (the oracle's hard value path: one-hot @ bins -> recover indices by nearest bin)
        code_o = hard_o['per'][k]['code'].numpy()[:, :, 0]
        idx_k = np.abs(code_o[:, :, None] - pairs[k][0].bins[None, None, :]).argmin(-1)
        eq = idx_k == hard_g['idx'][k].cpu().numpy()
        agree.append(eq.mean())
        same &= eq.all(1)
    print(f"cq2 hard-code agreement with the float32 oracle over {B} frames: codec 1 {agree[0]:.6f}, codec 2 {agree[1]:.6f}; "
          f"frames with every code equal: {same.mean():.4f}")
    assert agree[0] > 0.999 and agree[1] > 0.995
    assert same.mean() > 0.5
    e = _frame_err(hard_g['synthesized'].cpu().numpy()[same], np.asarray(hard_o['synthesized'])[same])
    assert np.quantile(e, 0.99) < 1e-4 and rel_err(hard_g['synthesized'].cpu().numpy()[same], np.asarray(hard_o['synthesized'])[same]) < 1e-4


def test_hard_code_agreement_vs_float64_truth():
    """What a user sees: which fraction of hard codes differs from an exact (float64) evaluation of the same network -- for the
    float32 oracle and for the GPU.  A code can only flip when the floating code sits within rounding noise of a bin mid-point, so
    both rates are tiny; the GPU must not be further from truth than the float32 reference (+ 5e-4 slack for counting noise)."""
    oc, gc = _pair(seed=8)
    B = 2072
    x = ar_frames(B, 512, seed=314, std=0.3)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        oc.ps._cursor = 0
        f64 = oc.encoder(torch.from_numpy(x).double()[:, :, None]).numpy()
        oc.ps._cursor = 0
        f32 = oc.encoder(torch.from_numpy(x)[:, :, None]).numpy()
    bins64 = oc.bins.astype(np.float64)
    idx_truth = np.abs(f64 - bins64[None, None, :]).argmin(-1)
    idx_f32 = ref_nn.quantizer_indices(torch.from_numpy(f32), oc.alpha, oc.bins).numpy()
    idx_gpu = gc.encode(cu(x))['idx'].cpu().numpy()
    d_f32 = float((idx_f32 != idx_truth).mean())
    d_gpu = float((idx_gpu != idx_truth).mean())
    d_pair = float((idx_gpu != idx_f32).mean())
    print(f"hard codes differing from float64 truth over {B * 256} codes: float32 oracle {d_f32:.2e}, GPU (tc_f16x3) {d_gpu:.2e}; "
          f"GPU vs float32 oracle {d_pair:.2e}")
    assert d_gpu <= d_f32 + 5e-4
    assert d_pair < 2e-3


def test_gated_bottleneck_decoder_vs_oracle():
    """nn_core_operator.py:115-137 (no caller in the reference; surface-only)."""
    from nsc_b200 import nn_core_operator as nn
    for wide, L, dil, flat in ((100, 256, 1, False), (50, 128, 2, True)):
        ps = ref_nn.ParamStream(seed=wide + dil)
        x = np.random.RandomState(3).randn(2, L, wide).astype(np.float32)
        ref = ref_nn.gated_bottleneck_decoder(torch.from_numpy(x), wide_layer=wide, narrow_layer=20, dilation_rate=dil,
                                              is_last_flat=flat, ps=ps).numpy()
        params = [tuple(cu(p) for p in t) for t in ps.params]
        got = nn.gated_bottleneck_decoder(cu(x), wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=flat, params=params)
        assert got.shape == ref.shape
        assert rel_err(got.cpu().numpy(), ref) < 5e-5


@pytest.mark.parametrize('B', [5, 32])
def test_gln_codec_on_the_plane_engine_per_frame_vs_oracle(B):
    """The reference's SHIPPED block type (constants.py:14 resnet_type = 'gln'; gated_bottleneck, nn_core_operator.py:82-112; separable
    up-conv, nscm.py:175-177) on the plane engine: fused k15 gate pair with the gate product in the epilogue, the dilation-2 gates on
    de-interleaved sub-images, depthwise + pointwise up-conv.  Per-frame error of the encoder's floating code vs float64 truth (the
    float32 oracle's own distance is the yardstick) and of the decoder on identical codes vs the float32 oracle; odd and even batches
    (CTA pairs need an even number of tiles)."""
    from nsc_b200 import _lib
    oc, gc = _pair(seed=11, rt='gln')
    x = ar_frames(B, 512, seed=B, std=0.3)
    x[1] *= 1e-2                                                      # a quiet frame must not hide behind the loud ones
    oc.ps._cursor = 0
    truth = oc.encoder(torch.from_numpy(x).double()[:, :, None])[:, :, 0].numpy()
    oc.ps._cursor = 0
    ref32 = oc.encoder(torch.from_numpy(x)[:, :, None])[:, :, 0].numpy()
    n_enc = oc.ps._cursor
    enc = gc.encode(cu(x))
    got = enc['floating_code'].cpu().numpy()
    e_gpu, e_ref = _frame_err(got, truth), _frame_err(ref32, truth)
    assert (e_gpu <= e_ref + 1e-4).all() and (e_gpu < 5e-5).all(), (e_gpu, e_ref)
    code = enc['code'].cpu().numpy()
    oc.ps._cursor = n_enc
    out_o = oc.decoder(torch.from_numpy(code)[:, :, None])[:, :, 0].numpy()
    out_g = gc.decode_indices(enc['idx']).cpu().numpy()
    assert _frame_err(out_g, out_o).max() < 1e-4
    lib = _lib.load()
    cfg = gc.cfg.to_struct()
    import ctypes as C
    assert lib.nsc_codec_on_plane_engine(C.byref(cfg)) == 1           # the tcgen05 plane path, not the layer-by-layer engines
