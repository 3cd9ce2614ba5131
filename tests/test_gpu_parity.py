"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json:north_star): quantiser codes bit-exact given identical pre-quantisation inputs; waveforms,
LPC coefficients and losses within 1e-4 relative error in fp32 (rel_err = max|a-b| / max|b|).
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_codec, ref_loss, ref_lpc, ref_nn
from util import ar_frames, quantizer_edge_codes, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TOL = 1e-4
DEV = 'cuda'


def cu(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(DEV)


def lsf_bins():
    return np.load(os.path.join(GOLD, 'lsf_bins_f64.npy')).astype(np.float32)


# ------------------------------------------------------------------------------------------------ quantiser
def _quant_case(x, bins, alpha, L):
    from nsc_b200 import nn_core_operator as nn
    n = len(bins)
    xt = torch.from_numpy(x)
    out = {}
    for share in (False, True):
        soft_o, code_o = ref_nn.scalar_softmax_quantization(xt, alpha, bins, 1.0, share, L, n)
        soft_g, code_g, idx_g = nn.scalar_softmax_quantization(cu(x), alpha, cu(bins), 1.0, share, L, n,
                                                               return_indices=True)
        idx_o = ref_nn.quantizer_indices(xt, alpha, bins)
        assert np.array_equal(idx_g.cpu().numpy(), idx_o.numpy().astype(np.uint8)), "codes not bit-exact"
        assert np.abs(soft_g.cpu().numpy() - soft_o.numpy()).max() < 1e-5
        if share:
            assert rel_err(code_g.cpu().numpy(), code_o.numpy()) < 1e-5
        else:
            assert np.array_equal(code_g.cpu().numpy(), code_o.numpy()), "hard value must be bins[idx] exactly"
        out[share] = (soft_g, code_g, idx_g)
    return out


def test_quantizer_32_bins_random_and_edges():
    bins = np.linspace(-1, 1, 32).astype(np.float32)
    rng = np.random.RandomState(0)
    x = np.concatenate([quantizer_edge_codes(bins), rng.uniform(-1.1, 1.1, 256 * 37).astype(np.float32)])
    x = x[: (len(x) // 256) * 256].reshape(-1, 256, 1)
    _quant_case(x, bins, -300.0, 256)


def test_quantizer_other_bin_counts_alphas_unsorted_duplicates():
    rng = np.random.RandomState(1)
    for n, L, alpha in ((64, 128, -300.0), (16, 256, -150.0), (33, 50, -40.0), (128, 16, -300.0)):
        bins = np.linspace(-1, 1, n).astype(np.float32)
        rng.shuffle(bins)                 # unsorted
        bins[3] = bins[7]                 # duplicate: the lower index must win
        x = np.concatenate([quantizer_edge_codes(bins), rng.uniform(-1.2, 1.2, L * 9).astype(np.float32)])
        x = x[: (len(x) // L) * L].reshape(-1, L, 1)
        _quant_case(x, bins, alpha, L)


def test_quantizer_lsf_codebook_256_unsorted():
    bins = lsf_bins()
    rng = np.random.RandomState(2)
    x = np.concatenate([quantizer_edge_codes(bins), rng.uniform(0, np.pi, 16 * 200).astype(np.float32)])
    x = x[: (len(x) // 16) * 16].reshape(-1, 16, 1)
    _quant_case(x, bins, -300.0, 16)


def test_quantizer_golden_fixture_and_pretrain_blend():
    from nsc_b200 import nn_core_operator as nn
    g = np.load(os.path.join(GOLD, 'quantizer.npz'))
    bins = np.linspace(-1, 1, 32).astype(np.float32)
    _, code, idx = nn.scalar_softmax_quantization(cu(g['x32']), -300.0, cu(bins), 1.0, False, 64, 32, return_indices=True)
    assert np.array_equal(idx.cpu().numpy(), g['idx32']) and np.array_equal(code.cpu().numpy(), g['code32'])
    _, code, idx = nn.scalar_softmax_quantization(cu(g['xl']), -300.0, cu(lsf_bins()), 1.0, False, 16, 256, return_indices=True)
    assert np.array_equal(idx.cpu().numpy(), g['idx_l']) and np.array_equal(code.cpu().numpy(), g['code_l'])
    # is_quan_on = 0 (pre-training epochs, nscm.py:560-568): the value path is the floating code itself
    _, code0 = nn.scalar_softmax_quantization(cu(g['x32']), -300.0, cu(bins), 0.0, True, 64, 32)
    assert np.array_equal(code0.cpu().numpy(), g['x32'])


def test_quantizer_stats_hist_qloss_entropy():
    from nsc_b200 import _lib, loss_terms_and_measures as lt
    bins = np.linspace(-1, 1, 32).astype(np.float32)
    x = np.random.RandomState(3).uniform(-1, 1, (9, 256, 1)).astype(np.float32)
    soft_o, _ = ref_nn.scalar_softmax_quantization(torch.from_numpy(x), -20.0, bins, 1.0, True, 256, 32)
    xt = cu(x)
    hist = torch.zeros(32, device=DEV)
    ql = torch.empty(9, device=DEV)
    alpha = torch.tensor([-20.0], device=DEV)
    rc = _lib.load().nsc_quantize_scalar(_lib.ptr(xt), 9, 256, _lib.ptr(cu(bins)), 32, _lib.ptr(alpha), 1.0, 1, None, None, None,
                                         _lib.ptr(hist), _lib.ptr(ql), _lib.stream_ptr())
    assert rc == 0
    assert rel_err(hist.cpu().numpy(), soft_o.reshape(-1, 32).sum(0).numpy()) < TOL
    assert rel_err(ql.cpu().numpy(), ref_loss.quan_loss(soft_o).numpy()) < TOL
    ent = lt.entropy_from_hist(hist)
    assert abs(float(ent) - float(ref_loss.entropy_coding_loss(soft_o))) < 1e-4
    # surface functions on a materialised soft tensor
    sg = cu(soft_o.numpy())
    assert rel_err(lt.quan_loss(sg).cpu().numpy(), ref_loss.quan_loss(soft_o).numpy()) < TOL
    assert abs(float(lt.entropy_coding_loss(sg)) - float(ref_loss.entropy_coding_loss(soft_o))) < 1e-4


# ------------------------------------------------------------------------------------------------ convs
CONV_CASES = [
    # B, L, cin, cout, k, dil, stride, act
    (3, 512, 1, 100, 55, 1, 1, None),
    (2, 512, 100, 20, 9, 1, 1, 'leaky_relu'),
    (2, 512, 20, 20, 9, 2, 1, 'leaky_relu'),
    (2, 256, 20, 100, 9, 1, 1, None),
    (2, 512, 100, 100, 9, 1, 2, 'leaky_relu'),
    (2, 256, 100, 1, 55, 1, 1, 'tanh'),
    (2, 512, 50, 1, 55, 1, 1, None),
    (2, 256, 1, 20, 9, 1, 1, None),
    (2, 512, 50, 20, 9, 1, 1, 'tanh'),
    (2, 128, 20, 20, 15, 2, 1, 'tanh'),
    (2, 128, 100, 20, 1, 1, 1, 'leaky_relu'),
    (1, 500, 7, 13, 9, 4, 1, None),      # generic fallback kernel, ragged length, odd channels
    (2, 37, 3, 5, 3, 1, 1, 'tanh'),      # tiny ragged
    (1, 511, 6, 10, 5, 1, 3, None),      # stride 3 generic
    (2, 100, 25, 25, 9, 1, 1, None),     # stride-[2,2] decoder width
]


@pytest.mark.parametrize('B,L,cin,cout,k,dil,stride,act', CONV_CASES)
def test_conv1d_vs_oracle(B, L, cin, cout, k, dil, stride, act):
    from nsc_b200 import nn_core_operator as nn
    rng = np.random.RandomState(k * 1000 + cin + cout)
    x = rng.randn(B, L, cin).astype(np.float32)
    w = (rng.randn(k, cin, cout) / np.sqrt(k * cin)).astype(np.float32)
    b = rng.randn(cout).astype(np.float32) * 0.1
    ref = ref_nn.conv1d_explicit(torch.from_numpy(x), w, b, dil, stride, act).numpy()
    got = nn.conv1d(cu(x), cout, k, dilation_rate=dil, strides=stride, activation=act, params=(cu(w), cu(b)))
    assert got.shape == ref.shape
    assert rel_err(got.cpu().numpy(), ref) < 2e-5


def test_conv1d_depth_vs_oracle():
    from nsc_b200 import nn_core_operator as nn
    rng = np.random.RandomState(5)
    x = rng.randn(2, 256, 100).astype(np.float32)
    dw = (rng.randn(9, 100, 1) / 3).astype(np.float32)
    pw = (rng.randn(1, 100, 100) / 10).astype(np.float32)
    b = rng.randn(100).astype(np.float32) * 0.1
    ref = ref_nn.conv1d_depth_explicit(torch.from_numpy(x), dw, pw, b, 1, 1, 'leaky_relu').numpy()
    got = nn.conv1d_depth(cu(x), 100, 9, activation='leaky_relu', params=(cu(dw), cu(pw), cu(b)))
    assert rel_err(got.cpu().numpy(), ref) < 2e-5


@pytest.mark.parametrize('gated', [False, True])
@pytest.mark.parametrize('cin,wide,L,dil,flat', [(100, 100, 512, 1, False), (100, 100, 256, 2, True),
                                                 (1, 100, 256, 1, False), (50, 50, 512, 2, True)])
def test_blocks_vs_oracle(gated, cin, wide, L, dil, flat):
    from nsc_b200 import nn_core_operator as nn
    ps = ref_nn.ParamStream(seed=wide + cin + dil)
    x = np.random.RandomState(8).randn(2, L, cin).astype(np.float32)
    fn_o = ref_nn.gated_bottleneck if gated else ref_nn.the_bottleneck
    ref = fn_o(torch.from_numpy(x), wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=flat, ps=ps).numpy()
    params = [tuple(cu(p) for p in t) for t in ps.params]
    fn_g = nn.gated_bottleneck if gated else nn.the_bottleneck
    got = fn_g(cu(x), wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=flat, params=params)
    assert rel_err(got.cpu().numpy(), ref) < 5e-5


# ------------------------------------------------------------------------------------------------ LPC
def test_lpc_golden_and_oracle():
    from nsc_b200 import lpc_utilities as lu
    g = np.load(os.path.join(GOLD, 'lpc.npz'))
    lsf = lu.lpc_analysis_windows(cu(g['windows']), 16, strict=True)
    assert lsf.dtype == torch.float64
    assert rel_err(lsf.cpu().numpy(), g['lsf']) < 1e-8
    poly = lu.lsf2poly_after_quan(cu(g['lsf'].astype(np.float32)), 16, strict=True)
    assert rel_err(poly.cpu().numpy(), g['poly']) < TOL          # the reference path is complex64 inside np.poly
    p64 = np.stack([ref_lpc.lsf2poly(r.astype(np.float64)) for r in g['lsf'].astype(np.float32)])
    assert rel_err(poly.cpu().numpy(), p64) < 1e-6                  # vs float64 truth: float32 output rounding only
    res = lu.lpc_analysis_get_residual(cu(g['frames'])[:, :, None], cu(g['poly']))
    assert np.abs(res.cpu().numpy() - g['res']).max() <= 1e-6 * np.abs(g['res']).max()
    syn = lu.lpc_synthesizer_tr(cu(g['poly']), cu(g['res']))
    assert rel_err(syn.cpu().numpy(), g['syn']) < 1e-6
    lsf_tr = lu.lpc_analysis_at_train(cu(g['frames'])[:, :, None], 16, strict=True)
    assert rel_err(lsf_tr.cpu().numpy(), g['lsf_train']) < 1e-8


def test_lpc_analysis_at_test_windowing_quirk():
    from nsc_b200 import lpc_utilities as lu
    seg = ar_frames(9, 512, seed=77)
    ref = ref_lpc.lpc_analysis_at_test(seg, 16)
    got = lu.lpc_analysis_at_test(cu(seg), 16)
    assert got.shape == ref.shape == (7, 16)
    assert rel_err(got.cpu().numpy(), ref) < 1e-8


def test_lpc_error_behaviour():
    from nsc_b200 import lpc_utilities as lu
    bad = np.full((2, 16), 0.5, dtype=np.float32); bad[1, 3] = 3.5     # > pi
    with pytest.raises(ValueError):
        lu.lsf2poly_after_quan(cu(bad), 16, strict=True)
    out = lu.lsf2poly_after_quan(cu(bad), 16)
    assert torch.isfinite(out[0]).all() and torch.isnan(out[1]).all()
    with pytest.raises(ZeroDivisionError):
        lu.lpc_analysis_windows(torch.zeros(1, 1024, device=DEV), 16, strict=True)
    assert lu.lpc_analysis_windows(torch.zeros(0, 1024, device=DEV)).shape == (0, 16)    # empty batch


def test_lpc_full_size_roundtrip_and_linearity():
    """Size-independent properties at a production batch: synthesis undoes a zero-state analysis FIR over the
    whole frame is NOT what the sub-framed residual is, so check (a) linearity of the residual in x and
    (b) analysis windows -> LSF sortedness, (c) lsf2poly(poly2lsf) fixed point through the GPU path."""
    from nsc_b200 import lpc_utilities as lu
    B = 4096
    win = cu(ar_frames(B, 1024, seed=101))
    lsf = lu.lpc_analysis_windows(win, 16, strict=True)
    assert bool((lsf[:, 1:] > lsf[:, :-1]).all()) and float(lsf.min()) > 0 and float(lsf.max()) < np.pi
    poly = lu.lsf2poly_after_quan(lsf.float(), 16)
    x1 = cu(ar_frames(B, 512, seed=102)); x2 = cu(ar_frames(B, 512, seed=103))
    r1 = lu.lpc_analysis_get_residual(x1, poly); r2 = lu.lpc_analysis_get_residual(x2, poly)
    r12 = lu.lpc_analysis_get_residual(x1 + x2, poly)
    assert float((r12 - (r1 + r2)).abs().max()) < 1e-4 * float(r12.abs().max())
    # synthesis is the inverse of the un-subframed analysis filter: check via the impulse response identity
    imp = torch.zeros(B, 512, device=DEV); imp[:, 0] = 1.0
    h = lu.lpc_synthesizer_tr(poly, imp)                       # h = 1/A
    # A * h = delta  (first 512 samples)
    a = poly.double(); hd = h.double()
    conv = torch.zeros(B, 64, dtype=torch.float64, device=DEV)
    for k in range(17):
        conv[:, k:] += a[:, k:k + 1] * hd[:, :64 - k]
    conv[:, 0] -= 1.0
    assert float(conv.abs().max()) < 1e-5


def test_lpc_analysis_batch_shapes_agree_and_match_the_oracle():
    """The analysis kernel works in batches of one or two frames per warp depending on the batch size, with ragged tails and silent
    frames in the middle: the same windows must give the same LSFs to the bit whichever shape runs, and the oracle's to 1e-8."""
    from nsc_b200 import lpc_utilities as lu
    N = 5003                                             # > 2 CTAs x 8 warps x 2 frames x 148 SMs: the two-frames-per-warp shape
    win_np = ar_frames(N, 1024, seed=211)
    win_np[17] = 0.0
    win_np[4999] = 0.0
    win = cu(win_np)
    whole = lu.lpc_analysis_windows(win, 16)
    parts = torch.cat([lu.lpc_analysis_windows(win[i:i + 997], 16) for i in range(0, N, 997)])
    assert torch.equal(torch.isnan(whole), torch.isnan(parts))
    assert torch.equal(torch.nan_to_num(whole), torch.nan_to_num(parts))
    assert bool(torch.isnan(whole[17]).all()) and bool(torch.isnan(whole[4999]).all()) and int(torch.isnan(whole).any(dim=1).sum()) == 2
    pick = [0, 1, 16, 18, 2500, 4998, 5002]
    ref = ref_lpc.lpc_analysis_windows(win_np[pick], 16)
    assert rel_err(whole[pick].cpu().numpy(), ref) < 1e-8


# ------------------------------------------------------------------------------------------------ losses
def test_losses_vs_oracle_and_golden():
    from nsc_b200 import loss_terms_and_measures as lt
    g = np.load(os.path.join(GOLD, 'losses.npz'))
    t, f = lt.losses(cu(g['dec']), cu(g['ori']))
    assert rel_err(t.cpu().numpy(), g['time_loss']) < TOL
    assert rel_err(f.cpu().numpy(), g['freq_loss']) < TOL
    assert rel_err(lt.mse_loss(cu(g['dec']), cu(g['ori'])).cpu().numpy(), g['time_loss']) < TOL
    assert rel_err(lt.mfcc_loss(cu(g['dec']), cu(g['ori'])).cpu().numpy(), g['freq_loss']) < TOL
    a = ar_frames(300, 512, seed=55, std=0.2)
    b = (a * 0.9 + 0.02 * np.random.RandomState(56).randn(*a.shape)).astype(np.float32)
    t, f = lt.losses(cu(b), cu(a))
    assert rel_err(t.cpu().numpy(), ref_loss.mse_loss(torch.from_numpy(b), torch.from_numpy(a)).numpy()) < TOL
    assert rel_err(f.cpu().numpy(), ref_loss.mfcc_loss(torch.from_numpy(b), torch.from_numpy(a)).numpy()) < TOL
    # identical signals: both losses are sqrt(1e-7)
    t, f = lt.losses(cu(a), cu(a))
    assert np.allclose(t.cpu().numpy(), np.sqrt(1e-7), rtol=1e-5) and np.allclose(f.cpu().numpy(), np.sqrt(1e-7), rtol=1e-5)


def test_mel_filterbank_matches_oracle():
    from nsc_b200 import _lib, loss_terms_and_measures as lt
    buf = lt.mel_filterbank(DEV).cpu().numpy()
    m = buf[:257 * 184].reshape(257, 184)
    ref = np.concatenate([ref_loss.linear_to_mel_weight_matrix(n, 257, 16000, 0.0, 8000.0) for n in (8, 16, 32, 128)], axis=1)
    assert np.abs(m - ref).max() < 1e-6


# ------------------------------------------------------------------------------------------------ codec / cascade / CQ
def _make_pair(rt, st, seed, nbins=32, precision='tc_f16x3'):
    from nsc_b200 import codec
    oc = ref_codec.OracleCodec(ref_codec.OracleCodecCfg(resnet_type=rt, strides=st, num_bins=nbins), seed=seed)
    cfg = codec.CodecConfig(resnet_type=rt, the_strides=st, num_bins=nbins, precision=precision)
    flat = codec.pack_params_numpy(cfg, oc.conv_params, oc.alpha, oc.bins)
    return oc, codec.NeuralCodec(cfg, torch.from_numpy(flat).to(DEV))


def _check_codec(oc, gc, x, the_share, tol=TOL):
    r_o = oc.forward(torch.from_numpy(x)[:, :, None], the_share, 1.0)
    r_g = gc.computational_graph_end2end_quan_on(cu(x), the_share, 1.0, want_soft=True, want_stats=True)
    fl_g = r_g['floating_code'].cpu().numpy()
    assert rel_err(fl_g, r_o['floating_code'].numpy()[:, :, 0]) < tol
    # codes: bit-exact given identical pre-quantisation inputs -> re-quantise the GPU's floating code with the oracle
    idx_o = ref_nn.quantizer_indices(torch.from_numpy(fl_g)[:, :, None], oc.alpha, oc.bins).numpy()
    assert np.array_equal(r_g['idx'].cpu().numpy(), idx_o.astype(np.uint8))
    # decoder parity on identical codes
    code_g = r_g['code'].cpu().numpy()
    oc.ps._cursor = _enc_layers(oc)
    out_o = oc.decoder(torch.from_numpy(code_g)[:, :, None])[:, :, 0].numpy()
    assert rel_err(r_g['out'].cpu().numpy(), out_o) < tol
    return r_o, r_g


def _enc_layers(oc):
    n = 1 + len(oc.cfg.strides) * (1) + 1
    per_block = 3 if oc.cfg.resnet_type == 'bottleneck' else 4
    nb = len(oc.cfg.bottleneck_kernel_and_dilation) - 4
    return n + (len(oc.cfg.strides) + 1) * nb * per_block


@pytest.mark.parametrize('precision', ['fp32', 'tc_f16x3'])
@pytest.mark.parametrize('rt,st', [('bottleneck', (2,)), ('gln', (2,)), ('bottleneck', (2, 2)), ('gln', (2, 2))])
@pytest.mark.parametrize('the_share', [False, True])
def test_codec_forward_vs_oracle(rt, st, the_share, precision):
    """Both fp32-class engines (FFMA and the tcgen05 fp16 hi/lo split) must meet the 1e-4 bar."""
    oc, gc = _make_pair(rt, st, seed=3, precision=precision)
    x = ar_frames(3, 512, seed=31, std=0.3)
    r_o, r_g = _check_codec(oc, gc, x, the_share)
    if the_share:
        # The soft path has no discontinuity, so the END-TO-END output can be compared directly -- but the alpha=-300
        # soft quantiser is steep (slope ~ alpha * bin spacing ~ 20), so fp32 rounding noise of the encoder is
        # amplified: measure the fp32 oracle's own distance from float64 truth and allow 1e-4 on top of it.
        r64 = oc.forward(torch.from_numpy(x).double()[:, :, None], the_share, 1.0)['out'].numpy()
        floor = rel_err(r_o['out'].numpy(), r64)
        assert rel_err(r_g['out'].cpu().numpy(), r64) < TOL + floor


def test_codec_reduced_precision_fp16_is_stated_separately():
    """precision='tc_f16' (plain fp16 tensor-core inputs) is NOT an fp32-parity mode: it is reported separately
    (BASELINE.json:north_star "stated separately if bf16 is used").  Bound here: 2e-2 on the decoder output."""
    oc, gc = _make_pair('bottleneck', (2,), seed=3, precision='tc_f16')
    x = ar_frames(3, 512, seed=31, std=0.3)
    r_o = oc.forward(torch.from_numpy(x)[:, :, None], True, 1.0)
    r_g = gc.computational_graph_end2end_quan_on(cu(x), True, 1.0)
    e_code = rel_err(r_g['floating_code'].cpu().numpy(), r_o['floating_code'].numpy()[:, :, 0])
    e_out = rel_err(r_g['out'].cpu().numpy(), r_o['out'].numpy())
    print(f"tc_f16 reduced precision: floating code {e_code:.2e}, decoder output {e_out:.2e}")
    assert e_code < 2e-2 and e_out < 2e-2


def test_precision_engines_agree_at_full_size():
    """Tensor engine (split) vs FFMA engine on a production batch: same codes except at quantiser boundaries,
    same waveform within 1e-4 where the codes agree."""
    oc, g32 = _make_pair('bottleneck', (2,), seed=4, precision='fp32')
    _, gtc = _make_pair('bottleneck', (2,), seed=4, precision='tc_f16x3')
    x = cu(ar_frames(512, 512, seed=62, std=0.3))
    a = g32.computational_graph_end2end_quan_on(x, True, 1.0)
    b = gtc.computational_graph_end2end_quan_on(x, True, 1.0)
    assert rel_err(b['floating_code'].cpu().numpy(), a['floating_code'].cpu().numpy()) < TOL
    assert rel_err(b['out'].cpu().numpy(), a['out'].cpu().numpy()) < TOL
    agree = float((a['idx'] == b['idx']).float().mean())
    assert agree > 0.999, agree


def test_codec_golden_fixture():
    g = np.load(os.path.join(GOLD, 'codec.npz'))
    for name, rt, st in [('bn2', 'bottleneck', (2,)), ('gln2', 'gln', (2,)), ('bn4', 'bottleneck', (2, 2))]:
        oc, gc = _make_pair(rt, st, seed=3)
        r = gc.computational_graph_end2end_quan_on(cu(g[name + '_x']), False, 1.0)
        assert rel_err(r['floating_code'].cpu().numpy(), g[name + '_floating']) < TOL
        # the fixture's inputs sit away from quantiser boundaries (tests/golden/make_golden.py), so every hard code must agree --
        # a flipped code fails here instead of silently skipping the decoder check
        same = r['code'].cpu().numpy() == g[name + '_code']
        assert same.all(), f"{name}: {int((~same).sum())} of {same.size} hard codes differ from the golden fixture"
        assert rel_err(r['out'].cpu().numpy(), g[name + '_out']) < TOL


def test_codec_encode_decode_split_equals_fused_and_chunking():
    """Properties at a production batch (B > the 2048-frame internal chunk, ragged last chunk)."""
    oc, gc = _make_pair('bottleneck', (2,), seed=4)
    B = 2048 + 300
    x = cu(ar_frames(B, 512, seed=61, std=0.3))
    full = gc.computational_graph_end2end_quan_on(x, False, 1.0)
    enc = gc.encode(x)
    assert torch.equal(enc['idx'], full['idx']) and torch.equal(enc['code'], full['code'])
    dec = gc.decode_indices(enc['idx'])
    assert torch.equal(dec, full['out'])
    # batch invariance: the first 7 frames alone give the same bits as inside the big batch
    small = gc.computational_graph_end2end_quan_on(x[:7].contiguous(), False, 1.0)
    assert torch.equal(small['idx'], full['idx'][:7]) and torch.equal(small['out'], full['out'][:7])
    # hard value is a codebook entry
    assert torch.equal(full['code'], gc.bins[full['idx'].long()])


def test_cascade_vs_oracle():
    from nsc_b200 import codec
    pairs = [_make_pair('bottleneck', (2,), seed=5), _make_pair('bottleneck', (2,), seed=6), _make_pair('gln', (2,), seed=7, nbins=64)]
    x = ar_frames(3, 512, seed=71, std=0.3)
    for lpc_variant, rs in ((False, 1.0), (False, 2.0), (True, 2.0)):
        cm = codec.CMRL([p[1] for p in pairs], res_scalar=rs)
        r = cm.all_modules_feedforward(cu(x), True, 1.0, lpc_variant=lpc_variant, want_stats=True, want_outs=True)
        dec_o, outs_o, per_o = ref_codec.cascade_forward([p[0] for p in pairs], torch.from_numpy(x)[:, :, None], True, 1.0,
                                                         res_scalar=rs, lpc_variant=lpc_variant)
        assert rel_err(r['decoded'].cpu().numpy(), dec_o.numpy()) < TOL
        for i in range(3):
            assert rel_err(r['outs'][i].cpu().numpy(), outs_o[i].numpy()) < 2 * TOL
            assert rel_err(r['hist'][i].cpu().numpy(), per_o[i]['soft'].reshape(-1, per_o[i]['soft'].shape[2]).sum(0).numpy()) < 1e-3
            assert rel_err(r['qloss'][i].cpu().numpy(), ref_loss.quan_loss(per_o[i]['soft']).numpy()) < 1e-3
        assert torch.allclose(r['decoded'], sum(r['outs']), atol=1e-6)


def test_cq_feedforward_vs_oracle_and_golden():
    from nsc_b200 import codec, loss_terms_and_measures as lt
    g = np.load(os.path.join(GOLD, 'cq.npz'))
    pairs = [_make_pair('bottleneck', (2,), seed=5), _make_pair('bottleneck', (2,), seed=6)]
    cm = codec.CMRL([p[1] for p in pairs], res_scalar=1.0)
    r = cm.feedforward_lpc(cu(g['x']), cu(g['lsf']), False, 1.0, want_stats=True)
    assert rel_err(r['poly'].cpu().numpy(), g['poly']) < TOL
    assert rel_err(r['res_x'].cpu().numpy(), g['res_x']) < TOL
    # LSF codes bit-exact
    idx_o = ref_nn.quantizer_indices(torch.from_numpy(g['lsf'])[:, :, None], -300.0, lsf_bins()).numpy()
    assert np.array_equal(r['lsf_idx'].cpu().numpy(), idx_o.astype(np.uint8))
    # soft path end to end (no discontinuity) against the oracle run on the same inputs
    rs = cm.feedforward_lpc(cu(g['x']), cu(g['lsf']), True, 1.0, want_stats=True)
    o = ref_codec.cq_feedforward([p[0] for p in pairs], -300.0, lsf_bins(), torch.from_numpy(g['x'])[:, :, None],
                                 torch.from_numpy(g['lsf'])[:, :, None], True, 1.0, res_scalar=1.0)
    assert rel_err(rs['decoded'].cpu().numpy(), o['decoded'].numpy()) < TOL
    assert rel_err(rs['synthesized'].cpu().numpy(), o['synthesized']) < TOL
    t, f = lt.losses(rs['decoded'], rs['res_x'])
    assert rel_err(t.cpu().numpy(), o['time_loss'].numpy()) < TOL
    assert rel_err(f.cpu().numpy(), o['freq_loss'].numpy()) < TOL
    assert abs(float(lt.entropy_from_hist(rs['lsf_hist'])) - float(o['ent_lpc'])) < 1e-3
    for i in range(2):
        assert abs(float(lt.entropy_from_hist(rs['hist'][i])) - float(o['ent'][i])) < 1e-3
    # hard path: synthesis parity given the GPU's own decoded residual and poly
    syn_o = ref_lpc.lpc_synthesizer_tr(r['poly'].cpu().numpy(), r['decoded'].cpu().numpy())
    assert rel_err(r['synthesized'].cpu().numpy(), syn_o) < 1e-5
    # golden: first codec's floating code (before any quantiser decision) on the golden residual
    e0 = pairs[0][1].encode(cu(g['res_x']))
    assert rel_err(e0['floating_code'].cpu().numpy(), g['floating0']) < TOL


def test_empty_batch_and_errors():
    from nsc_b200 import codec, nn_core_operator as nn
    oc, gc = _make_pair('bottleneck', (2,), seed=4)
    r = gc.computational_graph_end2end_quan_on(torch.zeros(0, 512, device=DEV), False, 1.0)
    assert r['out'].shape == (0, 512) and r['idx'].shape == (0, 256)
    with pytest.raises(ValueError):
        nn.conv1d(torch.zeros(1, 8, 3, device=DEV), 4, 3, params=(torch.zeros(3, 2, 4, device=DEV), torch.zeros(4, device=DEV)))
    with pytest.raises(ValueError):
        nn.conv1d(torch.zeros(1, 8, 3, device=DEV), 4, 3, dilation_rate=2, strides=2,
                  params=(torch.zeros(3, 3, 4, device=DEV), torch.zeros(4, device=DEV)))


def test_cuda_vs_reference_run():
    """The CUDA path, through the C ABI, against vectors produced by the REFERENCE'S OWN utilities.py / lpc_utilities.py run in the
    build container (tests/golden/make_ref_golden.py: unmodified source files, third-party libraries replaced by independent
    stand-ins; committed as tests/golden/reference_run.npz): framing, the utterance filters, LPC analysis at test / train time, LSF ->
    polynomial, the sub-framed residual filter and LPC synthesis."""
    from nsc_b200 import lpc_utilities as lu, utilities as ut
    g = dict(np.load(os.path.join(GOLD, 'reference_run.npz')))
    sig = cu(g['utt'])
    assert np.array_equal(ut.utterance_to_segment(sig, True).cpu().numpy(), g['seg_plain'].astype(np.float32))
    assert rel_err(ut.utterance_to_segment(sig, False).cpu().numpy(), g['seg_windowed']) < 1e-6
    assert rel_err(ut.highpass_filter(sig, out_f64=True).cpu().numpy(), g['highpass']) < 1e-9
    assert rel_err(ut.empha_filter(sig, out_f64=True).cpu().numpy(), g['empha']) < 1e-9
    lsf = lu.lpc_analysis_at_test(cu(g['at_test_in'])).cpu().numpy()
    assert lsf.shape == (6, 16) and np.abs(lsf - g['at_test_lsf']).max() < 1e-6
    tr = lu.lpc_analysis_at_train(cu(g['at_train_in'])[:, :, None]).cpu().numpy()
    assert np.abs(tr - g['at_train_lsf']).max() < 1e-6
    poly = lu.lsf2poly_after_quan(cu(g['at_train_lsf'].astype(np.float32)), 16).cpu().numpy()
    assert rel_err(poly, g['poly']) < 1e-5
    res = lu.lpc_analysis_get_residual(cu(g['at_train_in'])[:, :, None], cu(g['poly'])).cpu().numpy()
    assert rel_err(res, g['residual']) < 1e-5
    syn = lu.lpc_synthesizer_tr(cu(g['poly']), cu(g['residual'])).cpu().numpy()
    assert rel_err(syn, g['synth']) < 1e-5


@pytest.mark.parametrize('ti', range(4), ids=['bottleneck_s2', 'gln_s2', 'bottleneck_s4', 'gln_s4'])
@pytest.mark.parametrize('precision', ['fp32', 'tc_f16x3'])
def test_cuda_codec_vs_reference_run(ti, precision):
    """The CUDA codec, through the C ABI, against the output of the REFERENCE'S OWN graph code (nn_core_operator.py + the graph
    methods of neural_speech_coding_module.py, run on the TensorFlow stand-in of tests/golden/tf_shim.py; fixture
    tests/golden/reference_run_nn.npz).  Weights: the seeded stream of the generator, laid out through the library's own layer table
    -- a creation-order or shape disagreement with the reference graph shows up as garbage.  Floating code 1e-4; decoder 1e-4 on
    the reference run's own codes; hard codes identical wherever the floating codes are not within rounding of a mid-point."""
    import sys
    sys.path.insert(0, GOLD)
    try:
        import tf_shim
    finally:
        sys.path.remove(GOLD)
    from nsc_b200 import codec
    rt, st = [('bottleneck', (2,)), ('gln', (2,)), ('bottleneck', (2, 2)), ('gln', (2, 2))][ti]
    g = dict(np.load(os.path.join(GOLD, 'reference_run_nn.npz')))
    cfg = codec.CodecConfig(resnet_type=rt, the_strides=st, precision=precision)
    rng = np.random.RandomState(100 + ti)
    layers = []
    for L in codec.layer_table(cfg):          # the LIBRARY'S creation order and shapes
        shapes = ((L.k, L.cin, 1), (1, L.cin, L.cout), (L.cout,)) if L.separable else ((L.k, L.cin, L.cout), (L.cout,))
        layers.append(tf_shim.draw_layer(rng, shapes))
    flat = codec.pack_params_numpy(cfg, layers, -300.0, np.linspace(-1, 1, 32))
    gc = codec.NeuralCodec(cfg, torch.from_numpy(flat).to(DEV))
    x = cu(g['x'])
    tag = f"{rt}_{len(st)}_hard"
    enc = gc.encode(x)
    fl = enc['floating_code'].cpu().numpy()
    assert rel_err(fl, g[tag + '_floating']) < TOL
    code = enc['code'].cpu().numpy()
    near_mid = np.abs(np.abs(((g[tag + '_floating'] + 1) * 15.5) % 1.0 - 0.5)) < 1e-3      # within 1e-3 bin widths of a mid-point
    assert np.array_equal(code[~near_mid], g[tag + '_code'][~near_mid])
    out = gc.decode(cu(g[tag + '_code'])).cpu().numpy()
    assert rel_err(out, g[tag + '_out']) < TOL
    soft = gc.computational_graph_end2end_quan_on(x, True, 1.0)
    assert rel_err(soft['floating_code'].cpu().numpy(), g[f"{rt}_{len(st)}_soft_floating"]) < TOL


def test_cuda_losses_vs_reference_run():
    """mse_loss / mfcc_loss / quan_loss / entropy_coding_loss on the GPU against the values the reference's own
    loss_terms_and_measures.py produced on the TensorFlow stand-in (tests/golden/reference_run_nn.npz)."""
    from nsc_b200 import loss_terms_and_measures as lt
    g = dict(np.load(os.path.join(GOLD, 'reference_run_nn.npz')))
    dec, ori = cu(g['loss_dec']), cu(g['loss_ori'])
    assert rel_err(lt.mse_loss(dec, ori).cpu().numpy(), g['loss_mse']) < 1e-5
    assert rel_err(lt.mfcc_loss(dec, ori).cpu().numpy(), g['loss_mfcc']) < TOL
    soft = cu(g['loss_soft'])
    assert rel_err(lt.quan_loss(soft).cpu().numpy(), g['loss_quan']) < 1e-5
    assert abs(float(lt.entropy_coding_loss(soft)) - float(g['loss_ent'])) < 1e-4
    assert np.allclose([lt.entropy_to_bitrate(2.5, 2), lt.entropy_to_bitrate(2.5, 4)], g['bitrate'], rtol=1e-12)


@pytest.mark.parametrize('precision', ['fp32', 'tc_f16x3'])
def test_cuda_cq_feedforward_vs_reference_run(precision):
    """nsc_cq_forward against the collaborative-quantisation pass as the REFERENCE'S OWN CMRL.all_modules_feedforward_lpc built it
    (cmrl.py:770-839 run on the TensorFlow stand-in; tests/golden/reference_run_nn.npz): LSF codes bit-exact, quantised polynomial
    and residual 1e-5, and -- on the soft path, which has no discontinuity -- every codec output, the decoded sum and the synthesized
    frames end to end (2e-3: the alpha = -300 soft quantiser amplifies float32 rounding of the encoders)."""
    import sys
    sys.path.insert(0, GOLD)
    try:
        import tf_shim
    finally:
        sys.path.remove(GOLD)
    from nsc_b200 import codec
    g = dict(np.load(os.path.join(GOLD, 'reference_run_nn.npz')))
    cfg = codec.CodecConfig(resnet_type='bottleneck', precision=precision)
    rng = np.random.RandomState(400)
    gcs = []
    for _ in range(2):
        layers = [tf_shim.draw_layer(rng, ((L.k, L.cin, L.cout), (L.cout,))) for L in codec.layer_table(cfg)]
        gcs.append(codec.NeuralCodec(cfg, torch.from_numpy(codec.pack_params_numpy(cfg, layers, -300.0, np.linspace(-1, 1, 32))).to(DEV)))
    cm = codec.CMRL(gcs, res_scalar=2.0)
    x, lsf = cu(g['cq_x']), cu(g['cq_lsf'])
    soft = cm.feedforward_lpc(x, lsf, True, 1.0)
    hard = cm.feedforward_lpc(x, lsf, False, 1.0)
    torch.cuda.synchronize()
    for r, tag in ((soft, 'cq_soft'), (hard, 'cq_hard')):
        assert np.array_equal(r['lsf_idx'].cpu().numpy().astype(np.int64), g[tag + '_lsf_idx'])
        assert rel_err(r['poly'].cpu().numpy(), g[tag + '_poly']) < 1e-5
        assert rel_err(r['res_x'].cpu().numpy(), g[tag + '_res_x']) < 1e-5
    assert rel_err(soft['decoded'].cpu().numpy(), g['cq_soft_decoded']) < 2e-3
    assert rel_err(soft['synthesized'].cpu().numpy(), g['cq_soft_synth']) < 2e-3
    c = cm.all_modules_feedforward(x, True, 1.0, want_outs=True)
    assert rel_err(torch.stack(c['outs']).cpu().numpy(), g['cq_soft_plain_outs']) < 2e-3


@pytest.mark.parametrize('L,wide,dil', [(128, 100, 2), (384, 100, 2), (128, 50, 1), (512, 100, 2)])
def test_gated_block_surface_on_the_tensor_engine(L, wide, dil):
    """nn_core_operator.gated_bottleneck through nsc_gated_block_tc (tcgen05): dilation-2 gates on de-interleaved sub-images where
    a sub-image is a whole number of 128-row tiles (L = 512), else on the plain image with 16 halo rows (L = 128, 384); odd batch."""
    from nsc_b200 import nn_core_operator as nn
    ps = ref_nn.ParamStream(seed=L + wide + dil)
    x = np.random.RandomState(L).randn(3, L, wide).astype(np.float32)
    ref = ref_nn.gated_bottleneck(torch.from_numpy(x), wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, ps=ps).numpy()
    params = [tuple(cu(p) for p in t) for t in ps.params]
    got = nn.gated_bottleneck(cu(x), wide_layer=wide, narrow_layer=20, dilation_rate=dil, is_last_flat=False, params=params)
    assert nn.last_engine == 'tc'
    g, r = got.cpu().numpy(), ref
    assert rel_err(g, r) < 5e-5
    per = np.abs(g - r).reshape(3, -1).max(1) / np.abs(r).reshape(3, -1).max(1)
    assert per.max() < 5e-5


def test_cuda_blocks_vs_reference_run():
    """the_bottleneck / gated_bottleneck / gated_bottleneck_decoder / conv1d / conv1d_depth / change_channel on the GPU against the
    outputs of the reference's own nn_core_operator.py (tests/golden/reference_run_nn.npz)."""
    import sys
    sys.path.insert(0, GOLD)
    try:
        import tf_shim
    finally:
        sys.path.remove(GOLD)
    from nsc_b200 import nn_core_operator as nn
    g = dict(np.load(os.path.join(GOLD, 'reference_run_nn.npz')))
    xb = cu(g['block_x'])

    def params_like(layers, seed):
        rng = np.random.RandomState(seed)
        return [tuple(cu(a) for a in tf_shim.draw_layer(rng, tuple(l))) for l in layers]

    for name, fn, kw in (('the_bottleneck', nn.the_bottleneck, dict(wide_layer=100, narrow_layer=20, dilation_rate=2, is_last_flat=False)),
                         ('the_bottleneck_flat', nn.the_bottleneck, dict(wide_layer=100, narrow_layer=20, dilation_rate=1, is_last_flat=True)),
                         ('gated_bottleneck', nn.gated_bottleneck, dict(wide_layer=100, narrow_layer=20, dilation_rate=2, is_last_flat=False)),
                         ('gated_bottleneck_decoder', nn.gated_bottleneck_decoder, dict(wide_layer=100, narrow_layer=20, dilation_rate=1, is_last_flat=True))):
        layers = eval(str(g['block_' + name + '_layers']))
        y = fn(xb, params=params_like(layers, 200), **kw).cpu().numpy()
        assert rel_err(y, g['block_' + name]) < 5e-5, name
    for name, fn, kw, layers in (
            ('conv1d_s2', nn.conv1d, dict(num_filters=24, filter_size=9, strides=2, dilation_rate=1), [((9, 100, 24), (24,))]),
            ('conv1d_d3', nn.conv1d, dict(num_filters=8, filter_size=5, strides=1, dilation_rate=3, activation=None), [((5, 100, 8), (8,))]),
            ('conv1d_depth', nn.conv1d_depth, dict(num_filters=50, filter_size=9, activation=None), [((9, 100, 1), (1, 100, 50), (50,))]),
            ('change_channel', nn.change_channel, dict(the_channel=1, kernel_size=55, dilation_rate=7), [((55, 100, 1), (1,))])):
        y = fn(xb, params=params_like(layers, 300)[0], **kw).cpu().numpy()
        assert y.shape == g['op_' + name].shape and rel_err(y, g['op_' + name]) < 5e-5, name
