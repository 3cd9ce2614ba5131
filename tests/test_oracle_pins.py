"""Pins the CPU oracle (oracle/) against everything that exists to pin it with:
third-party doc-string known answers, the reference's literal constants, closed-form identities and the
committed golden fixtures.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_codec, ref_loss, ref_lpc, ref_nn
from util import ar_frames, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
REF = '/root/reference'


# --- spectrum's published poly2lsf / lsf2poly example (spectrum.linear_prediction doc-strings; MATLAB's known answer)
def test_spectrum_known_answer():
    lsf = [0.7842, 1.5605, 1.8776, 1.8984, 2.3593]
    a = [1.0000, 0.6149, 0.9899, 0.0000, 0.0031, -0.0082]
    np.testing.assert_allclose(ref_lpc.lsf2poly(lsf), a, atol=2e-4)
    np.testing.assert_allclose(ref_lpc.poly2lsf(a), lsf, atol=2e-4)


def test_lsf_poly_roundtrip_order16():
    x = ar_frames(4, 1024, seed=3)
    lsf = ref_lpc.lpc_analysis_windows(x, 16)
    assert np.all(np.diff(lsf, axis=1) > 0) and lsf.min() > 0 and lsf.max() < np.pi
    for row in lsf:
        a = ref_lpc.lsf2poly(row)
        assert a[0] == 1.0
        np.testing.assert_allclose(ref_lpc.poly2lsf(a), row, atol=1e-9)


def test_lpc_normal_equations():
    """levinson_durbin solves the Toeplitz system audiolazy.lpc builds."""
    from scipy.linalg import solve_toeplitz
    x = ar_frames(1, 1024, seed=5)[0].astype(np.float64)
    r = ref_lpc.acorr(x, 16)
    a = ref_lpc.levinson_durbin(r, 16)
    np.testing.assert_allclose(a[1:], solve_toeplitz(r[:-1], -r[1:]), rtol=1e-8, atol=1e-10)


# --- the reference's literal constants
def test_lsf_codebook_table():
    from nsc_b200 import constants
    gold = np.load(os.path.join(GOLD, 'lsf_bins_f64.npy'))
    assert gold.shape == (256,)
    assert np.array_equal(np.asarray(constants.lpc_coeff_lsf_bins, dtype=np.float32), gold.astype(np.float32))
    assert len(np.unique(gold)) == 256 and not np.all(np.diff(gold) > 0)      # unique but NOT monotone
    assert abs(np.diff(np.sort(gold)).min() - 7.4e-5) < 1e-5                  # min gap 7.4e-5


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
def test_constants_against_reference_source():
    import re
    src = open(os.path.join(REF, 'constants.py')).read()
    from nsc_b200 import constants
    key = 'lpc_coeff_lsf_bins_256_signal_preprosessed_final_flat = ['
    i = src.index(key)
    vals = [float(v) for v in re.findall(r'[-0-9.]+', src[i + len(key):src.index(']', i)])]
    assert np.array_equal(np.float32(vals), np.float32(constants.lpc_coeff_lsf_bins))
    for name in ('init_alpha', 'frame_length', 'overlap_each_side', 'sample_rate', 'empha_filter_coeff',
                 'lpc_perceptual_weighting_coeff', 'conv_mu'):
        m = re.search(r'^%s\s*=\s*([-0-9.]+)' % name, src, re.M)
        assert m and float(m.group(1)) == float(getattr(constants, name)), name


def test_residual_window_sum():
    """the overlap-added sub-frame windows do NOT sum to 1 (min 0.98764)."""
    w = ref_lpc.residual_windows()
    tot = np.zeros(512)
    for s in range(7):
        tot[s * 64:s * 64 + 128] += w[s]
    assert abs(tot.min() - 0.98764) < 1e-5 and tot.max() <= 1.0 + 1e-12
    assert tot[:64].min() == 1.0 and tot[-64:].min() == 1.0


# --- closed-form identities
def test_same_padding_table():
    """SURVEY.md 3.2 table."""
    assert ref_nn.same_padding(512, 55, 1, 1) == (512, 27, 27)
    assert ref_nn.same_padding(512, 9, 1, 1) == (512, 4, 4)
    assert ref_nn.same_padding(512, 9, 2, 1) == (512, 8, 8)
    assert ref_nn.same_padding(512, 9, 1, 2) == (256, 3, 4)
    assert ref_nn.same_padding(256, 15, 2, 1) == (256, 14, 14)


def test_param_counts():
    """458,052 conv parameters (enc 260,161 / dec 197,891); gln 350,152."""
    c = ref_codec.OracleCodec(ref_codec.OracleCodecCfg())
    sizes = [sum(int(np.prod(p.shape)) for p in t) for t in c.conv_params]
    assert len(sizes) == 29 and sum(sizes) == 458052 and sum(sizes[:15]) == 260161 and sum(sizes[15:]) == 197891
    g = ref_codec.OracleCodec(ref_codec.OracleCodecCfg(resnet_type='gln'))
    assert sum(int(np.prod(p.shape)) for t in g.conv_params for p in t) == 350152


def test_analysis_synthesis_roundtrip():
    """1/A(z) undoes A(z) when both start from zero state over the whole frame."""
    x = ar_frames(3, 512, seed=9).astype(np.float64)
    lsf = ref_lpc.lpc_analysis_windows(ar_frames(3, 1024, seed=10), 16)
    poly = ref_lpc.lsf2poly_after_quan(lsf.astype(np.float32), 16)
    for i in range(3):
        e = ref_lpc.fir_zero_state(poly[i].astype(np.float64), x[i])
        y = ref_lpc.lpc_synthesizer_tr(poly[i:i + 1], e[None, :].astype(np.float32))
        assert rel_err(y[0], x[i]) < 1e-4


def test_quantizer_tie_rule_and_hard_value():
    """x = 0 with the symmetric 32-bin init ties bins 15/16 -> 15 (tf.nn.top_k: lowest index)."""
    bins = np.linspace(-1, 1, 32).astype(np.float32)
    x = torch.zeros(1, 4, 1)
    idx = ref_nn.quantizer_indices(x, -300.0, bins)
    assert idx.tolist() == [[15, 15, 15, 15]]
    soft, code = ref_nn.scalar_softmax_quantization(x, -300.0, bins, 1.0, False, 4, 32)
    assert np.array_equal(code.numpy().ravel(), np.full(4, bins[15]))
    # softmax arg-max == logit arg-max on every fixture input
    g = np.load(os.path.join(GOLD, 'quantizer.npz'))
    for xs, b in ((g['x32'], bins), (g['xl'], np.load(os.path.join(GOLD, 'lsf_bins_f64.npy')).astype(np.float32))):
        a = ref_nn.quantizer_indices(torch.from_numpy(xs), -300.0, b)
        l = ref_nn.quantizer_indices_from_logits(torch.from_numpy(xs), -300.0, b)
        assert torch.equal(a, l)


def test_pixel_shuffle_definition():
    """out[b, s*l + r, c] = in[b, l, s*c + r] (nscm.py:158-167)."""
    x = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(2, 3, 4)
    y = ref_codec.OracleCodec._up_sampling_mod_helper(x, 2)
    assert y.shape == (2, 6, 2)
    for l in range(3):
        for r in range(2):
            for c in range(2):
                assert y[1, 2 * l + r, c] == x[1, l, 2 * c + r]


def test_mel_matrix_properties_and_parseval():
    for n in ref_loss.MEL_BANKS:
        m = ref_loss.linear_to_mel_weight_matrix(n, 257, 16000, 0.0, 8000.0)
        assert m.shape == (257, n) and np.all(m[0] == 0) and m.min() >= 0 and m.max() <= 1
        # interior bins: neighbouring triangles form a partition of unity
        s = m.sum(axis=1)
        lo, hi = np.argmax(m[:, 0]), np.argmax(m[:, -1])
        np.testing.assert_allclose(s[lo + 1:hi], 1.0, atol=1e-5)
    x = torch.from_numpy(ar_frames(2, 512, seed=2))
    st, mag = ref_loss.tf_stft(x)
    e_t = (x.double() ** 2).sum(-1)
    p = (st.abs().double() ** 2)
    e_f = (p[:, 0] + p[:, -1] + 2 * p[:, 1:-1].sum(-1)) / 512
    np.testing.assert_allclose(e_f.numpy(), e_t.numpy(), rtol=1e-5)


def test_loss_zero_distance():
    x = torch.from_numpy(ar_frames(2, 512, seed=4))
    np.testing.assert_allclose(ref_loss.mse_loss(x, x).numpy(), np.sqrt(1e-7), rtol=1e-6)
    np.testing.assert_allclose(ref_loss.mfcc_loss(x, x).numpy(), np.sqrt(1e-7), rtol=1e-6)
    soft = torch.zeros(2, 8, 4); soft[..., 1] = 1.0
    assert abs(float(ref_loss.entropy_coding_loss(soft))) < 1e-6
    uni = torch.full((2, 8, 4), 0.25)
    assert abs(float(ref_loss.entropy_coding_loss(uni)) - 2.0) < 1e-5
    np.testing.assert_allclose(ref_loss.quan_loss(uni).numpy(), 2.0, rtol=1e-6)


def test_cascade_algebra():
    """decoded = sum_i out_i and codec i sees res_scalar * (x - sum_{j<i} out_j) (cmrl.py:513-543)."""
    cfg = ref_codec.OracleCodecCfg()
    codecs = [ref_codec.OracleCodec(cfg, seed=1), ref_codec.OracleCodec(cfg, seed=2)]
    x = torch.from_numpy(ar_frames(1, 512, seed=6, std=0.3))[:, :, None]
    dec, outs, per = ref_codec.cascade_forward(codecs, x, False, 1.0, res_scalar=2.0)
    o0 = codecs[0].forward(x, False, 1.0)['out']
    o1 = codecs[1].forward(2.0 * (x - o0.unsqueeze(2)), False, 1.0)['out'] / 2.0
    assert torch.allclose(dec, o0 + o1, atol=1e-6)


def test_tf1_adam_step():
    th, m, v = ref_codec.tf1_adam_step(np.array([1.0]), np.array([0.5]), np.zeros(1), np.zeros(1), 1, 0.01)
    lr_t = 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9)
    np.testing.assert_allclose(th, 1.0 - lr_t * 0.05 / (np.sqrt(0.001 * 0.25) + 1e-8))


# --- committed golden fixtures (generated by tests/golden/make_golden.py)
def test_golden_lpc():
    g = np.load(os.path.join(GOLD, 'lpc.npz'))
    np.testing.assert_allclose(ref_lpc.lpc_analysis_windows(g['windows'], 16), g['lsf'], atol=1e-10)
    poly = ref_lpc.lsf2poly_after_quan(g['lsf'].astype(np.float32), 16)
    np.testing.assert_allclose(poly, g['poly'], atol=1e-6)
    res = ref_lpc.lpc_analysis_get_residual(g['frames'][:, :, None], g['poly'])
    np.testing.assert_allclose(res, g['res'], atol=1e-6)
    np.testing.assert_allclose(ref_lpc.lpc_synthesizer_tr(g['poly'], g['res']), g['syn'], atol=1e-5)


def test_golden_losses_and_quantizer():
    g = np.load(os.path.join(GOLD, 'losses.npz'))
    np.testing.assert_allclose(ref_loss.mse_loss(torch.from_numpy(g['dec']), torch.from_numpy(g['ori'])).numpy(),
                               g['time_loss'], rtol=1e-5)
    np.testing.assert_allclose(ref_loss.mfcc_loss(torch.from_numpy(g['dec']), torch.from_numpy(g['ori'])).numpy(),
                               g['freq_loss'], rtol=1e-4)
    q = np.load(os.path.join(GOLD, 'quantizer.npz'))
    bins = np.linspace(-1, 1, 32).astype(np.float32)
    assert np.array_equal(ref_nn.quantizer_indices(torch.from_numpy(q['x32']), -300.0, bins).numpy(), q['idx32'])
    assert np.array_equal(bins[q['idx32']], q['code32'][:, :, 0])


def test_golden_codec():
    g = np.load(os.path.join(GOLD, 'codec.npz'))
    for name, rt, st in [('bn2', 'bottleneck', (2,)), ('gln2', 'gln', (2,))]:
        oc = ref_codec.OracleCodec(ref_codec.OracleCodecCfg(resnet_type=rt, strides=st), seed=3)
        r = oc.forward(torch.from_numpy(g[name + '_x'])[:, :, None], False, 1.0)
        assert rel_err(r['out'].numpy(), g[name + '_out']) < 1e-4
        assert rel_err(r['floating_code'].numpy()[:, :, 0], g[name + '_floating']) < 1e-4


def test_fp32_vs_fp64_oracle_noise_floor():
    """How far the fp32 restatement sits from fp64 truth -- the floor under every 1e-4 claim."""
    oc = ref_codec.OracleCodec(ref_codec.OracleCodecCfg(), seed=3)
    x = torch.from_numpy(ar_frames(2, 512, seed=31, std=0.3))[:, :, None]
    oc.ps._cursor = 0
    f32 = oc.encoder(x)
    oc.ps._cursor = 0
    f64 = oc.encoder(x.double())
    assert rel_err(f32.numpy(), f64.numpy()) < 2e-5


# ---------------------------------------------------------------------------------------------- framing
def test_framing_oracle_counts_windows_and_filters():
    from oracle import ref_framing as rf
    # frame count of utilities.py:26: len(range(0, T - 512, 480))
    for T, n in ((512, 0), (513, 1), (992, 1), (993, 2), (48000, 99)):
        assert rf.utterance_to_segment(np.zeros(T), True).shape == (n, 512)
    the_w, first_w, last_w = rf.windows()
    assert the_w.shape == first_w.shape == last_w.shape == (512,)
    assert the_w[0] == 0.0 and the_w[31] == 1.0 and the_w[480] == 1.0 and the_w[511] == 0.0     # hanning(63) halves
    assert np.all(first_w[:480] == 1) and np.all(last_w[32:] == 1)
    # overlap-add of the middle windows: hanning(63)[k] + hanning(63)[31 + k] = 1 only at the ends -> seam error < 6 %
    seam = the_w[480:] + the_w[:32]
    assert abs(seam - 1).max() < 0.06
    # the flatten quirk of lpc_analysis_at_test: window 1 starts at frame 1 sample 0 = utterance sample 480
    x = np.arange(5000, dtype=np.float64)
    w = rf.lpc_windows_at_test(rf.utterance_to_segment(x, True))
    assert w.shape == (len(range(0, 5000 - 512, 480)) - 2, 1024)
    assert w[1, 0] == 480 and w[0, 512] == 480 and w[0, 511] == 511
    # filters: the asymmetric numerator taps (0.989502 vs 0.989592, lpc_utilities.py:10) leave a DC gain of
    # 9e-5 / 2.44e-4 = 0.3689 instead of 0 -- a reference quirk that is reproduced, not repaired
    hp = rf.highpass_filter(np.ones(20000))
    assert abs(hp[-1] - 9e-5 / 2.44e-4) < 1e-6
    y = rf.de_empha_filter(rf.empha_filter(x))
    assert np.allclose(y, x)
    idx = np.array([[1, 2, 3, 31, 0, 17, 8, 9]], dtype=np.uint8)
    pk = rf.pack_bits(idx, 5)
    assert pk.shape == (1, 5) and pk[0, 0] == (1 | (2 << 5)) & 0xFF
