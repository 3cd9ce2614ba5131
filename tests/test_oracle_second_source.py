"""Second-source pins of the oracle (CPU).  The reference ships no tests or golden vectors and TensorFlow / audiolazy / spectrum
cannot be installed here, so the oracle's [LIB] restatements are checked against INDEPENDENT implementations of the same published
algorithms that do exist in this image: transformers.audio_utils (mel filterbank with mel-space triangles, written to reproduce
tf.signal.linear_to_mel_weight_matrix), torchaudio (filterbank support; lfilter), scipy.signal (lfilter, lfiltic-free zero state),
numpy.linalg (dense Toeplitz solve, companion-matrix roots), torch.nn.functional.conv1d(padding='same'), numpy.fft.
Parity stays "unpinned by the reference itself" (oracle/__init__.py); these pins remove the failure mode "the oracle agrees only
with itself"."""
import warnings

import numpy as np
import pytest
import torch

from oracle import ref_codec, ref_loss, ref_lpc, ref_nn
from util import ar_frames, rel_err


def test_mel_matrix_vs_transformers_mel_space_triangles():
    """tf.signal.linear_to_mel_weight_matrix (loss_terms_and_measures.py:130-148): HTK mel scale, triangles linear IN MEL, DC row zero."""
    from transformers.audio_utils import mel_filter_bank
    for n in (8, 16, 32, 128):                     # constants.py:28 selected_ind
        ours = ref_loss.linear_to_mel_weight_matrix(n, 257, 16000, 0.0, 8000.0)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')        # (128 filters over 257 bins leave some filters empty -- in TensorFlow too)
            theirs = mel_filter_bank(num_frequency_bins=257, num_mel_filters=n, min_frequency=0.0, max_frequency=8000.0,
                                     sampling_rate=16000, norm=None, mel_scale='htk', triangularize_in_mel_space=True)
        assert ours.shape == theirs.shape == (257, n)
        assert np.abs(ours - theirs).max() < 1e-6
        assert np.all(ours[0] == 0.0)


def test_mel_matrix_support_vs_torchaudio():
    """torchaudio builds its triangles linear in HERTZ (a different published convention), so values differ by up to a few 1e-2; the band
    edges -- hence each filter's support -- are the same HTK mel grid."""
    import torchaudio
    for n in (8, 16, 32):
        ours = ref_loss.linear_to_mel_weight_matrix(n, 257, 16000, 0.0, 8000.0)
        theirs = torchaudio.functional.melscale_fbanks(257, 0.0, 8000.0, n, 16000, norm=None, mel_scale='htk').numpy()
        assert np.array_equal(ours > 0, theirs > 0)
        assert np.abs(ours - theirs).max() < 5e-2
        assert np.abs(ours.argmax(0) - theirs.argmax(0)).max() <= 1     # peaks on the same bin (or its neighbour: sampled triangles)


def test_iir_and_fir_vs_torchaudio_and_scipy():
    """audiolazy's ZFilter call = direct-form difference equation from zero state [LIB]; checked against torchaudio's and scipy's."""
    import torchaudio
    from scipy.signal import lfilter
    rng = np.random.RandomState(3)
    x = rng.randn(4, 600)
    poly = np.stack([ref_lpc.lpc_autocor(f, 16) for f in ar_frames(4, 600, seed=5).astype(np.float64)])
    for i in range(4):
        a = poly[i]
        y = ref_lpc.iir_zero_state([1.0], a, x[i])
        y_sp = lfilter([1.0], a, x[i])
        b_pad = np.zeros_like(a); b_pad[0] = 1.0
        y_ta = torchaudio.functional.lfilter(torch.from_numpy(x[i]), torch.from_numpy(a), torch.from_numpy(b_pad), clamp=False).numpy()
        assert rel_err(y, y_sp) < 1e-12
        assert rel_err(y, y_ta) < 1e-9
        e = ref_lpc.fir_zero_state(a, x[i])
        assert rel_err(e, np.convolve(x[i], a)[:x.shape[1]]) < 1e-12
    # the utterance filters (lpc_utilities.py:8-11): pre-emphasis then its inverse is the identity
    s = rng.randn(2000)
    assert rel_err(ref_lpc.de_empha_filter(ref_lpc.empha_filter(s)), s) < 1e-10


def test_levinson_vs_dense_solve_and_roots():
    """audiolazy.lpc 'autocor' [LIB] = Toeplitz normal equations; numpy's dense LU is a different algorithm for the same system."""
    for f in ar_frames(6, 1024, seed=9).astype(np.float64):
        r = ref_lpc.acorr(f, 16)
        a = ref_lpc.levinson_durbin(r, 16)
        R = np.array([[r[abs(i - j)] for j in range(16)] for i in range(16)])
        a_dense = np.concatenate(([1.0], np.linalg.solve(R, -r[1:17])))
        assert rel_err(a, a_dense) < 1e-8
        # minimum phase (autocorrelation method) and LSF <-> poly round trip through two different root finders
        assert np.abs(np.roots(a)).max() < 1.0
        lsf = ref_lpc.poly2lsf(a)
        assert np.all(np.diff(lsf) > 0) and lsf[0] > 0 and lsf[-1] < np.pi
        assert rel_err(ref_lpc.lsf2poly(lsf), a) < 1e-8
        # P / Q interlacing evaluated directly on the unit circle: A(e^{jw}) +- e^{-j17w} conj(A) vanishes at the LSFs
        w = lsf
        z = np.exp(-1j * np.outer(w, np.arange(17)))
        A = z @ a
        P = A + np.exp(-1j * 17 * w) * np.conj(A)
        Q = A - np.exp(-1j * 17 * w) * np.conj(A)
        assert np.minimum(np.abs(P), np.abs(Q)).max() < 1e-8


def test_lsf2poly_literal_complex64_switch():
    """spectrum.lsf2poly on the float32 row py_func hands it runs numpy.poly in complex64 [LIB]; the oracle's default (and the CUDA
    kernel) compute in float64.  MEASURED here: the literal path sits 4e-5 (median) to 1.3e-4 (worst frame of 300) away from the
    float64 result, relative to max|a| -- i.e. the reference's own rounding noise on the LPC polynomial is of the order of the 1e-4
    parity budget.  Parity of `poly` is therefore asserted against the float64 result (GPU within 1e-6, tests/test_gpu_parity.py),
    and this test pins the size of the reference's noise so that nobody mistakes it for a kernel error."""
    es = []
    for f in ar_frames(64, 1024, seed=11).astype(np.float64):
        lsf32 = ref_lpc.poly2lsf(ref_lpc.lpc_autocor(f, 16)).astype(np.float32)
        es.append(rel_err(ref_lpc.lsf2poly(lsf32, literal_dtype=True), ref_lpc.lsf2poly(lsf32)))
    es = np.array(es)
    assert es.min() > 0.0                          # the switch really changes the arithmetic
    assert np.median(es) < 1e-4 and es.max() < 4e-4


@pytest.mark.parametrize('k,d,cin,cout', [(9, 1, 100, 20), (9, 2, 20, 20), (55, 1, 1, 100), (15, 2, 20, 20), (1, 1, 100, 20)])
def test_same_conv_vs_torch_same_padding(k, d, cin, cout):
    """tf.layers.conv1d SAME at stride 1 (odd effective kernel): symmetric zero padding = torch's padding='same'."""
    rng = np.random.RandomState(k + d)
    x = torch.from_numpy(rng.randn(2, 96, cin).astype(np.float32))
    w = (rng.randn(k, cin, cout) / np.sqrt(k * cin)).astype(np.float32)
    b = rng.randn(cout).astype(np.float32)
    ours = ref_nn.conv1d_explicit(x, w, b, d, 1, None)
    theirs = torch.nn.functional.conv1d(x.transpose(1, 2), torch.from_numpy(w).permute(2, 1, 0).contiguous(), torch.from_numpy(b),
                                        padding='same', dilation=d).transpose(1, 2)
    assert rel_err(ours.numpy(), theirs.numpy()) < 1e-6


def test_strided_same_conv_vs_explicit_loop():
    """stride 2, k 9 (the down-sampling conv, nscm.py:152-156): TF pads (3, 4); a plain Python loop is the second source."""
    rng = np.random.RandomState(2)
    x = rng.randn(1, 32, 3).astype(np.float32)
    w = rng.randn(9, 3, 2).astype(np.float32)
    b = rng.randn(2).astype(np.float32)
    ours = ref_nn.conv1d_explicit(torch.from_numpy(x), w, b, 1, 2, None).numpy()
    xp = np.pad(x[0].astype(np.float64), [(3, 4), (0, 0)])
    ref = np.zeros((16, 2))
    for o in range(16):
        for t in range(9):
            ref[o] += xp[2 * o + t] @ w[t].astype(np.float64)
    ref += b
    assert ours.shape == (1, 16, 2)
    assert rel_err(ours[0], ref) < 1e-6


def test_rfft_and_loss_vs_numpy():
    """tf.signal.stft with frame_step = frame_length = 512 and no window is one rFFT-512 per frame (loss_terms_and_measures.py:178-183)."""
    x = ar_frames(3, 512, seed=21)
    st, mag = ref_loss.tf_stft(torch.from_numpy(x))
    ref = np.fft.rfft(x.astype(np.float64), axis=-1)
    assert rel_err(st.numpy(), ref) < 1e-5
    assert rel_err(mag.numpy(), np.sqrt(np.abs(ref) ** 2 + 1e-7)) < 1e-5
    y = x + 0.01 * ar_frames(3, 512, seed=22)
    t = ref_loss.mse_loss(torch.from_numpy(y), torch.from_numpy(x)).numpy()
    assert rel_err(t, np.sqrt(((y - x).astype(np.float64) ** 2).mean(-1) + 1e-7)) < 1e-5


def test_quantizer_vs_nearest_bin():
    """At alpha = -300 the soft-to-hard quantiser's hard code is the nearest bin (first index on exact ties) -- checked against a
    brute-force nearest-neighbour search in float64 on codes away from the mid-points."""
    bins = np.linspace(-1, 1, 32).astype(np.float32)
    rng = np.random.RandomState(4)
    x = rng.uniform(-1.1, 1.1, size=(4, 256)).astype(np.float32)
    mids = (bins[:-1].astype(np.float64) + bins[1:]) / 2
    keep = np.abs(x[..., None].astype(np.float64) - mids).min(-1) > 1e-4
    idx = ref_nn.quantizer_indices(torch.from_numpy(x)[:, :, None], -300.0, bins).numpy()
    nn_idx = np.abs(x[..., None].astype(np.float64) - bins.astype(np.float64)).argmin(-1)
    assert np.array_equal(idx[keep], nn_idx[keep])
