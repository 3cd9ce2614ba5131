"""Framing, overlap-add, utterance-level filters and code packing against the oracle,
including the end-to-end wav -> frames/LSFs -> codes -> frames -> wav pipeline of cmrl.py:666-737 with the codec replaced by
the identity (size-independent property: overlap-add of the analysis frames reproduces the signal where the windows sum to 1)."""
import numpy as np
import pytest
import torch

from oracle import ref_framing as rf
from util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def cu(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(DEV)


def _speech(T, seed=0):
    rng = np.random.RandomState(seed)
    from scipy.signal import lfilter
    x = lfilter([1.0], [1.0, -1.7, 0.81], rng.randn(T + 100))[100:]
    return (x / (x.std() if T > 8 else 1.0)).astype(np.float32)


@pytest.mark.parametrize('T', [513, 992, 993, 4000, 48000, 160123])
@pytest.mark.parametrize('post', [True, False])
def test_utterance_to_segment(T, post):
    from nsc_b200 import utilities as ut
    x = _speech(T, seed=T)
    want = rf.utterance_to_segment(x.astype(np.float64), post)
    got = ut.utterance_to_segment(cu(x), post).cpu().numpy()
    assert got.shape == want.shape
    if post:
        assert np.array_equal(got, want.astype(np.float32))     # pure copies: bit exact
    else:
        assert rel_err(got, want) < 1e-6


def test_short_and_empty_utterances():
    from nsc_b200 import utilities as ut
    for T in (0, 100, 512):
        assert ut.utterance_to_segment(torch.zeros(T, device=DEV), True).shape == (0, 512)
        assert rf.utterance_to_segment(np.zeros(T), True).shape == (0, 512)
    assert ut.overlap_add(torch.zeros((0, 512), device=DEV)).numel() == 32      # 512 + 480 * (0 - 1), like the reference's array


@pytest.mark.parametrize('T', [2000, 48000, 99999])
def test_lpc_windows_reproduce_the_flatten_quirk(T):
    from nsc_b200 import utilities as ut
    x = _speech(T, seed=1)
    want = rf.lpc_windows_at_test(rf.utterance_to_segment(x.astype(np.float64), True))
    got = ut.lpc_windows_at_test(cu(x)).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want.astype(np.float32))
    if want.shape[0] > 1:      # window 1 starts at flat index 512 = frame 1 sample 0 = utterance sample 480 (not 512)
        assert got[1, 0] == x[480]


@pytest.mark.parametrize('N', [1, 2, 3, 50, 333])
def test_overlap_add_both_loops(N):
    from nsc_b200 import utilities as ut
    rng = np.random.RandomState(N)
    fr = rng.randn(N, 512).astype(np.float32)
    # non-LPC loop: every frame, last frame gets last_window
    want = rf.overlap_add(fr.astype(np.float64), N, N, 512 + 480 * (N - 1))
    got = ut.overlap_add(cu(fr)).cpu().numpy()
    assert rel_err(got, want) < 1e-6
    # LPC loop: seg_amount = N, only N - 2 frames are accumulated into an array sized by another count
    if N >= 3:
        want = rf.overlap_add(fr.astype(np.float64), N, N - 2, 512 + 480 * (N - 1))
        got = ut.overlap_add(cu(fr), seg_amount=N, n_used=N - 2, out_len=512 + 480 * (N - 1)).cpu().numpy()
        assert rel_err(got, want) < 1e-6
    for j in (0, N // 2, N - 1):
        hp = ut.hann_process(cu(fr[j]), j, N).cpu().numpy()
        assert rel_err(hp, rf.hann_process(fr[j].astype(np.float64), j, N)) < 1e-6


def test_frame_then_overlap_add_is_identity_inside():
    from nsc_b200 import utilities as ut
    x = _speech(48000, seed=3)
    seg = ut.utterance_to_segment(cu(x), True)
    y = ut.overlap_add(seg).cpu().numpy()
    w, _, _ = rf.windows()
    # the trapezoid windows overlap-add to exactly 1 except in the seams of hanning(63) vs hanning(64) halves
    n = len(y)
    assert np.abs(y[32:n - 32] - x[32:n - 32]).max() < 0.06 * np.abs(x).max()
    assert np.allclose(y[:480 - 32], x[:480 - 32], atol=1e-6)


@pytest.mark.parametrize('T', [1, 2, 255, 256, 257, 5000, 160000])
def test_utterance_filters(T):
    from nsc_b200 import utilities as ut
    x = _speech(T, seed=T + 7)
    hp = ut.highpass_filter(cu(x), out_f64=True)
    assert rel_err(hp.cpu().numpy(), rf.highpass_filter(x)) < 1e-9
    pre = ut.empha_filter(ut.highpass_filter(cu(x)))
    want = rf.empha_filter(rf.highpass_filter(x).astype(np.float32))
    assert rel_err(pre.cpu().numpy(), want) < 1e-6
    de = ut.de_empha_filter(cu(x), out_f64=True)
    assert rel_err(de.cpu().numpy(), rf.de_empha_filter(x)) < 1e-9
    # de-emphasis inverts pre-emphasis (float64 round trip)
    rt = ut.de_empha_filter(ut.empha_filter(cu(x))).cpu().numpy()
    assert rel_err(rt, x) < 1e-5


def test_filters_batched_signals():
    from nsc_b200 import utilities as ut
    x = np.stack([_speech(7000, seed=s) for s in range(5)])
    got = ut.highpass_filter(cu(x)).cpu().numpy()
    for i in range(5):
        assert rel_err(got[i], rf.highpass_filter(x[i])) < 1e-6


def test_std_normalisation():
    from nsc_b200 import utilities as ut
    x = 3.7 * _speech(16000, seed=9)
    y, s = ut.load_sig_lpc(cu(x))
    assert abs(float(s) - np.std(x)) < 1e-4 * np.std(x)
    assert rel_err(y.cpu().numpy(), x / np.std(x)) < 1e-5


@pytest.mark.parametrize('L,nb', [(256, 32), (16, 256), (128, 64), (256, 2), (50, 33), (7, 5)])
def test_code_packing_round_trip(L, nb):
    from nsc_b200 import bitstream as bs
    rng = np.random.RandomState(L + nb)
    idx = rng.randint(0, nb, size=(301, L)).astype(np.uint8)
    packed = bs.pack_codes(cu(idx, torch.uint8), nb)
    assert np.array_equal(packed.cpu().numpy(), rf.pack_bits(idx, bs.bits_for(nb)))
    back = bs.unpack_codes(packed, L, nb)
    assert np.array_equal(back.cpu().numpy(), idx)


def test_frame_records():
    from nsc_b200 import bitstream as bs
    rng = np.random.RandomState(0)
    lsf = rng.randint(0, 256, size=(64, 16)).astype(np.uint8)
    c0 = rng.randint(0, 32, size=(64, 256)).astype(np.uint8)
    c1 = rng.randint(0, 32, size=(64, 256)).astype(np.uint8)
    rec = bs.pack_frames(cu(lsf, torch.uint8), [cu(c0, torch.uint8), cu(c1, torch.uint8)], [32, 32])
    assert rec.shape == (64, 16 + 160 + 160)
    l2, (d0, d1) = bs.unpack_frames(rec, [256, 256], [32, 32])
    assert np.array_equal(l2.cpu().numpy(), lsf) and np.array_equal(d0.cpu().numpy(), c0) and np.array_equal(d1.cpu().numpy(), c1)
    assert abs(bs.record_bitrate_kbps([256, 256], [32, 32]) - (128 + 2560) * 16000 / 480 / 1000) < 1e-9
