"""Algebra of the FOLDED narrow conv (nsc_b200/csrc/plane.cuh, fold2_weights_kernel in plane_conv.cu), in numpy: the k9 20 -> 20 conv of
a bottleneck block (nn_core_operator.py:64-68, SAME padding) equals a 48 -> 48 k5 conv on PAIRS of positions folded into the channel
axis -- dilation 1 directly, dilation 2 on each position parity separately, or as a block-diagonal k9 conv on the same folded image --
including the frame borders (the folded image's zero rows are the original's zero padding)."""
import numpy as np
import pytest

FC, NARROW = 24, 20          # channels per folded phase (20 + 4 zero), real channels


def same_conv(x, w, dil):
    """x (L, Cin), w (K, Cin, Cout): SAME-padded dilated conv, float64."""
    L, K = x.shape[0], w.shape[0]
    pad = (K - 1) * dil // 2
    xp = np.zeros((L + 2 * pad, x.shape[1]))
    xp[pad:pad + L] = x
    return sum(xp[t * dil:t * dil + L] @ w[t] for t in range(K))


def fold(x):
    """(L, 20) -> (L / 2, 48): row r = [position 2r | 4 zeros | position 2r + 1 | 4 zeros]"""
    L = x.shape[0]
    out = np.zeros((L // 2, 2 * FC))
    out[:, :NARROW] = x[0::2]
    out[:, FC:FC + NARROW] = x[1::2]
    return out


def unfold(y):
    L2 = y.shape[0]
    out = np.zeros((2 * L2, NARROW))
    out[0::2] = y[:, :NARROW]
    out[1::2] = y[:, FC:FC + NARROW]
    return out


def fold_weights(w, Kf):
    """The mapping of fold2_weights_kernel: Kf = 5 -> tap 2 s + ph' - ph; Kf = 9 -> block-diagonal, tap s."""
    out = np.zeros((Kf, 2 * FC, 2 * FC))
    for s in range(Kf):
        for phi in range(2):
            for pho in range(2):
                t = (s if phi == pho else -1) if Kf == 9 else 2 * s + phi - pho
                if 0 <= t <= 8:
                    out[s, phi * FC:phi * FC + NARROW, pho * FC:pho * FC + NARROW] = w[t]
    return out


@pytest.mark.parametrize('L', [256, 512])
def test_dilation_1_is_a_k5_conv_on_pairs(L):
    rng = np.random.RandomState(L)
    x, w = rng.randn(L, NARROW), rng.randn(9, NARROW, NARROW)
    want = same_conv(x, w, 1)
    got = unfold(same_conv(fold(x), fold_weights(w, 5), 1))
    assert np.allclose(got, want, rtol=0, atol=1e-12)
    # 18 of the 20 (super-tap, input phase, output phase) blocks carry a tap: the two outermost super-taps reach one way only
    wf = fold_weights(w, 5)
    blocks = [(s, a, b) for s in range(5) for a in range(2) for b in range(2) if np.any(wf[s, a * FC:(a + 1) * FC, b * FC:(b + 1) * FC])]
    assert len(blocks) == 18


def test_dilation_2_per_parity_and_block_diagonal():
    rng = np.random.RandomState(2)
    L = 512
    x, w = rng.randn(L, NARROW), rng.randn(9, NARROW, NARROW)
    want = same_conv(x, w, 2)
    # per position parity: each parity sub-frame is a dilation-1 problem of half the length, then folded as above
    got = np.zeros_like(want)
    for par in range(2):
        got[par::2] = unfold(same_conv(fold(x[par::2]), fold_weights(w, 5), 1))
    assert np.allclose(got, want, rtol=0, atol=1e-12)
    # block-diagonal k9 on the plain folded image (frames too short to split by parity): a position only meets its own phase
    got2 = unfold(same_conv(fold(x), fold_weights(w, 9), 1))
    assert np.allclose(got2, want, rtol=0, atol=1e-12)


def test_padding_channels_stay_zero():
    """Channels 20-23 of each phase carry zero weights in and out: whatever sits there (nothing is ever written there) cannot leak."""
    wf = fold_weights(np.ones((9, NARROW, NARROW)), 5)
    for ph in range(2):
        assert not wf[:, ph * FC + NARROW:(ph + 1) * FC, :].any() and not wf[:, :, ph * FC + NARROW:(ph + 1) * FC].any()
