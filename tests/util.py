"""Shared helpers of the test-suite: seeded synthetic audio/weights and error metrics."""
import numpy as np


def ar_frames(n_frames, length=512, seed=1234, std=1.0):
    """Gaussian noise through a stable AR(2) colouring (poles 0.9 at +-0.1*pi) -- speech-like, well-conditioned
    LPC."""
    from scipy.signal import lfilter
    rng = np.random.RandomState(seed)
    r, th = 0.9, 0.1 * np.pi
    a = [1.0, -2 * r * np.cos(th), r * r]
    x = lfilter([1.0], a, rng.randn(n_frames, length + 256), axis=-1)[:, 256:]
    x = x / x.std() * std
    return x.astype(np.float32)


def rel_err(a, b):
    """max|a-b| / max|b| (the 1e-4 criterion of BASELINE.json:north_star)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


def quantizer_edge_codes(bins):
    """Adversarial floating codes for a sorted-or-not codebook: exact bin hits, exact mid-points, +-0, |x|>1."""
    b = np.asarray(bins, dtype=np.float32)
    sb = np.sort(b)
    mids = ((sb[:-1].astype(np.float64) + sb[1:].astype(np.float64)) / 2).astype(np.float32)
    extra = np.array([0.0, -0.0, 1.5, -1.5, 1.0, -1.0, 0.999999, -0.999999], dtype=np.float32)
    return np.concatenate([b, mids, np.nextafter(mids, np.float32(4)), np.nextafter(mids, np.float32(-4)), extra])
