"""Oracle restatement of /root/reference/lpc_utilities.py (numpy float64).  TEST INFRASTRUCTURE ONLY.

The reference calls audiolazy (``lpc``, ``ZFilter``) and spectrum (``poly2lsf``, ``lsf2poly``); neither is
installed here, so their published algorithms are restated ([LIB] marks) -- see oracle/__init__.py for the
parity status.  All arithmetic is float64 exactly as the Python originals (Python floats), with the float32
cast at the same places as the reference (`.astype(np.float32)` at lpc_utilities.py:33, :77, :156).
"""
from __future__ import annotations

import numpy as np

FRAME_LENGTH = 512          # constants.py:25
EMPHA_COEFF = -0.68         # constants.py:64
HIGHPASS_B = (0.989502, -1.979004, 0.989592)   # lpc_utilities.py:10 (taps are asymmetric in the reference)
HIGHPASS_A = (1.0, -1.978882, 0.979126)        # lpc_utilities.py:11


# ----------------------------------------------------------------------------------------------
# direct-form filters with zero initial state == audiolazy ZFilter.__call__ on a fresh call [LIB]
# ----------------------------------------------------------------------------------------------
def fir_zero_state(b, x):
    b = np.asarray(b, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    y = np.zeros_like(x)
    for k in range(len(b)):
        if k == 0:
            y += b[0] * x
        else:
            y[k:] += b[k] * x[:-k]
    return y


def iir_zero_state(b, a, x):
    """y[n] = (sum_k b_k x[n-k] - sum_{k>=1} a_k y[n-k]) / a_0."""
    from scipy.signal import lfilter
    return lfilter(np.asarray(b, np.float64), np.asarray(a, np.float64), np.asarray(x, np.float64))


def empha_filter(x):
    """lpc_utilities.py:8: 1 - 0.68 z^-1."""
    return fir_zero_state([1.0, EMPHA_COEFF], x)


def de_empha_filter(x):
    """cmrl.py:735: 1 / (1 - 0.68 z^-1)."""
    return iir_zero_state([1.0], [1.0, EMPHA_COEFF], x)


def highpass_filter(x):
    """lpc_utilities.py:10-11."""
    return iir_zero_state(HIGHPASS_B, HIGHPASS_A, x)


# ----------------------------------------------------------------------------------------------
# audiolazy.lpc(blk, order), autocorrelation method [LIB]
# ----------------------------------------------------------------------------------------------
def acorr(blk, max_lag):
    """audiolazy.acorr: un-normalised r[tau] = sum_n blk[n] blk[n+tau], tau = 0..max_lag [LIB]."""
    blk = np.asarray(blk, dtype=np.float64)
    n = len(blk)
    return np.array([np.dot(blk[: n - tau], blk[tau:]) for tau in range(max_lag + 1)])


def levinson_durbin(r, order):
    """Solves the Toeplitz normal equations; returns a[0..order] with a[0] = 1 [LIB]."""
    a = np.zeros(order + 1, dtype=np.float64)
    a[0] = 1.0
    err = r[0]
    for m in range(1, order + 1):
        acc = r[m] + np.dot(a[1:m], r[m - 1:0:-1])
        k = -acc / err
        a_prev = a.copy()
        for i in range(1, m):
            a[i] = a_prev[i] + k * a_prev[m - i]
        a[m] = k
        err *= (1.0 - k * k)
    return a


def lpc_autocor(blk, order):
    return levinson_durbin(acorr(blk, order), order)


# ----------------------------------------------------------------------------------------------
# spectrum.poly2lsf / lsf2poly [LIB]
# ----------------------------------------------------------------------------------------------
def poly2lsf(a):
    a = np.array(a, dtype=np.float64)
    if a[0] != 1:
        a = a / a[0]
    if np.max(np.abs(np.roots(a))) >= 1.0:
        raise ValueError('The polynomial must have all roots inside of the unit circle.')
    p = len(a) - 1
    a1 = np.concatenate((a, [0.0]))
    a2 = a1[::-1]
    P1 = a1 - a2
    Q1 = a1 + a2
    if p % 2:
        P = np.polydiv(P1, [1.0, 0.0, -1.0])[0]
        Q = Q1
    else:
        P = np.polydiv(P1, [1.0, -1.0])[0]
        Q = np.polydiv(Q1, [1.0, 1.0])[0]
    aP = np.angle(np.roots(P))
    aQ = np.angle(np.roots(Q))
    # spectrum takes every second root of each conjugate pair and negates the angle; equivalently the
    # positive angle of every pair.
    lsf = np.sort(np.concatenate((aP[aP > 0], aQ[aQ > 0])))
    assert len(lsf) == p, (len(lsf), p)
    return lsf


def lsf2poly(lsf, literal_dtype=False):
    """spectrum.lsf2poly [LIB].  `literal_dtype` is the named switch for one library behaviour: spectrum does `numpy.array(lsf)`
    WITHOUT a dtype, so the float32 row that tf.py_func hands lsf2poly_after_quan (lpc_utilities.py:28-33) makes
    `exp(1j * lsf)` and `numpy.poly` run in complex64.  Default (False): float64 arithmetic, the mathematically intended
    result; True: the library's literal dtype propagation.  Measured (tests/test_oracle_second_source.py): the literal path is
    4e-5 (median) to 1.3e-4 (worst of 300 frames) away from the float64 result -- the reference's own rounding noise; the CUDA
    kernel computes in float64."""
    lsf = np.array(lsf) if (literal_dtype and np.asarray(lsf).dtype == np.float32) else np.array(lsf, dtype=np.float64)
    if np.max(lsf) > np.pi or np.min(lsf) < 0:
        raise ValueError('Line spectral frequencies must be between 0 and pi.')
    p = len(lsf)
    z = np.exp(1j * lsf)
    rQ = z[0::2]
    rP = z[1::2]
    rQ = np.concatenate((rQ, rQ.conjugate()))
    rP = np.concatenate((rP, rP.conjugate()))
    Q = np.real(np.poly(rQ))
    P = np.real(np.poly(rP))
    if p % 2:
        P1 = np.convolve(P, [1.0, 0.0, -1.0])
        Q1 = Q
    else:
        P1 = np.convolve(P, [1.0, -1.0])
        Q1 = np.convolve(Q, [1.0, 1.0])
    a = 0.5 * (P1 + Q1)
    return a[:-1]


# ----------------------------------------------------------------------------------------------
# the reference functions
# ----------------------------------------------------------------------------------------------
def lpc_analysis_at_train(raw_data_one_batch, order):
    """lpc_utilities.py:14-25 (no caller in the shipped code)."""
    raw = np.asarray(raw_data_one_batch)[:, :, 0]
    out = np.empty((raw.shape[0], order))
    for i in range(raw.shape[0]):
        frame = empha_filter(highpass_filter(raw[i, :]))
        out[i, :] = poly2lsf(lpc_autocor(frame, order))
    return out


def lsf2poly_after_quan(lpc_in_lsf, order, literal_dtype=False):
    """lpc_utilities.py:28-33 (literal_dtype: see lsf2poly)."""
    lpc_in_lsf = np.asarray(lpc_in_lsf)
    out = np.empty((lpc_in_lsf.shape[0], order + 1))
    for i in range(lpc_in_lsf.shape[0]):
        out[i, :] = lsf2poly(lpc_in_lsf[i, :], literal_dtype)
    return out.astype(np.float32)


def residual_windows():
    """The 7 sub-frame windows of lpc_utilities.py:46-75 as a (7,128) array."""
    sub = FRAME_LENGTH // 4
    half = sub // 2
    h = np.hanning(half * 2)
    w = np.tile(h, (7, 1))
    w[0] = np.append(np.ones(half), h[half:])
    w[6] = np.append(h[:half], np.ones(half))
    return w


def lpc_analysis_get_residual(raw_data_one_batch, quan_lpc_coeff):
    """lpc_utilities.py:37-77."""
    raw = np.asarray(raw_data_one_batch)
    coeff = np.asarray(quan_lpc_coeff)
    how_many = raw.shape[0]
    res = np.zeros((how_many, FRAME_LENGTH))
    sub = FRAME_LENGTH // 4
    half = sub // 2
    win = residual_windows()
    for i in range(how_many):
        sig = raw[i].reshape(-1).astype(np.float64)
        a = coeff[i, :].astype(np.float64)
        for s in range(7):
            seg = sig[s * half: s * half + sub]
            res[i, s * half: s * half + sub] += fir_zero_state(a, seg) * win[s]
    return res.astype(np.float32)


def analysis_window_1024():
    """lpc_utilities.py:120-121."""
    h = np.hanning(512)
    return np.concatenate((h[:256], np.ones(512), h[256:]))


def lpc_windows_at_test(raw_data):
    """lpc_utilities.py:98-104: flatten the hop-480 segment matrix, cut 1024-windows at hop 512."""
    flat = np.asarray(raw_data)[:, :].flatten()
    starts = range(0, len(flat) - FRAME_LENGTH * 2, FRAME_LENGTH)
    ret = np.empty((len(starts), FRAME_LENGTH * 2))
    for ind, i in enumerate(starts):
        ret[ind, :] = flat[i:i + FRAME_LENGTH * 2]
    return ret


def lpc_analysis_windows(windows_1024, order):
    """Body of the loop at lpc_utilities.py:112-124 for already-cut (N,1024) windows."""
    windows_1024 = np.asarray(windows_1024)
    w = analysis_window_1024()
    out = np.empty((windows_1024.shape[0], order))
    for i in range(windows_1024.shape[0]):
        frame = windows_1024[i, :].astype(np.float64) * w
        out[i, :] = poly2lsf(lpc_autocor(frame, order))
    return out


def lpc_analysis_at_test(raw_data, order):
    """lpc_utilities.py:94-125."""
    return lpc_analysis_windows(lpc_windows_at_test(raw_data), order)


def lpc_synthesizer_tr(lpc_coeff, lpc_res):
    """lpc_utilities.py:137-156 (forward; the custom gradient is unreachable)."""
    lpc_coeff = np.asarray(lpc_coeff)
    lpc_res = np.asarray(lpc_res)
    out = np.empty((lpc_res.shape[0], FRAME_LENGTH))
    for i in range(lpc_res.shape[0]):
        out[i, :] = iir_zero_state([1.0], lpc_coeff[i, :].astype(np.float64), lpc_res[i, :].reshape(-1))
    return out.astype(np.float32)
