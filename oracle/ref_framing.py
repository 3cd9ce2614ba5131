"""Oracle restatement of the framing / overlap-add half of /root/reference/utilities.py and of the utterance-level
filters of lpc_utilities.py:8-11 (numpy / scipy).  TEST INFRASTRUCTURE ONLY.

[LIB] audiolazy: calling a ZFilter on a sequence runs the direct-form difference equation from zero initial state in
Python floats (float64); scipy.signal.lfilter computes the same recursion."""
import numpy as np
from scipy.signal import lfilter

frame_length = 512        # constants.py:25
overlap_each_side = 32    # constants.py:26
empha_filter_coeff = -0.68  # constants.py:64
HIGHPASS_B = [0.989502, -1.979004, 0.989592]   # lpc_utilities.py:10
HIGHPASS_A = [1, -1.978882, 0.979126]          # lpc_utilities.py:11


def windows():
    """utilities.py:10-15: (the_window, first_window, last_window)."""
    o, L = overlap_each_side, frame_length
    the_window = np.append(np.append(np.hanning(o * 2 - 1)[:o], np.array([1] * (L - o * 2))), np.hanning(o * 2 - 1)[o - 1:])
    first_window = np.append(np.append(np.array([1] * o), np.array([1] * (L - o * 2))), np.hanning(o * 2)[o:])
    last_window = np.append(np.append(np.hanning(o * 2)[:o], np.array([1] * (L - o * 2))), np.array([1] * o))
    return the_window, first_window, last_window


def hann_process(utterance_seg, seg_ind, seg_amount):
    """utilities.py:7-22."""
    the_window, first_window, last_window = windows()
    if seg_ind == 0:
        return utterance_seg * first_window
    if seg_ind == seg_amount - 1:
        return utterance_seg * last_window
    return utterance_seg * the_window


def utterance_to_segment(utterance, post_window=False):
    """utilities.py:25-39."""
    hop = frame_length - overlap_each_side
    starts = range(0, len(utterance) - frame_length, hop)
    ret = np.empty((len(starts), frame_length))
    w = windows()[0]
    for ind, i in enumerate(starts):
        ret[ind, :] = utterance[i:i + frame_length] * (1 if post_window else w)
    return ret


def lpc_windows_at_test(segments):
    """lpc_utilities.py:98-104: flatten the (N, 512) hop-480 frame matrix, cut 1024-sample windows at hop 512."""
    raw = segments.flatten()
    starts = range(0, len(raw) - frame_length * 2, frame_length)
    ret = np.empty((len(starts), frame_length * 2))
    for ind, i in enumerate(starts):
        ret[ind, :] = raw[i:i + frame_length * 2]
    return ret


def overlap_add(frames, seg_amount, n_used, out_len):
    """cmrl.py:595-597 / :710-716."""
    hop = frame_length - overlap_each_side
    out = np.zeros(out_len)
    for j in range(n_used):
        out[j * hop: j * hop + frame_length] += hann_process(frames[j], j, seg_amount)
    return out


def highpass_filter(x):
    return lfilter(HIGHPASS_B, HIGHPASS_A, np.asarray(x, dtype=np.float64))


def empha_filter(x):
    return lfilter([1.0, empha_filter_coeff], [1.0], np.asarray(x, dtype=np.float64))


def de_empha_filter(x):
    return lfilter([1.0], [1.0, empha_filter_coeff], np.asarray(x, dtype=np.float64))


def pack_bits(idx, bits):
    """Little-endian fixed-width packing of rows of indices (pure numpy; the format nsc_pack_codes writes)."""
    rows, L = idx.shape
    b = ((idx[:, :, None].astype(np.uint32) >> np.arange(bits)[None, None, :]) & 1).reshape(rows, L * bits)
    pad = (-b.shape[1]) % 8
    b = np.concatenate([b, np.zeros((rows, pad), dtype=b.dtype)], axis=1).reshape(rows, -1, 8)
    return (b << np.arange(8)[None, None, :]).sum(axis=2).astype(np.uint8)
