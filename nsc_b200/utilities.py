"""Drop-in for the framing half of /root/reference/utilities.py plus the utterance-level filters of
lpc_utilities.py:8-11, on torch CUDA tensors:

    utterance_to_segment(utterance, post_window=False)   (T,) -> (N, 512), hop 480            utilities.py:25-39
    hann_process(seg, seg_ind, seg_amount)                one frame times its overlap-add window  utilities.py:7-22
    overlap_add(frames, seg_amount, n_used, out_len)      the accumulation loops of cmrl.py:595-597, :710-716
    lpc_windows_at_test(utterance)                        the 1024-sample windows of lpc_analysis_at_test  lpc_utilities.py:98-104
    highpass_filter / empha_filter / de_empha_filter      audiolazy ZFilter calls (zero state), float64 arithmetic
    load_sig_lpc(s)                                       per-file std normalisation of _load_sig_lpc (nscm.py:98-113)

eval_metrics (PESQ / STOI, shells out to an external binary) is out of scope.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .constants import empha_filter_coeff, frame_length, overlap_each_side

_lib.load()
hop_size = frame_length - overlap_each_side

# lpc_utilities.py:8-11
HIGHPASS_B = (0.989502, -1.979004, 0.989592)
HIGHPASS_A = (1.0, -1.978882, 0.979126)


def _sig(x: torch.Tensor, name: str) -> torch.Tensor:
    return _lib.require_f32(x, name).reshape(-1)


def segment_count(T: int) -> int:
    return int(_lib.load().nsc_segment_count(int(T)))


def utterance_to_segment(utterance, post_window=False, *, offset: int = 0):
    """utilities.py:25-39.  ``offset`` = first sample (the LPC path frames ``sig[256:]``, cmrl.py:695)."""
    u = _sig(utterance, 'utterance')
    T = u.numel()
    N = segment_count(T - offset)
    seg = torch.empty((N, frame_length), dtype=torch.float32, device=u.device)
    _lib.check(_lib.load().nsc_utterance_to_segment(_lib.ptr(u), T, offset, 1 if post_window else 0, _lib.ptr(seg),
                                                    _lib.stream_ptr()), 'utterance_to_segment')
    return seg


def lpc_windows_at_test(utterance):
    """The (N', 1024) windows ``lpc_analysis_at_test`` analyses (lpc_utilities.py:98-104), cut from the FLATTENED
    hop-480 frame matrix of the utterance; pass them to ``lpc_utilities.lpc_analysis_windows``."""
    u = _sig(utterance, 'utterance')
    T = u.numel()
    lib = _lib.load()
    Nw = int(lib.nsc_lpc_window_count(lib.nsc_segment_count(T)))
    win = torch.empty((Nw, 2 * frame_length), dtype=torch.float32, device=u.device)
    _lib.check(lib.nsc_lpc_windows(_lib.ptr(u), T, _lib.ptr(win), _lib.stream_ptr()), 'lpc_windows_at_test')
    return win


def overlap_add(frames, seg_amount=None, n_used=None, out_len=None):
    """``out[480 j : 480 j + 512] += hann_process(frames[j], j, seg_amount)`` for j < n_used.
    Defaults reproduce the non-LPC loop (cmrl.py:566-597): every frame is used, out_len = 512 + 480 (N - 1).
    The LPC loop (cmrl.py:674-716) is ``overlap_add(frames, seg_amount=N2, n_used=N2 - 2, out_len=512 + 480 (N - 2))``."""
    f = _lib.require_f32(frames, 'frames').reshape(-1, frame_length)
    N = f.shape[0]
    seg_amount = N if seg_amount is None else int(seg_amount)
    n_used = N if n_used is None else int(n_used)
    if n_used > N:
        raise ValueError("n_used exceeds the number of frames")
    out_len = frame_length + hop_size * (n_used - 1) if out_len is None else int(out_len)
    out = torch.empty((max(out_len, 0),), dtype=torch.float32, device=f.device)
    _lib.check(_lib.load().nsc_overlap_add(_lib.ptr(f), n_used, seg_amount, _lib.ptr(out), out.numel(), _lib.stream_ptr()),
               'overlap_add')
    return out


def hann_process(utterance_seg, seg_ind, seg_amount):
    """utilities.py:7-22 on one frame (the batched form is ``overlap_add``)."""
    f = _lib.require_f32(utterance_seg, 'utterance_seg').reshape(1, frame_length)
    # a single frame overlap-added at position seg_ind: read its slice back
    pad = torch.zeros((seg_ind + 1, frame_length), dtype=torch.float32, device=f.device)
    pad[seg_ind] = f[0]
    out = overlap_add(pad, seg_amount=max(seg_amount, seg_ind + 1), n_used=seg_ind + 1)
    return out[seg_ind * hop_size: seg_ind * hop_size + frame_length].clone()


def _iir(x, b, a, out_f64=False):
    xs = _lib.require_f32(x, 'signal')
    sig = xs.reshape(1, -1) if xs.dim() == 1 else xs.reshape(xs.shape[0], -1)
    n, T = sig.shape
    lib = _lib.load()
    ws_bytes = int(lib.nsc_iir_workspace_bytes(T, n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=sig.device)
    y = torch.empty((n, T), dtype=torch.float64 if out_f64 else torch.float32, device=sig.device)
    bb = (C.c_double * 3)(*[float(v) for v in b])
    aa = (C.c_double * 3)(*[float(v) for v in a])
    _lib.check(lib.nsc_iir_biquad(_lib.ptr(sig), T, n, bb, aa, None if out_f64 else _lib.ptr(y), _lib.ptr(y) if out_f64 else None,
                                  _lib.ptr(ws), ws_bytes, _lib.stream_ptr()), 'iir_biquad')
    return y.reshape(xs.shape)


def highpass_filter(x, out_f64=False):
    """lpc_utilities.py:10-11 (note the asymmetric numerator taps 0.989502 / 0.989592)."""
    return _iir(x, HIGHPASS_B, HIGHPASS_A, out_f64)


def empha_filter(x, out_f64=False):
    """lpc_utilities.py:8: 1 + empha_filter_coeff z^-1 (pre-emphasis, coefficient -0.68)."""
    return _iir(x, (1.0, empha_filter_coeff, 0.0), (1.0, 0.0, 0.0), out_f64)


def de_empha_filter(x, out_f64=False):
    """cmrl.py:735 ``(1 / empha_filter)(sig)``."""
    return _iir(x, (1.0, 0.0, 0.0), (1.0, empha_filter_coeff, 0.0), out_f64)


def load_sig_lpc(s):
    """The normalisation of _load_sig_lpc (nscm.py:111-113): divide by np.std (population std)."""
    x = _lib.require_f32(s, 'signal')
    scale = x.std(unbiased=False)
    return x / scale, scale


# ---- batches of equal-length utterances (one launch per step instead of one per utterance) -------------------------------------
def _batch(x: torch.Tensor, name: str) -> torch.Tensor:
    x = _lib.require_f32(x, name)
    if x.dim() != 2:
        raise ValueError(f"{name} must be (n_utterances, T)")
    return x


def utterances_to_segments(utterances, post_window=False, *, offset: int = 0, n_take=None):
    """utterance_to_segment over a batch (n, T) of equal-length utterances -> (n, n_take, 512)."""
    u = _batch(utterances, 'utterances')
    n, T = u.shape
    N = segment_count(T - offset)
    n_take = N if n_take is None else int(n_take)
    seg = torch.empty((n, n_take, frame_length), dtype=torch.float32, device=u.device)
    _lib.check(_lib.load().nsc_utterances_to_segments(_lib.ptr(u), T, n, offset, 1 if post_window else 0, n_take, _lib.ptr(seg),
                                                      _lib.stream_ptr()), 'utterances_to_segments')
    return seg


def lpc_windows_at_test_batch(utterances, n_take=None):
    """lpc_windows_at_test over a batch (n, T) -> (n, n_take, 1024)."""
    u = _batch(utterances, 'utterances')
    n, T = u.shape
    lib = _lib.load()
    Nw = int(lib.nsc_lpc_window_count(lib.nsc_segment_count(T)))
    n_take = Nw if n_take is None else int(n_take)
    win = torch.empty((n, n_take, 2 * frame_length), dtype=torch.float32, device=u.device)
    _lib.check(lib.nsc_lpc_windows_batch(_lib.ptr(u), T, n, n_take, _lib.ptr(win), _lib.stream_ptr()), 'lpc_windows_at_test_batch')
    return win


def overlap_add_batch(frames, seg_amount, n_used, out_len):
    """overlap_add over a batch: frames (n, n_used, 512) -> (n, out_len)."""
    f = _lib.require_f32(frames, 'frames')
    if f.dim() != 3 or f.shape[2] != frame_length or f.shape[1] != int(n_used):
        raise ValueError("frames must be (n, n_used, 512)")
    n = f.shape[0]
    out = torch.empty((n, max(int(out_len), 0)), dtype=torch.float32, device=f.device)
    _lib.check(_lib.load().nsc_overlap_add_batch(_lib.ptr(f), n, int(n_used), int(seg_amount), _lib.ptr(out), out.shape[1],
                                                 _lib.stream_ptr()), 'overlap_add_batch')
    return out
