"""Drop-in for /root/reference/lpc_utilities.py on torch CUDA tensors.

The reference's versions take/return numpy arrays and are spliced into the TF graph with tf.py_func
(SURVEY.md section 2.3); these take/return CUDA tensors with the same shapes and the same float32 results:
    lsf2poly_after_quan       (B,16)            -> (B,17) float32
    lpc_analysis_get_residual (B,512,1), (B,17) -> (B,512) float32
    lpc_synthesizer_tr        (B,17), (B,512)   -> (B,512) float32
    lpc_analysis_at_test      (N,512) segments  -> (N',16) float64 LSFs
    lpc_analysis_at_train     (B,512,1)         -> (B,16)  float64 LSFs
``strict=True`` reproduces the reference's error behaviour (spectrum raises on LSFs outside [0, pi]; audiolazy
divides by zero on silent frames) at the cost of one host synchronisation; the default leaves NaN rows.
"""
from __future__ import annotations

import torch

from . import _lib
from .constants import frame_length

_lib.load()


def _status(device):
    return torch.zeros(1, dtype=torch.int32, device=device)


def lsf2poly_after_quan(lpc_in_lsf, order=16, *, strict: bool = False):
    """lpc_utilities.py:28-33."""
    if order != _lib.LPC_ORDER:
        raise ValueError("the hot path is order 16 (neural_speech_coding_module.py:50)")
    lsf = _lib.require_f32(lpc_in_lsf, 'lpc_in_lsf').reshape(-1, order)
    B = lsf.shape[0]
    poly = torch.empty((B, order + 1), dtype=torch.float32, device=lsf.device)
    st = _status(lsf.device) if strict else None
    _lib.check(_lib.load().nsc_lsf2poly(_lib.ptr(lsf), B, _lib.ptr(poly), _lib.ptr(st), _lib.stream_ptr()), 'lsf2poly_after_quan')
    if strict and int(st.item()) != 0:
        raise ValueError('Line spectral frequencies must be between 0 and pi.')
    return poly


def lpc_analysis_get_residual(raw_data_one_batch, quan_lpc_coeff):
    """lpc_utilities.py:37-77."""
    x = _lib.require_f32(raw_data_one_batch, 'raw_data_one_batch').reshape(-1, frame_length)
    a = _lib.require_f32(quan_lpc_coeff, 'quan_lpc_coeff').reshape(-1, _lib.LPC_ORDER + 1)
    if a.shape[0] != x.shape[0]:
        raise ValueError("batch mismatch between frames and LPC coefficients")
    res = torch.empty_like(x)
    _lib.check(_lib.load().nsc_lpc_residual(_lib.ptr(x), _lib.ptr(a), x.shape[0], _lib.ptr(res), _lib.stream_ptr()),
               'lpc_analysis_get_residual')
    return res


def lpc_synthesizer_tr(lpc_coeff, lpc_res):
    """lpc_utilities.py:137-156 (forward; its custom gradient is unreachable through py_func)."""
    a = _lib.require_f32(lpc_coeff, 'lpc_coeff').reshape(-1, _lib.LPC_ORDER + 1)
    r = _lib.require_f32(lpc_res, 'lpc_res').reshape(-1, frame_length)
    if a.shape[0] != r.shape[0]:
        raise ValueError("batch mismatch between LPC coefficients and residual")
    y = torch.empty_like(r)
    _lib.check(_lib.load().nsc_lpc_synth(_lib.ptr(a), _lib.ptr(r), r.shape[0], _lib.ptr(y), _lib.stream_ptr()),
               'lpc_synthesizer_tr')
    return y


def lpc_windows_at_test(raw_data):
    """The window cutting of lpc_utilities.py:98-104: flatten the (N,512) hop-480 segment matrix (the 32-sample
    overlaps are therefore repeated -- a reference quirk kept on purpose) and cut 1024-long windows at hop 512."""
    flat = raw_data.reshape(-1)
    n = flat.numel()
    count = len(range(0, n - frame_length * 2, frame_length))
    if count <= 0:
        return flat.new_empty((0, frame_length * 2))
    return flat.unfold(0, frame_length * 2, frame_length)[:count].contiguous()


def lpc_analysis_windows(windows, order=16, *, strict: bool = False, dtype=torch.float64):
    """Loop body of lpc_utilities.py:112-124 on already-cut (N,1024) windows -> (N,16) LSFs.
    float64 like the reference's array by default; dtype=torch.float32 returns the cast that the float32
    `lpc_x` placeholder receives (cmrl.py:699-702) straight from the kernel."""
    if order != _lib.LPC_ORDER:
        raise ValueError("the hot path is order 16")
    w = _lib.require_f32(windows, 'windows').reshape(-1, frame_length * 2)
    N = w.shape[0]
    lsf = torch.empty((N, order), dtype=dtype, device=w.device)
    st = _status(w.device) if strict else None
    p64, p32 = (_lib.ptr(lsf), None) if dtype == torch.float64 else (None, _lib.ptr(lsf))
    _lib.check(_lib.load().nsc_lpc_analyze(_lib.ptr(w), N, p64, p32, _lib.ptr(st), _lib.stream_ptr()), 'lpc_analysis')
    if strict and int(st.item()) != 0:
        raise ZeroDivisionError('LPC analysis failed on %d frame(s) (silent or non-minimum-phase)' % int(st.item()))
    return lsf


def lpc_analysis_at_test(raw_data, order=16, *, strict: bool = False):
    """lpc_utilities.py:94-125."""
    return lpc_analysis_windows(lpc_windows_at_test(_lib.require_f32(raw_data, 'raw_data')), order, strict=strict)


def lpc_analysis_at_train(raw_data_one_batch, order=16, *, strict: bool = False):
    """lpc_utilities.py:14-25 (no caller in the shipped reference; surface parity only)."""
    if order != _lib.LPC_ORDER:
        raise ValueError("the hot path is order 16")
    x = _lib.require_f32(raw_data_one_batch, 'raw_data_one_batch').reshape(-1, frame_length)
    B = x.shape[0]
    lsf = torch.empty((B, order), dtype=torch.float64, device=x.device)
    st = _status(x.device) if strict else None
    _lib.check(_lib.load().nsc_lpc_analyze_train(_lib.ptr(x), B, _lib.ptr(lsf), None, _lib.ptr(st), _lib.stream_ptr()),
               'lpc_analysis_at_train')
    if strict and int(st.item()) != 0:
        raise ZeroDivisionError('LPC analysis failed on %d frame(s)' % int(st.item()))
    return lsf
