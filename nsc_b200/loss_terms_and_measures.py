"""Drop-in for the hot-path rows of /root/reference/loss_terms_and_measures.py.

Only the functions wired into a training / evaluation graph exist here (mse_loss, mse_loss_v1, mfcc_loss,
quan_loss, entropy_coding_loss, entropy_to_bitrate, vec_l2norm_tf); si_snr, mu-law, mdct, psd/SMR/MNR and the
`pesq` subprocess are outside the path.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .constants import frame_length, overlap_each_side, sample_rate

_lib.load()
_MELW = {}


def mel_filterbank(device) -> torch.Tensor:
    """The 4 HTK banks of mfcc_transform (:130-148) as one (257*184 + 2*184) float32 device buffer."""
    key = str(device)
    if key not in _MELW:
        buf = torch.empty(_lib.MEL_BUFFER_FLOATS, dtype=torch.float32, device=device)
        _lib.check(_lib.load().nsc_mel_filterbank(_lib.ptr(buf), _lib.stream_ptr()), 'mel_filterbank')
        _MELW[key] = buf
    return _MELW[key]


def _pair(decoded_sig, original_sig):
    d = _lib.require_f32(decoded_sig, 'decoded_sig').reshape(-1, frame_length)
    o = _lib.require_f32(original_sig, 'original_sig').reshape(-1, frame_length)
    if d.shape != o.shape:
        raise ValueError("decoded / original shape mismatch")
    return d, o


def mse_loss(decoded_sig, original_sig, kai_re_mat=1):
    """:77-79 -- a per-frame RMSE of shape (B,)."""
    d, o = _pair(decoded_sig, original_sig)
    out = torch.empty(d.shape[0], dtype=torch.float32, device=d.device)
    _lib.check(_lib.load().nsc_losses_forward(_lib.ptr(d), _lib.ptr(o), d.shape[0], None, _lib.ptr(out), None,
                                              _lib.stream_ptr()), 'mse_loss')
    return out


mse_loss_v1 = mse_loss   # :82-84, identical body


def mfcc_loss(decoded_sig, original_sig, is_finetuning=False):
    """:151-175 -- mean over the 4 mel resolutions of the per-frame log-mel RMSE, shape (B,)."""
    d, o = _pair(decoded_sig, original_sig)
    out = torch.empty(d.shape[0], dtype=torch.float32, device=d.device)
    _lib.check(_lib.load().nsc_losses_forward(_lib.ptr(d), _lib.ptr(o), d.shape[0], _lib.ptr(mel_filterbank(d.device)),
                                              None, _lib.ptr(out), _lib.stream_ptr()), 'mfcc_loss')
    return out


def losses(decoded_sig, original_sig):
    """mse_loss and mfcc_loss in ONE launch (they share the two signal reads)."""
    d, o = _pair(decoded_sig, original_sig)
    t = torch.empty(d.shape[0], dtype=torch.float32, device=d.device)
    f = torch.empty_like(t)
    _lib.check(_lib.load().nsc_losses_forward(_lib.ptr(d), _lib.ptr(o), d.shape[0], _lib.ptr(mel_filterbank(d.device)),
                                              _lib.ptr(t), _lib.ptr(f), _lib.stream_ptr()), 'losses')
    return t, f


def quan_loss(softmax_assignment):
    """:257-259 on a materialised (B, L, n) soft assignment.  (The codec path gets the same number from the
    quantiser kernel without ever writing the soft tensor -- see codec.NeuralCodec.)"""
    s = _lib.require_f32(softmax_assignment, 'softmax_assignment')
    B, L, n = s.shape
    out = torch.empty(B, dtype=torch.float32, device=s.device)
    _lib.check(_lib.load().nsc_quan_loss(_lib.ptr(s), B, L, n, _lib.ptr(out), _lib.stream_ptr()), 'quan_loss')
    return out


def entropy_coding_loss(soft_assignment):
    """:262-267 on a materialised soft assignment: one scalar over the whole batch."""
    s = _lib.require_f32(soft_assignment, 'soft_assignment')
    n = s.shape[2]
    hist = torch.zeros(n, dtype=torch.float32, device=s.device)
    _lib.check(_lib.load().nsc_soft_histogram(_lib.ptr(s), s.numel() // n, n, _lib.ptr(hist), _lib.stream_ptr()),
               'entropy_coding_loss')
    return entropy_from_hist(hist)


def entropy_from_hist(hist):
    h = _lib.require_f32(hist, 'hist')
    out = torch.empty((), dtype=torch.float32, device=h.device)
    _lib.check(_lib.load().nsc_entropy_from_hist(_lib.ptr(h), h.numel(), _lib.ptr(out), _lib.stream_ptr()),
               'entropy_coding_loss')
    return out


def vec_l2norm_tf(x):
    """:9-10."""
    return torch.sqrt(torch.sum(x * x, dim=-1))


def entropy_to_bitrate(total_entropy, the_strides):
    """:63-67."""
    code_len_val = 128 if the_strides == 4 else 256
    return ((sample_rate / 1024.0) / (frame_length - overlap_each_side)) * code_len_val * total_entropy
