"""Oracle restatement of the hot-path rows of /root/reference/loss_terms_and_measures.py (torch-CPU).

TEST INFRASTRUCTURE ONLY.  Only the functions wired into a training/eval graph are restated
(SURVEY.md 2.1 row 3): mse_loss(_v1), tf_stft, mfcc_transform, mfcc_loss, quan_loss,
entropy_coding_loss, entropy_to_bitrate.
"""
from __future__ import annotations

import math

import numpy as np
import torch

FRAME_LENGTH = 512
SAMPLE_RATE = 16000
OVERLAP_EACH_SIDE = 32
MEL_BANKS = (8, 16, 32, 128)     # loss_terms_and_measures.py:133


def mse_loss(decoded_sig, original_sig, kai_re_mat=1):
    """loss_terms_and_measures.py:77-79: a per-frame RMSE of shape (B,)."""
    mse = torch.mean((decoded_sig - original_sig) ** 2, dim=-1)
    return torch.sqrt(mse + 1e-07)


mse_loss_v1 = mse_loss  # :82-84 identical body


def hertz_to_mel(f):
    return 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)  # HTK [LIB]


def linear_to_mel_weight_matrix(num_mel_bins, num_spectrogram_bins, sample_rate, lower_edge_hertz, upper_edge_hertz):
    """tf.signal.linear_to_mel_weight_matrix [LIB]: HTK triangles, DC row zero, float64 then cast."""
    nyquist = sample_rate / 2.0
    bands_to_zero = 1
    lin = np.linspace(0.0, nyquist, num_spectrogram_bins)[bands_to_zero:]
    spec_mel = hertz_to_mel(lin)[:, None]
    edges = np.linspace(hertz_to_mel(lower_edge_hertz), hertz_to_mel(upper_edge_hertz), num_mel_bins + 2)
    lower = edges[:-2][None, :]
    center = edges[1:-1][None, :]
    upper = edges[2:][None, :]
    lower_slopes = (spec_mel - lower) / (center - lower)
    upper_slopes = (upper - spec_mel) / (upper - center)
    m = np.maximum(0.0, np.minimum(lower_slopes, upper_slopes))
    return np.pad(m, [[bands_to_zero, 0], [0, 0]]).astype(np.float32)


def tf_stft(sig, the_frame_length=FRAME_LENGTH):
    """:178-183: ONE un-windowed length-512 rFFT per frame; mag = sqrt(re^2 + im^2 + 1e-7)."""
    x = sig.reshape(-1, FRAME_LENGTH)
    st = torch.fft.rfft(x, n=the_frame_length, dim=-1)
    mag = torch.sqrt(st.real ** 2 + st.imag ** 2 + 1e-7)
    return st, mag


def mfcc_transform(the_stft, the_spectrum, is_finetuning=False):
    """:130-148."""
    nbins = the_stft.shape[-1]
    out = []
    for n in MEL_BANKS:
        m = torch.as_tensor(linear_to_mel_weight_matrix(n, nbins, 16000, 0.0, 8000.0)).to(the_spectrum.dtype)
        out.append(torch.log(the_spectrum @ m + 1e-7))
    return out


def mfcc_loss(decoded_sig, original_sig, is_finetuning=False):
    """:151-175.  The `shape[0] == 128` branch tests the batch dimension (None in every graph) and is dead."""
    dec_st, dec_sp = tf_stft(decoded_sig)
    ori_st, ori_sp = tf_stft(original_sig)
    ori_sp = 1.0 / FRAME_LENGTH * ori_sp ** 2
    dec_sp = 1.0 / FRAME_LENGTH * dec_sp ** 2
    pred = mfcc_transform(dec_st, dec_sp)
    true = mfcc_transform(ori_st, ori_sp)
    d = [mse_loss_v1(p, t).unsqueeze(-1) for p, t in zip(pred, true)]
    return torch.mean(torch.cat(d, dim=-1), dim=-1)


def quan_loss(softmax_assignment):
    """:257-259."""
    return torch.mean(torch.sum(torch.sqrt(softmax_assignment + 1e-20), dim=-1), dim=-1)


def entropy_coding_loss(soft_assignment):
    """:262-267: ONE scalar over the whole batch."""
    s = soft_assignment.reshape(-1, soft_assignment.shape[2])
    hist = torch.sum(s, dim=0)
    hist = hist / torch.sum(hist)
    return -torch.sum(hist * torch.log(hist + 1e-7) / math.log(2.0))


def entropy_to_bitrate(total_entropy, the_strides):
    """:63-67."""
    code_len_val = 128 if the_strides == 4 else 256
    return ((SAMPLE_RATE / 1024.0) / (FRAME_LENGTH - OVERLAP_EACH_SIDE)) * code_len_val * total_entropy
