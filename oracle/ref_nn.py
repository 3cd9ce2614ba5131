"""Oracle restatement of /root/reference/nn_core_operator.py (torch-CPU).  TEST INFRASTRUCTURE ONLY.

Layout everywhere is the reference's channels-last (B, L, C).  ``dtype`` selects the arithmetic:
torch.float32 = like-for-like with the TF graph, torch.float64 = "truth" used to bound fp32 noise.

TensorFlow creates conv weights implicitly in call order inside a variable scope
(tf.compat.v1.layers.conv1d, nn_core_operator.py:6-14).  The oracle reproduces that with a
``ParamStream``: each conv call pulls the next (kernel[k,cin,cout], bias[cout]) pair, or -- in
init mode -- creates it Glorot-uniform / zeros the way TF would [LIB] and records it.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.2  # tf.nn.leaky_relu default alpha [LIB]; nn_core_operator.py:30


# ----------------------------------------------------------------------------------------------
# implicit-variable emulation
# ----------------------------------------------------------------------------------------------
class ParamStream:
    """Hands out conv parameters in creation order (TF variable-scope emulation)."""

    def __init__(self, params: Optional[List[Tuple[np.ndarray, ...]]] = None, seed: int = 0):
        self.init_mode = params is None
        self.params: List[Tuple[np.ndarray, ...]] = [] if params is None else list(params)
        self._cursor = 0
        self._rng = np.random.RandomState(seed)

    def _glorot(self, shape, fan_in, fan_out):
        limit = math.sqrt(6.0 / (fan_in + fan_out))  # glorot_uniform [LIB]
        return self._rng.uniform(-limit, limit, size=shape).astype(np.float32)

    def next_conv(self, k: int, cin: int, cout: int):
        if self.init_mode:
            w = self._glorot((k, cin, cout), k * cin, k * cout)
            # TF initialises biases to zero; the oracle uses small non-zero biases so that a
            # kernel that forgets the bias cannot pass parity.
            b = self._rng.uniform(-0.05, 0.05, size=(cout,)).astype(np.float32)
            self.params.append((w, b))
        w, b = self.params[self._cursor]
        assert w.shape == (k, cin, cout), (w.shape, (k, cin, cout), self._cursor)
        assert b.shape == (cout,)
        self._cursor += 1
        return w, b

    def next_sepconv(self, k: int, cin: int, cout: int):
        """Keras SeparableConv1D: depthwise [k,cin,1], pointwise [1,cin,cout], bias [cout] [LIB]."""
        if self.init_mode:
            dw = self._glorot((k, cin, 1), k * cin, k * 1)
            pw = self._glorot((1, cin, cout), cin, cout)
            b = self._rng.uniform(-0.05, 0.05, size=(cout,)).astype(np.float32)
            self.params.append((dw, pw, b))
        dw, pw, b = self.params[self._cursor]
        assert dw.shape == (k, cin, 1) and pw.shape == (1, cin, cout) and b.shape == (cout,)
        self._cursor += 1
        return dw, pw, b

    def done(self):
        assert self._cursor == len(self.params), (self._cursor, len(self.params))


# ----------------------------------------------------------------------------------------------
# SAME padding [LIB]: out = ceil(L/s); pad = max((out-1)s + (k-1)d + 1 - L, 0); left = pad // 2
# ----------------------------------------------------------------------------------------------
def same_padding(length: int, k: int, dilation: int, stride: int):
    out = -(-length // stride)
    total = max((out - 1) * stride + (k - 1) * dilation + 1 - length, 0)
    left = total // 2
    return out, left, total - left


def _apply_activation(x: torch.Tensor, activation):
    if activation is None:
        return x
    if activation == 'tanh':
        return torch.tanh(x)
    if activation == 'leaky_relu':
        return F.leaky_relu(x, LEAKY_SLOPE)
    raise ValueError(activation)


def conv1d(inputs: torch.Tensor, num_filters: int, filter_size: int, padding='SAME', dilation_rate=1,
           strides=1, activation='tanh', *, ps: ParamStream):
    """nn_core_operator.py:6-14.  Cross-correlation, channels-last, bias always on [LIB]."""
    assert padding == 'SAME'
    B, L, cin = inputs.shape
    w, b = ps.next_conv(filter_size, cin, num_filters)
    return conv1d_explicit(inputs, w, b, dilation_rate, strides, activation)


def conv1d_explicit(inputs: torch.Tensor, w, b, dilation_rate=1, strides=1, activation=None):
    B, L, cin = inputs.shape
    k = w.shape[0]
    _, left, right = same_padding(L, k, dilation_rate, strides)
    x = inputs.transpose(1, 2)  # (B, C, L)
    x = F.pad(x, (left, right))
    wt = torch.as_tensor(w, dtype=inputs.dtype).permute(2, 1, 0).contiguous()  # (cout, cin, k)
    bt = torch.as_tensor(b, dtype=inputs.dtype)
    y = F.conv1d(x, wt, bt, stride=strides, dilation=dilation_rate)
    y = y.transpose(1, 2).contiguous()
    return _apply_activation(y, activation)


def conv1d_depth(inputs: torch.Tensor, num_filters: int, filter_size: int, padding='SAME', dilation_rate=1,
                 strides=1, activation='tanh', *, ps: ParamStream):
    """nn_core_operator.py:17-21: Keras SeparableConv1D (depth multiplier 1) [LIB]."""
    assert padding == 'SAME'
    B, L, cin = inputs.shape
    dw, pw, b = ps.next_sepconv(filter_size, cin, num_filters)
    return conv1d_depth_explicit(inputs, dw, pw, b, dilation_rate, strides, activation)


def conv1d_depth_explicit(inputs, dw, pw, b, dilation_rate=1, strides=1, activation=None):
    B, L, cin = inputs.shape
    k = dw.shape[0]
    _, left, right = same_padding(L, k, dilation_rate, strides)
    x = F.pad(inputs.transpose(1, 2), (left, right))
    dwt = torch.as_tensor(dw, dtype=inputs.dtype).permute(1, 2, 0).contiguous()  # (cin, 1, k)
    y = F.conv1d(x, dwt, None, stride=strides, dilation=dilation_rate, groups=cin)
    pwt = torch.as_tensor(pw, dtype=inputs.dtype).permute(2, 1, 0).contiguous()  # (cout, cin, 1)
    y = F.conv1d(y, pwt, torch.as_tensor(b, dtype=inputs.dtype))
    return _apply_activation(y.transpose(1, 2).contiguous(), activation)


def activation_func(x):
    """nn_core_operator.py:24-31."""
    return F.leaky_relu(x, LEAKY_SLOPE)


def batch_norm(x, training=None):
    """nn_core_operator.py:34-42: identity."""
    return x


def change_channel(the_input, wide_layer=30, the_channel=1, kernel_size=9, dilation_rate=1, strides=1,
                   activation=None, *, ps: ParamStream):
    """nn_core_operator.py:45-54: dilation is forced to 1 whatever the argument says."""
    return conv1d(the_input, the_channel, filter_size=kernel_size, padding='SAME', dilation_rate=1,
                  strides=strides, activation=activation, ps=ps)


def the_bottleneck(the_input, wide_layer=30, narrow_layer=10, non_dilated_neck_kernel_size=9,
                   dilated_neck_kernel_size=9, dilation_rate=1, is_last_flat=False, *, ps: ParamStream):
    """nn_core_operator.py:57-79."""
    y = conv1d(the_input, narrow_layer, non_dilated_neck_kernel_size, dilation_rate=1, activation=None, ps=ps)
    y = activation_func(y)
    y = conv1d(y, narrow_layer, dilated_neck_kernel_size, dilation_rate=dilation_rate, activation=None, ps=ps)
    y = activation_func(y)
    y = conv1d(y, wide_layer, non_dilated_neck_kernel_size, dilation_rate=1, activation=None, ps=ps)
    if not is_last_flat:
        return activation_func(y + the_input)  # a 1-channel input broadcasts (nn_core_operator.py:77)
    return y + the_input


def gated_bottleneck(the_input, wide_layer=30, narrow_layer=10, non_dilated_neck_kernel_size=9,
                     dilated_neck_kernel_size=9, dilation_rate=1, is_last_flat=False, the_share=False,
                     *, ps: ParamStream):
    """nn_core_operator.py:82-112.  Kernel 15 of the gate convs is hard-coded (:92, :97)."""
    y = conv1d(the_input, narrow_layer, 1, dilation_rate=1, activation=None, ps=ps)
    y = activation_func(y)
    left = conv1d(y, narrow_layer, 15, dilation_rate=dilation_rate, activation=None, ps=ps)
    right = conv1d(y, narrow_layer, 15, dilation_rate=dilation_rate, activation='tanh', ps=ps)
    y = left * right
    y = conv1d(y, wide_layer, non_dilated_neck_kernel_size, dilation_rate=1, activation=None, ps=ps)
    if not is_last_flat:
        return activation_func(y + the_input)
    return y + the_input


def gated_bottleneck_decoder(the_input, wide_layer=30, narrow_layer=10, non_dilated_neck_kernel_size=9,
                             dilated_neck_kernel_size=9, dilation_rate=1, is_last_flat=False, the_share=False,
                             *, ps: ParamStream):
    """nn_core_operator.py:115-137 (no caller in the reference)."""
    y = conv1d(the_input, narrow_layer, 1, dilation_rate=1, activation=None, ps=ps)
    y = activation_func(y)
    left = conv1d(y, narrow_layer, dilated_neck_kernel_size, dilation_rate=dilation_rate, activation=None, ps=ps)
    right = conv1d(y, narrow_layer, dilated_neck_kernel_size, dilation_rate=dilation_rate, activation='tanh', ps=ps)
    y = left * right
    y = conv1d_depth(y, wide_layer, non_dilated_neck_kernel_size, dilation_rate=1, activation=None, ps=ps)
    if not is_last_flat:
        return activation_func(y + the_input)
    return y + the_input


def scalar_softmax_quantization(floating_code, alpha, bins, is_quan_on, the_share, code_length, num_kmean_kernels):
    """nn_core_operator.py:140-164.

    Returns (soft_assignment (B,L,n) -- always the SOFT one, bit_code (B,L,1)).
    ``the_share`` True selects the soft assignment for the value path, False the one-hot
    (tf.cond, :154-158).  top_k returns the lowest index among ties [LIB].
    """
    dt = floating_code.dtype
    bins_t = torch.as_tensor(bins, dtype=dt).reshape(1, 1, -1)
    alpha_t = torch.as_tensor(alpha, dtype=dt)
    dist = torch.abs(floating_code - bins_t)
    logits = alpha_t * dist
    soft = torch.softmax(logits, dim=-1)
    idx = first_argmax(soft)
    hard = F.one_hot(idx, num_kmean_kernels).to(dt).reshape(-1, code_length, num_kmean_kernels)
    sel = soft if the_share else hard
    bit_code = torch.matmul(sel, bins_t.reshape(-1, 1)).reshape(-1, floating_code.shape[1], 1)
    iq = torch.as_tensor(is_quan_on, dtype=dt)
    bit_code = (1 - iq) * floating_code + iq * bit_code
    return soft, bit_code.reshape(-1, floating_code.shape[1], 1)


def first_argmax(t: torch.Tensor) -> torch.Tensor:
    """Lowest index attaining the maximum along the last axis (tf.nn.top_k tie rule [LIB])."""
    m = t.max(dim=-1, keepdim=True).values
    n = t.shape[-1]
    ar = torch.arange(n).expand_as(t)
    cand = torch.where(t == m, ar, torch.full_like(ar, n))
    return cand.min(dim=-1).values


def quantizer_indices(floating_code, alpha, bins):
    """Implicit integer code of scalar_softmax_quantization (the argmax of the literal softmax)."""
    dt = floating_code.dtype
    bins_t = torch.as_tensor(bins, dtype=dt).reshape(1, 1, -1)
    logits = torch.as_tensor(alpha, dtype=dt) * torch.abs(floating_code - bins_t)
    return first_argmax(torch.softmax(logits, dim=-1))


def quantizer_indices_from_logits(floating_code, alpha, bins):
    """First arg-max of the fp32 logits -- the form the CUDA kernel uses."""
    dt = floating_code.dtype
    bins_t = torch.as_tensor(bins, dtype=dt).reshape(1, 1, -1)
    logits = torch.as_tensor(alpha, dtype=dt) * torch.abs(floating_code - bins_t)
    return first_argmax(logits)


def vec_l2norm(x):
    """loss_terms_and_measures.py:9-10."""
    return torch.sqrt(torch.sum(x * x, dim=-1))
