"""Drop-in for /root/reference/nn_core_operator.py on torch CUDA tensors -- same function names, positional
order and defaults.

The one signature extension: TensorFlow creates the conv weights implicitly inside a variable scope
(`tf.compat.v1.layers.conv1d`, nn_core_operator.py:6-14); eager code cannot, so every conv-bearing function
takes a trailing keyword ``params`` -- a tuple/list of float32 CUDA tensors in TF's creation order
(kernel (k, cin, cout), bias (cout); SeparableConv1D: depthwise (k, cin, 1), pointwise (1, cin, cout), bias).
Tensors are channels-last (B, L, C) float32, exactly like the reference graph.  Activations are named
(`None`, ``'tanh'``, ``'leaky_relu'``); `torch.tanh` / `activation_func` are accepted as aliases.

All arithmetic runs in libnsc_b200.so (hand-written sm_100a kernels); there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib
from . import constants as _c

_lib.load()   # fail loudly at import time if the CUDA library is missing

# Conv arithmetic of the surface functions, same modes as CodecConfig.precision:
#   'tc_f16x3' (default)  tcgen05 tensor cores, fp16 hi/lo split, fp32 accumulate: fp32-class results (<= 1e-4, tests/test_gpu_plane.py)
#   'tc_f16'              tcgen05, plain fp16 inputs: REDUCED precision, stated separately
#   'fp32'                CUDA-core FFMA
# Shapes the tensor engine does not cover (anything but the codec's layer shapes) run on the FFMA engine in every mode.
_ENGINE = 'tc_f16x3'
_PRECISION_CODE = {'tc_f16x3': 1, 'tc_f16': 2}
last_engine = None      # 'tc' / 'tc_folded' / 'tc_fused' / 'ffma': which engine the most recent conv-bearing call ran on (for tests)


def set_engine(name: str) -> str:
    """Selects the conv arithmetic of conv1d / change_channel / the_bottleneck; returns the previous setting."""
    global _ENGINE
    if name not in ('tc_f16x3', 'tc_f16', 'fp32'):
        raise ValueError(f"unknown engine {name!r}")
    prev, _ENGINE = _ENGINE, name
    return prev


def _act_code(activation) -> int:
    if activation is None:
        return _lib.ACT_NONE
    if activation in ('tanh', torch.tanh) or getattr(activation, '__name__', '') == 'tanh':
        return _lib.ACT_TANH
    if activation in ('leaky_relu', 'lrelu') or getattr(activation, '__name__', '') in ('activation_func', 'leaky_relu'):
        return _lib.ACT_LRELU
    raise ValueError(f"unsupported activation {activation!r}")


def _same_out(length: int, stride: int) -> int:
    return -(-length // stride)


def _check_cl(x: torch.Tensor, name: str) -> torch.Tensor:
    if x.dim() != 3:
        raise ValueError(f"{name} must be (B, L, C), got {tuple(x.shape)}")
    return _lib.require_f32(x, name)


def conv1d(inputs, num_filters, filter_size, padding='SAME', dilation_rate=1, strides=1, activation='tanh', *,
           params: Sequence[torch.Tensor]):
    """nn_core_operator.py:6-14."""
    if padding != 'SAME':
        raise ValueError("only padding='SAME' exists on the reference's path")
    x = _check_cl(inputs, 'inputs')
    w, b = params
    B, L, cin = x.shape
    if tuple(w.shape) != (filter_size, cin, num_filters) or tuple(b.shape) != (num_filters,):
        raise ValueError(f"conv1d params {tuple(w.shape)}/{tuple(b.shape)} do not match "
                         f"({filter_size},{cin},{num_filters})")
    global last_engine
    y = torch.empty((B, _same_out(L, strides), num_filters), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    w, b = _lib.require_f32(w, 'kernel'), _lib.require_f32(b, 'bias')
    prec = _PRECISION_CODE.get(_ENGINE)
    if prec is not None and B > 0 and (num_filters == 1 or _act_code(activation) != _lib.ACT_TANH):
        # the codec's layer shapes run on the tensor engine (the same kernels the whole-codec entry points launch)
        ws_bytes = lib.nsc_conv1d_tc_workspace_bytes(B, L, cin, num_filters, filter_size, dilation_rate, strides, 0, 1, prec)
        if ws_bytes > 0:
            ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=x.device)
            rc = lib.nsc_conv1d_tc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), None, _lib.ptr(y), B, L, cin, num_filters, filter_size,
                                   dilation_rate, strides, _act_code(activation), 0, _lib.ACT_NONE, 1, prec, _lib.ptr(ws), ws_bytes,
                                   _lib.stream_ptr())
            _lib.check(rc, 'conv1d')
            last_engine = 'tc'
            return y
    rc = lib.nsc_conv1d(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), B, L, cin, num_filters, filter_size, dilation_rate,
                        strides, _act_code(activation), _lib.stream_ptr())
    _lib.check(rc, 'conv1d')
    last_engine = 'ffma'
    return y


def conv1d_depth(inputs, num_filters, filter_size, padding='SAME', dilation_rate=1, strides=1, activation='tanh', *,
                 params: Sequence[torch.Tensor]):
    """nn_core_operator.py:17-21 (Keras SeparableConv1D, depth multiplier 1)."""
    if padding != 'SAME':
        raise ValueError("only padding='SAME' exists on the reference's path")
    x = _check_cl(inputs, 'inputs')
    dw, pw, b = params
    B, L, cin = x.shape
    if tuple(dw.shape) != (filter_size, cin, 1) or tuple(pw.shape) != (1, cin, num_filters):
        raise ValueError("conv1d_depth params do not match the layer shape")
    Lout = _same_out(L, strides)
    tmp = torch.empty((B, Lout, cin), dtype=torch.float32, device=x.device)
    y = torch.empty((B, Lout, num_filters), dtype=torch.float32, device=x.device)
    rc = _lib.load().nsc_conv1d_depth(_lib.ptr(x), _lib.ptr(_lib.require_f32(dw, 'dw')), _lib.ptr(_lib.require_f32(pw, 'pw')),
                                      _lib.ptr(_lib.require_f32(b, 'bias')), _lib.ptr(tmp), _lib.ptr(y), B, L, cin,
                                      num_filters, filter_size, dilation_rate, strides, _act_code(activation),
                                      _lib.stream_ptr())
    _lib.check(rc, 'conv1d_depth')
    return y


def activation_func(_x):
    """nn_core_operator.py:24-31: leaky ReLU, slope 0.2.  (Inside the codec it is fused into the conv epilogue.)"""
    return torch.nn.functional.leaky_relu(_x, 0.2)


def batch_norm(_x, training=None):
    """nn_core_operator.py:34-42: identity."""
    return _x


def change_channel(the_input, wide_layer=30, the_channel=1, kernel_size=9, dilation_rate=1, strides=1,
                   activation=None, *, params):
    """nn_core_operator.py:45-54: the dilation argument is ignored (forced to 1) exactly like the reference."""
    return conv1d(the_input, the_channel, filter_size=kernel_size, padding='SAME', dilation_rate=1, strides=strides,
                  activation=activation, params=params)


def _flatten_params(params, n_expected: int) -> torch.Tensor:
    flat = []
    for p in params:
        flat.extend(p)
    if len(flat) != n_expected:
        raise ValueError(f"expected {n_expected // 2} conv (kernel, bias) pairs, got {len(flat)} tensors")
    return torch.cat([_lib.require_f32(t, 'param').reshape(-1) for t in flat])


def _block(the_input, wide_layer, narrow_layer, k_plain, k_dilated, dilation_rate, is_last_flat, gated, params, fused=None):
    x = _check_cl(the_input, 'the_input')
    B, L, cin = x.shape
    flat = _flatten_params(params, 8 if gated else 6)
    global last_engine
    y = torch.empty((B, L, wide_layer), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    prec = _PRECISION_CODE.get(_ENGINE)
    if prec is not None and not gated and B > 0 and cin == wide_layer:
        # the_bottleneck on the tensor engine (three launches as in the codec program, or the one-launch fused kernel on request)
        ws_bytes = lib.nsc_bottleneck_block_tc_workspace_bytes(B, L, wide_layer, narrow_layer, prec)
        if ws_bytes > 0 and k_plain == 9 and k_dilated == 9 and dilation_rate in (1, 2):
            import ctypes as C
            ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=x.device)
            # fused: None = the codec program's form -- three launches, the 20 -> 20 conv on folded images where the frame is long
            # enough (hi/lo planes, 256 positions up); 'folded' / False force the folded / taps-in-N form; True = the ONE-launch fused
            # kernel (opt-in: an intermittent wrong result on the first launch of a large uneven batch is open, DESIGN.md finding 11)
            if fused is None:
                fused = 'folded' if (prec == 1 and L >= 256) else False
            want_folded = fused == 'folded'
            fused = C.c_int32(-2 if want_folded else (0 if fused else -1))
            rc = lib.nsc_bottleneck_block_tc(_lib.ptr(x), _lib.ptr(flat), _lib.ptr(y), B, L, wide_layer, narrow_layer, k_plain, k_dilated,
                                             dilation_rate, int(bool(is_last_flat)), prec, C.byref(fused), _lib.ptr(ws), ws_bytes,
                                             _lib.stream_ptr())
            _lib.check(rc, 'the_bottleneck')
            last_engine = 'tc_fused' if fused.value else ('tc_folded' if want_folded else 'tc')
            return y
    if prec is not None and gated and B > 0 and cin == wide_layer and k_plain == 9 and dilation_rate in (1, 2):
        # the gated block of the codec path: k1 conv, fused gate pair (product in the epilogue), k9 conv + residual on tcgen05
        ws_bytes = lib.nsc_gated_block_tc_workspace_bytes(B, L, wide_layer, narrow_layer, dilation_rate, prec)
        if ws_bytes > 0:
            ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=x.device)
            rc = lib.nsc_gated_block_tc(_lib.ptr(x), _lib.ptr(flat), _lib.ptr(y), B, L, wide_layer, narrow_layer, k_plain, dilation_rate,
                                        int(bool(is_last_flat)), prec, _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
            _lib.check(rc, 'gated_bottleneck')
            last_engine = 'tc'
            return y
    last_engine = 'ffma'
    ws_bytes = lib.nsc_block_workspace_bytes(B, L, wide_layer, narrow_layer)
    ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=x.device)
    rc = lib.nsc_bottleneck_block(_lib.ptr(x), _lib.ptr(flat), _lib.ptr(y), B, L, cin, wide_layer, narrow_layer, k_plain,
                                  k_dilated, dilation_rate, int(bool(is_last_flat)), int(gated), _lib.ptr(ws), ws_bytes,
                                  _lib.stream_ptr())
    _lib.check(rc, 'gated_bottleneck' if gated else 'the_bottleneck')
    return y


def the_bottleneck(the_input, wide_layer=30, narrow_layer=10, non_dilated_neck_kernel_size=9,
                   dilated_neck_kernel_size=9, dilation_rate=1, is_last_flat=False, *, params, fused=None):
    """nn_core_operator.py:57-79.  params = [(w1,b1), (w2,b2), (w3,b3)].
    On the tensor engine the block runs as the codec program runs it: three launches, the second conv on folded images where the
    frame is long enough (`fused='folded'` / `False` force that / the taps-in-N form; `fused=True`: the opt-in one-launch kernel)."""
    return _block(the_input, wide_layer, narrow_layer, non_dilated_neck_kernel_size, dilated_neck_kernel_size,
                  dilation_rate, is_last_flat, False, params, fused)


def gated_bottleneck(the_input, wide_layer=30, narrow_layer=10, non_dilated_neck_kernel_size=9,
                     dilated_neck_kernel_size=9, dilation_rate=1, is_last_flat=False, the_share=False, *, params):
    """nn_core_operator.py:82-112 (gate kernel size 15 is hard-coded there; `the_share` is unused there too).
    params = [(w_1x1,b), (w_left,b), (w_right,b), (w_out,b)]."""
    return _block(the_input, wide_layer, narrow_layer, non_dilated_neck_kernel_size, dilated_neck_kernel_size,
                  dilation_rate, is_last_flat, True, params)


def gated_bottleneck_decoder(the_input, wide_layer=30, narrow_layer=10, non_dilated_neck_kernel_size=9,
                             dilated_neck_kernel_size=9, dilation_rate=1, is_last_flat=False, the_share=False, *,
                             params):
    """nn_core_operator.py:115-137 -- defined but never called by the reference; composed from the surface ops.
    params = [(w_1x1,b), (w_left,b), (w_right,b), (dw, pw, b)]."""
    y = conv1d(the_input, narrow_layer, 1, dilation_rate=1, activation=None, params=params[0])
    y = activation_func(y)
    left = conv1d(y, narrow_layer, dilated_neck_kernel_size, dilation_rate=dilation_rate, activation=None, params=params[1])
    right = conv1d(y, narrow_layer, dilated_neck_kernel_size, dilation_rate=dilation_rate, activation='tanh', params=params[2])
    y = conv1d_depth(left * right, wide_layer, non_dilated_neck_kernel_size, dilation_rate=1, activation=None,
                     params=params[3])
    y = y + the_input
    return y if is_last_flat else activation_func(y)


def scalar_softmax_quantization(floating_code, alpha, bins, is_quan_on, the_share, code_length, num_kmean_kernels, *,
                                return_indices: bool = False):
    """nn_core_operator.py:140-164.

    floating_code (B, L, 1); alpha 0-d CUDA tensor (or float); bins (n,) CUDA tensor.
    Returns (soft_assignment (B, L, n) -- always the SOFT one, bit_code (B, L, 1)); with
    ``return_indices=True`` additionally the implicit integer code (B, L) uint8 the reference never surfaces.
    ``the_share`` True -> the value path uses the soft assignment, False -> the one-hot (tf.cond, :154-158).
    """
    x = _check_cl(floating_code, 'floating_code')
    B, L, one = x.shape
    if one != 1:
        raise ValueError("floating_code must have one channel")
    if L != code_length:
        raise ValueError(f"code_length {code_length} does not match the tensor ({L})")
    bins = _lib.require_f32(bins, 'bins')
    if bins.numel() != num_kmean_kernels:
        raise ValueError("num_kmean_kernels does not match bins")
    alpha_t = alpha if isinstance(alpha, torch.Tensor) else torch.tensor(float(alpha), dtype=torch.float32, device=x.device)
    alpha_t = _lib.require_f32(alpha_t.reshape(1), 'alpha')
    soft = torch.empty((B, L, num_kmean_kernels), dtype=torch.float32, device=x.device)
    out = torch.empty((B, L, 1), dtype=torch.float32, device=x.device)
    idx = torch.empty((B, L), dtype=torch.uint8, device=x.device)
    rc = _lib.load().nsc_quantize_scalar(_lib.ptr(x), B, L, _lib.ptr(bins), num_kmean_kernels, _lib.ptr(alpha_t),
                                         float(is_quan_on), int(bool(the_share)), _lib.ptr(out), _lib.ptr(idx),
                                         _lib.ptr(soft), None, None, _lib.stream_ptr())
    _lib.check(rc, 'scalar_softmax_quantization')
    if return_indices:
        return soft, out, idx
    return soft, out


def vector_softmax_quantization(floating_code, alpha, bins, is_quan_on, is_share, top_k, code_len):
    """nn_core_operator.py:167-195 -- dead path in the reference (its only caller `one_ae_vq` is commented out,
    cmrl.py:912,920); kept on the surface as a stub."""
    raise NotImplementedError("vector_softmax_quantization is out of the hot path (no caller in the reference)")
