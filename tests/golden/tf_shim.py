"""A minimal stand-in for the TensorFlow 1.x-style API surface that /root/reference/nn_core_operator.py and the graph-building methods
of /root/reference/neural_speech_coding_module.py (:152-335) touch, backed by torch-CPU float32.  TEST INFRASTRUCTURE ONLY, used by
tests/golden/make_ref_golden.py to EXECUTE the reference's own, unmodified graph code (which layers, in which order, with which
activations, residuals, reshapes and permutes) without TensorFlow, which is not installable here.

Written from TensorFlow's documented behaviour, independently of oracle/ (nothing here imports it):
  * tf.compat.v1.layers.conv1d / tf.keras.layers.SeparableConv1D: channels_last, 'SAME' padding = total
    max((ceil(L / s) - 1) s + (k - 1) d + 1 - L, 0) zeros, floor(total / 2) of them on the left; variables are created in call order
    (kernel (k, cin, cout) + bias (cout); depthwise (k, cin, 1) + pointwise (1, cin, cout) + bias) -- here they are DRAWN from a
    caller-supplied list in that order, and a shape mismatch raises: the creation order and the shapes of the reference's graph are
    part of what gets checked;
  * tf.nn.leaky_relu slope 0.2; tf.nn.softmax over the last axis; tf.nn.top_k(x).indices = index of the largest entry, lowest
    index on ties; tf.one_hot; tf.cond on a Python bool; reshape / permute / matmul / expand_dims / cast as named.
"""
import contextlib
import math
import types

import numpy as np
import torch
import torch.nn.functional as F


class VariableFeed:
    """Variables in TF creation order: a list of tuples of numpy arrays ((kernel, bias) or (depthwise, pointwise, bias))."""

    def __init__(self, conv_params):
        self.params = list(conv_params)
        self.cursor = 0
        self.created = []          # shapes, in creation order (what tf.compat.v1.trainable_variables() would list)

    def take(self, shapes):
        if self.cursor >= len(self.params):
            raise AssertionError(f"the reference graph creates more than {len(self.params)} conv layers")
        arrs = self.params[self.cursor]
        self.cursor += 1
        got = tuple(tuple(np.asarray(a).shape) for a in arrs)
        if got != tuple(shapes):
            raise AssertionError(f"layer {self.cursor - 1}: the reference graph creates {shapes}, the supplied variables are {got}")
        self.created.extend(shapes)
        return [torch.as_tensor(np.asarray(a), dtype=torch.float32) for a in arrs]


def draw_layer(rng, shapes):
    """Deterministic variables of one layer from a numpy RandomState: Glorot-uniform kernels, small NON-zero biases (so that a
    misplaced bias shows).  Used by the generator (shapes as the reference graph asks for them) and by the tests (shapes of the
    oracle's layer table): the two agree only if creation order and shapes agree."""
    out = []
    for shp in shapes:
        if len(shp) == 1:
            out.append(rng.uniform(-0.1, 0.1, size=shp).astype(np.float32))
        else:
            fan_in, fan_out = shp[0] * shp[1], shp[0] * shp[2]
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            out.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
    return tuple(out)


class SeededFeed(VariableFeed):
    """Creates every variable on demand, in the order the graph asks for it, from RandomState(seed)."""

    def __init__(self, seed):
        super().__init__([])
        self.rng = np.random.RandomState(seed)
        self.layers = []           # per layer: tuple of shapes

    def take(self, shapes):
        arrs = draw_layer(self.rng, shapes)
        self.params.append(arrs)
        self.layers.append(tuple(shapes))
        self.created.extend(shapes)
        self.cursor += 1
        return [torch.as_tensor(a) for a in arrs]


_FEED = [None]
PLACEHOLDERS = {}          # name -> value, for tf.compat.v1.placeholder


def set_feed(feed):
    _FEED[0] = feed


def _same_pad(L, k, d, s):
    out = -(-L // s)
    total = max((out - 1) * s + (k - 1) * d + 1 - L, 0)
    return total // 2, total - total // 2


def _conv_cl(x, w, b, d, s):
    """x (B, L, Cin) channels-last, w (k, Cin, Cout) -> (B, Lout, Cout), SAME padding."""
    k = w.shape[0]
    pl, pr = _same_pad(x.shape[1], k, d, s)
    xt = F.pad(x.permute(0, 2, 1), (pl, pr))
    y = F.conv1d(xt, w.permute(2, 1, 0).contiguous(), b, stride=s, dilation=d)
    return y.permute(0, 2, 1)


def _layers_conv1d(inputs, filters, padding='valid', kernel_size=1, activation=None, dilation_rate=1, strides=1,
                   data_format='channels_last'):
    assert padding.upper() == 'SAME' and data_format == 'channels_last'
    cin = int(inputs.shape[-1])
    w, b = _FEED[0].take(((int(kernel_size), cin, int(filters)), (int(filters),)))
    y = _conv_cl(inputs, w, b, int(dilation_rate), int(strides))
    return activation(y) if activation is not None else y


class _SeparableConv1D:
    def __init__(self, filters, padding='valid', kernel_size=1, activation=None, dilation_rate=1, strides=1, data_format='channels_last'):
        assert padding.upper() == 'SAME' and data_format == 'channels_last' and int(strides) == 1
        self.f, self.k, self.act, self.d = int(filters), int(kernel_size), activation, int(dilation_rate)

    def __call__(self, inputs):
        cin = int(inputs.shape[-1])
        dw, pw, b = _FEED[0].take(((self.k, cin, 1), (1, cin, self.f), (self.f,)))
        pl, pr = _same_pad(inputs.shape[1], self.k, self.d, 1)
        xt = F.pad(inputs.permute(0, 2, 1), (pl, pr))
        y = F.conv1d(xt, dw[:, :, 0].t().unsqueeze(1).contiguous(), None, dilation=self.d, groups=cin)      # depthwise, multiplier 1
        y = F.conv1d(y, pw[0].t().unsqueeze(2).contiguous(), b).permute(0, 2, 1)                              # pointwise + bias
        return self.act(y) if self.act is not None else y


class _TopK:
    def __init__(self, x):
        # index of the maximum, lowest index on ties (torch.max does not promise a tie rule: make it explicit)
        m = x.max(dim=-1, keepdim=True).values
        n = x.shape[-1]
        idx = torch.where(x == m, torch.arange(n).expand_as(x), torch.full_like(x, n, dtype=torch.long)).min(dim=-1, keepdim=True).values
        self.indices = idx
        self.values = m


class _Var:
    def __init__(self, shape):
        self._s = list(shape)

    def get_shape(self):
        return types.SimpleNamespace(as_list=lambda: list(self._s))

    @property
    def shape(self):
        return tuple(self._s)


def _variable(value, dtype=None, name=None):
    return torch.as_tensor(np.asarray(value, dtype=np.float32))


def build():
    """-> a module object to install as sys.modules['tensorflow']."""
    tf = types.ModuleType('tensorflow')
    tf.float32 = torch.float32
    tf.custom_gradient = lambda f: f
    tf.nn = types.SimpleNamespace(
        tanh=torch.tanh, relu=torch.relu, elu=F.elu,
        leaky_relu=lambda x, alpha=0.2: F.leaky_relu(x, negative_slope=alpha),
        softmax=lambda x, axis=-1: torch.softmax(x, dim=axis),
        top_k=lambda x, k=1: _TopK(x))
    tf.abs = torch.abs
    tf.multiply = lambda a, b: torch.as_tensor(a) * b
    tf.matmul = lambda a, b: torch.matmul(a, b)
    tf.expand_dims = lambda x, axis: torch.as_tensor(x).unsqueeze(axis)
    tf.reshape = lambda x, shape: torch.as_tensor(x).reshape(tuple(int(v) for v in shape))
    tf.cast = lambda x, dtype: torch.as_tensor(x).to(dtype)
    tf.one_hot = lambda idx, depth: F.one_hot(idx.long(), int(depth)).to(torch.float32)
    tf.cond = lambda pred, f1, f2: f1() if bool(pred) else f2()
    tf.Variable = _variable
    tf.constant = _variable
    tf.ones = lambda shape=None, **k: torch.ones(shape)
    tf.sqrt, tf.square, tf.reduce_sum, tf.reduce_mean = torch.sqrt, torch.square, None, None
    v1 = types.SimpleNamespace()
    v1.layers = types.SimpleNamespace(conv1d=_layers_conv1d, batch_normalization=lambda inputs, **k: inputs)
    v1.variable_scope = lambda name, *a, **k: contextlib.nullcontext()
    v1.trainable_variables = lambda *a, **k: [_Var(s) for s in (_FEED[0].created if _FEED[0] else [])]
    # graph-mode plumbing of cmrl.py, executed eagerly: a placeholder IS the value fed under its name; py_func calls the Python
    # function on numpy arrays right away
    v1.placeholder = lambda dtype=None, shape=None, name=None: PLACEHOLDERS[name]

    def _py_func(fn, inp, Tout):
        args = [a.detach().numpy() if isinstance(a, torch.Tensor) else a for a in inp]
        r = fn(*args)
        r = r[0] if isinstance(r, tuple) else r
        return [torch.as_tensor(np.asarray(r, dtype=np.float32))]
    v1.py_func = _py_func
    tf.bool = torch.bool
    tf.compat = types.SimpleNamespace(v1=v1)
    tf.keras = types.SimpleNamespace(
        layers=types.SimpleNamespace(SeparableConv1D=_SeparableConv1D, BatchNormalization=None),
        backend=types.SimpleNamespace(permute_dimensions=lambda x, perm: x.permute(*perm)))
    # ---- loss_terms_and_measures.py:63-84, :130-183, :257-267 (mse / mel / quantisation / entropy terms)
    def _red(fn):
        def f(input_tensor=None, axis=None, **k):
            x = torch.stack(list(input_tensor)) if isinstance(input_tensor, (list, tuple)) else torch.as_tensor(input_tensor)
            return fn(x) if axis is None else fn(x, dim=axis)
        return f
    tf.reduce_mean, tf.reduce_sum = _red(torch.mean), _red(torch.sum)
    tf.square, tf.sqrt, tf.subtract, tf.sign = torch.square, torch.sqrt, lambda a, b: a - b, torch.sign
    tf.concat = lambda xs, axis=0: torch.cat(list(xs), dim=axis)
    tf.math = types.SimpleNamespace(real=lambda z: z.real, imag=lambda z: z.imag,
                                    log=lambda x: torch.log(torch.as_tensor(x, dtype=torch.float32)))

    def _stft(signals, frame_length, frame_step, fft_length=None, window_fn=None, pad_end=False):
        # window_fn=None: rectangular window; frames of frame_length every frame_step samples, no padding; rfft of fft_length
        assert window_fn is None and not pad_end
        fr = signals.unfold(-1, int(frame_length), int(frame_step))
        return torch.fft.rfft(fr, n=int(fft_length or frame_length), dim=-1)

    def _mel_matrix(num_mel_bins=20, num_spectrogram_bins=129, sample_rate=8000, lower_edge_hertz=125.0, upper_edge_hertz=3800.0, dtype=None):
        # tf.signal.linear_to_mel_weight_matrix as documented: HTK mel scale 1127 ln(1 + f / 700); the DC bin is dropped from the
        # triangle computation and re-added as a zero row; triangles max(0, min(lower slope, upper slope)) between band edges that
        # are linearly spaced in mel (num_mel_bins + 2 of them)
        def mel(f):
            return 1127.0 * np.log1p(np.asarray(f, np.float64) / 700.0)
        nyq = sample_rate / 2.0
        freqs = np.linspace(0.0, nyq, int(num_spectrogram_bins))[1:]
        sm = mel(freqs)[:, None]
        edges = np.linspace(mel(lower_edge_hertz), mel(upper_edge_hertz), int(num_mel_bins) + 2)
        lo, ce, up = edges[None, :-2], edges[None, 1:-1], edges[None, 2:]
        w = np.maximum(0.0, np.minimum((sm - lo) / (ce - lo), (up - sm) / (up - ce)))
        return torch.as_tensor(np.pad(w, [[1, 0], [0, 0]]).astype(np.float32))

    v2 = types.SimpleNamespace(signal=types.SimpleNamespace(stft=_stft, linear_to_mel_weight_matrix=_mel_matrix, rfft=torch.fft.rfft))
    tf.compat.v2 = v2
    tf.signal = v2.signal
    return tf
