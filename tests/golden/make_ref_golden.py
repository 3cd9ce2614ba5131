"""Golden vectors from the REFERENCE'S OWN CODE, run here: tests/golden/reference_run.npz.

The reference (/root/reference) cannot run as shipped -- TensorFlow, audiolazy, spectrum, pystoi, soundfile and mdct are not
installed and there is no network.  But the bodies of its framing / LPC helpers are plain numpy around a handful of third-party
calls.  This script imports the reference's *unmodified* `utilities.py` and `lpc_utilities.py` from where they lie, with those
third-party modules replaced by small behavioural stand-ins written here from the libraries' documented behaviour -- INDEPENDENTLY
of oracle/ (nothing below imports it):

  audiolazy.ZFilter(list)       a linear time-invariant filter object: calling it on a sequence yields the zero-state response;
                                `a / b`, `1 / a` compose transfer functions; `.numlist`          -> scipy.signal.lfilter
  audiolazy.lpc(block, order)   autocorrelation-method LPC ("autocor" is audiolazy's default strategy): numlist = [1, a_1 .. a_p]
                                solving the Toeplitz normal equations                           -> scipy.linalg.solve_toeplitz
  spectrum.poly2lsf / lsf2poly  textbook sum / difference polynomial roots                      -> numpy.roots / numpy.poly
  tensorflow                    only `tf.custom_gradient` is touched at import (identity decorator here); nothing TF runs

What these vectors pin is therefore the reference's OWN code -- window constructions, sub-frame weighting, hop arithmetic, the
flatten quirk of lpc_analysis_at_test, frame loops -- executed for real; the third-party semantics stay [LIB] assumptions, now
with a second, independent implementation behind them.  tests/test_reference_run_pins.py checks oracle/ against the fixture
(always) and re-runs this generator against /root/reference when it is present (here; not on the GPU box).

    python tests/golden/make_ref_golden.py            # rewrites tests/golden/reference_run.npz
"""
import importlib
import os
import sys
import types

import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_run.npz')


# ------------------------------------------------------------------------------------------------ third-party stand-ins
class ZFilter:
    """audiolazy.ZFilter as the reference uses it: ZFilter(list) is the FIR filter with those taps; division builds rational
    transfer functions; calling the filter on a sequence returns (an iterable of) the zero-state response."""

    def __init__(self, num, den=None):
        self.numlist = [float(v) for v in num]
        self.denlist = [1.0] if den is None else [float(v) for v in den]

    def __call__(self, seq):
        from scipy.signal import lfilter
        return lfilter(np.asarray(self.numlist, np.float64), np.asarray(self.denlist, np.float64), np.asarray(list(seq), np.float64))

    def __truediv__(self, other):
        if isinstance(other, ZFilter):
            return ZFilter(np.convolve(self.numlist, other.denlist), np.convolve(self.denlist, other.numlist))
        return ZFilter(np.asarray(self.numlist) / other, self.denlist)

    def __rtruediv__(self, other):          # number / filter
        return ZFilter(np.asarray(self.denlist) * other, self.numlist)


def lpc(block, order):
    from scipy.linalg import solve_toeplitz
    x = np.asarray(list(block), np.float64)
    r = np.array([np.dot(x[:len(x) - k], x[k:]) for k in range(order + 1)])
    a = solve_toeplitz(r[:-1], -r[1:])
    return ZFilter(np.concatenate([[1.0], a]))


def poly2lsf(a):
    a = np.asarray(a, np.float64)
    a = a / a[0]
    p = np.concatenate([a, [0.0]]) + np.concatenate([[0.0], a[::-1]])      # sum polynomial (root at z = -1)
    q = np.concatenate([a, [0.0]]) - np.concatenate([[0.0], a[::-1]])      # difference polynomial (root at z = +1)
    ang = np.concatenate([np.angle(np.roots(p)), np.angle(np.roots(q))])
    ang = np.sort(ang[(ang > 1e-9) & (ang < np.pi - 1e-9)])
    return ang


def lsf2poly(lsf):
    lsf = np.asarray(lsf, np.float64)
    z = np.exp(1j * lsf)
    rp, rq = z[0::2], z[1::2]
    p = np.poly(np.concatenate([rp, rp.conj()]))
    q = np.poly(np.concatenate([rq, rq.conj()]))
    p = np.convolve(p, [1.0, 1.0])          # even order: P carries the root at -1, Q the root at +1
    q = np.convolve(q, [1.0, -1.0])
    return (0.5 * (p + q)).real[:-1]


class _Anything(types.ModuleType):
    """A module whose every attribute is callable and returns another stand-in (nothing of it is ever executed for values)."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Anything(name)

    def __call__(self, *a, **k):
        return _Anything('call')


def install_stubs():
    tf = _Anything('tensorflow')
    tf.custom_gradient = lambda f: f
    mods = {'tensorflow': tf, 'tensorflow.python': _Anything('tensorflow.python'),
            'tensorflow.python.framework': _Anything('tensorflow.python.framework'),
            'tensorflow.python.framework.ops': _Anything('ops'), 'tensorflow_probability': _Anything('tfp'),
            'mdct': _Anything('mdct'), 'pystoi': _Anything('pystoi'), 'pystoi.stoi': _Anything('pystoi.stoi'),
            'soundfile': _Anything('soundfile'), 'pesq': _Anything('pesq'), 'pypesq': _Anything('pypesq')}
    mods['pystoi.stoi'].stoi = lambda *a, **k: 0.0
    al = types.ModuleType('audiolazy')
    al.ZFilter, al.lpc = ZFilter, lpc
    al.__all__ = ['ZFilter', 'lpc']
    sp = types.ModuleType('spectrum')
    sp.poly2lsf, sp.lsf2poly = poly2lsf, lsf2poly
    mods['audiolazy'], mods['spectrum'] = al, sp
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    return saved


def load_reference():
    """-> (utilities, lpc_utilities) modules of the reference, executed from their own source files."""
    saved = install_stubs()
    own = {k: sys.modules.pop(k, None) for k in ('utilities', 'lpc_utilities', 'constants', 'loss_terms_and_measures')}
    sys.path.insert(0, REF)
    try:
        lu = importlib.import_module('lpc_utilities')
        ut = importlib.import_module('utilities')
    finally:
        sys.path.remove(REF)
        for k in ('utilities', 'lpc_utilities', 'constants', 'loss_terms_and_measures'):
            sys.modules.pop(k, None)
            if own[k] is not None:
                sys.modules[k] = own[k]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert os.path.dirname(os.path.abspath(lu.__file__)) == REF and os.path.dirname(os.path.abspath(ut.__file__)) == REF
    return ut, lu


# ------------------------------------------------------------------------------------------------ inputs (seeded, self-contained)
def ar_signal(n, seed):
    """A speech-like AR(2) + noise signal (so that the LPC systems are well conditioned)."""
    rng = np.random.RandomState(seed)
    e = rng.randn(n + 64)
    y = np.zeros(n + 64)
    for i in range(2, n + 64):
        y[i] = 1.6 * y[i - 1] - 0.8 * y[i - 2] + e[i]
    y = y[64:]
    return (0.1 * y / np.abs(y).max()).astype(np.float32)


def generate():
    ut, lu = load_reference()
    out = {}
    # utilities.py:7-22, :25-39 -- framing windows
    sig = ar_signal(512 + 480 * 9 + 137, 11)
    out['utt'] = sig
    out['seg_windowed'] = ut.utterance_to_segment(sig, False)
    out['seg_plain'] = ut.utterance_to_segment(sig, True)
    n = out['seg_plain'].shape[0]
    out['hann_first'] = ut.hann_process(out['seg_plain'][0], 0, n)
    out['hann_mid'] = ut.hann_process(out['seg_plain'][3], 3, n)
    out['hann_last'] = ut.hann_process(out['seg_plain'][n - 1], n - 1, n)
    # lpc_utilities.py:8-11 -- the two module-level filters (zero state, whole signal)
    out['highpass'] = np.asarray(list(lu.highpass_filter(sig.astype(np.float64))))
    out['empha'] = np.asarray(list(lu.empha_filter(sig.astype(np.float64))))
    # lpc_utilities.py:94-129 -- window cutting (flatten quirk: a (rows, cols) input is flattened), trapezoid-Hann window, LSFs
    raw = ar_signal(1024 * 3 + 512 * 2, 12).reshape(4, 1024)
    out['at_test_in'] = raw
    out['at_test_lsf'] = lu.lpc_analysis_at_test(raw, 16)
    # lpc_utilities.py:14-25 -- per-frame analysis at train time (high-pass + emphasis per frame, no window)
    fr = np.stack([ar_signal(512, 20 + i) for i in range(5)])[:, :, None]
    out['at_train_in'] = fr[:, :, 0]
    out['at_train_lsf'] = lu.lpc_analysis_at_train(fr, 16)
    # lpc_utilities.py:28-33, :37-77, :137-156 -- lsf -> poly, sub-framed residual, synthesis
    lsf32 = out['at_train_lsf'].astype(np.float32)
    out['poly'] = lu.lsf2poly_after_quan(lsf32, 16)
    out['residual'] = lu.lpc_analysis_get_residual(fr, out['poly'])
    syn = lu.lpc_synthesizer_tr(out['poly'], out['residual'])
    out['synth'] = syn[0] if isinstance(syn, tuple) else syn
    return out


# ------------------------------------------------------------------------------------------------ the neural part
OUT_NN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_run_nn.npz')
TOPOLOGIES = [('bottleneck', (2,)), ('gln', (2,)), ('bottleneck', (2, 2)), ('gln', (2, 2))]


def load_reference_nn():
    """-> (nn_core_operator, neural_speech_coding_module) of the reference, executed from their own source files on tf_shim."""
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import tf_shim
    sys.path.remove(here)
    saved = install_stubs()
    tf = tf_shim.build()
    extra = {'tensorflow': tf, 'librosa': _Anything('librosa'),
             'tensorflow.python.ops': _Anything('ops'), 'tensorflow.python.ops.math_ops': _Anything('math_ops'),
             'tensorflow.python.ops.random_ops': _Anything('random_ops'), 'tensorflow.python.framework.dtypes': _Anything('dtypes')}
    saved.update({k: sys.modules.get(k) for k in extra if k not in saved})
    sys.modules.update(extra)
    names = ('utilities', 'lpc_utilities', 'constants', 'loss_terms_and_measures', 'nn_core_operator', 'neural_speech_coding_module', 'cmrl')
    own = {k: sys.modules.pop(k, None) for k in names}
    sys.path.insert(0, REF)
    try:
        nn = importlib.import_module('nn_core_operator')
        nscm = importlib.import_module('neural_speech_coding_module')
        nscm.cmrl_module = importlib.import_module('cmrl')
    finally:
        sys.path.remove(REF)
        for k in names:
            sys.modules.pop(k, None)
            if own[k] is not None:
                sys.modules[k] = own[k]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert os.path.dirname(os.path.abspath(nn.__file__)) == REF and os.path.dirname(os.path.abspath(nscm.__file__)) == REF
    return tf_shim, nn, nscm


def nn_inputs():
    x = np.stack([ar_signal(512, 40 + i) for i in range(3)]) * 3.0
    x[1] *= 0.05                         # one quiet frame
    return x.astype(np.float32)


def generate_nn():
    import contextlib
    import io
    import torch
    tf_shim, nn, nscm = load_reference_nn()
    out = {'x': nn_inputs()}
    xt = torch.from_numpy(out['x'])[:, :, None]
    sink = io.StringIO()                 # the reference prints shapes while it builds the graph
    for ti, (rt, strides) in enumerate(TOPOLOGIES):
        # the reference selects the block type by a module-level switch it star-imports from constants.py (constants.py:13-14)
        nscm.resnet_type = rt
        m = object.__new__(nscm.neuralSpeechCodingModule)          # __init__ loads the training corpus from the author's disk
        m._bottleneck_kernel_and_dilation = [9, 9, 100, 20, 1, 2]  # README.md:75
        captured = {}
        real_q = nn.scalar_softmax_quantization

        def spy(floating_code, alpha, bins, is_quan_on, the_share, code_length, n, _c=captured):
            _c['floating'] = floating_code
            soft, code = real_q(floating_code, alpha, bins, is_quan_on, the_share, code_length, n)
            _c['soft'], _c['code'] = soft, code
            return soft, code

        nscm.scalar_softmax_quantization = spy
        for share in (False, True):
            feed = tf_shim.SeededFeed(seed=100 + ti)
            tf_shim.set_feed(feed)
            with contextlib.redirect_stdout(sink), torch.no_grad():
                r = m.computational_graph_end2end_quan_on(xt, share, 1.0, 32, 'scope_1', list(strides))
            tag = f"{rt}_{len(strides)}_{'soft' if share else 'hard'}"
            out[tag + '_out'] = r[4].numpy()
            out[tag + '_code0'] = r[3].numpy()                      # the_final_code[0, :, 0]: what the reference returns
            out[tag + '_floating'] = captured['floating'][:, :, 0].numpy()
            out[tag + '_code'] = captured['code'][:, :, 0].numpy()
            out[tag + '_softsum'] = captured['soft'].sum(0).sum(0).numpy()
        out[f"{rt}_{len(strides)}_layers"] = np.array(repr(feed.layers))
        nscm.scalar_softmax_quantization = real_q
    # single functions of nn_core_operator.py on their own
    rng = np.random.RandomState(7)
    xb = torch.from_numpy(rng.randn(2, 128, 100).astype(np.float32))
    out['block_x'] = xb.numpy()
    for name, fn, kw in (('the_bottleneck', nn.the_bottleneck, dict(wide_layer=100, narrow_layer=20, dilation_rate=2, is_last_flat=False)),
                         ('the_bottleneck_flat', nn.the_bottleneck, dict(wide_layer=100, narrow_layer=20, dilation_rate=1, is_last_flat=True)),
                         ('gated_bottleneck', nn.gated_bottleneck, dict(wide_layer=100, narrow_layer=20, dilation_rate=2, is_last_flat=False)),
                         ('gated_bottleneck_decoder', nn.gated_bottleneck_decoder, dict(wide_layer=100, narrow_layer=20, dilation_rate=1, is_last_flat=True))):
        feed = tf_shim.SeededFeed(seed=200)
        tf_shim.set_feed(feed)
        with torch.no_grad():
            out['block_' + name] = fn(xb, **kw).numpy()
        out['block_' + name + '_layers'] = np.array(repr(feed.layers))
    for name, fn, kw in (('conv1d_s2', nn.conv1d, dict(num_filters=24, filter_size=9, strides=2, dilation_rate=1)),
                         ('conv1d_d3', nn.conv1d, dict(num_filters=8, filter_size=5, strides=1, dilation_rate=3, activation=None)),
                         ('conv1d_depth', nn.conv1d_depth, dict(num_filters=50, filter_size=9, activation=None)),
                         ('change_channel', nn.change_channel, dict(the_channel=1, kernel_size=55, dilation_rate=7))):   # dilation is ignored (:52)
        feed = tf_shim.SeededFeed(seed=300)
        tf_shim.set_feed(feed)
        with torch.no_grad():
            out['op_' + name] = fn(xb, **kw).numpy()
    # loss terms (loss_terms_and_measures.py:63-84, :130-183, :257-267), star-imported into nn_core_operator's namespace
    ori = np.stack([ar_signal(512, 60 + i) for i in range(5)]) * 4.0
    dec = (ori + 0.02 * rng.randn(5, 512)).astype(np.float32)
    out['loss_ori'], out['loss_dec'] = ori.astype(np.float32), dec
    with contextlib.redirect_stdout(sink), torch.no_grad():
        out['loss_mse'] = nn.mse_loss(torch.from_numpy(dec), torch.from_numpy(out['loss_ori'])).numpy()
        out['loss_mfcc'] = nn.mfcc_loss(torch.from_numpy(dec), torch.from_numpy(out['loss_ori'])).numpy()
        soft = torch.softmax(torch.from_numpy(rng.randn(3, 64, 32).astype(np.float32)) * 3.0, dim=-1)
        out['loss_soft'] = soft.numpy()
        out['loss_quan'] = nn.quan_loss(soft).numpy()
        out['loss_ent'] = np.asarray(nn.entropy_coding_loss(soft).numpy())
        out['bitrate'] = np.array([nn.entropy_to_bitrate(2.5, 2), nn.entropy_to_bitrate(2.5, 4)], np.float64)
    # the cascade graphs of cmrl.py, built by the reference's own CMRL.all_modules_feedforward (:513-543) and
    # all_modules_feedforward_lpc (:770-830) -- placeholders fed eagerly, tf.py_func bodies from the reference's lpc_utilities.py --
    # followed by the two lines of _feedforward_lpc that finish the pass (:836-839: decoded = sum of the codec outputs, synthesis)
    cm_mod = nscm.cmrl_module
    nscm.resnet_type = 'bottleneck'
    cq_x = np.stack([ar_signal(512, 80 + i) for i in range(4)]) * 5.0
    out['cq_x'] = cq_x.astype(np.float32)
    lsf = cm_mod.lpc_analysis_at_train(out['cq_x'][:, :, None], 16).astype(np.float32)
    out['cq_lsf'] = lsf
    for share in (False, True):
        c = object.__new__(cm_mod.CMRL)
        c._bottleneck_kernel_and_dilation = [9, 9, 100, 20, 1, 2]
        c._res_scalar, c._num_resnets, c._the_strides, c._num_bins_for_follower, c._lpc_order = 2.0, 2, [2], [32, 32], 16
        tf_shim.PLACEHOLDERS.clear()
        tf_shim.PLACEHOLDERS.update({'x': torch.from_numpy(out['cq_x'])[:, :, None], 'x_': torch.from_numpy(out['cq_x'])[:, :, None],
                                     'lr': 0.0, 'the_share': share, 'tau': 0.0, 'is_quan_on': 1.0,
                                     'lpc_x': torch.from_numpy(lsf)[:, :, None]})
        tag = 'cq_soft' if share else 'cq_hard'
        tf_shim.set_feed(tf_shim.SeededFeed(seed=400))
        with contextlib.redirect_stdout(sink), torch.no_grad():
            r = c.all_modules_feedforward_lpc(2)
        res_x, poly, soft_lpc, outs = r[4], r[5], r[9], r[15]
        decoded = np.sum([o.numpy() for o in outs], axis=0)
        syn = cm_mod.lpc_synthesizer_tr(poly.numpy(), decoded)
        out[tag + '_poly'], out[tag + '_res_x'] = poly.numpy(), res_x[:, :, 0].numpy()
        out[tag + '_lsf_idx'] = soft_lpc.argmax(-1).numpy().astype(np.int64)
        out[tag + '_outs'] = np.stack([o.numpy() for o in outs])
        out[tag + '_decoded'] = decoded
        out[tag + '_synth'] = syn[0] if isinstance(syn, tuple) else syn
        # plain cascade (no LPC): codec 0 is neither scaled nor divided
        tf_shim.set_feed(tf_shim.SeededFeed(seed=400))
        with contextlib.redirect_stdout(sink), torch.no_grad():
            r = c.all_modules_feedforward(2)
        out[tag + '_plain_outs'] = np.stack([o.numpy() for o in r[9]])
    # quantiser incl. exact ties (mid-points between bins) and out-of-range values
    bins = np.linspace(-1, 1, 32).astype(np.float32)
    fc = rng.uniform(-1.2, 1.2, size=(2, 256, 1)).astype(np.float32)
    fc[0, :31, 0] = (bins[:-1] + bins[1:]) / 2
    out['q_in'] = fc[:, :, 0]
    for share in (False, True):
        with contextlib.redirect_stdout(sink), torch.no_grad():
            soft, code = nn.scalar_softmax_quantization(torch.from_numpy(fc), torch.tensor(-300.0), torch.from_numpy(bins), 1.0, share, 256, 32)
        out['q_code_' + ('soft' if share else 'hard')] = code[:, :, 0].numpy()
        out['q_soft_argmax'] = soft.argmax(-1).numpy().astype(np.int64)
    return out


if __name__ == '__main__':
    vec = generate()
    np.savez_compressed(OUT, **vec)
    print('wrote', OUT, {k: v.shape for k, v in vec.items()})
    vec = generate_nn()
    np.savez_compressed(OUT_NN, **vec)
    print('wrote', OUT_NN, {k: v.shape for k, v in vec.items() if not k.endswith('_layers')})
