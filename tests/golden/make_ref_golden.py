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


if __name__ == '__main__':
    vec = generate()
    np.savez_compressed(OUT, **vec)
    print('wrote', OUT, {k: v.shape for k, v in vec.items()})
