"""Pins against the REFERENCE'S OWN CODE run in this container (tests/golden/make_ref_golden.py): the unmodified
/root/reference/utilities.py and lpc_utilities.py executed from their source files with the absent third-party libraries replaced
by independent behavioural stand-ins (scipy lfilter for audiolazy.ZFilter, solve_toeplitz for audiolazy.lpc, numpy.roots / numpy.poly
for spectrum.poly2lsf / lsf2poly).  Everything the reference itself wrote around those calls -- window constructions, hop
arithmetic, the flatten quirk, sub-frame weighting, frame loops -- ran for real.

  * always:   oracle/ against the committed fixture (tests/golden/reference_run.npz);
  * here:     the generator re-run against /root/reference reproduces the committed fixture (skipped on the GPU box, which has no
              reference tree);
  * -m gpu:   the CUDA path, through the C ABI, against the same fixture (tests/test_gpu_parity.py::test_cuda_vs_reference_run)."""
import os
import sys

import numpy as np
import pytest

from oracle import ref_framing, ref_lpc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
FIX = os.path.join(GOLD, 'reference_run.npz')


def _fix():
    return dict(np.load(FIX))


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason="reference tree not present on this box")
def test_generator_reproduces_the_committed_fixture():
    sys.path.insert(0, GOLD)
    try:
        import make_ref_golden
        fresh = make_ref_golden.generate()
    finally:
        sys.path.remove(GOLD)
    g = _fix()
    assert sorted(fresh) == sorted(g)
    for k in g:
        assert fresh[k].shape == g[k].shape and _rel(fresh[k], g[k]) < 1e-9, k        # (root finders: not bit-stable across BLAS builds)


def test_framing_oracle_equals_the_reference_run():
    """utilities.py:7-39 executed by the reference itself."""
    g = _fix()
    sig = g['utt']
    assert np.array_equal(ref_framing.utterance_to_segment(sig, True), g['seg_plain'])
    assert _rel(ref_framing.utterance_to_segment(sig, False), g['seg_windowed']) < 1e-15
    n = g['seg_plain'].shape[0]
    for k, i in (('hann_first', 0), ('hann_mid', 3), ('hann_last', n - 1)):
        assert _rel(ref_framing.hann_process(g['seg_plain'][i], i, n), g[k]) < 1e-15, k


def test_filters_oracle_equals_the_reference_run():
    """lpc_utilities.py:8-11: the module-level high-pass and pre-emphasis filters on a whole signal."""
    g = _fix()
    x = g['utt'].astype(np.float64)
    assert _rel(ref_lpc.highpass_filter(x), g['highpass']) < 1e-10
    assert _rel(ref_lpc.empha_filter(x), g['empha']) < 1e-12
    assert _rel(ref_framing.highpass_filter(x), g['highpass']) < 1e-10


def test_lpc_analysis_oracle_equals_the_reference_run():
    """lpc_analysis_at_test (flatten quirk, 1024-sample windows at hop 512, trapezoid-Hann window; lpc_utilities.py:94-129) and
    lpc_analysis_at_train (per-frame high-pass + emphasis, :14-25).  The reference's own code ran; LPC solve and root finding behind
    it were scipy / numpy.roots, the oracle uses Levinson-Durbin and its own LSF routine -- two independent routes to the same LSFs."""
    g = _fix()
    lsf = ref_lpc.lpc_analysis_at_test(g['at_test_in'], 16)
    assert lsf.shape == g['at_test_lsf'].shape == (6, 16)            # 4 x 1024 samples flattened -> 6 windows, not 4
    assert np.abs(lsf - g['at_test_lsf']).max() < 1e-7
    tr = ref_lpc.lpc_analysis_at_train(g['at_train_in'][:, :, None], 16)
    assert np.abs(tr - g['at_train_lsf']).max() < 1e-7


def test_lsf2poly_residual_synthesis_oracle_equals_the_reference_run():
    """lsf2poly_after_quan (:28-33), lpc_analysis_get_residual (seven half-overlapping 128-sample sub-frames, each filtered from
    zero state and Hann-weighted, :37-77) and lpc_synthesizer_tr (:137-156), chained exactly as cmrl.py:793-843 chains them."""
    g = _fix()
    lsf32 = g['at_train_lsf'].astype(np.float32)
    poly = ref_lpc.lsf2poly_after_quan(lsf32, 16)
    assert poly.dtype == np.float32 and _rel(poly, g['poly']) < 2e-6
    res = ref_lpc.lpc_analysis_get_residual(g['at_train_in'][:, :, None], g['poly'])
    assert res.dtype == np.float32 and _rel(res, g['residual']) < 1e-6
    syn = ref_lpc.lpc_synthesizer_tr(g['poly'], g['residual'])
    syn = syn[0] if isinstance(syn, tuple) else syn
    assert _rel(syn, g['synth']) < 1e-6
    # the sub-frame weights sum to one, so residual -> synthesis returns the frame (up to the sub-frame state resets)
    assert _rel(g['synth'], g['at_train_in']) < 0.5
