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


# ------------------------------------------------------------------------------------------------ the neural part
import torch

from oracle import ref_codec, ref_nn

FIX_NN = os.path.join(GOLD, 'reference_run_nn.npz')
TOPOLOGIES = [('bottleneck', (2,)), ('gln', (2,)), ('bottleneck', (2, 2)), ('gln', (2, 2))]


def _shim():
    sys.path.insert(0, GOLD)
    try:
        import tf_shim
    finally:
        sys.path.remove(GOLD)
    return tf_shim


def seeded_params_like(conv_params, seed):
    """Variables drawn from RandomState(seed) in the ORACLE'S creation order and shapes -- the generator drew them in the REFERENCE
    graph's creation order and shapes, so the two sets are identical only if order and shapes agree layer by layer."""
    rng = np.random.RandomState(seed)
    draw = _shim().draw_layer
    return [draw(rng, tuple(tuple(np.asarray(a).shape) for a in layer)) for layer in conv_params]


def oracle_codec_for(rt, strides, ti):
    cfg = ref_codec.OracleCodecCfg(resnet_type=rt, strides=strides)
    shapes = ref_codec.OracleCodec(cfg, seed=0).conv_params
    return ref_codec.OracleCodec(cfg, conv_params=seeded_params_like(shapes, 100 + ti))


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason="reference tree not present on this box")
def test_nn_generator_reproduces_the_committed_fixture():
    sys.path.insert(0, GOLD)
    try:
        import make_ref_golden
        fresh = make_ref_golden.generate_nn()
    finally:
        sys.path.remove(GOLD)
    g = dict(np.load(FIX_NN))
    assert sorted(fresh) == sorted(g)
    for k in g:
        if k.endswith('_layers'):
            assert str(fresh[k]) == str(g[k]), k
        else:
            assert np.allclose(fresh[k], g[k], rtol=0, atol=1e-6), k


@pytest.mark.parametrize('ti', range(4), ids=['bottleneck_s2', 'gln_s2', 'bottleneck_s4', 'gln_s4'])
def test_codec_graph_oracle_equals_the_reference_run(ti):
    """computational_graph_end2end_quan_on (nscm.py:262-295) built by the reference's OWN code -- _the_encoder_in_each_module,
    _stack_bottleneck_blocks, the_bottleneck / gated_bottleneck, _down_sampling_mod, _up_sampling_mod(_helper),
    scalar_softmax_quantization, _the_decoder_in_each_module -- on the TF stand-in, vs the oracle's restatement of it, same weights.
    Layer count, creation order and shapes must match (the variables are drawn from one seeded stream in creation order);
    floating code, codes and decoder output must agree to float32 rounding."""
    rt, strides = TOPOLOGIES[ti]
    g = dict(np.load(FIX_NN))
    oc = oracle_codec_for(rt, strides, ti)
    layers = eval(str(g[f"{rt}_{len(strides)}_layers"]))
    assert [tuple(tuple(np.asarray(a).shape) for a in layer) for layer in oc.conv_params] == [tuple(l) for l in layers]
    x = torch.from_numpy(g['x'])[:, :, None]
    for share in (False, True):
        tag = f"{rt}_{len(strides)}_{'soft' if share else 'hard'}"
        with torch.no_grad():
            r = oc.forward(x, share, 1.0)
        fl = r['floating_code'][:, :, 0].numpy()
        assert _rel(fl, g[tag + '_floating']) < 2e-6, tag
        # identical floating codes -> the hard codes must be the same bins; where the float32 codes differ in the last bits the
        # argmax may sit on the other side of a mid-point, so compare on the entries whose floating codes agree exactly
        code = r['code'][:, :, 0].numpy()
        if share:
            assert _rel(code, g[tag + '_code']) < 2e-3, tag          # soft code: slope ~ alpha * bin spacing amplifies 1e-6
        else:
            same = fl == g[tag + '_floating']
            assert same.mean() > 0.5 and np.array_equal(code[same], g[tag + '_code'][same]), tag
            assert (code != g[tag + '_code']).mean() < 0.01, tag
        assert np.array_equal(code[0], g[tag + '_code0']) or not share and (code[0] != g[tag + '_code0']).mean() < 0.01
        # decoder on the REFERENCE RUN'S codes (so that a flipped code cannot hide or fake a decoder difference)
        oc.ps._cursor = 0
        with torch.no_grad():
            oc.encoder(x)
            out = oc.decoder(torch.from_numpy(g[tag + '_code'])[:, :, None])[:, :, 0].numpy()
        assert _rel(out, g[tag + '_out']) < 5e-6, tag


def test_blocks_and_ops_oracle_equals_the_reference_run():
    """nn_core_operator.py's functions one by one, executed by the reference: conv1d (stride 2; dilation 3), conv1d_depth,
    change_channel (which IGNORES its dilation_rate argument, :52), the_bottleneck (both is_last_flat), gated_bottleneck,
    gated_bottleneck_decoder, scalar_softmax_quantization (exact mid-point ties, out-of-range values)."""
    g = dict(np.load(FIX_NN))
    xb = torch.from_numpy(g['block_x'])
    draw = _shim().draw_layer

    def ps_like(layers, seed):
        rng = np.random.RandomState(seed)
        return ref_nn.ParamStream([draw(rng, tuple(l)) for l in layers])

    for name, fn, kw in (('the_bottleneck', ref_nn.the_bottleneck, dict(wide_layer=100, narrow_layer=20, dilation_rate=2, is_last_flat=False)),
                         ('the_bottleneck_flat', ref_nn.the_bottleneck, dict(wide_layer=100, narrow_layer=20, dilation_rate=1, is_last_flat=True)),
                         ('gated_bottleneck', ref_nn.gated_bottleneck, dict(wide_layer=100, narrow_layer=20, dilation_rate=2, is_last_flat=False)),
                         ('gated_bottleneck_decoder', ref_nn.gated_bottleneck_decoder, dict(wide_layer=100, narrow_layer=20, dilation_rate=1, is_last_flat=True))):
        layers = eval(str(g['block_' + name + '_layers']))
        with torch.no_grad():
            y = fn(xb, ps=ps_like(layers, 200), **kw).numpy()
        assert _rel(y, g['block_' + name]) < 2e-6, name
    for name, fn, kw, layers in (
            ('conv1d_s2', ref_nn.conv1d, dict(num_filters=24, filter_size=9, strides=2, dilation_rate=1), [((9, 100, 24), (24,))]),
            ('conv1d_d3', ref_nn.conv1d, dict(num_filters=8, filter_size=5, strides=1, dilation_rate=3, activation=None), [((5, 100, 8), (8,))]),
            ('conv1d_depth', ref_nn.conv1d_depth, dict(num_filters=50, filter_size=9, activation=None), [((9, 100, 1), (1, 100, 50), (50,))]),
            ('change_channel', ref_nn.change_channel, dict(the_channel=1, kernel_size=55, dilation_rate=7), [((55, 100, 1), (1,))])):
        with torch.no_grad():
            y = fn(xb, ps=ps_like(layers, 300), **kw).numpy()
        assert y.shape == g['op_' + name].shape and _rel(y, g['op_' + name]) < 2e-6, name
    bins = np.linspace(-1, 1, 32).astype(np.float32)
    fc = torch.from_numpy(g['q_in'])[:, :, None]
    for share in (False, True):
        soft, code = ref_nn.scalar_softmax_quantization(fc, np.float32(-300.0), bins, 1.0, share, 256, 32)
        assert np.array_equal(code[:, :, 0].numpy(), g['q_code_' + ('soft' if share else 'hard')]) or \
            _rel(code[:, :, 0].numpy(), g['q_code_' + ('soft' if share else 'hard')]) < 1e-6
    idx = ref_nn.quantizer_indices(fc, np.float32(-300.0), bins).numpy().astype(np.int64).reshape(g['q_soft_argmax'].shape)
    assert np.array_equal(idx, g['q_soft_argmax'])            # the implicit integer code: argmax of the literal softmax, lowest index on ties


def test_loss_terms_oracle_equals_the_reference_run():
    """mse_loss, mfcc_loss (rectangular-window STFT of the 512-sample frame, power / 512, four HTK mel resolutions 8 / 16 / 32 / 128,
    log, per-resolution RMS distance, mean), quan_loss, entropy_coding_loss and entropy_to_bitrate, executed by the reference
    (loss_terms_and_measures.py:63-84, :130-183, :257-267)."""
    from oracle import ref_loss
    g = dict(np.load(FIX_NN))
    dec, ori = torch.from_numpy(g['loss_dec']), torch.from_numpy(g['loss_ori'])
    assert _rel(ref_loss.mse_loss(dec, ori).numpy(), g['loss_mse']) < 1e-6
    assert _rel(ref_loss.mfcc_loss(dec, ori).numpy(), g['loss_mfcc']) < 1e-5
    soft = torch.from_numpy(g['loss_soft'])
    assert _rel(ref_loss.quan_loss(soft).numpy(), g['loss_quan']) < 1e-6
    assert abs(float(ref_loss.entropy_coding_loss(soft)) - float(g['loss_ent'])) < 1e-5
    assert np.allclose([ref_loss.entropy_to_bitrate(2.5, 2), ref_loss.entropy_to_bitrate(2.5, 4)], g['bitrate'], rtol=1e-12)


def _cq_oracle_codecs():
    cfg = ref_codec.OracleCodecCfg()
    shapes = ref_codec.OracleCodec(cfg, seed=0).conv_params
    rng = np.random.RandomState(400)                      # ONE stream across both scopes, as the reference graph creates them
    draw = _shim().draw_layer
    ocs = []
    for _ in range(2):
        ocs.append(ref_codec.OracleCodec(cfg, conv_params=[draw(rng, tuple(tuple(np.asarray(a).shape) for a in layer)) for layer in shapes]))
    return ocs


@pytest.mark.parametrize('share', [False, True], ids=['hard', 'soft'])
def test_cascade_and_cq_feedforward_oracle_equals_the_reference_run(share):
    """CMRL.all_modules_feedforward_lpc (cmrl.py:770-830: 256-bin LSF quantiser, lsf2poly and residual through tf.py_func, codec 0 on
    res_x * res_scalar, codec i on res_scalar * (res_x - sum of the earlier outputs), every output divided by res_scalar) followed by
    _feedforward_lpc's sum and synthesis (:836-839), and CMRL.all_modules_feedforward (:513-543: codec 0 neither scaled nor divided) --
    graphs built by the reference's own code, two codecs, res_scalar = 2."""
    g = dict(np.load(FIX_NN))
    tag = 'cq_soft' if share else 'cq_hard'
    ocs = _cq_oracle_codecs()
    bins = np.load(os.path.join(GOLD, 'lsf_bins_f64.npy')).astype(np.float32)
    x = torch.from_numpy(g['cq_x'])[:, :, None]
    lsf = torch.from_numpy(g['cq_lsf'])[:, :, None]
    with torch.no_grad():
        o = ref_codec.cq_feedforward(ocs, -300.0, bins, x, lsf, share, 1.0, res_scalar=2.0)
    idx = ref_nn.quantizer_indices(lsf, np.float32(-300.0), bins).numpy().astype(np.int64).reshape(g[tag + '_lsf_idx'].shape)
    assert np.array_equal(idx, g[tag + '_lsf_idx'])
    assert _rel(o['poly'], g[tag + '_poly']) < 2e-6
    assert _rel(o['res_x'][:, :, 0].numpy(), g[tag + '_res_x']) < 2e-6
    if share:       # no discontinuity on the soft path: the whole pass can be compared end to end
        assert _rel(torch.stack(o['outs']).numpy(), g[tag + '_outs']) < 2e-3      # alpha = -300 soft quantiser amplifies float32 rounding
        assert _rel(o['decoded'].numpy(), g[tag + '_decoded']) < 2e-3
        assert _rel(o['synthesized'], g[tag + '_synth']) < 2e-3
    else:           # hard path: codec 0 sees identical input; a code flipped at a mid-point changes its output locally
        d0 = np.abs(o['outs'][0].numpy() - g[tag + '_outs'][0]).max(1) / np.abs(g[tag + '_outs'][0]).max()
        assert np.median(d0) < 1e-5
    # the plain cascade: codec 0 unscaled and undivided
    with torch.no_grad():
        dec, outs, _ = ref_codec.cascade_forward(ocs, x, share, 1.0, res_scalar=2.0, lpc_variant=False)
    if share:
        assert _rel(torch.stack(outs).numpy(), g[tag + '_plain_outs']) < 2e-3
    else:
        d0 = np.abs(outs[0].numpy() - g[tag + '_plain_outs'][0]).max(1) / np.abs(g[tag + '_plain_outs'][0]).max()
        assert np.median(d0) < 1e-5
