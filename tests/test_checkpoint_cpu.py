"""TF-variable importer and schedule shell (host logic only, no GPU): names follow TF's creation-order convention, the
flat image round-trips, shape mismatches fail loudly, the entropy controller follows nscm.py:630-639."""
import numpy as np
import pytest

from nsc_b200 import checkpoint as ck, codec
from oracle import ref_codec


@pytest.mark.parametrize('rt,st', [('bottleneck', (2,)), ('gln', (2,)), ('bottleneck', (2, 2))])
def test_import_matches_oracle_parameter_stream(rt, st):
    cfg = codec.CodecConfig(resnet_type=rt, the_strides=st)
    oc = ref_codec.OracleCodec(ref_codec.OracleCodecCfg(resnet_type=rt, strides=st), seed=11)
    names = ck.tf_variable_names(cfg, 'scope_2')
    assert names[0] == ('scope_2/conv1d/kernel', 'scope_2/conv1d/bias')
    assert names[1][0] == 'scope_2/conv1d_1/kernel'
    variables = {}
    for ns, arrs in zip(names, oc.conv_params):
        for n, a in zip(ns, arrs):
            variables[n + ':0'] = a                       # TF's tensor names carry ':0'
    variables['scope_2/alpha'] = np.float32(oc.alpha)
    variables['scope_2/bins'] = np.asarray(oc.bins, dtype=np.float32)
    variables['scope_2/conv1d/kernel/Adam'] = np.zeros(3)  # optimiser slots are ignored
    flat = ck.params_from_tf_variables(cfg, 'scope_2', variables)
    want = codec.pack_params_numpy(cfg, oc.conv_params, oc.alpha, oc.bins)
    assert np.array_equal(flat, want)
    back = ck.tf_variables_from_params(cfg, 'scope_2', flat)
    for ns, arrs in zip(names, oc.conv_params):
        for n, a in zip(ns, arrs):
            assert np.array_equal(back[n], a)
    assert back['scope_2/alpha'] == np.float32(oc.alpha)


def test_import_errors():
    cfg = codec.CodecConfig(resnet_type='bottleneck')
    with pytest.raises(KeyError):
        ck.params_from_tf_variables(cfg, 'scope_1', {})
    oc = ref_codec.OracleCodec(ref_codec.OracleCodecCfg(), seed=1)
    variables = {n: a for ns, arrs in zip(ck.tf_variable_names(cfg, 'scope_1'), oc.conv_params) for n, a in zip(ns, arrs)}
    variables['scope_1/alpha'] = -300.0
    variables['scope_1/bins'] = np.linspace(-1, 1, 64)     # wrong codebook size
    with pytest.raises(ValueError):
        ck.params_from_tf_variables(cfg, 'scope_1', variables)
    variables['scope_1/bins'] = np.linspace(-1, 1, 32)
    variables['scope_1/conv1d_3/kernel'] = np.zeros((9, 20, 99), dtype=np.float32)
    with pytest.raises(ValueError):
        ck.params_from_tf_variables(cfg, 'scope_1', variables)


def test_lsf_codebook_and_schedule():
    v = {'lpc_quan/alpha:0': np.float32(-250.0), 'lpc_quan/bins:0': np.array([0.3, 0.1, 0.2], dtype=np.float32)}
    p = ck.lsf_params_from_tf_variables(v)
    assert p.dtype == np.float32 and np.allclose(p, [-250.0, 0.3, 0.1, 0.2])
    assert np.allclose(ck.sorted_lsf_bins(p), [-250.0, 0.1, 0.2, 0.3])
    c = ck.EntropyController(target_entropy=2.0, tau=0.1)
    assert abs(c.update(2.2) - 0.115) < 1e-12            # above target + 0.05
    assert abs(c.update(2.03) - 0.115) < 1e-12           # inside the dead band
    assert abs(c.update(1.9) - 0.07) < 1e-12             # below target: three steps down
    assert abs(c.update(5.0, is_quan_on=0.0) - 0.07) < 1e-12
    assert abs(c.update_finetune(2.5) - 0.085) < 1e-12
    s = ck.schedule(3, pretrain_step=5, epochs=100)
    assert s == {'loss': 'loss_no_quan', 'is_quan_on': 0.0, 'optimizer': 'no_quan', 'update_lpc_residual': False, 'skip_training': False}
    assert ck.schedule(30, 5, 100)['update_lpc_residual'] and not ck.schedule(31, 5, 100)['update_lpc_residual']
    # nscm.py:578-584: `is_cq and (i % 30 == 0 and i != 0 or i == epoch - 3)` -- also DURING pre-training, and that epoch skips training
    pre = ck.schedule(30, pretrain_step=50, epochs=100)
    assert pre['loss'] == 'loss_no_quan' and pre['update_lpc_residual'] and pre['skip_training']
    assert ck.schedule(97, 5, epochs=100)['update_lpc_residual'] and not ck.schedule(96, 5, epochs=100)['update_lpc_residual']
    assert not ck.schedule(30, 5, 100, is_cq=False)['update_lpc_residual'] and not ck.schedule(0, 5, 100)['update_lpc_residual']


def test_separable_names_are_per_graph_in_multi_codec_graphs():
    """Keras layer names are uniquified per graph: the second 'gln' codec's up-conv is scope_2/separable_conv1d_1 [LIB, unverified].
    The importer finds it by the predicted name (sep_start) and, failing that, by ascending suffix inside the scope."""
    cfg = codec.CodecConfig(resnet_type='gln', the_strides=(2,))
    assert ck.separable_count(cfg) == 1
    assert ck.separable_count(codec.CodecConfig(resnet_type='gln', the_strides=(2, 2))) == 2
    ocs = [ref_codec.OracleCodec(ref_codec.OracleCodecCfg(resnet_type='gln', strides=(2,)), seed=s) for s in (1, 2)]
    variables, start = {}, 0
    for i, oc in enumerate(ocs):
        scope = f'scope_{i + 1}'
        names = ck.tf_variable_names(cfg, scope, sep_start=start)
        for ns, arrs in zip(names, oc.conv_params):
            for n, a in zip(ns, arrs):
                variables[n] = a
        variables[f'{scope}/alpha'] = np.float32(oc.alpha)
        variables[f'{scope}/bins'] = np.asarray(oc.bins, dtype=np.float32)
        start += ck.separable_count(cfg)
    assert 'scope_1/separable_conv1d/depthwise_kernel' in variables and 'scope_2/separable_conv1d_1/depthwise_kernel' in variables
    for i, oc in enumerate(ocs):
        want = codec.pack_params_numpy(cfg, oc.conv_params, oc.alpha, oc.bins)
        assert np.array_equal(ck.params_from_tf_variables(cfg, f'scope_{i + 1}', variables, sep_start=i), want)
        # without the hint (sep_start = 0) the per-scope pattern search still finds the layer
        assert np.array_equal(ck.params_from_tf_variables(cfg, f'scope_{i + 1}', variables), want)
