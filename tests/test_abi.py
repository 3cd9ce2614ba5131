"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/nsc_b200.h declares,
the Python binding table covers them, and the host-side layout logic (layer tables, error codes) matches
the oracle.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nsc_b200 import _lib, codec
from oracle import ref_codec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'nsc_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nsc_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libnsc_b200.so does not export {s}"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes table and header disagree"
    assert lib.nsc_version() == 100


def test_no_torch_types_in_header():
    src = open(os.path.join(ROOT, 'include', 'nsc_b200.h')).read()
    assert 'torch' not in src.lower() and 'at::' not in src and '#include <cuda' not in src


@pytest.mark.parametrize('rt,st', [('bottleneck', (2,)), ('gln', (2,)), ('bottleneck', (2, 2)), ('gln', (2, 2))])
def test_layer_table_matches_oracle_creation_order(rt, st):
    cfg = codec.CodecConfig(resnet_type=rt, the_strides=st)
    oc = ref_codec.OracleCodec(ref_codec.OracleCodecCfg(resnet_type=rt, strides=st))
    tab = codec.layer_table(cfg)
    assert len(tab) == len(oc.conv_params)
    off = 0
    for L, t in zip(tab, oc.conv_params):
        assert L.offset == off
        assert L.separable == (len(t) == 3)
        assert t[0].shape == ((L.k, L.cin, 1) if L.separable else (L.k, L.cin, L.cout))
        off += sum(int(np.prod(p.shape)) for p in t)
    assert codec.param_count(cfg) == off + 1 + cfg.num_bins


def test_pack_params_roundtrip():
    cfg = codec.CodecConfig(resnet_type='bottleneck')
    oc = ref_codec.OracleCodec(ref_codec.OracleCodecCfg(), seed=3)
    flat = codec.pack_params_numpy(cfg, oc.conv_params, oc.alpha, oc.bins)
    tab = codec.layer_table(cfg)
    L = tab[7]
    w = flat[L.offset:L.offset + L.k * L.cin * L.cout].reshape(L.k, L.cin, L.cout)
    assert np.array_equal(w, oc.conv_params[7][0])
    assert flat[-33] == np.float32(-300.0) and np.array_equal(flat[-32:], oc.bins)


def test_from_args_matches_readme_flags():
    cfg = codec.CodecConfig.from_args('9 9 100 20 1 2', '2', 32)
    assert cfg.the_strides == (2,) and cfg.code_length == 256
    assert codec.CodecConfig.from_args('9 9 100 20 1 2', '4', 32).code_length == 128


def test_invalid_config_is_reported_not_crashed():
    lib = _lib.load()
    st = codec.CodecConfig(resnet_type='bottleneck').to_struct()
    st.num_bins = 1000
    assert lib.nsc_codec_param_count(C.byref(st)) == -1
    assert 'num_bins' in _lib.last_error()
    st = codec.CodecConfig(resnet_type='bottleneck').to_struct()
    st.wide = 101          # decoder channels not divisible by the stride (nscm.py:187 assert)
    assert lib.nsc_codec_workspace_bytes(C.byref(st), 4) == -1
    with pytest.raises(ValueError):
        codec.CodecConfig(resnet_type='resnet').to_struct()


def test_cpu_tensors_are_rejected_loudly():
    import torch
    from nsc_b200 import nn_core_operator as nn
    with pytest.raises(ValueError, match='CUDA'):
        nn.scalar_softmax_quantization(torch.zeros(1, 4, 1), -300.0, torch.linspace(-1, 1, 32), 1.0, False, 4, 32)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'nsc_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', txt, re.M), f
