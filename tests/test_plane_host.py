"""Host-side logic of the plane engine, no GPU needed: which layer shapes the tensor engine plans (and which it refuses, loudly),
workspace sizes of the codec entry points on the plane path, bit-packing sizes, frame counts."""
import numpy as np
import pytest

from nsc_b200 import _lib, codec


def _ws(B, Lin, Cin, Cout, k, dil=1, stride=1, res_mode=0, shuffle=1, precision=1):
    return _lib.load().nsc_conv1d_tc_workspace_bytes(B, Lin, Cin, Cout, k, dil, stride, res_mode, shuffle, precision)


def test_codec_layer_shapes_are_planned():
    # every conv shape of the '9 9 100 20 1 2' / stride-2 codec, both precisions
    shapes = [
        (512, 1, 100, 55, 1, 1, 0, 1), (512, 100, 20, 9, 1, 1, 0, 1), (512, 20, 20, 9, 2, 1, 0, 1), (512, 20, 100, 9, 1, 1, 1, 1),
        (512, 100, 100, 9, 1, 2, 0, 1), (256, 100, 1, 55, 1, 1, 0, 1), (256, 1, 20, 9, 1, 1, 0, 1), (256, 20, 100, 9, 1, 1, 2, 1),
        (256, 100, 100, 9, 1, 1, 0, 2), (512, 50, 20, 9, 1, 1, 0, 1), (512, 20, 50, 9, 1, 1, 1, 1), (512, 50, 1, 55, 1, 1, 0, 1),
    ]
    for prec in (1, 2):
        for (L, cin, cout, k, dil, stride, res, sh) in shapes:
            n = _ws(64, L, cin, cout, k, dil, stride, res, sh, prec)
            assert n > 0, ((L, cin, cout, k, dil, stride, res, sh, prec), _lib.last_error())
    # hi/lo planes need more workspace than fp16 planes
    assert _ws(64, 512, 100, 20, 9, precision=1) > _ws(64, 512, 100, 20, 9, precision=2)


@pytest.mark.parametrize('bad', [
    dict(Lin=500, Cin=100, Cout=20, k=9),                 # length not a multiple of 128
    dict(Lin=512, Cin=100, Cout=100, k=15, dil=2),        # 14-row halo does not fit the 8 zero rows of an image
    dict(Lin=512, Cin=100, Cout=100, k=9, stride=3),      # stride
    dict(Lin=512, Cin=200, Cout=100, k=9),                # more than 128 input channels
    dict(Lin=512, Cin=100, Cout=100, k=9, precision=0),   # fp32 is not a tensor-engine mode
])
def test_unsupported_shapes_are_refused(bad):
    assert _ws(8, **bad) < 0
    assert _lib.last_error() != ''


PLAN_KEYS = ('kind', 'staged', 'pair', 'mt', 'n_iss', 'resident', 'wslots', 'stages', 'smem', 'tmem_cols', 'grid', 'units')


def _plan(B, Lin, Cin, Cout, k, dil=1, stride=1, res_mode=0, shuffle=1, precision=1):
    import ctypes as C
    out = (C.c_int64 * 12)()
    rc = _lib.load().nsc_conv1d_tc_plan_info(B, Lin, Cin, Cout, k, dil, stride, res_mode, shuffle, precision, out)
    assert rc == 0, _lib.last_error()
    return dict(zip(PLAN_KEYS, list(out)))


def test_launch_plans_of_the_codec_layers():
    """What the plane engine would launch for every layer shape of the codec (148 SMs assumed without a device): kernel family,
    shared memory within one SM, TMEM within 512 columns, the rings the DESIGN.md kernel table describes."""
    smem_max = 227 * 1024
    t1 = _plan(2072, 512, 100, 20, 9)
    assert t1['kind'] == 0 and t1['stages'] == 5 and t1['wslots'] == 4 and t1['tmem_cols'] == 512      # 4 weight slabs resident, 5 input stages
    t2 = _plan(2072, 512, 20, 20, 9, dil=2)
    assert t2['kind'] == 0 and t2['staged'] == 3 and t2['stages'] == 7 and t2['wslots'] == 3 and t2['tmem_cols'] == 256   # grouped form: 3 tap groups, 5 tap slots
    head = _plan(2072, 256, 100, 1, 55)
    assert head['kind'] == 0 and head['tmem_cols'] == 128
    x3 = _plan(2072, 512, 20, 100, 9, res_mode=1)
    assert x3['kind'] == 1 and x3['staged'] == 1 and x3['pair'] == 1 and x3['mt'] == 2 and x3['resident'] == 0 and x3['wslots'] >= 4
    x3_odd = _plan(2071, 256, 20, 100, 9, res_mode=1)        # odd number of work units: one-CTA kernel, full-size weight slots
    assert x3_odd['pair'] == 0 and x3_odd['staged'] == 1 and x3_odd['smem'] < x3['smem']
    stem = _plan(2072, 512, 1, 100, 55)
    assert stem['kind'] == 2 and stem['staged'] == 1
    down = _plan(2072, 512, 100, 100, 9, stride=2)
    assert down['kind'] == 1 and down['mt'] == 1 and down['pair'] == 0 and down['resident'] == 0 and down['units'] == 2 * 2072
    up = _plan(2072, 256, 100, 100, 9, shuffle=2)
    assert up['kind'] == 1 and up['mt'] == 2 and up['n_iss'] == 2 and up['pair'] == 1 and up['grid'] == 148 and up['units'] == 2072
    for p in (t1, t2, head, x3, stem, down, up):
        assert 0 < p['smem'] <= smem_max and p['tmem_cols'] <= 512, p


def test_cta_pairs_need_an_even_number_of_work_units():
    """The up-sampling conv runs as CTA pairs only when every CTA of a pair gets the same number of tiles; an odd batch falls back
    to the one-CTA kernel (same weights image, same result).  A pair's ring holds half-size units, so it has more slots."""
    even = _plan(302, 256, 100, 100, 9, shuffle=2)
    odd = _plan(301, 256, 100, 100, 9, shuffle=2)
    assert even['pair'] == 1 and even['grid'] % 2 == 0 and even['grid'] == 148
    assert odd['pair'] == 0 and odd['grid'] == 148
    assert even['wslots'] > odd['wslots'] and even['smem'] <= 227 * 1024
    tiny = _plan(2, 256, 100, 100, 9, shuffle=2)
    assert tiny['pair'] == 1 and tiny['grid'] == 2
    one = _plan(1, 256, 100, 100, 9, shuffle=2)
    assert one['pair'] == 0 and one['grid'] == 1


def test_codec_workspace_sizes_on_the_plane_path():
    lib = _lib.load()
    cfg = codec.CodecConfig(resnet_type='bottleneck', precision='tc_f16x3').to_struct()
    cfg32 = codec.CodecConfig(resnet_type='bottleneck', precision='fp32').to_struct()
    import ctypes as C
    one = lib.nsc_codec_workspace_bytes(C.byref(cfg), 1)
    big = lib.nsc_codec_workspace_bytes(C.byref(cfg), 100000)
    npass = lib.nsc_pass_frames(C.byref(cfg), 1)
    assert npass == 28 * 148 and lib.nsc_pass_frames(C.byref(cfg32), 1) > 0 and lib.nsc_pass_frames(None, 1) == -1
    chunk = lib.nsc_codec_workspace_bytes(C.byref(cfg), npass)
    assert 0 < one < lib.nsc_codec_workspace_bytes(C.byref(cfg), npass // 2) < chunk == big      # capped at one pass of 28 x 148 frames
    per_frame = (chunk - one) / (npass - 1)
    assert 1.5e6 < per_frame < 1.8e6                        # 11 plane buffers + 3 folded narrow images: 1.76 MB per frame
    assert lib.nsc_codec_workspace_bytes(C.byref(cfg32), 2048) < chunk   # fp32 NCL buffers are smaller
    # the reference's shipped 'gln' blocks run on the plane path too: three more buffers (two de-interleaved narrow twins for the
    # dilation-2 gate convs, a third half-length wide image for the depthwise half of the separable up-conv)
    gln = codec.CodecConfig(resnet_type='gln', precision='tc_f16x3').to_struct()
    gchunk = lib.nsc_codec_workspace_bytes(C.byref(gln), npass)
    assert chunk < gchunk < 1.2 * chunk and gchunk == lib.nsc_codec_workspace_bytes(C.byref(gln), 100000)
    # two stride-2 stages (the_strides = '4', cmrl.py:804) with bottleneck blocks: on the plane path as well (three resolution levels)
    s4 = codec.CodecConfig(resnet_type='bottleneck', the_strides=(2, 2), precision='tc_f16x3').to_struct()
    assert lib.nsc_codec_on_plane_engine(C.byref(s4)) == 1 and lib.nsc_codec_workspace_bytes(C.byref(s4), npass) > chunk
    g4 = codec.CodecConfig(resnet_type='gln', the_strides=(2, 2), precision='tc_f16x3').to_struct()
    assert lib.nsc_codec_on_plane_engine(C.byref(g4)) == 1      # (its dilation-2 gates at 128 positions stage 16 halo rows)
    # configurations outside the plane path keep the layer-by-layer workspace: the one-plane mode with two stages (wide / 4 channels
    # would be a packed image), other narrow widths
    for cfg_out in (codec.CodecConfig(bottleneck_kernel_and_dilation=(9, 9, 100, 24, 1, 2), precision='tc_f16x3'),
                    codec.CodecConfig(resnet_type='bottleneck', the_strides=(2, 2), precision='tc_f16')):
        st = cfg_out.to_struct()
        assert lib.nsc_codec_on_plane_engine(C.byref(st)) == 0
        assert lib.nsc_codec_workspace_bytes(C.byref(st), 2048) < chunk
    assert lib.nsc_codec_on_plane_engine(C.byref(gln)) == 1 and lib.nsc_codec_on_plane_engine(C.byref(cfg)) == 1
    assert lib.nsc_codec_on_plane_engine(C.byref(cfg32)) == 0


def test_framing_and_packing_sizes():
    lib = _lib.load()
    for T in (0, 512, 513, 992, 993, 48000):
        assert lib.nsc_segment_count(T) == len(range(0, T - 512, 480))
    assert lib.nsc_lpc_window_count(99) == 97 and lib.nsc_lpc_window_count(2) == 0
    assert lib.nsc_packed_row_bytes(256, 5) == 160 and lib.nsc_packed_row_bytes(16, 8) == 16 and lib.nsc_packed_row_bytes(7, 3) == 3
    assert lib.nsc_iir_workspace_bytes(1000, 2) >= 2 * 1000 * 8


def test_narrow_conv_plan_follows_the_frame_length():
    """The block's k9 20 -> 20 conv as the codec program plans it (host logic): folded images (pairs of positions in the channel axis,
    three MMA-issuing threads with one accumulator each = 3 x 2 x 48 TMEM columns -> 512, resident weights, a staging window for the
    unfolding epilogue inside the shared-memory budget) from 256 positions up -- dilation 2 per position parity at 512 positions
    (twice the frames, paired tiles), block-diagonal k9 at 256 -- and the taps-in-N kernel on shorter frames."""
    import ctypes as C
    lib = _lib.load()
    keys = ('kind', 'staged', 'pair', 'mt', 'n_iss', 'resident', 'wslots', 'stages', 'smem', 'tmem_cols', 'grid', 'units')

    def plan(B, L, d):
        out = (C.c_int64 * 12)()
        assert lib.nsc_narrow_conv_plan_info(B, L, d, out) == 0, _lib.last_error()
        return dict(zip(keys, list(out)))
    for L, d, form, wslots, units in ((512, 1, 5, 10, 2 * 2072), (256, 1, 5, 10, 2072), (512, 2, 5, 10, 2 * 2072), (256, 2, 6, 18, 2072)):
        p = plan(2072, L, d)
        assert (p['kind'], p['staged'], p['pair'], p['mt'], p['n_iss'], p['resident'], p['wslots']) == (1, form, 0, 1, 3, 1, wslots), (L, d, p)
        assert p['tmem_cols'] == 512 and p['smem'] <= 227 * 1024 and p['stages'] >= 4 and p['units'] == units
        assert p['grid'] == 148
    assert plan(2072, 512, 2)['grid'] == 148 and plan(7, 512, 2)['grid'] == 7      # dilation 2 by parity: a CTA's unit of work is a frame
    short = plan(2072, 128, 1)
    assert short['kind'] == 0 and short['staged'] == 3                          # taps-in-N, three tap groups
    assert lib.nsc_narrow_conv_plan_info(1, 100, 1, (C.c_int64 * 12)()) != 0


def test_plane_kernels_do_not_spill():
    """ptxas -v logs of the in-tree build (nsc_b200/csrc/*.ptxas.log): no kernel of the plane engine may spill registers.  (The
    unfolding epilogue of the folded narrow conv once shared a kernel with the other tap-shift layers: every instantiation spilled and
    the HBM-bound k1 layers of 'gln' lost 30 % -- it now has its own instantiation.)"""
    import os
    import re
    log = os.path.join(os.path.dirname(os.path.abspath(_lib.__file__)), 'csrc', 'plane_conv.ptxas.log')
    if not os.path.exists(log):
        pytest.skip('no ptxas log (library built elsewhere)')
    t = open(log).read()
    found = re.findall(r"Compiling entry function '([^']+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", t)
    assert len(found) >= 10
    for name, stack, st, ld in found:
        if 'plane_' in name and 'plane_block_kernel' not in name:      # (the opt-in fused block kernel runs 18 warps at 96 registers: a few bytes)
            assert (int(st), int(ld)) == (0, 0), name
        if 'plane_block_kernel' in name:
            assert int(st) <= 128 and int(ld) <= 128, name
