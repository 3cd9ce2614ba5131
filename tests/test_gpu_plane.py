"""Plane engine (tcgen05 conv layers on fp16 plane images, nsc_b200/csrc/plane_conv.cu) against the CPU oracle, layer by
layer and through the C ABI (nsc_conv1d_tc), on every layer shape the codec uses.

precision 1 (fp16 hi/lo split, 3 MMAs) must meet the fp32 bar; precision 2 (plain fp16) is checked against a looser,
stated bound."""
import numpy as np
import pytest
import torch

from oracle import ref_nn
from util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'
ACTS = {None: 0, 'tanh': 1, 'lrelu': 2}


def _oracle(x, w, b, dil, stride, act, res, res_mode, post_act, shuffle):
    y = ref_nn.conv1d_explicit(torch.from_numpy(x), w, b, dil, stride, None)
    y = y.numpy().astype(np.float64)
    f = {None: lambda v: v, 'tanh': np.tanh, 'lrelu': lambda v: np.where(v > 0, v, 0.2 * v)}
    y = f[act](y)
    if res_mode == 1:
        y = y + res
    elif res_mode == 2:
        y = y + res[:, :, None]
    y = f[post_act](y)
    if shuffle > 1:
        B, L, C = y.shape
        y = y.reshape(B, L, C // shuffle, shuffle).transpose(0, 1, 3, 2).reshape(B, L * shuffle, C // shuffle)
    return y


def _run(B, Lin, Cin, Cout, k, dil=1, stride=1, act=None, res_mode=0, post_act=None, shuffle=1, precision=1, seed=0):
    from nsc_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(seed)
    x = rng.randn(B, Lin, Cin).astype(np.float32)
    lim = np.sqrt(6.0 / (k * Cin + k * Cout))
    w = rng.uniform(-lim, lim, (k, Cin, Cout)).astype(np.float32)
    b = (0.1 * rng.randn(Cout)).astype(np.float32)
    Lout = (Lin + stride - 1) // stride
    res = None
    if res_mode == 1:
        res = rng.randn(B, Lout, Cout).astype(np.float32)
    elif res_mode == 2:
        res = rng.randn(B, Lout).astype(np.float32)
    want = _oracle(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), dil, stride, act,
                   None if res is None else res.astype(np.float64), res_mode, post_act, shuffle)
    xt, wt, bt = (torch.from_numpy(a).to(DEV) for a in (x, w, b))
    rt = None if res is None else torch.from_numpy(res).to(DEV)
    y = torch.full((B, Lout * shuffle, Cout // shuffle), float('nan'), device=DEV)
    ws_bytes = lib.nsc_conv1d_tc_workspace_bytes(B, Lin, Cin, Cout, k, dil, stride, res_mode, shuffle, precision)
    assert ws_bytes > 0, _lib.last_error()
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
    rc = lib.nsc_conv1d_tc(_lib.ptr(xt), _lib.ptr(wt), _lib.ptr(bt), _lib.ptr(rt), _lib.ptr(y), B, Lin, Cin, Cout, k, dil,
                           stride, ACTS[act], res_mode, ACTS[post_act], shuffle, precision, _lib.ptr(ws), ws_bytes,
                           _lib.stream_ptr())
    _lib.check(rc, 'nsc_conv1d_tc')
    torch.cuda.synchronize()
    got = y.cpu().numpy()
    assert np.isfinite(got).all()
    return rel_err(got, want.reshape(got.shape))


# every conv shape of the '9 9 100 20 1 2' / stride-2 bottleneck codec
T_LAYERS = [
    dict(Lin=512, Cin=100, Cout=20, k=9, act='lrelu'),             # block conv 1 @512
    dict(Lin=256, Cin=100, Cout=20, k=9, act='lrelu'),             # block conv 1 @256
    dict(Lin=512, Cin=50, Cout=20, k=9, act='lrelu'),              # decoder block conv 1 on 50 channels
    dict(Lin=512, Cin=20, Cout=20, k=9, dil=1, act='lrelu'),       # block conv 2, dilation 1
    dict(Lin=512, Cin=20, Cout=20, k=9, dil=2, act='lrelu'),       # block conv 2, dilation 2
    dict(Lin=256, Cin=20, Cout=20, k=9, dil=2, act='lrelu'),
    dict(Lin=256, Cin=100, Cout=1, k=55, act='tanh'),              # code head
    dict(Lin=512, Cin=50, Cout=1, k=55, act=None),                 # output head
]
X_LAYERS = [
    dict(Lin=512, Cin=20, Cout=100, k=9, res_mode=1, post_act='lrelu'),    # block conv 3 + residual + lrelu
    dict(Lin=256, Cin=20, Cout=100, k=9, res_mode=1, post_act=None),       # last block of a stack: flat
    dict(Lin=256, Cin=20, Cout=100, k=9, res_mode=2, post_act='lrelu'),    # decoder block 1: broadcast residual
    dict(Lin=512, Cin=20, Cout=50, k=9, res_mode=1, post_act='lrelu'),
    dict(Lin=512, Cin=100, Cout=100, k=9, stride=2, act='lrelu'),          # down-sampling conv
    dict(Lin=256, Cin=100, Cout=100, k=9, act='lrelu', shuffle=2),         # up-sampling conv + sub-pixel shuffle
    dict(Lin=512, Cin=1, Cout=100, k=55, act='lrelu'),                     # stem (Toeplitz)
    dict(Lin=256, Cin=1, Cout=20, k=9, act='lrelu'),                       # decoder block 1 conv 1 (Toeplitz)
]


@pytest.mark.parametrize('layer', T_LAYERS, ids=lambda d: 'k%d_d%d_%dto%d_L%d' % (d['k'], d.get('dil', 1), d['Cin'], d['Cout'], d['Lin']))
@pytest.mark.parametrize('precision', [1, 2])
def test_taps_in_n_layers(layer, precision):
    # 5 frames on 148 CTAs: one frame per CTA; 301 frames: CTAs walk several frames (spill slots recycle)
    for B in (5, 301):
        e = _run(B=B, precision=precision, seed=B, **layer)
        assert e < (2e-5 if precision == 1 else 4e-3), (B, e)


@pytest.mark.parametrize('layer', X_LAYERS, ids=lambda d: 'k%d_s%d_%dto%d_L%d_r%d_sh%d' % (d['k'], d.get('stride', 1), d['Cin'], d['Cout'], d['Lin'], d.get('res_mode', 0), d.get('shuffle', 1)))
@pytest.mark.parametrize('precision', [1, 2])
def test_tap_shift_layers(layer, precision):
    for B in (3, 301):
        e = _run(B=B, precision=precision, seed=B, **layer)
        assert e < (2e-5 if precision == 1 else 4e-3), (B, e)


def _child(env_extra, body):
    """Engine switches are read once per process: runs `body` (python source using test_gpu_plane as t) in a child process."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\nimport test_gpu_plane as t\n" % (root, os.path.join(root, 'tests'))) + body
    subprocess.run([sys.executable, '-c', code], check=True, env=dict(os.environ, **env_extra), timeout=600)


PAIR_LAYERS = [X_LAYERS[5], X_LAYERS[0], X_LAYERS[1], X_LAYERS[2], X_LAYERS[3]]


@pytest.mark.parametrize('layer', PAIR_LAYERS, ids=['up_shuffle', '20to100_L512', '20to100_L256_flat', '20to100_L256_bcast', '20to50_L512'])
@pytest.mark.parametrize('precision', [1, 2])
def test_cta_pair_layers(layer, precision):
    """An even number of work units runs the up-sampling 100 -> 100 conv and the staged narrow-input layers as CTA pairs
    (cta_group::2: M = 256 instructions issued by the leader, each CTA stages its own tile and half of every weight unit; the staged
    kernel uses the freed shared memory for a fourth residual / output unit).  2 frames = one or two pairs; 302 = every pair busy,
    ragged; 1184 frames = several units per CTA so ring slots, accumulator slots, staging units and the peer's landed-reports wrap."""
    for B in (2, 302, 1184):
        e = _run(B=B, precision=precision, seed=B, **layer)
        assert e < (2e-5 if precision == 1 else 4e-3), (B, e)


@pytest.mark.parametrize('pair', ['0', '2'])
def test_cta_pair_knob(pair):
    """NSC_PLANE_PAIR=0: the one-CTA kernel on even batches; =2: pairs wherever the shape allows (stride-2 conv, and 20 -> 20 on
    the tap-shift kernel via NSC_PLANE_NARROW=X, and the taps-in-N 100 -> 20 layer with half of every weight slab per CTA and eight
    input stages).  Results must not depend on the choice."""
    _child({'NSC_PLANE_PAIR': pair, 'NSC_PLANE_NARROW': 'X'},
           "for layer in (t.X_LAYERS[4], t.X_LAYERS[5], t.X_LAYERS[0], t.X_LAYERS[3], t.T_LAYERS[3], t.T_LAYERS[5], t.T_LAYERS[0], t.T_LAYERS[1]):\n"
           "    for prec in (1, 2):\n"
           "        for B in (2, 302, 1184):\n"
           "            e = t._run(B=B, precision=prec, seed=B, **layer)\n"
           "            assert e < (2e-5 if prec == 1 else 4e-3), (layer, prec, B, e)\n")


@pytest.mark.parametrize('groups', ['0', '5'])
def test_narrow_layers_with_other_tap_groupings(groups):
    """20 -> 20 (k9, dilation 1 / 2): the default is three tap groups by descriptor row shift with five tap slots across TMEM lanes;
    NSC_PLANE_TGROUPS=0 keeps all nine taps in N, =5 uses five groups with three slots.  Same results, frame borders included
    (5 frames: every tile touches a border; 301: CTAs walk several frames)."""
    _child({'NSC_PLANE_TGROUPS': groups},
           "for layer in (t.T_LAYERS[3], t.T_LAYERS[4], t.T_LAYERS[5]):\n"
           "    for prec in (1, 2):\n"
           "        for B in (5, 301):\n"
           "            e = t._run(B=B, precision=prec, seed=B, **layer)\n"
           "            assert e < (2e-5 if prec == 1 else 4e-3), (layer, prec, B, e)\n")


def test_unsupported_shape_fails_loudly():
    from nsc_b200 import _lib
    lib = _lib.load()
    assert lib.nsc_conv1d_tc_workspace_bytes(4, 500, 100, 20, 9, 1, 1, 0, 1, 1) < 0     # 500 is not a multiple of 128
    assert 'tensor engine' in _lib.last_error()


# ------------------------------------------------------------------------------------------------ full-size properties
def _cq2(seed0=5, seed1=6, precision='tc_f16x3'):
    from nsc_b200 import codec
    cfg = codec.CodecConfig(resnet_type='bottleneck', precision=precision)
    return codec.CMRL([codec.NeuralCodec(cfg, device=DEV, seed=seed0), codec.NeuralCodec(cfg, device=DEV, seed=seed1)], res_scalar=1.0)


def test_full_size_batch_independence_across_chunk_boundaries():
    """BASELINE-size batch (4,500 frames: a full 4,144-frame pass and a ragged one): every frame's codes and waveform are
    bit-identical to coding it in a 7-frame batch -- no dependence on the chunk it fell in, the CTA that took it, or its
    neighbours (frames are independent units)."""
    from util import ar_frames
    from nsc_b200 import lpc_utilities as lu
    cm = _cq2()
    win = torch.from_numpy(ar_frames(4500, 1024, seed=77)).to(DEV)
    x = win[:, 256:768].contiguous()
    lsf = lu.lpc_analysis_windows(win, 16, dtype=torch.float32)
    big = cm.feedforward_lpc(x, lsf, False, 1.0)
    for lo in (0, 2068, 2072, 4140, 4144, 4493):      # around the pass boundary (4,144) and the ragged tail
        small = cm.feedforward_lpc(x[lo:lo + 7].contiguous(), lsf[lo:lo + 7].contiguous(), False, 1.0)
        for k in range(2):
            assert torch.equal(small['idx'][k], big['idx'][k][lo:lo + 7])
        assert torch.equal(small['lsf_idx'], big['lsf_idx'][lo:lo + 7])
        assert torch.equal(small['synthesized'], big['synthesized'][lo:lo + 7])
    assert torch.isfinite(big['synthesized']).all()


def test_full_size_cascade_algebra_and_hard_round_trip():
    """decoded = sum of the codecs' outputs; decoding the hard indices reproduces the forward pass output bit for bit."""
    from util import ar_frames
    cm = _cq2(7, 8)
    x = torch.from_numpy(ar_frames(2500, 512, seed=78, std=0.3)).to(DEV)
    r = cm.all_modules_feedforward(x, False, 1.0, want_outs=True)
    total = r['outs'][0] + r['outs'][1]
    assert rel_err(r['decoded'].cpu().numpy(), total.cpu().numpy()) < 1e-6
    # codec 0 sees x itself: its decoder applied to the transmitted indices gives outs[0] exactly
    dec0 = cm.codecs[0].decode_indices(r['idx'][0])
    assert torch.equal(dec0, r['outs'][0])
    # codec 1 codes the residual x - outs[0]
    enc1 = cm.codecs[1].encode((x - r['outs'][0]).contiguous(), False, 1.0)
    assert torch.equal(enc1['idx'], r['idx'][1])


@pytest.mark.parametrize('resnet_type', ['bottleneck', 'gln'])
def test_plane_path_matches_layer_by_layer_engine_at_full_size(monkeypatch, resnet_type):
    """Plane engine vs the first tensor engine (same hi/lo arithmetic, fp32 activations between layers) on 1,000 frames --
    both block types: the_bottleneck and the reference's shipped default gated_bottleneck (constants.py:14)."""
    import subprocess, sys, os, json
    from util import ar_frames
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, json, numpy as np, torch\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from util import ar_frames\n"
        "from nsc_b200 import codec\n"
        "cfg = codec.CodecConfig(resnet_type=%r); gc = codec.NeuralCodec(cfg, device='cuda', seed=9)\n"
        "x = torch.from_numpy(ar_frames(1000, 512, seed=79, std=0.3)).cuda()\n"
        "r = gc.computational_graph_end2end_quan_on(x, True, 1.0)\n"
        "np.save(sys.argv[1], np.concatenate([r['floating_code'].cpu().numpy().ravel(), r['out'].cpu().numpy().ravel()]))\n"
    ) % (root, os.path.join(root, 'tests'), resnet_type)
    outs = []
    for tag, env in (('plane', {}), ('layered', {'NSC_PLANE': '0'})):
        path = '/tmp/nsc_plane_vs_layered_%s.npy' % tag
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, '-c', code, path], check=True, env=e, timeout=300)
        outs.append(np.load(path))
    assert rel_err(outs[0], outs[1]) < 1e-4


def test_folded_cascade_and_quantiser_are_bit_identical_to_the_stand_alone_kernels():
    """The plane path folds the cascade's input arithmetic into the stem (in = res_scalar * (x - decoded)), the hard quantiser into the
    code head's epilogue and the accumulation decoded += out / res_scalar into the output head's (cmrl.py:522-531, :810-830;
    nn_core_operator.py:140-164).  NSC_PLANE_FOLD=0 runs the stand-alone kernels instead: every output must agree to the bit, and the
    folded program launches 3 kernels fewer per codec and pass (2 for the first codec of a cascade with res_scalar = 1)."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from util import ar_frames\n"
        "from nsc_b200 import codec, _lib\n"
        "from oracle import ref_lpc\n"
        "cfg = codec.CodecConfig(resnet_type='bottleneck')\n"
        "cm = codec.CMRL([codec.NeuralCodec(cfg, device='cuda', seed=5 + i) for i in range(3)], res_scalar=2.0)\n"
        "x = torch.from_numpy(ar_frames(300, 512, seed=4, std=0.3)).cuda()\n"
        "lsf = torch.from_numpy(ref_lpc.lpc_analysis_windows(ar_frames(300, 1024, seed=5), 16).astype(np.float32)).cuda()\n"
        "l0 = _lib.load().nsc_launch_count()\n"
        "r = cm.feedforward_lpc(x, lsf, False, 1.0)\n"
        "n = _lib.load().nsc_launch_count() - l0\n"
        "c = cm.all_modules_feedforward(x, False, 1.0, want_outs=True)\n"
        "e = cm.codecs[0].computational_graph_end2end_quan_on(x, False, 1.0)\n"
        "torch.cuda.synchronize()\n"
        "np.savez(sys.argv[1], launches=n, dec=r['decoded'].cpu().numpy(), syn=r['synthesized'].cpu().numpy(),\n"
        "         idx=np.stack([i.cpu().numpy() for i in r['idx']]), cdec=c['decoded'].cpu().numpy(),\n"
        "         couts=np.stack([o.cpu().numpy() for o in c['outs']]), eidx=e['idx'].cpu().numpy(), eout=e['out'].cpu().numpy(),\n"
        "         ecode=e['code'].cpu().numpy())\n"
    ) % (root, os.path.join(root, 'tests'))
    res = {}
    for tag, env in (('fold', {}), ('plain', {'NSC_PLANE_FOLD': '0'})):
        path = '/tmp/nsc_fold_%s.npz' % tag
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, '-c', code, path], check=True, env=e, timeout=300)
        res[tag] = dict(np.load(path))
    for k in ('dec', 'syn', 'idx', 'cdec', 'couts', 'eidx', 'eout', 'ecode'):
        assert np.array_equal(res['fold'][k], res['plain'][k]), k
    assert int(res['plain']['launches']) - int(res['fold']['launches']) == 3 * 3


def test_prepared_workspace_skips_the_per_weights_work_and_changes_no_bit():
    """nsc_prepare (include/nsc_b200.h): zero rows + packed weights once per weights, not once per call.  Same bits as the plain call,
    about half the launches at batch 128 (one pack launch per layer gone); a changed configuration, batch size or parameter pointer
    falls back to the full call; in-place weight changes need a new prepare (documented contract)."""
    from nsc_b200 import codec, _lib
    from oracle import ref_lpc
    from util import ar_frames
    lib = _lib.load()
    cfg = codec.CodecConfig(resnet_type='bottleneck')
    gc = codec.NeuralCodec(cfg, device=DEV, seed=21)
    x = torch.from_numpy(ar_frames(128, 512, seed=2, std=0.3)).to(DEV)
    l0 = lib.nsc_launch_count()
    a = gc.computational_graph_end2end_quan_on(x, False, 1.0)
    n_plain = lib.nsc_launch_count() - l0
    gc.prepare(128)
    l0 = lib.nsc_launch_count()
    b = gc.computational_graph_end2end_quan_on(x, False, 1.0)
    n_prep = lib.nsc_launch_count() - l0
    torch.cuda.synchronize()
    assert all(torch.equal(a[k], b[k]) for k in ('out', 'idx', 'code', 'floating_code'))
    assert n_prep <= n_plain - 25, (n_plain, n_prep)
    # encode / decode share the codec's workspace layout
    e = gc.encode(x)
    assert torch.equal(e['idx'], a['idx']) and torch.equal(gc.decode_indices(e['idx']), a['out'])
    # another batch size below one pass: full call again (and the registration is dropped)
    c = gc.computational_graph_end2end_quan_on(x[:64], False, 1.0)
    assert torch.equal(c['out'], a['out'][:64])
    assert lib.nsc_release(_lib.ptr(gc._ws)) == 0
    # weights changed in place: prepare again
    gc.prepare(128)
    gc.params.mul_(1.01)
    gc.prepare(128)
    d = gc.computational_graph_end2end_quan_on(x, False, 1.0)
    gc.release()
    f = gc.computational_graph_end2end_quan_on(x, False, 1.0)
    torch.cuda.synchronize()
    assert torch.equal(d['out'], f['out']) and not torch.equal(d['out'], a['out'])
    # collaborative quantisation entry point
    cm = codec.CMRL([codec.NeuralCodec(cfg, device=DEV, seed=5 + i) for i in range(2)], res_scalar=1.0)
    lsf = torch.from_numpy(ref_lpc.lpc_analysis_windows(ar_frames(128, 1024, seed=5), 16).astype(np.float32)).to(DEV)
    r0 = cm.feedforward_lpc(x, lsf, False, 1.0)
    cm.prepare(128)
    r1 = cm.feedforward_lpc(x, lsf, False, 1.0)
    r2 = cm.feedforward_lpc(x, lsf, False, 1.0)
    torch.cuda.synchronize()
    for k in ('decoded', 'synthesized', 'lsf_idx'):
        assert torch.equal(r0[k], r1[k]) and torch.equal(r0[k], r2[k])
    assert all(torch.equal(i0, i1) for i0, i1 in zip(r0['idx'], r1['idx']))
    assert lib.nsc_release(_lib.ptr(cm._ws)) == 1 and lib.nsc_release(_lib.ptr(cm._ws)) == 0


def test_cuda_graph_replay_of_a_fixed_batch_call_is_bit_identical():
    """codec.GraphedCall: the batch-128 codec call and the collaborative-quantisation call captured once (after nsc_prepare) and
    replayed -- same bits as the plain calls, for new inputs on every replay."""
    from nsc_b200 import codec
    from oracle import ref_lpc
    from util import ar_frames
    cfg = codec.CodecConfig(resnet_type='bottleneck')
    gc = codec.NeuralCodec(cfg, device=DEV, seed=31)
    g = gc.graphed_forward(128)
    for seed in (1, 2):
        x = torch.from_numpy(ar_frames(128, 512, seed=seed, std=0.3)).to(DEV)
        r = g.run(x)
        got = {k: r[k].clone() for k in ('out', 'idx', 'code')}
        gc.release()
        ref = gc.computational_graph_end2end_quan_on(x, False, 1.0)
        gc.prepare(128)
        torch.cuda.synchronize()
        assert all(torch.equal(got[k], ref[k]) for k in got)
    cm = codec.CMRL([codec.NeuralCodec(cfg, device=DEV, seed=5 + i) for i in range(2)], res_scalar=1.0)
    gq = cm.graphed_feedforward_lpc(64)
    x = torch.from_numpy(ar_frames(64, 512, seed=3, std=0.3)).to(DEV)
    lsf = torch.from_numpy(ref_lpc.lpc_analysis_windows(ar_frames(64, 1024, seed=4), 16).astype(np.float32)).to(DEV)
    r = gq.run(x, lsf)
    got = {k: r[k].clone() for k in ('decoded', 'synthesized', 'lsf_idx')}
    gidx = [t.clone() for t in r['idx']]
    cm.release()
    ref = cm.feedforward_lpc(x, lsf, False, 1.0)
    torch.cuda.synchronize()
    assert all(torch.equal(got[k], ref[k]) for k in got) and all(torch.equal(a, b) for a, b in zip(gidx, ref['idx']))


def test_staged_epilogue_ring_is_bit_identical_to_the_direct_epilogue_under_load():
    """compute-sanitizer's racecheck cannot model the mbarrier / async-proxy chain of the staged epilogue (epilogue writes a unit ->
    st_done -> bulk store -> wait_group.read -> st_empty -> the residual loader's bulk copy refills the unit) and reports a potential
    WAW hazard on the unit (profiles/r01_sanitizer_racecheck_plane.txt).  Evidence instead of argument: NSC_PLANE_NOSTAGE=1 runs the
    same MMAs and the same epilogue arithmetic with per-thread global loads / stores and NO shared-memory staging.  Over 3,000 frames
    x 2 codecs (~1.1 million unit hand-offs across all 148 persistent CTAs), three repetitions, every code and every output sample of
    the two variants must agree to the bit -- one unit refilled before it was stored, or stored before it was written, would show."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from util import ar_frames\n"
        "from nsc_b200 import codec\n"
        "from oracle import ref_lpc\n"
        "cfg = codec.CodecConfig(resnet_type='bottleneck')\n"
        "cm = codec.CMRL([codec.NeuralCodec(cfg, device='cuda', seed=5 + i) for i in range(2)], res_scalar=1.0)\n"
        "x = torch.from_numpy(ar_frames(3000, 512, seed=4, std=0.3)).cuda()\n"
        "lsf = torch.from_numpy(ref_lpc.lpc_analysis_windows(ar_frames(3000, 1024, seed=5), 16).astype(np.float32)).cuda()\n"
        "outs = [cm.feedforward_lpc(x, lsf, True, 1.0) for _ in range(3)]\n"
        "torch.cuda.synchronize()\n"
        "assert all(torch.equal(outs[0]['decoded'], o['decoded']) for o in outs[1:])\n"
        "h = cm.feedforward_lpc(x, lsf, False, 1.0)\n"
        "np.savez(sys.argv[1], dec=outs[0]['decoded'].cpu().numpy(), syn=outs[0]['synthesized'].cpu().numpy(),\n"
        "         idx=np.stack([i.cpu().numpy() for i in h['idx']]), hdec=h['decoded'].cpu().numpy())\n"
    ) % (root, os.path.join(root, 'tests'))
    res = {}
    for tag, env in (('staged', {}), ('direct', {'NSC_PLANE_NOSTAGE': '1'})):
        path = '/tmp/nsc_stage_%s.npz' % tag
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, '-c', code, path], check=True, env=e, timeout=600)
        res[tag] = dict(np.load(path))
    for k in ('dec', 'syn', 'idx', 'hdec'):
        assert np.array_equal(res['staged'][k], res['direct'][k]), k
