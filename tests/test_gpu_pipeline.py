"""wav -> codes -> wav through nsc_b200.pipeline against the oracle's composition of the same reference steps
(cmrl.py:666-737): std normalisation, high-pass, pre-emphasis, LPC windows, frames of sig[256:], CQ pass, overlap-add,
de-emphasis."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_codec, ref_framing as rf, ref_lpc, ref_nn
from util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _utterance(T, seed):
    from scipy.signal import lfilter
    rng = np.random.RandomState(seed)
    x = lfilter([1.0], [1.0, -1.6, 0.8], rng.randn(T + 200))[200:]
    return (0.05 * x).astype(np.float32)


def _oracle_pipeline(x, ocs, bins, the_share):
    std = np.std(x)
    s = x / std
    f = rf.empha_filter(rf.highpass_filter(s))
    seg = rf.utterance_to_segment(f, True)
    lsf_all = ref_lpc.lpc_analysis_at_test(seg, 16)
    seg2 = rf.utterance_to_segment(f[256:], True)
    n_used = seg2.shape[0] - 2
    fr = torch.from_numpy(seg2[:n_used].astype(np.float32))[:, :, None]
    lsf = torch.from_numpy(lsf_all[:n_used].astype(np.float32))[:, :, None]
    o = ref_codec.cq_feedforward(ocs, -300.0, bins, fr, lsf, the_share, 1.0)
    out_len = 512 + 480 * (seg.shape[0] - 2)
    syn = rf.overlap_add(np.asarray(o['synthesized'], dtype=np.float64), seg2.shape[0], n_used, out_len)
    return rf.de_empha_filter(syn) * std, n_used


def test_code_utterances_matches_oracle_composition():
    from nsc_b200 import codec, pipeline
    ocfg = ref_codec.OracleCodecCfg()
    cfg = codec.CodecConfig(resnet_type='bottleneck')
    ocs = [ref_codec.OracleCodec(ocfg, seed=5), ref_codec.OracleCodec(ocfg, seed=6)]
    gcs = [codec.NeuralCodec(cfg, torch.from_numpy(codec.pack_params_numpy(cfg, o.conv_params, o.alpha, o.bins)).to(DEV)) for o in ocs]
    cm = codec.CMRL(gcs, res_scalar=1.0)
    bins = np.load(os.path.join(GOLD, 'lsf_bins_f64.npy')).astype(np.float32)
    sigs = [_utterance(6000, 1), _utterance(9137, 2)]
    got = pipeline.code_utterances(cm, [torch.from_numpy(s).to(DEV) for s in sigs], the_share=True)
    for s, g in zip(sigs, got):
        want, n_used = _oracle_pipeline(s, ocs, bins, True)
        assert g['n_frames'] == n_used
        assert g['synthesized'].shape[0] == want.shape[0]
        assert rel_err(g['synthesized'].cpu().numpy(), want) < 2e-4    # soft path end to end (1e-4 per stage: analysis LSFs, codec, filters)


def test_hard_codes_pack_and_survive_the_round_trip():
    from nsc_b200 import bitstream, codec, pipeline
    cfg = codec.CodecConfig(resnet_type='bottleneck')
    cm = codec.CMRL([codec.NeuralCodec(cfg, device=DEV, seed=5), codec.NeuralCodec(cfg, device=DEV, seed=6)], res_scalar=1.0)
    sigs = [torch.from_numpy(_utterance(16000, 3)).to(DEV)]
    out = pipeline.code_utterances(cm, sigs, the_share=False, pack=True)[0]
    n = out['n_frames']
    assert out['records'].shape == (n, 16 + 160 + 160)
    lsf_idx, codes = bitstream.unpack_frames(out['records'], [256, 256], [32, 32])
    assert torch.equal(lsf_idx, out['lsf_idx']) and all(torch.equal(a, b) for a, b in zip(codes, out['idx']))
    assert torch.isfinite(out['synthesized']).all()
    assert out['synthesized'].shape[0] == 512 + 480 * (pipeline.ut.segment_count(16000) - 2)


def test_digital_silence_does_not_poison_the_utterance():
    """A silent stretch has no LPC solution (the reference raises in poly2lsf).  strict=True raises like the reference; the default
    substitutes the previous frame's LSFs, counts the frames, and every output sample stays finite -- before the fix a single NaN
    frame went through the de-emphasis IIR into the whole rest of the utterance."""
    from nsc_b200 import codec, pipeline
    cfg = codec.CodecConfig(resnet_type='bottleneck')
    cm = codec.CMRL([codec.NeuralCodec(cfg, device=DEV, seed=5), codec.NeuralCodec(cfg, device=DEV, seed=6)], res_scalar=1.0)
    x = _utterance(24000, 7)
    x[:6000] = 0.0                     # leading digital silence: the zero-state filters keep it exactly zero
    sig = torch.from_numpy(x).to(DEV)
    out = pipeline.code_utterances(cm, [sig], the_share=False, pack=True)[0]
    assert int(out['n_failed_lpc_frames']) > 0
    assert torch.isfinite(out['synthesized']).all() and torch.isfinite(out['decoded']).all()
    with pytest.raises(ValueError):
        pipeline.code_utterances(cm, [sig], strict=True)
    clean = pipeline.code_utterances(cm, [torch.from_numpy(_utterance(24000, 7)).to(DEV)])[0]
    assert int(clean['n_failed_lpc_frames']) == 0


def test_batched_corpus_path_is_bit_identical_to_the_per_utterance_path():
    """code_utterances analyses / synthesises utterances of equal length as one batch (one launch per step instead of one per
    utterance).  Same arithmetic per utterance: every output must match the one-by-one path to the bit, including an utterance
    with a silent stretch (forward-fill must not look across utterance boundaries) and groups of one."""
    from nsc_b200 import codec, pipeline
    cfg = codec.CodecConfig(resnet_type='bottleneck')
    cm = codec.CMRL([codec.NeuralCodec(cfg, device=DEV, seed=5), codec.NeuralCodec(cfg, device=DEV, seed=6)], res_scalar=1.0)
    quiet = _utterance(16000, 13)
    quiet[:5000] = 0.0
    sigs = [_utterance(16000, 11), _utterance(9137, 12), quiet, _utterance(6000, 14), _utterance(9137, 15), _utterance(16000, 16)]
    sigs = [torch.from_numpy(s).to(DEV) for s in sigs]
    a = pipeline.code_utterances(cm, sigs, the_share=False, pack=True)
    b = pipeline.code_utterances_one_by_one(cm, sigs, the_share=False, pack=True)
    assert len(a) == len(b) == len(sigs)
    for x, y in zip(a, b):
        assert x['n_frames'] == y['n_frames'] and int(x['n_failed_lpc_frames']) == int(y['n_failed_lpc_frames'])
        for k in ('synthesized', 'decoded', 'lsf_idx', 'records'):
            assert torch.equal(x[k], y[k]), k
        assert all(torch.equal(p, q) for p, q in zip(x['idx'], y['idx']))
    assert int(a[2]['n_failed_lpc_frames']) > 0 and int(a[0]['n_failed_lpc_frames']) == 0


def test_batch_framing_entry_points_match_the_single_utterance_ones():
    from nsc_b200 import utilities as ut
    rng = np.random.RandomState(3)
    u = torch.from_numpy(rng.randn(5, 7013).astype(np.float32)).to(DEV)
    seg = ut.utterances_to_segments(u, True, offset=256)
    win = ut.lpc_windows_at_test_batch(u)
    for i in range(5):
        assert torch.equal(seg[i], ut.utterance_to_segment(u[i], True, offset=256))
        assert torch.equal(win[i], ut.lpc_windows_at_test(u[i]))
    n2 = seg.shape[1]
    y = ut.overlap_add_batch(seg[:, :n2 - 2].contiguous(), seg_amount=n2, n_used=n2 - 2, out_len=512 + 480 * (n2 - 2))
    for i in range(5):
        assert torch.equal(y[i], ut.overlap_add(seg[i, :n2 - 2], seg_amount=n2, n_used=n2 - 2, out_len=512 + 480 * (n2 - 2)))
    # a truncated take keeps the leading frames
    assert torch.equal(ut.utterances_to_segments(u, False, n_take=3), torch.stack([ut.utterance_to_segment(r, False)[:3] for r in u]))
