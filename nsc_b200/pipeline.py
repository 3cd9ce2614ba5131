"""Utterance-level collaborative-quantisation coding: the loop of CMRL.cmrl_eval_lpc (cmrl.py:666-737) with every step on
the GPU and the per-frame ``sess.run`` replaced by ONE batched pass over the frames of all utterances.

    per utterance:  s / std(s)                                  _load_sig_lpc           nscm.py:98-113
                    high-pass -> pre-emphasis                   cmrl.py:671
                    1024-sample LPC windows -> 16 LSFs          lpc_analysis_at_test    cmrl.py:693
                    frames of sig[256:] (hop 480)               cmrl.py:695
    all frames:     LSF codebook -> residual -> CMRL cascade -> synthesis   (CMRL.feedforward_lpc)
    per utterance:  trapezoid-Hann overlap-add of the first N2 - 2 frames   cmrl.py:698-716
                    de-emphasis, * std                          cmrl.py:735-737
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from . import bitstream, lpc_utilities as lu, utilities as ut
from .constants import frame_length

HOP = ut.hop_size


def analysis(sig: torch.Tensor, strict: bool = False) -> Dict[str, object]:
    """One utterance (T,) float32 on the GPU -> filtered signal, frames to code, their LSFs and the bookkeeping counts.
    A silent (all-zero) LPC window has no LPC solution: the reference raises inside poly2lsf there.  strict=True raises like the
    reference; otherwise such frames take the previous frame's LSFs (the first one: the uniform LSF grid of A(z) = 1) -- with a
    silent residual the choice is inaudible -- and their count is returned as 'n_failed', so that a NaN never reaches the
    de-emphasis IIR (which would spread it over the rest of the utterance)."""
    s, std = ut.load_sig_lpc(sig)
    f = ut.empha_filter(ut.highpass_filter(s))
    T = f.numel()
    n_seg = ut.segment_count(T)                     # frames of the whole signal (sizes the output arrays, cmrl.py:674-676)
    n_seg2 = ut.segment_count(T - 256)              # frames of sig[256:]                                 (cmrl.py:695)
    n_used = max(n_seg2 - 2, 0)                     # the loop runs range(N2 - 2)                           (cmrl.py:698)
    frames = ut.utterance_to_segment(f, True, offset=256)[:n_used]
    lsf = lu.lpc_analysis_windows(ut.lpc_windows_at_test(f), dtype=torch.float32)[:n_used]
    bad = ~torch.isfinite(lsf).all(dim=1)
    n_failed = bad.sum()                               # stays on the device: no host synchronisation per utterance
    if strict and lsf.numel() and int(n_failed):
        raise ValueError(f"LPC analysis failed on {int(n_failed)} silent frame(s) (the reference raises in poly2lsf)")
    if lsf.numel():
        # forward-fill: index of the last good frame at or before each frame (-1 -> the neutral grid); a no-op without failures
        idx = torch.arange(lsf.shape[0], device=lsf.device)
        last = torch.cummax(torch.where(bad, torch.full_like(idx, -1), idx), dim=0).values
        neutral = (torch.arange(1, lsf.shape[1] + 1, device=lsf.device, dtype=lsf.dtype) * (torch.pi / (lsf.shape[1] + 1)))
        filled = torch.where((last >= 0)[:, None], torch.nan_to_num(lsf[last.clamp(min=0)]), neutral[None, :].expand_as(lsf))
        lsf = torch.where(bad[:, None], filled, lsf)
    return {'std': std, 'filtered': f, 'frames': frames, 'lsf': lsf, 'n_seg': n_seg, 'n_seg2': n_seg2, 'n_used': n_used,
            'n_failed': n_failed}


def synthesis(frames: torch.Tensor, n_seg: int, n_seg2: int, std, de_emphasis: bool = True) -> torch.Tensor:
    """Overlap-add the coded frames of one utterance back into a signal (cmrl.py:710-716, :735-737)."""
    out_len = frame_length + HOP * (n_seg - 2)
    y = ut.overlap_add(frames, seg_amount=n_seg2, n_used=frames.shape[0], out_len=max(out_len, 0))
    if de_emphasis:
        y = ut.de_empha_filter(y)
    return y * std


def _fill_failed_lsf(lsf: torch.Tensor):
    """lsf (n, m, 16): frames without an LPC solution take the last good frame's LSFs OF THEIR OWN UTTERANCE (else the neutral
    grid); returns (filled, failures per utterance).  Same rule as ``analysis`` above, one row per utterance."""
    n, m, order = lsf.shape
    bad = ~torch.isfinite(lsf).all(dim=2)
    idx = torch.arange(m, device=lsf.device).expand(n, m)
    last = torch.cummax(torch.where(bad, torch.full_like(idx, -1), idx), dim=1).values
    neutral = torch.arange(1, order + 1, device=lsf.device, dtype=lsf.dtype) * (torch.pi / (order + 1))
    prev = torch.nan_to_num(torch.gather(lsf, 1, last.clamp(min=0)[:, :, None].expand(n, m, order)))
    filled = torch.where((last >= 0)[:, :, None], prev, neutral.expand(n, m, order))
    return torch.where(bad[:, :, None], filled, lsf), bad.sum(dim=1)


def analysis_batch(sigs: torch.Tensor, strict: bool = False) -> Dict[str, object]:
    """``analysis`` over a batch (n, T) of EQUAL-LENGTH utterances: the same arithmetic per utterance, one launch per step
    for the whole batch (filters, framing, window cutting, LPC analysis) instead of one per utterance.  'frames' (n, n_used, 512),
    'lsf' (n, n_used, 16), 'std' (n,), 'n_failed' (n,)."""
    if sigs.dim() != 2:
        raise ValueError("analysis_batch expects (n_utterances, T)")
    std = torch.stack([s.std(unbiased=False) for s in sigs])        # per utterance exactly as load_sig_lpc computes it
    f = ut.empha_filter(ut.highpass_filter(sigs / std[:, None]))
    n, T = f.shape
    n_seg = ut.segment_count(T)
    n_seg2 = ut.segment_count(T - 256)
    n_used = max(n_seg2 - 2, 0)
    frames = ut.utterances_to_segments(f, True, offset=256, n_take=n_used)
    win = ut.lpc_windows_at_test_batch(f, n_take=n_used)
    lsf = lu.lpc_analysis_windows(win, dtype=torch.float32).reshape(n, n_used, -1)
    if n_used:
        lsf, n_failed = _fill_failed_lsf(lsf)
        if strict and int(n_failed.sum()):
            raise ValueError(f"LPC analysis failed on {int(n_failed.sum())} silent frame(s) (the reference raises in poly2lsf)")
    else:
        n_failed = torch.zeros(n, dtype=torch.long, device=f.device)
    return {'std': std, 'filtered': f, 'frames': frames, 'lsf': lsf, 'n_seg': n_seg, 'n_seg2': n_seg2, 'n_used': n_used,
            'n_failed': n_failed}


def synthesis_batch(frames: torch.Tensor, n_seg: int, n_seg2: int, std, de_emphasis: bool = True) -> torch.Tensor:
    """``synthesis`` over a batch: frames (n, n_used, 512) -> (n, out_len); std a scalar or (n,)."""
    out_len = frame_length + HOP * (n_seg - 2)
    y = ut.overlap_add_batch(frames, seg_amount=n_seg2, n_used=frames.shape[1], out_len=max(out_len, 0))
    if de_emphasis and y.numel():
        y = ut.de_empha_filter(y)
    return y * (std[:, None] if torch.is_tensor(std) and std.dim() == 1 else std)


def code_utterances(cm, signals: Sequence[torch.Tensor], the_share: bool = False, pack: bool = False,
                    strict: bool = False) -> List[Dict[str, object]]:
    """Encode + decode a list of utterances with ``cm`` (a codec.CMRL).  Returns per utterance:
    'synthesized' (time domain, de-emphasised, rescaled), 'decoded' (overlap-added residual-domain signal), 'lsf_idx',
    'idx' (per codec) and, with ``pack``, the fixed-width 'records' of bitstream.pack_frames.
    Utterances of equal length are analysed and synthesised as one batch (the per-utterance arithmetic is unchanged; at 10 s per
    utterance the 14 small launches per utterance were a third of a corpus pass); ALL frames go through the codec in one call."""
    groups: Dict[int, List[int]] = {}
    for i, s in enumerate(signals):
        groups.setdefault(int(s.numel()), []).append(i)
    ana = []
    for T, members in groups.items():
        a = analysis_batch(torch.stack([signals[i].reshape(-1) for i in members]), strict)
        a['members'] = members
        ana.append(a)
    total = sum(a['n_used'] * len(a['members']) for a in ana)
    if total == 0:
        return [{'synthesized': torch.zeros(0, device=s.device), 'decoded': torch.zeros(0, device=s.device)} for s in signals]
    frames = torch.cat([a['frames'].reshape(-1, frame_length) for a in ana])
    lsf = torch.cat([a['lsf'].reshape(-1, a['lsf'].shape[-1]) for a in ana])
    r = cm.feedforward_lpc(frames, lsf, the_share, 1.0)
    records = None
    if pack:
        records = bitstream.pack_frames(r['lsf_idx'], r['idx'], [c.cfg.num_bins for c in cm.codecs], cm.n_lsf_bins)
    out: List[Dict[str, object]] = [None] * len(signals)
    o = 0
    for a in ana:
        m, n = len(a['members']), a['n_used']
        sl = slice(o, o + m * n)
        syn = synthesis_batch(r['synthesized'][sl].reshape(m, n, frame_length), a['n_seg'], a['n_seg2'], a['std'])
        dec = synthesis_batch(r['decoded'][sl].reshape(m, n, frame_length), a['n_seg'], a['n_seg2'], 1.0, de_emphasis=False)
        for k, i in enumerate(a['members']):
            u = slice(o + k * n, o + (k + 1) * n)
            d = {'synthesized': syn[k], 'decoded': dec[k], 'lsf_idx': r['lsf_idx'][u], 'idx': [c[u] for c in r['idx']],
                 'n_frames': n, 'n_failed_lpc_frames': a['n_failed'][k]}
            if records is not None:
                d['records'] = records[u]
            out[i] = d
        o += m * n
    return out


def code_utterances_one_by_one(cm, signals: Sequence[torch.Tensor], the_share: bool = False, pack: bool = False,
                               strict: bool = False) -> List[Dict[str, object]]:
    """The same as ``code_utterances`` with per-utterance analysis / synthesis launches (the batched form is tested against it)."""
    ana = [analysis(s, strict) for s in signals]
    counts = [a['n_used'] for a in ana]
    if sum(counts) == 0:
        return [{'synthesized': torch.zeros(0, device=s.device), 'decoded': torch.zeros(0, device=s.device)} for s in signals]
    frames = torch.cat([a['frames'] for a in ana])
    lsf = torch.cat([a['lsf'] for a in ana])
    r = cm.feedforward_lpc(frames, lsf, the_share, 1.0)
    records = None
    if pack:
        records = bitstream.pack_frames(r['lsf_idx'], r['idx'], [c.cfg.num_bins for c in cm.codecs], cm.n_lsf_bins)
    out, o = [], 0
    for a, n in zip(ana, counts):
        sl = slice(o, o + n)
        d = {
            'synthesized': synthesis(r['synthesized'][sl], a['n_seg'], a['n_seg2'], a['std']),
            'decoded': synthesis(r['decoded'][sl], a['n_seg'], a['n_seg2'], 1.0, de_emphasis=False),
            'lsf_idx': r['lsf_idx'][sl], 'idx': [i[sl] for i in r['idx']], 'n_frames': n, 'n_failed_lpc_frames': a['n_failed'],
        }
        if records is not None:
            d['records'] = records[sl]
        out.append(d)
        o += n
    return out
