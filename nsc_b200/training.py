"""Data-parallel training step of the CQ cascade over the C ABI.

Mirrors the loss assembly and optimiser of the reference graphs -- not their epoch loops, dataset paths or Saver:
    one_ae_lpc        nscm.py:1033-1059   quan/entropy weights {16, 256}/272, tau fed per step
    _finetuning_lpc   cmrl.py:464-490     unweighted quan terms, entropy term dropped (commented out at :485)
Facts reproduced: soft value path; per-frame loss VECTORS whose SUM is differentiated; the scalar entropy term counted
once per frame of the (global) batch; no gradient through lsf2poly / residual / synthesis; res_x is fed; TF1 Adam.

Multi-GPU: one process per GPU, frames sharded by rank (equal shard sizes).  Every codec's gradient, the LSF codebook's gradient
and the soft histograms live in ONE flat fp32 buffer (< 1 M floats) that is SUM all-reduced ONCE per step over torch.distributed
(NCCL on B200) -- `collectives_per_step == 1`.  Only when an entropy weight is non-zero (one_ae_lpc; `_finetuning_lpc` drops the
term, cmrl.py:485) does the backward pass need the batch-GLOBAL histograms before it starts (the entropy is a function of the global
histogram); then the few hundred histogram floats are all-reduced ahead of it (`collectives_per_step == 2`).
Nothing in the step synchronises the host: the global batch is world_size x the per-rank batch.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib
from .codec import CMRL, FRAME
from .loss_terms_and_measures import entropy_from_hist, mel_filterbank


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class CQTrainer:
    def __init__(self, cmrl: CMRL, coeff_term: Sequence[float] = (60.0, 10.0, 10.0, 0.0),
                 quan_w: Optional[Sequence[float]] = None, ent_w: Optional[Sequence[float]] = None,
                 trainable: Optional[Sequence[bool]] = None, train_lsf: bool = True,
                 lr: float = 2e-4, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
        """coeff_term = (c_time, c_freq, c_quan, tau0) as in `--coeff_term '60 10 10 0'` (README.md:78).
        quan_w / ent_w: per-quantiser weights, index 0 = LSF codebook, 1.. = codecs."""
        self.cm = cmrl
        n = len(cmrl.codecs)
        self.c = [float(v) for v in coeff_term]
        self.quan_w = [1.0] * (n + 1) if quan_w is None else [float(v) for v in quan_w]
        self.ent_w = [0.0] * (n + 1) if ent_w is None else [float(v) for v in ent_w]
        tr = [True] * n if trainable is None else [bool(v) for v in trainable]
        self.trainable = [bool(train_lsf)] + tr
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        dev = cmrl.lsf_params.device
        # ONE flat buffer: [codec 0 grads | codec 1 grads | ... | LSF codebook grads | LSF histogram | codec histograms]
        sizes = [c.params.numel() for c in cmrl.codecs] + [cmrl.lsf_params.numel()]
        hsizes = [cmrl.n_lsf_bins] + [c.cfg.num_bins for c in cmrl.codecs]
        # (every segment starts on a 256-byte boundary, like the separately allocated tensors it replaces; the padding stays zero)
        al = lambda v: (v + 63) // 64 * 64
        self._n_grad = sum(al(sz) for sz in sizes)
        self.flat = torch.zeros(self._n_grad + sum(al(sz) for sz in hsizes), dtype=torch.float32, device=dev)
        views, off = [], 0
        for sz in sizes + hsizes:
            views.append(self.flat[off:off + sz]); off += al(sz)
        self.grads = views[:n]
        self.lsf_grad = views[n]
        self.hists = views[n + 1:]
        # two optimizers with separate slots, like the reference's two AdamOptimizer(...).minimize ops (nscm.py:1055-1059):
        # 'quan' minimises loss_quan, 'no_quan' (pre-training epochs) minimises c0 time + c1 freq only
        self._slots = {name: {'m': [torch.zeros_like(c.params) for c in cmrl.codecs] + [torch.zeros_like(cmrl.lsf_params)],
                              'v': [torch.zeros_like(c.params) for c in cmrl.codecs] + [torch.zeros_like(cmrl.lsf_params)], 't': 0}
                       for name in ('quan', 'no_quan')}
        self.m, self.v = self._slots['quan']['m'], self._slots['quan']['v']
        self._ws: Optional[torch.Tensor] = None
        self._dev = dev

    @property
    def needs_global_hist(self) -> bool:
        """The backward pass differentiates tau * entropy(global histogram) only when an entropy weight is non-zero."""
        return any(w != 0.0 for w in self.ent_w)

    @property
    def collectives_per_step(self) -> int:
        return 0 if not _dist_on() else (2 if self.needs_global_hist else 1)

    @staticmethod
    def one_ae_lpc(cmrl: CMRL, coeff_term=(60.0, 10.0, 10.0, 0.0), is_cq: bool = True, **kw) -> "CQTrainer":
        """nscm.py:1033-1053 (first codec + LSF codebook; weights 16/272 and 256/272 at stride 2)."""
        Lc = cmrl.codecs[0].cfg.code_length
        w = [16.0 / (16.0 + Lc), Lc / (16.0 + Lc)]
        return CQTrainer(cmrl, coeff_term, quan_w=w, ent_w=w, train_lsf=is_cq, **kw)

    @staticmethod
    def finetuning_lpc(cmrl: CMRL, coeff_term=(60.0, 10.0, 10.0, 0.0), **kw) -> "CQTrainer":
        """cmrl.py:464-490: all scopes trainable, quan terms unweighted, no entropy term."""
        n = len(cmrl.codecs)
        return CQTrainer(cmrl, coeff_term, quan_w=[1.0] * (n + 1), ent_w=[0.0] * (n + 1), **kw)

    # ------------------------------------------------------------------------------------------
    def _workspace(self, B: int) -> torch.Tensor:
        need = int(_lib.load().nsc_train_workspace_bytes(self.cm._cfgs, len(self.cm.codecs), B))
        if need < 0:
            _lib.check(-1, 'nsc_train_workspace_bytes')
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self._dev)
        return self._ws

    def loss_and_grads(self, res_x, lpc_x, tau: Optional[float] = None, is_quan_on: float = 1.0, quan_terms: bool = True) -> Dict[str, object]:
        """One forward + backward on this rank's shard.  Gradients (of the batch-SUM objective over the GLOBAL batch)
        land in self.grads / self.lsf_grad, already all-reduced when torch.distributed is initialised.
        quan_terms=False differentiates loss_no_quan = c0 time + c1 freq (the pre-training op, nscm.py:1049): c2 and tau are zero."""
        lib = _lib.load()
        cm, n = self.cm, len(self.cm.codecs)
        x = _lib.require_f32(res_x, 'res_x').reshape(-1, FRAME)
        lsf = _lib.require_f32(lpc_x, 'lpc_x').reshape(-1, _lib.LPC_ORDER)
        B, dev = x.shape[0], x.device
        tau = self.c[3] if tau is None else float(tau)
        c2 = self.c[2]
        if not quan_terms:
            c2, tau = 0.0, 0.0
        ws = self._workspace(B)
        melw = mel_filterbank(dev)
        decoded = torch.empty((B, FRAME), dtype=torch.float32, device=dev)
        time_l = torch.empty(B, dtype=torch.float32, device=dev)
        freq_l = torch.empty(B, dtype=torch.float32, device=dev)
        qloss = [torch.empty(B, dtype=torch.float32, device=dev) for _ in range(n + 1)]
        hists = self.hists
        self.flat[self._n_grad:].zero_()
        params = _lib.ptr_array([c.params for c in cm.codecs])
        rc = lib.nsc_train_forward(cm._cfgs, n, params, _lib.ptr(cm.lsf_params), cm.n_lsf_bins, _lib.ptr(x), _lib.ptr(lsf), B,
                                   cm.res_scalar, float(is_quan_on), _lib.ptr(melw), _lib.ptr(decoded), _lib.ptr(time_l),
                                   _lib.ptr(freq_l), _lib.ptr_array(qloss), _lib.ptr_array(hists), _lib.ptr(ws), ws.numel(),
                                   _lib.stream_ptr())
        _lib.check(rc, 'nsc_train_forward')
        # equal shards: the global batch needs no collective and no host synchronisation
        global_B = B * (dist.get_world_size() if _dist_on() else 1)
        hists_early = _dist_on() and self.needs_global_hist
        if hists_early:
            dist.all_reduce(self.flat[self._n_grad:], op=dist.ReduceOp.SUM)   # batch-global soft histograms BEFORE the backward pass
        coeff = (C.c_float * 4)(self.c[0], self.c[1], c2, tau)
        qw = (C.c_float * (n + 1))(*self.quan_w)
        ew = (C.c_float * (n + 1))(*self.ent_w)
        tr = (C.c_int32 * (n + 1))(*[int(v) for v in self.trainable])
        rc = lib.nsc_train_backward(cm._cfgs, n, params, _lib.ptr(cm.lsf_params), cm.n_lsf_bins, _lib.ptr(x), _lib.ptr(lsf), B,
                                    cm.res_scalar, float(is_quan_on), _lib.ptr(melw), _lib.ptr(decoded), coeff, qw, ew, global_B,
                                    _lib.ptr_array(hists), tr, _lib.ptr_array(self.grads), _lib.ptr(self.lsf_grad), _lib.ptr(ws),
                                    ws.numel(), _lib.stream_ptr())
        _lib.check(rc, 'nsc_train_backward')
        if _dist_on():
            # THE all-reduce of the step (SUM, not mean): all gradients -- and the histograms when the backward did not need them
            dist.all_reduce(self.flat[:self._n_grad] if hists_early else self.flat, op=dist.ReduceOp.SUM)
        ent = [entropy_from_hist(h) for h in hists]
        quan = sum(w * q for w, q in zip(self.quan_w, qloss))
        ent_term = sum(w * e for w, e in zip(self.ent_w, ent))
        return {'decoded': decoded, 'time_loss': time_l, 'freq_loss': freq_l, 'quan_loss': quan, 'ent_loss': ent_term,
                'entropies': ent, 'hists': [h.clone() for h in hists], 'global_batch': global_B,
                'loss_vector': self.c[0] * time_l + self.c[1] * freq_l + c2 * quan + tau * ent_term}

    def apply_adam(self, lr: Optional[float] = None, optimizer: str = 'quan') -> None:
        """TF1 AdamOptimizer update of every trainable scope (separate slots per scope, like tf's per-variable slots; one set of
        slots and one step counter per optimizer, 'quan' or 'no_quan')."""
        lib = _lib.load()
        sl = self._slots[optimizer]
        sl['t'] += 1
        self.t = self._slots['quan']['t']
        lr = self.lr if lr is None else lr
        items = [(c.params, g) for c, g in zip(self.cm.codecs, self.grads)] + [(self.cm.lsf_params, self.lsf_grad)]
        flags = self.trainable[1:] + [self.trainable[0]]
        for (p, g), m, v, on in zip(items, sl['m'], sl['v'], flags):
            if not on:
                continue
            _lib.check(lib.nsc_adam_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), p.numel(), float(lr), sl['t'],
                                         self.b1, self.b2, self.eps, _lib.stream_ptr()), 'nsc_adam_step')

    def step(self, res_x, lpc_x, tau: Optional[float] = None, is_quan_on: float = 1.0, lr: Optional[float] = None,
             optimizer: str = 'quan'):
        """One training step.  optimizer='no_quan' is the pre-training op of the reference (loss_no_quan, fed is_quan_on = 0,
        checkpoint.schedule(...)['optimizer'])."""
        out = self.loss_and_grads(res_x, lpc_x, tau, is_quan_on, quan_terms=(optimizer == 'quan'))
        self.apply_adam(lr, optimizer)
        return out


def update_lpc_residual(raw_frames: torch.Tensor, lsf: torch.Tensor, lsf_params: torch.Tensor, chunk: int = 50000,
                        init_alpha: Optional[float] = None) -> torch.Tensor:
    """_update_lpc_residual (nscm.py:1075-1122): every 30 epochs (and at epoch - 3; checkpoint.schedule) a CQ run recomputes the LPC
    residual of the WHOLE training set with the codebook it has learned so far, and trains on those residuals from then on.

    As the reference does it: the learned LSF bins are re-read SORTED (:1084-1087), alpha is a fresh `init_alpha` (not the learned
    one, :1083), the assignment is HARD (`the_share: False`, is_quan_on = 1, :1088-1094), then lsf2poly_after_quan and
    lpc_analysis_get_residual per chunk of 50,000 frames (:1103-1119).  raw_frames (N, 512) float32, lsf (N, 16) float32 (the stored
    un-quantised LSFs), lsf_params = {alpha, bins[n]} of the `lpc_quan` scope -> residual (N, 512) float32.  Device in, device out."""
    from . import constants as _c
    from . import lpc_utilities as lu
    from . import nn_core_operator as nn
    x = _lib.require_f32(raw_frames, 'raw_frames').reshape(-1, FRAME)
    l = _lib.require_f32(lsf, 'lsf').reshape(-1, _lib.LPC_ORDER)
    if x.shape[0] != l.shape[0]:
        raise ValueError("raw_frames / lsf batch mismatch")
    bins = torch.sort(_lib.require_f32(lsf_params, 'lsf_params')[1:]).values.contiguous()
    alpha = float(_c.init_alpha if init_alpha is None else init_alpha)
    out = torch.empty_like(x)
    for s0 in range(0, x.shape[0], int(chunk)):
        s1 = min(s0 + int(chunk), x.shape[0])
        _, q = nn.scalar_softmax_quantization(l[s0:s1, :, None].contiguous(), alpha, bins, 1.0, False, _lib.LPC_ORDER, bins.numel())
        poly = lu.lsf2poly_after_quan(q[:, :, 0].contiguous(), _lib.LPC_ORDER)
        out[s0:s1] = lu.lpc_analysis_get_residual(x[s0:s1, :, None].contiguous(), poly)
    return out
