"""Oracle restatement of the codec topology and CMRL cascade (torch-CPU).  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/neural_speech_coding_module.py:152-335 (one codec) and
/root/reference/cmrl.py:513-543, :770-858 (cascade, CQ feed-forward), plus the loss assembly of
nscm.py:1033-1059 / cmrl.py:464-490 and TF1's Adam update [LIB].
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ref_loss, ref_lpc, ref_nn
from .ref_nn import ParamStream

FRAME_LENGTH = 512
INIT_ALPHA = -300.0     # constants.py:5
LPC_ORDER = 16          # nscm.py:50


@dataclass
class OracleCodecCfg:
    bottleneck_kernel_and_dilation: Sequence[int] = (9, 9, 100, 20, 1, 2)   # README.md:75
    strides: Sequence[int] = (2,)                                           # already expanded ([2] or [2,2])
    resnet_type: str = 'bottleneck'                                        # constants.py:13-14
    num_bins: int = 32


class OracleCodec:
    """One `scope_k` of the reference graph: conv parameters in creation order + alpha + bins."""

    def __init__(self, cfg: OracleCodecCfg, seed: int = 0, conv_params=None, alpha=None, bins=None):
        self.cfg = cfg
        self.ps = ParamStream(conv_params, seed=seed)
        self.alpha = np.float32(INIT_ALPHA if alpha is None else alpha)
        # nscm.py:269 / :308 -- np.linspace(-1, 1, n) cast to float32
        self.bins = (np.linspace(-1, 1, cfg.num_bins) if bins is None else np.asarray(bins)).astype(np.float32)
        if self.ps.init_mode:
            # run the graph once on a dummy frame to create the variables in TF's order
            self.forward(torch.zeros(1, FRAME_LENGTH, 1), the_share=False, is_quan_on=1.0)
            self.ps.init_mode = False

    @property
    def conv_params(self):
        return self.ps.params

    # --- nscm.py:152-156
    def _down_sampling_mod(self, x, the_stride=2):
        y = ref_nn.conv1d(x, self.cfg.bottleneck_kernel_and_dilation[2], 9, dilation_rate=1, strides=the_stride,
                          activation=None, ps=self.ps)
        return ref_nn.activation_func(y)

    # --- nscm.py:158-167
    @staticmethod
    def _up_sampling_mod_helper(x, the_stride=2):
        B, L, C = x.shape
        r = x.reshape(B, L, C // the_stride, the_stride).permute(0, 1, 3, 2)
        return r.reshape(B, L * the_stride, C // the_stride)

    # --- nscm.py:169-181
    def _up_sampling_mod(self, x, the_stride=2):
        if self.cfg.resnet_type == 'bottleneck':
            y = ref_nn.conv1d(x, x.shape[-1], 9, dilation_rate=1, strides=1, activation=None, ps=self.ps)
        else:
            y = ref_nn.conv1d_depth(x, x.shape[-1], 9, dilation_rate=1, strides=1, activation=None, ps=self.ps)
        y = ref_nn.activation_func(y)
        return self._up_sampling_mod_helper(y, the_stride)

    # --- nscm.py:183-217
    def _stack_bottleneck_blocks(self, x, strides=1, is_post_up_samling=True):
        cfg = self.cfg.bottleneck_kernel_and_dilation
        assert cfg[2] % strides == 0
        if x.shape[-1] == 1:
            wide = cfg[2]
        else:
            wide = int(x.shape[-1] / strides) if is_post_up_samling else x.shape[-1]
        for i in range(len(cfg) - 4):
            flag = i == (len(cfg) - 5)
            block = ref_nn.the_bottleneck if self.cfg.resnet_type == 'bottleneck' else ref_nn.gated_bottleneck
            x = block(x, non_dilated_neck_kernel_size=cfg[1], dilated_neck_kernel_size=cfg[0], wide_layer=wide,
                      narrow_layer=cfg[3], dilation_rate=cfg[i + 4], is_last_flat=flag, ps=self.ps)
        return x

    # --- nscm.py:219-237
    def encoder(self, x):
        cfg = self.cfg.bottleneck_kernel_and_dilation
        y = ref_nn.change_channel(x, the_channel=cfg[2], kernel_size=55, activation=None, ps=self.ps)
        y = ref_nn.activation_func(y)
        for s in self.cfg.strides:
            y = self._stack_bottleneck_blocks(y, is_post_up_samling=False)
            y = self._down_sampling_mod(y, the_stride=s)
        y = self._stack_bottleneck_blocks(y, is_post_up_samling=False)
        return ref_nn.change_channel(y, the_channel=1, kernel_size=55, activation='tanh', ps=self.ps)

    # --- nscm.py:239-260
    def decoder(self, code):
        y = code
        for s in self.cfg.strides:
            y = self._stack_bottleneck_blocks(y, is_post_up_samling=False)
            y = self._up_sampling_mod(y, the_stride=s)
        y = self._stack_bottleneck_blocks(y, is_post_up_samling=False)
        return ref_nn.change_channel(y, the_channel=1, kernel_size=55, activation=None, ps=self.ps)

    # --- nscm.py:262-295 / :297-335
    def forward(self, x, the_share, is_quan_on, alpha=None, bins=None):
        """x (B,512,1) -> dict(soft (B,Lc,n), floating_code (B,Lc,1), code (B,Lc,1), out (B,512))."""
        self.ps._cursor = 0
        alpha = self.alpha if alpha is None else alpha
        bins = self.bins if bins is None else bins
        floating = self.encoder(x)
        code_len = FRAME_LENGTH // (2 ** len(self.cfg.strides))
        soft, code = ref_nn.scalar_softmax_quantization(floating, alpha, bins, is_quan_on, the_share, code_len,
                                                        self.cfg.num_bins)
        out = self.decoder(code)
        self.ps.done()
        return dict(soft=soft, floating_code=floating, code=code, out=out[:, :, 0])


def cascade_forward(codecs: List[OracleCodec], x, the_share, is_quan_on, res_scalar=1.0, lpc_variant=False):
    """cmrl.py:513-543 (lpc_variant=False) and the loop of cmrl.py:806-830 (lpc_variant=True).

    The two differ only in codec 0: the LPC variant multiplies its input by res_scalar and divides its
    output (cmrl.py:810, :818); the plain variant feeds x unscaled and does not divide (cmrl.py:522-528).
    """
    outs, per = [], []
    for i, c in enumerate(codecs):
        if i == 0:
            inp = x * res_scalar if lpc_variant else x
        else:
            inp = res_scalar * (x - torch.stack(outs, 0).sum(0).unsqueeze(2))
        r = c.forward(inp, the_share, is_quan_on)
        o = r['out']
        if i > 0 or lpc_variant:
            o = o / res_scalar
        outs.append(o)
        per.append(r)
    decoded = torch.stack(outs, 0).sum(0)
    return decoded, outs, per


def cq_feedforward(codecs, lsf_alpha, lsf_bins, x, lpc_x, the_share, is_quan_on, res_scalar=1.0):
    """cmrl.py:770-858: LSF quantiser -> lsf2poly -> residual -> cascade -> synthesis (+ report losses)."""
    dt = x.dtype
    soft_lpc, q_lsf = ref_nn.scalar_softmax_quantization(lpc_x, lsf_alpha, lsf_bins, is_quan_on, the_share,
                                                         LPC_ORDER, len(lsf_bins))
    q_lsf = q_lsf[:, :, 0].reshape(-1, LPC_ORDER)
    poly = ref_lpc.lsf2poly_after_quan(q_lsf.detach().numpy(), LPC_ORDER)            # py_func, float32 out
    res = ref_lpc.lpc_analysis_get_residual(x.detach().numpy(), poly)                 # py_func, float32 out
    res_x = torch.as_tensor(res).to(dt).reshape(-1, FRAME_LENGTH, 1)
    decoded, outs, per = cascade_forward(codecs, res_x, the_share, is_quan_on, res_scalar, lpc_variant=True)
    synthesized = ref_lpc.lpc_synthesizer_tr(poly, decoded.detach().to(torch.float32).numpy())
    time_loss = ref_loss.mse_loss(decoded, res_x[:, :, 0])
    freq_loss = ref_loss.mfcc_loss(decoded, res_x[:, :, 0])
    ent_lpc = ref_loss.entropy_coding_loss(soft_lpc)
    ent = [ref_loss.entropy_coding_loss(p['soft']) for p in per]
    return dict(soft_lpc=soft_lpc, q_lsf=q_lsf, poly=poly, res_x=res_x, decoded=decoded, outs=outs, per=per,
                synthesized=synthesized, time_loss=time_loss, freq_loss=freq_loss, ent_lpc=ent_lpc, ent=ent)


# ----------------------------------------------------------------------------------------------
# training-step semantics
# ----------------------------------------------------------------------------------------------
def tf1_adam_step(theta, grad, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """TF1 AdamOptimizer [LIB]: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps). t starts at 1."""
    m = beta1 * m + (1.0 - beta1) * grad
    v = beta2 * v + (1.0 - beta2) * grad * grad
    lr_t = lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    theta = theta - lr_t * m / (np.sqrt(v) + eps)
    return theta, m, v


def cq_training_objective(codecs, lsf_alpha, lsf_bins, res_x, lpc_x, is_quan_on, coeff, quan_w, ent_w, tau, res_scalar=1.0,
                          global_batch=None):
    """The scalar whose gradient `minimize` applies: sum over the batch of the per-frame loss vector,
    with the scalar entropy term broadcast (counted once per frame).  res_x is FED (nscm.py:586-595); the LSF codebook only
    enters through quan_loss / entropy of its own soft assignment.  All arguments may be torch tensors that require grad.
    quan_w / ent_w: index 0 = LSF codebook, 1.. = codecs."""
    B = res_x.shape[0]
    soft_lpc, _ = ref_nn.scalar_softmax_quantization(lpc_x, lsf_alpha, lsf_bins, is_quan_on, True, LPC_ORDER, len(lsf_bins))
    decoded, outs, per = cascade_forward(codecs, res_x, True, is_quan_on, res_scalar, lpc_variant=True)
    time_loss = ref_loss.mse_loss(decoded, res_x[:, :, 0])
    freq_loss = ref_loss.mfcc_loss(decoded, res_x[:, :, 0])
    softs = [soft_lpc] + [p['soft'] for p in per]
    quan = sum(w * ref_loss.quan_loss(s) for w, s in zip(quan_w, softs))
    ent = sum(w * ref_loss.entropy_coding_loss(s) for w, s in zip(ent_w, softs))
    vec = coeff[0] * time_loss + coeff[1] * freq_loss + coeff[2] * quan + tau * ent
    gb = B if global_batch is None else global_batch
    total = (coeff[0] * time_loss + coeff[1] * freq_loss + coeff[2] * quan).sum() + gb * tau * ent
    return total, dict(vec=vec, time=time_loss, freq=freq_loss, quan=quan, ent=ent, decoded=decoded)
