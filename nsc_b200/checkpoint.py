"""TF-variable import and the training-schedule shell.

The reference keeps its weights in `tf.compat.v1.train.Saver` checkpoints (nscm.py:548, :652; cmrl.py:64-73) -- none ship,
and TensorFlow is not installable here, so the importer works on a plain ``{variable name: ndarray}`` mapping that a
maintainer dumps on a TensorFlow machine with

    r = tf.train.load_checkpoint('./check/model_bnn_ac_<id>_<save_id>.ckpt')
    np.savez('nsc_vars.npz', **{n: r.get_tensor(n) for n in r.get_variable_to_shape_map()})

[LIB] variable names of `tf.compat.v1.layers.conv1d` inside `variable_scope(scope)`: the i-th conv created in the scope is
``<scope>/conv1d/{kernel,bias}`` for i = 0 and ``<scope>/conv1d_<i>/{kernel,bias}`` after that (names are uniquified per enclosing
variable scope).  Keras `SeparableConv1D` (the 'gln' up-conv, nscm.py:175-177) is different: Keras layer names are uniquified per
GRAPH, not per variable scope, so in a multi-codec graph (cmrl.py all_modules_feedforward[_lpc], follower training) the j-th
separable conv created ANYWHERE in the graph is ``<scope>/separable_conv1d[_<j>]/{depthwise_kernel,pointwise_kernel,bias}`` --
codec 2's up-conv is ``scope_2/separable_conv1d_1`` (and further suffixes with the_strides = '4').  `sep_start` carries that
running count; `params_from_tf_variables` additionally resolves the layer by pattern inside the scope when the exact name is absent.
UNVERIFIED against a real TensorFlow 2.0 dump (TensorFlow is not installable here): derived from TF's naming rules.
Quantiser variables ``<scope>/alpha``, ``<scope>/bins`` (nscm.py:267-269, :305-308); the LSF codebook lives in scope
``lpc_quan`` (nscm.py:994-997).
Adam slot variables (``.../Adam``, ``.../Adam_1``) are ignored, like the reference's stage-to-stage restores (cmrl.py:116-119).

The schedule shell reproduces the entropy controller of the training loops (nscm.py:630-639, :494-518) and the
sorted re-read of the learned LSF bins by `_update_lpc_residual` (nscm.py:1084-1087).
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Sequence, Tuple

import numpy as np

from .codec import CodecConfig, layer_table, pack_params_numpy


def separable_count(cfg: CodecConfig) -> int:
    """Separable convs one codec creates (one per up-sampling stage of a 'gln' codec) -- the increment of `sep_start` per codec."""
    return sum(1 for spec in layer_table(cfg) if spec.separable)


def tf_variable_names(cfg: CodecConfig, scope: str, sep_start: int = 0) -> List[Tuple[str, ...]]:
    """Names of one codec's conv variables in creation order: (kernel, bias) or (depthwise, pointwise, bias) per layer.
    sep_start: separable convs created EARLIER IN THE SAME GRAPH (Keras names are per graph): 0 for the first codec built,
    separable_count(cfg_1) for the second, ..."""
    names, n_conv, n_sep = [], 0, int(sep_start)
    for spec in layer_table(cfg):
        if spec.separable:
            base = f"{scope}/separable_conv1d" + (f"_{n_sep}" if n_sep else "")
            names.append((base + "/depthwise_kernel", base + "/pointwise_kernel", base + "/bias"))
            n_sep += 1
        else:
            base = f"{scope}/conv1d" + (f"_{n_conv}" if n_conv else "")
            names.append((base + "/kernel", base + "/bias"))
            n_conv += 1
    return names


def _get(variables: Mapping[str, np.ndarray], name: str) -> np.ndarray:
    for key in (name, name + ":0"):
        if key in variables:
            return np.asarray(variables[key])
    raise KeyError(f"variable {name!r} is not in the checkpoint dump")


def _resolve_separable(variables: Mapping[str, np.ndarray], scope: str, nth: int) -> str:
    """Base name of the nth (0-based, ascending suffix) separable conv that exists inside ``scope`` in the dump."""
    import re
    pat = re.compile(re.escape(scope) + r"/separable_conv1d(?:_(\d+))?/depthwise_kernel(?::0)?$")
    found = sorted((int(m.group(1) or 0) for m in (pat.match(k) for k in variables) if m))
    if nth >= len(found):
        raise KeyError(f"separable conv #{nth} of scope {scope!r} is not in the checkpoint dump (found suffixes {found})")
    return f"{scope}/separable_conv1d" + (f"_{found[nth]}" if found[nth] else "")


def params_from_tf_variables(cfg: CodecConfig, scope: str, variables: Mapping[str, np.ndarray], sep_start: int = 0) -> np.ndarray:
    """Flat float32 parameter image (the layout of nsc_codec_layer_info) of the codec that lives in ``scope`` ('scope_1', ...).
    Separable layers are looked up under the name `sep_start` predicts and, failing that, by ascending suffix inside the scope."""
    conv = []
    nth_sep = 0
    for spec, names in zip(layer_table(cfg), tf_variable_names(cfg, scope, sep_start)):
        if spec.separable:
            if not any(k in variables for k in (names[0], names[0] + ":0")):
                base = _resolve_separable(variables, scope, nth_sep)
                names = (base + "/depthwise_kernel", base + "/pointwise_kernel", base + "/bias")
            nth_sep += 1
        arrs = [_get(variables, n).astype(np.float32) for n in names]
        want = ([(spec.k, spec.cin, 1), (1, spec.cin, spec.cout), (spec.cout,)] if spec.separable
                else [(spec.k, spec.cin, spec.cout), (spec.cout,)])
        for n, a, w in zip(names, arrs, want):
            # Keras stores the depthwise kernel as (k, cin, 1) and conv kernels as (k, cin, cout) [LIB]
            if tuple(a.shape) != tuple(w):
                raise ValueError(f"{n}: shape {tuple(a.shape)} does not match the configured layer {tuple(w)}")
        conv.append(tuple(arrs))
    alpha = float(np.asarray(_get(variables, f"{scope}/alpha")).reshape(-1)[0])
    bins = _get(variables, f"{scope}/bins").astype(np.float32).reshape(-1)
    if bins.size != cfg.num_bins:
        raise ValueError(f"{scope}/bins has {bins.size} entries, the config says {cfg.num_bins}")
    return pack_params_numpy(cfg, conv, alpha, bins)


def lsf_params_from_tf_variables(variables: Mapping[str, np.ndarray], scope: str = 'lpc_quan') -> np.ndarray:
    """{alpha, bins[...]} of the LSF codebook (cmrl.py:782-788) as the float32 vector nsc_cq_forward takes."""
    alpha = float(np.asarray(_get(variables, f"{scope}/alpha")).reshape(-1)[0])
    bins = _get(variables, f"{scope}/bins").astype(np.float32).reshape(-1)
    return np.concatenate([[np.float32(alpha)], bins]).astype(np.float32)


def tf_variables_from_params(cfg: CodecConfig, scope: str, params: np.ndarray, sep_start: int = 0) -> Dict[str, np.ndarray]:
    """Inverse of params_from_tf_variables (export towards a TensorFlow `assign`)."""
    out, p = {}, np.asarray(params, dtype=np.float32)
    for spec, names in zip(layer_table(cfg), tf_variable_names(cfg, scope, sep_start)):
        o = spec.offset
        shapes = ([(spec.k, spec.cin, 1), (1, spec.cin, spec.cout), (spec.cout,)] if spec.separable
                  else [(spec.k, spec.cin, spec.cout), (spec.cout,)])
        for n, sh in zip(names, shapes):
            size = int(np.prod(sh))
            out[n] = p[o:o + size].reshape(sh).copy()
            o += size
    tail = p[-(1 + cfg.num_bins):]
    out[f"{scope}/alpha"] = np.float32(tail[0])
    out[f"{scope}/bins"] = tail[1:].copy()
    return out


class EntropyController:
    """The tau schedule of the training loops: after every epoch tau moves by +-0.015 towards the target entropy
    (nscm.py:630-639: +0.015 above target + 0.05, -0.045 below target, only while quantisation is on; the finetuning loop
    nscm.py:494-518 uses symmetric +-0.015 per codec)."""

    def __init__(self, target_entropy: float, tau: float = 0.0, ent_change: float = 0.015):
        self.target, self.tau, self.ent_change = float(target_entropy), float(tau), float(ent_change)

    def update(self, entropy: float, is_quan_on: float = 1.0) -> float:
        if is_quan_on == 1.0:
            if entropy > self.target + 0.05:
                self.tau += self.ent_change
            elif entropy < self.target:
                self.tau -= self.ent_change * 3
        return self.tau

    def update_finetune(self, entropy: float) -> float:
        if entropy > self.target:
            self.tau += self.ent_change
        elif entropy < self.target:
            self.tau -= self.ent_change
        return self.tau


def schedule(epoch: int, pretrain_step: int, epochs: int = 0, is_cq: bool = True, update_every: int = 30) -> Dict[str, object]:
    """What epoch ``epoch`` of model_training_lpc does (nscm.py:560-598).
      * the first ``pretrain_step`` epochs run the op that minimises loss_no_quan = c0 time + c1 freq with is_quan_on = 0, the rest
        the op that minimises loss_quan (:560-568) -- two AdamOptimizer instances with SEPARATE slots (nscm.py:1055-1059);
      * a CQ run recomputes all LPC residuals with the current (sorted) LSF codebook when ``i % 30 == 0 and i != 0`` or
        ``i == epoch - 3`` -- regardless of the pre-training phase -- and that epoch SKIPS its training pass (:578-584)."""
    pre = epoch < pretrain_step
    update = bool(is_cq) and ((epoch % update_every == 0 and epoch != 0) or (epochs > 0 and epoch == epochs - 3))
    return {'loss': 'loss_no_quan' if pre else 'loss_quan', 'is_quan_on': 0.0 if pre else 1.0, 'optimizer': 'no_quan' if pre else 'quan',
            'update_lpc_residual': update, 'skip_training': update}


def sorted_lsf_bins(lsf_params: np.ndarray) -> np.ndarray:
    """_update_lpc_residual re-reads the learned LSF bins SORTED (nscm.py:1084-1087); the whole step -- sorted bins, fresh alpha,
    hard assignment, lsf2poly, residual over the training set -- is nsc_b200.training.update_lpc_residual."""
    p = np.asarray(lsf_params, dtype=np.float32).copy()
    p[1:] = np.sort(p[1:])
    return p
