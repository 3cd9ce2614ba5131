"""ctypes binding of libnsc_b200.so (the C ABI declared in include/nsc_b200.h).

There is NO fallback: if the shared library is missing or a symbol cannot be bound, importing the
operator modules raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C nsc_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnsc_b200.so")

MAX_BLOCKS, MAX_STRIDES, MAX_CODECS = 8, 4, 8
FRAME_LENGTH, LPC_ORDER = 512, 16
MEL_BINS, MEL_TOTAL = 257, 184
MEL_BUFFER_FLOATS = MEL_BINS * MEL_TOTAL + 2 * MEL_TOTAL
ACT_NONE, ACT_TANH, ACT_LRELU = 0, 1, 2


class CodecCfgStruct(C.Structure):
    """struct nsc_codec_cfg (include/nsc_b200.h)."""
    _fields_ = [
        ("k_dilated", C.c_int32), ("k_plain", C.c_int32), ("wide", C.c_int32), ("narrow", C.c_int32),
        ("n_blocks", C.c_int32), ("dilations", C.c_int32 * MAX_BLOCKS),
        ("n_strides", C.c_int32), ("strides", C.c_int32 * MAX_STRIDES),
        ("resnet_type", C.c_int32), ("num_bins", C.c_int32), ("precision", C.c_int32),
    ]


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_cfgp = C.POINTER(CodecCfgStruct)
_ppv = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/nsc_b200.h declares
SIGNATURES = {
    "nsc_version": (_i32, []),
    "nsc_last_error": (C.c_char_p, []),
    "nsc_launch_count": (C.c_longlong, []),
    "nsc_profile_begin": (_i32, [_i32]),
    "nsc_profile_end": (_i32, [C.POINTER(_i32), C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_double), _i32]),
    "nsc_conv1d": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "nsc_conv1d_tc_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "nsc_conv1d_tc_plan_info": (_i32, [_i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, C.POINTER(C.c_int64)]),
    "nsc_narrow_conv_plan_info": (_i32, [_i64, _i32, _i32, C.POINTER(C.c_int64)]),
    "nsc_conv1d_tc": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "nsc_conv1d_depth": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "nsc_block_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32]),
    "nsc_bottleneck_block": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "nsc_bottleneck_block_tc_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32, _i32]),
    "nsc_bottleneck_block_tc": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, C.POINTER(_i32), _vp, _i64, _vp]),
    "nsc_gated_block_tc_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32, _i32, _i32]),
    "nsc_gated_block_tc": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "nsc_debug_block_stats": (_i32, [C.POINTER(C.c_ulonglong), _i32]),
    "nsc_quantize_scalar": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nsc_dequantize_scalar": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp]),
    "nsc_quan_loss": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "nsc_soft_histogram": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "nsc_entropy_from_hist": (_i32, [_vp, _i32, _vp, _vp]),
    "nsc_lpc_analyze": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "nsc_lpc_analyze_train": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "nsc_lsf2poly": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "nsc_lpc_residual": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "nsc_lpc_synth": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "nsc_mel_filterbank": (_i32, [_vp, _vp]),
    "nsc_losses_forward": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "nsc_codec_param_count": (_i64, [_cfgp]),
    "nsc_codec_layer_info": (_i32, [_cfgp, _i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64)]),
    "nsc_codec_workspace_bytes": (_i64, [_cfgp, _i64]),
    "nsc_codec_on_plane_engine": (_i32, [_cfgp]),
    "nsc_codec_forward": (_i32, [_cfgp, _vp, _vp, _i64, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "nsc_codec_encode": (_i32, [_cfgp, _vp, _vp, _i64, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "nsc_codec_decode": (_i32, [_cfgp, _vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "nsc_prepare": (_i32, [_i32, _cfgp, _i32, _ppv, _i64, _vp, _i64, _vp]),
    "nsc_release": (_i32, [_vp]),
    "nsc_cascade_workspace_bytes": (_i64, [_cfgp, _i32, _i64]),
    "nsc_cascade_forward": (_i32, [_cfgp, _i32, _ppv, _vp, _i64, _f32, _i32, _f32, _i32, _ppv, _ppv, _ppv, _ppv, _vp, _vp, _i64, _vp]),
    "nsc_cq_workspace_bytes": (_i64, [_cfgp, _i32, _i64]),
    "nsc_pass_frames": (_i64, [_cfgp, _i32]),
    "nsc_cq_forward": (_i32, [_cfgp, _i32, _ppv, _vp, _i32, _vp, _vp, _i64, _f32, _f32, _i32, _vp, _vp, _vp, _ppv, _ppv, _ppv,
                              _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "nsc_train_workspace_bytes": (_i64, [_cfgp, _i32, _i64]),
    "nsc_train_forward": (_i32, [_cfgp, _i32, _ppv, _vp, _i32, _vp, _vp, _i64, _f32, _f32, _vp, _vp, _vp, _vp, _ppv, _ppv, _vp, _i64, _vp]),
    "nsc_train_backward": (_i32, [_cfgp, _i32, _ppv, _vp, _i32, _vp, _vp, _i64, _f32, _f32, _vp, _vp, C.POINTER(_f32), C.POINTER(_f32),
                                  C.POINTER(_f32), _i64, _ppv, C.POINTER(_i32), _ppv, _vp, _vp, _i64, _vp]),
    "nsc_segment_count": (_i64, [_i64]),
    "nsc_utterance_to_segment": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp]),
    "nsc_lpc_window_count": (_i64, [_i64]),
    "nsc_lpc_windows": (_i32, [_vp, _i64, _vp, _vp]),
    "nsc_overlap_add": (_i32, [_vp, _i64, _i64, _vp, _i64, _vp]),
    "nsc_utterances_to_segments": (_i32, [_vp, _i64, _i64, _i64, _i32, _i64, _vp, _vp]),
    "nsc_lpc_windows_batch": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "nsc_overlap_add_batch": (_i32, [_vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "nsc_iir_workspace_bytes": (_i64, [_i64, _i64]),
    "nsc_iir_biquad": (_i32, [_vp, _i64, _i64, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _vp, _vp, _i64, _vp]),
    "nsc_packed_row_bytes": (_i32, [_i32, _i32]),
    "nsc_pack_codes": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "nsc_unpack_codes": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "nsc_adam_step": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _i64, _f32, _f32, _f32, _vp]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Loads the library once and binds every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library is not built. nsc_b200 has no CPU fallback; "
            "run `make -C nsc_b200/csrc` (or __graft_entry__.build()).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover
            raise RuntimeError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().nsc_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    """Maps the C return code to the Python exception the reference's own code would raise."""
    if rc == 0:
        return
    msg = f"{what}: {last_error()} (rc={rc})"
    if rc == -1:
        raise ValueError(msg)
    raise RuntimeError(msg)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError("nsc_b200 operates on CUDA tensors only (no CPU path)")
    if not t.is_contiguous():
        raise ValueError("nsc_b200 needs contiguous tensors")
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr_array(ts: Optional[Sequence[Optional[torch.Tensor]]]):
    """Host array of device pointers (NULL entries allowed); None -> NULL array."""
    if ts is None:
        return None
    arr = (C.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = ptr(t)
    return arr


def require_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()
