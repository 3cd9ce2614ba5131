"""Fixed-width packing of the hard codes.  The reference has no bitstream -- it reports
`entropy_to_bitrate` estimates (loss_terms_and_measures.py:63-67); this gives "encode to hard codes" an actual byte
format: per frame 16 LSF indices (8 bits for the 256-entry codebook) followed by each codec's code indices at
ceil(log2(num_bins)) bits, little-endian bit order inside a row."""
from __future__ import annotations

import math
from typing import Sequence

import torch

from . import _lib

_lib.load()


def bits_for(num_bins: int) -> int:
    return max(1, math.ceil(math.log2(num_bins)))


def pack_codes(idx: torch.Tensor, num_bins: int) -> torch.Tensor:
    """(rows, L) uint8 indices -> (rows, ceil(L * bits / 8)) uint8."""
    if idx.dtype != torch.uint8:
        raise ValueError("indices must be uint8")
    rows, L = idx.shape
    bits = bits_for(num_bins)
    rb = int(_lib.load().nsc_packed_row_bytes(L, bits))
    out = torch.empty((rows, rb), dtype=torch.uint8, device=idx.device)
    _lib.check(_lib.load().nsc_pack_codes(_lib.ptr(idx.contiguous()), rows, L, bits, _lib.ptr(out), _lib.stream_ptr()), 'pack_codes')
    return out


def unpack_codes(packed: torch.Tensor, L: int, num_bins: int) -> torch.Tensor:
    rows = packed.shape[0]
    bits = bits_for(num_bins)
    if packed.shape[1] != int(_lib.load().nsc_packed_row_bytes(L, bits)):
        raise ValueError("packed row length does not match (L, bits)")
    out = torch.empty((rows, L), dtype=torch.uint8, device=packed.device)
    _lib.check(_lib.load().nsc_unpack_codes(_lib.ptr(packed.contiguous()), rows, L, bits, _lib.ptr(out), _lib.stream_ptr()), 'unpack_codes')
    return out


def pack_frames(lsf_idx: torch.Tensor, code_idx: Sequence[torch.Tensor], num_bins: Sequence[int], lsf_bins: int = 256) -> torch.Tensor:
    """One record per frame: packed LSF indices, then every codec's packed codes.  (B, record_bytes) uint8."""
    parts = [pack_codes(lsf_idx, lsf_bins)] + [pack_codes(c, n) for c, n in zip(code_idx, num_bins)]
    return torch.cat(parts, dim=1)


def unpack_frames(records: torch.Tensor, code_len: Sequence[int], num_bins: Sequence[int], lsf_order: int = 16, lsf_bins: int = 256):
    lib = _lib.load()
    off = 0
    n = int(lib.nsc_packed_row_bytes(lsf_order, bits_for(lsf_bins)))
    lsf = unpack_codes(records[:, off:off + n].contiguous(), lsf_order, lsf_bins)
    off += n
    codes = []
    for L, nb in zip(code_len, num_bins):
        n = int(lib.nsc_packed_row_bytes(L, bits_for(nb)))
        codes.append(unpack_codes(records[:, off:off + n].contiguous(), L, nb))
        off += n
    return lsf, codes


def record_bitrate_kbps(code_len: Sequence[int], num_bins: Sequence[int], lsf_order: int = 16, lsf_bins: int = 256,
                        hop: int = 480, sample_rate: int = 16000) -> float:
    """Bitrate of the fixed-width records (an upper bound on what an entropy coder over the same indices needs)."""
    bits = lsf_order * bits_for(lsf_bins) + sum(L * bits_for(n) for L, n in zip(code_len, num_bins))
    return bits * sample_rate / hop / 1000.0
