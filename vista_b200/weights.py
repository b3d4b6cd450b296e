"""Weight repacking: reference ``state_dict`` tensors (OIHW conv / (out,in) linear, fp32) ->
the layouts the kernels consume (fp16 [N, taps*Cin] with tap-major K; GEGLU tile interleave)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def conv_weight_to_taps(w: torch.Tensor) -> torch.Tensor:
    """(O, I, kh, kw) or (O, I, kt, 1, 1) -> (O, taps*I) with K index = tap*I + i (tap = kh*3+kw or kt)."""
    if w.dim() == 5:
        assert w.shape[3] == 1 and w.shape[4] == 1
        w = w[:, :, :, 0, 0].permute(0, 2, 1)          # O, kt, I
    elif w.dim() == 4:
        w = w.permute(0, 2, 3, 1)                      # O, kh, kw, I
    else:
        raise ValueError(w.shape)
    return w.reshape(w.shape[0], -1).contiguous()


def permute_geglu(w: torch.Tensor, b: Optional[torch.Tensor], tile_n: int) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """GEGLU.proj weight (2*inner, C): rows [0, inner) are values, [inner, 2*inner) gates
    (vwm/modules/attention.py:90-92).  Re-order rows so that every tile of ``tile_n`` output columns
    holds tile_n/2 value rows followed by their gate rows (epilogue pairs column j with j + tile_n/2)."""
    two_inner = w.shape[0]
    inner = two_inner // 2
    h = tile_n // 2
    assert inner % h == 0, (inner, tile_n)
    idx = []
    for t in range(inner // h):
        idx.extend(range(t * h, (t + 1) * h))
        idx.extend(range(inner + t * h, inner + (t + 1) * h))
    idx = torch.tensor(idx, device=w.device)
    return w.index_select(0, idx).contiguous(), (None if b is None else b.index_select(0, idx).contiguous())
