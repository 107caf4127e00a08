"""AP2POH tail: what ``AP2POH.forward`` does between ``propagate_AP2C_backward`` (F-6) and
the POH it returns (ap2poh.py:104-116): channel-wise symmetric k x k convolution of the real and the imaginary
part (nn.py:35-95), ``amplitude_normalizor`` (per-plane max * 1.01, util.py:53-66), ``angle`` and the
double-phase checkerboard encoding (ap2poh.py:86-95).  The reference runs ~14 full-size element-wise / reduction
kernels; here the field is read twice (max pass, encode pass) and the POH written once.

Inference only (``generatePOH.py`` runs the generator under ``no_grad``): calling it on tensors that require
grad raises, the training step keeps differentiating the reference's torch ops.
"""

from __future__ import annotations

import torch

from . import _cabi_next as N
from ._next_common import lib, ptr, scratch, staged, stream_handle


def symmetric_kernels(conv) -> tuple:
    """(weights [3,k,k], bias [3]) of a reference ``ChannelWiseSymmetricConv`` (nn.py:35-95): each colour's kernel is
    ``params[distance_map]`` (one parameter per squared distance from the centre)."""
    ws, bs = [], []
    for sub in (conv.conv_r, conv.conv_g, conv.conv_b):
        ws.append(sub.params.detach()[sub.distance_map.to(sub.params.device)])
        bs.append(sub.bias.detach().reshape(()))
    return torch.stack(ws), torch.stack(bs)


def ap2poh_tail(complex_field: torch.Tensor, weights: torch.Tensor, bias: torch.Tensor, return_plane_max=False):
    """``[B,3,R,C]`` complex64 -> POH ``[B,3,R,C]`` fp32 (ap2poh.py:107-116).

    ``weights`` ``[3,k,k]`` and ``bias`` ``[3]`` as returned by :func:`symmetric_kernels`."""
    if torch.is_grad_enabled() and (complex_field.requires_grad or weights.requires_grad or bias.requires_grad):
        raise RuntimeError("ap2poh_tail is the inference path (no autograd); wrap the call in torch.no_grad()")
    f = staged(complex_field, torch.complex64)
    if f.dim() != 4 or f.shape[1] != 3:
        raise ValueError(f"expected a [B,3,R,C] complex field, got {tuple(complex_field.shape)}")
    w, b = staged(weights), staged(bias)
    k = int(w.shape[-1])
    if w.shape != (3, k, k) or b.shape != (3,):
        raise ValueError(f"weights {tuple(w.shape)} / bias {tuple(b.shape)}: expected [3,k,k] and [3]")
    B, _, rows, cols = (int(s) for s in f.shape)
    planes = B * 3
    partial = scratch(lib().lhg_next_partial_floats(planes, rows, cols), f.device)
    plane_max = torch.empty(planes, dtype=torch.float32, device=f.device)
    poh = torch.empty(B, 3, rows, cols, dtype=torch.float32, device=f.device)
    N.check(lib().lhg_ap2poh_tail(ptr(f), ptr(w), ptr(b), k, planes, rows, cols, ptr(partial), partial.numel(),
                                  ptr(plane_max), ptr(poh), stream_handle()))
    poh = poh.to(complex_field.device)
    return (poh, plane_max.view(B, 3)) if return_plane_max else poh
