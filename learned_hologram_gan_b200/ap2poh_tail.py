"""AP2POH tail: what ``AP2POH.forward`` does between ``propagate_AP2C_backward`` (F-6) and
the POH it returns (ap2poh.py:104-116): channel-wise symmetric k x k convolution of the real and the imaginary
part (nn.py:35-95), ``amplitude_normalizor`` (per-plane max * 1.01, util.py:53-66), ``angle`` and the
double-phase checkerboard encoding (ap2poh.py:86-95).  The reference runs ~14 full-size element-wise / reduction
kernels; here the field is read twice (max pass, encode pass) and the POH written once.

Differentiable: the training step (``AP2POH.forward`` inside ``watermelon.train``) reaches the same kernels through a
``torch.autograd.Function`` whose backward (``lhg_ap2poh_tail_backward``) returns the gradients of the complex field
(``dL/dre + i dL/dim``), of the three k x k kernels and of the three biases; the gradient of the per-plane maximum
goes to the arg-max pixel, as ``torch.max`` does.
"""

from __future__ import annotations

import torch

from . import _cabi_next as N
from ._next_common import lib, ptr, scratch, staged, stream_handle


def symmetric_kernels(conv, differentiable: bool = False) -> tuple:
    """(weights [3,k,k], bias [3]) of a reference ``ChannelWiseSymmetricConv`` (nn.py:35-95): each colour's kernel is
    ``params[distance_map]`` (one parameter per squared distance from the centre).  ``differentiable=True`` keeps the
    graph to ``params`` / ``bias`` (training); the default detaches (inference)."""
    ws, bs = [], []
    for sub in (conv.conv_r, conv.conv_g, conv.conv_b):
        params, bias = (sub.params, sub.bias) if differentiable else (sub.params.detach(), sub.bias.detach())
        ws.append(params[sub.distance_map.to(params.device)])
        bs.append(bias.reshape(()))
    return torch.stack(ws), torch.stack(bs)


def _check(f, w, b):
    if f.dim() != 4 or f.shape[1] != 3:
        raise ValueError(f"expected a [B,3,R,C] complex field, got {tuple(f.shape)}")
    k = int(w.shape[-1])
    if w.shape != (3, k, k) or b.shape != (3,):
        raise ValueError(f"weights {tuple(w.shape)} / bias {tuple(b.shape)}: expected [3,k,k] and [3]")
    return k


def _forward(f, w, b, k):
    B, _, rows, cols = (int(s) for s in f.shape)
    planes = B * 3
    partial = scratch(lib().lhg_next_partial_floats(planes, rows, cols), f.device)
    plane_max = torch.empty(planes, dtype=torch.float32, device=f.device)
    poh = torch.empty(B, 3, rows, cols, dtype=torch.float32, device=f.device)
    N.check(lib().lhg_ap2poh_tail(ptr(f), ptr(w), ptr(b), k, planes, rows, cols, ptr(partial), partial.numel(),
                                  ptr(plane_max), ptr(poh), stream_handle()))
    return poh, plane_max.view(B, 3)


class _Tail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, field, weights, bias):
        f, w, b = staged(field, torch.complex64), staged(weights), staged(bias)
        k = _check(f, w, b)
        poh, _ = _forward(f, w, b, k)
        ctx.save_for_backward(f, w, b)
        ctx.k, ctx.devices = k, (field.device, weights.device, bias.device)
        return poh.to(field.device)

    @staticmethod
    def backward(ctx, g):
        f, w, b = ctx.saved_tensors
        k = ctx.k
        B, _, rows, cols = (int(s) for s in f.shape)
        planes = B * 3
        g_d = staged(g)
        need = lib().lhg_ap2poh_tail_backward_floats(k, planes, rows, cols)
        work = scratch(need, f.device)
        gf = torch.empty_like(f)
        gw, gb = torch.empty_like(w), torch.empty_like(b)
        N.check(lib().lhg_ap2poh_tail_backward(ptr(f), ptr(w), ptr(b), k, ptr(g_d), planes, rows, cols, ptr(work),
                                               work.numel(), ptr(gf), ptr(gw), ptr(gb), stream_handle()))
        return gf.to(ctx.devices[0]), gw.to(ctx.devices[1]), gb.to(ctx.devices[2])


def ap2poh_tail(complex_field: torch.Tensor, weights: torch.Tensor, bias: torch.Tensor, return_plane_max=False):
    """``[B,3,R,C]`` complex64 -> POH ``[B,3,R,C]`` fp32 (ap2poh.py:107-116).

    ``weights`` ``[3,k,k]`` and ``bias`` ``[3]`` as returned by :func:`symmetric_kernels`.  Differentiable with
    respect to all three when they require grad."""
    needs_grad = torch.is_grad_enabled() and (complex_field.requires_grad or weights.requires_grad or bias.requires_grad)
    if needs_grad:
        if return_plane_max:
            raise ValueError("return_plane_max is an inference-only diagnostic")
        return _Tail.apply(complex_field, weights, bias)
    f, w, b = staged(complex_field, torch.complex64), staged(weights), staged(bias)
    k = _check(f, w, b)
    poh, plane_max = _forward(f, w, b, k)
    poh = poh.to(complex_field.device)
    return (poh, plane_max) if return_plane_max else poh
