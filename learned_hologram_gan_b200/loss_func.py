"""Loss epilogues that consume the propagation path's outputs, B200-native.

Mirror of the reference's ``learnedMethodForHologram/watermelon_hologram/loss_func.py`` for the functions of
``watermelon.G_loss`` (watermelon.py:418-445): same names, argument order and values, but each loss is ONE
streaming pass over its inputs (the reference re-reads every plane 5-10 times) and ONE pass for its gradient,
through the C ABI of ``include/lhg_next_b200.h``.  No CPU or torch-op fallback.
"""

from __future__ import annotations

import torch

from . import _cabi_next as N
from ._next_common import lib, partial_for, planes_of, ptr, staged, stream_handle


class _AmpTerms(torch.autograd.Function):
    """terms = [mse, TV(hat), TV(target), |TV(hat)-TV(target)|, mse + alpha*|...|] from one pass (loss.py:66-103)."""

    @staticmethod
    def forward(ctx, hat, target, alpha):
        hat_d = staged(hat)
        tgt_d = None if target is None else staged(target)
        if tgt_d is not None and tgt_d.shape != hat_d.shape:
            raise RuntimeError(f"The size of tensor a {tuple(hat_d.shape)} must match the size of tensor b "
                               f"{tuple(tgt_d.shape)}")
        planes, rows, cols = planes_of(hat_d)
        dev = hat_d.device
        partial = partial_for(planes, rows, cols, dev)
        terms = torch.empty(5, dtype=torch.float32, device=dev)
        N.check(lib().lhg_amp_loss_terms(ptr(hat_d), ptr(tgt_d), planes, rows, cols, float(alpha), ptr(partial),
                                         partial.numel(), ptr(terms), stream_handle()))
        ctx.save_for_backward(hat_d, tgt_d, terms)
        ctx.alpha, ctx.in_device, ctx.shape = float(alpha), hat.device, (planes, rows, cols)
        return terms.to(hat.device)

    @staticmethod
    def backward(ctx, g):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("loss_func: gradients flow to the estimate only (the target is data)")
        hat_d, tgt_d, terms = ctx.saved_tensors
        g = g.to(device=hat_d.device, dtype=torch.float32).contiguous()
        planes, rows, cols = ctx.shape
        grad = torch.empty_like(hat_d)
        N.check(lib().lhg_amp_loss_backward(ptr(hat_d), ptr(tgt_d), ptr(g), ptr(terms), ctx.alpha, planes, rows, cols,
                                            ptr(grad), stream_handle()))
        return grad.to(ctx.in_device), None, None


def amp_loss_terms(amp_hat, amp, alpha=1.0):
    """All five terms of one pass: ``[mse, TV(hat), TV(target), total_variation_loss, amp_loss]`` (a [5] tensor).
    ``watermelon.G_loss`` takes pixel_loss = terms[0] and TV_loss = terms[3] from the same pass."""
    return _AmpTerms.apply(amp_hat, amp, alpha)


def mse_loss(hat, target):
    """``F.mse_loss(hat, target)`` (watermelon.py:436)."""
    return _AmpTerms.apply(hat, target, 0.0)[0]


def total_variation(tensor):
    """loss.py:66-77: mean|x[..., 1:] - x[..., :-1]| + mean|x[..., 1:, :] - x[..., :-1, :]|."""
    return _AmpTerms.apply(tensor, None, 0.0)[1]


def total_variation_loss(y_hat, y):
    """loss.py:92-96."""
    return _AmpTerms.apply(y_hat, y, 0.0)[3]


def amp_loss(amp_hat, amp, alpha=1.0):
    """loss.py:99-103: mse + alpha * total_variation_loss."""
    return _AmpTerms.apply(amp_hat, amp, alpha)[4]


class _FocalPhase(torch.autograd.Function):
    """``weighted=True``: focal_sincos_phase_gradient_loss; ``False``: phase_sincos_gradient_loss (same pass, terms[3])."""

    @staticmethod
    def forward(ctx, fake_phase, real_phase, weighted=True):
        fake_d, real_d = staged(fake_phase), staged(real_phase)
        if fake_d.shape != real_d.shape:
            raise RuntimeError(f"The size of tensor a {tuple(fake_d.shape)} must match the size of tensor b "
                               f"{tuple(real_d.shape)}")
        planes, rows, cols = planes_of(fake_d)
        dev = fake_d.device
        partial = partial_for(planes, rows, cols, dev)
        terms = torch.empty(4, dtype=torch.float32, device=dev)
        N.check(lib().lhg_focal_phase_loss_terms(ptr(fake_d), ptr(real_d), planes, rows, cols, ptr(partial),
                                                 partial.numel(), ptr(terms), stream_handle()))
        ctx.save_for_backward(fake_d, real_d, terms)
        ctx.in_device, ctx.shape, ctx.weighted = fake_phase.device, (planes, rows, cols), bool(weighted)
        return terms[2 if weighted else 3].clone().to(fake_phase.device)

    @staticmethod
    def backward(ctx, g):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("loss_func: gradients flow to the estimate only (the target is data)")
        fake_d, real_d, terms = ctx.saved_tensors
        g1 = g.to(device=fake_d.device, dtype=torch.float32).reshape(1).contiguous()
        planes, rows, cols = ctx.shape
        grad = torch.empty_like(fake_d)
        if ctx.weighted:
            N.check(lib().lhg_focal_phase_loss_backward(ptr(fake_d), ptr(real_d), ptr(terms), ptr(g1), planes, rows,
                                                        cols, ptr(grad), stream_handle()))
        else:
            N.check(lib().lhg_phase_gradient_loss_backward(ptr(fake_d), ptr(real_d), ptr(g1), planes, rows, cols,
                                                           ptr(grad), stream_handle()))
        return grad.to(ctx.in_device), None, None


def focal_sincos_phase_gradient_loss(fake_phase, real_phase):
    """loss.py:135-163.  The focal weights are constants of the graph (``torch.no_grad``), so
    ``mean(d * d/max d) = sum d^2 / (max d * count)``: sum and max come out of the same pass, and the gradient
    is a 5-point stencil of ``sin/cos(fake) - sin/cos(real)``."""
    return _FocalPhase.apply(fake_phase, real_phase, True)


def phase_sincos_gradient_loss(fake_phase, real_phase):
    """loss.py:165-183 (the un-weighted variant, ``watermelon.py:921``): mean d1 + mean d2 of the same differences."""
    return _FocalPhase.apply(fake_phase, real_phase, False)


class _PhasePoint(torch.autograd.Function):
    """``focal=True``: focal_sincos_phase_loss; ``False``: plain_phase_loss (same pass, terms[2])."""

    @staticmethod
    def forward(ctx, fake_phase, real_phase, focal=True):
        fake_d, real_d = staged(fake_phase), staged(real_phase)
        if fake_d.shape != real_d.shape:
            raise RuntimeError(f"The size of tensor a {tuple(fake_d.shape)} must match the size of tensor b "
                               f"{tuple(real_d.shape)}")
        planes, rows, cols = planes_of(fake_d)
        dev = fake_d.device
        partial = partial_for(planes, rows, cols, dev)
        terms = torch.empty(3, dtype=torch.float32, device=dev)
        N.check(lib().lhg_phase_point_loss_terms(ptr(fake_d), ptr(real_d), planes, rows, cols, ptr(partial),
                                                 partial.numel(), ptr(terms), stream_handle()))
        ctx.save_for_backward(fake_d, real_d, terms)
        ctx.in_device, ctx.shape, ctx.focal = fake_phase.device, (planes, rows, cols), bool(focal)
        return terms[1 if focal else 2].clone().to(fake_phase.device)

    @staticmethod
    def backward(ctx, g):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("loss_func: gradients flow to the estimate only (the target is data)")
        fake_d, real_d, terms = ctx.saved_tensors
        g1 = g.to(device=fake_d.device, dtype=torch.float32).reshape(1).contiguous()
        planes, rows, cols = ctx.shape
        grad = torch.empty_like(fake_d)
        N.check(lib().lhg_phase_point_loss_backward(ptr(fake_d), ptr(real_d), ptr(terms), ptr(g1), int(ctx.focal),
                                                    planes, rows, cols, ptr(grad), stream_handle()))
        return grad.to(ctx.in_device), None, None


def focal_sincos_phase_loss(fake_phase, real_phase):
    """loss.py:186-203: ``mean(d * d / max d)`` over d = |sin f - sin r|, |cos f - cos r| -- sum d^2 and max d come
    out of one pass; the gradient is ``(u cos f - v sin f) / (max d * count)`` (the focal weight is detached)."""
    return _PhasePoint.apply(fake_phase, real_phase, True)


def plain_phase_loss(fake_phase, real_phase):
    """loss.py:206-208: ``mean |fake - real|``."""
    return _PhasePoint.apply(fake_phase, real_phase, False)
