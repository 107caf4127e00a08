"""Host-side engine: a plan per geometry, raw launches of the C ABI, and the autograd
Functions whose backward is the adjoint propagation.

PyTorch is used here for device memory, streams and autograd plumbing only; every FFT,
transfer-function evaluation, pad/crop and epilogue runs in libasm_b200.so.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

import torch

from . import _cabi as A

_WORKSPACE_CAP = int(os.environ.get("LHG_WORKSPACE_MB", "2048")) * (1 << 20)
_FUSED_STEP = os.environ.get("LHG_FUSED_STEP", "1") != "0"  # 0: forward and adjoint of the L2 step as two calls
LOSS_PARTIALS = 1024


def compute_device() -> torch.device:
    """The device every propagation runs on: the CURRENT CUDA device (one process per GPU)."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            "learned_hologram_gan_b200 needs a CUDA device: the propagation path has no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def host_radial_grid(prow: int, pcol: int) -> torch.Tensor:
    """sqrt(u^2+v^2)*min(Rp,Cp) on the un-shifted unit-less fftfreq grid, host fp32 (util.py:232-234)"""
    u = torch.fft.fftfreq(prow).unsqueeze(-1)
    v = torch.fft.fftfreq(pcol).unsqueeze(0)
    return torch.sqrt(u**2 + v**2) * min(prow, pcol)


def host_w_grid(prow: int, pcol: int, pitch: float, wavelengths: torch.Tensor) -> torch.Tensor:
    """w = sqrt(max(1/lambda^2 - fx^2 - fy^2, 0)) with the reference's host ops (asm.py:56-57, :163-171)."""
    fr = torch.fft.fftfreq(prow, pitch)
    fc = torch.fft.fftfreq(pcol, pitch)
    sq = fr.unsqueeze(1) ** 2 + fc.unsqueeze(0) ** 2
    inv_l2 = (1 / wavelengths**2).unsqueeze(1).unsqueeze(2)
    return torch.sqrt(torch.clamp(inv_l2 - sq.unsqueeze(0), min=0))


def host_wm_grid(prow, pcol, pitch, wavelengths, mask_radius) -> torch.Tensor:
    """The grid the kernels read: |wm| = w_grid, sign bit = outside the circular mask.

    Built on the HOST with the very torch ops the reference constructor runs (asm.py:60-63):
    torch's CPU sqrt goes through MKL VML, which is not correctly rounded for ~1 % of the bins,
    and one ulp of w is up to 1.6e-3 rad of transfer-function phase -- an IEEE-exact device
    builder is 2e-5 away from the reference, above the 1e-5 parity gate.  One-off, per geometry."""
    w = host_w_grid(prow, pcol, pitch, wavelengths)
    outside = host_radial_grid(prow, pcol) > mask_radius
    sign = torch.where(outside, -torch.ones((), dtype=torch.float32), torch.ones((), dtype=torch.float32))
    return torch.copysign(w, sign.unsqueeze(0)).contiguous()


class Plan:
    """Opaque asm_plan + geometry.  Immutable after creation; safe to share across threads."""

    def __init__(self, rows, cols, pad_rows, pad_cols, pitch, wavelengths, mask_radius, device=None):
        self.lib = A.load()
        self.device = compute_device() if device is None else torch.device(device)
        self.rows, self.cols = int(rows), int(cols)
        self.pad_rows, self.pad_cols = int(pad_rows), int(pad_cols)
        self.prow = self.rows + 2 * self.pad_rows
        self.pcol = self.cols + 2 * self.pad_cols
        wl = torch.as_tensor(wavelengths, dtype=torch.float32).detach().cpu().contiguous()
        self.n_colour = int(wl.numel())
        arr = (C.c_float * self.n_colour)(*wl.tolist())
        handle = C.c_void_p()
        A.check(
            self.lib.asm_plan_create(
                C.byref(handle), int(self.device.index or 0), self.rows, self.cols, self.pad_rows,
                self.pad_cols, float(pitch), arr, self.n_colour, float(mask_radius),
            )
        )
        self.handle = handle
        self.inv_n = 1.0 / float(self.prow * self.pcol)
        self.pitch, self.wavelengths, self.mask_radius = float(pitch), wl, float(mask_radius)
        self.wm = None
        if os.environ.get("LHG_DEVICE_GRIDS", "0") != "1":
            self.wm = host_wm_grid(self.prow, self.pcol, self.pitch, wl, self.mask_radius).to(self.device)
        self.fused_step = False
        # the same grid in the tile order of the compile-time planned column kernel (0 bytes: no such kernel)
        self.wm_tiled = None
        nbytes = int(self.lib.asm_wm_tiled_bytes(self.handle))
        if nbytes > 0:
            self.wm_tiled = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                stream = torch.cuda.current_stream(self.device).cuda_stream
                A.check(self.lib.asm_build_wm_tiled(self.handle, _ptr(self.wm), _ptr(self.wm_tiled),
                                                    C.c_void_p(stream)))

        self.fused_step = self.wm_tiled is not None and bool(self.lib.asm_fused_step_supported(self.handle))

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                self.lib.asm_plan_destroy(h)
            except Exception:
                pass
            self.handle = None

    # ---- attribute grids -----------------------------------------------------------------
    def build_grid(self, kind, z: Optional[torch.Tensor] = None, flags: int = 0) -> torch.Tensor:
        dev = self.device
        nz = 0 if z is None else int(z.numel())
        if kind == A.GRID_W:
            out = torch.empty((self.n_colour, self.prow, self.pcol), dtype=torch.float32, device=dev)
        elif kind in (A.GRID_CIRC_MASK, A.GRID_RADIAL):
            out = torch.empty((self.prow, self.pcol), dtype=torch.float32, device=dev)
        elif kind == A.GRID_H:
            out = torch.empty((nz, self.n_colour, self.prow, self.pcol), dtype=torch.complex64, device=dev)
        elif kind == A.GRID_BAND_LIMIT:
            out = torch.empty((nz, self.n_colour, self.prow, self.pcol), dtype=torch.uint8, device=dev)
        else:
            raise ValueError(kind)
        if z is not None:
            z = z.to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            A.check(self.lib.asm_build_grid(self.handle, kind, _ptr(self.wm), _ptr(z), nz, flags, _ptr(out),
                                            C.c_void_p(stream)))
        return out

    # ---- one pass of the hot path ---------------------------------------------------------
    def run(self, *, n_samples, n_depth=1, reduce_depth=False, in_kind, in0=None, in1=None,
            cot_abs=None, cot_angle=None, cot_abs2=None, cot_target=None, cot_scale=0.0,
            phase_scale=1.0, filter_kind=A.FILTER_NONE, filter_flags=0, z=None, depth_index=None,
            out_kind, out0, out1=None, save_field=None, aux_phase=None, aux_amp=None,
            out_scale=1.0, loss_target=None, loss_partial=None, adj_grad_phase=None, adj_cot_scale=0.0,
            loss_target_u8=False):
        io = A.AsmIO()
        io.struct_bytes = C.sizeof(A.AsmIO)
        io.n_samples, io.n_depth, io.reduce_depth = int(n_samples), int(n_depth), int(bool(reduce_depth))
        io.in_kind, io.filter_kind, io.filter_flags, io.out_kind = in_kind, filter_kind, filter_flags, out_kind
        io.in0, io.in1 = _ptr(in0), _ptr(in1)
        io.cot_abs, io.cot_angle, io.cot_abs2, io.cot_target = (
            _ptr(cot_abs), _ptr(cot_angle), _ptr(cot_abs2), _ptr(cot_target))
        io.cot_scale, io.phase_scale = float(cot_scale), float(phase_scale)
        io.wm_grid = _ptr(self.wm)
        io.wm_tiled = _ptr(self.wm_tiled)
        io.z_dev = _ptr(z)
        io.n_z = 0 if z is None else int(z.numel())
        io.depth_index = _ptr(depth_index)
        io.out0, io.out1, io.save_field = _ptr(out0), _ptr(out1), _ptr(save_field)
        io.aux_phase, io.aux_amp = _ptr(aux_phase), _ptr(aux_amp)
        io.out_scale = float(out_scale)
        io.loss_target, io.loss_partial = _ptr(loss_target), _ptr(loss_partial)
        io.loss_partial_len = 0 if loss_partial is None else int(loss_partial.numel())
        io.adj_grad_phase, io.adj_cot_scale = _ptr(adj_grad_phase), float(adj_cot_scale)
        io.loss_target_u8 = int(bool(loss_target_u8))
        need = int(self.lib.asm_workspace_bytes(self.handle, C.byref(io)))
        if need == 0:
            A.check(-1)
        per_sample = (need + max(io.n_samples, 1) - 1) // max(io.n_samples, 1) + 1024
        ws_bytes = min(need, max(per_sample, _WORKSPACE_CAP))
        dev = self.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            ws = _workspace(ws_bytes, dev, stream)
            io.workspace, io.workspace_bytes = _ptr(ws), ws_bytes
            A.check(self.lib.asm_propagate(self.handle, C.byref(io), C.c_void_p(stream.cuda_stream)))
        return io


# Scratch (W1/W2) is kept between calls, one grow-only buffer per (device, stream, host thread): calls on one stream
# are ordered by the stream, and a multi-GB block that goes back to torch's caching allocator after every call gets
# split for smaller tensors, which costs a synchronising cudaMalloc whenever the next call no longer fits it
# (measured: 3-6 ms per step in bench.py).  Not cached while a CUDA graph is being captured (the block would
# belong to the graph's private pool).
_WS_CACHE = {}
_WS_LOCK = threading.Lock()


def _workspace(nbytes: int, dev: torch.device, stream) -> torch.Tensor:
    if torch.cuda.is_current_stream_capturing():
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ws.record_stream(stream)
        return ws
    key = (dev.index, stream.cuda_stream, threading.get_ident())
    with _WS_LOCK:
        ws = _WS_CACHE.get(key)
        if ws is None or ws.numel() < nbytes:
            _WS_CACHE.pop(key, None)
            ws = None  # release the old block before asking for the larger one
            if len(_WS_CACHE) >= 16:  # streams / threads that are gone
                _WS_CACHE.clear()
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _WS_CACHE[key] = ws
    return ws


def finish_loss(plan: "Plan", partial: torch.Tensor, scale: float) -> torch.Tensor:
    """scale * sum(partial) as a device scalar, by the library's fixed-order finishing block (asm_loss_finish)."""
    out = torch.empty((), dtype=torch.float32, device=partial.device)
    with torch.cuda.device(partial.device):
        stream = torch.cuda.current_stream(partial.device).cuda_stream
        A.check(plan.lib.asm_loss_finish(_ptr(partial), int(partial.numel()), float(scale), _ptr(out),
                                         C.c_void_p(stream)))
    return out


def release_workspaces() -> None:
    """Give the cached scratch buffers back to torch's allocator (e.g. before a memory-hungry phase)."""
    with _WS_LOCK:
        _WS_CACHE.clear()


def upload_small(t: torch.Tensor, dtype, dev) -> torch.Tensor:
    """Distances / depth indices to the device without stalling the host: a pageable host-to-device copy returns
    only once the stream has drained up to it, so host tensors go through (cached) pinned memory."""
    t = t.detach()
    if t.device.type == "cuda":
        return t.to(device=dev, dtype=dtype).contiguous()
    return t.to(dtype).contiguous().pin_memory().to(dev, non_blocking=True)


def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.to(device=dev, dtype=torch.float32).contiguous()


def _c64(t: torch.Tensor, dev) -> torch.Tensor:
    return t.to(device=dev, dtype=torch.complex64).contiguous()


class FilterSpec:
    """What multiplies the spectrum between the column transforms."""

    def __init__(self, use_h: bool, conj: bool, mask: bool, z: Optional[torch.Tensor],
                 depth_index: Optional[torch.Tensor] = None):
        self.kind = A.FILTER_H if use_h else A.FILTER_NONE
        self.flags = (A.FLAG_CONJ if conj else 0) | (A.FLAG_CIRC_MASK if mask else 0)
        self.z = z
        self.depth_index = depth_index

    def adjoint_flags(self):
        return self.flags ^ A.FLAG_CONJ


OUT_KINDS = {"abs": A.OUT_ABS, "angle": A.OUT_ANGLE, "abs_angle": A.OUT_ABS_ANGLE,
             "complex": A.OUT_COMPLEX, "abs2": A.OUT_ABS2}


class _FieldToField(torch.autograd.Function):
    """a*exp(i*s*phase) [S,c,R,C] -> D filtered planes per sample -> abs/angle/complex/|.|^2.

    forward  = asm.py:87-92 (and :328-335, :382-383, :511-522)
    backward = the same pipeline with conj(filter), summed over depth in the column pass
    """

    @staticmethod
    def forward(ctx, plan: Plan, filt: FilterSpec, n_depth: int, out: str, phase_scale: float,
                amp: Optional[torch.Tensor], phase: torch.Tensor):
        dev = plan.device
        S = phase.shape[0]
        phase_d = _f32(phase, dev)
        amp_d = None if amp is None else _f32(amp, dev)
        shape = (S * n_depth, plan.n_colour, plan.rows, plan.cols)
        need_grad = any(ctx.needs_input_grad)
        kind = OUT_KINDS[out]
        out0 = torch.empty(shape, dtype=torch.complex64 if out == "complex" else torch.float32, device=dev)
        out1 = torch.empty(shape, dtype=torch.float32, device=dev) if out == "abs_angle" else None
        field = None
        if need_grad and out != "complex":
            field = torch.empty(shape, dtype=torch.complex64, device=dev)
        plan.run(n_samples=S, n_depth=n_depth, in_kind=A.IN_PHASE if amp_d is None else A.IN_AMP_PHASE,
                 in0=amp_d, in1=phase_d, phase_scale=phase_scale, filter_kind=filt.kind,
                 filter_flags=filt.flags, z=filt.z, depth_index=filt.depth_index, out_kind=kind,
                 out0=out0, out1=out1, save_field=field, out_scale=plan.inv_n)
        ctx.plan, ctx.filt, ctx.n_depth, ctx.out, ctx.phase_scale = plan, filt, n_depth, out, phase_scale
        ctx.has_amp = amp is not None
        ctx.amp_needs = amp is not None and ctx.needs_input_grad[5]
        ctx.in_devices = (phase.device, None if amp is None else amp.device)
        ctx.save_for_backward(phase_d, amp_d, field)
        if out == "abs_angle":
            return out0, out1
        return out0

    @staticmethod
    def backward(ctx, *grads):
        plan, filt = ctx.plan, ctx.filt
        dev = plan.device
        phase_d, amp_d, field = ctx.saved_tensors
        S = phase_d.shape[0]
        g_phase = torch.empty_like(phase_d)
        g_amp = torch.empty_like(phase_d) if ctx.amp_needs else None
        kw = dict(n_samples=S, n_depth=ctx.n_depth, reduce_depth=True, filter_kind=filt.kind,
                  filter_flags=filt.adjoint_flags(), z=filt.z, depth_index=filt.depth_index,
                  out_kind=A.OUT_GRAD_PHASE, out0=g_phase, out1=g_amp, aux_phase=phase_d, aux_amp=amp_d,
                  phase_scale=ctx.phase_scale, out_scale=plan.inv_n)
        if ctx.out == "complex":
            plan.run(in_kind=A.IN_COMPLEX, in0=_c64(grads[0], dev), **kw)
        else:
            g0 = None if grads[0] is None else _f32(grads[0], dev)
            g1 = None if len(grads) < 2 or grads[1] is None else _f32(grads[1], dev)
            cot = dict(cot_abs=None, cot_angle=None, cot_abs2=None)
            if ctx.out == "abs":
                cot["cot_abs"] = g0
            elif ctx.out == "angle":
                cot["cot_angle"] = g0
            elif ctx.out == "abs2":
                cot["cot_abs2"] = g0
            else:
                cot["cot_abs"], cot["cot_angle"] = g0, g1
            plan.run(in_kind=A.IN_COTANGENT, in0=field, **cot, **kw)
        gp = g_phase.to(ctx.in_devices[0])
        ga = None if g_amp is None else g_amp.to(ctx.in_devices[1])
        return None, None, None, None, None, ga, gp


class _FieldToSpectrum(torch.autograd.Function):
    """a*exp(i*s*phase) -> fft2(pad(.)) * filter, padded spectrum out (asm.py:386-392, :548-552)."""

    @staticmethod
    def forward(ctx, plan: Plan, filt: FilterSpec, phase_scale: float, amp, phase):
        dev = plan.device
        S = phase.shape[0]
        phase_d = _f32(phase, dev)
        amp_d = None if amp is None else _f32(amp, dev)
        out0 = torch.empty((S, plan.n_colour, plan.prow, plan.pcol), dtype=torch.complex64, device=dev)
        plan.run(n_samples=S, in_kind=A.IN_PHASE if amp_d is None else A.IN_AMP_PHASE, in0=amp_d, in1=phase_d,
                 phase_scale=phase_scale, filter_kind=filt.kind, filter_flags=filt.flags, z=filt.z,
                 depth_index=filt.depth_index, out_kind=A.OUT_SPECTRUM, out0=out0, out_scale=1.0)
        ctx.plan, ctx.filt, ctx.phase_scale = plan, filt, phase_scale
        ctx.amp_needs = amp is not None and ctx.needs_input_grad[3]
        ctx.in_devices = (phase.device, None if amp is None else amp.device)
        ctx.save_for_backward(phase_d, amp_d)
        return out0

    @staticmethod
    def backward(ctx, g):
        plan, filt = ctx.plan, ctx.filt
        dev = plan.device
        phase_d, amp_d = ctx.saved_tensors
        g_phase = torch.empty_like(phase_d)
        g_amp = torch.empty_like(phase_d) if ctx.amp_needs else None
        # (fft2 o pad)^H = crop o (Rp*Cp * ifft2): the un-normalised inverse, scale 1
        plan.run(n_samples=phase_d.shape[0], in_kind=A.IN_SPECTRUM, in0=_c64(g, dev), filter_kind=filt.kind,
                 filter_flags=filt.adjoint_flags(), z=filt.z, depth_index=filt.depth_index,
                 out_kind=A.OUT_GRAD_PHASE, out0=g_phase, out1=g_amp, aux_phase=phase_d, aux_amp=amp_d,
                 phase_scale=ctx.phase_scale, out_scale=1.0)
        gp = g_phase.to(ctx.in_devices[0])
        ga = None if g_amp is None else g_amp.to(ctx.in_devices[1])
        return None, None, None, ga, gp


class _SpectrumToField(torch.autograd.Function):
    """padded spectrum [S,c,Rp,Cp] -> D filtered planes per sample -> crop(ifft2(.)) -> abs, angle
    (asm.py:524-546).  backward sums conj(filter_d) * fft2(pad(cotangent_d)) over depth."""

    @staticmethod
    def forward(ctx, plan: Plan, filt: FilterSpec, n_depth: int, out: str, spec):
        dev = plan.device
        S = spec.shape[0]
        spec_d = _c64(spec, dev)
        shape = (S * n_depth, plan.n_colour, plan.rows, plan.cols)
        kind = OUT_KINDS[out]
        out0 = torch.empty(shape, dtype=torch.complex64 if out == "complex" else torch.float32, device=dev)
        out1 = torch.empty(shape, dtype=torch.float32, device=dev) if out == "abs_angle" else None
        field = None
        if ctx.needs_input_grad[4] and out != "complex":
            field = torch.empty(shape, dtype=torch.complex64, device=dev)
        plan.run(n_samples=S, n_depth=n_depth, in_kind=A.IN_SPECTRUM, in0=spec_d, filter_kind=filt.kind,
                 filter_flags=filt.flags, z=filt.z, depth_index=filt.depth_index, out_kind=kind, out0=out0,
                 out1=out1, save_field=field, out_scale=plan.inv_n)
        ctx.plan, ctx.filt, ctx.n_depth, ctx.out, ctx.S = plan, filt, n_depth, out, S
        ctx.in_device = spec.device
        ctx.save_for_backward(field)
        if out == "abs_angle":
            return out0, out1
        return out0

    @staticmethod
    def backward(ctx, *grads):
        plan, filt = ctx.plan, ctx.filt
        dev = plan.device
        (field,) = ctx.saved_tensors
        g_spec = torch.empty((ctx.S, plan.n_colour, plan.prow, plan.pcol), dtype=torch.complex64, device=dev)
        kw = dict(n_samples=ctx.S, n_depth=ctx.n_depth, reduce_depth=True, filter_kind=filt.kind,
                  filter_flags=filt.adjoint_flags(), z=filt.z, depth_index=filt.depth_index,
                  out_kind=A.OUT_SPECTRUM, out0=g_spec, out_scale=plan.inv_n)
        if ctx.out == "complex":
            plan.run(in_kind=A.IN_COMPLEX, in0=_c64(grads[0], dev), **kw)
        else:
            g0 = None if grads[0] is None else _f32(grads[0], dev)
            g1 = None if len(grads) < 2 or grads[1] is None else _f32(grads[1], dev)
            cot = dict(cot_abs=None, cot_angle=None, cot_abs2=None)
            if ctx.out == "abs":
                cot["cot_abs"] = g0
            elif ctx.out == "angle":
                cot["cot_angle"] = g0
            elif ctx.out == "abs2":
                cot["cot_abs2"] = g0
            else:
                cot["cot_abs"], cot["cot_angle"] = g0, g1
            plan.run(in_kind=A.IN_COTANGENT, in0=field, **cot, **kw)
        return None, None, None, None, g_spec.to(ctx.in_device)


class _AmplitudeMSE(torch.autograd.Function):
    """Fused bench/training loss: mean((|propagate(phase)| - target)^2) over all D planes.

    The squared-error partial sums are reduced inside the last row pass (fixed order, no float
    atomics); backward feeds cot_scale*(|y|-target)*y/|y| straight into the adjoint's first pass,
    so neither the amplitudes nor their gradient are materialised."""

    @staticmethod
    def forward(ctx, plan: Plan, filt: FilterSpec, n_depth: int, amp, phase, target):
        dev = plan.device
        S = phase.shape[0]
        phase_d = _f32(phase, dev)
        amp_d = None if amp is None else _f32(amp, dev)
        target_d = _f32(target, dev)
        shape = (S * n_depth, plan.n_colour, plan.rows, plan.cols)
        if tuple(target_d.shape) != shape:
            raise ValueError(f"target shape {tuple(target_d.shape)} != {shape}")
        amp_hat = torch.empty(shape, dtype=torch.float32, device=dev)
        partial = torch.empty(LOSS_PARTIALS, dtype=torch.float32, device=dev)
        numel = amp_hat.numel()
        ctx.plan, ctx.filt, ctx.n_depth, ctx.numel = plan, filt, n_depth, numel
        ctx.amp_needs = amp is not None and ctx.needs_input_grad[3]
        ctx.in_devices = (phase.device, None if amp is None else amp.device)
        ctx.mark_non_differentiable(amp_hat)
        common = dict(n_samples=S, n_depth=n_depth, in_kind=A.IN_PHASE if amp_d is None else A.IN_AMP_PHASE,
                      in0=amp_d, in1=phase_d, filter_kind=filt.kind, filter_flags=filt.flags, z=filt.z,
                      depth_index=filt.depth_index, out_kind=A.OUT_ABS, out0=amp_hat, out_scale=plan.inv_n,
                      loss_target=target_d, loss_partial=partial)
        ctx.eager_grad = False
        if plan.fused_step and _FUSED_STEP and ctx.needs_input_grad[4] and not ctx.amp_needs:
            # the phase gradient of the un-weighted loss is computed right here by the fused step (one call, no
            # saved field); backward only multiplies it by the upstream scalar
            g_phase = torch.empty_like(phase_d)
            try:
                plan.run(adj_grad_phase=g_phase, adj_cot_scale=2.0 / numel, **common)
                ctx.eager_grad = True
                ctx.save_for_backward(g_phase)
            except A.AsmError as e:
                if e.code != -2:  # ASM_EUNSUPPORTED_SIZE (e.g. unaligned views): the two-call form
                    raise
        if not ctx.eager_grad:
            field = torch.empty(shape, dtype=torch.complex64, device=dev)
            plan.run(save_field=field, **common)
            ctx.save_for_backward(phase_d, amp_d, field, target_d)
        loss = finish_loss(plan, partial, 1.0 / numel)
        return loss.to(phase.device), amp_hat

    @staticmethod
    def backward(ctx, g_loss, _g_amp_hat):
        plan, filt = ctx.plan, ctx.filt
        if ctx.eager_grad:
            (g_phase,) = ctx.saved_tensors
            gp = (g_phase * g_loss.to(device=plan.device, dtype=torch.float32)).to(ctx.in_devices[0])
            return None, None, None, None, gp, None
        phase_d, amp_d, field, target_d = ctx.saved_tensors
        g_phase = torch.empty_like(phase_d)
        g_amp = torch.empty_like(phase_d) if ctx.amp_needs else None
        scale = 2.0 / ctx.numel  # the upstream scalar is applied on the device below (no host sync)
        plan.run(n_samples=phase_d.shape[0], n_depth=ctx.n_depth, reduce_depth=True, in_kind=A.IN_COTANGENT,
                 in0=field, cot_target=target_d, cot_scale=scale, filter_kind=filt.kind,
                 filter_flags=filt.adjoint_flags(), z=filt.z, depth_index=filt.depth_index,
                 out_kind=A.OUT_GRAD_PHASE, out0=g_phase, out1=g_amp, aux_phase=phase_d, aux_amp=amp_d,
                 out_scale=plan.inv_n)
        gl = g_loss.to(device=plan.device, dtype=torch.float32)
        g_phase.mul_(gl)
        if g_amp is not None:
            g_amp.mul_(gl)
        gp = g_phase.to(ctx.in_devices[0])
        ga = None if g_amp is None else g_amp.to(ctx.in_devices[1])
        return None, None, None, ga, gp, None


def amplitude_mse_direct(plan: Plan, filt: FilterSpec, n_depth: int, phase: torch.Tensor, target: torch.Tensor,
                         grad_scale: float, grad_out: Optional[torch.Tensor] = None):
    """Forward + adjoint of the amplitude-L2 workload WITHOUT autograd (training loops that own their
    optimiser step, the sharded focal stack, the bench): two asm_propagate calls.

    Returns (sum over all planes of (|y| - target)^2 as a device scalar,
             grad_scale/2 * d(sum_sq)/d(phase) written into ``grad_out`` (contiguous, like ``phase``)).
    With grad_scale = 2/numel the gradient is that of the mean squared error."""
    dev = plan.device
    if phase.dim() != 4 or tuple(phase.shape[1:]) != (plan.n_colour, plan.rows, plan.cols):
        # the kernels read S*n_colour*R*C floats from the phase and write as many into the gradient
        raise ValueError(f"phase shape {tuple(phase.shape)} != [N, {plan.n_colour}, {plan.rows}, {plan.cols}]")
    S = phase.shape[0]
    phase_d = _f32(phase, dev)
    # 8-bit targets (uint8 samples v, target = v / 255 like the reference's image loader, util.py:44) are converted
    # by the fused row kernel as it reads them; every other path gets the same values as fp32
    u8 = target.dtype == torch.uint8
    target_d = target.to(dev).contiguous() if u8 else _f32(target, dev)
    shape = (S * n_depth, plan.n_colour, plan.rows, plan.cols)
    if tuple(target_d.shape) != shape:
        raise ValueError(f"target shape {tuple(target_d.shape)} != {shape}")
    partial = torch.empty(LOSS_PARTIALS, dtype=torch.float32, device=dev)
    if plan.fused_step and _FUSED_STEP:
        # one call: the row-inverse pass of the forward and the row-forward pass of the adjoint are one kernel,
        # |y| and the saved field are never written (asm_io.adj_grad_phase, include/asm_b200.h)
        g_phase = grad_out if grad_out is not None else torch.empty_like(phase_d)
        if not g_phase.is_contiguous() or g_phase.shape != phase_d.shape:
            raise ValueError("grad_out must be contiguous and shaped like phase")
        try:
            plan.run(n_samples=S, n_depth=n_depth, in_kind=A.IN_PHASE, in1=phase_d, filter_kind=filt.kind,
                     filter_flags=filt.flags, z=filt.z, depth_index=filt.depth_index, out_kind=A.OUT_ABS, out0=None,
                     out_scale=plan.inv_n, loss_target=target_d, loss_partial=partial, adj_grad_phase=g_phase,
                     adj_cot_scale=float(grad_scale), loss_target_u8=u8)
            return finish_loss(plan, partial, 1.0), g_phase
        except A.AsmError as e:  # e.g. a view that is not 16-byte aligned: the two-call form below takes it
            if e.code != -2:  # ASM_EUNSUPPORTED_SIZE
                raise
    if u8:
        # fl(v / 255) as the IEEE quotient (torch's CUDA div by a scalar multiplies by the reciprocal, 1 ulp off for
        # some v): the fp64 quotient rounds to the same fp32 value for all 256 inputs (checked exhaustively)
        target_d = target_d.to(torch.float64).div_(255).to(torch.float32)
    amp_hat = torch.empty(shape, dtype=torch.float32, device=dev)
    field = torch.empty(shape, dtype=torch.complex64, device=dev)
    plan.run(n_samples=S, n_depth=n_depth, in_kind=A.IN_PHASE, in1=phase_d, filter_kind=filt.kind,
             filter_flags=filt.flags, z=filt.z, depth_index=filt.depth_index, out_kind=A.OUT_ABS, out0=amp_hat,
             save_field=field, out_scale=plan.inv_n, loss_target=target_d, loss_partial=partial)
    g_phase = grad_out if grad_out is not None else torch.empty_like(phase_d)
    if not g_phase.is_contiguous() or g_phase.shape != phase_d.shape:
        raise ValueError("grad_out must be contiguous and shaped like phase")
    plan.run(n_samples=S, n_depth=n_depth, reduce_depth=True, in_kind=A.IN_COTANGENT, in0=field,
             cot_target=target_d, cot_scale=float(grad_scale), filter_kind=filt.kind,
             filter_flags=filt.adjoint_flags(), z=filt.z, depth_index=filt.depth_index,
             out_kind=A.OUT_GRAD_PHASE, out0=g_phase, aux_phase=phase_d, out_scale=plan.inv_n)
    return finish_loss(plan, partial, 1.0), g_phase


def field_to_field(plan, filt, n_depth, out, amp, phase, phase_scale=1.0):
    return _FieldToField.apply(plan, filt, n_depth, out, phase_scale, amp, phase)


def field_to_spectrum(plan, filt, amp, phase, phase_scale=1.0):
    return _FieldToSpectrum.apply(plan, filt, phase_scale, amp, phase)


def spectrum_to_field(plan, filt, n_depth, out, spec):
    return _SpectrumToField.apply(plan, filt, n_depth, out, spec)


def amplitude_mse(plan, filt, n_depth, amp, phase, target):
    return _AmplitudeMSE.apply(plan, filt, n_depth, amp, phase, target)
