"""Timing of the 8(f) stages (N1-N4) on one GPU: CUDA-event time per call, algorithmic bytes per call and the
achieved GB/s, beside the same stage written with the torch ops the reference uses (run on the same GPU).

    python tools/time_next.py [c4|c2]      # stack [8,3,2160,3840] (config 4) or [40,3,384,384] (config 2)
"""

from __future__ import annotations

import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from learned_hologram_gan_b200 import ap2poh_tail as T  # noqa: E402
from learned_hologram_gan_b200 import focal_stack_export as E  # noqa: E402
from learned_hologram_gan_b200 import loss_func as L  # noqa: E402


def timed(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def tv(x):
    return (x[..., :, 1:] - x[..., :, :-1]).abs().mean() + (x[..., 1:, :] - x[..., :-1, :]).abs().mean()


def torch_amp_terms(h, t):
    return F.mse_loss(h, t), (tv(h) - tv(t)).abs()


def torch_focal(f, r):
    sf, sr = torch.cat((f.sin(), f.cos()), 1), torch.cat((r.sin(), r.cos()), 1)
    d1 = ((sf[..., :, 1:] - sf[..., :, :-1]) - (sr[..., :, 1:] - sr[..., :, :-1])).abs()
    d2 = ((sf[..., 1:, :] - sf[..., :-1, :]) - (sr[..., 1:, :] - sr[..., :-1, :])).abs()
    with torch.no_grad():
        w1, w2 = d1 / d1.max(), d2 / d2.max()
    return (d1 * w1).mean() + (d2 * w2).mean()


def torch_export(x):
    mn, mx = x.amin((-2, -1), keepdim=True), x.amax((-2, -1), keepdim=True)
    n = (x - mn) / (mx - mn)
    rgba = torch.cat((n, torch.ones_like(n[:, :1])), 1).permute(0, 2, 3, 1)
    return (rgba * 255).to(torch.uint8).contiguous()


def torch_tail(field, w, b):
    def conv(x):
        return torch.cat([F.conv2d(x[:, c:c + 1], w[c][None, None], b[c:c + 1], padding=1) for c in range(3)], 1)
    m = torch.complex(conv(field.real), conv(field.imag))
    a, p = m.abs(), m.angle()
    a = a / (a.amax((-2, -1), keepdim=True) * 1.01)
    ac = a.acos()
    rows, cols = field.shape[-2:]
    odd = ((torch.arange(rows, device=field.device)[:, None] + torch.arange(cols, device=field.device)[None]) % 2).float()
    return (1 - odd) * (p + ac) + odd * (p - ac)


def measure(which="c4"):
    shape = (8, 3, 2160, 3840) if which == "c4" else (40, 3, 384, 384)
    gen = torch.Generator(device="cuda").manual_seed(0)
    h = torch.rand(shape, device="cuda", generator=gen)
    t = torch.rand(shape, device="cuda", generator=gen)
    n = h.numel()
    rows = []

    def row(name, ours, theirs, bytes_per_call):
        ms = timed(ours)
        ref_ms = timed(theirs, iters=3, warmup=1) if theirs is not None else None
        rows.append({"stage": name, "key": name.split(" ")[1], "ms": round(ms, 4), "algorithmic_GB": round(bytes_per_call / 1e9, 4),
                     "GB_per_s": round(bytes_per_call / ms / 1e6, 1),
                     "torch_ops_ms": None if ref_ms is None else round(ref_ms, 4)})

    row("N1 amp_loss_terms (mse + TV loss, forward)", lambda: L.amp_loss_terms(h, t, 1.0),
        lambda: torch_amp_terms(h, t), 8 * n)
    hg = h.clone().requires_grad_(True)

    def ours_fb():
        hg.grad = None
        L.amp_loss(hg, t, 1.0).backward()

    def torch_fb():
        hg.grad = None
        a, b = torch_amp_terms(hg, t)
        (a + b).backward()

    row("N1 amp_loss forward + backward", ours_fb, torch_fb, 8 * n + 12 * n)
    row("N1 focal_sincos_phase_gradient_loss (forward)", lambda: L.focal_sincos_phase_gradient_loss(h, t),
        lambda: torch_focal(h, t), 8 * n)

    def ours_ffb():
        hg.grad = None
        L.focal_sincos_phase_gradient_loss(hg, t).backward()

    def torch_ffb():
        hg.grad = None
        torch_focal(hg, t).backward()

    row("N1 focal loss forward + backward", ours_ffb, torch_ffb, 8 * n + 12 * n)
    row("N4 focal_stack_to_u8 (min/max + normalise + RGBA pack)", lambda: E.focal_stack_to_u8(h),
        lambda: torch_export(h), 4 * n + 4 * n + 4 * n // 3)
    row("N4 tensor_normalizor_2D", lambda: E.tensor_normalizor_2D(h),
        lambda: (h - h.amin((-2, -1), keepdim=True)) / (h.amax((-2, -1), keepdim=True) - h.amin((-2, -1), keepdim=True)),
        4 * n + 8 * n)
    nb = shape[0] if which == "c2" else 1
    field = torch.complex(torch.randn(nb, 3, *shape[-2:], device="cuda", generator=gen),
                          torch.randn(nb, 3, *shape[-2:], device="cuda", generator=gen))
    w = torch.rand(3, 3, 3, device="cuda", generator=gen)
    b = torch.zeros(3, device="cuda")
    with torch.no_grad():
        row("N2 ap2poh_tail (3x3 symmetric conv, normalise, double phase)", lambda: T.ap2poh_tail(field, w, b),
            lambda: torch_tail(field, w, b), field.numel() * 20)
    return {"workload": which, "shape": list(shape), "device": torch.cuda.get_device_name(0), "stages": rows}


def main():
    print(json.dumps(measure(sys.argv[1] if len(sys.argv) > 1 else "c4"), indent=1))


if __name__ == "__main__":
    main()
