"""Focal-stack export: the step after the multi-distance propagation in ``generatePOH.py:72-78``.

``utilities.tensor_normalizor_2D`` (util.py:69-84) reads every plane five times (two max, two min, the affine
map) and ``multi_sample_plotter`` (util.py:179-203) moves fp32 planes to the host to have ``plt.imsave`` turn
them into 8-bit RGBA.  Here: one min/max reduction pass, then one pass that normalises and packs straight to
uint8 on the device, so the device-to-host copy is 4 bytes per pixel instead of 12.
"""

from __future__ import annotations

import os

import torch

from . import _cabi_next as N
from ._next_common import lib, planes_of, ptr, scratch, staged, stream_handle


def plane_minmax(t: torch.Tensor) -> torch.Tensor:
    """``[planes, 2]`` = (min, max) over the last two dims of every plane (device tensor)."""
    x = staged(t)
    planes, rows, cols = planes_of(x)
    out = torch.empty(planes, 2, dtype=torch.float32, device=x.device)
    partial = scratch(lib().lhg_next_partial_floats(planes, rows, cols), x.device)
    N.check(lib().lhg_plane_minmax(ptr(x), planes, rows * cols, ptr(partial), partial.numel(), ptr(out),
                                   stream_handle()))
    return out


def tensor_normalizor_2D(tensor_to_normalize: torch.Tensor) -> torch.Tensor:
    """util.py:69-84: (x - min) / (max - min) per plane, same fp32 operations, result on the input's device."""
    x = staged(tensor_to_normalize)
    planes, rows, cols = planes_of(x)
    mm = plane_minmax(x)
    out = torch.empty_like(x)
    N.check(lib().lhg_normalize_planes(ptr(x), ptr(mm), planes, rows * cols, ptr(out), stream_handle()))
    return out.to(tensor_to_normalize.device)


def amplitude_normalizor(amp: torch.Tensor) -> torch.Tensor:
    """util.py:53-66: amp / (1.01 * per-plane max), same fp32 operations, result on the input's device (forward only:
    inside the differentiable AP2POH tail the normalisation is part of ``ap2poh_tail``)."""
    x = staged(amp)
    planes, rows, cols = planes_of(x)
    mm = plane_minmax(x)
    out = torch.empty_like(x)
    N.check(lib().lhg_amplitude_normalize(ptr(x), ptr(mm), planes, rows * cols, ptr(out), stream_handle()))
    return out.to(amp.device)


def focal_stack_to_u8(amp: torch.Tensor, normalize: bool = True, alpha_channel: bool = True) -> torch.Tensor:
    """``[N,3,R,C]`` fp32 -> ``[N,R,C,4]`` (or 3) uint8 on the device: what ``plt.imsave`` stores for
    ``tensor_normalizor_2D(amp)[i].permute(1,2,0)``, i.e. ``(x*255).astype(uint8)`` with alpha 255."""
    x = staged(amp)
    if x.dim() == 3:
        x = x.unsqueeze(0)
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"expected [N,3,R,C], got {tuple(amp.shape)}")
    n, _, rows, cols = (int(s) for s in x.shape)
    oc = 4 if alpha_channel else 3
    mm = plane_minmax(x) if normalize else None
    out = torch.empty(n, rows, cols, oc, dtype=torch.uint8, device=x.device)
    N.check(lib().lhg_pack_rgb_u8(ptr(x), ptr(mm), n, rows, cols, oc, ptr(out), stream_handle()))
    return out


def save_focal_stack(amp: torch.Tensor, save_dir: str, titles=None, normalize: bool = True):
    """``multi_sample_plotter(tensor_normalizor_2D(amp), titles, rgb_img=True, save_dir)`` (generatePOH.py:72-78):
    one ``{title}.png`` (RGBA, 8 bit) per sample; titles default to ``range(N)``.  Returns the file names."""
    from PIL import Image

    u8 = focal_stack_to_u8(amp, normalize=normalize, alpha_channel=True)
    host = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=True)
    host.copy_(u8, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    if titles is None:
        titles = range(host.shape[0])
    os.makedirs(save_dir, exist_ok=True)
    names = []
    for i in range(host.shape[0]):
        name = os.path.join(save_dir, f"{titles[i]}.png")
        Image.fromarray(host[i].numpy(), mode="RGBA").save(name)
        names.append(name)
    return names
