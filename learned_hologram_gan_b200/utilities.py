"""The few helpers of ``learnedMethodForHologram/utilities.py`` that sit on the propagation
path or that its callers need to reach it (util.py:30-50, :69-84, :206-243, :276-296, :410-415).

Like the reference these mask helpers are one-off host functions (their bits depend on the
host's sqrt, see engine.host_wm_grid); plotting, dataset and seeding helpers of the reference
are out of scope and, when the reference tree is available, are served from there by ``overlay``.
"""

from __future__ import annotations

import math

import torch


def num_gpus() -> int:
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def try_gpu(i=None) -> torch.device:
    """CUDA device to compute on (util.py:410-415).  ``i=None`` means the CURRENT device, so one
    process per GPU works under torchrun; on a single GPU this is the reference's ``cuda:0``."""
    n = num_gpus()
    if i is None:
        if n > 0:
            return torch.device("cuda", torch.cuda.current_device())
        i = 0
    if n > i:
        return torch.device(f"cuda:{i}")
    print(f"gpu with index '{i}' is not available")
    return torch.device("cpu")


def _radial(rows: int, cols: int) -> torch.Tensor:
    u = torch.fft.fftfreq(rows).unsqueeze(-1)
    v = torch.fft.fftfreq(cols).unsqueeze(0)
    return torch.sqrt(u**2 + v**2) * min(rows, cols)


def prepare_circular_frequency_mask_grid(samplingRowNum, samplingColNum):
    """sqrt(u^2+v^2)*min(rows, cols) on the un-shifted fftfreq grid (util.py:276-296)."""
    return _radial(samplingRowNum, samplingColNum)


def generate_circular_frequency_mask(sample_row_num=192, sample_col_num=192, radius=60, decay_rate=None):
    """Circular low-pass, optional exponential skirt (util.py:206-243)."""
    shorter_edge = min(sample_row_num, sample_col_num)
    if radius > shorter_edge / 2:
        raise ValueError(
            f"The radius {radius} is larger than the half of the sample size {shorter_edge/2}"
        )
    dist = _radial(sample_row_num, sample_col_num)
    mask = torch.ones_like(dist)
    outside = dist > radius
    if decay_rate is not None:
        mask[outside] = torch.exp(-decay_rate * (dist[outside] - radius))
    else:
        mask[outside] = 0.0
    return mask


def phase_tensor_generator(image_path_or_tensor):
    """PNG path -> [C,H,W] phase in [0, 2 pi]; tensors pass through (util.py:30-50)."""
    if isinstance(image_path_or_tensor, str):
        import numpy as np
        from PIL import Image

        img = np.asarray(Image.open(image_path_or_tensor))
        t = torch.from_numpy(img.copy())
        if t.dim() == 2:
            t = t.unsqueeze(-1)
        t = t.permute(2, 0, 1).contiguous()
        if t.dtype == torch.uint8:
            t = t.to(torch.float32).div(255)
        else:
            t = t.to(torch.float32)
        return t * 2 * math.pi
    if isinstance(image_path_or_tensor, torch.Tensor):
        return image_path_or_tensor
    raise ValueError("The input should be a string or a tensor.")


def tensor_normalizor_2D(tensor_to_normalize):
    """Per-plane min/max normalisation to [0,1] (util.py:69-84): one min/max reduction pass and one affine pass of
    ``next_stages.cu`` (``generatePOH.py:72-78`` calls this right after the multi-distance propagation)."""
    from .focal_stack_export import tensor_normalizor_2D as impl

    return impl(tensor_to_normalize)


def amplitude_normalizor(amp):
    """amp / (1.01 * per-plane max) (util.py:53-66), forward only (tensors that require grad keep torch's graph)."""
    if torch.is_grad_enabled() and amp.requires_grad:
        from .engine import compute_device

        a = amp.to(compute_device())  # differentiable; host tensors are staged, nothing is computed on the CPU
        return (a / (a.amax(dim=(-2, -1), keepdim=True) * 1.01)).to(amp.device)
    from .focal_stack_export import amplitude_normalizor as impl

    return impl(amp)
