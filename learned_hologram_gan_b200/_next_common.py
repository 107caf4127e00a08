"""Host plumbing shared by the 8(f) stages: device staging, per-stream scratch, plane views."""

from __future__ import annotations

import ctypes as C
import threading
from typing import Optional

import torch

from . import _cabi_next as N
from .engine import compute_device

_LOCK = threading.Lock()
_SCRATCH: dict = {}


def lib():
    return N.load()


def stream_handle():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def staged(t: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """Contiguous copy-or-view of ``t`` on the compute device (host tensors are staged; nothing runs on the CPU)."""
    return t.detach().to(device=compute_device(), dtype=dtype).contiguous()


def planes_of(t: torch.Tensor):
    """(planes, rows, cols) of a tensor whose last two dims are the image."""
    if t.dim() < 2:
        raise ValueError(f"expected a tensor with at least 2 dims, got {tuple(t.shape)}")
    rows, cols = int(t.shape[-2]), int(t.shape[-1])
    if rows == 0 or cols == 0:
        raise ValueError(f"empty image dims in {tuple(t.shape)}")
    return t.numel() // (rows * cols), rows, cols


def scratch(n_floats: int, dev: torch.device) -> torch.Tensor:
    """Grow-only fp32 scratch per (device, stream, thread): the reductions' per-block partials."""
    n_floats = max(int(n_floats), 16)
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(n_floats, dtype=torch.float32, device=dev)
    key = (dev.index, torch.cuda.current_stream().cuda_stream, threading.get_ident())
    with _LOCK:
        buf = _SCRATCH.get(key)
        if buf is None or buf.numel() < n_floats:
            if len(_SCRATCH) >= 16:
                _SCRATCH.clear()
            buf = torch.empty(n_floats, dtype=torch.float32, device=dev)
            _SCRATCH[key] = buf
    return buf


def partial_for(planes: int, rows: int, cols: int, dev: torch.device) -> torch.Tensor:
    return scratch(lib().lhg_next_partial_floats(planes, rows, cols), dev)
