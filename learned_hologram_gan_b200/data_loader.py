"""Raw ``.bin`` dataset reader: mirror of ``watermelon_hologram/data_loader.py``.

The reference memory-maps fp32 ``[N,C,H,W]`` files and, per item, builds ``torch.tensor(memmap[idx])`` (a pageable
host copy) followed by a blocking ``.to(device)`` for each of the 3-4 arrays (dl.py:39-54).  Same classes, same
``__getitem__`` values here; in addition ``fetch(indices)`` builds a whole batch with ONE multi-threaded native
gather into pinned memory (``lhg_bin_gather``), ONE asynchronous upload per array and the assembly (RGB + first
depth plane, 2*pi*phase) on the device.  Only the first depth plane is read from disk and uploaded (the reference
reads all three and keeps one, dl.py:46).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch.utils.data import Dataset

from . import _cabi_next as N
from ._next_common import lib, ptr, stream_handle
from .engine import compute_device


class _BinFile:
    def __init__(self, path, shape):
        self.shape = tuple(int(s) for s in shape)
        self.map = np.memmap(path, dtype=np.float32, mode="r", shape=self.shape)
        self.item_floats = int(np.prod(self.shape[1:]))
        self.plane_floats = int(self.shape[2] * self.shape[3])
        self._pinned = None

    def gather(self, idx: np.ndarray, floats_per_item: int, threads: int = 0) -> torch.Tensor:
        """Pinned ``[n, floats_per_item]`` holding the leading floats of the selected items."""
        n = int(idx.shape[0])
        need = n * floats_per_item
        if self._pinned is None or self._pinned.numel() < need:
            self._pinned = torch.empty(max(need, 1), dtype=torch.float32, pin_memory=torch.cuda.is_available())
        dst = self._pinned[:need]
        N.check(lib().lhg_bin_gather(C.c_void_p(self.map.ctypes.data), self.shape[0], self.item_floats * 4,
                                     floats_per_item * 4, C.c_void_p(idx.ctypes.data), n,
                                     C.c_void_p(dst.data_ptr()), threads))
        return dst.view(n, floats_per_item)

    def upload(self, idx, floats_per_item, dev, threads=0):
        host = self.gather(idx, floats_per_item, threads)
        out = torch.empty(host.shape, dtype=torch.float32, device=dev)
        out.copy_(host, non_blocking=True)
        # the pinned staging buffer is reused by the next fetch: the copy must have left it
        self._event = torch.cuda.Event()
        self._event.record()
        return out

    def wait(self):
        ev = getattr(self, "_event", None)
        if ev is not None:
            ev.synchronize()
            self._event = None


def _indices(indices, n):
    idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int64).reshape(-1))
    if idx.size and (idx.min() < 0 or idx.max() >= n):
        raise IndexError("Index out of range")
    return idx


class _Base(Dataset):
    def _init(self, samplesNum, channlesNum, height, width, cuda):
        self.dataShape = (samplesNum, channlesNum, height, width)
        # cuda=False in the reference means "items stay on the host"; the batch path below always delivers to the
        # compute device, __getitem__ honours the flag
        self.device = compute_device() if cuda else torch.device("cpu")

    def __len__(self):
        return self.dataShape[0]

    def _check(self, idx):
        if idx < 0 or idx >= len(self):
            raise IndexError("Index out of range")

    def _item(self, f: _BinFile, idx, planes=None):
        a = f.map[idx] if planes is None else f.map[idx][:planes]
        return torch.from_numpy(np.array(a)).to(self.device)


def _rgbd(img_d, depth_d, n, plane, dev):
    out = torch.empty(n, 4, plane, dtype=torch.float32, device=dev)
    N.check(lib().lhg_assemble_rgbd(ptr(img_d), ptr(depth_d), 1, n, plane, ptr(out), stream_handle()))
    return out


class dataloaderImgDepthAmpPhs(_Base):
    """dl.py:8-54: item = (cat(img, depth[0:1]) [4,H,W], amp [3,H,W], phs [3,H,W])."""

    def __init__(self, img_path, depth_path, amp_path, phs_path, samplesNum=3800, channlesNum=3, height=192,
                 width=192, cuda=False):
        self._init(samplesNum, channlesNum, height, width, cuda)
        self.img, self.depth = _BinFile(img_path, self.dataShape), _BinFile(depth_path, self.dataShape)
        self.amp, self.phs = _BinFile(amp_path, self.dataShape), _BinFile(phs_path, self.dataShape)

    def __getitem__(self, idx):
        self._check(idx)
        return (torch.cat((self._item(self.img, idx), self._item(self.depth, idx, 1)), dim=0),
                self._item(self.amp, idx), self._item(self.phs, idx))

    def fetch(self, indices, threads=0):
        """Batch ``([n,4,H,W], [n,C,H,W], [n,C,H,W])`` on the compute device (the default collate of the items)."""
        idx = _indices(indices, len(self))
        dev, n = compute_device(), int(idx.shape[0])
        _, c, h, w = self.dataShape
        for f in (self.img, self.depth, self.amp, self.phs):
            f.wait()
        img_d = self.img.upload(idx, self.img.item_floats, dev, threads)
        dep_d = self.depth.upload(idx, self.depth.plane_floats, dev, threads)
        amp_d = self.amp.upload(idx, self.amp.item_floats, dev, threads)
        phs_d = self.phs.upload(idx, self.phs.item_floats, dev, threads)
        if c != 3:
            raise ValueError("RGB + depth assembly needs 3 image channels")
        rgbd = _rgbd(img_d, dep_d, n, h * w, dev).view(n, 4, h, w)
        return rgbd, amp_d.view(n, c, h, w), phs_d.view(n, c, h, w)


class dataloaderAmpPIPhs(_Base):
    """dl.py:57-86: item = (amp, 2*pi*phs)."""

    def __init__(self, amp_path, phs_path, samplesNum=3800, channlesNum=3, height=192, width=192, cuda=False):
        self._init(samplesNum, channlesNum, height, width, cuda)
        self.amp, self.phs = _BinFile(amp_path, self.dataShape), _BinFile(phs_path, self.dataShape)

    def __getitem__(self, idx):
        self._check(idx)
        return self._item(self.amp, idx), 2 * torch.pi * self._item(self.phs, idx)

    def fetch(self, indices, threads=0):
        idx = _indices(indices, len(self))
        dev, n = compute_device(), int(idx.shape[0])
        _, c, h, w = self.dataShape
        self.amp.wait()
        self.phs.wait()
        amp_d = self.amp.upload(idx, self.amp.item_floats, dev, threads)
        phs_d = self.phs.upload(idx, self.phs.item_floats, dev, threads)
        out = torch.empty_like(phs_d)
        N.check(lib().lhg_scale_two_pi(ptr(phs_d), phs_d.numel(), ptr(out), stream_handle()))
        return amp_d.view(n, c, h, w), out.view(n, c, h, w)


class dataloaderImgDepth(_Base):
    """dl.py:89-123: item = cat(img, depth[0:1]) [4,H,W]."""

    def __init__(self, img_path, depth_path, samplesNum=3800, channlesNum=3, height=192, width=192, cuda=False):
        self._init(samplesNum, channlesNum, height, width, cuda)
        self.img, self.depth = _BinFile(img_path, self.dataShape), _BinFile(depth_path, self.dataShape)

    def __getitem__(self, idx):
        self._check(idx)
        return torch.cat((self._item(self.img, idx), self._item(self.depth, idx, 1)), dim=0)

    def fetch(self, indices, threads=0):
        idx = _indices(indices, len(self))
        dev, n = compute_device(), int(idx.shape[0])
        _, c, h, w = self.dataShape
        if c != 3:
            raise ValueError("RGB + depth assembly needs 3 image channels")
        self.img.wait()
        self.depth.wait()
        img_d = self.img.upload(idx, self.img.item_floats, dev, threads)
        dep_d = self.depth.upload(idx, self.depth.plane_floats, dev, threads)
        return _rgbd(img_d, dep_d, n, h * w, dev).view(n, 4, h, w)
