"""Multi-GPU sharding of a focal stack: one process per GPU, planes partitioned, no data-path
collective in the forward; the backward needs one all-reduce of the phase gradient (planes of
one colour may live on several ranks) and one of the scalar loss.

A "plane" is one (colour, depth) pair of one RGB hologram.  Planes are flattened colour-major
(p = colour * n_depth + depth) and rank r owns the contiguous range
[r*P//world, (r+1)*P//world).  A rank recomputes the forward FFT of every colour it touches
from the replicated phase (cheaper than broadcasting a half-transformed spectrum).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Segment:
    colour: int
    d0: int
    d1: int  # exclusive

    @property
    def n_depth(self) -> int:
        return self.d1 - self.d0


def plane_shards(n_colour: int, n_depth: int, world: int, rank: int) -> List[Segment]:
    """Contiguous colour-major plane range of ``rank`` split into per-colour depth segments."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    total = n_colour * n_depth
    p0, p1 = rank * total // world, (rank + 1) * total // world
    segs = []
    p = p0
    while p < p1:
        c = p // n_depth
        d0 = p - c * n_depth
        d1 = min(n_depth, d0 + (p1 - p))
        segs.append(Segment(c, d0, d1))
        p += d1 - d0
    return segs


SegmentFn = Callable[[Segment, torch.Tensor, torch.Tensor], "tuple[torch.Tensor, torch.Tensor]"]


class ShardedFocalStack:
    """Amplitude-L2 loss of a focal stack and its phase gradient, planes sharded over ranks.

    ``segment_fn(segment, phase_c [B,1,R,C], target [B*n,1,R,C]) -> (sum_sq_err, grad_phase_c)``
    computes one segment on the local device.  The default runs the CUDA path; tests inject a CPU
    stand-in to exercise the partitioning and the reductions under gloo.
    """

    def __init__(self, rows, cols, distances, pad_size, filter_radius_coefficient, pixel_pitch,
                 wave_length, world: Optional[int] = None, rank: Optional[int] = None,
                 segment_fn: Optional[SegmentFn] = None, group=None):
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.distances = torch.as_tensor(distances, dtype=torch.float32)
        self.wave_length = torch.as_tensor(wave_length, dtype=torch.float32)
        self.n_colour = int(self.wave_length.numel())
        self.n_depth = int(self.distances.numel())
        self.segments = plane_shards(self.n_colour, self.n_depth, self.world, self.rank)
        self.rows, self.cols = rows, cols
        self._geom = dict(sample_row_num=rows, sample_col_num=cols, pad_size=pad_size,
                          filter_radius_coefficient=filter_radius_coefficient, pixel_pitch=pixel_pitch)
        self._props = {}
        self._segment_fn = segment_fn or self._cuda_segment

    # ---- local planes -------------------------------------------------------------------------
    def local_planes(self) -> int:
        return sum(s.n_depth for s in self.segments)

    def local_target_shape(self, batch: int, seg: Segment):
        return (batch * seg.n_depth, 1, self.rows, self.cols)

    def _prop(self, colour: int):
        if colour not in self._props:
            from .angular_spectrum_method import bandLimitedAngularSpectrumMethod_for_multiple_distances as M

            self._props[colour] = M(distances=self.distances, wave_length=self.wave_length[colour:colour + 1],
                                    band_limit=False, cuda=True, **self._geom)
        return self._props[colour]

    def _cuda_segment(self, seg: Segment, phase_c: torch.Tensor, target: torch.Tensor):
        # gradient of the SUM of squared errors (the caller divides by the global element count)
        return self._prop(seg.colour).amplitude_mse_and_phase_gradient(
            phase_c, self.distances[seg.d0:seg.d1], target, 2.0)

    # ---- single-rank fast path ----------------------------------------------------------------------
    def _full_prop(self):
        if "full" not in self._props:
            from .angular_spectrum_method import bandLimitedAngularSpectrumMethod_for_multiple_distances as M

            self._props["full"] = M(distances=self.distances, wave_length=self.wave_length, band_limit=False,
                                    cuda=True, **self._geom)
        return self._props["full"]

    def loss_and_grad_full(self, phase: torch.Tensor, target: torch.Tensor):
        """world == 1: every (colour, depth) plane is local, so the whole RGB stack goes through ONE
        forward and ONE adjoint call (target [B*n_depth, n_colour, R, C], index b*n_depth + d as in
        asm.py:516-518) instead of one pair per colour segment."""
        if self.world != 1:
            raise RuntimeError("loss_and_grad_full is the single-rank path")
        numel = target.numel()
        sum_sq, grad = self._full_prop().amplitude_mse_and_phase_gradient(phase, self.distances, target, 2.0 / numel)
        return sum_sq / numel, grad

    # ---- one step ---------------------------------------------------------------------------------
    def loss_and_grad(self, phase: torch.Tensor, targets: Sequence[torch.Tensor]):
        """phase [B,n_colour,R,C] (replicated on every rank); targets[i] belongs to segments[i].
        Returns (mean squared error over ALL planes of all ranks, d loss / d phase) on every rank."""
        if len(targets) != len(self.segments):
            raise ValueError("one target tensor per local segment")
        batch = phase.shape[0]
        numel = batch * self.n_colour * self.n_depth * self.rows * self.cols
        if self._segment_fn == self._cuda_segment and batch == 1:
            # library path, one sample: every segment writes its (already 1/numel-scaled) gradient straight
            # into its colour plane of the result; planes no local segment touches are zeroed
            grad = torch.empty_like(phase)
            sum_sq = torch.zeros((), dtype=torch.float32, device=phase.device)
            touched = set()
            for seg, tgt in zip(self.segments, targets):
                s, _ = self._prop(seg.colour).amplitude_mse_and_phase_gradient(
                    phase[:, seg.colour:seg.colour + 1], self.distances[seg.d0:seg.d1], tgt, 2.0 / numel,
                    grad_out=grad[:, seg.colour:seg.colour + 1])
                sum_sq = sum_sq + s
                touched.add(seg.colour)
            for c in range(self.n_colour):
                if c not in touched:
                    grad[:, c].zero_()
            if self.world > 1:
                dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group)
                dist.all_reduce(sum_sq, op=dist.ReduceOp.SUM, group=self.group)
            return sum_sq / numel, grad
        grad = torch.zeros_like(phase)
        sum_sq = torch.zeros((), dtype=torch.float32, device=phase.device)
        for seg, tgt in zip(self.segments, targets):
            s, g = self._segment_fn(seg, phase[:, seg.colour:seg.colour + 1].contiguous(), tgt)
            sum_sq = sum_sq + s.to(phase.device)
            grad[:, seg.colour:seg.colour + 1] += g.to(phase.device)
        if self.world > 1:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(sum_sq, op=dist.ReduceOp.SUM, group=self.group)
        return sum_sq / numel, grad / numel
