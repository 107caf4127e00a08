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


def segment_cost(n_depth: int, fixed: float, per_depth: float) -> float:
    return fixed + per_depth * n_depth if n_depth > 0 else 0.0


def balanced_shards(n_colour: int, n_depth: int, world: int, fixed: float = 1.65,
                    per_depth: float = 1.0) -> List[List[Segment]]:
    """Partition of the (colour, depth) planes that minimises the slowest rank under a cost model.

    Every segment (one colour, a depth range) pays a fixed cost -- the forward transform of its colour, recomputed
    from the replicated phase, plus the last inverse transform and row pass of the adjoint -- and a per-depth cost.
    Equal plane counts (``plane_shards``) are therefore not equal times: a rank whose range straddles two colours
    pays the fixed cost twice.  The plan is found by next-fit packing of the colour-major plane sequence under a
    makespan T (a rank takes depths of the current colour while ``cost <= T``), T searched over the finite set of
    attainable costs.  ``fixed / per_depth = 1.65`` is the measured ratio on the 4320 x 7680 grid.
    Returns one segment list per rank (possibly empty: at world = 8 with 3 colours x 8 planes the best makespan is
    reached with two planes fewer than an even split would give the last ranks)."""

    def pack(T: float):
        ranks: List[List[Segment]] = [[] for _ in range(world)]
        r, used = 0, 0.0
        for c in range(n_colour):
            d = 0
            while d < n_depth:
                if r >= world:
                    return None
                room = T - used - fixed
                n = min(n_depth - d, int((room + 1e-9) // per_depth)) if room >= per_depth - 1e-9 else 0
                if n <= 0:
                    r, used = r + 1, 0.0
                    continue
                ranks[r].append(Segment(c, d, d + n))
                used += segment_cost(n, fixed, per_depth)
                d += n
        return ranks

    cands = sorted({k * fixed + m * per_depth for k in range(1, n_colour + 1) for m in range(1, n_colour * n_depth + 1)})
    best = None
    for T in cands:
        best = pack(T)
        if best is not None:
            break
    assert best is not None

    def makespan(ranks):
        return max(sum(segment_cost(s.n_depth, fixed, per_depth) for s in r) for r in ranks)

    # second candidate when there are at least as many ranks as colours: every rank works on ONE colour, colour c
    # gets k_c ranks (sum k_c = world) and splits its depths evenly among them.  Same makespan as the packing in the
    # cases that matter (world = 8: 3+3+2 ranks), but no rank is left without planes.
    if world >= n_colour and any(not r for r in best):
        k = [1] * n_colour
        for _ in range(world - n_colour):
            worst = max(range(n_colour), key=lambda c: (-(-n_depth // k[c]), -c))
            k[worst] += 1
        excl: List[List[Segment]] = []
        for c in range(n_colour):
            kc = min(k[c], n_depth)
            for i in range(kc):
                excl.append([Segment(c, i * n_depth // kc, (i + 1) * n_depth // kc)])
        excl += [[] for _ in range(world - len(excl))]
        if makespan(excl) <= makespan(best) + 1e-9:
            best = excl
    return best


def colour_owners(shards: Sequence[Sequence[Segment]], n_colour: int) -> List[List[int]]:
    """ranks that hold at least one plane of each colour (the sub-group its phase gradient is reduced in)"""
    return [[r for r, segs in enumerate(shards) if any(s.colour == c for s in segs)] for c in range(n_colour)]


SegmentFn = Callable[[Segment, torch.Tensor, torch.Tensor], "tuple[torch.Tensor, torch.Tensor]"]


class ShardedFocalStack:
    """Amplitude-L2 loss of a focal stack and its phase gradient, planes sharded over ranks.

    ``segment_fn(segment, phase_c [B,1,R,C], target [B*n,1,R,C]) -> (sum_sq_err, grad_phase_c)``
    computes one segment on the local device.  The default runs the CUDA path; tests inject a CPU
    stand-in to exercise the partitioning and the reductions under gloo.
    """

    def __init__(self, rows, cols, distances, pad_size, filter_radius_coefficient, pixel_pitch,
                 wave_length, world: Optional[int] = None, rank: Optional[int] = None,
                 segment_fn: Optional[SegmentFn] = None, group=None, balanced: bool = False,
                 colour_groups: bool = False):
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.distances = torch.as_tensor(distances, dtype=torch.float32)
        self.wave_length = torch.as_tensor(wave_length, dtype=torch.float32)
        self.n_colour = int(self.wave_length.numel())
        self.n_depth = int(self.distances.numel())
        # balanced: cost-model partition (balanced_shards) instead of equal plane counts
        self.shards = (balanced_shards(self.n_colour, self.n_depth, self.world) if balanced else
                       [plane_shards(self.n_colour, self.n_depth, self.world, r) for r in range(self.world)])
        self.segments = self.shards[self.rank]
        self.owners = colour_owners(self.shards, self.n_colour)
        self.owned_colours = sorted({s.colour for s in self.segments})
        # colour_groups: one process sub-group per colour held by more than one rank (created collectively, in
        # colour order, by every rank of the default group -- torch.distributed's rule for new_group)
        self.colour_group = [None] * self.n_colour
        if colour_groups and self.world > 1 and dist.is_initialized():
            for c, ranks in enumerate(self.owners):
                if len(ranks) > 1:
                    g = dist.new_group(ranks=ranks)
                    if self.rank in ranks:
                        self.colour_group[c] = g
        self.rows, self.cols = rows, cols
        self._geom = dict(sample_row_num=rows, sample_col_num=cols, pad_size=pad_size,
                          filter_radius_coefficient=filter_radius_coefficient, pixel_pitch=pixel_pitch)
        self._props = {}
        self._segment_fn = segment_fn or self._cuda_segment
        # measurement hook: set to a list to have every sharded step append (start, end) CUDA events recorded on the
        # compute stream around its collectives (the time the step waits for NCCL and for its slower peers)
        self.collective_events = None

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

    # ---- one step, colour-sharded result -----------------------------------------------------------------
    def loss_and_grad_sharded(self, phase: torch.Tensor, targets: Sequence[torch.Tensor], grads=None,
                              reduce_loss: bool = True):
        """The step of a colour-sharded optimisation loop: a rank only ever needs the phase planes of the colours it
        holds, so the phase gradient of colour c is summed among the ranks that hold planes of c
        (``self.owners[c]``, 33 MB at 4K: 0.08 ms over NVLink) instead of all-reducing the whole [B,3,R,C] gradient
        over every rank.  The reductions are issued (asynchronously, one NCCL stream per sub-group) only AFTER the
        last local segment has been enqueued: a collective issued earlier sits on the GPU waiting for its slower
        peers, and its resident CTAs take the shared memory the persistent one-CTA-per-SM column kernel counts on --
        measured at N = 2: the rank that reduced its first colour early ran its second segment at half speed
        (10.3 instead of 5.2 ms per step).  The step ends with stream-side waits only (no host sync).

        Returns (mean squared error over all planes of all ranks -- local partial if reduce_loss is False --,
        {colour: d loss / d phase[:, colour] as [B,1,R,C]}) for the colours in ``self.owned_colours``.  Without
        ``grads`` the gradients live in buffers owned by this object and are overwritten by the next call."""
        if len(targets) != len(self.segments):
            raise ValueError("one target tensor per local segment")
        batch = phase.shape[0]
        numel = batch * self.n_colour * self.n_depth * self.rows * self.cols
        dev = phase.device
        if grads is None:
            # persistent gradient buffers, like a parameter's .grad: a fresh 33 MB tensor per colour and step that is
            # then handed to NCCL's stream cannot be recycled by the caching allocator right away
            key = (batch, str(dev))
            if getattr(self, "_grad_key", None) != key:
                self._grad_bufs = {c: torch.empty((batch, 1, self.rows, self.cols), dtype=torch.float32, device=dev)
                                   for c in self.owned_colours}
                self._grad_key = key
            grads = self._grad_bufs
        sum_sq = torch.zeros((), dtype=torch.float32, device=dev)
        seen = set()
        for seg, tgt in zip(self.segments, targets):
            c = seg.colour
            phase_c = phase[:, c:c + 1]
            if self._segment_fn == self._cuda_segment:
                first = c not in seen
                out = grads[c] if first else torch.empty_like(grads[c])
                s, _ = self._prop(c).amplitude_mse_and_phase_gradient(
                    phase_c, self.distances[seg.d0:seg.d1], tgt, 2.0 / numel, grad_out=out)
                if not first:
                    grads[c] += out
            else:
                s, g = self._segment_fn(seg, phase_c.contiguous(), tgt)
                g = g.to(dev) / numel
                if c in seen:
                    grads[c] += g
                else:
                    grads[c].copy_(g)
            seen.add(c)
            sum_sq = sum_sq + s.to(dev)
        works = []
        probe = self.collective_events is not None
        if probe:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        for c in self.owned_colours:  # colour order on every rank
            if self.colour_group[c] is not None:
                works.append(dist.all_reduce(grads[c], op=dist.ReduceOp.SUM, group=self.colour_group[c], async_op=True))
        if self.world > 1 and reduce_loss and dist.is_initialized():
            works.append(dist.all_reduce(sum_sq, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()  # NCCL: a stream-side dependency, the host does not block
        if probe:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.collective_events.append((e0, e1))
        return sum_sq / numel, grads

    # ---- one step, replicated result -----------------------------------------------------------------------
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
