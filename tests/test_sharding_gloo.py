"""world_size-2 run of the sharded focal-stack step on CPU (gloo): the partition, the phase-gradient
all-reduce and the loss all-reduce.  The per-segment compute is a CPU stand-in (the oracle) because
this container has no GPU; the CUDA segment function is exercised by bench.py --gpus N on the box."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import asm_oracle as O

from learned_hologram_gan_b200.sharding import ShardedFocalStack

ROWS, COLS, PAD, COEF, PITCH = 24, 40, 12, 0.45, 3.74e-6
WL = torch.tensor([638e-9, 520e-9, 450e-9])
Z = torch.linspace(4e-4, 10e-4, 5)
BATCH = 2


def inputs():
    gen = torch.Generator().manual_seed(122731)
    phase = 2 * torch.pi * torch.rand(BATCH, 3, ROWS, COLS, generator=gen)
    target = torch.rand(BATCH * Z.numel(), 3, ROWS, COLS, generator=gen)  # index b*D+d
    return phase, target


def oracle_segment(seg, phase_c, target):
    g = O.Geometry(rows=ROWS, cols=COLS, pad=PAD, radius_coef=COEF, pitch=PITCH,
                   wavelengths=WL[seg.colour:seg.colour + 1])
    p = phase_c.detach().clone().requires_grad_(True)
    g0 = O.spectrum_of(g, None, p)
    h = O.transfer_function(g, Z[seg.d0:seg.d1]) * O.diffraction_limited_mask(g)
    amp = torch.abs(O.field_from_spectrum(g, (g0.unsqueeze(1) * h).view(-1, 1, g.prow, g.pcol)))
    s = ((amp - target) ** 2).sum()
    s.backward()
    return s.detach(), p.grad


def worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    phase, target = inputs()
    stack = ShardedFocalStack(ROWS, COLS, Z, PAD, COEF, PITCH, WL, segment_fn=oracle_segment)
    D = Z.numel()
    tgts = []
    for seg in stack.segments:
        t = target.view(BATCH, D, 3, ROWS, COLS)[:, seg.d0:seg.d1, seg.colour:seg.colour + 1]
        tgts.append(t.reshape(BATCH * seg.n_depth, 1, ROWS, COLS).contiguous())
    loss, grad = stack.loss_and_grad(phase, tgts)
    if rank == 0:
        out.put((loss, grad, stack.local_planes()))
    dist.barrier()
    dist.destroy_process_group()


def worker_sharded(rank, world, port, out):
    """colour-sharded step: cost-model partition, per-colour sub-group reduction, async scalar-loss reduce"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    phase, target = inputs()
    stack = ShardedFocalStack(ROWS, COLS, Z, PAD, COEF, PITCH, WL, segment_fn=oracle_segment, balanced=True,
                              colour_groups=True)
    D = Z.numel()
    tgts = []
    for seg in stack.segments:
        t = target.view(BATCH, D, 3, ROWS, COLS)[:, seg.d0:seg.d1, seg.colour:seg.colour + 1]
        tgts.append(t.reshape(BATCH * seg.n_depth, 1, ROWS, COLS).contiguous())
    loss, grads = stack.loss_and_grad_sharded(phase, tgts)
    assert sorted(grads) == stack.owned_colours
    out.put((rank, loss, {c: g.clone() for c, g in grads.items()}, stack.owners))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_step_matches_the_single_process_oracle():
    phase, target = inputs()
    g = O.Geometry(rows=ROWS, cols=COLS, pad=PAD, radius_coef=COEF, pitch=PITCH, wavelengths=WL)
    loss_ref, grad_ref, _ = O.amp_mse_forward_backward(g, phase, Z, target)
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    loss, grad, local = out.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert local in (7, 8)  # 15 (colour, depth) planes over 2 ranks
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * loss_ref.item()
    assert O.rel_l2(grad, grad_ref) <= 1e-5


@pytest.mark.parametrize("world", [2, 4])
def test_colour_sharded_step_matches_the_single_process_oracle(world):
    """Every rank ends with the FULL gradient of the colours it holds (summed inside that colour's sub-group) and
    with the global loss; colours it does not hold are not communicated at all."""
    phase, target = inputs()
    g = O.Geometry(rows=ROWS, cols=COLS, pad=PAD, radius_coef=COEF, pitch=PITCH, wavelengths=WL)
    loss_ref, grad_ref, _ = O.amp_mse_forward_backward(g, phase, Z, target)
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = free_port()
    procs = [ctx.Process(target=worker_sharded, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get() for _ in range(world)]
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    covered = set()
    for rank, loss, grads, owners in results:
        assert abs(loss.item() - loss_ref.item()) <= 1e-5 * loss_ref.item()
        for c, gc in grads.items():
            assert rank in owners[c]
            assert O.rel_l2(gc, grad_ref[:, c:c + 1]) <= 1e-5, (rank, c)
            covered.add(c)
    assert covered == {0, 1, 2}


def test_balanced_partition_covers_every_plane_once_and_beats_the_even_split():
    from learned_hologram_gan_b200.sharding import balanced_shards, plane_shards, segment_cost

    for n_depth in (1, 5, 8, 64):
        for world in (1, 2, 3, 4, 5, 8, 16):
            shards = balanced_shards(3, n_depth, world)
            assert len(shards) == world
            seen = {}
            for segs in shards:
                for s_ in segs:
                    for d in range(s_.d0, s_.d1):
                        assert (s_.colour, d) not in seen
                        seen[(s_.colour, d)] = 1
            assert len(seen) == 3 * n_depth
            cost = lambda segs: sum(segment_cost(s_.n_depth, 1.65, 1.0) for s_ in segs)  # noqa: E731
            even = max(cost(plane_shards(3, n_depth, world, r)) for r in range(world))
            assert max(cost(segs) for segs in shards) <= even + 1e-9
    # world = 8, 3 colours x 8 planes: one colour per rank, nobody idle
    shards = balanced_shards(3, 8, 8)
    assert all(len(segs) == 1 for segs in shards)
