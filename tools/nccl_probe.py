"""Time NCCL all-reduces the sharded step issues (diagnosis): torchrun --nproc-per-node N tools/nccl_probe.py"""
import os, sys, time
import torch
import torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sub = dist.new_group(ranks=list(range(min(world, 3))))
pair = dist.new_group(ranks=[0, 1])
x = torch.ones(1, 1, 2160, 3840, device=dev)
s = torch.zeros((), device=dev)
big = torch.ones(1, 3, 2160, 3840, device=dev)


def t(name, fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{name:50s} {a.elapsed_time(b) / n:8.3f} ms", flush=True)


def async_ar(buf, group):
    w = dist.all_reduce(buf, group=group, async_op=True)
    w.wait()


t("world all_reduce 100 MB", lambda: dist.all_reduce(big))
t("world all_reduce 33 MB", lambda: dist.all_reduce(x))
t("world all_reduce scalar", lambda: dist.all_reduce(s))
if rank < min(world, 3):
    t("sub-group(<=3) all_reduce 33 MB", lambda: dist.all_reduce(x, group=sub))
    t("sub-group(<=3) all_reduce 33 MB async+wait", lambda: async_ar(x, sub))
else:
    for _ in range(2):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
if rank < 2:
    t("pair all_reduce 33 MB", lambda: dist.all_reduce(x, group=pair))
else:
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
t("world all_reduce scalar async+wait", lambda: async_ar(s, None))


def both():
    w1 = dist.all_reduce(x, group=sub if rank < min(world, 3) else None, async_op=True) if rank < min(world, 3) else None
    w2 = dist.all_reduce(s, async_op=True)
    if w1 is not None:
        w1.wait()
    w2.wait()


t("sub 33 MB async + world scalar async, then waits", both)
dist.destroy_process_group()
