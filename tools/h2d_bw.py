"""Pinned host -> device copy bandwidth of this box (what bounds bench.py's e2e numbers: 896 MB of fp32 inputs per C4
step and rank).  Alone: `python tools/h2d_bw.py`.  All ranks at once (the aggregate ceiling of the host the N GPUs
share): `python -m torch.distributed.run --nproc-per-node N tools/h2d_bw.py` -- every rank copies from its own pinned,
first-touched buffer at the same time; rank 0 prints per-rank and aggregate GB/s."""
import json
import os

import torch

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 796 * 1024 * 1024 // 4
h = torch.empty(n, dtype=torch.float32, pin_memory=True)
h.fill_(1.0)  # first touch by this rank
d = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    d.copy_(h, non_blocking=True)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
gbs = n * 4 / ms / 1e6
if world > 1:
    t = torch.tensor([gbs], device="cuda")
    all_t = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(all_t, t)
    if rank == 0:
        per = [round(float(x), 1) for x in all_t]
        print(json.dumps({"ranks": world, "h2d_bytes_per_rank": n * 4, "GB_per_s_per_rank": per,
                          "GB_per_s_aggregate": round(sum(per), 1)}))
    dist.destroy_process_group()
else:
    print(json.dumps({"h2d_bytes": n * 4, "ms": ms, "GB_per_s": gbs}))
