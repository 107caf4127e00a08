#!/usr/bin/env python
"""bench.py -- RGB depth-plane propagations/s (forward + adjoint backward) of the band-limited
angular-spectrum path, with the HBM-roofline fraction of the dominant kernel and the
reference's CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|stages-c4|stages-c2] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of synthetic input: multi-distance
propagation of a random-phase RGB POH to D depth planes, amplitude-L2 loss against random
targets, and the adjoint back to the phase (BASELINE.json configs[3] "c4" by default: 3840x2160
POH, 2x zero-padded to 7680x4320, 8 planes = 24 propagations per step; "c2" = configs[1]:
batch 4 of 384x384 padded to 1024x1024, 10 planes = 120 propagations per step).

N > 1 (launched by torchrun, one rank per GPU), default `--scaling weak`: the units of the path -- whole
holograms with all their (colour, depth) planes -- are partitioned over the ranks, one hologram per rank and
step, no data-path collective (NCCL carries the scalar loss only); `value` = the propagations all ranks
processed / the slowest rank's time.  `--scaling strong` shards the 24 planes of ONE hologram over the ranks
instead (BASELINE config 4's wording); NCCL then also all-reduces the phase gradient.
`--workload stages-c4|stages-c2` measures the stages either side of the path (loss epilogues,
AP2POH tail, focal-stack export) at the focal-stack size of that config instead (run_stages below).
`--impl reference` times the CPU oracle port of the reference (oracle/asm_oracle.py) on the
host cores on a bounded sample (fewer depth planes) of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WL = [638e-9, 520e-9, 450e-9]
PITCH = 3.74e-6
WORKLOADS = {
    # BASELINE.json configs[3]
    "c4": dict(rows=2160, cols=3840, pad=1080, coef=0.45, batch=1, depths=8, z0=4e-4, z1=10e-4,
               cpu_depths=1, name="C4 4K POH 3840x2160 -> 7680x4320 padded, RGB x 8 planes, fwd+L2+adjoint"),
    # BASELINE.json configs[1]
    "c2": dict(rows=384, cols=384, pad=320, coef=0.35, batch=4, depths=10, z0=4e-4, z1=10e-4,
               cpu_depths=10, name="C2 batch 4 of 384x384 -> 1024x1024 padded, RGB x 10 planes, fwd+L2+adjoint"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc, self.first = index, [], None, 0

    def mark(self):
        """Samples from here on belong to the timed region (the warm-up ones are kept only if none follow)."""
        self.first = len(self.samples)

    def start(self):
        """In-process NVML polling (nvidia_ml_py) every 25 ms; the nvidia-smi -lms loop is the fallback
        (LHG_CLOCK_SAMPLER=smi forces it).  A polling nvidia-smi process was seen to stretch single steps of the timed
        region by 10-40 ms in about one run in five (profiles/r01_final.md)."""
        self.nvml = None
        if os.environ.get("LHG_CLOCK_SAMPLER", "nvml") != "smi":
            try:
                import pynvml

                pynvml.nvmlInit()
                pr = torch.cuda.get_device_properties(self.index)
                bus = f"{pr.pci_domain_id:08X}:{pr.pci_bus_id:02X}:{pr.pci_device_id:02X}.0"
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
                self.nvml = pynvml
                self.stop_flag = False
                self.proc = True
                threading.Thread(target=self._poll_nvml, daemon=True).start()
                return
            except Exception:
                self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        names = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                bits = int(reasons_fn(self.handle))
                flags = ["Active" if bits & b else "Not Active" for b, _ in names]
                self.samples.append(", ".join([str(sm), str(mx)] + flags))
            except Exception:
                pass
            time.sleep(0.025)

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def wait_ready(self, timeout=5.0):
        """Block until the first sample has arrived: a sampler's start-up (NVML initialisation) stalls the GPU for tens
        of milliseconds, which must not land inside the timed region."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if getattr(self, "nvml", None) is not None:
            self.stop_flag = True
        else:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        timed = self.samples[self.first:] or self.samples
        for s in timed:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "sampler": "nvml" if getattr(self, "nvml", None) is not None else "nvidia-smi"}


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run this rank (and therefore allocate its pinned staging buffers, first touch) on the
    NUMA node its GPU hangs off, so the per-step host -> device copies of the e2e loop do not cross the socket
    interconnect.  Best effort: returns a description, never raises."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return f"gpu {bus}: no NUMA node reported"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"gpu {bus}: node {node} has no CPU this process may use"
        os.sched_setaffinity(0, cpus)
        return f"gpu {bus}: bound to NUMA node {node} ({len(cpus)} cpus)"
    except Exception as e:  # noqa: BLE001
        return f"no NUMA binding ({type(e).__name__}: {e})"


def make_inputs(wl, world, rank):
    """Seeded synthetic POH phase and target amplitudes, generated on the host (pinned)."""
    gen = torch.Generator().manual_seed(122731 + rank)  # rank > 0 only under weak scaling: its own hologram
    B, R, C, D = wl["batch"], wl["rows"], wl["cols"], wl["depths"]
    phase = (2 * torch.pi * torch.rand(B, 3, R, C, generator=gen)).pin_memory()
    return phase, gen


def algorithmic_bytes(wl, seg_depths, fused):
    """SURVEY.md 8(d) streaming-pass floor per (sample, colour) group with D planes.  fused: the row-inverse
    pass of the forward and the row-forward pass of the adjoint are one kernel (k4): |y| and the saved field are
    neither written nor read back, the target is read once."""
    R, C = wl["rows"], wl["cols"]
    Cp = C + 2 * int(wl["pad"] * (C / R))
    out = {"k1": 0, "k2": 0, "k3": 0, "k4": 0, "total": 0}
    for D in seg_depths:
        k2 = (8 * R * Cp + D * 8 * R * Cp) * 2                                   # fwd 1->D, bwd D->1
        if fused:
            k1 = 4 * R * C + 8 * R * Cp                                          # fwd prologue
            k3 = 8 * R * Cp + 8 * R * C                                          # bwd epilogue (phase in, grad out)
            k4 = D * (8 * R * Cp + 4 * R * C + 8 * R * Cp)                       # W2 row + target in, W1' row out
        else:
            k1 = (4 * R * C + 8 * R * Cp) + D * (12 * R * C + 8 * R * Cp)          # fwd + bwd prologue
            k3 = D * (8 * R * Cp + 4 * R * C + 8 * R * C) + (8 * R * Cp + 8 * R * C)  # fwd (+save) + bwd
            k4 = 0
        out["k1"] += k1 * wl["batch"]
        out["k2"] += k2 * wl["batch"]
        out["k3"] += k3 * wl["batch"]
        out["k4"] += k4 * wl["batch"]
    out["total"] = out["k1"] + out["k2"] + out["k3"] + out["k4"]
    return out


def cpu_step_factory(wl):
    """One bounded-sample step of the reference's CPU path (oracle port): forward, L2, backward.
    w_grid and the mask are built once, as the reference does in its constructor."""
    from oracle import asm_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    D = wl["cpu_depths"]
    z = torch.linspace(wl["z0"], wl["z1"], wl["depths"])[:D]
    gen = torch.Generator().manual_seed(122731)
    phase = 2 * torch.pi * torch.rand(wl["batch"], 3, wl["rows"], wl["cols"], generator=gen)
    target = torch.rand(wl["batch"] * D, 3, wl["rows"], wl["cols"], generator=gen)
    g = O.Geometry(rows=wl["rows"], cols=wl["cols"], pad=wl["pad"], radius_coef=wl["coef"], pitch=PITCH,
                   wavelengths=torch.tensor(WL))
    w = O.w_grid(g)
    mask = O.diffraction_limited_mask(g)

    def step():
        p = phase.clone().requires_grad_(True)
        g0 = O.spectrum_of(g, torch.ones_like(p), p)
        h = O.transfer_function(g, z, w) * mask
        gz = (g0.unsqueeze(1) * h).view(-1, 3, g.prow, g.pcol)
        amp = torch.abs(O.field_from_spectrum(g, gz))
        loss = torch.nn.functional.mse_loss(amp, target)
        loss.backward()
        return loss

    props = wl["batch"] * 3 * D
    sample = (f"{wl['name']}; CPU sample = first {D} of {wl['depths']} depth planes per step "
              f"({props} propagations/step)")
    return step, props, sample


def time_cpu(step, warm, steps):
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def run_reference(args, wl, rank, world):
    """The reference's own CPU implementation of the path (oracle port), all host threads."""
    if rank != 0:
        return
    step, props, sample = cpu_step_factory(wl)
    warm = min(args.warmup, 1)
    steps = max(1, min(args.steps, 3 if args.workload == "c4" else 5))
    dt = time_cpu(step, warm, steps)
    value = props / dt
    line = {
        "impl": "reference", "metric": "rgb_depth_plane_propagations_per_s_fwd_bwd", "value": value,
        "unit": "propagations/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "key": args.workload},
        "cpu_baseline": {"value": value, "unit": "propagations/s", "cores": torch.get_num_threads(),
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "propagations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(wl):
    step, props, sample = cpu_step_factory(wl)
    dt = time_cpu(step, 1, 2)
    return {"value": props / dt, "unit": "propagations/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": sample + f", {dt:.2f} s/step"}


def run_stages(args, rank):
    """`--workload stages-c4|stages-c2`: the stages either side of the path at the
    config-4 / config-2 focal-stack size: CUDA-event time per call, algorithmic GB/s against the measured HBM
    peak, the same stage written with the reference's torch ops on the same GPU, and the reference's CPU code
    (oracle port) on the host cores on a bounded sample."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import time_next

    which = args.workload.split("-")[1]
    res = time_next.measure(which)
    peak, peak_src = peaks()
    cpu = {}
    if not args.no_cpu_baseline:
        from oracle import next_oracle as NO

        torch.set_num_threads(os.cpu_count() or 1)
        shape = (2, 3) + tuple(res["shape"][-2:]) if which == "c4" else tuple(res["shape"])
        gen = torch.Generator().manual_seed(122731)
        h, t = torch.rand(shape, generator=gen), torch.rand(shape, generator=gen)
        hg = h.clone().requires_grad_(True)
        field = torch.complex(torch.randn((1, 3) + shape[-2:], generator=gen), torch.randn((1, 3) + shape[-2:], generator=gen))
        w, b = torch.rand(3, 3, 3, generator=gen), torch.zeros(3)
        n = h.numel()

        def fb(fn):
            def f():
                hg.grad = None
                fn(hg, t).backward()
            return f

        jobs = {
            "amp_loss_terms": (lambda: NO.amp_loss(h, t), 8 * n),
            "amp_loss": (fb(NO.amp_loss), 20 * n),
            "focal_sincos_phase_gradient_loss": (lambda: NO.focal_sincos_phase_gradient_loss(h, t), 8 * n),
            "focal": (fb(NO.focal_sincos_phase_gradient_loss), 20 * n),
            "focal_stack_to_u8": (lambda: NO.focal_stack_u8(h), 4 * n + 4 * n + 4 * n // 3),
            "tensor_normalizor_2D": (lambda: NO.tensor_normalizor_2D(h), 12 * n),
            "ap2poh_tail": (lambda: NO.ap2poh_tail(field, w, b), 20 * field.numel()),
        }
        for key, (fn, nbytes) in jobs.items():
            dt = time_cpu(fn, 1, 2)
            cpu[key] = {"value": nbytes / dt / 1e9, "unit": "GB/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{list(shape) if key != 'ap2poh_tail' else list(field.shape)} on the host, {dt * 1e3:.1f} ms/call"}
    total_b = total_ms = 0.0
    for st in res["stages"]:
        st["frac_of_hbm_peak"] = st["GB_per_s"] / peak
        if st["key"] in cpu:
            st["cpu_baseline"] = cpu[st["key"]]
        total_b += st["algorithmic_GB"]
        total_ms += st["ms"]
    line = {"metric": "adjacent_stage_algorithmic_GB_per_s", "value": total_b / (total_ms * 1e-3), "unit": "GB/s",
            "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": total_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"8(f) stages N1/N2/N4 on a {res['shape']} focal stack", "key": args.workload},
            "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                         "achieved": total_b / (total_ms * 1e-3), "frac": total_b / (total_ms * 1e-3) / peak,
                         "traffic": None},
            "stages": res["stages"]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["stages-c4", "stages-c2"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.workload.startswith("stages"):
        if args.impl == "reference":  # the CPU port of every stage is timed inside the stages line itself
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "stage workloads report their CPU port in "
                                  "stages[*].cpu_baseline of the b200 line"}), flush=True)
            return
        run_stages(args, rank)
        return
    wl = WORKLOADS[args.workload]
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch.distributed as dist

    # stdout carries exactly ONE JSON line: everything libraries print there (NCCL's version banner at
    # communicator creation) is sent to stderr; the line itself goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    from learned_hologram_gan_b200 import _cabi
    from learned_hologram_gan_b200.sharding import ShardedFocalStack

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single rank: not bound"
    print(f"rank {rank}: {numa}", file=sys.stderr)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()
    warmup = max(args.warmup, 3)

    z = torch.linspace(wl["z0"], wl["z1"], wl["depths"])
    # weak scaling: every rank owns a whole hologram (all planes local); strong: the planes of one are sharded
    weak = world > 1 and args.scaling == "weak"
    local_world, local_rank_id = (1, 0) if weak else (world, rank)
    stack = ShardedFocalStack(wl["rows"], wl["cols"], z, wl["pad"], wl["coef"], PITCH, torch.tensor(WL),
                              world=local_world, rank=local_rank_id)
    phase_h, gen = make_inputs(wl, world, rank if weak else 0)
    B = wl["batch"]
    # every rank draws the full target stream so the global job is independent of N; keeps its planes
    targets_h = []
    all_t = {}
    for c in range(3):
        for d in range(wl["depths"]):
            t = torch.rand(B, 1, wl["rows"], wl["cols"], generator=gen)
            all_t[(c, d)] = t
    for seg in stack.segments:
        # target layout of one segment: index b*n_depth + d  (asm.py:516-518)
        t = torch.stack([all_t[(seg.colour, d)] for d in range(seg.d0, seg.d1)], dim=1)
        targets_h.append(t.reshape(B * seg.n_depth, 1, wl["rows"], wl["cols"]).contiguous().pin_memory())
    full_h = None
    single = local_world == 1
    if single:
        # one rank owns every plane: the whole RGB stack is ONE forward + ONE adjoint call
        full_h = torch.stack([torch.stack([all_t[(c, d)][:, 0] for c in range(3)], dim=1)
                              for d in range(wl["depths"])], dim=1)  # [B, D, 3, R, C]
        full_h = full_h.reshape(B * wl["depths"], 3, wl["rows"], wl["cols"]).contiguous().pin_memory()
        targets_h = [full_h]
    del all_t
    phase_d = phase_h.to(dev)
    targets_d = [t.to(dev) for t in targets_h]

    def reduce_loss(loss):
        if weak:  # the only collective of the weak-scaling job: the scalar loss
            loss = loss.clone()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
            loss /= world
        return loss

    def step_resident():
        if single:
            loss, grad = stack.loss_and_grad_full(phase_d, targets_d[0])
            return reduce_loss(loss), grad
        return stack.loss_and_grad(phase_d, targets_d)

    # end to end: every step copies ITS inputs from pinned host memory and reads its loss back.  The copies
    # of step i+1 run on a second stream into the other of two device buffers while step i computes
    # (what a training loop's prefetching loader does); all of them lie inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [(torch.empty_like(phase_d), [torch.empty_like(t) for t in targets_d]) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]   # slot filled (recorded on the copy stream)
    free = [torch.cuda.Event(), torch.cuda.Event()]    # slot consumed (recorded on the compute stream)

    def stage(slot):
        copy_stream.wait_event(free[slot])
        with torch.cuda.stream(copy_stream):
            slots[slot][0].copy_(phase_h, non_blocking=True)
            for dst, src in zip(slots[slot][1], targets_h):
                dst.copy_(src, non_blocking=True)
            ready[slot].record(copy_stream)

    def run_e2e(steps):
        out = None
        main = torch.cuda.current_stream(dev)
        for ev in free:
            ev.record(main)
        stage(0)
        for i in range(steps):
            slot = i & 1
            if i + 1 < steps:
                stage(slot ^ 1)
            main.wait_event(ready[slot])
            p, ts = slots[slot]
            if single:
                loss, grad = stack.loss_and_grad_full(p, ts[0])
                loss = reduce_loss(loss)
            else:
                loss, grad = stack.loss_and_grad(p, ts)
            free[slot].record(main)
            out = (loss.item(), grad)  # device -> host read of the step's result
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps, out

    clocks = ClockSampler(local_rank)
    clocks.start()  # started before the warm-up ...
    clocks.wait_ready()  # ... and past its start-up before anything is timed
    for _ in range(warmup):
        step_resident()
    clocks.mark()
    lib.asm_profile_enable(1)
    l0 = lib.asm_launch_count()
    ms_step, (loss, grad) = timed(step_resident, args.steps)
    l1 = lib.asm_launch_count()
    lib.asm_profile_enable(0)
    import ctypes as C

    kms = (C.c_double * 4)(0, 0, 0, 0)
    kn = (C.c_longlong * 4)(0, 0, 0, 0)
    lib.asm_profile_collect(kms, kn, 4)
    clk = clocks.stop()

    run_e2e(2)
    n_e2e = max(2, args.steps)  # the first copy of the pipeline has nothing to overlap with: amortised over K steps
    ms_e2e, _ = timed(lambda: run_e2e(n_e2e), 1)
    ms_e2e /= n_e2e

    if os.environ.get("LHG_E2E_PROBE"):  # diagnosis: the copies of the e2e loop alone (no compute), to stderr
        def copies_only():
            main = torch.cuda.current_stream(dev)
            for ev in free:
                ev.record(main)
            for i in range(n_e2e):
                stage(i & 1)
                free[i & 1].record(copy_stream)
            main.wait_stream(copy_stream)
        ms_copy, _ = timed(copies_only, 1)
        print(f"e2e probe: copies alone {ms_copy / n_e2e:.3f} ms/step, e2e {ms_e2e:.3f} ms/step, resident {ms_step:.3f} ms/step",
              file=sys.stderr)

    props = B * 3 * wl["depths"] * (world if weak else 1)  # whole job, all ranks
    value = props / (ms_step * 1e-3)
    e2e = props / (ms_e2e * 1e-3)
    h2d = phase_h.numel() * 4 + sum(t.numel() * 4 for t in targets_h)

    # ---- roofline of the dominant kernel on this rank ----
    peak, peak_src = peaks()
    seg_depths = [s.n_depth for s in stack.segments]  # per-colour segments: the byte model is per (sample, colour) group
    ab = algorithmic_bytes(wl, seg_depths, fused=kn[3] > 0)
    names = ["row_forward_kernel", "column_kernel", "row_inverse_kernel", "row_inverse_forward_fused_kernel"]
    keys = ["k1", "k2", "k3", "k4"]
    per_kernel = {}
    for i in range(4):
        if kn[i]:
            gbs = ab[keys[i]] * args.steps / (kms[i] * 1e-3) / 1e9
            per_kernel[names[i]] = {"ms_per_step": kms[i] / args.steps, "launches_per_step": kn[i] / args.steps,
                                    "algorithmic_GBps": gbs, "frac": gbs / peak}
    dom = max(range(4), key=lambda i: kms[i])
    dom_bytes_per_launch = ab[keys[dom]] * args.steps / max(kn[dom], 1)
    dom_ms_per_launch = kms[dom] / max(kn[dom], 1)
    achieved = dom_bytes_per_launch / (dom_ms_per_launch * 1e-3) / 1e9
    traffic, fp32_busy = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.isfile(tpath) and single:
        with open(tpath) as f:
            rec = json.load(f).get(args.workload, {}).get(names[dom], {})
        traffic, fp32_busy = rec.get("traffic_bytes_per_launch"), rec.get("fp32_pipe_busy_frac")
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r01_final.md (r01t)"
                if traffic else None,
                "peak_source": peak_src,
                "fp32_pipe_busy_frac": fp32_busy,  # ncu, same capture: the pipe that actually bounds this kernel
                "bytes_per_launch": dom_bytes_per_launch, "ms_per_launch": dom_ms_per_launch,
                "step_frac_of_hbm_floor": (ab["total"] / (ms_step * 1e-3) / 1e9) / peak,
                "note": "the column kernel is bound on chip (FP32 pipe / latency at 18 warps per SM, DESIGN.md 3.4), "
                        "not by HBM: 9 column transforms per strip read",
                "per_kernel": per_kernel}

    if rank == 0:
        line = {
            "metric": "rgb_depth_plane_propagations_per_s_fwd_bwd", "value": value, "unit": "propagations/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak" if (weak or world == 1) else "strong", "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl["name"], "key": args.workload, "propagations_per_step": props,
                       "sharding": (f"{world} rank(s) x 1 hologram x {stack.local_planes()} (colour,depth) planes, "
                                    "no data-path collective" if (weak or world == 1) else
                                    f"{world} rank(s) x {stack.local_planes()} (colour,depth) planes of one hologram"),
                       "l2": "working set per step >> 126 MB L2 (no flush needed)", "numa": numa},
            "roofline": roofline,
            "e2e": {"value": e2e, "unit": "propagations/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
            "gpu_launches": int(l1 - l0),
            "clocks": clk,
            "loss": float(loss),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
