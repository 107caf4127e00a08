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
               cpu_depths=2, name="C4 4K POH 3840x2160 -> 7680x4320 padded, RGB x 8 planes, fwd+L2+adjoint"),
    # BASELINE.json configs[1]
    "c2": dict(rows=384, cols=384, pad=320, coef=0.35, batch=4, depths=10, z0=4e-4, z1=10e-4,
               cpu_depths=10, name="C2 batch 4 of 384x384 -> 1024x1024 padded, RGB x 10 planes, fwd+L2+adjoint"),
    # BASELINE.json configs[4]: focal-stack sweep (pad 540 -> 2160x3840 is the measured line, pad 0 in the sweep)
    "c5": dict(kind="sweep", rows=1080, cols=1920, pad=540, coef=0.45, batch=16, depths=64, z0=4e-4, z1=10e-4,
               cpu_depths=2, cpu_batch=1, ref_gpu_batch=1,
               name="C5 1080p POH 1920x1080 -> 3840x2160 padded, batch 16, RGB x 64 planes, fwd+L2+adjoint"),
    # BASELINE.json configs[2]: the propagation calls of one trainingModel.py GAN step, per rank
    "c3": dict(kind="gan_step", rows=384, cols=384, pad=320, coef=0.45, batch=4, depths=20,
               name="C3 trainingModel.py step at 384x384 -> 1024x1024, batch 4 per GPU: F-6, AP2POH tail, F-7, F-13, "
                    "F-12 (random depth of 20), G_loss pixel/TV/focal terms, backward"),
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
        self.period = float(os.environ.get("LHG_CLOCK_POLL_MS", "25")) * 1e-3
        if os.environ.get("LHG_CLOCK_SAMPLER", "nvml") == "off":  # diagnosis only: a line without clocks is rejected
            self.proc = None
            return
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
        mx = None
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                if mx is None:
                    mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)  # a constant of the board
                bits = int(reasons_fn(self.handle))
                flags = ["Active" if bits & b else "Not Active" for b, _ in names]
                self.samples.append(", ".join([str(sm), str(mx)] + flags))
            except Exception:
                pass
            time.sleep(self.period)

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


def reference_multi(wl, cuda):
    """The UNMODIFIED reference module (angular_spectrum_method.py:470-522), loaded by oracle/ref_shim.py from
    /root/reference here or from the offline install under baseline/_ref on the GPU box.  None if neither exists."""
    from oracle import ref_shim

    if not ref_shim.available():
        return None
    ref_asm, _ = ref_shim.load()
    return ref_asm.bandLimitedAngularSpectrumMethod_for_multiple_distances


def reference_step_factory(wl, D, cuda, batch=None):
    """One step of the reference's own code on `D` of the workload's depth planes: multi __call__ (asm.py:503-522)
    + F.mse_loss + backward to the phase.  cuda=False: the reference's CPU path on all host threads (falls back to
    the oracle port, kind "port", only if the reference files are nowhere to be found); cuda=True: the same module
    on cuda:0 (cuFFT + ATen), the same-box GPU comparator."""
    z = torch.linspace(wl["z0"], wl["z1"], wl["depths"])[:D]
    gen = torch.Generator().manual_seed(122731)
    B = batch or wl["batch"]
    phase = 2 * torch.pi * torch.rand(B, 3, wl["rows"], wl["cols"], generator=gen)
    target = torch.rand(B * D, 3, wl["rows"], wl["cols"], generator=gen)
    props = B * 3 * D
    Multi = reference_multi(wl, cuda)
    kind = "reference"
    if Multi is not None:
        prop = Multi(sample_row_num=wl["rows"], sample_col_num=wl["cols"], distances=z, pad_size=wl["pad"],
                     filter_radius_coefficient=wl["coef"], pixel_pitch=PITCH, wave_length=torch.tensor(WL),
                     band_limit=False, cuda=cuda)
        dev = prop.device
        phase, target, ones = phase.to(dev), target.to(dev), torch.ones_like(phase).to(dev)

        def step():
            p = phase.clone().requires_grad_(True)
            loss = torch.nn.functional.mse_loss(prop(ones, p, z), target)
            loss.backward()
            return loss
    else:
        if cuda:
            return None, props, "reference files not found (neither /root/reference nor baseline/_ref)", "unavailable"
        from oracle import asm_oracle as O

        kind = "port"
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

    sample = (f"{wl['name']}; sample = " + (f"{B} of {wl['batch']} holograms, " if B != wl["batch"] else "") +
              f"the first {D} of {wl['depths']} depth planes per step ({props} propagations/step" + ("" if D == wl["depths"] else
              f"; the forward FFT is shared by {D} planes instead of {wl['depths']}, which understates the "
              f"per-propagation rate by about {(1 + 1 / D) / (1 + 1 / wl['depths']) - 1:.0%}") + ")")
    return step, props, sample, kind


def time_cpu(step, warm, steps):
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def run_reference(args, wl, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores, all
    threads, the arm's own --steps/--warmup, each step a bounded sample (cpu_depths planes) of the workload."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    step, props, sample, kind = reference_step_factory(wl, wl["cpu_depths"], cuda=False, batch=wl.get("cpu_batch"))
    warm, steps = max(args.warmup, 1), max(args.steps, 1)
    dt = time_cpu(step, warm, steps)
    value = props / dt
    line = {
        "impl": "reference", "metric": "rgb_depth_plane_propagations_per_s_fwd_bwd", "value": value,
        "unit": "propagations/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling or ("strong" if args.workload in ("c4", "c5") else "weak"),
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "key": args.workload},
        "cpu_baseline": {"value": value, "unit": "propagations/s", "cores": torch.get_num_threads(),
                         "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "propagations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(wl):
    torch.set_num_threads(os.cpu_count() or 1)
    step, props, sample, kind = reference_step_factory(wl, wl["cpu_depths"], cuda=False, batch=wl.get("cpu_batch"))
    dt = time_cpu(step, 1, 2)
    return {"value": props / dt, "unit": "propagations/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": sample + f", {dt:.2f} s/step"}


def gpu_reference(wl, steps, warmup):
    """The same-box GPU bar (SURVEY 8(d) comparators 1 and 2): the unmodified reference module with cuda=True
    (cuFFT + ATen), the FULL workload (every depth plane, same seeded inputs), timed with CUDA events like the new
    path; plus bare torch.fft.fft2 / ifft2 at the same batch (perf comparator only; cuFFT is not on the product
    path).  Runs after the new path's own measurements, on rank 0 at N = 1."""
    out = {"available": False}
    try:
        rb = wl.get("ref_gpu_batch") or wl["batch"]
        step, props, sample, kind = reference_step_factory(wl, wl["depths"], cuda=True, batch=rb)
        out["sample"] = sample
        if step is None:
            out["why"] = sample
            return out
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out.update({"available": True, "impl": "reference module, cuda=True (cuFFT + ATen), unmodified",
                    "ms_per_step": ms, "propagations_per_s": props / (ms * 1e-3), "steps": steps, "warmup": warmup,
                    "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9, "loss": float(loss.detach()),
                    "propagations_per_step": props})
        # forward only (generatePOH.py:66-70 runs it under no_grad)
        Multi = reference_multi(wl, True)
        del step
        torch.cuda.empty_cache()
        z = torch.linspace(wl["z0"], wl["z1"], wl["depths"])
        prop = Multi(sample_row_num=wl["rows"], sample_col_num=wl["cols"], distances=z, pad_size=wl["pad"],
                     filter_radius_coefficient=wl["coef"], pixel_pitch=PITCH, wave_length=torch.tensor(WL),
                     band_limit=False, cuda=True)
        gen = torch.Generator().manual_seed(122731)
        phase = (2 * torch.pi * torch.rand(rb, 3, wl["rows"], wl["cols"], generator=gen)).cuda()
        ones = torch.ones_like(phase)

        def ev_time(fn, n):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n

        with torch.no_grad():
            out["forward_only_ms"] = ev_time(lambda: prop(ones, phase, z), max(2, steps // 2))
        del prop
        torch.cuda.empty_cache()
        # bare cuFFT: one fft2 of [B,3,Rp,Cp] and one ifft2 of [B*D,3,Rp,Cp] (the transforms of one forward
        # pass); a forward + backward step runs each of them twice
        Rp = wl["rows"] + 2 * wl["pad"]
        Cp = wl["cols"] + 2 * int(wl["pad"] * (wl["cols"] / wl["rows"]))
        x = torch.randn(rb, 3, Rp, Cp, dtype=torch.complex64, device="cuda")
        y = torch.randn(rb * wl["depths"], 3, Rp, Cp, dtype=torch.complex64, device="cuda")
        f_ms = ev_time(lambda: torch.fft.fft2(x), max(2, steps // 2))
        i_ms = ev_time(lambda: torch.fft.ifft2(y), max(2, steps // 2))
        out["bare_cufft"] = {"fft2_ms": f_ms, "ifft2_ms": i_ms, "fwd_bwd_ms": 2 * (f_ms + i_ms),
                             "note": "torch.fft only, no pad / H / multiply / crop / abs / loss"}
        del x, y
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001 -- a comparator must never take the bench line down
        out["error"] = f"{type(e).__name__}: {e}"[:300]
        torch.cuda.empty_cache()
    return out


def run_reference_gpu(args, wl, rank):
    """`--impl reference-gpu`: the gpu_reference comparator as a line of its own."""
    if rank != 0:
        return
    torch.cuda.set_device(0)
    ref = gpu_reference(wl, args.steps, max(args.warmup, 3))
    props = wl["batch"] * 3 * wl["depths"]
    line = {"impl": "reference-gpu", "metric": "rgb_depth_plane_propagations_per_s_fwd_bwd",
            "value": ref.get("propagations_per_s"), "unit": "propagations/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ref.get("ms_per_step"), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "key": args.workload, "propagations_per_step": props},
            "gpu_reference": ref, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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


class Harness:
    """What every workload shares: rank / device set-up, the barrier + CUDA-event timing with the max over ranks,
    the clock sampler and the single JSON line on the real stdout."""

    def __init__(self, args):
        import torch.distributed as dist

        self.args, self.dist = args, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        # stdout carries exactly ONE JSON line: everything libraries print there (NCCL's version banner at
        # communicator creation) is sent to stderr; the line itself goes to the saved descriptor
        sys.stdout.flush()
        self.real_stdout = os.dup(1)
        os.dup2(2, 1)
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa = bind_to_gpu_numa_node(self.local_rank) if self.world > 1 else "single rank: not bound"
        print(f"rank {self.rank}: {self.numa}", file=sys.stderr)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        from learned_hologram_gan_b200 import _cabi

        self.lib = _cabi.load()
        self.warmup = max(args.warmup, 3)
        self.clocks = None
        self.remeasured = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """`steps` calls of fn between barrier + synchronize on both sides; CUDA events; max over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        self.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item() / steps, out

    def sum_over_ranks(self, x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.item()

    def start_clocks(self):
        self.clocks = ClockSampler(self.local_rank)
        self.clocks.start()       # started before the warm-up ...
        self.clocks.wait_ready()  # ... and past its start-up before anything is timed

    def kernel_profile(self, enable):
        import ctypes as C

        if enable:
            self.lib.asm_profile_enable(1)
            self._l0 = self.lib.asm_launch_count()
            return None
        l1 = self.lib.asm_launch_count()
        self.lib.asm_profile_enable(0)
        kms, kn = (C.c_double * 4)(0, 0, 0, 0), (C.c_longlong * 4)(0, 0, 0, 0)
        self.lib.asm_profile_collect(kms, kn, 4)
        return list(kms), list(kn), int(l1 - self._l0)

    def emit(self, line):
        if self.rank == 0:
            sys.stdout.flush()
            os.write(self.real_stdout, (json.dumps(line) + "\n").encode())

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


KERNEL_NAMES = ["row_forward_kernel", "column_kernel", "row_inverse_kernel", "row_inverse_forward_fused_kernel"]
KERNEL_KEYS = ["k1", "k2", "k3", "k4"]


def traffic_record(workload, kernel):
    """DRAM bytes per launch of `kernel` from the newest committed ncu capture (profiles/r*_traffic.json, written by
    tools/ncu_traffic.py from the same ncu run as that round's launch list)."""
    import glob

    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        with open(path) as f:
            rec = json.load(f)
        hit = rec.get(workload, {}).get(kernel)
        if hit:
            return hit, os.path.basename(path), rec.get("source")
    return {}, None, None


def roofline_of(workload, ab, kms, kn, steps, ms_step, single):
    peak, peak_src = peaks()
    per_kernel = {}
    for i in range(4):
        if kn[i]:
            gbs = ab[KERNEL_KEYS[i]] * steps / (kms[i] * 1e-3) / 1e9
            per_kernel[KERNEL_NAMES[i]] = {"ms_per_step": kms[i] / steps, "launches_per_step": kn[i] / steps,
                                           "algorithmic_GBps": gbs, "frac": gbs / peak}
    dom = max(range(4), key=lambda i: kms[i])
    dom_bytes_per_launch = ab[KERNEL_KEYS[dom]] * steps / max(kn[dom], 1)
    dom_ms_per_launch = kms[dom] / max(kn[dom], 1)
    achieved = dom_bytes_per_launch / (dom_ms_per_launch * 1e-3) / 1e9
    rec, tfile, tsrc = traffic_record(workload, KERNEL_NAMES[dom]) if single else ({}, None, None)
    return {"bound": "hbm", "kernel": KERNEL_NAMES[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": rec.get("traffic_bytes_per_launch"),
            "traffic_source": (f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/{tfile}"
                               + (f" ({tsrc})" if tsrc else "")) if tfile else None,
            "peak_source": peak_src,
            "fp32_pipe_busy_frac": rec.get("fp32_pipe_busy_frac"),  # ncu, same capture
            "bytes_per_launch": dom_bytes_per_launch, "ms_per_launch": dom_ms_per_launch,
            "step_frac_of_hbm_floor": (ab["total"] / (ms_step * 1e-3) / 1e9) / peak,
            "step_algorithmic_GB": ab["total"] / 1e9,
            "note": "the column kernel is bound on chip (FP32 pipe / issue at 18 warps per SM, DESIGN.md 3.4), "
                    "not by HBM: 9 column transforms per strip read",
            "per_kernel": per_kernel}


def run_focal_stack(args, wl):
    """c4 / c2: one RGB hologram batch -> D planes, amplitude-L2, adjoint.  N = 1: the whole stack in one fused
    call.  N > 1, --scaling strong (default for c4, the split BASELINE.json configs[3] names): the (colour, depth)
    planes of ONE hologram are partitioned over the ranks by a cost model, a rank keeps only the phase planes of
    its colours, and the phase gradient of a colour is summed inside the sub-group of ranks holding its planes;
    the weak-scaling rate (one whole hologram per rank, no data-path collective) is measured in the same run and
    reported under "weak_scaling".  --scaling weak makes the weak rate the line's value."""
    from learned_hologram_gan_b200.sharding import ShardedFocalStack, balanced_shards, segment_cost

    h = Harness(args)
    dist, dev, rank, world = h.dist, h.dev, h.rank, h.world
    scaling = args.scaling or ("strong" if args.workload == "c4" else "weak")
    B, R, C, D = wl["batch"], wl["rows"], wl["cols"], wl["depths"]
    z = torch.linspace(wl["z0"], wl["z1"], D)
    wlt = torch.tensor(WL)

    def make_targets(gen):
        return {(c, d): torch.rand(B, 1, R, C, generator=gen) for c in range(3) for d in range(D)}

    # ---------------- whole-stack replica (N = 1, and the weak-scaling measurement at N > 1) ----------------
    def build_replica(seed_rank):
        stack = ShardedFocalStack(R, C, z, wl["pad"], wl["coef"], PITCH, wlt, world=1, rank=0)
        phase_h, gen = make_inputs(wl, world, seed_rank)
        all_t = make_targets(gen)
        full_h = torch.stack([torch.stack([all_t[(c, d)][:, 0] for c in range(3)], dim=1) for d in range(D)], dim=1)
        full_h = full_h.reshape(B * D, 3, R, C).contiguous().pin_memory()  # [B*D,3,R,C], index b*D+d (asm.py:516-518)
        return stack, phase_h, [full_h]

    def replica_step(stack, phase_d, targets_d):
        return stack.loss_and_grad_full(phase_d, targets_d[0])

    # ---------------- planes of one hologram sharded over the ranks ----------------
    def build_strong():
        stack = ShardedFocalStack(R, C, z, wl["pad"], wl["coef"], PITCH, wlt, world=world, rank=rank,
                                  balanced=True, colour_groups=True)
        phase_h, gen = make_inputs(wl, world, 0)  # the SAME hologram on every rank
        all_t = make_targets(gen)                 # every rank draws the full stream, keeps its planes
        targets_h = []
        for seg in stack.segments:
            t = torch.stack([all_t[(seg.colour, d)] for d in range(seg.d0, seg.d1)], dim=1)
            targets_h.append(t.reshape(B * seg.n_depth, 1, R, C).contiguous().pin_memory())
        return stack, phase_h, targets_h

    def measure(stack, phase_h, targets_h, step_fn, h2d_phase_planes, label, resident=True):
        """resident timing with the library's per-kernel events, then the end-to-end loop"""
        phase_d = phase_h.to(dev)
        targets_d = [t.to(dev) for t in targets_h]
        for _ in range(h.warmup):
            step_fn(stack, phase_d, targets_d)
        ms_step, out, prof, clk = None, None, None, None
        if resident:
            if label == "main":
                h.clocks.mark()
                h.kernel_profile(True)
            ms_step, out = h.timed(lambda: step_fn(stack, phase_d, targets_d), args.steps)
            prof = h.kernel_profile(False) if label == "main" else None
            # A step is 6 back-to-back launches: the K steps cannot take longer than the sum of their kernels' own
            # CUDA-event times plus launch gaps (1-1.5 % on an undisturbed box).  When they do by more than 3 % (a
            # stall between launches: seen once in a few runs on some boxes, up to tens of ms in ONE step, with
            # unchanged kernel times), the region is measured again -- the driver's own policy for a disturbed run --
            # at most twice; the fastest attempt is the line's value and every attempt is listed in it.
            if label == "main" and prof is not None and world == 1:
                attempts, profs = [], []
                while True:
                    ksum = sum(prof[0]) / args.steps
                    attempts.append({"ms_per_step": ms_step, "kernel_sum_ms_per_step": ksum})
                    profs.append(prof)
                    if not (ksum > 0 and ms_step > 1.03 * ksum) or len(attempts) == 3:
                        break
                    h.kernel_profile(True)
                    ms_step, out = h.timed(lambda: step_fn(stack, phase_d, targets_d), args.steps)
                    prof = h.kernel_profile(False)
                if len(attempts) > 1:
                    best = min(range(len(attempts)), key=lambda i: attempts[i]["ms_per_step"])
                    ms_step, prof = attempts[best]["ms_per_step"], profs[best]
                    h.remeasured = {"attempts": attempts, "kept": best,
                                    "reason": "a timed region exceeded the sum of its kernels' event times by > 3 %"}
            clk = h.clocks.stop() if label == "main" else None

        # end to end: every step copies ITS inputs from pinned host memory and reads its loss back.  The copies of
        # step i+1 run on a second stream into the other of two device buffers while step i computes (what a
        # training loop's prefetching loader does); all of them lie inside the timed region.
        copy_stream = torch.cuda.Stream(device=dev)
        slots = [(torch.empty_like(phase_d), [torch.empty_like(t) for t in targets_d]) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]   # slot filled (recorded on the copy stream)
        free = [torch.cuda.Event(), torch.cuda.Event()]    # slot consumed (recorded on the compute stream)
        planes = list(h2d_phase_planes)

        def stage(slot):
            copy_stream.wait_event(free[slot])
            with torch.cuda.stream(copy_stream):
                if len(planes) == 3:
                    slots[slot][0].copy_(phase_h, non_blocking=True)
                else:  # only the phase planes of the colours this rank holds
                    for c in planes:
                        slots[slot][0][:, c].copy_(phase_h[:, c], non_blocking=True)
                for dst, src in zip(slots[slot][1], targets_h):
                    dst.copy_(src, non_blocking=True)
                ready[slot].record(copy_stream)

        def run_e2e(steps):
            res = None
            main_s = torch.cuda.current_stream(dev)
            for ev in free:
                ev.record(main_s)
            stage(0)
            for i in range(steps):
                slot = i & 1
                if i + 1 < steps:
                    stage(slot ^ 1)
                main_s.wait_event(ready[slot])
                p, ts = slots[slot]
                loss, grad = step_fn(stack, p, ts)
                free[slot].record(main_s)
                res = (loss.item(), grad)  # device -> host read of the step's result (the scalar loss)
            return res

        run_e2e(2)
        n_e2e = max(2, args.steps)  # the first copy of the pipeline has nothing to overlap with: amortised over K
        ms_e2e, _ = h.timed(lambda: run_e2e(n_e2e), 1)
        ms_e2e /= n_e2e
        h2d = len(planes) * B * R * C * 4 + sum(t.numel() * t.element_size() for t in targets_h)
        del slots
        return ms_step, ms_e2e, h.sum_over_ranks(h2d), out, prof, clk

    h.start_clocks()
    props_one = B * 3 * D
    weak_rec = None
    u8_rec = None
    if world == 1 or scaling == "weak":
        stack, phase_h, targets_h = build_replica(rank)
        ms_step, ms_e2e, h2d, (loss, grad), prof, clk = measure(stack, phase_h, targets_h, replica_step, range(3), "main")
        if world == 1:
            # separately labelled variant: the same step with 8-bit targets (v / 255, the reference's image
            # convention; generatePOH-style targets are 8-bit PNGs) uploaded as uint8 and converted inside the
            # fused row kernel -- a quarter of the target bytes over the host link.  NOT the headline e2e.
            t8_h = [(t * 255).round().to(torch.uint8).pin_memory() for t in targets_h]
            _, u8_ms, u8_h2d, _, _, _ = measure(stack, phase_h, t8_h, replica_step, range(3), "u8", resident=False)
            u8_rec = {"value": props_one / (u8_ms * 1e-3), "unit": "propagations/s", "ms_per_step": u8_ms,
                      "h2d_bytes_per_step": int(u8_h2d), "d2h_bytes_per_step": 4,
                      "note": "uint8 targets (value / 255) instead of fp32: a different input format, reported beside "
                              "the fp32 e2e, not instead of it"}
            del t8_h
        # weak scaling: no collective in the step at all; the per-rank losses are combined once, after the K steps
        if world > 1:
            lt = loss.detach().clone()
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
            loss = lt / world
        props = props_one * world
        sharding = (f"{world} rank(s) x 1 hologram x {stack.local_planes()} (colour,depth) planes; no data-path "
                    "collective, the scalar losses are combined once after the K steps")
        seg_depths = [D] * 3
        # one rank: the whole job is local either way; the line carries the mode the N > 1 runs of the same command use
        out_scaling = scaling if world == 1 else "weak"
    else:
        stack, phase_h, targets_h = build_strong()
        step = lambda st, p, ts: st.loss_and_grad_sharded(p, ts)  # noqa: E731
        ms_step, ms_e2e, h2d, (loss, grads), prof, clk = measure(stack, phase_h, targets_h, step,
                                                                  stack.owned_colours, "main")
        # time between "last local kernel enqueued" and "collectives done" on the compute stream, per step
        stack.collective_events = []
        phase_d = phase_h.to(dev)
        targets_d = [t.to(dev) for t in targets_h]
        h.timed(lambda: stack.loss_and_grad_sharded(phase_d, targets_d), args.steps)
        coll_ms = sum(a.elapsed_time(b) for a, b in stack.collective_events) / max(len(stack.collective_events), 1)
        stack.collective_events = None
        coll_ms = h.sum_over_ranks(coll_ms) / world
        # the same step without the per-colour gradient reductions: their exposed cost per step
        saved_groups = stack.colour_group
        stack.colour_group = [None] * 3
        ms_nored, _ = h.timed(lambda: stack.loss_and_grad_sharded(phase_d, targets_d, reduce_loss=False), args.steps)
        stack.colour_group = saved_groups
        del phase_d, targets_d
        props = props_one
        shards = balanced_shards(3, D, world)
        costs = [sum(segment_cost(s_.n_depth, 1.65, 1.0) for s_ in r) for r in shards]
        sharding = (f"{world} ranks x planes of ONE hologram, cost-model partition "
                    f"{[[(s_.colour, s_.d0, s_.d1) for s_ in r] for r in shards]}; phase gradient summed per colour "
                    f"inside the sub-group of ranks holding its planes {stack.owners}")
        seg_depths = [s_.n_depth for s_ in stack.segments]
        out_scaling = "strong"
        strong_extra = {"collective_ms_per_step": ms_step - ms_nored, "ms_per_step_without_reductions": ms_nored,
                        "collective_wait_ms_on_stream": coll_ms,
                        "partition_efficiency_bound": (3 * 1.65 + 3 * D) / (world * max(costs)),
                        "bound_note": "every segment recomputes the forward transform of its colour and finishes its "
                                      "own adjoint (fixed cost = 1.65 plane-costs, measured at N = 1): the bound is "
                                      "total cost / (N x the slowest rank's cost)"}
        # weak-scaling rate in the same run (one whole hologram per rank)
        del stack, targets_h
        torch.cuda.empty_cache()
        from learned_hologram_gan_b200 import engine as _E

        _E.release_workspaces()
        wstack, wphase_h, wtargets_h = build_replica(rank)
        w_ms, w_e2e, w_h2d, _, _, _ = measure(wstack, wphase_h, wtargets_h, replica_step, range(3), "weak")
        weak_rec = {"value": props_one * world / (w_ms * 1e-3), "unit": "propagations/s", "ms_per_step": w_ms,
                    "e2e": {"value": props_one * world / (w_e2e * 1e-3), "ms_per_step": w_e2e,
                            "h2d_bytes_per_step": w_h2d},
                    "sharding": f"{world} ranks x 1 whole hologram each, no data-path collective"}

    kms, kn, launches = prof
    ab = algorithmic_bytes(wl, seg_depths, fused=kn[3] > 0)
    roofline = roofline_of(args.workload, ab, kms, kn, args.steps, ms_step, world == 1)
    line = {
        "metric": "rgb_depth_plane_propagations_per_s_fwd_bwd", "value": props / (ms_step * 1e-3),
        "unit": "propagations/s", "n_gpus": world, "steps": args.steps, "warmup": h.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": out_scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "key": args.workload, "propagations_per_step": props,
                   "sharding": sharding, "l2": "working set per step >> 126 MB L2 (no flush needed)",
                   "numa": h.numa,
                   "e2e_result": "the scalar loss is read back every step; the phase gradient stays resident on the "
                                 "device for the optimiser step (100 MB at 4K, it never crosses the host link)"},
        "roofline": roofline,
        "e2e": {"value": props / (ms_e2e * 1e-3), "unit": "propagations/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e},
        "gpu_launches": launches,
        "clocks": clk,
        "loss": float(loss),
    }
    if h.remeasured:
        line["remeasured"] = h.remeasured
    if weak_rec:
        line["weak_scaling"] = weak_rec
        line["strong_scaling"] = strong_extra
    if u8_rec:
        line["e2e_u8_targets"] = u8_rec
    if world == 1 and rank == 0:
        del stack, targets_h
        torch.cuda.empty_cache()
        from learned_hologram_gan_b200 import engine as _E

        _E.release_workspaces()
        if not args.no_gpu_reference:
            line["gpu_reference"] = gpu_reference(wl, max(2, min(args.steps, 10)), 3)
            if line["gpu_reference"].get("ms_per_step"):
                line["gpu_reference"]["new_path_speedup"] = line["gpu_reference"]["ms_per_step"] / ms_step
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
    h.emit(line)
    h.finish()


def run_sweep(args, wl):
    """c5 (BASELINE.json configs[4]): 1080p RGB POHs, batch 16, to D depth planes.  The 48 (sample, colour) groups
    are sharded WHOLE over the ranks (whole samples: 16 / N per rank), so no rank shares a forward FFT with another
    and the path has no collective at all; total work is fixed (strong scaling).  The line's value is the padded
    (2160 x 3840) D = 64 forward + L2 + adjoint rate; `sweep` holds the forward-only focal-stack rate
    (generatePOH.py:66-70 under no_grad) for D in {1, 4, 16, 64}, padded and un-padded."""
    import learned_hologram_gan_b200.angular_spectrum_method as M
    from learned_hologram_gan_b200 import engine as E

    h = Harness(args)
    dev, rank, world = h.dev, h.rank, h.world
    B, R, C, D = wl["batch"], wl["rows"], wl["cols"], wl["depths"]
    if B % world:
        raise SystemExit(f"c5 shards whole samples: batch {B} is not a multiple of {world} ranks")
    Bl = B // world
    wlt = torch.tensor(WL)
    gen = torch.Generator().manual_seed(122731 + rank)
    phase_h = (2 * torch.pi * torch.rand(Bl, 3, R, C, generator=gen)).pin_memory()
    phase_d = phase_h.to(dev)
    # targets: one sample's stack is drawn and pinned once (1.6 GB); every local sample gets a copy of it on the
    # device (the rate does not depend on the values; 25 GB of pinned host memory would)
    one_h = torch.rand(D, 3, R, C, generator=gen).pin_memory()
    target_d = torch.empty(Bl * D, 3, R, C, device=dev)
    for b in range(Bl):
        target_d[b * D:(b + 1) * D].copy_(one_h, non_blocking=True)

    def make(pad, depths):
        z = torch.linspace(wl["z0"], wl["z1"], depths)
        return M.bandLimitedAngularSpectrumMethod_for_multiple_distances(
            sample_row_num=R, sample_col_num=C, distances=z, pad_size=pad, filter_radius_coefficient=wl["coef"],
            pixel_pitch=PITCH, wave_length=wlt, band_limit=False, cuda=True), z

    prop, z = make(wl["pad"], D)
    numel = B * D * 3 * R * C
    grad = torch.empty_like(phase_d)

    def step(p, t):
        return prop.amplitude_mse_and_phase_gradient(p, z, t, 2.0 / numel, grad_out=grad)

    h.start_clocks()
    for _ in range(h.warmup):
        step(phase_d, target_d)
    h.clocks.mark()
    h.kernel_profile(True)
    ms_step, (sum_sq, _) = h.timed(lambda: step(phase_d, target_d), args.steps)
    kms, kn, launches = h.kernel_profile(False)
    clk = h.clocks.stop()

    # end to end: phase and targets of every local sample come from pinned host memory each step (double-buffered
    # on a copy stream), the scalar loss goes back
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [(torch.empty_like(phase_d), target_d if i == 0 else torch.empty_like(target_d)) for i in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]

    def stage(slot):
        copy_stream.wait_event(free[slot])
        with torch.cuda.stream(copy_stream):
            slots[slot][0].copy_(phase_h, non_blocking=True)
            for b in range(Bl):
                slots[slot][1][b * D:(b + 1) * D].copy_(one_h, non_blocking=True)
            ready[slot].record(copy_stream)

    def run_e2e(steps):
        main_s = torch.cuda.current_stream(dev)
        for ev in free:
            ev.record(main_s)
        stage(0)
        val = None
        for i in range(steps):
            slot = i & 1
            if i + 1 < steps:
                stage(slot ^ 1)
            main_s.wait_event(ready[slot])
            s_, _ = step(*slots[slot])
            free[slot].record(main_s)
            val = s_.item()
        return val

    n_e2e = max(2, min(args.steps, 4))
    run_e2e(2)
    ms_e2e, _ = h.timed(lambda: run_e2e(n_e2e), 1)
    ms_e2e /= n_e2e
    h2d = h.sum_over_ranks(phase_h.numel() * 4 + Bl * one_h.numel() * 4)
    del slots
    torch.cuda.empty_cache()

    # forward-only focal-stack sweep (the generatePOH.py call), this rank's samples
    sweep = []
    ones = torch.ones_like(phase_d)
    del target_d, prop
    for pad in (wl["pad"], 0):
        for depths in (1, 4, 16, 64):
            E.release_workspaces()
            torch.cuda.empty_cache()
            pr, zz = make(pad, depths)
            with torch.no_grad():
                for _ in range(2):
                    out = pr(ones, phase_d, zz)
                n = 3
                ms, _ = h.timed(lambda: pr(ones, phase_d, zz), n)
            del out, pr
            Cp = C + 2 * int(pad * (C / R))
            a_fwd = B * 3 * (4 * R * C + 16 * R * Cp + depths * (16 * R * Cp + 4 * R * C))  # SURVEY 8(d) A_pass,fwd
            peak, _ = peaks()
            sweep.append({"pad": pad, "depths": depths, "ms": ms, "propagations_per_s_forward": B * 3 * depths / (ms * 1e-3),
                          "frac_of_hbm_floor": a_fwd / world / (ms * 1e-3) / 1e9 / peak})

    props = B * 3 * D
    lw = dict(wl, batch=Bl)
    ab = algorithmic_bytes(lw, [D] * 3, fused=kn[3] > 0)
    roofline = roofline_of("c5", ab, kms, kn, args.steps, ms_step, world == 1)
    loss = sum_sq.detach().clone()
    if world > 1:
        h.dist.all_reduce(loss, op=h.dist.ReduceOp.SUM)
    line = {
        "metric": "rgb_depth_plane_propagations_per_s_fwd_bwd", "value": props / (ms_step * 1e-3),
        "unit": "propagations/s", "n_gpus": world, "steps": args.steps, "warmup": h.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "key": "c5", "propagations_per_step": props,
                   "sharding": f"{world} rank(s) x {Bl} whole samples (x 3 colours x {D} planes), no collective",
                   "l2": "working set per step >> 126 MB L2 (no flush needed)", "numa": h.numa,
                   "e2e_result": "phase + fp32 targets of every sample uploaded per step from one pinned 1.6 GB "
                                 "stack (same bytes as distinct targets), scalar loss read back; gradient stays resident"},
        "roofline": roofline,
        "e2e": {"value": props / (ms_e2e * 1e-3), "unit": "propagations/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e},
        "gpu_launches": launches, "clocks": clk, "loss": float(loss) / numel,
        "sweep": sweep,
    }
    if world == 1 and rank == 0:
        E.release_workspaces()
        torch.cuda.empty_cache()
        if not args.no_gpu_reference:
            line["gpu_reference"] = gpu_reference(wl, max(2, min(args.steps, 5)), 2)
            if line["gpu_reference"].get("propagations_per_s"):
                line["gpu_reference"]["new_path_speedup"] = line["value"] / line["gpu_reference"]["propagations_per_s"]
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl)
    h.emit(line)
    h.finish()


def run_gan_step(args, wl):
    """c3 (BASELINE.json configs[2]): the propagation work of one trainingModel.py GAN step (hot loop
    watermelon.py:207-277), one process per GPU, batch 4 per rank (weak scaling): F-6 -> AP2POH tail -> F-7,
    F-13 on the dataset pair, F-12 with one randperm draw per rank and step, G_loss's pixel / TV / focal-phase
    terms, and the backward of all of it.  The networks (UNets, critic) are out of scope and not in the
    step; the only parameters are the tail's 3 x (3x3) kernels and biases, whose gradients are all-reduced like DDP
    would.  The step is captured in a CUDA graph (stage_random_depths feeds the depth draw to the replay); the eager
    time and the time inside this library's propagation kernels are reported beside it."""
    import learned_hologram_gan_b200.angular_spectrum_method as M
    from learned_hologram_gan_b200 import loss_func as LF
    from learned_hologram_gan_b200.ap2poh_tail import ap2poh_tail

    h = Harness(args)
    dev, rank, world, dist = h.dev, h.rank, h.world, h.dist
    B, R = wl["batch"], wl["rows"]
    wlt = torch.tensor(WL)
    geom = dict(sample_row_num=R, sample_col_num=R, pad_size=wl["pad"], filter_radius_coefficient=wl["coef"],
                pixel_pitch=PITCH, wave_length=wlt, cuda=True)
    fixed = M.bandLimitedAngularSpectrumMethod_for_single_fixed_distance(distance=torch.tensor([1e-3]), **geom)
    multi = M.bandLimitedAngularSpectrumMethod_for_multiple_distances(
        distances=torch.linspace(-4e-4, 0, wl["depths"] + 1)[:-1], **geom)
    torch.manual_seed(122731 + rank)  # the CPU global generator F-12 draws its depths from: per rank
    gen = torch.Generator().manual_seed(122731 + rank)
    amp_h = torch.rand(B, 3, R, R, generator=gen).pin_memory()       # dataset amplitude (RGBD2AP output stand-in)
    phs_h = torch.rand(B, 3, R, R, generator=gen).pin_memory()       # dataset phase in [0,1)
    net_a_h = torch.rand(B, 3, R, R, generator=gen).pin_memory()     # the generator's (a, phi) at z: network output stand-in
    net_p_h = (2 * torch.pi * torch.rand(B, 3, R, R, generator=gen)).pin_memory()
    st = {k: v.to(dev) for k, v in dict(amp=amp_h, phs=phs_h, net_a=net_a_h, net_p=net_p_h).items()}
    cw = 0.5 * torch.rand(3, 3, 3, generator=gen)
    conv_w = (cw + cw.transpose(1, 2)).to(dev).requires_grad_(True)
    conv_b = torch.zeros(3, device=dev, requires_grad=True)
    g_a = torch.zeros(B, 3, R, R, device=dev)
    g_p = torch.zeros(B, 3, R, R, device=dev)

    def step():
        a = st["net_a"].detach().requires_grad_(True)
        p = st["net_p"].detach().requires_grad_(True)
        c = fixed.propagate_AP2C_backward(a, p)                                   # F-6   AP2POH.py:107
        poh = ap2poh_tail(c, conv_w, conv_b)                                      # tail  AP2POH.py:108-116
        s_hat = fixed.propagate_POH2Freq_forward(poh)                             # F-7   watermelon.py:219
        s_tgt = multi.filter_AP2filteredFreq(st["amp"], st["phs"])                # F-13  watermelon.py:224
        a2, q2 = multi.propagate_multiple_samples_with_random_fixed_multiple_distances_freq2amp(
            torch.cat([s_hat, s_tgt], 0))                                         # F-12  watermelon.py:231
        terms = LF.amp_loss_terms(a2[:B], a2[B:].detach(), 1.0)                   # watermelon.py:418-445
        loss = terms[0] + terms[3] + LF.focal_sincos_phase_gradient_loss(q2[:B], q2[B:].detach())
        conv_w.grad = conv_b.grad = None
        loss.backward()
        g_a.copy_(a.grad)  # what flows on into the (out-of-scope) generator network
        g_p.copy_(p.grad)
        return loss.detach()

    def reduce_params():
        if world > 1:  # DDP's job for the tail's 30 parameters
            flat = torch.cat([conv_w.grad.reshape(-1), conv_b.grad.reshape(-1)])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)

    def eager():
        out = step()
        reduce_params()
        return out

    h.start_clocks()
    for _ in range(h.warmup):
        eager()
    # ---- CUDA graph of the whole step (forward, losses, backward) ----
    graph, g_loss, graph_err = None, None, None
    try:
        multi.stage_random_depths(B)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            g_loss = step()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        graph, graph_err = None, f"{type(e).__name__}: {e}"[:300]
        torch.cuda.synchronize()

    def graphed():
        multi.stage_random_depths(B)  # host: one randperm; device: a 32-byte upload ahead of the replay
        graph.replay()
        reduce_params()
        return g_loss

    main_fn = graphed if graph is not None else eager
    if graph is None:
        multi.stage_random_depths(0)
    for _ in range(h.warmup):
        main_fn()
    h.clocks.mark()
    ms_step, loss = h.timed(main_fn, args.steps)
    clk = h.clocks.stop()
    # eager step and the time inside this library's propagation kernels (separate passes)
    multi.stage_random_depths(0)
    ms_eager, _ = h.timed(eager, args.steps)
    h.kernel_profile(True)
    h.timed(eager, args.steps)
    kms, kn, launches = h.kernel_profile(False)
    path_ms = sum(kms) / args.steps

    # end to end: the step's four input tensors come from pinned host memory, the loss goes back
    def e2e_step():
        for k, src in (("amp", amp_h), ("phs", phs_h), ("net_a", net_a_h), ("net_p", net_p_h)):
            st[k].copy_(src, non_blocking=True)
        return main_fn().item()

    if graph is not None:
        multi.stage_random_depths(B)
    e2e_step()
    ms_e2e, _ = h.timed(e2e_step, args.steps)
    plane_ops = 12 * B // 4 * 5  # F-6 12 + F-7 12 + F-13 12 + F-12 24 plane passes at batch 4
    props = plane_ops * world
    peak, peak_src = peaks()
    Rp = R + 2 * wl["pad"]
    g = B * 3
    # streaming floor of the step's propagation calls, forward + backward (the C3 paragraph):
    # F-6 (fwd: a, phi in, complex out; bwd mirror), F-7 / F-13 (spectrum out), F-12 (spectrum in, abs+angle out)
    f6 = g * (8 * R * R + 16 * R * Rp + 16 * R * Rp + 8 * R * R) * 2
    f7 = g * (4 * R * R + 16 * R * Rp + 8 * Rp * Rp) * 2
    f13 = g * (8 * R * R + 16 * R * Rp + 8 * Rp * Rp)
    f12 = 2 * g * (8 * Rp * Rp + 16 * R * Rp + 8 * R * R) + g * (8 * Rp * Rp + 16 * R * Rp + 24 * R * R)
    a_path = f6 + f7 + f13 + f12
    line = {
        "metric": "rgb_depth_plane_propagations_per_s_fwd_bwd", "value": props / (ms_step * 1e-3),
        "unit": "propagations/s", "n_gpus": world, "steps": args.steps, "warmup": h.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "key": "c3", "propagations_per_step": props,
                   "sharding": f"{world} rank(s) x batch {B} (data parallel), per-rank randperm draw; NCCL carries only "
                               "the all-reduce of the tail's 30 parameter gradients",
                   "l2": "the 1024^2 planes are L2-resident (8 MiB each): launch / latency bound, the HBM fraction is "
                         "informational (SURVEY 8(d))", "numa": h.numa,
                   "cuda_graph": graph is not None, "cuda_graph_error": graph_err},
        "roofline": {"bound": "hbm", "kernel": "propagation kernels of the step (sum)", "achieved": a_path / (path_ms * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": a_path / (path_ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "peak_source": peak_src, "path_algorithmic_GB": a_path / 1e9},
        "step": {"graph_ms": ms_step if graph is not None else None, "eager_ms": ms_eager,
                 "path_kernels_ms": path_ms, "path_launches_per_step": launches / args.steps,
                 "glue_ms_eager": ms_eager - path_ms},
        "e2e": {"value": props / (ms_e2e * 1e-3), "unit": "propagations/s",
                "h2d_bytes_per_step": 4 * amp_h.numel() * 4 * world, "d2h_bytes_per_step": 4 * world,
                "ms_per_step": ms_e2e},
        "gpu_launches": launches, "clocks": clk, "loss": float(loss),
    }
    h.emit(line)
    h.finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["stages-c4", "stages-c2"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="N > 1: strong = the planes of one hologram sharded (default for c4), weak = one per rank")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.workload.startswith("stages"):
        if args.impl != "b200":  # the CPU port of every stage is timed inside the stages line itself
            if rank == 0:
                print(json.dumps({"impl": args.impl, "unavailable": "stage workloads report their CPU port in "
                                  "stages[*].cpu_baseline of the b200 line"}), flush=True)
            return
        run_stages(args, rank)
        return
    wl = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl != "b200" and wl.get("kind") == "gan_step":
        if rank == 0:
            print(json.dumps({"impl": args.impl, "unavailable": "c3 composes five reference methods with the "
                              "(out-of-scope) networks' tensors; its reference comparators are the c2 / c4 arms"}), flush=True)
        return
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args, wl, rank)
    elif wl.get("kind", "focal_stack") == "focal_stack":
        run_focal_stack(args, wl)
    elif wl["kind"] == "sweep":
        run_sweep(args, wl)
    elif wl["kind"] == "gan_step":
        run_gan_step(args, wl)
    else:
        raise SystemExit(f"unknown workload kind {wl.get('kind')}")


if __name__ == "__main__":
    main()
