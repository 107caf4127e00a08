"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run here (the container that has /root/reference):  python tests/golden/make_golden.py
It imports the reference's angular_spectrum_method through oracle/ref_shim.py, runs every
method on the hot path on small seeded inputs (CPU, fp32) and stores inputs, outputs and
reference-autograd gradients as .npz.  It also copies the reference's only known-answer
data (output/test_output/terminalTest/poh.pt + 0..9.png, README.md:123-132) as data files.
Nothing here is reference source code.
"""

from __future__ import annotations

import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_shim  # noqa: E402

CASES = {
    # name: rows, cols, pad, coef  (all padded sizes 2/3/5-smooth)
    "sq48p8": dict(rows=48, cols=48, pad=8, coef=0.45),  # -> 64 x 64
    "r40c60p10": dict(rows=40, cols=60, pad=10, coef=0.35),  # -> 60 x 90
    "r24c36p0": dict(rows=24, cols=36, pad=0, coef=0.5),  # no padding
}
WAVELENGTHS = torch.tensor([638e-9, 520e-9, 450e-9])
PITCH = 3.74e-6


def npify(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def run_case(asm, name, rows, cols, pad, coef):
    gen = torch.Generator().manual_seed(122731)
    kw = dict(
        sample_row_num=rows,
        sample_col_num=cols,
        pad_size=pad,
        filter_radius_coefficient=coef,
        pixel_pitch=PITCH,
        wave_length=WAVELENGTHS,
        band_limit=False,
        cuda=False,
    )
    B, D = 2, 3
    z_multi = torch.linspace(4e-4, 10e-4, D)
    z_stack = torch.linspace(-4e-4, 0, 6)[:-1]
    z_fixed = torch.tensor([1e-3])
    z_base = torch.linspace(-1e-3, 2.5e-3, 4)

    base = asm.bandLimitedAngularSpectrumMethod(**kw)
    fixed = asm.bandLimitedAngularSpectrumMethod_for_single_fixed_distance(distance=z_fixed, **kw)
    multi = asm.bandLimitedAngularSpectrumMethod_for_multiple_distances(distances=z_stack, **kw)
    Rp, Cp = base.samplingRowNum, base.samplingColNum

    phase = 2 * torch.pi * torch.rand(B, 3, rows, cols, generator=gen)
    amp = torch.rand(B, 3, rows, cols, generator=gen)
    phase3 = 2 * torch.pi * torch.rand(3, rows, cols, generator=gen)
    phs01 = torch.rand(B, 3, rows, cols, generator=gen)
    out = dict(
        rows=rows, cols=cols, pad=pad, coef=coef, pitch=PITCH, wavelengths=WAVELENGTHS,
        z_multi=z_multi, z_stack=z_stack, z_fixed=z_fixed, z_base=z_base,
        phase=phase, amp=amp, phase3=phase3, phs01=phs01,
        mask=base.diffraction_limited_mask, w_grid=base.w_grid,
        H_multi=multi.H, H_fixed=fixed.H,
        soft_grid=fixed.circular_frequency_mask_differentiable_grid,
        soft_mask_040=fixed.generate_circular_frequency_mask_differentiable(torch.tensor(0.4)),
        band_mask=base.generate_band_limited_mask(z_base),
    )

    # F-1 base call: 3-D input broadcast over 4 distances, and the paired (dim0 == D) form
    out["f1_bcast"] = base(torch.ones_like(phase3), phase3, z_base)
    ph4 = 2 * torch.pi * torch.rand(4, 3, rows, cols, generator=gen)
    am4 = torch.rand(4, 3, rows, cols, generator=gen)
    out["f1_ph4"], out["f1_am4"] = ph4, am4
    out["f1_paired"] = base(am4, ph4, z_base)
    # F-3 intensity
    out["f3"] = base.propagate_P2I(phase3.unsqueeze(0), z_base)
    # F-2 / F-5 need unpadded geometry to be coherent; only run when pad == 0
    if pad == 0:
        ap = torch.rand(B, 6, rows, cols, generator=gen)
        out["f2_in"] = ap
        out["f2"] = base.propagate_AP2AP(ap, z_base[:B])
        out["f5"] = fixed.propagate_AP2AP(ap)
    # F-4
    out["f4"] = fixed(amp, phase)
    # F-6 with grads
    a6 = amp.clone().requires_grad_(True)
    p6 = phase.clone().requires_grad_(True)
    y6 = fixed.propagate_AP2C_backward(a6, p6)
    cot6 = torch.complex(
        torch.randn(y6.shape, generator=gen), torch.randn(y6.shape, generator=gen)
    )
    (torch.view_as_real(y6) * torch.view_as_real(cot6)).sum().backward()
    out["f6"], out["f6_cot"], out["f6_ga"], out["f6_gp"] = y6, cot6, a6.grad, p6.grad
    # F-7 with grad
    p7 = phase.clone().requires_grad_(True)
    y7 = fixed.propagate_POH2Freq_forward(p7)
    cot7 = torch.complex(
        torch.randn(y7.shape, generator=gen), torch.randn(y7.shape, generator=gen)
    )
    (torch.view_as_real(y7) * torch.view_as_real(cot7)).sum().backward()
    out["f7"], out["f7_cot"], out["f7_gp"] = y7, cot7, p7.grad
    # F-8
    p8 = phase.clone().requires_grad_(True)
    c8 = torch.tensor(0.4, requires_grad=True)
    a8, q8, l8 = fixed.propagate_POH2AP_forward_with_spectrum_loss(p8, c8)
    w8 = torch.rand(a8.shape, generator=gen)
    ((a8 * w8).sum() + 3.0 * l8).backward()
    out["f8_amp"], out["f8_ang"], out["f8_loss"], out["f8_w"] = a8, q8, l8, w8
    out["f8_gp"], out["f8_gc"] = p8.grad, c8.grad
    # F-9
    out["f9_amp"], out["f9_ang"] = fixed.propagate_POH2AP_forward(phase)
    # F-10 with the bench's loss (MSE on amplitudes) and grad
    p10 = phase.clone().requires_grad_(True)
    y10 = multi(torch.ones_like(p10), p10, z_multi)
    tgt10 = torch.rand(y10.shape, generator=gen)
    l10 = torch.nn.functional.mse_loss(y10, tgt10)
    l10.backward()
    out["f10"], out["f10_tgt"], out["f10_loss"], out["f10_gp"] = y10, tgt10, l10, p10.grad
    # F-10 with amplitude and amp grad
    a10 = amp.clone().requires_grad_(True)
    p10b = phase.clone().requires_grad_(True)
    y10b = multi(a10, p10b, z_multi)
    (y10b * tgt10).sum().backward()
    out["f10b"], out["f10b_ga"], out["f10b_gp"] = y10b, a10.grad, p10b.grad
    # F-13, F-7 spectra feeding F-11 / F-12
    spec_t = multi.filter_AP2filteredFreq(amp, phs01)
    out["f13"] = spec_t
    spec_in = torch.cat((y7.detach(), spec_t), dim=0)  # [2B,3,Rp,Cp] like watermelon.py:229
    s11 = spec_in.clone()
    out["f11_amp"], out["f11_ang"] = (
        multi.propagate_multiple_samples_with_all_fixed_multiple_distances_freq2amp(s11)
    )
    s12 = spec_in.clone().requires_grad_(True)
    torch.manual_seed(7)
    idx = torch.randperm(multi.H.size(0))[0 : s12.size(0) // 2]
    torch.manual_seed(7)
    a12, q12 = multi.propagate_multiple_samples_with_random_fixed_multiple_distances_freq2amp(s12)
    wa = torch.rand(a12.shape, generator=gen)
    wq = torch.rand(a12.shape, generator=gen)
    # weight the angle by the amplitude^2 so the ill-conditioned |y|~0 pixels do not dominate
    ((a12 * wa).sum() + (torch.sin(q12) * wq * a12.detach() ** 2).sum()).backward()
    out["f12_seed"], out["f12_idx"] = 7, idx
    out["f12_amp"], out["f12_ang"], out["f12_wa"], out["f12_wq"] = a12, q12, wa, wq
    out["f12_gspec"] = s12.grad
    out["spec_in"] = spec_in
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **npify(out))
    print(name, "Rp x Cp =", Rp, Cp, "keys", len(out))


def main():
    asm, _ = ref_shim.load()
    for name, c in CASES.items():
        run_case(asm, name, **c)
    # known-answer data of the reference (README command): copy as data
    src = os.path.join(ref_shim.REFERENCE_ROOT, "output", "test_output", "terminalTest")
    dst = os.path.join(HERE, "terminalTest")
    os.makedirs(dst, exist_ok=True)
    poh = torch.load(os.path.join(src, "poh.pt"), map_location="cpu")
    np.save(os.path.join(dst, "poh.npy"), poh.detach().cpu().numpy())
    for i in range(10):
        shutil.copyfile(os.path.join(src, f"{i}.png"), os.path.join(dst, f"{i}.png"))
        os.chmod(os.path.join(dst, f"{i}.png"), 0o644)
    print("terminalTest copied:", tuple(poh.shape), poh.dtype)


if __name__ == "__main__":
    main()
