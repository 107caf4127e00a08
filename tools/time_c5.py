"""Forward focal-stack sweep of BASELINE config 5 (1080p RGB POH, batch B, D planes) on cuda:0: propagations/s
vs plane count, padded (2160 x 3840) and un-padded.  Diagnosis helper: python tools/time_c5.py [B]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learned_hologram_gan_b200.angular_spectrum_method as m

WL = torch.tensor([638e-9, 520e-9, 450e-9])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
R, C = 1080, 1920
g = torch.Generator().manual_seed(1)
phs = (6.28 * torch.rand(B, 3, R, C, generator=g)).cuda()
amp = torch.ones_like(phs)
for pad in (540, 0):
    for D in (1, 4, 16, 64):
        z = torch.linspace(4e-4, 10e-4, D)
        prop = m.bandLimitedAngularSpectrumMethod_for_multiple_distances(
            sample_row_num=R, sample_col_num=C, pad_size=pad, filter_radius_coefficient=0.45, pixel_pitch=3.74e-6,
            wave_length=WL, distances=z, cuda=True)
        with torch.no_grad():
            for _ in range(2):
                out = prop(amp, phs, z)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 3
            for _ in range(n):
                out = prop(amp, phs, z)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"pad {pad:4d}  B {B}  D {D:3d}  {ms:9.3f} ms  {B * 3 * D / ms * 1e3:10.0f} propagations/s (forward)")
        del out, prop
        torch.cuda.empty_cache()
