"""Forward focal-stack sweep of BASELINE config 5 (1080p RGB POH, batch B, D planes) on cuda:0: propagations/s
vs plane count, padded (2160 x 3840) and un-padded.  Diagnosis helper: python tools/time_c5.py [B]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learned_hologram_gan_b200.angular_spectrum_method as m
from learned_hologram_gan_b200 import focal_stack_export as E

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
            e2 = torch.cuda.Event(enable_timing=True)
            for _ in range(n):
                u8 = E.focal_stack_to_u8(out)  # generatePOH.py:72-78: the export that follows the focal stack
            e2.record()
            torch.cuda.synchronize()
        ms, ms_exp = e0.elapsed_time(e1) / n, e1.elapsed_time(e2) / n
        print(f"pad {pad:4d}  B {B}  D {D:3d}  {ms:9.3f} ms  {B * 3 * D / ms * 1e3:10.0f} propagations/s (forward)"
              f"   + 8-bit export {ms_exp:8.3f} ms ({out.numel() * (8 + 4 / 3) / ms_exp / 1e6:6.0f} GB/s)")
        del u8
        del out, prop
        torch.cuda.empty_cache()
