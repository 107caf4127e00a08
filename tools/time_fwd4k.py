"""Forward-only focal stack at the C4 geometry (the generatePOH.py call: amplitudes of 8 planes, no loss, no adjoint),
CUDA-event time per call and the library's per-kernel times.  Diagnosis helper (A/B of LHG_ROWS_PAIR / LHG_ROWS_TMA)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learned_hologram_gan_b200.angular_spectrum_method as m  # noqa: E402
from learned_hologram_gan_b200 import _cabi  # noqa: E402

WL = torch.tensor([638e-9, 520e-9, 450e-9])
z = torch.linspace(4e-4, 10e-4, 8)
prop = m.bandLimitedAngularSpectrumMethod_for_multiple_distances(sample_row_num=2160, sample_col_num=3840, distances=z,
                                                                 pad_size=1080, filter_radius_coefficient=0.45,
                                                                 wave_length=WL, cuda=True)
g = torch.Generator().manual_seed(4)
phase = (6.28 * torch.rand(1, 3, 2160, 3840, generator=g)).cuda()
ones = torch.ones_like(phase)
lib = _cabi.load()
with torch.no_grad():
    for _ in range(3):
        out = prop(ones, phase, z)
    torch.cuda.synchronize()
    lib.asm_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        out = prop(ones, phase, z)
    e1.record()
    torch.cuda.synchronize()
    lib.asm_profile_enable(0)
kms, kn = (C.c_double * 4)(0, 0, 0, 0), (C.c_longlong * 4)(0, 0, 0, 0)
lib.asm_profile_collect(kms, kn, 4)
print({k: os.environ.get(k) for k in ("LHG_ROWS_PAIR", "LHG_ROWS_TMA")}, "ms/call", round(e0.elapsed_time(e1) / n, 3),
      "kernels (row fwd, column, row inv, fused)", [round(x / n, 3) for x in kms], "checksum", float(out.double().sum()))
