"""CPU oracle for the band-limited angular-spectrum propagation path.

TEST INFRASTRUCTURE ONLY.  This file is a restatement, in plain torch-on-CPU
arithmetic, of the algorithm in the reference's
``learnedMethodForHologram/angular_spectrum_method.py`` and the three helpers
it uses from ``learnedMethodForHologram/utilities.py``.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it; the product package
(``learned_hologram_gan_b200``) never does and fails loudly without its CUDA
library.

Where the arithmetic lives: the reference vendors none of it.  Every number is
produced by PyTorch (pinned ``torch==2.1.2`` in the reference's
``pyproject.toml:7``; 2.11.0 in this image): ``torch.fft.fft2/ifft2/fftfreq``
plus ATen element-wise kernels.  The restatement therefore calls the same
torch CPU primitives, in the same order and dtype (fp32 / complex64), but is
organised as free functions over a small ``Geometry`` record instead of the
reference's three classes.

Pinning (see ``tests/test_oracle_pinned.py``):
  * against the reference module itself, imported from ``/root/reference``
    through the shim in ``oracle/ref_shim.py`` (only where that tree exists);
  * against the fixtures the reference was run on to produce
    ``tests/golden/*.npz`` (script: ``tests/golden/make_golden.py``);
  * against the reference's only known-answer data,
    ``output/test_output/terminalTest/{poh.pt,0..9.png}`` (copied as data to
    ``tests/golden/terminalTest/``).

Each function cites the reference lines it follows as ``asm.py:L`` for
``learnedMethodForHologram/angular_spectrum_method.py`` and ``util.py:L`` for
``learnedMethodForHologram/utilities.py``.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import torch

DEFAULT_WAVELENGTHS = (639e-9, 515e-9, 473e-9)  # asm.py:37


# --------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------
@dataclass
class Geometry:
    """Sizes and physical constants of one propagator (asm.py:30-63)."""

    rows: int
    cols: int
    pad: int = 0
    radius_coef: float = 0.5
    pitch: float = 3.74e-6
    wavelengths: torch.Tensor = field(
        default_factory=lambda: torch.tensor(DEFAULT_WAVELENGTHS)
    )

    def __post_init__(self):
        self.pad_r = self.pad  # asm.py:45
        self.pad_c = int(self.pad * (self.cols / self.rows))  # asm.py:46
        self.prow = self.rows + 2 * self.pad_r  # asm.py:48
        self.pcol = self.cols + 2 * self.pad_c  # asm.py:49
        self.wavelengths = torch.as_tensor(self.wavelengths, dtype=torch.float32)
        # asm.py:56-57 -- note: "freq_x" runs along ROWS, "freq_y" along columns
        self.freq_r = torch.fft.fftfreq(self.prow, self.pitch)
        self.freq_c = torch.fft.fftfreq(self.pcol, self.pitch)


# --------------------------------------------------------------------------
# masks and transfer function
# --------------------------------------------------------------------------
def radial_grid(prow: int, pcol: int) -> torch.Tensor:
    """Unitless radial frequency scaled by the shorter edge (util.py:290-296)."""
    u = torch.fft.fftfreq(prow).unsqueeze(-1)
    v = torch.fft.fftfreq(pcol).unsqueeze(0)
    return torch.sqrt(u**2 + v**2) * min(prow, pcol)


def circular_mask(prow: int, pcol: int, radius: float) -> torch.Tensor:
    """Hard circular low-pass, DC at [0,0] (util.py:206-243, decay_rate=None)."""
    if radius > min(prow, pcol) / 2:  # util.py:225-229
        raise ValueError(
            f"The radius {radius} is larger than the half of the sample size "
            f"{min(prow, pcol) / 2}"
        )
    dist = radial_grid(prow, pcol)
    mask = torch.ones_like(dist)
    mask[dist > radius] = 0.0
    return mask


def diffraction_limited_mask(g: Geometry) -> torch.Tensor:
    """asm.py:141-153: radius = min(Rp, Cp) * coefficient."""
    return circular_mask(g.prow, g.pcol, min(g.prow, g.pcol) * g.radius_coef)


def soft_circular_mask(g: Geometry, coef) -> torch.Tensor:
    """Sigmoid-edged mask with a (possibly tensor) coefficient (asm.py:426-436)."""
    radius = min(g.prow, g.pcol) * coef
    return torch.sigmoid(1.0 * (radius - radial_grid(g.prow, g.pcol)))


def w_grid(g: Geometry) -> torch.Tensor:
    """Axial frequency sqrt(max(1/lambda^2 - fx^2 - fy^2, 0)), fp32 (asm.py:155-171)."""
    sq = g.freq_r.unsqueeze(1) ** 2 + g.freq_c.unsqueeze(0) ** 2
    inv_l2 = (1 / g.wavelengths**2).unsqueeze(1).unsqueeze(2)
    return torch.sqrt(torch.clamp(inv_l2 - sq.unsqueeze(0), min=0))


def transfer_function(g: Geometry, distances: torch.Tensor, w: Optional[torch.Tensor] = None) -> torch.Tensor:
    """H[d,c,:,:] = exp(-2j*pi*z_d*w_c), complex64 (asm.py:195-213)."""
    if w is None:
        w = w_grid(g)
    z = distances.to(torch.float32).reshape(-1, 1, 1, 1)
    return torch.exp(-2j * torch.pi * z * w)


def transfer_function_fixed(g: Geometry, distance: torch.Tensor) -> torch.Tensor:
    """Fixed-distance flavour: a [1]-shaped distance broadcasts, H is [3,Rp,Cp] (asm.py:464-466)."""
    return torch.exp(-2j * torch.pi * distance * w_grid(g))


def band_limit_mask(g: Geometry, distances: torch.Tensor) -> torch.Tensor:
    """Matsushima band limit [D,3,Rp,Cp] bool (asm.py:173-193).  Built but never applied
    by the reference (asm.py:65-66, :332)."""
    d_r = 1 / (g.prow * g.pitch)
    d_c = 1 / (g.pcol * g.pitch)
    z = distances.unsqueeze(1)
    lam = g.wavelengths.unsqueeze(0)
    lim_r = 1 / (torch.sqrt((2 * d_r * z) ** 2 + 1) * lam)
    lim_c = 1 / (torch.sqrt((2 * d_c * z) ** 2 + 1) * lam)
    m_r = torch.abs(g.freq_r)[None, None, :, None] < lim_r[:, :, None, None]
    m_c = torch.abs(g.freq_c)[None, None, None, :] < lim_c[:, :, None, None]
    return m_r & m_c


# --------------------------------------------------------------------------
# pad / crop / field construction
# --------------------------------------------------------------------------
def pad(g: Geometry, x: torch.Tensor) -> torch.Tensor:
    """Centred zero-pad of the last two dims (asm.py:215-239)."""
    if g.pad_r == 0:
        return x
    return torch.nn.functional.pad(x, (g.pad_c, g.pad_c, g.pad_r, g.pad_r))


def crop(g: Geometry, x: torch.Tensor) -> torch.Tensor:
    """Inverse of pad; the reference slices a 4-D tensor (asm.py:241-260)."""
    if g.pad_r == 0:
        return x
    return x[..., g.pad_r : -g.pad_r, g.pad_c : -g.pad_c]


def phasor(amp: Optional[torch.Tensor], phase: torch.Tensor) -> torch.Tensor:
    """amp * exp(1j*phase) (asm.py:88); amp None means the phase-only form (asm.py:136)."""
    e = torch.exp(1j * phase)
    return e if amp is None else amp * e


def spectrum_of(g: Geometry, amp, phase) -> torch.Tensor:
    """fft2(pad(a*exp(i*phi))) over the last two dims (asm.py:87-89)."""
    return torch.fft.fft2(pad(g, phasor(amp, phase)))


def field_from_spectrum(g: Geometry, spec: torch.Tensor) -> torch.Tensor:
    """crop(ifft2(G)) (asm.py:92)."""
    return crop(g, torch.fft.ifft2(spec))


# --------------------------------------------------------------------------
# the reference's methods, one function each
# --------------------------------------------------------------------------
def base_call(g: Geometry, amp, phase, distances, mask=None) -> torch.Tensor:
    """F-1  bandLimitedAngularSpectrumMethod.__call__ (asm.py:68-94): amplitude out."""
    mask = diffraction_limited_mask(g) if mask is None else mask
    g0 = spectrum_of(g, amp, phase)
    gz = g0 * transfer_function(g, distances) * mask
    return torch.abs(field_from_spectrum(g, gz))


def base_ap2ap(g: Geometry, amp_phs, distances) -> torch.Tensor:
    """F-2  base propagate_AP2AP (asm.py:96-129): interleaved (amp,phs) in, planar out, x H, no mask."""
    v = amp_phs.view(-1, 3, 2, g.prow, g.pcol)
    g0 = torch.fft.fft2(pad(g, v[:, :, 0] * torch.exp(1j * v[:, :, 1])))
    gz = field_from_spectrum(g, g0 * transfer_function(g, distances))
    return torch.cat((torch.abs(gz), torch.angle(gz)), dim=1)


def base_p2i(g: Geometry, phase, distances, mask=None) -> torch.Tensor:
    """F-3  propagate_P2I (asm.py:131-139): intensity |.|^2 out."""
    mask = diffraction_limited_mask(g) if mask is None else mask
    gz = spectrum_of(g, None, phase) * transfer_function(g, distances) * mask
    return torch.abs(field_from_spectrum(g, gz)) ** 2


def fixed_call(g: Geometry, distance, amp, phase) -> torch.Tensor:
    """F-4  fixed-distance __call__ (asm.py:323-336)."""
    h = transfer_function_fixed(g, distance)
    gz = spectrum_of(g, amp, phase) * h * diffraction_limited_mask(g)
    return torch.abs(field_from_spectrum(g, gz))


def fixed_ap2ap(g: Geometry, distance, amp_phs) -> torch.Tensor:
    """F-5  fixed-distance propagate_AP2AP (asm.py:338-368): divides by H."""
    h = transfer_function_fixed(g, distance)
    v = amp_phs.view(-1, 3, 2, g.prow, g.pcol)
    g0 = torch.fft.fft2(pad(g, v[:, :, 0] * torch.exp(1j * v[:, :, 1])))
    gz = field_from_spectrum(g, g0 / h)
    return torch.cat((torch.abs(gz), torch.angle(gz)), dim=1)


def fixed_ap2c_backward(g: Geometry, distance, amp, phase) -> torch.Tensor:
    """F-6  propagate_AP2C_backward (asm.py:374-384): complex field out, / H, no mask."""
    h = transfer_function_fixed(g, distance)
    return field_from_spectrum(g, spectrum_of(g, amp, phase) / h)


def fixed_poh2freq(g: Geometry, distance, poh) -> torch.Tensor:
    """F-7  propagate_POH2Freq_forward (asm.py:386-392): padded spectrum out."""
    h = transfer_function_fixed(g, distance)
    return spectrum_of(g, None, poh) * h * diffraction_limited_mask(g)


def fixed_poh2ap_spectrum_loss(g: Geometry, distance, phase, coef=torch.tensor(0.5)):
    """F-8  propagate_POH2AP_forward_with_spectrum_loss (asm.py:394-412)."""
    h = transfer_function_fixed(g, distance)
    g0 = spectrum_of(g, None, phase)
    gf = g0 * h * soft_circular_mask(g, coef)
    loss = torch.mean(torch.abs(g0) - torch.abs(gf))
    gz = field_from_spectrum(g, gf)
    return torch.abs(gz), torch.angle(gz), loss


def fixed_poh2ap(g: Geometry, distance, phase):
    """F-9  propagate_POH2AP_forward (asm.py:414-424)."""
    gz = field_from_spectrum(g, fixed_poh2freq(g, distance, phase))
    return torch.abs(gz), torch.angle(gz)


def multi_call(g: Geometry, amp, phase, distances) -> torch.Tensor:
    """F-10 multi-distance __call__ (asm.py:503-522): [B*D,3,R,C], index b*D+d."""
    g0 = spectrum_of(g, amp, phase)
    h = transfer_function(g, distances) * diffraction_limited_mask(g)
    gz = (g0.unsqueeze(1) * h).view(-1, 3, g.prow, g.pcol)
    return torch.abs(field_from_spectrum(g, gz))


def multi_all_freq2amp(g: Geometry, ctor_distances, spec):
    """F-11 ..._all_fixed_multiple_distances_freq2amp (asm.py:524-531)."""
    h = transfer_function(g, ctor_distances)
    gz = spec.unsqueeze(1) * h * diffraction_limited_mask(g)
    f = field_from_spectrum(g, gz.view(-1, 3, g.prow, g.pcol))
    return torch.abs(f), torch.angle(f)


def multi_random_freq2amp(g: Geometry, ctor_distances, spec, indices=None):
    """F-12 ..._random_fixed_multiple_distances_freq2amp (asm.py:533-546).

    ``indices`` None draws ``torch.randperm`` from the CPU global generator exactly as
    the reference does (asm.py:536)."""
    h_all = transfer_function(g, ctor_distances)
    if indices is None:
        indices = torch.randperm(h_all.size(0))[0 : spec.size(0) // 2]
    gz = spec.view(2, -1, 3, g.prow, g.pcol) * h_all[indices] * diffraction_limited_mask(g)
    f = field_from_spectrum(g, gz.view(-1, 3, g.prow, g.pcol))
    return torch.abs(f), torch.angle(f)


def multi_filter_ap2freq(g: Geometry, amp, phs) -> torch.Tensor:
    """F-13 filter_AP2filteredFreq (asm.py:548-552): phase is in [0,1), scaled by 2*pi."""
    phs = 2 * torch.pi * phs
    return spectrum_of(g, amp, phs) * diffraction_limited_mask(g)


# --------------------------------------------------------------------------
# helpers used by benches / fixtures
# --------------------------------------------------------------------------
def normalize_planes(x: torch.Tensor) -> torch.Tensor:
    """Per-plane min/max normalisation to [0,1] (util.py:69-84)."""
    hi = x.amax(dim=(-2, -1), keepdim=True)
    lo = x.amin(dim=(-2, -1), keepdim=True)
    return (x - lo) / (hi - lo)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| over all elements, computed in fp64 (works for complex)."""
    a = a.detach().cpu()
    b = b.detach().cpu()
    if a.is_complex() or b.is_complex():
        a = a.to(torch.complex128)
        b = b.to(torch.complex128)
    else:
        a = a.double()
        b = b.double()
    den = torch.linalg.vector_norm(b).item()
    num = torch.linalg.vector_norm(a - b).item()
    return num / den if den > 0 else num


def amp_mse_forward_backward(g: Geometry, phase, distances, target):
    """Bench workload (BASELINE config 2/4): multi_call -> MSE against target -> d/d(phase)."""
    phase = phase.detach().clone().requires_grad_(True)
    amp_hat = multi_call(g, torch.ones_like(phase), phase, distances)
    loss = torch.nn.functional.mse_loss(amp_hat, target)
    loss.backward()
    return loss.detach(), phase.grad.detach(), amp_hat.detach()
