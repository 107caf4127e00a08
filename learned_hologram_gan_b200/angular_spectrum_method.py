"""Drop-in replacement for ``learnedMethodForHologram/angular_spectrum_method.py``.

Same three classes, constructor keywords, method names, argument order, shapes, dtypes and
error behaviour as the reference module, so ``trainingModel.py``, ``generatePOH.py`` and
``tests/test_angular_spectrum_method.py`` of the reference run against it unchanged
(see ``learned_hologram_gan_b200.overlay``).  Nothing is computed with torch.fft here:
each method is one or two calls of the fused CUDA pipeline behind ``include/asm_b200.h``
through an autograd Function whose backward is the adjoint propagation.

Behaviour kept on purpose (cites are reference lines, ``asm.py`` = angular_spectrum_method.py):
  * ``band_limit`` is stored and ignored (asm.py:53, :65-66); the band-limit mask can be
    generated but is never applied (asm.py:332 is commented out in the reference).
  * The circular "diffraction limited" mask is what every masked method applies (asm.py:60-63).
  * ``__call__`` returns the AMPLITUDE although the reference names it intensity (asm.py:92).
  * ``torch.randperm`` is drawn from the CPU global generator once per call (asm.py:536).
  * ``cuda=False`` keeps attributes and results on the host; the arithmetic still runs on the
    current CUDA device (inputs are staged), because this path has no CPU implementation.
"""

from __future__ import annotations

import torch

from . import _cabi as A
from . import engine as E
from . import utilities


def _device_for(cuda: bool) -> torch.device:
    return utilities.try_gpu() if cuda else torch.device("cpu")


class bandLimitedAngularSpectrumMethod:
    """Band-limited angular spectrum propagation, any distances (asm.py:5-260).

    dim 0 of the inputs is either 1 (broadcast over the distances) or equal to the number of
    distances (paired); the class does not combine batch and multi-distance (asm.py:11).
    """

    def __init__(
        self,
        sample_row_num=192,
        sample_col_num=192,
        pad_size=0,
        filter_radius_coefficient=0.5,
        pixel_pitch=3.74e-6,
        wave_length=torch.tensor([639e-9, 515e-9, 473e-9]),
        band_limit=False,
        cuda=False,
    ):
        self.originalRowNum = sample_row_num
        self.originalColNum = sample_col_num
        self.pad_size_row = pad_size
        self.pad_size_col = int(pad_size * (sample_col_num / sample_row_num))
        self.samplingRowNum = sample_row_num + 2 * self.pad_size_row
        self.samplingColNum = sample_col_num + 2 * self.pad_size_col
        self.pixel_pitch = pixel_pitch
        self.wave_length = wave_length
        self.band_limit = band_limit
        self.device = _device_for(cuda)
        self.freq_x = torch.fft.fftfreq(self.samplingRowNum, self.pixel_pitch)
        self.freq_y = torch.fft.fftfreq(self.samplingColNum, self.pixel_pitch)

        shorter = min(self.samplingRowNum, self.samplingColNum)
        radius = shorter * filter_radius_coefficient
        if radius > shorter / 2:  # util.py:225-229, raised from the constructor like the reference
            raise ValueError(
                f"The radius {radius} is larger than the half of the sample size {shorter / 2}"
            )
        self._radius = radius
        self._coef = filter_radius_coefficient
        self._plan = E.Plan(
            sample_row_num, sample_col_num, self.pad_size_row, self.pad_size_col,
            pixel_pitch, wave_length, radius,
        )
        self._cache = {}

    # ---- lazily materialised attributes (the reference builds them eagerly) ---------------
    def _cached(self, key, make):
        if key not in self._cache:
            self._cache[key] = make().to(self.device)
        return self._cache[key]

    @property
    def diffraction_limited_mask(self):
        return self._cached("mask", lambda: self.generate_diffraction_limited_mask(self._coef))

    @property
    def w_grid(self):
        return self._cached("w", self.generate_w_grid)

    # ---- helpers --------------------------------------------------------------------------
    def _z(self, distances):
        return E.upload_small(torch.as_tensor(distances).reshape(-1), torch.float32, self._plan.device)

    def _prep(self, t):
        """[3,R,C] -> [1,3,R,C]; shape check against the plan."""
        if t.dim() == 3:
            t = t.unsqueeze(0)
        if t.dim() != 4 or t.shape[1] != self._plan.n_colour or tuple(t.shape[2:]) != (
            self.originalRowNum, self.originalColNum,
        ):
            raise RuntimeError(
                f"expected [N,{self._plan.n_colour},{self.originalRowNum},{self.originalColNum}], "
                f"got {tuple(t.shape)}"
            )
        return t

    def _broadcast_depth(self, n_in, n_z):
        """dim-0 broadcasting of G_0 [n_in,3,..] against H [n_z,3,..] (asm.py:91):
        returns (n_depth per sample, depth_index or None)."""
        if n_in == 1:
            return n_z, None
        if n_z == 1:
            return 1, torch.zeros(n_in, dtype=torch.int32, device=self._plan.device)
        if n_in == n_z:
            return 1, torch.arange(n_in, dtype=torch.int32, device=self._plan.device)
        raise RuntimeError(
            f"The size of tensor a ({n_in}) must match the size of tensor b ({n_z}) at non-singleton dimension 0"
        )

    def _pair(self, amp, phase):
        phase = self._prep(phase)
        if amp is None:
            return None, phase
        amp, phase = torch.broadcast_tensors(self._prep(amp), phase)
        return amp, phase

    def _propagate(self, amp, phase, distances, out, mask=True, conj=False):
        amp, phase = self._pair(amp, phase)
        z = self._z(distances)
        n_depth, index = self._broadcast_depth(phase.shape[0], z.numel())
        filt = E.FilterSpec(True, conj, mask, z, index)
        res = E.field_to_field(self._plan, filt, n_depth, out, amp, phase)
        dev = phase.device
        if isinstance(res, tuple):
            return tuple(r.to(dev) for r in res)
        return res.to(dev)

    # ---- reference surface ----------------------------------------------------------------
    def __call__(self, amplitute_tensor, phase_tensor, distances):
        """|crop(ifft2(fft2(pad(a e^{i phi})) H mask))| -> [D,3,R,C]  (asm.py:68-94)."""
        return self._propagate(amplitute_tensor, phase_tensor, distances, "abs")

    def propagate_AP2AP(self, amp_phs_tensor_0, distances):
        """Interleaved (amp,phase) per colour in, planar [amp x3, angle x3] out, x H, no mask
        (asm.py:96-129).  Like the reference it is only coherent without padding."""
        if self.pad_size_row != 0:
            raise RuntimeError("propagate_AP2AP is only defined for pad_size == 0 (asm.py:113-127)")
        v = amp_phs_tensor_0.view(-1, 3, 2, self.samplingRowNum, self.samplingColNum)
        amp, ang = self._propagate(v[:, :, 0], v[:, :, 1], distances, "abs_angle", mask=False)
        return torch.cat((amp, ang), dim=1)

    def propagate_P2I(self, phase_tensor, distances):
        """Intensity |.|^2 (asm.py:131-139)."""
        return self._propagate(None, phase_tensor, distances, "abs2")

    def generate_diffraction_limited_mask(self, radius_coefficient):
        """Circular low-pass [Rp,Cp] f32 (asm.py:141-153 -> util.py:206-243)."""
        shorter = min(self.samplingRowNum, self.samplingColNum)
        radius = shorter * radius_coefficient
        if radius > shorter / 2:
            raise ValueError(
                f"The radius {radius} is larger than the half of the sample size {shorter / 2}"
            )
        grid = E.host_radial_grid(self.samplingRowNum, self.samplingColNum)
        mask = torch.ones_like(grid)
        mask[grid > radius] = 0.0
        return mask.to(self.device)

    def generate_w_grid(self):
        """sqrt(max(1/lambda^2 - fx^2 - fy^2, 0)) [3,Rp,Cp] f32 (asm.py:155-171).  Host-built once per
        geometry with the reference's own ops (see engine.host_wm_grid for why)."""
        if self._plan.wm is not None:
            return torch.abs(self._plan.wm).to(self.device)
        return self._plan.build_grid(A.GRID_W).to(self.device)

    def _band_limit(self, distances):
        d_r = 1 / (self.samplingRowNum * self.pixel_pitch)
        d_c = 1 / (self.samplingColNum * self.pixel_pitch)
        z = distances.detach().cpu().unsqueeze(1)
        lam = self.wave_length.detach().cpu().unsqueeze(0)
        lim_r = 1 / (torch.sqrt((2 * d_r * z) ** 2 + 1) * lam)
        lim_c = 1 / (torch.sqrt((2 * d_c * z) ** 2 + 1) * lam)
        keep_r = torch.abs(self.freq_x)[None, None, :, None] < lim_r[:, :, None, None]
        keep_c = torch.abs(self.freq_y)[None, None, None, :] < lim_c[:, :, None, None]
        return (keep_r & keep_c).to(self.device)

    def generate_band_limited_mask(self, distances):
        """Band-limit mask [D,3,Rp,Cp] bool (asm.py:173-193); generated, never applied (host ops)."""
        return self._band_limit(distances)

    def generate_transfer_function(self, distances):
        """H = exp(-2 pi i z w) [D,3,Rp,Cp] c64, fp32-faithful (asm.py:195-213)."""
        z = self._z(distances)
        return self._plan.build_grid(A.GRID_H, z).to(self.device)

    def padding(self, tensor):
        """Centred zero-pad (asm.py:215-239)."""
        if self.pad_size_row == 0:
            return tensor
        return torch.nn.functional.pad(
            tensor,
            (self.pad_size_col, self.pad_size_col, self.pad_size_row, self.pad_size_row),
            mode="constant",
            value=0,
        )

    def cropping(self, tensor):
        """Inverse of padding (asm.py:241-260)."""
        if self.pad_size_row == 0:
            return tensor
        return tensor[
            :, :, self.pad_size_row : -self.pad_size_row, self.pad_size_col : -self.pad_size_col
        ]


class bandLimitedAngularSpectrumMethod_for_single_fixed_distance(bandLimitedAngularSpectrumMethod):
    """One fixed distance, batched 4-D inputs only (asm.py:263-466)."""

    def __init__(
        self,
        sample_row_num=192,
        sample_col_num=192,
        pad_size=0,
        filter_radius_coefficient=0.5,
        pixel_pitch=3.74e-6,
        wave_length=torch.tensor([639e-9, 515e-9, 473e-9]),
        band_limit=False,
        cuda=False,
        distance=torch.tensor([1e-3]),
    ):
        super().__init__(
            sample_row_num, sample_col_num, pad_size, filter_radius_coefficient,
            pixel_pitch, wave_length, band_limit, cuda,
        )
        self.distance = distance
        self._zdev = self._z(distance)

    @property
    def circular_frequency_mask_differentiable_grid(self):
        return self._cached("radial", lambda: E.host_radial_grid(self.samplingRowNum, self.samplingColNum))

    @property
    def band_limited_mask(self):
        return self._cached("band", lambda: self._band_limit(self.distance))

    @property
    def H(self):
        """Cached transfer function [3,Rp,Cp] (asm.py:321)."""
        return self._cached("H", lambda: self._plan.build_grid(A.GRID_H, self._zdev)[0])

    def _fixed(self, amp, phase, out, mask, conj=False):
        amp, phase = self._pair(amp, phase)
        filt = E.FilterSpec(True, conj, mask, self._zdev, None)
        res = E.field_to_field(self._plan, filt, 1, out, amp, phase)
        dev = phase.device
        if isinstance(res, tuple):
            return tuple(r.to(dev) for r in res)
        return res.to(dev)

    def __call__(self, amplitute_tensor, phase_tensor):
        """asm.py:323-336."""
        return self._fixed(amplitute_tensor, phase_tensor, "abs", mask=True)

    def propagate_AP2AP(self, amp_phs_tensor_0):
        """Backward direction: divides by H (asm.py:338-368)."""
        if self.pad_size_row != 0:
            raise RuntimeError("propagate_AP2AP is only defined for pad_size == 0 (asm.py:352-366)")
        v = amp_phs_tensor_0.view(-1, 3, 2, self.samplingRowNum, self.samplingColNum)
        amp, ang = self._fixed(v[:, :, 0], v[:, :, 1], "abs_angle", mask=False, conj=True)
        return torch.cat((amp, ang), dim=1)

    def propagate_AP2C_backward(self, amp_z, phs_z):
        """crop(ifft2(fft2(pad(a e^{i phi})) / H)) complex field (asm.py:374-384).
        1/H is evaluated as conj(H): |H| = 1 to 6e-8."""
        return self._fixed(amp_z, phs_z, "complex", mask=False, conj=True)

    def propagate_POH2Freq_forward(self, POH):
        """fft2(pad(e^{i POH})) H mask, padded spectrum out (asm.py:386-392)."""
        POH = self._prep(POH)
        filt = E.FilterSpec(True, False, True, self._zdev, None)
        return E.field_to_spectrum(self._plan, filt, None, POH).to(POH.device)

    def propagate_POH2AP_forward_with_spectrum_loss(self, phs_0, filter_radius_coefficient=torch.tensor(0.5)):
        """asm.py:394-412.  The FFTs run in the CUDA pipeline; the soft (sigmoid) mask, which is
        differentiable in its coefficient, and the spectrum-mean loss are point-wise torch ops."""
        phs_0 = self._prep(phs_0)
        dev = self._plan.device
        none = E.FilterSpec(False, False, False, None, None)
        G_0 = E.field_to_spectrum(self._plan, none, None, phs_0)
        soft = self.generate_circular_frequency_mask_differentiable(filter_radius_coefficient).to(dev)
        G_z_filtered = G_0 * self.H.to(dev) * soft
        spectrum_mean_loss = torch.mean(torch.abs(G_0) - torch.abs(G_z_filtered))
        amp, ang = E.spectrum_to_field(self._plan, none, 1, "abs_angle", G_z_filtered)
        out_dev = phs_0.device
        return amp.to(out_dev), ang.to(out_dev), spectrum_mean_loss.to(out_dev)

    def propagate_POH2AP_forward(self, phs_0):
        """asm.py:414-424."""
        return self._fixed(None, phs_0, "abs_angle", mask=True)

    def generate_circular_frequency_mask_differentiable(self, filter_radius_coefficient):
        """sigmoid(radius - D) (asm.py:426-436); differentiable in the coefficient."""
        shorter_edge = min(self.samplingRowNum, self.samplingColNum)
        radius = shorter_edge * filter_radius_coefficient
        grid = self.circular_frequency_mask_differentiable_grid
        if isinstance(radius, torch.Tensor):
            radius = radius.to(grid.device)
        return torch.sigmoid(1.0 * (radius - grid))

    def generate_band_limited_mask(self):
        """asm.py:442-462: [1,3,Rp,Cp] bool for the fixed distance."""
        return self._band_limit(self.distance)

    def generate_transfer_function(self):
        """asm.py:464-466: [3,Rp,Cp] c64."""
        return self._plan.build_grid(A.GRID_H, self._zdev)[0].to(self.device)


class bandLimitedAngularSpectrumMethod_for_multiple_distances(bandLimitedAngularSpectrumMethod):
    """Batch x multiple distances (asm.py:469-552).  Output index = sample*D + depth."""

    def __init__(
        self,
        sample_row_num=192,
        sample_col_num=192,
        distances=None,
        pad_size=160,
        filter_radius_coefficient=0.5,
        pixel_pitch=3.74e-6,
        wave_length=torch.tensor([639e-9, 515e-9, 473e-9]),
        band_limit=False,
        cuda=True,
    ):
        super().__init__(
            sample_row_num, sample_col_num, pad_size, filter_radius_coefficient,
            pixel_pitch, wave_length, band_limit, cuda,
        )
        self.distances = distances.to(self.device)  # AttributeError on None, like asm.py:500
        self._zdev = self._z(distances)

    @property
    def H(self):
        """Cached stack [D,3,Rp,Cp] (asm.py:501); materialised on first read only."""
        return self._cached("H", lambda: self._plan.build_grid(A.GRID_H, self._zdev))

    def __call__(self, amplitute_tensor, phase_tensor, distances):
        """asm.py:503-522: every sample to every distance, amplitude out [B*D,3,R,C]."""
        amp, phase = self._pair(amplitute_tensor, phase_tensor)
        z = self._z(distances)
        filt = E.FilterSpec(True, False, True, z, None)
        return E.field_to_field(self._plan, filt, int(z.numel()), "abs", amp, phase).to(phase.device)

    def propagate_with_amplitude_mse(self, amplitute_tensor, phase_tensor, distances, target_amplitude):
        """Extension (not in the reference): the bench workload ``mse_loss(self(a, phi, z), target)``
        with the L2 reduction fused into the last pass and a fused adjoint.  Returns (loss, amplitude)."""
        amp, phase = self._pair(amplitute_tensor, phase_tensor)
        z = self._z(distances)
        filt = E.FilterSpec(True, False, True, z, None)
        loss, amp_hat = E.amplitude_mse(self._plan, filt, int(z.numel()), amp, phase, target_amplitude)
        return loss, amp_hat.to(phase.device)

    def amplitude_mse_and_phase_gradient(self, phase_tensor, distances, target_amplitude, grad_scale, grad_out=None):
        """Extension: forward + adjoint of ``sum((self(1, phi, z) - target)^2)`` in two library calls, no
        autograd graph.  Returns (sum of squared errors, grad_scale/2 * its phase gradient)."""
        phase_tensor = self._prep(phase_tensor)  # [N,n_colour,R,C] of THIS plan: the kernels trust the shape
        z = self._z(distances)
        filt = E.FilterSpec(True, False, True, z, None)
        return E.amplitude_mse_direct(self._plan, filt, int(z.numel()), phase_tensor, target_amplitude,
                                      grad_scale, grad_out)

    def _check_spectrum(self, G_0):
        if G_0.dim() != 4 or tuple(G_0.shape[1:]) != (3, self.samplingRowNum, self.samplingColNum):
            raise RuntimeError(
                f"expected [N,3,{self.samplingRowNum},{self.samplingColNum}] spectrum, got {tuple(G_0.shape)}"
            )

    def propagate_multiple_samples_with_all_fixed_multiple_distances_freq2amp(self, G_0):
        """asm.py:524-531: spectrum in, all constructor distances, (abs, angle) out."""
        self._check_spectrum(G_0)
        filt = E.FilterSpec(True, False, True, self._zdev, None)
        amp, ang = E.spectrum_to_field(self._plan, filt, int(self._zdev.numel()), "abs_angle", G_0)
        return amp.to(G_0.device), ang.to(G_0.device)

    def propagate_multiple_samples_with_random_fixed_multiple_distances_freq2amp(self, G_0):
        """asm.py:533-546: sample i and sample i+N/2 share one random constructor distance."""
        self._check_spectrum(G_0)
        n = G_0.size(0)
        staged = getattr(self, "_staged_index", None)
        if staged is not None and staged.numel() == n and n % 2 == 0:
            index = staged  # drawn by stage_random_depths (the step is being captured / replayed as a CUDA graph)
        else:
            indices = torch.randperm(self._zdev.numel())[0 : n // 2]  # CPU global generator, asm.py:536
            if n % 2 != 0 or indices.numel() != n // 2:
                raise RuntimeError(
                    f"need an even number of spectra, at most twice the {self._zdev.numel()} distances; got {n}"
                )
            index = E.upload_small(torch.cat((indices, indices)), torch.int32, self._plan.device)
        filt = E.FilterSpec(True, False, True, self._zdev, index)
        amp, ang = E.spectrum_to_field(self._plan, filt, 1, "abs_angle", G_0)
        return amp.to(G_0.device), ang.to(G_0.device)

    def stage_random_depths(self, n_half):
        """Extension for CUDA-graph training steps.  The reference draws ``torch.randperm(D)[:N/2]`` on the host inside
        every call (asm.py:536), which a captured graph would freeze.  This draws the SAME thing (one randperm from
        the CPU global generator) ahead of the call into a persistent device buffer; the next calls of
        ``propagate_multiple_samples_with_random_fixed_multiple_distances_freq2amp`` with N = 2*n_half spectra use
        that buffer instead of drawing, so a replayed graph sees fresh depths after every ``stage_random_depths``.
        ``stage_random_depths(0)`` returns to the reference behaviour."""
        if n_half <= 0:
            self._staged_index = None
            return None
        D = int(self._zdev.numel())
        if n_half > D:
            raise RuntimeError(f"need at most {D} depths, got {n_half}")
        indices = torch.randperm(D)[0:n_half]
        dev = self._plan.device
        if getattr(self, "_staged_index", None) is None or self._staged_index.numel() != 2 * n_half:
            self._staged_index = torch.zeros(2 * n_half, dtype=torch.int32, device=dev)
            self._staged_pin = torch.zeros(2 * n_half, dtype=torch.int32).pin_memory()
            self._staged_evt = None
        if self._staged_evt is not None:
            self._staged_evt.synchronize()  # the previous upload has left the pinned buffer
        self._staged_pin.copy_(torch.cat((indices, indices)).to(torch.int32))
        with torch.cuda.device(dev):
            self._staged_index.copy_(self._staged_pin, non_blocking=True)
            self._staged_evt = torch.cuda.Event()
            self._staged_evt.record()
        return indices

    def filter_AP2filteredFreq(self, amp, phs):
        """fft2(pad(a e^{i 2 pi phs})) mask (asm.py:548-552)."""
        amp, phs = self._pair(amp, phs)
        filt = E.FilterSpec(False, False, True, None, None)
        two_pi = float(torch.tensor(2 * torch.pi, dtype=torch.float32))
        return E.field_to_spectrum(self._plan, filt, amp, phs, phase_scale=two_pi).to(phs.device)
