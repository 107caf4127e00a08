"""Time the calls of one trainingModel.py-style step (BASELINE config 3: 384x384, batch 4, pad 320) on cuda:0.
Diagnosis helper: python tools/time_c3.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learned_hologram_gan_b200.angular_spectrum_method as m

WL = torch.tensor([638e-9, 520e-9, 450e-9])
B, R = 4, 384
fixed = m.bandLimitedAngularSpectrumMethod_for_single_fixed_distance(
    sample_row_num=R, sample_col_num=R, pad_size=320, filter_radius_coefficient=0.45, pixel_pitch=3.74e-6,
    wave_length=WL, distance=torch.tensor([1e-3]), cuda=True)
multi = m.bandLimitedAngularSpectrumMethod_for_multiple_distances(
    sample_row_num=R, sample_col_num=R, pad_size=320, filter_radius_coefficient=0.45, pixel_pitch=3.74e-6,
    wave_length=WL, distances=torch.linspace(-4e-4, 0, 21)[:-1], cuda=True)
g = torch.Generator().manual_seed(1)
amp = torch.rand(B, 3, R, R, generator=g).cuda()
phs = (6.28 * torch.rand(B, 3, R, R, generator=g)).cuda()


def t(name, fn, n=20):
    import ctypes as C
    from learned_hologram_gan_b200 import _cabi
    lib = _cabi.load()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / n
    t0 = time.perf_counter()  # host time to enqueue one call (no synchronisation inside the loop)
    for _ in range(n):
        fn()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    # device time inside this library's kernels (event pairs around every launch; a separate pass)
    lib.asm_profile_enable(1)
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    lib.asm_profile_enable(0)
    kms, kn = (C.c_double * 3)(), (C.c_longlong * 3)()
    lib.asm_profile_collect(kms, kn, 3)
    ks = " ".join(f"{nm} {kms[i] / n:6.3f} ({kn[i] // n})" for i, nm in enumerate(("K1", "K2", "K3")))
    print(f"{name:60s} {total:8.3f} ms  host {host:6.3f}  kernels: {ks}  sum {sum(kms) / n:6.3f}")


def step():
    a = amp.clone().requires_grad_(True)
    p = phs.clone().requires_grad_(True)
    c = fixed.propagate_AP2C_backward(a, p)                      # F-6
    poh = torch.angle(c)                                         # stand-in for the conv + double-phase tail
    s_hat = fixed.propagate_POH2Freq_forward(poh)                # F-7
    s_tgt = multi.filter_AP2filteredFreq(amp, phs / 6.28)        # F-13
    a2, q2 = multi.propagate_multiple_samples_with_random_fixed_multiple_distances_freq2amp(
        torch.cat([s_hat, s_tgt], 0))                            # F-12
    loss = (a2[:B] - a2[B:]).pow(2).mean() + (torch.sin(q2[:B]) - torch.sin(q2[B:])).pow(2).mean()
    loss.backward()
    return loss


from learned_hologram_gan_b200 import loss_func as LF  # noqa: E402
from learned_hologram_gan_b200.ap2poh_tail import ap2poh_tail  # noqa: E402

conv_w = (0.5 * torch.rand(3, 3, 3, generator=g)).cuda()
conv_w = (conv_w + conv_w.transpose(1, 2)).requires_grad_(True)
conv_b = torch.zeros(3, device="cuda", requires_grad=True)


def step_with_stages(fused=True):
    """The same step with the real AP2POH tail between F-6 and F-7 and watermelon.G_loss's pixel / TV / focal phase
    terms (watermelon.py:418-445); fused=False writes the tail and the losses with the reference's torch ops."""
    a = amp.clone().requires_grad_(True)
    p = phs.clone().requires_grad_(True)
    c = fixed.propagate_AP2C_backward(a, p)
    if fused:
        poh = ap2poh_tail(c, conv_w, conv_b)
    else:
        F = torch.nn.functional
        conv = lambda x: torch.cat([F.conv2d(x[:, i:i + 1], conv_w[i][None, None], conv_b[i:i + 1], padding=1) for i in range(3)], 1)
        mfield = torch.complex(conv(c.real), conv(c.imag))
        am, ph = mfield.abs(), mfield.angle()
        am = am / (am.amax((-2, -1), keepdim=True) * 1.01)
        odd = ((torch.arange(R, device="cuda")[:, None] + torch.arange(R, device="cuda")[None]) % 2).float()
        poh = (1 - odd) * (ph + am.acos()) + odd * (ph - am.acos())
    s_hat = fixed.propagate_POH2Freq_forward(poh)
    s_tgt = multi.filter_AP2filteredFreq(amp, phs / 6.28)
    a2, q2 = multi.propagate_multiple_samples_with_random_fixed_multiple_distances_freq2amp(torch.cat([s_hat, s_tgt], 0))
    hat_a, tgt_a, hat_q, tgt_q = a2[:B], a2[B:].detach(), q2[:B], q2[B:].detach()
    if fused:
        terms = LF.amp_loss_terms(hat_a, tgt_a, 1.0)
        loss = terms[0] + terms[3] + LF.focal_sincos_phase_gradient_loss(hat_q, tgt_q)
    else:
        tv = lambda x: (x[..., :, 1:] - x[..., :, :-1]).abs().mean() + (x[..., 1:, :] - x[..., :-1, :]).abs().mean()
        sf, sr = torch.cat((hat_q.sin(), hat_q.cos()), 1), torch.cat((tgt_q.sin(), tgt_q.cos()), 1)
        d1 = ((sf[..., :, 1:] - sf[..., :, :-1]) - (sr[..., :, 1:] - sr[..., :, :-1])).abs()
        d2 = ((sf[..., 1:, :] - sf[..., :-1, :]) - (sr[..., 1:, :] - sr[..., :-1, :])).abs()
        loss = (torch.nn.functional.mse_loss(hat_a, tgt_a) + (tv(hat_a) - tv(tgt_a)).abs()
                + (d1 * (d1 / d1.max()).detach()).mean() + (d2 * (d2 / d2.max()).detach()).mean())
    conv_w.grad = conv_b.grad = None
    loss.backward()
    return loss


t("F-6  propagate_AP2C_backward (fwd only)", lambda: fixed.propagate_AP2C_backward(amp, phs))
t("F-7  propagate_POH2Freq_forward (fwd only)", lambda: fixed.propagate_POH2Freq_forward(phs))
t("F-13 filter_AP2filteredFreq", lambda: multi.filter_AP2filteredFreq(amp, phs))
spec = fixed.propagate_POH2Freq_forward(phs)
t("F-12 random_fixed_multiple_distances_freq2amp (fwd only)",
  lambda: multi.propagate_multiple_samples_with_random_fixed_multiple_distances_freq2amp(torch.cat([spec, spec], 0)))
t("F-10 multi __call__ D=20 (fwd only)", lambda: multi(amp, phs, multi.distances))
t("config-3 style step (F-6, F-7, F-13, F-12, backward)", step, n=10)
t("  + AP2POH tail + G_loss terms, this package's stages", lambda: step_with_stages(True), n=10)
t("  + AP2POH tail + G_loss terms, torch ops for the stages", lambda: step_with_stages(False), n=10)
print("loss of the two variants:", float(step_with_stages(True)), float(step_with_stages(False)))

if "--profile" in sys.argv:  # where the host time of the step goes
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        step()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
