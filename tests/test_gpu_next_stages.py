"""GPU parity of the stages either side of the path against oracle/next_oracle.py and the
golden vectors made by the unmodified reference.  Gates: bit-exact for the byte / min-max / normalise work,
rel <= 1e-4 on losses and their gradients (the tolerance of BASELINE.json's north_star), 1e-5 on fields."""

import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, load_golden

from oracle import asm_oracle as O
from oracle import next_oracle as NO

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-4


@pytest.fixture(scope="module")
def gold():
    return load_golden("next_small")


@pytest.fixture(scope="module")
def LF():
    from learned_hologram_gan_b200 import loss_func

    return loss_func


def rel(a, b):
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def grad_of(fn, x, *rest):
    x = x.clone().requires_grad_(True)
    y = fn(x, *rest)
    y.backward()
    return y.detach(), x.grad


@pytest.mark.parametrize("tag", ["a", "b"])
def test_losses_match_the_reference_golden(gold, LF, tag):
    alpha = float(gold["alpha"])
    hat, tgt = gold.t(f"{tag}_hat").cuda(), gold.t(f"{tag}_tgt").cuda()
    terms = LF.amp_loss_terms(hat, tgt, alpha)
    for i, key in enumerate(("mse", "tv_hat", "tv_tgt", "tv_loss", "amp_loss")):
        assert rel(terms[i], gold.t(f"{tag}_{key}")) <= LOSS_TOL, key
    loss, g = grad_of(LF.amp_loss, hat, tgt, alpha)
    assert rel(loss, gold.t(f"{tag}_amp_loss")) <= LOSS_TOL
    assert rel(g, gold.t(f"{tag}_amp_loss_grad")) <= LOSS_TOL
    tv, g = grad_of(LF.total_variation, hat)
    assert rel(tv, gold.t(f"{tag}_tv_hat")) <= LOSS_TOL
    assert rel(g, gold.t(f"{tag}_tv_grad")) <= LOSS_TOL
    fl, g = grad_of(LF.focal_sincos_phase_gradient_loss, gold.t(f"{tag}_fake").cuda(), gold.t(f"{tag}_real").cuda())
    assert rel(fl, gold.t(f"{tag}_focal")) <= LOSS_TOL
    assert rel(g, gold.t(f"{tag}_focal_grad")) <= LOSS_TOL


@pytest.mark.parametrize("shape", [(4, 3, 384, 384), (2, 3, 37, 53), (1, 3, 16, 516), (3, 3, 33, 4), (1, 1, 5, 1)])
def test_losses_match_the_oracle(LF, shape):
    gen = torch.Generator().manual_seed(sum(shape))
    hat, tgt = torch.rand(shape, generator=gen), torch.rand(shape, generator=gen)
    # G_loss style combination (watermelon.py:436-439): both weights reach the same backward pass
    def combo(lossmod):
        def f(h, t):
            if lossmod is LF:
                terms = LF.amp_loss_terms(h, t, 0.0)
                return 2.0 * terms[0] + 3.0 * terms[3]
            return 2.0 * torch.nn.functional.mse_loss(h, t) + 3.0 * NO.total_variation_loss(h, t)
        return f
    want, gw = grad_of(combo(NO), hat, tgt)
    got, gg = grad_of(combo(LF), hat.cuda(), tgt.cuda())
    if shape[-1] == 1:  # no horizontal neighbours: mean of an empty tensor is NaN in the reference, and here
        assert torch.isnan(want) and torch.isnan(got.cpu())
        return
    assert rel(got, want) <= LOSS_TOL
    assert rel(gg, gw) <= LOSS_TOL
    fake, real = 7.0 * hat - 0.5, 2 * torch.pi * tgt
    want, gw = grad_of(NO.focal_sincos_phase_gradient_loss, fake, real)
    got, gg = grad_of(LF.focal_sincos_phase_gradient_loss, fake.cuda(), real.cuda())
    assert rel(got, want) <= LOSS_TOL
    assert rel(gg, gw) <= LOSS_TOL
    want, gw = grad_of(NO.phase_sincos_gradient_loss, fake, real)  # loss.py:165-183, the un-weighted variant
    got, gg = grad_of(LF.phase_sincos_gradient_loss, fake.cuda(), real.cuda())
    assert rel(got, want) <= LOSS_TOL
    assert rel(gg, gw) <= LOSS_TOL
    for name in ("focal_sincos_phase_loss", "plain_phase_loss"):  # loss.py:186-208, the point-wise phase losses
        want, gw = grad_of(getattr(NO, name), fake, real)
        got, gg = grad_of(getattr(LF, name), fake.cuda(), real.cuda())
        assert rel(got, want) <= LOSS_TOL, name
        assert rel(gg, gw) <= LOSS_TOL, name


def test_losses_scalar_path_for_unaligned_storage(LF):
    """A contiguous tensor whose storage is only 4-byte aligned goes through the scalar kernels: same values."""
    gen = torch.Generator().manual_seed(11)
    shape = (2, 3, 24, 40)
    n = int(np.prod(shape))
    hat, tgt = torch.rand(shape, generator=gen), torch.rand(shape, generator=gen)
    buf = torch.empty(n + 1, device="cuda")
    off = buf[1:].view(shape)
    off.copy_(hat)
    assert off.data_ptr() % 16 != 0 and off.is_contiguous()
    a = LF.amp_loss_terms(off, tgt.cuda(), 0.5)
    b = LF.amp_loss_terms(hat.cuda(), tgt.cuda(), 0.5)
    assert rel(a, b) <= 1e-6
    _, ga = grad_of(LF.amp_loss, off, tgt.cuda(), 0.5)
    _, gb = grad_of(LF.amp_loss, hat.cuda(), tgt.cuda(), 0.5)
    assert torch.equal(ga, gb)


def test_losses_are_deterministic_and_follow_the_input_device(LF):
    gen = torch.Generator().manual_seed(5)
    hat, tgt = torch.rand(3, 3, 96, 128, generator=gen), torch.rand(3, 3, 96, 128, generator=gen)
    runs = [LF.amp_loss_terms(hat.cuda(), tgt.cuda(), 1.0) for _ in range(5)]
    assert all(torch.equal(runs[0], r) for r in runs[1:])
    f = [LF.focal_sincos_phase_gradient_loss(hat.cuda(), tgt.cuda()) for _ in range(5)]
    assert all(torch.equal(f[0], r) for r in f[1:])
    host = LF.amp_loss(hat, tgt)  # host tensors are staged to the GPU, the result comes back to the host
    assert host.device.type == "cpu" and rel(host, NO.amp_loss(hat, tgt)) <= LOSS_TOL
    same = LF.focal_sincos_phase_gradient_loss(hat.cuda(), hat.cuda())
    assert torch.isnan(same)  # 0/0, as in the reference (loss.py:152-156)


def test_losses_at_the_config4_stack_size(LF):
    """24 x 3 planes of 2160 x 3840: value against torch's own reductions on the device, gradient by its
    closed form properties (sum of the TV gradient over a plane is 0; the mse part is 2(h-t)/N)."""
    gen = torch.Generator(device="cuda").manual_seed(1)
    hat = torch.rand(8, 3, 2160, 3840, device="cuda", generator=gen)
    tgt = torch.rand(8, 3, 2160, 3840, device="cuda", generator=gen)
    terms = LF.amp_loss_terms(hat, tgt, 1.0)
    want_mse = torch.nn.functional.mse_loss(hat, tgt)
    assert rel(terms[0], want_mse) <= LOSS_TOL
    assert rel(terms[1], NO.total_variation(hat)) <= LOSS_TOL
    _, g = grad_of(LF.mse_loss, hat, tgt)
    assert rel(g[0, 0], (2.0 / hat.numel()) * (hat[0, 0] - tgt[0, 0])) <= 1e-6
    _, g = grad_of(LF.total_variation, hat)
    assert abs(float(g[3, 1].double().sum())) <= 1e-9


# ---- N4 ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b"])
def test_export_matches_golden_bit_for_bit(gold, tag):
    from learned_hologram_gan_b200 import focal_stack_export as E

    x = gold.t(f"{tag}_stack")
    mm = E.plane_minmax(x.cuda()).cpu()
    assert torch.equal(mm[:, 0], x.amin(dim=(-2, -1)).reshape(-1)) and torch.equal(mm[:, 1], x.amax(dim=(-2, -1)).reshape(-1))
    assert torch.equal(E.tensor_normalizor_2D(x.cuda()).cpu(), gold.t(f"{tag}_stack_norm"))
    want = NO.focal_stack_u8(x)
    assert np.array_equal(E.focal_stack_to_u8(x.cuda()).cpu().numpy(), want)
    assert np.array_equal(E.focal_stack_to_u8(x.cuda(), alpha_channel=False).cpu().numpy(), want[..., :3])
    norm = NO.tensor_normalizor_2D(x)
    raw = E.focal_stack_to_u8(norm.cuda(), normalize=False).cpu().numpy()
    assert np.array_equal(raw, want)
    assert E.tensor_normalizor_2D(x).device.type == "cpu"


def test_export_reproduces_the_reference_pngs(tmp_path):
    """README.md:123-132 end to end on the CUDA path: poh -> focal stack -> PNG files, against the reference's PNGs."""
    from PIL import Image

    from learned_hologram_gan_b200 import bandLimitedAngularSpectrumMethod_for_multiple_distances as Multi
    from learned_hologram_gan_b200 import focal_stack_export as E

    d = os.path.join(GOLDEN_DIR, "terminalTest")
    poh = torch.from_numpy(np.load(os.path.join(d, "poh.npy"))).unsqueeze(0).cuda()
    z = torch.linspace(4e-4, 10e-4, 10)
    prop = Multi(sample_row_num=384, sample_col_num=384, distances=z, pad_size=320, filter_radius_coefficient=0.35,
                 pixel_pitch=3.74e-6, wave_length=torch.tensor([638e-9, 520e-9, 450e-9]), band_limit=False, cuda=True)
    amp = prop(torch.ones_like(poh), poh, z)
    names = E.save_focal_stack(amp, str(tmp_path))
    assert [os.path.basename(n) for n in names] == [f"{i}.png" for i in range(10)]
    for i, name in enumerate(names):
        got = np.asarray(Image.open(name)).astype(np.int32)
        want = np.asarray(Image.open(os.path.join(d, f"{i}.png"))).astype(np.int32)
        assert got.shape == want.shape == (384, 384, 4)
        diff = np.abs(got - want)
        assert diff.max() <= 1 and (diff > 0).mean() <= 0.01


def test_export_at_the_config4_and_config5_sizes():
    from learned_hologram_gan_b200 import focal_stack_export as E

    gen = torch.Generator(device="cuda").manual_seed(2)
    for shape in ((8, 3, 2160, 3840), (64, 3, 1080, 1920), (2, 3, 1081, 1919)):
        x = 5.0 * torch.rand(shape, device="cuda", generator=gen) + 0.25
        mn, mx = x.amin(dim=(-2, -1), keepdim=True), x.amax(dim=(-2, -1), keepdim=True)
        want = (((x - mn) / (mx - mn)) * 255).to(torch.uint8).permute(0, 2, 3, 1)
        got = E.focal_stack_to_u8(x, alpha_channel=False)
        assert torch.equal(got, want)
        rgba = E.focal_stack_to_u8(x)
        assert torch.equal(rgba[..., :3], want) and bool((rgba[..., 3] == 255).all())


# ---- N2 ----------------------------------------------------------------------------------------------------------
def phasor(p):
    p = torch.as_tensor(p).cpu()
    return torch.polar(torch.ones_like(p), p)


def test_ap2poh_tail_matches_the_reference_golden(gold):
    from learned_hologram_gan_b200.ap2poh_tail import ap2poh_tail

    with torch.no_grad():
        poh, pmax = ap2poh_tail(gold.t("tail_field").cuda(), gold.t("tail_weights").cuda(), gold.t("tail_bias").cuda(),
                                return_plane_max=True)
    assert O.rel_l2(phasor(poh), phasor(gold.t("tail_poh"))) <= 1e-5
    assert poh.shape == gold.t("tail_poh").shape and pmax.shape == (2, 3)


@pytest.mark.parametrize("k,shape", [(3, (4, 3, 384, 384)), (3, (1, 3, 45, 67)), (3, (2, 3, 17, 130)), (3, (1, 3, 1, 1)), (3, (1, 3, 33, 250)), (3, (2, 3, 2, 2)),
                                     (5, (1, 3, 45, 67)), (1, (2, 3, 8, 8)), (7, (1, 3, 20, 24))])
def test_ap2poh_tail_matches_the_oracle(k, shape):
    from learned_hologram_gan_b200.ap2poh_tail import ap2poh_tail

    gen = torch.Generator().manual_seed(k)
    field = torch.complex(torch.randn(shape, generator=gen), torch.randn(shape, generator=gen))
    w = torch.rand(3, k, k, generator=gen)
    w = 0.5 * (w + w.transpose(1, 2))
    b = 0.1 * torch.randn(3, generator=gen)
    want = NO.ap2poh_tail(field, w, b)
    with torch.no_grad():
        got = ap2poh_tail(field.cuda(), w.cuda(), b.cuda())
    assert O.rel_l2(phasor(got), phasor(want)) <= 1e-5


@pytest.mark.parametrize("k,shape", [(3, (4, 3, 384, 384)), (3, (2, 3, 17, 130)), (5, (1, 3, 45, 67)), (1, (2, 3, 8, 8))])
def test_ap2poh_tail_gradients_match_reference_autograd(k, shape):
    """The training step differentiates AP2POH.forward (ap2poh.py:104-116): gradients of a 2*pi-periodic loss of the
    POH with respect to the complex field, the kernels and the biases against torch autograd of the oracle."""
    from learned_hologram_gan_b200.ap2poh_tail import ap2poh_tail

    gen = torch.Generator().manual_seed(100 + k)
    field = torch.complex(torch.randn(shape, generator=gen), torch.randn(shape, generator=gen))
    w = torch.rand(3, k, k, generator=gen)
    w = 0.5 * (w + w.transpose(1, 2))
    b = 0.1 * torch.randn(3, generator=gen)
    r1, r2 = torch.randn(shape, generator=gen), torch.randn(shape, generator=gen)

    def run(fn, dev):
        f, ww, bb = (t.to(dev).clone().requires_grad_(True) for t in (field, w, b))
        poh = fn(f, ww, bb)
        loss = (torch.cos(poh) * r1.to(dev) + torch.sin(poh) * r2.to(dev)).sum()
        loss.backward()
        return loss.detach().cpu(), f.grad.cpu(), ww.grad.cpu(), bb.grad.cpu()

    want = run(NO.ap2poh_tail, "cpu")
    got = run(ap2poh_tail, "cuda")
    assert rel(got[0], want[0]) <= LOSS_TOL
    assert O.rel_l2(got[1], want[1]) <= LOSS_TOL
    assert rel(got[2], want[2]) <= LOSS_TOL
    assert rel(got[3], want[3]) <= LOSS_TOL


# ---- N3 ----------------------------------------------------------------------------------------------------------
def test_bin_reader_batches_are_byte_exact(tmp_path):
    from learned_hologram_gan_b200 import data_loader as DL

    shape = (40, 3, 48, 64)
    files = {}
    for name in ("img", "depth", "amp", "phs"):
        arr = np.random.default_rng(len(name) + 7).random(shape, dtype=np.float32)
        arr.tofile(tmp_path / f"{name}.bin")
        files[name] = arr
    kw = dict(samplesNum=40, channlesNum=3, height=48, width=64, cuda=True)
    ds = DL.dataloaderImgDepthAmpPhs(*(str(tmp_path / f"{n}.bin") for n in ("img", "depth", "amp", "phs")), **kw)
    for idx in ([3], [5, 1, 39, 0, 5, 17, 22, 8], list(range(40))):
        rgbd, amp, phs = ds.fetch(idx)
        want = torch.stack([NO.rgbd_item(files["img"], files["depth"], i) for i in idx])
        assert torch.equal(rgbd.cpu(), want)
        assert torch.equal(amp.cpu(), torch.from_numpy(files["amp"][idx]))
        assert torch.equal(phs.cpu(), torch.from_numpy(files["phs"][idx]))
    item = ds[7]
    assert item[0].is_cuda and torch.equal(item[0].cpu(), NO.rgbd_item(files["img"], files["depth"], 7))
    ds2 = DL.dataloaderAmpPIPhs(str(tmp_path / "amp.bin"), str(tmp_path / "phs.bin"), **kw)
    amp, phs = ds2.fetch([9, 2, 30])
    assert torch.equal(phs.cpu(), torch.stack([NO.pi_phase_item(files["phs"], i) for i in (9, 2, 30)]))
    assert torch.equal(amp.cpu(), torch.from_numpy(files["amp"][[9, 2, 30]]))
    ds3 = DL.dataloaderImgDepth(str(tmp_path / "img.bin"), str(tmp_path / "depth.bin"), **kw)
    assert torch.equal(ds3.fetch([4, 4]).cpu(), torch.stack([NO.rgbd_item(files["img"], files["depth"], 4)] * 2))
    with pytest.raises(IndexError):
        ds.fetch([40])
    assert ds.fetch([])[0].shape == (0, 4, 48, 64)


# ---- the stages composed with the propagation path: a config-3 style generator step -----------------------------
@pytest.mark.parametrize("rows,cols,pad,coef,B,D", [(384, 384, 320, 0.45, 2, 4), (48, 64, 8, 0.45, 2, 3)])
def test_generator_step_composition_vs_oracle(rows, cols, pad, coef, B, D):
    """trainingModel.py's generator step around the path (watermelon.py:219-231,418-445, AP2POH.py:104-116):
    (amp_z, phs_z) -> F-6 -> AP2POH tail -> POH -> F-7 -> F-11 -> amp_loss + focal phase loss, backward to the
    inputs, the symmetric kernels and the biases.  Every arrow is one of this package's autograd Functions; the same
    chain of oracle functions under torch autograd on the CPU is the reference.

    Inputs are smooth fields: with random phases the fields have zeros (min |field| 3e-4, min amplitude 9e-5 at 384^2),
    angle's gradient is ~1/|y| there, and the ORACLE's own gradients move by 5e-4 under 1e-7 relative input noise
    (measured, CPU fp32) -- no implementation can agree with it to 1e-4.  On these inputs the same probe gives 2e-5."""
    from learned_hologram_gan_b200 import angular_spectrum_method as M
    from learned_hologram_gan_b200 import loss_func as LF
    from learned_hologram_gan_b200.ap2poh_tail import ap2poh_tail

    wl = torch.tensor([638e-9, 520e-9, 450e-9])
    gen = torch.Generator().manual_seed(31 + rows)
    kw = dict(sample_row_num=rows, sample_col_num=cols, pad_size=pad, filter_radius_coefficient=coef,
              wave_length=wl, cuda=True)
    zf = torch.tensor([1e-3])
    zs = torch.linspace(-4e-4, 0, D + 1)[:-1].contiguous()
    fixed = M.bandLimitedAngularSpectrumMethod_for_single_fixed_distance(distance=zf, **kw)
    multi = M.bandLimitedAngularSpectrumMethod_for_multiple_distances(distances=zs, **kw)
    g = O.Geometry(rows=rows, cols=cols, pad=pad, radius_coef=coef, wavelengths=wl)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, rows), torch.linspace(0, 1, cols), indexing="ij")
    base = torch.stack([torch.sin(2 * torch.pi * (1 + c) * xx) * torch.cos(2 * torch.pi * (2 - 0.5 * c) * yy)
                        for c in range(3)])
    phs_z = 1.5 * base[None] + 0.2 * torch.rand(B, 3, rows, cols, generator=gen)
    amp_z = 0.6 + 0.2 * torch.cos(3 * base[None] + 1.0) + 0.05 * torch.rand(B, 3, rows, cols, generator=gen)
    w = torch.rand(3, 3, 3, generator=gen)
    w = 0.5 * (w + w.transpose(1, 2))
    b = 0.05 * torch.randn(3, generator=gen)
    tgt_amp = torch.rand(B * D, 3, rows, cols, generator=gen)
    tgt_phs = 2 * torch.pi * torch.rand(B * D, 3, rows, cols, generator=gen) - torch.pi

    def ref_step():
        a, p, ww, bb = (t.clone().requires_grad_(True) for t in (amp_z, phs_z, w, b))
        field = O.fixed_ap2c_backward(g, zf, a, p)
        poh = NO.ap2poh_tail(field, ww, bb)
        spec = O.fixed_poh2freq(g, zf, poh)
        amps, phss = O.multi_all_freq2amp(g, zs, spec)
        loss = NO.amp_loss(amps, tgt_amp, 0.5) + 0.1 * NO.focal_sincos_phase_gradient_loss(phss, tgt_phs)
        loss.backward()
        return loss.detach(), a.grad, p.grad, ww.grad, bb.grad

    def gpu_step():
        a, p, ww, bb = (t.cuda().requires_grad_(True) for t in (amp_z, phs_z, w, b))
        field = fixed.propagate_AP2C_backward(a, p)
        poh = ap2poh_tail(field, ww, bb)
        spec = fixed.propagate_POH2Freq_forward(poh)
        amps, phss = multi.propagate_multiple_samples_with_all_fixed_multiple_distances_freq2amp(spec)
        loss = LF.amp_loss(amps, tgt_amp.cuda(), 0.5) + 0.1 * LF.focal_sincos_phase_gradient_loss(phss, tgt_phs.cuda())
        loss.backward()
        return loss.detach(), a.grad, p.grad, ww.grad, bb.grad

    want, got = ref_step(), gpu_step()
    names = ("loss", "d amp_z", "d phs_z", "d kernels", "d biases")
    for name, x, y in zip(names, got, want):
        assert rel(x, y) <= LOSS_TOL, (name, rel(x, y))


# ---- no write outside the output tensors (compute-sanitizer is not available on the pool: guard bands instead) ----
@pytest.mark.parametrize("shape", [(2, 3, 17, 130), (1, 3, 33, 250), (2, 3, 9, 11), (1, 3, 16, 512), (3, 3, 1, 4)])
def test_outputs_stay_inside_their_tensors(shape):
    """Every output of the C ABI is a slice of a larger buffer filled with a sentinel; after the call the guard bands
    on both sides are untouched and the payload is fully written (no sentinel left)."""
    import ctypes as C

    from learned_hologram_gan_b200 import _cabi_next as N

    lib = N.load()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    planes, rows, cols = shape[0] * shape[1], shape[2], shape[3]
    n = planes * rows * cols
    G = 1024
    gen = torch.Generator(device="cuda").manual_seed(n)
    x = torch.rand(shape, device="cuda", generator=gen)
    y = torch.rand(shape, device="cuda", generator=gen)
    partial = torch.empty(lib.lhg_next_partial_floats(planes, rows, cols), device="cuda")
    SENT = -12345.0

    def guarded(numel, dtype=torch.float32, sentinel=SENT):
        buf = torch.full((numel + 2 * G,), sentinel, dtype=dtype, device="cuda")
        return buf, buf[G:G + numel]

    def check(buf, numel, sentinel=SENT, full=True):
        assert bool((buf[:G] == sentinel).all()) and bool((buf[G + numel:] == sentinel).all())
        if full:
            assert not bool((buf[G:G + numel] == sentinel).any())

    p = lambda t: C.c_void_p(t.data_ptr())
    terms = torch.empty(5, device="cuda")
    N.check(lib.lhg_amp_loss_terms(p(x), p(y), planes, rows, cols, 1.0, p(partial), partial.numel(), p(terms), stream))
    g5 = torch.ones(5, device="cuda")
    buf, out = guarded(n)
    N.check(lib.lhg_amp_loss_backward(p(x), p(y), p(g5), p(terms), 1.0, planes, rows, cols, p(out), stream))
    check(buf, n)
    fterms = torch.empty(4, device="cuda")
    N.check(lib.lhg_focal_phase_loss_terms(p(x), p(y), planes, rows, cols, p(partial), partial.numel(), p(fterms), stream))
    buf, out = guarded(n)
    N.check(lib.lhg_focal_phase_loss_backward(p(x), p(y), p(fterms), p(g5), planes, rows, cols, p(out), stream))
    check(buf, n)
    mm = torch.empty(planes, 2, device="cuda")
    N.check(lib.lhg_plane_minmax(p(x), planes, rows * cols, p(partial), partial.numel(), p(mm), stream))
    buf, out = guarded(n)
    N.check(lib.lhg_normalize_planes(p(x), p(mm), planes, rows * cols, p(out), stream))
    check(buf, n)
    for oc in (3, 4):
        nb = shape[0] * rows * cols * oc
        buf, out = guarded(nb, torch.uint8, 77)
        N.check(lib.lhg_pack_rgb_u8(p(x), p(mm), shape[0], rows, cols, oc, p(out), stream))
        check(buf, nb, 77, full=False)
    field = torch.complex(x, y).contiguous()
    for k in (3, 5):
        w = torch.rand(3, k, k, device="cuda", generator=gen)
        b = torch.zeros(3, device="cuda")
        pmax = torch.empty(planes, device="cuda")
        buf, out = guarded(n)
        N.check(lib.lhg_ap2poh_tail(p(field), p(w), p(b), k, planes, rows, cols, p(partial), partial.numel(), p(pmax),
                                    p(out), stream))
        check(buf, n)
        need = lib.lhg_ap2poh_tail_backward_floats(k, planes, rows, cols)
        work = torch.empty(need, device="cuda")
        gbuf, gout = guarded(2 * n)
        gw, gb = torch.empty(3, k, k, device="cuda"), torch.empty(3, device="cuda")
        N.check(lib.lhg_ap2poh_tail_backward(p(field), p(w), p(b), k, p(x), planes, rows, cols, p(work), need, p(gout),
                                             p(gw), p(gb), stream))
        check(gbuf, 2 * n)
    torch.cuda.synchronize()


def test_utilities_normalizers_run_on_the_stage_kernels(gold):
    """learnedMethodForHologram.utilities.tensor_normalizor_2D / amplitude_normalizor (the import shim's module when
    the reference's own utilities cannot be imported): bit-identical to the oracle, through next_stages.cu."""
    from learned_hologram_gan_b200 import _cabi_next as N
    from learned_hologram_gan_b200 import utilities as U

    lib = N.load()
    for tag in ("a", "b"):
        x = gold.t(f"{tag}_stack")
        n0 = lib.lhg_next_launch_count()
        got = U.tensor_normalizor_2D(x.cuda())
        amp = U.amplitude_normalizor(x.abs().cuda())
        assert lib.lhg_next_launch_count() - n0 >= 4  # two reductions + two affine passes, all ours
        assert torch.equal(got.cpu(), gold.t(f"{tag}_stack_norm"))
        assert torch.equal(amp.cpu(), NO.amplitude_normalizor(x.abs()))
    a = gold.t("a_stack").abs().cuda().requires_grad_(True)
    U.amplitude_normalizor(a).sum().backward()  # training keeps torch's graph
    assert a.grad is not None and a.grad.shape == a.shape
