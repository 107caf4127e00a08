"""CPU tests for the 8(f) stages: the oracle (oracle/next_oracle.py) pinned to the reference, the C-ABI header of
the stages, and the host-side gather of the .bin reader (no GPU needed)."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT, load_golden

from oracle import asm_oracle as O
from oracle import next_oracle as NO
from oracle import ref_shim

from learned_hologram_gan_b200 import _cabi_next


@pytest.fixture(scope="module")
def gold():
    return load_golden("next_small")


def same(a, b, tol=0.0):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if tol == 0.0:
        assert torch.equal(a, b)
    else:
        assert O.rel_l2(a, b) <= tol, O.rel_l2(a, b)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_losses_match_golden(gold, tag):
    alpha = float(gold["alpha"])
    hat = gold.t(f"{tag}_hat").requires_grad_(True)
    tgt = gold.t(f"{tag}_tgt")
    loss = NO.amp_loss(hat, tgt, alpha)
    loss.backward()
    same(loss.detach(), gold.t(f"{tag}_amp_loss"))
    same(hat.grad, gold.t(f"{tag}_amp_loss_grad"))
    same(NO.total_variation(hat.detach()), gold.t(f"{tag}_tv_hat"))
    same(NO.total_variation_loss(hat.detach(), tgt), gold.t(f"{tag}_tv_loss"))
    fake = gold.t(f"{tag}_fake").requires_grad_(True)
    fl = NO.focal_sincos_phase_gradient_loss(fake, gold.t(f"{tag}_real"))
    fl.backward()
    same(fl.detach(), gold.t(f"{tag}_focal"))
    same(fake.grad, gold.t(f"{tag}_focal_grad"))
    same(NO.tensor_normalizor_2D(gold.t(f"{tag}_stack")), gold.t(f"{tag}_stack_norm"))


def test_oracle_ap2poh_tail_matches_golden(gold):
    poh = NO.ap2poh_tail(gold.t("tail_field"), gold.t("tail_weights"), gold.t("tail_bias"))
    # same torch ops in the same order as AP2POH.forward: identical up to conv2d's algorithm choice
    same(torch.polar(torch.ones_like(poh), poh), torch.polar(torch.ones_like(poh), gold.t("tail_poh")), tol=1e-6)


def test_oracle_imsave_bytes_reproduce_the_reference_pngs():
    """README.md:123-132: the PNGs the reference wrote with plt.imsave are RGBA with alpha 255 and equal
    (x*255).astype(uint8) of the normalised focal stack to 1 LSB (the oracle's restatement of matplotlib 3.8.1)."""
    from PIL import Image

    d = os.path.join(GOLDEN_DIR, "terminalTest")
    poh = torch.from_numpy(np.load(os.path.join(d, "poh.npy"))).unsqueeze(0)
    g = O.Geometry(rows=384, cols=384, pad=320, radius_coef=0.35, pitch=3.74e-6,
                   wavelengths=torch.tensor([638e-9, 520e-9, 450e-9]))
    amp = O.multi_call(g, torch.ones_like(poh), poh, torch.linspace(4e-4, 10e-4, 10))
    got = NO.focal_stack_u8(amp)
    for i in range(10):
        img = Image.open(os.path.join(d, f"{i}.png"))
        assert img.mode == "RGBA"
        want = np.asarray(img).astype(np.int32)
        assert (want[:, :, 3] == 255).all() and (got[i][:, :, 3] == 255).all()
        diff = np.abs(got[i].astype(np.int32) - want)
        assert diff.max() <= 1 and (diff > 0).mean() <= 0.01


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present (GPU box)")
def test_oracle_against_live_reference(tmp_path):
    ref = ref_shim.load_next()
    L, U, DL = ref["loss_func"], ref["utilities"], ref["data_loader"]
    gen = torch.Generator().manual_seed(3)
    a, b = torch.rand(2, 3, 17, 23, generator=gen), torch.rand(2, 3, 17, 23, generator=gen)
    same(NO.amp_loss(a, b, 0.3), L.amp_loss(a, b, 0.3))
    same(NO.total_variation_loss(a, b), L.total_variation_loss(a, b))
    same(NO.focal_sincos_phase_gradient_loss(6 * a, 6 * b), L.focal_sincos_phase_gradient_loss(6 * a, 6 * b))
    same(NO.phase_sincos_gradient_loss(6 * a, 6 * b), L.phase_sincos_gradient_loss(6 * a, 6 * b))
    same(NO.focal_sincos_phase_loss(6 * a, 6 * b), L.focal_sincos_phase_loss(6 * a, 6 * b))
    same(NO.plain_phase_loss(6 * a, 6 * b), L.plain_phase_loss(6 * a, 6 * b))
    fa = (6 * a).clone().requires_grad_(True)
    fb = (6 * a).clone().requires_grad_(True)
    (NO.focal_sincos_phase_loss(fa, 6 * b) + 2 * NO.plain_phase_loss(fa, 6 * b)).backward()
    (L.focal_sincos_phase_loss(fb, 6 * b) + 2 * L.plain_phase_loss(fb, 6 * b)).backward()
    same(fa.grad, fb.grad)
    same(NO.tensor_normalizor_2D(a), U.tensor_normalizor_2D(a))
    same(NO.amplitude_normalizor(a), U.amplitude_normalizor(a))
    same(NO.checkerboard(6, 9, True), U.generate_checkerboard_mask(6, 9, 1, True))
    same(NO.checkerboard(6, 9, False), U.generate_checkerboard_mask(6, 9, 1, False))
    shape = (5, 3, 6, 8)
    files = {}
    for name in ("img", "depth", "amp", "phs"):
        arr = np.random.default_rng(len(name)).random(shape, dtype=np.float32)
        arr.tofile(tmp_path / f"{name}.bin")
        files[name] = arr
    kw = dict(samplesNum=5, channlesNum=3, height=6, width=8, cuda=False)
    ds = DL.dataloaderImgDepthAmpPhs(*(str(tmp_path / f"{n}.bin") for n in ("img", "depth", "amp", "phs")), **kw)
    rgbd, amp, phs = ds[3]
    same(rgbd, NO.rgbd_item(files["img"], files["depth"], 3))
    same(amp, torch.tensor(files["amp"][3]))
    ds2 = DL.dataloaderAmpPIPhs(str(tmp_path / "amp.bin"), str(tmp_path / "phs.bin"), **kw)
    same(ds2[1][1], NO.pi_phase_item(files["phs"], 1))


# ---- the C ABI of the stages -----------------------------------------------------------------------------------
def declared():
    text = open(os.path.join(ROOT, "include", "lhg_next_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lhg_[a-z0-9_]+)\s*\(", text)))


def test_next_header_binding_and_library_agree():
    names = declared()
    assert names and sorted(_cabi_next.EXPORTS) == names
    lib = _cabi_next.load()
    for name in names:
        assert hasattr(lib, name), name
    header = open(os.path.join(ROOT, "include", "lhg_next_b200.h")).read()
    assert lib.lhg_next_version() == int(re.search(r"#define LHG_NEXT_VERSION (\d+)", header).group(1))
    assert lib.lhg_next_partial_floats(3, 16, 128) == 3 * 2 * 6 + 148 * 16  # focal strips (124 columns per CTA, 6 partials) + staged doubles
    assert lib.lhg_next_partial_floats(0, 16, 128) == 0


def test_next_arguments_are_validated_without_a_device():
    lib = _cabi_next.load()
    before = lib.lhg_next_launch_count()
    assert lib.lhg_pack_rgb_u8(None, None, 1, 4, 4, 5, None, None) == -1
    assert b"out_channels" in lib.lhg_next_last_error()
    assert lib.lhg_amp_loss_terms(None, None, 1, 4, 4, 0.0, None, 0, None, None) == -1
    assert lib.lhg_ap2poh_tail(None, None, None, 4, 3, 4, 4, None, 0, None, None, None) == -1
    assert b"kernel size" in lib.lhg_next_last_error()
    assert lib.lhg_plane_minmax(None, 0, 16, None, 0, None, None) == 0  # empty batch: nothing to do
    assert lib.lhg_next_launch_count() == before


def test_bin_gather_is_byte_exact(tmp_path):
    """dl.py:39-54 host side: dst[i] = file[idx[i]] (whole items, and the first plane only)."""
    lib = _cabi_next.load()
    rng = np.random.default_rng(0)
    data = rng.random((37, 3, 40, 56), dtype=np.float32)
    path = tmp_path / "x.bin"
    data.tofile(path)
    mm = np.memmap(path, dtype=np.float32, mode="r", shape=data.shape)
    item = data[0].nbytes
    for n, threads in ((1, 0), (9, 1), (64, 4), (400, 0)):
        idx = rng.integers(0, 37, size=n).astype(np.int64)
        for copy in (item, item // 3):
            dst = np.zeros((n, copy // 4), dtype=np.float32)
            rc = lib.lhg_bin_gather(ctypes.c_void_p(mm.ctypes.data), 37, item, copy,
                                    ctypes.c_void_p(idx.ctypes.data), n, ctypes.c_void_p(dst.ctypes.data), threads)
            assert rc == 0
            want = data[idx].reshape(n, -1)[:, : copy // 4]
            assert dst.tobytes() == np.ascontiguousarray(want).tobytes()
    bad = np.array([0, 37], dtype=np.int64)
    dst = np.zeros((2, item // 4), dtype=np.float32)
    assert lib.lhg_bin_gather(ctypes.c_void_p(mm.ctypes.data), 37, item, item, ctypes.c_void_p(bad.ctypes.data), 2,
                              ctypes.c_void_p(dst.ctypes.data), 0) == -1
    assert b"out of range" in lib.lhg_next_last_error()


def test_data_loader_items_match_the_oracle(tmp_path):
    from learned_hologram_gan_b200 import data_loader as DL

    shape = (6, 3, 8, 12)
    files = {}
    for name in ("img", "depth", "amp", "phs"):
        arr = np.random.default_rng(len(name) + 1).random(shape, dtype=np.float32)
        arr.tofile(tmp_path / f"{name}.bin")
        files[name] = arr
    kw = dict(samplesNum=6, channlesNum=3, height=8, width=12, cuda=False)
    ds = DL.dataloaderImgDepthAmpPhs(*(str(tmp_path / f"{n}.bin") for n in ("img", "depth", "amp", "phs")), **kw)
    assert len(ds) == 6
    rgbd, amp, phs = ds[4]
    same(rgbd, NO.rgbd_item(files["img"], files["depth"], 4))
    same(amp, torch.tensor(files["amp"][4]))
    same(phs, torch.tensor(files["phs"][4]))
    with pytest.raises(IndexError):
        ds[6]
    ds2 = DL.dataloaderAmpPIPhs(str(tmp_path / "amp.bin"), str(tmp_path / "phs.bin"), **kw)
    same(ds2[2][1], NO.pi_phase_item(files["phs"], 2))
    ds3 = DL.dataloaderImgDepth(str(tmp_path / "img.bin"), str(tmp_path / "depth.bin"), **kw)
    same(ds3[0], NO.rgbd_item(files["img"], files["depth"], 0))


def test_symmetric_kernels_follow_the_reference_parametrisation():
    """nn.py:35-73: one parameter per squared distance from the centre, kernel = params[distance_map]."""
    from learned_hologram_gan_b200.ap2poh_tail import symmetric_kernels

    class Sub:
        def __init__(self, seed):
            g = torch.Generator().manual_seed(seed)
            self.params = torch.rand(3, generator=g, requires_grad=True)  # distances 0, 1, 2 of a 3 x 3 kernel
            self.bias = torch.rand(1, generator=g, requires_grad=True)
            self.distance_map = torch.tensor([[2, 1, 2], [1, 0, 1], [2, 1, 2]])

    class Conv:
        conv_r, conv_g, conv_b = Sub(1), Sub(2), Sub(3)

    w, b = symmetric_kernels(Conv)
    assert w.shape == (3, 3, 3) and b.shape == (3,) and not w.requires_grad
    for c, sub in enumerate((Conv.conv_r, Conv.conv_g, Conv.conv_b)):
        assert torch.equal(w[c], sub.params.detach()[sub.distance_map]) and torch.equal(w[c], w[c].T)
        assert b[c] == sub.bias.detach()[0]
    w, b = symmetric_kernels(Conv, differentiable=True)
    (w.sum() + b.sum()).backward()
    assert torch.equal(Conv.conv_r.params.grad, torch.tensor([1.0, 4.0, 4.0]))  # multiplicity of each distance
    if ref_shim.available():
        ref = ref_shim.load_next()["neural_network_components"]
        torch.manual_seed(0)
        conv = ref.ChannelWiseSymmetricConv(kernel_size=3, padding=1)
        w, b = symmetric_kernels(conv)
        x = torch.rand(2, 3, 9, 11)
        same(NO.channelwise_symmetric_conv(x, w, b), conv(x).detach(), tol=1e-6)
