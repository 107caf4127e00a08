"""UNMODIFIED reference caller code on top of the drop-in module (north_star "drops into
trainingModel.py, generatePOH.py and tests/test_angular_spectrum_method.py unchanged").

The reference files are the offline install under baseline/_ref (made by __graft_entry__.build() from
/root/reference, git-ignored, travels to the GPU box) or /root/reference itself where it exists.  Everything in
this file is fixture set-up: a matplotlib stand-in (absent from the image; its imsave is pinned by the reference's
own PNGs in test_next_oracle_pinned.py), the synthetic PNG the reference test opens, argparse values of the README
command.  The code that RUNS is the reference's.
"""

import ast
import os
import runpy
import sys
import types

import numpy as np
import pytest
import torch

from oracle import asm_oracle as O
from oracle import next_oracle as NO
from oracle import ref_shim

from conftest import GOLDEN_DIR, ROOT

pytestmark = pytest.mark.gpu

CALLERS = [os.path.join(ROOT, "baseline", "_ref", "_callers"), "/root/reference"]


def caller_file(rel):
    for base in CALLERS:
        path = os.path.join(base, rel)
        if os.path.isfile(path):
            return path
    pytest.skip(f"reference caller {rel} is staged neither under baseline/_ref/_callers nor /root/reference")


@pytest.fixture()
def overlay(monkeypatch):
    """learnedMethodForHologram = this repo's propagation module + the reference's own other modules."""
    if not ref_shim.available():
        pytest.skip("no reference tree")
    if "matplotlib" not in sys.modules:  # stand-in: only imsave is ever reached (utilities.py:140-151)
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")

        def imsave(path, arr, **_kw):
            from PIL import Image

            Image.fromarray(NO.imsave_bytes(np.asarray(arr)), mode="RGBA").save(path)

        plt.imsave = imsave
        plt.show = lambda *a, **k: None
        mpl.pyplot = plt
        monkeypatch.setitem(sys.modules, "matplotlib", mpl)
        monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    for name in [n for n in sys.modules if n == "learnedMethodForHologram" or n.startswith("learnedMethodForHologram.")]:
        monkeypatch.delitem(sys.modules, name)
    from learned_hologram_gan_b200 import overlay as ov

    pkg = ov.install(reference_root=ref_shim.REFERENCE_ROOT)
    import learned_hologram_gan_b200.angular_spectrum_method as ours

    assert sys.modules["learnedMethodForHologram.angular_spectrum_method"] is ours
    yield pkg
    for name in [n for n in sys.modules if n == "learnedMethodForHologram" or n.startswith("learnedMethodForHologram.")]:
        sys.modules.pop(name, None)


def test_the_reference_test_file_runs_as_a_file(overlay, tmp_path, monkeypatch):
    """tests/test_angular_spectrum_method.py:6-31 executed with runpy (its __main__ block calls test()): a
    2400 x 4094 PNG at the CWD-relative path it opens, base class, keyword arguments, band_limit=True, CPU tensors.
    The file asserts nothing, so the value it hands to tensor_normalizor_2D is recorded and compared here."""
    from PIL import Image

    path = caller_file(os.path.join("tests", "test_angular_spectrum_method.py"))
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, size=(2400, 4094, 3), dtype=np.uint8)
    (tmp_path / "data" / "images").mkdir(parents=True)
    Image.fromarray(img, mode="RGB").save(tmp_path / "data" / "images" / "sample_hologram.png")
    monkeypatch.chdir(tmp_path)
    util = sys.modules["learnedMethodForHologram.utilities"]
    seen = {}
    real = util.tensor_normalizor_2D

    def spy(t):
        seen["intensities"] = t
        seen["normalized"] = real(t)
        return seen["normalized"]

    monkeypatch.setattr(util, "tensor_normalizor_2D", spy)
    with pytest.raises(SystemExit) as ex:
        runpy.run_path(path, run_name="__main__")
    assert ex.value.code in (None, 0)
    got = seen["intensities"]
    assert tuple(got.shape) == (4, 3, 2400, 4094) and got.device.type == "cpu"
    assert float(seen["normalized"].min()) == 0.0 and float(seen["normalized"].max()) == 1.0
    wl = torch.tensor([639e-9, 515e-9, 473e-9])
    phase = torch.from_numpy(img.copy()).permute(2, 0, 1).contiguous().to(torch.float32).div(255) * 2 * torch.pi
    g = O.Geometry(rows=2400, cols=4094, pad=0, radius_coef=0.5, wavelengths=wl)
    want = O.base_call(g, torch.ones_like(phase), phase, torch.linspace(-1e-3, 2.5e-3, 4))
    assert O.rel_l2(got, want) <= 1e-5


def test_unmodified_ap2poh_forward_on_the_overlay(overlay):
    """AP2POH.py:16-116 instantiated as is (cuda=True): its propagator is this repo's class, its conv / masks /
    normaliser are the reference's.  Against the SAME class loaded beside it with the reference's own propagator
    on the CPU (ref_shim alias), same weights."""
    import importlib

    mod = importlib.import_module("learnedMethodForHologram.watermelon_hologram.AP2POH")
    assert mod.fixed_distance_propogator.__module__ == "learned_hologram_gan_b200.angular_spectrum_method"
    ref_mod = ref_shim.load_next()["AP2POH"]
    assert os.path.samefile(os.path.dirname(mod.__file__), os.path.dirname(ref_mod.__file__))
    R, C, B = 96, 128, 2
    kw = dict(input_shape=(1, 6, R, C), pad_size=48, filter_radius_coefficient=0.45, pixel_pitch=3.74e-6,
              wave_length=torch.tensor([638e-9, 520e-9, 450e-9]), distance=torch.tensor([1e-3]), kernel_size=3)
    torch.manual_seed(5)
    ours = mod.AP2POH(cuda=True, **kw)
    ref = ref_mod.AP2POH(cuda=False, **kw)
    ref.load_state_dict({k: v.cpu() for k, v in ours.state_dict().items()})
    gen = torch.Generator().manual_seed(9)
    amp = torch.rand(B, 3, R, C, generator=gen)
    phs = 2 * torch.pi * torch.rand(B, 3, R, C, generator=gen)
    a_g = amp.cuda().requires_grad_(True)
    p_g = phs.cuda().requires_grad_(True)
    poh = ours(a_g, p_g)
    a_r = amp.clone().requires_grad_(True)
    p_r = phs.clone().requires_grad_(True)
    poh_ref = ref(a_r, p_r)
    assert poh.device.type == "cuda" and tuple(poh.shape) == (B, 3, R, C)
    # POH = angle +- acos(amplitude): compared on the unit circle (angle wraps at +-pi)
    assert O.rel_l2(torch.polar(torch.ones_like(poh_ref), poh.detach().cpu()),
                    torch.polar(torch.ones_like(poh_ref), poh_ref.detach())) <= 1e-4
    # gradients flow to both inputs through the adjoint (watermelon.py trains the UNet through this)
    w = torch.rand(B, 3, R, C, generator=gen)
    (torch.cos(poh) * w.cuda()).sum().backward()
    (torch.cos(poh_ref) * w).sum().backward()
    assert O.rel_l2(a_g.grad.cpu(), a_r.grad) <= 2e-3  # acos' / the per-plane max amplify the 1e-5 field error
    assert O.rel_l2(p_g.grad.cpu(), p_r.grad) <= 2e-3


def test_generatepoh_propagate_block_reproduces_the_readme_pngs(overlay, tmp_path):
    """generatePOH.py:51-78 (the body of `if args.propagate:`), cut out of the staged file with ast and executed
    unmodified on the README's inputs (README.md:123-132: poh.pt of sample 99, pad 320, coefficient 0.35,
    10 planes in [0.4, 1.0] mm): writes 0..9.png, which must equal the reference's own PNGs to 1 LSB."""
    from PIL import Image

    path = caller_file("generatePOH.py")
    tree = ast.parse(open(path).read())
    main = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main")
    block = next(n for n in main.body if isinstance(n, ast.If) and ast.unparse(n.test) == "args.propagate")
    code = compile(ast.Module(body=block.body, type_ignores=[]), path, "exec")
    util = sys.modules["learnedMethodForHologram.utilities"]
    from learnedMethodForHologram.angular_spectrum_method import (
        bandLimitedAngularSpectrumMethod_for_multiple_distances as BLASM_v4,
    )

    d = os.path.join(GOLDEN_DIR, "terminalTest")
    poh = torch.from_numpy(np.load(os.path.join(d, "poh.npy"))).unsqueeze(0).cuda()
    args = types.SimpleNamespace(sample_row_num=384, sample_col_num=384, pad_size=320, min_distance=4e-4,
                                 max_distance=10e-4, num_intervals=10, filter_radius_coefficient=0.35,
                                 pixel_pitch=3.74e-6, wave_length=[638e-9, 520e-9, 450e-9],
                                 output_image_dir=str(tmp_path / "out"))
    exec(code, {"torch": torch, "utilities": util, "BLASM_v4": BLASM_v4, "args": args, "POH": poh, "print": print})
    for i in range(10):
        want = np.asarray(Image.open(os.path.join(d, f"{i}.png")).convert("RGB")).astype(np.int32)
        got = np.asarray(Image.open(tmp_path / "out" / f"{i}.png").convert("RGB")).astype(np.int32)
        diff = np.abs(got - want)
        assert diff.max() <= 1, (i, diff.max())
        assert (diff > 0).mean() <= 0.01, (i, (diff > 0).mean())
