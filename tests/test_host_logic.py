"""Host-side logic that runs without a GPU: the reference-faithful grids, the import shim, the
geometry of the drop-in constructors' arguments, and the plane partition of the sharded path."""

import os
import sys

import pytest
import torch

from conftest import ROOT

from learned_hologram_gan_b200 import engine as E
from learned_hologram_gan_b200.sharding import Segment, plane_shards


def test_host_grids_are_bit_identical_to_the_reference(golden):
    prow = int(golden["rows"]) + 2 * int(golden["pad"])
    pcol = int(golden["cols"]) + 2 * int(int(golden["pad"]) * (int(golden["cols"]) / int(golden["rows"])))
    radius = min(prow, pcol) * float(golden["coef"])
    wm = E.host_wm_grid(prow, pcol, float(golden["pitch"]), golden.t("wavelengths"), radius)
    assert torch.equal(wm.abs(), golden.t("w_grid"))
    for c in range(3):
        assert torch.equal((~torch.signbit(wm[c])).float(), golden.t("mask"))
    assert torch.equal(E.host_radial_grid(prow, pcol), golden.t("soft_grid"))


def test_import_lines_of_the_reference_resolve_to_the_b200_module():
    import learnedMethodForHologram.utilities  # noqa: F401  (the reference test imports only this)
    import learnedMethodForHologram

    mod = learnedMethodForHologram.angular_spectrum_method
    assert mod.__name__ == "learned_hologram_gan_b200.angular_spectrum_method"
    from learnedMethodForHologram.angular_spectrum_method import (  # noqa: F401
        bandLimitedAngularSpectrumMethod,
        bandLimitedAngularSpectrumMethod_for_multiple_distances as BLASM_v4,
        bandLimitedAngularSpectrumMethod_for_single_fixed_distance as fixed_distance_propogator,
    )
    for name in ("try_gpu", "phase_tensor_generator", "tensor_normalizor_2D",
                 "generate_circular_frequency_mask", "prepare_circular_frequency_mask_grid"):
        assert hasattr(learnedMethodForHologram.utilities, name)


def test_constructor_signatures_match_the_reference_surface():
    import inspect

    import learned_hologram_gan_b200.angular_spectrum_method as m

    base = list(inspect.signature(m.bandLimitedAngularSpectrumMethod.__init__).parameters)
    assert base == ["self", "sample_row_num", "sample_col_num", "pad_size", "filter_radius_coefficient",
                    "pixel_pitch", "wave_length", "band_limit", "cuda"]
    fixed = list(inspect.signature(
        m.bandLimitedAngularSpectrumMethod_for_single_fixed_distance.__init__).parameters)
    assert fixed == base + ["distance"]
    multi = inspect.signature(m.bandLimitedAngularSpectrumMethod_for_multiple_distances.__init__).parameters
    assert list(multi) == ["self", "sample_row_num", "sample_col_num", "distances", "pad_size",
                           "filter_radius_coefficient", "pixel_pitch", "wave_length", "band_limit", "cuda"]
    assert multi["pad_size"].default == 160 and multi["cuda"].default is True
    call = list(inspect.signature(m.bandLimitedAngularSpectrumMethod.__call__).parameters)
    assert call == ["self", "amplitute_tensor", "phase_tensor", "distances"]  # sic, asm.py:70


def test_no_cpu_fallback():
    """Without a CUDA device the product path must refuse to run, not silently compute on the host."""
    if torch.cuda.is_available():
        pytest.skip("needs a GPU-less host")
    import learned_hologram_gan_b200.angular_spectrum_method as m

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.bandLimitedAngularSpectrumMethod(sample_row_num=64, sample_col_num=64)
    src = open(os.path.join(ROOT, "learned_hologram_gan_b200", "engine.py")).read()
    assert "oracle" not in src and "torch.fft.fft2" not in src and "torch.fft.ifft2" not in src
    # the stages either side of the path: no module of the package imports the oracle or transforms with torch.fft,
    # and every stage refuses to run without a CUDA device
    import glob
    import re

    for path in glob.glob(os.path.join(ROOT, "learned_hologram_gan_b200", "*.py")):
        text = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), path
        assert not re.search(r"torch\.fft\.(i?r?fft2?|i?fftn|hfft|ihfft)\b", text), path  # fftfreq (host grids) is fine
    from learned_hologram_gan_b200 import ap2poh_tail, focal_stack_export, loss_func

    x = torch.rand(1, 3, 8, 8)
    for call in (lambda: loss_func.amp_loss(x, x), lambda: loss_func.focal_sincos_phase_gradient_loss(x, x),
                 lambda: focal_stack_export.focal_stack_to_u8(x), lambda: focal_stack_export.tensor_normalizor_2D(x),
                 lambda: ap2poh_tail.ap2poh_tail(torch.complex(x, x), torch.rand(3, 3, 3), torch.zeros(3))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_plane_shards_cover_every_plane_once(world):
    n_colour, n_depth = 3, 8
    seen = []
    for rank in range(world):
        for seg in plane_shards(n_colour, n_depth, world, rank):
            assert 0 <= seg.colour < n_colour and 0 <= seg.d0 < seg.d1 <= n_depth
            seen += [(seg.colour, d) for d in range(seg.d0, seg.d1)]
    assert sorted(seen) == [(c, d) for c in range(n_colour) for d in range(n_depth)]
    sizes = [sum(s.n_depth for s in plane_shards(n_colour, n_depth, world, r)) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
    assert plane_shards(3, 8, 8, 2) == [Segment(0, 6, 8), Segment(1, 0, 1)]


def test_every_runtime_knob_is_documented():
    """INTEGRATION.md section 5 lists every LHG_* environment variable the library, the Python face and bench.py read."""
    import glob
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = (glob.glob(os.path.join(root, "learned_hologram_gan_b200", "csrc", "*.cu*"))
             + glob.glob(os.path.join(root, "learned_hologram_gan_b200", "*.py"))
             + glob.glob(os.path.join(root, "learnedMethodForHologram", "*.py")) + [os.path.join(root, "bench.py")])
    used = set()
    for f in files:
        with open(f) as fh:
            src = fh.read()
        used |= set(re.findall(r'getenv\("(LHG_[A-Z0-9_]+)"\)', src))
        used |= set(re.findall(r'environ(?:\.get\(|\[)"(LHG_[A-Z0-9_]+)"', src))
    assert used, "no knob found: the patterns are stale"
    with open(os.path.join(root, "INTEGRATION.md")) as fh:
        doc = fh.read()
    missing = sorted(k for k in used if k not in doc)
    assert not missing, f"undocumented run-time knobs: {missing}"
