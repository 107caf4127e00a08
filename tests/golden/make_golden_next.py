"""Golden fixtures for the stages either side of the path, made by the UNMODIFIED reference.

Run here (the container that has /root/reference):  python tests/golden/make_golden_next.py
Imports the reference's loss_func / utilities / AP2POH / neural_network_components through oracle/ref_shim.py,
runs them on small seeded CPU inputs and stores inputs, outputs and reference-autograd gradients in
next_small.npz.  matplotlib is absent in this image, so the 8-bit export has no live reference here: it is pinned
by the reference's own PNGs (terminalTest/0..9.png, already in this directory).  Nothing here is reference source.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_shim  # noqa: E402

SHAPES = {"a": (2, 3, 20, 28), "b": (1, 3, 9, 11)}  # cols % 4 == 0 (16-byte path) and an odd width (scalar path)
ALPHA = 0.7


def main():
    ref = ref_shim.load_next()
    L, U = ref["loss_func"], ref["utilities"]
    out = {"alpha": np.float32(ALPHA)}
    for tag, shape in SHAPES.items():
        gen = torch.Generator().manual_seed(122731 + len(tag) + shape[-1])
        hat = torch.rand(shape, generator=gen).requires_grad_(True)
        tgt = torch.rand(shape, generator=gen)
        loss = L.amp_loss(hat, tgt, ALPHA)
        loss.backward()
        out.update({f"{tag}_hat": hat.detach(), f"{tag}_tgt": tgt, f"{tag}_amp_loss": loss.detach(),
                    f"{tag}_amp_loss_grad": hat.grad.clone(),
                    f"{tag}_mse": torch.nn.functional.mse_loss(hat.detach(), tgt),
                    f"{tag}_tv_hat": L.total_variation(hat.detach()), f"{tag}_tv_tgt": L.total_variation(tgt),
                    f"{tag}_tv_loss": L.total_variation_loss(hat.detach(), tgt)})
        hat.grad = None
        L.total_variation(hat).backward()
        out[f"{tag}_tv_grad"] = hat.grad.clone()
        fake = (2 * torch.pi * torch.rand(shape, generator=gen)).requires_grad_(True)
        real = 2 * torch.pi * torch.rand(shape, generator=gen)
        fl = L.focal_sincos_phase_gradient_loss(fake, real)
        fl.backward()
        out.update({f"{tag}_fake": fake.detach(), f"{tag}_real": real, f"{tag}_focal": fl.detach(),
                    f"{tag}_focal_grad": fake.grad.clone()})
        amp = 3.0 * torch.rand(shape, generator=gen) - 0.5
        out.update({f"{tag}_stack": amp, f"{tag}_stack_norm": U.tensor_normalizor_2D(amp)})

    # AP2POH tail: the reference module itself (CPU), its own random symmetric kernels
    torch.manual_seed(122731)
    net = ref["AP2POH"].AP2POH(input_shape=(1, 6, 24, 32), cuda=False, pad_size=12, filter_radius_coefficient=0.45,
                               distance=torch.tensor([1e-3]))
    for sub, b in zip((net.part1.conv_r, net.part1.conv_g, net.part1.conv_b), (0.01, -0.02, 0.03)):
        sub.bias.data.fill_(b)
    gen = torch.Generator().manual_seed(7)
    amp_z = torch.rand(2, 3, 24, 32, generator=gen)
    phs_z = 2 * torch.pi * torch.rand(2, 3, 24, 32, generator=gen)
    with torch.no_grad():
        field = net.propagator.propagate_AP2C_backward(amp_z, phs_z)
        poh = net(amp_z, phs_z)
    ws = torch.stack([s.params.detach()[s.distance_map] for s in (net.part1.conv_r, net.part1.conv_g, net.part1.conv_b)])
    bs = torch.stack([s.bias.detach().reshape(()) for s in (net.part1.conv_r, net.part1.conv_g, net.part1.conv_b)])
    out.update({"tail_field": field, "tail_weights": ws, "tail_bias": bs, "tail_poh": poh})

    np.savez_compressed(os.path.join(HERE, "next_small.npz"),
                        **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                           for k, v in out.items()})
    print("wrote next_small.npz:", sorted(out))


if __name__ == "__main__":
    main()
