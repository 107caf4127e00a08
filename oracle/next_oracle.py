"""CPU restatement of the stages either side of the propagation path.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing else); the product path never touches it.
Each function follows the reference lines it cites ("loss.py" = watermelon_hologram/loss_func.py, "util.py" =
utilities.py, "ap2poh.py" = watermelon_hologram/AP2POH.py, "nn.py" = neural_network_components.py, "dl.py" =
watermelon_hologram/data_loader.py).  Pinned by tests/test_next_oracle_pinned.py: bit-identical to the live
reference functions in this container and to tests/golden/next_small.npz (made by tests/golden/make_golden_next.py
from the unmodified reference).
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ---- N1 -----------------------------------------------------------------------------------------------------
def total_variation(t):  # loss.py:66-77
    d1 = t[:, :, :, 1:] - t[:, :, :, :-1]
    d2 = t[:, :, 1:, :] - t[:, :, :-1, :]
    return torch.mean(torch.abs(d1)) + torch.mean(torch.abs(d2))


def total_variation_loss(y_hat, y):  # loss.py:92-96
    return torch.abs(total_variation(y_hat) - total_variation(y))


def amp_loss(amp_hat, amp, alpha=1.0):  # loss.py:99-103
    return F.mse_loss(amp_hat, amp) + alpha * total_variation_loss(amp_hat, amp)


def focal_sincos_phase_gradient_loss(fake_phase, real_phase):  # loss.py:135-163
    sf = torch.cat((torch.sin(fake_phase), torch.cos(fake_phase)), dim=1)
    sr = torch.cat((torch.sin(real_phase), torch.cos(real_phase)), dim=1)
    d1 = torch.abs((sf[:, :, :, 1:] - sf[:, :, :, :-1]) - (sr[:, :, :, 1:] - sr[:, :, :, :-1]))
    d2 = torch.abs((sf[:, :, 1:, :] - sf[:, :, :-1, :]) - (sr[:, :, 1:, :] - sr[:, :, :-1, :]))
    with torch.no_grad():
        w1 = torch.pow(d1, 1)
        w1 = w1 / torch.max(w1)
        w2 = torch.pow(d2, 1)
        w2 = w2 / torch.max(w2)
    return torch.mean(d1 * w1) + torch.mean(d2 * w2)


def phase_sincos_gradient_loss(fake_phase, real_phase):  # loss.py:165-183
    sf = torch.cat((torch.sin(fake_phase), torch.cos(fake_phase)), dim=1)
    sr = torch.cat((torch.sin(real_phase), torch.cos(real_phase)), dim=1)
    d1 = torch.abs((sf[:, :, :, 1:] - sf[:, :, :, :-1]) - (sr[:, :, :, 1:] - sr[:, :, :, :-1]))
    d2 = torch.abs((sf[:, :, 1:, :] - sf[:, :, :-1, :]) - (sr[:, :, 1:, :] - sr[:, :, :-1, :]))
    return torch.mean(d1) + torch.mean(d2)


def focal_sincos_phase_loss(fake_phase, real_phase):  # loss.py:186-203
    sf = torch.cat((torch.sin(fake_phase), torch.cos(fake_phase)), dim=1)
    sr = torch.cat((torch.sin(real_phase), torch.cos(real_phase)), dim=1)
    d1 = torch.abs(sf - sr)
    with torch.no_grad():
        w = torch.pow(d1, 1)
        w = w / torch.max(w)
    return torch.mean(d1 * w)


def plain_phase_loss(fake_phase, real_phase):  # loss.py:206-208
    return torch.mean(torch.abs(fake_phase - real_phase))


# ---- N4 -----------------------------------------------------------------------------------------------------
def tensor_normalizor_2D(t):  # util.py:69-84
    mx, _ = torch.max(t, dim=-1, keepdim=True)
    mx, _ = torch.max(mx, dim=-2, keepdim=True)
    mn, _ = torch.min(t, dim=-1, keepdim=True)
    mn, _ = torch.min(mn, dim=-2, keepdim=True)
    return (t - mn) / (mx - mn)


def imsave_bytes(rgb):
    """What ``plt.imsave(name, rgb)`` stores for a float32 [R,C,3] array in [0,1] (util.py:147-151): matplotlib's
    ``ScalarMappable.to_rgba(bytes=True)`` float branch, ``(xx * 255).astype(np.uint8)`` on a float32 RGBA array
    whose alpha is 1."""
    rgb = np.asarray(rgb, dtype=np.float32)
    xx = np.empty(rgb.shape[:2] + (4,), dtype=np.float32)
    xx[:, :, :3] = rgb
    xx[:, :, 3] = 1
    return (xx * 255).astype(np.uint8)


def focal_stack_u8(amp):  # generatePOH.py:72-78 -> util.py:179-203 -> util.py:147-151
    norm = tensor_normalizor_2D(amp)
    return np.stack([imsave_bytes(norm[i].permute(1, 2, 0).numpy()) for i in range(norm.shape[0])])


# ---- N2 -----------------------------------------------------------------------------------------------------
def amplitude_normalizor(amp):  # util.py:53-66
    mx, _ = torch.max(amp, dim=-1, keepdim=True)
    mx, _ = torch.max(mx, dim=-2, keepdim=True)
    return amp / (mx * 1.01)


def channelwise_symmetric_conv(x, weights, bias):  # nn.py:59-95: one k x k kernel + bias per colour, zero padding
    k = weights.shape[-1]
    outs = [F.conv2d(x[:, c:c + 1], weights[c][None, None], bias[c].reshape(1), padding=(k - 1) // 2)
            for c in range(3)]
    return torch.cat(outs, dim=1)


def checkerboard(height, width, reserve):  # util.py:354-382 with cell_size 1
    x = np.arange(width).reshape(1, -1)
    y = np.arange(height).reshape(-1, 1)
    cb = torch.tensor(((x + y) % 2).astype(np.float32))
    return 1 - cb if reserve else cb


def ap2poh_tail(field, weights, bias):  # ap2poh.py:86-95,107-116
    m = torch.complex(channelwise_symmetric_conv(torch.real(field), weights, bias),
                      channelwise_symmetric_conv(torch.imag(field), weights, bias))
    amp, phs = amplitude_normalizor(torch.abs(m)), torch.angle(m)
    acos_amp = torch.acos(amp)
    m1 = checkerboard(field.shape[-2], field.shape[-1], True)
    m2 = checkerboard(field.shape[-2], field.shape[-1], False)
    return m1 * (phs + acos_amp) + m2 * (phs - acos_amp)


# ---- N3 -----------------------------------------------------------------------------------------------------
def rgbd_item(img, depth, idx):  # dl.py:43-51
    return torch.cat((torch.tensor(img[idx]), torch.tensor(depth[idx][0]).unsqueeze(0)), dim=0)


def pi_phase_item(phs, idx):  # dl.py:85
    return 2 * torch.pi * torch.tensor(phs[idx])
