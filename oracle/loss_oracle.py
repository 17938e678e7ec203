"""CPU/GPU torch restatement of the reference's image losses — TEST INFRASTRUCTURE, not product code (only tests/,
__graft_entry__.smoke() and bench.py's reference arm may import it).  Restates, citing /root/reference:
  lib/utils/loss_utils.py:21-37   l1_loss  (mean over masked pixels x channels)
  lib/utils/loss_utils.py:84-126  gaussian / create_window / ssim / _ssim (11x11, sigma 1.5, zero padding, masked pixels zeroed in both images)
  train.py:101-104                loss = (1 - l) * l1w * L1 + l * (1 - SSIM)
  train.py:107-113                sky loss on the clamped accumulation map
Pinned by tests/golden/callsite/losses.npz, which the reference's OWN l1_loss / ssim produced (tests/golden/make_loss_golden.py).
"""
from __future__ import annotations

from math import exp

import torch
import torch.nn.functional as F


def l1_loss(out, gt, mask=None):
    a, b = out.permute(1, 2, 0), gt.permute(1, 2, 0)
    if mask is not None:
        a, b = a[mask.squeeze(0)], b[mask.squeeze(0)]
    return (a - b).abs().mean()


def _window(channel, dtype, device):
    g = torch.tensor([exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).float()[None, None].expand(channel, 1, 11, 11).contiguous().to(dtype=dtype, device=device)


def ssim(img1, img2, mask=None):
    ch = img1.size(-3)
    w = _window(ch, img1.dtype, img1.device)
    if mask is not None:
        img1 = torch.where(mask, img1, torch.zeros_like(img1))
        img2 = torch.where(mask, img2, torch.zeros_like(img2))
    conv = lambda t: F.conv2d(t, w, padding=5, groups=ch)
    mu1, mu2 = conv(img1), conv(img2)
    s1, s2, s12 = conv(img1 * img1) - mu1 * mu1, conv(img2 * img2) - mu2 * mu2, conv(img1 * img2) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))).mean()


def photometric_loss(image, gt, mask, lambda_l1, lambda_dssim):
    return (1.0 - lambda_dssim) * lambda_l1 * l1_loss(image, gt, mask) + lambda_dssim * (1.0 - ssim(image, gt, mask))


def sky_loss(acc, sky_mask, weight=1.0):
    a = torch.clamp(acc, min=1e-6, max=1.0 - 1e-6)
    return weight * torch.where(sky_mask, -torch.log(1 - a), -torch.log(a)).mean()
