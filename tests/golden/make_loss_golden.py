"""Generates tests/golden/callsite/losses.npz with the reference's OWN l1_loss / ssim (lib/utils/loss_utils.py:21-37, 91-126,
imported unmodified through tests/refharness.py on the CPU build container): values and autograd gradients for seeded
[3, 70, 93] image pairs with and without a mask.  python tests/golden/make_loss_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refharness as H  # noqa: E402


def case(seed, Hh=70, Ww=93):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(3, Hh, Ww, generator=g)
    img = (gt + 0.15 * torch.randn(3, Hh, Ww, generator=g)).clamp(0, 1.2)
    mask = torch.rand(1, Hh, Ww, generator=g) > 0.3
    return img, gt, mask


def main():
    ns = H.load()
    lu = ns.loss_utils
    out = {}
    for seed in (0, 1):
        img, gt, mask = case(seed)
        for tag, m in (("nomask", None), ("mask", mask)):
            x = img.clone().requires_grad_(True)
            l1 = lu.l1_loss(x, gt, m)
            (g_l1,) = torch.autograd.grad(l1, x)
            x = img.clone().requires_grad_(True)
            ss = lu.ssim(x, gt, mask=m)
            (g_ss,) = torch.autograd.grad(ss, x)
            k = f"s{seed}_{tag}_"
            out[k + "l1"], out[k + "ssim"] = float(l1), float(ss)
            out[k + "g_l1"], out[k + "g_ssim"] = g_l1.numpy(), g_ss.numpy()
    path = os.path.join(HERE, "callsite", "losses.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), {k: v for k, v in out.items() if not hasattr(v, "shape")})


if __name__ == "__main__":
    main()
