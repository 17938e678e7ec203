"""CPU: the loss oracle (oracle/loss_oracle.py) against values and gradients of the reference's own l1_loss / ssim
(tests/golden/callsite/losses.npz, tests/golden/make_loss_golden.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_loss_golden import case  # noqa: E402  (pure-torch input generator; the reference import happens only in its main())
from oracle import loss_oracle as LO  # noqa: E402

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "callsite", "losses.npz")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def test_loss_oracle_matches_reference_functions():
    z = np.load(FIX)
    for seed in (0, 1):
        img, gt, mask = case(seed)
        for tag, m in (("nomask", None), ("mask", mask)):
            k = f"s{seed}_{tag}_"
            x = img.clone().requires_grad_(True)
            l1 = LO.l1_loss(x, gt, m)
            (g,) = torch.autograd.grad(l1, x)
            assert abs(float(l1) - float(z[k + "l1"])) < 1e-7 and rel(g.numpy(), z[k + "g_l1"]) < 1e-6
            x = img.clone().requires_grad_(True)
            ss = LO.ssim(x, gt, m)
            (g,) = torch.autograd.grad(ss, x)
            assert abs(float(ss) - float(z[k + "ssim"])) < 1e-6 and rel(g.numpy(), z[k + "g_ssim"]) < 1e-5
