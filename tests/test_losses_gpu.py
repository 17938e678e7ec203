"""GPU: sgr_image_loss / sgr_sky_loss / sgr_densify_stats / sgr_adam_step through their Python mirrors, against the reference's
own outputs (tests/golden/callsite/losses.npz), the torch oracles and torch.optim.Adam."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_loss_golden import case  # noqa: E402
import street_gaussians_b200 as sgb  # noqa: E402
from oracle import loss_oracle as LO  # noqa: E402
from street_gaussians_b200 import losses, training  # noqa: E402
from test_losses_cpu import FIX, rel  # noqa: E402

pytestmark = pytest.mark.gpu


def test_l1_and_ssim_vs_reference_fixture():
    z = np.load(FIX)
    for seed in (0, 1):
        img, gt, mask = case(seed)
        for tag, m in (("nomask", None), ("mask", mask)):
            k = f"s{seed}_{tag}_"
            md = m.cuda() if m is not None else None
            x = img.cuda().requires_grad_(True)
            l1 = losses.l1_loss(x, gt.cuda(), md)
            l1.backward()
            assert abs(float(l1) - float(z[k + "l1"])) < 2e-7 and rel(x.grad.cpu().numpy(), z[k + "g_l1"]) < 1e-6, k
            x = img.cuda().requires_grad_(True)
            ss = losses.ssim(x, gt.cuda(), mask=md)
            ss.backward()
            assert abs(float(ss) - float(z[k + "ssim"])) < 2e-6, (k, float(ss), float(z[k + "ssim"]))
            assert rel(x.grad.cpu().numpy(), z[k + "g_ssim"]) < 2e-4, (k, rel(x.grad.cpu().numpy(), z[k + "g_ssim"]))


@pytest.mark.parametrize("H,W", [(1280, 1920), (37, 16), (16, 5)])
def test_photometric_and_sky_loss_vs_oracle(H, W):
    g = torch.Generator().manual_seed(H + W)
    gt = torch.rand(3, H, W, generator=g).cuda()
    img0 = (gt + 0.1 * torch.randn(3, H, W, generator=g).cuda()).clamp(0, 1)
    mask = (torch.rand(1, H, W, generator=g) > 0.2).cuda()
    for m in (None, mask):
        x = img0.clone().requires_grad_(True)
        a = losses.photometric_loss(x, gt, m, lambda_l1=1.0, lambda_dssim=0.2)
        (2.5 * a).backward()  # a non-unit upstream gradient
        y = img0.clone().requires_grad_(True)
        b = LO.photometric_loss(y, gt, m, 1.0, 0.2)
        (2.5 * b).backward()
        assert abs(float(a) - float(b)) < 2e-6 * max(1.0, abs(float(b)))
        assert rel(x.grad.cpu().numpy(), y.grad.cpu().numpy()) < 3e-4
    acc0 = torch.rand(1, H, W, generator=g).cuda()
    acc0[0, 0, :3] = torch.tensor([0.0, 1.0, 5e-7])[: min(3, W)].cuda()  # values the clamp catches: zero gradient there
    sky = (torch.rand(1, H, W, generator=g) > 0.5).cuda()
    x = acc0.clone().requires_grad_(True)
    a = losses.sky_loss(x, sky, 0.05)
    a.backward()
    y = acc0.clone().requires_grad_(True)
    b = LO.sky_loss(y, sky, 0.05)
    b.backward()
    assert abs(float(a) - float(b)) < 1e-6 * max(1.0, abs(float(b))) and rel(x.grad.cpu().numpy(), y.grad.cpu().numpy()) < 1e-5


def test_densification_stats_match_reference_semantics():
    """street_gaussian_model.py:551-571 replayed with torch ops on the same tensors."""
    g = torch.Generator().manual_seed(4)
    counts = [1001, 37, 0, 500]
    P = sum(counts)
    radii = (torch.randint(-1, 40, (P,), generator=g, dtype=torch.int32)).cuda()
    grad = torch.randn(P, 3, generator=g).cuda()

    def fresh():
        return [dict(max_radii2D=torch.rand(n, generator=torch.Generator().manual_seed(n)).cuda() * 30,
                     xyz_gradient_accum=torch.rand(n, 2, generator=torch.Generator().manual_seed(n + 1)).cuda(),
                     denom=torch.ones(n, 1).cuda()) for n in counts]

    mine, ref = fresh(), fresh()
    training.add_densification_stats(mine, radii, grad)
    vis, start = radii > 0, 0
    for m, n in zip(ref, counts):  # the reference's per-model slicing
        v, r, gm = vis[start:start + n], radii[start:start + n].float(), grad[start:start + n]
        m["max_radii2D"][v] = torch.max(m["max_radii2D"][v], r[v])
        m["xyz_gradient_accum"][v, 0:1] += torch.norm(gm[v, :2], dim=-1, keepdim=True)
        m["xyz_gradient_accum"][v, 1:2] += torch.norm(gm[v, 2:], dim=-1, keepdim=True)
        m["denom"][v] += 1
        start += n
    for a, b in zip(mine, ref):
        for k in a:
            assert torch.allclose(a[k], b[k], rtol=1e-6, atol=1e-7), k


def test_fused_adam_matches_torch_adam():
    g = torch.Generator().manual_seed(6)
    shapes = [(1001, 3), (1001, 1, 3), (1001, 15, 3), (1001, 1), (1001, 3), (1001, 4), (37, 3), (70001,)]
    lrs = [1.6e-4, 2.5e-3, 1.25e-4, 5e-2, 5e-3, 1e-3, 1e-2, 3e-3]
    base = [torch.randn(*s, generator=g) for s in shapes]
    pa = [torch.nn.Parameter(b.clone().cuda()) for b in base]
    pb = [torch.nn.Parameter(b.clone().cuda()) for b in base]
    oa = training.FusedAdam([dict(params=[p], lr=lr, name=str(i)) for i, (p, lr) in enumerate(zip(pa, lrs))], lr=0.0, eps=1e-15)
    ob = torch.optim.Adam([dict(params=[p], lr=lr, name=str(i)) for i, (p, lr) in enumerate(zip(pb, lrs))], lr=0.0, eps=1e-15)
    for it in range(5):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).cuda() * (10.0 ** (it - 2))
            a.grad, b.grad = gr.clone(), gr.clone()
        if it == 3:
            pa[1].grad = None  # a parameter without a gradient is skipped, like torch
            pb[1].grad = None
            oa.param_groups[0]["lr"] = ob.param_groups[0]["lr"] = 7e-5  # update_learning_rate (gaussian_model.py:320-325)
        oa.step()
        ob.step()
        oa.zero_grad(set_to_none=True)
        ob.zero_grad(set_to_none=True)
    for i, (a, b) in enumerate(zip(pa, pb)):
        assert rel(a.detach().cpu().numpy(), b.detach().cpu().numpy()) < 2e-6, i
        assert rel(oa.state[a]["exp_avg_sq"].cpu().numpy(), ob.state[b]["exp_avg_sq"].cpu().numpy()) < 2e-6
        assert int(oa.state[a]["step"]) == int(ob.state[b]["step"])
