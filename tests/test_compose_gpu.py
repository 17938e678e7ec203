"""GPU: sgr_compose_forward / sgr_compose_backward (through street_gaussians_b200.compose) against what the reference's own
StreetGaussianModel composed on the build container (tests/golden/callsite/compose_sh*.npz) and against the torch oracle."""
import numpy as np
import pytest
import torch

import compose_case as CC
import street_gaussians_b200 as sgb
import util
from oracle import compose_oracle as CO
from street_gaussians_b200 import synthetic
from test_compose_cpu import FIX, load_case, rel

pytestmark = pytest.mark.gpu
FWD_TOL, GRAD_TOL = 2e-6, 3e-5  # fp32 elementwise math (expf / sqrtf / rounding order); the rasterizer bars are 1e-4 / 1e-3


def to_dev(models, dev="cuda"):
    return [{k: v.detach().to(dev).requires_grad_(True) for k, v in m.items()} for m in models]


@pytest.mark.parametrize("path", FIX, ids=[p.split("/")[-1] for p in FIX])
def test_compose_vs_reference_model_fixture(path):
    z, M, models, poses, idft, flip, fq = load_case(path)
    dm = to_dev(models)
    dposes = poses.detach().cuda().requires_grad_(True)
    out = sgb.compose(dm, dposes, idft.cuda(), flip.cuda().bool(), fq.cuda())
    names = ("xyz", "rotation", "scaling", "opacity", "features")
    for k, t in zip(names, out):
        assert t.shape == z["ref_" + k].shape and t.is_cuda
        assert rel(t.detach().cpu().numpy(), z["ref_" + k]) < FWD_TOL, k
    up = CC.upstream(int(z["seed"]) + 1, out[0].shape[0], M)
    torch.autograd.backward(list(out), [up[k].cuda() for k in names])
    for i, m in enumerate(dm):
        for k in CC.KEYS:
            assert rel(m[k].grad.cpu().numpy(), z[f"ref_g{i}_{k}"]) < GRAD_TOL, (i, k)
    assert rel(dposes.grad.cpu().numpy(), z["ref_dposes"]) < GRAD_TOL


def test_compose_edge_cases_vs_oracle():
    """background only; an EMPTY actor between two real ones; no flip mask; segment boundaries inside warps."""
    g = torch.Generator().manual_seed(3)
    for actors, use_flip in (((), False), ((37, 0, 501), False), ((64, 33), True)):
        models = CC.make_case(11, 1001, list(actors), 16, 3)
        n_act = len(actors)
        poses = torch.randn(n_act, 7, generator=g) if n_act else None
        idft = torch.randn(n_act, 3, generator=g) if n_act else None
        flip = (torch.rand(sum(actors), generator=g) < 0.5) if use_flip else None
        fq = torch.tensor([0.0, 0.0, 1.0, 0.0])
        cm = [{k: v.clone().requires_grad_(True) for k, v in m.items()} for m in models]
        cp = poses.clone().requires_grad_(True) if n_act else None
        ref = CO.compose(cm, cp, idft, flip, fq) if n_act else CO.compose(cm, torch.zeros(0, 7), torch.zeros(0, 3), None, fq)
        dm = to_dev(models)
        dp = poses.cuda().requires_grad_(True) if n_act else None
        out = sgb.compose(dm, dp, idft.cuda() if n_act else None, flip.cuda() if use_flip else None, fq.cuda() if use_flip else None)
        names = ("xyz", "rotation", "scaling", "opacity", "features")
        for k, t in zip(names, out):
            assert rel(t.detach().cpu().numpy(), ref[k].detach().numpy()) < FWD_TOL, (actors, k)
        up = CC.upstream(5, out[0].shape[0], 16)
        torch.autograd.backward(list(out), [up[k].cuda() for k in names])
        torch.autograd.backward([ref[k] for k in names], [up[k] for k in names])
        for i in range(len(models)):
            for k in CC.KEYS:
                if cm[i][k].numel():
                    assert rel(dm[i][k].grad.cpu().numpy(), cm[i][k].grad.numpy()) < GRAD_TOL, (actors, i, k)
        if n_act:
            live = [a for a, n in enumerate(actors) if n > 0]
            assert rel(dp.grad.cpu().numpy()[live], cp.grad.numpy()[live]) < GRAD_TOL
            dead = [a for a, n in enumerate(actors) if n == 0]
            assert float(dp.grad.cpu()[dead].abs().sum()) == 0.0 if dead else True


def test_compose_then_rasterize_matches_materialised_inputs():
    """The composed tensors go straight into GaussianRasterizer: rendering them equals rendering the oracle's composed tensors
    (images to 1e-4, every raw-parameter gradient to 1e-3 — BASELINE.json's bars), i.e. the composer is a drop-in for the
    reference's get_* properties in front of the render call."""
    z, M, models, poses, idft, flip, fq = load_case(FIX[-1])
    dev = "cuda"
    scene = synthetic.make_scene(P=8, width=320, height=208, sh_degree=int(z["sh_degree"]), seed=2)
    cam = scene["cam"]
    st = util.settings_from(sgb, cam, dev)
    rast = sgb.GaussianRasterizer(st)
    g = torch.Generator().manual_seed(9)
    ups = [torch.randn(c, 208, 320, generator=g).to(dev) / (208 * 320) for c in (3, 1, 1)]

    def render(xyz, rot, scale, opac, sh):
        m2d = torch.zeros_like(xyz, requires_grad=True)
        col, radii, dep, alp, _ = rast(means3D=xyz, means2D=m2d, opacities=opac, shs=sh, scales=scale, rotations=rot)
        torch.autograd.backward([col, dep, alp], ups)
        return col.detach(), int((radii > 0).sum())

    dm = to_dev(models)
    dp = poses.detach().cuda().requires_grad_(True)
    col_a, vis = render(*sgb.compose(dm, dp, idft.cuda(), flip.cuda(), fq.cuda()))
    assert vis > 500
    om = to_dev(models)
    op = poses.detach().cuda().requires_grad_(True)
    o = CO.compose(om, op, idft.cuda(), flip.cuda(), fq.cuda())
    col_b, _ = render(o["xyz"], o["rotation"], o["scaling"], o["opacity"], o["features"])
    assert float((col_a - col_b).abs().max()) <= 1e-4
    for i in range(len(models)):
        for k in CC.KEYS:
            assert rel(dm[i][k].grad.cpu().numpy(), om[i][k].grad.cpu().numpy()) < 1e-3, (i, k)
    assert rel(dp.grad.cpu().numpy(), op.grad.cpu().numpy()) < 1e-3
