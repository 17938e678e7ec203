"""the reference's OWN render call site runs unchanged against this package.

`StreetGaussianRenderer.render_kernel` (lib/models/street_gaussian_renderer.py:122-280) and `make_rasterizer`
(lib/utils/camera_utils.py:194-227) are imported from /root/reference without edits (tests/refharness.py lists the fakes a
CPU-only box needs), with `diff_gaussian_rasterization` resolved to street_gaussians_b200's shim, and driven with a REAL
`StreetGaussianModel` (its compose code, street_gaussian_model.py:287-449, runs as shipped).  There is no GPU here and the
product has no CPU fallback, so the three C-ABI stages are replaced by recorders: the test pins everything ABOVE the C ABI —
import names, the 12 settings fields, keyword names, None handling, result tuple, gradient slots.  The recorded call is
committed as tests/golden/callsite/render_kernel.npz (tests/golden/make_callsite_golden.py) and replayed on the GPU box
through the compiled reference and through libsgr.so by tests/test_parity_gpu.py::test_callsite_replay_vs_reference.
/root/reference does not exist on the GPU box: this module is skipped there.
"""
import numpy as np
import pytest
import torch

import refharness as H

pytestmark = pytest.mark.skipif(not H.available() or torch.cuda.is_available(),
                                reason="needs /root/reference (build container only); the GPU half is test_callsite_replay_vs_reference")


class Recorder:
    """Stands in for the three C-ABI stages of street_gaussians_b200.rasterizer (forward, backward blend, backward geom)."""

    def __init__(self):
        self.calls = []

    def install(self, monkeypatch):
        from street_gaussians_b200 import rasterizer as R
        rec = self

        def fwd(means3D, sh, colors_precomp, semantics, opacities, scales, rotations, cov3Ds_precomp, settings, band, capacity=None):
            P = means3D.shape[0]
            H_, W_ = int(settings.image_height), int(settings.image_width)
            S = int(semantics.shape[1]) if semantics is not None and semantics.dim() == 2 else 0
            rec.calls.append(dict(means3D=means3D, sh=sh, colors_precomp=colors_precomp, semantics=semantics, opacities=opacities,
                                  scales=scales, rotations=rotations, cov3Ds_precomp=cov3Ds_precomp, settings=settings))
            st = R._ForwardState()
            st.geom = st.img = st.binning = None
            st.num_instances = 0
            none_if_empty = lambda t: None if t is None or t.numel() == 0 else t
            tensors = dict(means3D=means3D, opacities=opacities, sh=none_if_empty(sh), colors_precomp=none_if_empty(colors_precomp),
                           scales=none_if_empty(scales), rotations=none_if_empty(rotations), cov3Ds_precomp=none_if_empty(cov3Ds_precomp),
                           semantics=semantics if S > 0 else None)
            z = lambda c: torch.zeros((c, H_, W_))
            return z(3), torch.arange(P, dtype=torch.int32) % 3, z(1), z(1), z(S), st, tensors

        def bwd_blend(settings, band, st, tensors, alpha, gc, gd, ga, gs, grad2d_out=None):
            P = tensors["means3D"].shape[0]
            S = int(tensors["semantics"].shape[1]) if tensors["semantics"] is not None else 0
            return torch.ones((P, 12)), torch.ones((P, S))

        def bwd_geom(settings, band, st, tensors, radii, grad2d):
            P = tensors["means3D"].shape[0]
            M = tensors["sh"].shape[1] if tensors["sh"] is not None else 0
            f = lambda *s: torch.full(s, 0.5)
            return (f(P, 3), torch.cat([f(P, 2), torch.full((P, 1), 7.0)], 1), f(P, M, 3) if M else None, None, f(P, 1), f(P, 3), f(P, 4), None)

        monkeypatch.setattr(R, "_forward_impl", fwd)
        monkeypatch.setattr(R, "_backward_blend_impl", bwd_blend)
        monkeypatch.setattr(R, "_backward_geom_impl", bwd_geom)


def test_reference_render_kernel_runs_unchanged_on_the_shim(monkeypatch):
    ns = H.load()
    sgb = ns.sgb
    # the reference's import statement (camera_utils.py:13) resolved to this package through the shim
    assert ns.camera_utils.GaussianRasterizer is sgb.GaussianRasterizer
    assert ns.camera_utils.GaussianRasterizationSettings is sgb.GaussianRasterizationSettings
    rec = Recorder()
    rec.install(monkeypatch)
    cam = H.make_camera(ns)
    model = H.make_street_model(ns)
    model.set_visibility(["background"] + model.obj_list)
    model.parse_camera(cam)
    out = ns.renderer.StreetGaussianRenderer().render_kernel(cam, model, white_background=False)
    assert len(rec.calls) == 1
    c = rec.calls[0]
    P = model.num_gaussians
    # what the call site passed: the composed tensors of the real model, in the reference's keyword slots
    assert torch.equal(c["means3D"], model.get_xyz) and c["means3D"].shape == (P, 3)
    assert torch.equal(c["opacities"], model.get_opacity) and torch.equal(c["scales"], model.get_scaling)
    assert torch.equal(c["sh"], model.get_features) and c["sh"].shape == (P, (model.max_sh_degree + 1) ** 2, 3)
    assert c["rotations"].shape == (P, 4)
    assert c["colors_precomp"].numel() == 0 and c["cov3Ds_precomp"].numel() == 0  # None -> empty CPU tensors (DGR __init__.py:207-217)
    assert c["semantics"].shape == (P, 0)  # None -> zeros(P, 0) (DGR __init__.py:219)
    s = c["settings"]
    assert s._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
                         "sh_degree", "campos", "prefiltered", "debug")
    assert (s.image_height, s.image_width) == (cam.image_height, cam.image_width) and s.sh_degree == model.max_sh_degree
    assert torch.equal(s.viewmatrix, cam.world_view_transform) and torch.equal(s.projmatrix, cam.full_proj_transform)
    assert torch.equal(s.campos, cam.camera_center) and s.prefiltered is False and s.debug is False
    assert abs(s.tanfovx - np.tan(cam.FoVx * 0.5)) < 1e-12 and s.scale_modifier == 1.0
    # result dict of the call site (street_gaussian_renderer.py:266-280)
    assert set(out) >= {"rgb", "acc", "depth", "viewspace_points", "visibility_filter", "radii"}
    assert out["rgb"].shape == (3, cam.image_height, cam.image_width) and out["radii"].dtype == torch.int32
    assert torch.equal(out["visibility_filter"], out["radii"] > 0)
    # backward: the 3-column means2D gradient reaches viewspace_points (.grad[:, :2] and [:, 2:] are both read by
    # street_gaussian_model.py:569-570) and the parameter gradients flow back through the reference's compose code
    (out["rgb"].sum() + out["depth"].sum() + out["acc"].sum()).backward()
    vg = out["viewspace_points"].grad
    assert vg.shape == (P, 3) and float(vg[0, 2]) == 7.0 and float(vg[0, 0]) == 0.5
    assert model.background._xyz.grad is not None and model.background._opacity.grad is not None
    obj = getattr(model, model.obj_list[0])
    assert obj._xyz.grad is not None and obj._features_dc.grad.shape == obj._features_dc.shape
    assert model.actor_pose.opt_trans.grad is not None  # tracked-pose refinement receives gradients through get_xyz


def test_eval_mode_means2D_none_and_empty_scene(monkeypatch):
    """means2D may be None in eval (street_gaussian_renderer.py:170-173): the shim must accept it."""
    ns = H.load()
    rec = Recorder()
    rec.install(monkeypatch)
    cam = H.make_camera(ns)
    model = H.make_street_model(ns, n_bkgd=50, n_obj=1, per_obj=20)
    model.set_visibility(["background"] + model.obj_list)
    model.parse_camera(cam)
    rast = ns.camera_utils.make_rasterizer(cam, model.max_sh_degree, torch.zeros(3), 1.0)
    with torch.no_grad():
        color, radii, depth, alpha, sem = rast(means3D=model.get_xyz, means2D=None, opacities=model.get_opacity, shs=model.get_features,
                                               scales=model.get_scaling, rotations=model.get_rotation)
    assert color.shape[0] == 3 and sem.shape[0] == 0 and len(rec.calls) == 1
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        rast(means3D=model.get_xyz, means2D=None, opacities=model.get_opacity, scales=model.get_scaling, rotations=model.get_rotation)
