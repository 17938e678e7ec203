"""Generates tests/golden/callsite/render_kernel.npz: the exact rasterizer call the reference's UNMODIFIED
StreetGaussianRenderer.render_kernel (lib/models/street_gaussian_renderer.py:122-280) makes for a real StreetGaussianModel
(background + 2 actors, config configs/example/waymo_train_002.yaml: SH degree 1, fourier_dim 5), recorded on the CPU build
container through tests/refharness.py.  Run from the repo root:  python tests/golden/make_callsite_golden.py
The GPU box (no /root/reference) replays the file through oracle/_ref and libsgr.so:
tests/test_parity_gpu.py::test_callsite_replay_vs_reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refharness as H  # noqa: E402


def main():
    ns = H.load()
    from street_gaussians_b200 import rasterizer as R
    calls = []

    def fwd(means3D, sh, colors_precomp, semantics, opacities, scales, rotations, cov3Ds_precomp, settings, band, capacity=None):
        calls.append(dict(means3D=means3D, shs=sh, opacities=opacities, scales=scales, rotations=rotations, settings=settings))
        raise StopIteration  # the call has been captured; nothing below the C ABI exists on this box

    R._forward_impl = fwd
    torch.manual_seed(0)
    cam = H.make_camera(ns, width=400, height=240)
    model = H.make_street_model(ns, n_bkgd=12000, n_obj=2, per_obj=2000)
    model.set_visibility(["background"] + model.obj_list)
    model.parse_camera(cam)
    try:
        ns.renderer.StreetGaussianRenderer().render_kernel(cam, model, white_background=False)
    except StopIteration:
        pass
    c = calls[0]
    s = c["settings"]
    H_, W_ = int(s.image_height), int(s.image_width)
    # same key convention as make_golden.py ("in_*"), so tests/test_oracle_cpu.py::scene_from_npz reads it; the upstream
    # gradients are regenerated from a seed by the test instead of being stored (incompressible noise)
    out = {"in_" + k: c[k].detach().numpy().astype(np.float32) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    out.update(image_height=H_, image_width=W_, tanfovx=float(s.tanfovx), tanfovy=float(s.tanfovy), bg=s.bg.numpy().astype(np.float32),
               scale_modifier=float(s.scale_modifier), viewmatrix=s.viewmatrix.numpy().astype(np.float32),
               projmatrix=s.projmatrix.numpy().astype(np.float32), sh_degree=int(s.sh_degree), campos=s.campos.numpy().astype(np.float32))
    os.makedirs(os.path.join(HERE, "callsite"), exist_ok=True)
    path = os.path.join(HERE, "callsite", "render_kernel.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
