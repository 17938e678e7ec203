"""Generates tests/golden/callsite/compose_sh{1,3}.npz: what the reference's UNMODIFIED StreetGaussianModel
(lib/models/street_gaussian_model.py:225-449: parse_camera, get_xyz / get_rotation / get_scaling / get_opacity / get_features;
lib/models/gaussian_model.py:224-251 activations; lib/models/gaussian_model_actor.py:71-80 Fourier DC) composes from the seeded
raw parameters of tests/compose_case.py, and the gradients torch autograd sends back through that code for seeded upstream
gradients — including the tracked-pose gradients (obj_rots / obj_trans).  Run on the build container (CPU, tests/refharness.py):
    python tests/golden/make_compose_golden.py 1 ; python tests/golden/make_compose_golden.py 3
The fixture stores the per-frame inputs the reference derived (actor poses, IDFT rows, the random flip mask, the flip
quaternion) and the reference's outputs; the raw parameters are regenerated from the seed on the GPU box.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import compose_case as CC  # noqa: E402
import refharness as H  # noqa: E402

SEED, N_BKGD, ACTORS = 5, 1500, (700, 500)


def main():
    deg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    ns = H.load(extra_opts=["model.gaussian.sh_degree", str(deg)])
    torch.manual_seed(0)  # the flip mask is torch.rand_like inside parse_camera (street_gaussian_model.py:275-284)
    cam = H.make_camera(ns, frame=3)
    model = H.make_street_model(ns, n_bkgd=8, n_obj=len(ACTORS), per_obj=8)
    M, C = (model.max_sh_degree + 1) ** 2, int(model.fourier_dim)
    assert model.max_sh_degree == deg
    raw = CC.make_case(SEED, N_BKGD, ACTORS, M, C)
    subs = [model.background] + [getattr(model, n) for n in model.obj_list]
    for sub, r in zip(subs, raw):
        for k in CC.KEYS:
            setattr(sub, "_" + k, torch.nn.Parameter(r[k].clone()))
    model.set_visibility(["background"] + model.obj_list)
    model.parse_camera(cam)
    assert model.graph_obj_list == model.obj_list
    model.obj_rots.retain_grad()
    model.obj_trans.retain_grad()
    out = dict(xyz=model.get_xyz, rotation=model.get_rotation, scaling=model.get_scaling, opacity=model.get_opacity, features=model.get_features)
    P = out["xyz"].shape[0]
    up = CC.upstream(SEED + 1, P, M)
    torch.autograd.backward([out[k] for k in up], [up[k] for k in up])
    starts = np.cumsum([0] + [r["xyz"].shape[0] for r in raw])
    # per-actor pose: the reference expands one pose to every Gaussian of the actor (street_gaussian_model.py:269-273)
    poses, dposes, idft = [], [], []
    for a, name in enumerate(model.obj_list):
        lo, hi = starts[a + 1] - N_BKGD, starts[a + 2] - N_BKGD
        rot, tr = model.obj_rots[lo], model.obj_trans[lo]
        assert torch.equal(model.obj_rots[lo:hi], rot.expand(hi - lo, 4)) and torch.equal(model.obj_trans[lo:hi], tr.expand(hi - lo, 3))
        poses.append(torch.cat([rot, tr]).detach().numpy())
        dposes.append(torch.cat([model.obj_rots.grad[lo:hi].sum(0), model.obj_trans.grad[lo:hi].sum(0)]).numpy())
        obj = getattr(model, name)
        t = obj.fourier_scale * (model.frame - obj.start_frame) / (obj.end_frame - obj.start_frame)
        idft.append(ns.sh_utils.IDFT(t, obj.fourier_dim)[0].numpy())
    fx = dict(sh_degree=deg, fourier_dim=C, seed=SEED, n_bkgd=N_BKGD, actors=np.array(ACTORS), poses=np.stack(poses).astype(np.float32),
              idft=np.stack(idft).astype(np.float32), flip=model.flip_mask.numpy().astype(np.uint8),
              flip_quat=model.flip_matrix.reshape(4).numpy().astype(np.float32), ref_dposes=np.stack(dposes).astype(np.float32))
    for k, v in out.items():
        fx["ref_" + k] = v.detach().numpy().astype(np.float32)
    for i, sub in enumerate(subs):
        for k in CC.KEYS:
            fx[f"ref_g{i}_{k}"] = getattr(sub, "_" + k).grad.numpy().astype(np.float32)
    os.makedirs(os.path.join(HERE, "callsite"), exist_ok=True)
    path = os.path.join(HERE, "callsite", f"compose_sh{deg}.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path), "bytes; P =", P, "M =", M, "C =", C, "flipped:", int(model.flip_mask.sum()))


if __name__ == "__main__":
    main()
