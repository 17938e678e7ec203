"""CPU restatement (torch, differentiable) of the reference's scene-graph composer — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's baselines may import this.  It restates, citing /root/reference:
  lib/models/street_gaussian_model.py:287-304  get_scaling   (cat of per-model exp)
  lib/models/street_gaussian_model.py:305-333  get_rotation  (normalize; optional flip quaternion; q_obj (x) q_local; normalize)
  lib/models/street_gaussian_model.py:335-363  get_xyz       (optional y-flip; R(q_obj) x + t_obj)
  lib/models/street_gaussian_model.py:365-381  get_features  (background: cat(dc, rest); actors: Fourier DC)
  lib/models/street_gaussian_model.py:431-449  get_opacity   (cat of per-model sigmoid)
  lib/models/gaussian_model.py:224-251         activations   (exp / sigmoid / F.normalize)
  lib/models/gaussian_model_actor.py:71-80     get_features_fourier (sum_c dc[:, c] * IDFT(t)[c])
  lib/utils/general_utils.py:125-146, 220-238  quaternion_to_matrix (normalises its input), quaternion_raw_multiply
Pinned by tests/golden/callsite/compose_sh{1,3}.npz, which the reference's own StreetGaussianModel produced
(tests/golden/make_compose_golden.py); tests/test_compose_cpu.py checks this file against them to 1e-6.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def quat_mul(a, b):  # general_utils.py:220-238
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)


def quat_to_matrix(r):  # general_utils.py:125-146 (normalises first)
    q = r / torch.sqrt((r * r).sum(-1, keepdim=True))
    w, x, y, z = torch.unbind(q, -1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(r.shape[:-1] + (3, 3))


def compose(models, poses, idft, flip, flip_quat):
    """models[0] = background, models[1:] = actors (dicts of raw tensors: xyz, rotation, scaling, opacity, features_dc, features_rest).
    poses[a] = (qw, qx, qy, qz, tx, ty, tz) of actor a; idft[a] = its IDFT row; flip = bool [sum of actor counts] or None.
    Returns dict(xyz, rotation, scaling, opacity, features) exactly like StreetGaussianModel.get_*."""
    bk, actors = models[0], models[1:]
    xyz, rot, feat = [bk["xyz"]], [F.normalize(bk["rotation"])], [torch.cat((bk["features_dc"], bk["features_rest"]), 1)]
    scal = [torch.exp(m["scaling"]) for m in models]
    opac = [torch.sigmoid(m["opacity"]) for m in models]
    off = 0
    for a, m in enumerate(actors):
        n = m["xyz"].shape[0]
        q_obj, t_obj = poses[a, :4], poses[a, 4:7]
        fm = flip[off:off + n].bool() if flip is not None else torch.zeros(n, dtype=torch.bool, device=m["xyz"].device)
        off += n
        x_l = torch.where(fm[:, None], m["xyz"] * torch.tensor([1.0, -1.0, 1.0], dtype=m["xyz"].dtype, device=m["xyz"].device), m["xyz"])  # :347-349 (flip_axis = 1)
        xyz.append(torch.einsum("ij,bj->bi", quat_to_matrix(q_obj[None])[0], x_l) + t_obj)                     # :350-351
        r_l = F.normalize(m["rotation"])
        if flip is not None:
            r_l = torch.where(fm[:, None], quat_mul(flip_quat.to(r_l.dtype).expand(n, 4), r_l), r_l)              # :319-323
        rot.append(F.normalize(quat_mul(q_obj.expand(n, 4), r_l)))                                               # :324-325
        dc = (m["features_dc"] * idft[a][None, :, None].to(m["features_dc"].dtype)).sum(1, keepdim=True)        # gaussian_model_actor.py:76-77
        feat.append(torch.cat((dc, m["features_rest"]), 1))
    return dict(xyz=torch.cat(xyz), rotation=torch.cat(rot), scaling=torch.cat(scal), opacity=torch.cat(opac), features=torch.cat(feat))
