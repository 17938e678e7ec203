"""Seeded raw parameters of a small street scene (background + actors) for the composer tests — pure torch, no reference
import, so the GPU box regenerates the SAME tensors that tests/golden/make_compose_golden.py fed to the reference's
StreetGaussianModel on the build container."""
from __future__ import annotations

import torch

KEYS = ("xyz", "rotation", "scaling", "opacity", "features_dc", "features_rest")


def raw_model(g: torch.Generator, n: int, M: int, C: int, centre, spread):
    rn = lambda *s, sd=1.0: torch.randn(*s, generator=g) * sd
    return dict(xyz=rn(n, 3) * torch.tensor(spread) + torch.tensor(centre), features_dc=rn(n, C, 3), features_rest=rn(n, M - 1, 3, sd=0.2),
                scaling=torch.log(torch.tensor(0.08)) + rn(n, 3, sd=0.5), rotation=rn(n, 4), opacity=rn(n, 1, sd=2.0))


def make_case(seed: int, n_bkgd: int, actors, M: int, C: int):
    """[background, actor_0, actor_1, ...] as dicts of raw tensors (names = the reference's nn.Parameters without the underscore)."""
    g = torch.Generator().manual_seed(seed)
    models = [raw_model(g, n_bkgd, M, 1, [0.0, 0.0, 20.0], [8.0, 3.0, 10.0])]
    for n in actors:
        models.append(raw_model(g, n, M, C, [0.0, 0.0, 0.0], [1.5, 0.6, 0.5]))
    return models


def upstream(seed: int, P: int, M: int):
    """Seeded upstream gradients on the five composed tensors."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    return dict(xyz=rn(P, 3), rotation=rn(P, 4), scaling=rn(P, 3), opacity=rn(P, 1), features=rn(P, M, 3))
