"""Post-backward bookkeeping of a training iteration over sgr_densify_stats / sgr_adam_step.

    add_densification_stats(models, radii, viewspace_point_grad)
        = StreetGaussianModel.set_max_radii2D + add_densification_stats (lib/models/street_gaussian_model.py:551-571) for all
        sub-models in one kernel.  `models`: objects / mappings exposing max_radii2D [n], xyz_gradient_accum [n,2], denom [n,1]
        (the attributes lib/models/gaussian_model.py:49-51 creates) in composition order (background, then the frame's actors).
    FusedAdam(param_groups, ...)
        torch.optim.Adam's interface (param_groups with per-group "lr" / "name", .step(), .zero_grad(), state_dict) for the way the
        reference uses it (gaussian_model.py:300-303: lr per group, eps=1e-15, no weight decay / amsgrad); ONE kernel updates every
        tensor of every group — pass the groups of all sub-models to a single FusedAdam to get a single launch per iteration.
CUDA tensors only; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import torch

from . import _capi
from .rasterizer import _ptr, _stream


def _attr(model, name):
    return model[name] if isinstance(model, dict) else getattr(model, name)


def add_densification_stats(models: Sequence, radii: torch.Tensor, viewspace_point_grad: torch.Tensor) -> None:
    L = _capi.lib()
    if not radii.is_cuda:
        raise _capi.SgrError("add_densification_stats needs CUDA tensors (there is no CPU fallback)")
    dev = radii.device
    n = len(models)
    segs = (_capi.SgrStatSegment * n)()
    start, keep = 0, []
    for k, m in enumerate(models):
        mr, ga, dn = _attr(m, "max_radii2D"), _attr(m, "xyz_gradient_accum"), _attr(m, "denom")
        for t, shape in ((mr, 1), (ga, 2), (dn, 1)):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                raise _capi.SgrError("densification statistics must be contiguous fp32 CUDA tensors (they are updated in place)")
        cnt = int(mr.shape[0])
        if ga.numel() != 2 * cnt or dn.numel() != cnt:
            raise ValueError(f"model {k}: xyz_gradient_accum must be [{cnt}, 2] and denom [{cnt}, 1]")
        s = segs[k]
        s.start, s.count = start, cnt
        s.max_radii2D, s.xyz_gradient_accum, s.denom = (t.data_ptr() if cnt else None for t in (mr, ga, dn))
        start += cnt
    if radii.numel() != start or viewspace_point_grad.shape != (start, 3):
        raise ValueError(f"radii must be [{start}] and the viewspace gradient [{start}, 3] for these models")
    r = radii if (radii.dtype == torch.int32 and radii.is_contiguous()) else radii.to(torch.int32).contiguous()
    g = viewspace_point_grad if (viewspace_point_grad.dtype == torch.float32 and viewspace_point_grad.is_contiguous()) \
        else viewspace_point_grad.to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        rc = L.sgr_densify_stats(segs, n, _ptr(r), _ptr(g), _stream(dev))
    _capi.check(rc, "sgr_densify_stats")


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas=(0.9, 0.999), eps) semantics, every parameter of every group updated by one kernel."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        L = _capi.lib()
        buckets = {}  # (device, betas, eps) -> [(param, group)]
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise _capi.SgrError("FusedAdam needs CUDA parameters (there is no CPU fallback)")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise _capi.SgrError("FusedAdam updates contiguous fp32 parameters in place")
                buckets.setdefault((p.device, tuple(group["betas"]), float(group["eps"])), []).append((p, group))
        for (dev, betas, eps), items in buckets.items():
            tab = (_capi.SgrAdamTensor * len(items))()
            keep = []
            for k, (p, group) in enumerate(items):
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = int(st["step"]) + 1
                g = p.grad if (p.grad.dtype == torch.float32 and p.grad.is_contiguous()) else p.grad.to(torch.float32).contiguous()
                keep.append(g)
                t = tab[k]
                t.param, t.grad, t.exp_avg, t.exp_avg_sq = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                t.numel, t.lr, t.step = p.numel(), float(group["lr"]), st["step"]
            with torch.cuda.device(dev):
                rc = L.sgr_adam_step(tab, len(items), float(betas[0]), float(betas[1]), eps, _stream(dev))
            _capi.check(rc, "sgr_adam_step")
            del keep
        return loss
