"""Host side of the scene-graph composer: the reference's StreetGaussianModel.get_xyz / get_rotation /
get_scaling / get_opacity / get_features (lib/models/street_gaussian_model.py:287-449) as ONE autograd.Function over
sgr_compose_forward / sgr_compose_backward (include/sgr.h).

    models : [background, actor_0, actor_1, ...] — each a mapping (or object) with the reference's raw parameters
             _xyz [n,3], _rotation [n,4], _scaling [n,3], _opacity [n,1], _features_dc [n,C,3], _features_rest [n,M-1,3]
             (lib/models/gaussian_model.py:41-47); the leading underscore is optional in mappings.
    poses  : [num_actors, 7] = (qw, qx, qy, qz, tx, ty, tz): the rows parse_camera expands into obj_rots / obj_trans (:258-273)
    idft   : [num_actors, C] IDFT(t, C) rows (lib/utils/sh_utils.py:120-130)
    flip   : optional bool [sum of actor counts] (the flip_mask of :275-284) with flip_quat [4] (the reference's flip_matrix)

`compose(...)` returns (means3D, rotations, scales, opacities, shs) ready for GaussianRasterizer.forward; gradients flow to every
raw parameter and to `poses`.  PyTorch is plumbing (allocation, autograd bookkeeping); all arithmetic is in libsgr.so, and there
is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _capi
from .rasterizer import _ptr, _stream

RAW_KEYS = ("xyz", "rotation", "scaling", "opacity", "features_dc", "features_rest")


def _raw(model, key):
    if isinstance(model, dict):
        return model[key] if key in model else model["_" + key]
    return getattr(model, "_" + key)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()


class _Compose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, poses, idft, flip, flip_quat, n_models, *raw):
        L = _capi.lib()
        if not raw[0].is_cuda:
            raise _capi.SgrError("street_gaussians_b200 composer needs CUDA tensors (there is no CPU fallback)")
        dev = raw[0].device
        raw = tuple(_f32c(t) for t in raw)
        models = [raw[6 * k: 6 * k + 6] for k in range(n_models)]
        M = int(models[0][5].shape[1]) + 1
        nseg = n_models
        segs = (_capi.SgrSegment * nseg)()
        start = 0
        for k, m in enumerate(models):
            n, Cdim = int(m[0].shape[0]), int(m[4].shape[1])
            if Cdim > _capi.MAX_FOURIER:
                raise _capi.SgrError(f"fourier_dim {Cdim} exceeds SGR_MAX_FOURIER = {_capi.MAX_FOURIER}")
            s = segs[k]
            s.start, s.count, s.fourier_dim, s.posed = start, n, Cdim, int(k > 0)
            s.xyz, s.rotation, s.scaling, s.opacity, s.features_dc, s.features_rest = (t.data_ptr() if t.numel() else None for t in m)
            start += n
        P = start
        # per-segment pose / IDFT rows (row 0 = background, ignored by the kernels)
        pose_tab = torch.zeros((nseg, 8), device=dev, dtype=torch.float32)
        idft_tab = torch.zeros((nseg, _capi.MAX_FOURIER), device=dev, dtype=torch.float32)
        if nseg > 1:
            pose_tab[1:, :7] = poses.to(device=dev, dtype=torch.float32)
            idft_tab[1:, : idft.shape[1]] = idft.to(device=dev, dtype=torch.float32)
        flip_full = None
        if flip is not None and nseg > 1:
            flip_full = torch.zeros((P,), device=dev, dtype=torch.uint8)
            flip_full[int(models[0][0].shape[0]):] = flip.to(device=dev, dtype=torch.uint8)
            flip_quat = _f32c(flip_quat.to(dev).reshape(4))
        e = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        xyz, rot, scale, opac, sh = e(P, 3), e(P, 4), e(P, 3), e(P, 1), e(P, M, 3)
        with torch.cuda.device(dev):
            rc = L.sgr_compose_forward(segs, nseg, M, _ptr(pose_tab), _ptr(idft_tab), _ptr(flip_full), _ptr(flip_quat) if flip_full is not None else None,
                                       _ptr(xyz), _ptr(rot), _ptr(scale), _ptr(opac), _ptr(sh), _stream(dev))
        _capi.check(rc, "sgr_compose_forward")
        ctx.segs, ctx.nseg, ctx.M, ctx.n_models = segs, nseg, M, n_models
        ctx.flip_full, ctx.flip_quat = flip_full, (flip_quat if flip_full is not None else None)
        ctx.pose_shape = None if poses is None else (tuple(poses.shape), poses.dtype, poses.device)
        ctx.save_for_backward(pose_tab, idft_tab, *raw)
        return xyz, rot, scale, opac, sh

    @staticmethod
    def backward(ctx, g_xyz, g_rot, g_scale, g_opac, g_sh):
        L = _capi.lib()
        pose_tab, idft_tab, *raw = ctx.saved_tensors
        dev = pose_tab.device
        nseg, M = ctx.nseg, ctx.M
        P = sum(int(raw[6 * k].shape[0]) for k in range(nseg))
        z = lambda g, *shape: _f32c(g) if g is not None else torch.zeros(shape, device=dev, dtype=torch.float32)
        g_xyz, g_rot, g_scale, g_opac, g_sh = z(g_xyz, P, 3), z(g_rot, P, 4), z(g_scale, P, 3), z(g_opac, P, 1), z(g_sh, P, M, 3)
        grads = [torch.empty_like(t) for t in raw]
        gtab = (_capi.SgrSegmentGrads * nseg)()
        for k in range(nseg):
            g = gtab[k]
            g.xyz, g.rotation, g.scaling, g.opacity, g.features_dc, g.features_rest = (t.data_ptr() if t.numel() else None for t in grads[6 * k: 6 * k + 6])
        dposes = torch.empty((nseg, 8), device=dev, dtype=torch.float32)
        scratch = torch.empty((nseg, 16), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            rc = L.sgr_compose_backward(ctx.segs, gtab, nseg, M, _ptr(pose_tab), _ptr(idft_tab), _ptr(ctx.flip_full), _ptr(ctx.flip_quat),
                                        _ptr(g_xyz), _ptr(g_rot), _ptr(g_scale), _ptr(g_opac), _ptr(g_sh), _ptr(dposes), _ptr(scratch),
                                        _stream(dev))
        _capi.check(rc, "sgr_compose_backward")
        g_poses = None
        if ctx.pose_shape is not None:
            shape, dtype, device = ctx.pose_shape
            g_poses = dposes[1:, :7].to(device=device, dtype=dtype).reshape(shape)
        return (g_poses, None, None, None, None) + tuple(grads)


def compose(models: Sequence, poses: Optional[torch.Tensor] = None, idft: Optional[torch.Tensor] = None,
            flip: Optional[torch.Tensor] = None, flip_quat: Optional[torch.Tensor] = None):
    """-> (means3D[P,3], rotations[P,4], scales[P,3], opacities[P,1], shs[P,M,3]) of background + actors, differentiable."""
    raw = [_raw(m, k) for m in models for k in RAW_KEYS]
    n_act = len(models) - 1
    if n_act > 0:
        if poses is None or tuple(poses.shape) != (n_act, 7):
            raise ValueError(f"poses must be [{n_act}, 7] (qw, qx, qy, qz, tx, ty, tz per actor)")
        if idft is None or idft.shape[0] != n_act:
            raise ValueError(f"idft must be [{n_act}, fourier_dim]")
    if flip is not None and flip_quat is None:
        raise ValueError("flip mask given without flip_quat (the reference's flip_matrix)")
    return _Compose.apply(poses, idft, flip, flip_quat, len(models), *raw)
