"""Image-space losses of the training step over sgr_image_loss / sgr_sky_loss (include/sgr.h).

Same signatures as the reference's lib/utils/loss_utils.py so `from street_gaussians_b200.losses import l1_loss, ssim` replaces
`from lib.utils.loss_utils import l1_loss, ssim` (train.py:16) unchanged:
    l1_loss(network_output, gt, mask=None)                                   loss_utils.py:21-37
    ssim(img1, img2, window_size=11, size_average=True, mask=None)            loss_utils.py:91-126
and the fused form of train.py:101-104,
    photometric_loss(image, gt, mask, lambda_l1, lambda_dssim) = (1 - l) * l1w * L1 + l * (1 - SSIM),
which produces the value and dL/dimage in two kernels; `sky_loss(acc, sky_mask, weight)` is train.py:107-113.
The gradient image is computed in the forward call (it costs one more kernel) and handed to autograd in backward, so the
rasterizer's backward receives it without replaying ~20 PyTorch kernels.  CUDA tensors only: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _capi
from .rasterizer import _ptr, _stream


def _prep(img: torch.Tensor) -> torch.Tensor:
    if not img.is_cuda:
        raise _capi.SgrError("street_gaussians_b200.losses needs CUDA tensors (there is no CPU fallback)")
    return img if (img.dtype == torch.float32 and img.is_contiguous()) else img.to(torch.float32).contiguous()


class _ImageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, mask, w_l1: float, w_ssim: float, which: int):
        L = _capi.lib()
        img, g = _prep(image), _prep(gt.detach())
        if img.dim() != 3 or img.shape != g.shape:
            raise ValueError(f"image / gt must both be [C, H, W], got {tuple(image.shape)} and {tuple(gt.shape)}")
        Cn, H, W = (int(v) for v in img.shape)
        dev = img.device
        m = None
        if mask is not None:
            m = mask.reshape(-1).to(device=dev, dtype=torch.uint8).contiguous()
            if m.numel() != H * W:
                raise ValueError("mask must be [1, H, W]")
        need_grad = image.requires_grad
        grad = torch.empty_like(img) if need_grad else None
        scalars = torch.empty(4, device=dev, dtype=torch.float32)
        nbytes = int(L.sgr_image_loss_scratch_bytes(Cn, H, W))
        scratch = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            rc = L.sgr_image_loss(Cn, H, W, _ptr(img), _ptr(g), _ptr(m), float(w_l1), float(w_ssim), _ptr(grad), _ptr(scalars), _ptr(scratch),
                                  nbytes, _stream(dev))
        _capi.check(rc, "sgr_image_loss")
        ctx.save_for_backward(grad) if need_grad else None
        ctx.has_grad, ctx.in_dtype = need_grad, image.dtype
        return scalars[which]

    @staticmethod
    def backward(ctx, g_out):
        if not ctx.has_grad:
            return None, None, None, None, None, None
        (grad,) = ctx.saved_tensors
        return (grad * g_out).to(ctx.in_dtype), None, None, None, None, None


def l1_loss(network_output, gt, mask=None):
    """mean |network_output - gt| over the masked pixels (all pixels without a mask); inputs [C, H, W], mask [1, H, W] bool."""
    return _ImageLoss.apply(network_output, gt, mask, 1.0, 0.0, 0)


def l2_loss(network_output, gt, mask=None):
    """Not on the hot path (the reference's train.py never calls it); plain torch like the reference."""
    a, b = network_output.permute(1, 2, 0), gt.permute(1, 2, 0)
    if mask is not None:
        a, b = a[mask.squeeze(0)], b[mask.squeeze(0)]
    return ((a - b) ** 2).mean()


def ssim(img1, img2, window_size=11, size_average=True, mask=None):
    """Mean SSIM with the reference's 11x11 / sigma 1.5 Gaussian window, zero padding and masking convention."""
    if window_size != 11 or not size_average:
        raise _capi.SgrError("the fused SSIM implements the configuration the reference trains with: window_size=11, size_average=True")
    return _ImageLoss.apply(img1, img2, mask, 0.0, 1.0, 0)


def photometric_loss(image, gt, mask=None, lambda_l1: float = 1.0, lambda_dssim: float = 0.2):
    """(1 - lambda_dssim) * lambda_l1 * L1 + lambda_dssim * (1 - SSIM)   (train.py:101-104), one fused evaluation."""
    return _ImageLoss.apply(image, gt, mask, (1.0 - lambda_dssim) * lambda_l1, -lambda_dssim, 0) + lambda_dssim


class _SkyLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acc, sky_mask, weight: float):
        L = _capi.lib()
        a = _prep(acc)
        dev = a.device
        m = sky_mask.reshape(-1).to(device=dev, dtype=torch.uint8).contiguous()
        if m.numel() != a.numel():
            raise ValueError("sky_mask must have the shape of acc")
        need_grad = acc.requires_grad
        grad = torch.empty_like(a) if need_grad else None
        scalars = torch.empty(2, device=dev, dtype=torch.float32)
        scratch = torch.empty(64, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            rc = L.sgr_sky_loss(a.numel(), _ptr(a), _ptr(m), float(weight), _ptr(grad), _ptr(scalars), _ptr(scratch), _stream(dev))
        _capi.check(rc, "sgr_sky_loss")
        ctx.save_for_backward(grad) if need_grad else None
        ctx.has_grad, ctx.in_dtype = need_grad, acc.dtype
        return scalars[0]

    @staticmethod
    def backward(ctx, g_out):
        if not ctx.has_grad:
            return None, None, None
        (grad,) = ctx.saved_tensors
        return (grad * g_out).to(ctx.in_dtype), None, None


def sky_loss(acc, sky_mask, weight: float = 1.0):
    """weight * mean( sky ? -log(1 - acc) : -log(acc) ) with acc clamped to [1e-6, 1 - 1e-6]   (train.py:107-113)."""
    return _SkyLoss.apply(acc, sky_mask, weight)
