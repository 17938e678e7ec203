"""Tile-row sharding of one frame across the GPUs of a node (no counterpart in the reference, which is
single-GPU).  One process per GPU (torch.distributed); Gaussian parameters are replicated, pixels are partitioned:

  forward : every rank preprocesses all P Gaussians (HBM-bound, ~0.1 ms) but bins / sorts / blends only the 16-px tile
            rows it owns -> no collective.  Each rank's output images are zero outside its rows.
  backward: each rank's blend_bwd produces PARTIAL per-Gaussian screen-space sums grad2d[P,12] (+ semantics[P,S]);
            ONE sum-all-reduce over NVLink (NCCL) makes them global, then the cheap per-Gaussian chain rule
            (preprocess_bwd) runs replicated, so every rank ends with the full, identical parameter gradients
            ("Variant B" of SURVEY.md §8e: 48 B/Gaussian on the wire instead of a 248 B/Gaussian all-gather).

Rows are dealt cyclically (row r -> rank r % world) so a horizon-heavy street scene balances without a histogram.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer, TileRowBand


def cyclic_band(image_height: int, rank: int, world: int) -> TileRowBand:
    rows = (int(image_height) + 15) // 16
    return TileRowBand(begin=rank, end=rows, step=world) if rank < rows else TileRowBand(begin=0, end=0, step=1)


def contiguous_band(image_height: int, rank: int, world: int) -> TileRowBand:
    rows = (int(image_height) + 15) // 16
    per = (rows + world - 1) // world
    return TileRowBand(begin=min(rows, rank * per), end=min(rows, (rank + 1) * per), step=1)


class ShardedGaussianRasterizer(GaussianRasterizer):
    """GaussianRasterizer whose forward covers this rank's tile rows and whose backward all-reduces the per-Gaussian
    screen-space sums.  With world == 1 it is exactly GaussianRasterizer."""

    def __init__(self, raster_settings: GaussianRasterizationSettings, group: Optional[dist.ProcessGroup] = None,
                 layout: str = "cyclic", capacity=None):
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
        band = None
        reduce = None
        if world > 1:
            mk = cyclic_band if layout == "cyclic" else contiguous_band
            band = mk(raster_settings.image_height, rank, world)

            def reduce(grad2d: torch.Tensor, g_sem: torch.Tensor):
                dist.all_reduce(grad2d, op=dist.ReduceOp.SUM, group=group)
                if g_sem.numel():
                    dist.all_reduce(g_sem, op=dist.ReduceOp.SUM, group=group)
                return grad2d, g_sem

        super().__init__(raster_settings, band=band, grad_reduce=reduce, capacity=capacity)
        self.group, self.world, self.rank = group, world, rank

    def gather_images(self, *images: torch.Tensor):
        """Sum the zero-padded per-band images into full frames on every rank (only needed when a full image is wanted
        on one device; the loss can be evaluated band-locally)."""
        if self.world == 1:
            return images
        out = []
        for im in images:
            full = im.detach().clone()
            dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
            out.append(full)
        return tuple(out)


def band_of_rows(image_height: int, rank: int, world: int, layout: str = "cyclic") -> torch.Tensor:
    """Boolean mask [H] of the pixel rows owned by `rank` (host-side helper for tests and band-local losses)."""
    band = (cyclic_band if layout == "cyclic" else contiguous_band)(image_height, rank, world)
    rows = torch.arange((int(image_height) + 15) // 16)
    own = (rows >= band.begin) & (rows < band.end) & (((rows - band.begin) % max(band.step, 1)) == 0)
    return own.repeat_interleave(16)[: int(image_height)]
