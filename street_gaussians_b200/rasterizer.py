"""Host side of the drop-in: the reference's Python surface over the C-ABI CUDA library.

Mirrors /root/reference/submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py (DGR below):
  * ``GaussianRasterizationSettings`` — same 12 fields in the same order (DGR :167-179);
  * ``GaussianRasterizer`` — ``forward`` / ``markVisible`` / ``visible_filter`` with the same signatures, defaults,
    argument checks (same two ``Exception`` messages, DGR :201-205) and the same 5-tuple result
    ``(color, radii, depth, alpha, semantic)`` (DGR :197-233, 186-195, 235-260);
  * ``rasterize_gaussians`` — the 10-positional-argument functional entry (DGR :21-44);
  * gradients for the same 8 inputs in the same order (DGR :152-163), including the 3-column ``means2D`` convention
    (x, y = NDC-scaled screen gradient, z = sum |gx|+|gy|).

PyTorch is plumbing here: it owns device memory (caching allocator), the current stream and autograd bookkeeping.
All arithmetic happens in libsgr.so (hand-written CUDA, sm_100a) through include/sgr.h.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Tuple

import torch
import torch.nn as nn

from . import _capi


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class TileRowBand(NamedTuple):
    """Tile rows (16 px each) owned by this process: r in [begin, end) with (r - begin) % step == 0."""
    begin: int
    end: int
    step: int = 1


def _ptr(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None
    return C.c_void_p(t.data_ptr())


def _dev_f32(t: torch.Tensor, device) -> torch.Tensor:
    """contiguous fp32 on `device` (the reference calls .contiguous() on every argument, DGR/rasterize_points.cu:92-118)."""
    if t.device != device or t.dtype != torch.float32:
        t = t.to(device=device, dtype=torch.float32)
    return t.contiguous()


def _none_if_empty(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None or t.numel() == 0 else t


def _make_frame(s: GaussianRasterizationSettings, P: int, M: int, S: int, device, band: Optional[TileRowBand]):
    keep = [_dev_f32(s.bg, device), _dev_f32(s.viewmatrix, device), _dev_f32(s.projmatrix, device), _dev_f32(s.campos, device)]
    fr = _capi.SgrFrame()
    fr.P, fr.D, fr.M, fr.S = int(P), int(s.sh_degree), int(M), int(S)
    fr.width, fr.height = int(s.image_width), int(s.image_height)
    fr.tan_fovx, fr.tan_fovy, fr.scale_modifier = float(s.tanfovx), float(s.tanfovy), float(s.scale_modifier)
    fr.prefiltered, fr.debug = int(bool(s.prefiltered)), int(bool(s.debug))
    if band is None:
        fr.row_begin = fr.row_end = fr.row_step = 0
    else:
        fr.row_begin, fr.row_end, fr.row_step = int(band.begin), int(band.end), int(band.step)
    fr.bg, fr.viewmatrix, fr.projmatrix, fr.campos = (k.data_ptr() for k in keep)
    return fr, keep


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dump(path: str, args) -> None:
    """CPU deep copy of the call's arguments, written with torch.save — the reference's debug post-mortem."""
    try:
        torch.save(tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args), path)
    except Exception:
        pass


class _ForwardState:
    """Caller-owned forward->backward state (the reference keeps geomBuffer/binningBuffer/imgBuffer byte tensors)."""
    __slots__ = ("geom", "img", "binning", "num_instances")


class InstanceCapacity:
    """Opt-in, sync-free binning (include/sgr.h: sgr_forward_bounded).  The reference — and this library by default —
    blocks the host once per forward to read the instance count R back and size the sort buffers.  With a capacity
    object the buffers are sized from the largest R seen so far (x `headroom`), nothing is read back inside forward, and
    the true count of each frame arrives asynchronously in pinned memory; `check()` (called at the start of the next
    forward and from `GaussianRasterizer.synchronize_capacity()`) raises if a frame did not fit, after growing the capacity.
    The first frame(s) run in exact mode to learn R."""

    def __init__(self, headroom: float = 1.25, initial: Optional[int] = None):
        self.headroom, self.capacity = float(headroom), (int(initial) if initial else None)
        self.gaussian_capacity = None  # depth-order slots of the compacted Gaussian-sharded forward (sgr_sharded_forward)
        self.frozen = False            # freeze(): capacities fixed, no per-frame status copy / event (CUDA-graph capture of a step)
        self._pending = []  # (pinned int32[4], cuda event) in submission order
        self._pool = []     # pinned status words ready for reuse: no pin_memory() (a cudaHostAlloc) inside the steady-state step

    def observe(self, R: int):
        want = int(R * self.headroom) + 4096
        if self.capacity is None or want > self.capacity:
            self.capacity = want

    def freeze(self, frozen: bool = True):
        """Stop tracking: forwards neither copy their status words back nor record events, so a whole step (forward + backward) is a
        fixed sequence of stream operations that torch.cuda.graph can capture and replay.  Overflows are then only visible through
        an explicit sgr_forward_status; unfreeze to resume tracking."""
        if frozen:
            self.check(wait=True)
        self.frozen = bool(frozen)
        return self

    def observe_gaussians(self, n: int):
        want = int(n * self.headroom) + 1024
        if self.gaussian_capacity is None or want > self.gaussian_capacity:
            self.gaussian_capacity = want

    def status_word(self) -> torch.Tensor:
        """A pinned int32[8] for sgr_forward_status_async; recycled once its frame has been checked."""
        return self._pool.pop() if self._pool else torch.zeros(8, dtype=torch.int32).pin_memory()

    def track(self, host_status: torch.Tensor, event):
        self._pending.append((host_status, event))

    def check(self, wait: bool = False):
        """Examine the frames whose status has arrived (all of them with wait=True).  Raises on the FIRST overflowed frame
        after growing the capacity; the frames after it stay pending, so a later call still reports them.  An overflowed
        frame was rendered incomplete: call check(wait=True) (GaussianRasterizer.synchronize_capacity()) before the
        optimiser step when that matters, and re-render."""
        pending, self._pending = self._pending, []
        for i, (host_status, ev) in enumerate(pending):
            if wait:
                ev.synchronize()
            if not ev.query():
                self._pending.append((host_status, ev))
                continue
            R, overflow, n_sel, timed_out = int(host_status[0]), int(host_status[1]), int(host_status[3]), int(host_status[4])
            self._pool.append(host_status)
            if timed_out:
                self._pending.extend(pending[i + 1:])
                raise _capi.SgrError(f"the device barrier of epoch {timed_out} timed out (a peer rank never arrived within 2 s): that frame "
                                     "was rendered from incomplete peer data")
            old, old_g = self.capacity, self.gaussian_capacity
            self.observe(R)
            if n_sel:
                self.observe_gaussians(n_sel)
            if overflow:
                self._pending.extend(pending[i + 1:])
                what = (f"instance capacity {old} overflowed (frame needed {R}); capacity raised to {self.capacity}" if overflow & 1 else
                        f"Gaussian capacity {old_g} overflowed (band holds {n_sel}); capacity raised to {self.gaussian_capacity}")
                raise _capi.SgrError(what + " — re-render that frame")


def _forward_impl(means3D, sh, colors_precomp, semantics, opacities, scales, rotations, cov3Ds_precomp,
                  settings: GaussianRasterizationSettings, band: Optional[TileRowBand], capacity: Optional["InstanceCapacity"] = None):
    L = _capi.lib()
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # DGR/rasterize_points.cu:58-60
    if not means3D.is_cuda:
        raise _capi.SgrError("street_gaussians_b200 rasterizer needs CUDA tensors (there is no CPU fallback)")
    device = means3D.device
    P = means3D.shape[0]
    H, W = int(settings.image_height), int(settings.image_width)
    sh, colors_precomp = _none_if_empty(sh), _none_if_empty(colors_precomp)
    scales, rotations, cov3Ds_precomp = _none_if_empty(scales), _none_if_empty(rotations), _none_if_empty(cov3Ds_precomp)
    S = int(semantics.shape[1]) if (semantics is not None and semantics.dim() == 2) else 0
    M = int(sh.shape[1]) if sh is not None else 0

    f32 = dict(device=device, dtype=torch.float32)
    whole = band is None
    alloc_img = torch.empty if whole else torch.zeros  # a partial band leaves foreign rows untouched -> hand back zeros there
    st = _ForwardState()
    st.num_instances = 0
    st.binning = None
    if P == 0:  # the reference short-circuits and returns zero-filled images (DGR/rasterize_points.cu:86)
        st.geom = st.img = None
        z = lambda c: torch.zeros((c, H, W), **f32)
        return z(3), torch.zeros((0,), device=device, dtype=torch.int32), z(1), z(1), z(S), st, None

    tensors = dict(means3D=_dev_f32(means3D, device), opacities=_dev_f32(opacities, device))
    for name, t in (("sh", sh), ("colors_precomp", colors_precomp), ("scales", scales), ("rotations", rotations),
                    ("cov3Ds_precomp", cov3Ds_precomp), ("semantics", semantics if S > 0 else None)):
        tensors[name] = _dev_f32(t, device) if t is not None else None

    color = alloc_img((3, H, W), **f32)
    depth = alloc_img((1, H, W), **f32)
    alpha = alloc_img((1, H, W), **f32)
    semantic = alloc_img((S, H, W), **f32)
    radii = torch.empty((P,), device=device, dtype=torch.int32)

    fr, keep = _make_frame(settings, P, M, S, device, band)
    gb, ib = C.c_size_t(0), C.c_size_t(0)
    _capi.check(L.sgr_state_sizes(C.byref(fr), C.byref(gb), C.byref(ib)), "sgr_state_sizes")
    st.geom = torch.empty((gb.value,), device=device, dtype=torch.uint8)
    st.img = torch.empty((ib.value,), device=device, dtype=torch.uint8)

    if capacity is not None and not capacity.frozen:
        capacity.check()
    if capacity is not None and capacity.capacity is not None:
        # bounded mode: no host synchronisation anywhere in this call
        cap = int(capacity.capacity)
        nbytes = int(L.sgr_binning_bytes(cap))
        st.binning = torch.empty((nbytes,), device=device, dtype=torch.uint8)
        with torch.cuda.device(device):
            rc = L.sgr_forward_bounded(C.byref(fr), _ptr(tensors["means3D"]), _ptr(tensors["sh"]), _ptr(tensors["colors_precomp"]),
                                       _ptr(tensors["semantics"]), _ptr(tensors["opacities"]), _ptr(tensors["scales"]),
                                       _ptr(tensors["rotations"]), _ptr(tensors["cov3Ds_precomp"]), _ptr(color), _ptr(depth), _ptr(alpha),
                                       _ptr(semantic), _ptr(radii), _ptr(st.geom), gb.value, _ptr(st.img), ib.value, _ptr(st.binning), nbytes,
                                       cap, _stream(device))
            _capi.check(rc, "sgr_forward_bounded")
            if not capacity.frozen:
                host_status = capacity.status_word()
                rc = L.sgr_forward_status_async(C.byref(fr), _ptr(st.geom), C.c_void_p(host_status.data_ptr()), _stream(device))
                _capi.check(rc, "sgr_forward_status_async")
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(device))
                capacity.track(host_status, ev)
        st.num_instances = cap
        del keep
        return color, radii, depth, alpha, semantic, st, tensors

    def _alloc(_user, nbytes):
        st.binning = torch.empty((int(nbytes),), device=device, dtype=torch.uint8)
        return st.binning.data_ptr()

    cb = _capi.ALLOC_FN(_alloc)
    bin_ptr, n_inst = C.c_void_p(), C.c_int64(0)
    with torch.cuda.device(device):
        rc = L.sgr_forward(C.byref(fr), _ptr(tensors["means3D"]), _ptr(tensors["sh"]), _ptr(tensors["colors_precomp"]),
                           _ptr(tensors["semantics"]), _ptr(tensors["opacities"]), _ptr(tensors["scales"]),
                           _ptr(tensors["rotations"]), _ptr(tensors["cov3Ds_precomp"]), _ptr(color), _ptr(depth), _ptr(alpha),
                           _ptr(semantic), _ptr(radii), _ptr(st.geom), gb.value, _ptr(st.img), ib.value, cb, None,
                           C.byref(bin_ptr), C.byref(n_inst), _stream(device))
    _capi.check(rc, "sgr_forward")
    st.num_instances = int(n_inst.value)
    if capacity is not None:
        capacity.observe(st.num_instances)
    del keep
    return color, radii, depth, alpha, semantic, st, tensors


def _backward_blend_impl(settings, band, st: _ForwardState, tensors, alpha, grad_color, grad_depth, grad_alpha, grad_semantic,
                         grad2d_out: Optional[torch.Tensor] = None):
    """Stage 1: returns (grad2d[P,12], dL_dsemantics[P,S]) — the per-rank partial sums under tile-row sharding.
    grad2d_out: write the sums there instead of a fresh tensor (peer-mapped workspace of the Gaussian-sharded mode)."""
    L = _capi.lib()
    means3D = tensors["means3D"]
    device, P = means3D.device, means3D.shape[0]
    S = int(tensors["semantics"].shape[1]) if tensors["semantics"] is not None else 0
    M = int(tensors["sh"].shape[1]) if tensors["sh"] is not None else 0
    fr, keep = _make_frame(settings, P, M, S, device, band)
    grad2d = grad2d_out if grad2d_out is not None else torch.empty((P, 12), device=device, dtype=torch.float32)
    g_sem = torch.empty((P, S), device=device, dtype=torch.float32)
    gc, gd, ga = _dev_f32(grad_color, device), _dev_f32(grad_depth, device), _dev_f32(grad_alpha, device)
    gs = _dev_f32(grad_semantic, device) if S > 0 else None
    with torch.cuda.device(device):
        rc = L.sgr_backward_blend(C.byref(fr), st.num_instances, _ptr(tensors["semantics"]), _ptr(st.geom), _ptr(st.binning),
                                  _ptr(st.img), _ptr(alpha), _ptr(gc), _ptr(gd), _ptr(ga), _ptr(gs), _ptr(grad2d), _ptr(g_sem),
                                  _stream(device))
    _capi.check(rc, "sgr_backward_blend")
    del keep
    return grad2d, g_sem


def _backward_geom_impl(settings, band, st: _ForwardState, tensors, radii, grad2d):
    """Stage 2: per-Gaussian chain rule.  Outputs are torch.empty — the kernel writes every element."""
    L = _capi.lib()
    means3D = tensors["means3D"]
    device, P = means3D.device, means3D.shape[0]
    S = int(tensors["semantics"].shape[1]) if tensors["semantics"] is not None else 0
    sh, colors, scales, rots, cov = (tensors[k] for k in ("sh", "colors_precomp", "scales", "rotations", "cov3Ds_precomp"))
    M = int(sh.shape[1]) if sh is not None else 0
    fr, keep = _make_frame(settings, P, M, S, device, band)
    e = lambda *shape: torch.empty(shape, device=device, dtype=torch.float32)
    g_means3D, g_means2D, g_opac = e(P, 3), e(P, 3), e(P, 1)
    g_sh = e(P, M, 3) if sh is not None else None
    g_colors = e(P, 3) if colors is not None else None
    g_scales = e(P, 3) if cov is None else None
    g_rots = e(P, 4) if cov is None else None
    g_cov = e(P, 6) if cov is not None else None
    with torch.cuda.device(device):
        rc = L.sgr_backward_geom(C.byref(fr), _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(scales), _ptr(rots), _ptr(cov), _ptr(radii),
                                 _ptr(st.geom), _ptr(grad2d), _ptr(g_means3D), _ptr(g_means2D), _ptr(g_sh), _ptr(g_colors),
                                 _ptr(g_opac), _ptr(g_scales), _ptr(g_rots), _ptr(g_cov), _stream(device))
    _capi.check(rc, "sgr_backward_geom")
    del keep
    return g_means3D, g_means2D, g_sh, g_colors, g_opac, g_scales, g_rots, g_cov


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, semantics, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, band=None, grad_reduce=None, capacity=None):
        try:
            color, radii, depth, alpha, semantic, st, tensors = _forward_impl(means3D, sh, colors_precomp, semantics, opacities, scales,
                                                                              rotations, cov3Ds_precomp, raster_settings, band, capacity)
        except Exception:
            if raster_settings.debug:  # same post-mortem as the reference (DGR/diff_gaussian_rasterization/__init__.py:87-94)
                _dump("snapshot_fw.dump", (raster_settings.bg, means3D, colors_precomp, semantics, opacities, scales, rotations,
                                           raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                                           raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy,
                                           raster_settings.image_height, raster_settings.image_width, sh, raster_settings.sh_degree,
                                           raster_settings.campos, raster_settings.prefiltered, raster_settings.debug))
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
            raise
        ctx.raster_settings, ctx.band, ctx.state, ctx.grad_reduce = raster_settings, band, st, grad_reduce
        ctx.shapes = tuple(None if t is None else (tuple(t.shape), t.device, t.dtype)
                           for t in (means3D, means2D, sh, colors_precomp, semantics, opacities, scales, rotations, cov3Ds_precomp))
        # The fp32-contiguous inputs go through save_for_backward like the reference's (DGR __init__.py:101), so autograd's
        # version counters catch an in-place update of a parameter between forward and backward instead of silently
        # differentiating the mutated values.
        ctx.tensor_keys = None if tensors is None else tuple(k for k, v in tensors.items() if v is not None)
        ctx.tensor_none = None if tensors is None else tuple(k for k, v in tensors.items() if v is None)
        ctx.save_for_backward(radii, alpha, *([] if tensors is None else [tensors[k] for k in ctx.tensor_keys]))
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha, semantic

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha, grad_semantic):
        radii, alpha, *saved = ctx.saved_tensors
        tensors = None
        if ctx.tensor_keys is not None:
            tensors = dict(zip(ctx.tensor_keys, saved))
            tensors.update({k: None for k in ctx.tensor_none})
        st, settings, band = ctx.state, ctx.raster_settings, ctx.band
        shapes = ctx.shapes

        def zeros_like_input(i):
            if shapes[i] is None:
                return None
            shape, device, dtype = shapes[i]
            return torch.zeros(shape, device=device, dtype=dtype)

        if tensors is None:  # P == 0
            return tuple(zeros_like_input(i) for i in range(9)) + (None, None, None, None)
        dev = tensors["means3D"].device
        H, W = int(settings.image_height), int(settings.image_width)
        S = int(tensors["semantics"].shape[1]) if tensors["semantics"] is not None else 0
        zimg = lambda c: torch.zeros((c, H, W), device=dev, dtype=torch.float32)
        grad_color = grad_color if grad_color is not None else zimg(3)
        grad_depth = grad_depth if grad_depth is not None else zimg(1)
        grad_alpha = grad_alpha if grad_alpha is not None else zimg(1)
        grad_semantic = grad_semantic if grad_semantic is not None else zimg(S)
        try:
            grad2d, g_sem = _backward_blend_impl(settings, band, st, tensors, alpha, grad_color, grad_depth, grad_alpha, grad_semantic)
            if ctx.grad_reduce is not None:  # multi-GPU: sum the per-band partial sums across ranks (one collective)
                grad2d, g_sem = ctx.grad_reduce(grad2d, g_sem)
            g_means3D, g_means2D, g_sh, g_colors, g_opac, g_scales, g_rots, g_cov = _backward_geom_impl(settings, band, st, tensors, radii, grad2d)
        except Exception:
            if settings.debug:  # DGR/diff_gaussian_rasterization/__init__.py:141-148
                _dump("snapshot_bw.dump", (settings.bg, tensors["means3D"], radii, tensors["colors_precomp"], tensors["scales"],
                                           tensors["rotations"], settings.scale_modifier, tensors["cov3Ds_precomp"], settings.viewmatrix,
                                           settings.projmatrix, settings.tanfovx, settings.tanfovy, grad_color, grad_depth, grad_alpha,
                                           grad_semantic, tensors["sh"], settings.sh_degree, settings.campos, alpha, tensors["semantics"],
                                           settings.debug))
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
            raise

        def fit(g, i):
            """cast/reshape a gradient to the original input's shape, dtype and device (None stays None)."""
            if shapes[i] is None:
                return None
            shape, device, dtype = shapes[i]
            if g is None:  # the input was an empty placeholder tensor
                return torch.zeros(shape, device=device, dtype=dtype)
            return g.reshape(shape).to(device=device, dtype=dtype)

        return (fit(g_means3D, 0), fit(g_means2D, 1), fit(g_sh, 2), fit(g_colors, 3), fit(g_sem if S > 0 else None, 4),
                fit(g_opac, 5), fit(g_scales, 6), fit(g_rots, 7), fit(g_cov, 8), None, None, None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, semantics, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, semantics, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, None, None, None)


class GaussianRasterizer(nn.Module):
    """Same surface as the reference class (DGR :181-260).  ``band`` / ``grad_reduce`` (tile-row sharding across GPUs,
    street_gaussians_b200.sharded) and ``capacity`` (sync-free binning, InstanceCapacity) are the only additions:
    keyword-only, default off."""

    def __init__(self, raster_settings, *, band: Optional[TileRowBand] = None, grad_reduce=None,
                 capacity: Optional[InstanceCapacity] = None):
        super().__init__()
        self.raster_settings = raster_settings
        self.band = band
        self.grad_reduce = grad_reduce
        self.capacity = capacity  # opt-in sync-free binning; share ONE InstanceCapacity across the rasterizers of a training loop

    def synchronize_capacity(self):
        """Wait for the outstanding frame statuses of the sync-free mode and raise if any frame overflowed."""
        if self.capacity is not None:
            self.capacity.check(wait=True)

    def markVisible(self, positions):
        L = _capi.lib()
        s = self.raster_settings
        with torch.no_grad():
            if not positions.is_cuda:
                raise _capi.SgrError("markVisible needs a CUDA tensor")
            dev = positions.device
            pos = _dev_f32(positions, dev)
            P = pos.shape[0]
            visible = torch.zeros((P,), device=dev, dtype=torch.bool)
            if P:
                view, proj = _dev_f32(s.viewmatrix, dev), _dev_f32(s.projmatrix, dev)
                with torch.cuda.device(dev):
                    rc = L.sgr_mark_visible(P, _ptr(pos), _ptr(view), _ptr(proj), C.c_void_p(visible.data_ptr()), _stream(dev))
                _capi.check(rc, "sgr_mark_visible")
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, semantics=None):
        raster_settings = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = lambda: torch.Tensor([])
        shs = empty() if shs is None else shs
        colors_precomp = empty() if colors_precomp is None else colors_precomp
        scales = empty() if scales is None else scales
        rotations = empty() if rotations is None else rotations
        cov3D_precomp = empty() if cov3D_precomp is None else cov3D_precomp
        if semantics is None:
            semantics = torch.zeros(means3D.shape[0], 0, dtype=torch.float32, device=means3D.device)
        return _RasterizeGaussians.apply(means3D, means2D, shs, colors_precomp, semantics, opacities, scales, rotations,
                                         cov3D_precomp, raster_settings, self.band, self.grad_reduce, self.capacity)

    def visible_filter(self, means3D, scales=None, rotations=None, cov3D_precomp=None) -> Tuple[torch.Tensor, torch.Tensor]:
        L = _capi.lib()
        s = self.raster_settings
        with torch.no_grad():
            if means3D.dim() != 2 or means3D.shape[1] != 3:
                raise RuntimeError("means3D must have dimensions (num_points, 3)")
            if not means3D.is_cuda:
                raise _capi.SgrError("visible_filter needs CUDA tensors")
            dev = means3D.device
            P = means3D.shape[0]
            radii = torch.zeros((P,), device=dev, dtype=torch.int32)
            means2D = torch.zeros((P, 2), device=dev, dtype=torch.float32)
            if P:
                m = _dev_f32(means3D, dev)
                sc, ro, cv = _none_if_empty(scales), _none_if_empty(rotations), _none_if_empty(cov3D_precomp)
                sc = _dev_f32(sc, dev) if sc is not None else None
                ro = _dev_f32(ro, dev) if ro is not None else None
                cv = _dev_f32(cv, dev) if cv is not None else None
                fr, keep = _make_frame(s, P, 0, 0, dev, None)
                with torch.cuda.device(dev):
                    rc = L.sgr_visible_filter(C.byref(fr), _ptr(m), _ptr(sc), _ptr(ro), _ptr(cv), _ptr(radii), _ptr(means2D), _stream(dev))
                _capi.check(rc, "sgr_visible_filter")
                del keep
        return radii, means2D


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """simple_knn._C.distCUDA2 (KNN/spatial.cu:14-26): mean squared distance to the 3 nearest neighbours, [P] fp32."""
    L = _capi.lib()
    if not points.is_cuda:
        raise _capi.SgrError("distCUDA2 needs a CUDA tensor")
    dev = points.device
    pts = _dev_f32(points, dev)
    P = pts.shape[0]
    out = torch.zeros((P,), device=dev, dtype=torch.float32)
    if P:
        nbytes = L.sgr_knn_scratch_bytes(P)
        scratch = torch.empty((nbytes,), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            rc = L.sgr_knn_mean_dist2(P, _ptr(pts), _ptr(out), _ptr(scratch), nbytes, _stream(dev))
        _capi.check(rc, "sgr_knn_mean_dist2")
    return out
