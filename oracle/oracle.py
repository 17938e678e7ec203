"""ctypes binding of the CPU oracle (oracle/sgr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package (street_gaussians_b200/) never imports this module.

All arrays are numpy float32, laid out exactly like the tensors the reference API takes
(DGR/diff_gaussian_rasterization/__init__.py:197-233).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsgr_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile sgr_oracle.c with gcc (seconds). Called by __graft_entry__.build()."""
    src = os.path.join(_HERE, "sgr_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-fopenmp", "-shared", "-fPIC", "-ffp-contract=off", src, "-o", _LIB_PATH, "-lm"])
    return _LIB_PATH


class _Params(C.Structure):
    _fields_ = [("P", C.c_int32), ("D", C.c_int32), ("M", C.c_int32), ("S", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
                ("bg", C.c_float * 3), ("view", C.c_float * 16), ("proj", C.c_float * 16), ("campos", C.c_float * 3)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.or_forward.restype = C.c_void_p
        _lib.or_forward.argtypes = [C.POINTER(_Params)] + [C.c_void_p] * 13
        _lib.or_backward.restype = C.c_int
        _lib.or_backward.argtypes = [C.c_void_p] * 21
        _lib.or_free.argtypes = [C.c_void_p]
        _lib.or_num_rendered.restype = C.c_int64
        _lib.or_num_rendered.argtypes = [C.c_void_p]
        _lib.or_pairs_evaluated.restype = C.c_int64
        _lib.or_pairs_evaluated.argtypes = [C.c_void_p]
        _lib.or_pairs_blended.restype = C.c_int64
        _lib.or_pairs_blended.argtypes = [C.c_void_p]
        _lib.or_get_geom.argtypes = [C.c_void_p] * 8
        _lib.or_get_n_contrib.argtypes = [C.c_void_p, C.c_void_p]
        _lib.or_mark_visible.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.or_knn_mean_dist2.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        _lib.or_num_threads.restype = C.c_int
        _lib.or_set_num_threads.argtypes = [C.c_int]
    return _lib


def _f32(a):
    if a is None:
        return None
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return None if a is None or a.size == 0 else a.ctypes.data_as(C.c_void_p)


@dataclass
class Camera:
    """The 12 fields of GaussianRasterizationSettings (DGR/diff_gaussian_rasterization/__init__.py:167-179) as numpy."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: np.ndarray
    scale_modifier: float
    viewmatrix: np.ndarray
    projmatrix: np.ndarray
    sh_degree: int
    campos: np.ndarray


class Forward:
    """Result of one oracle forward; keeps the C state needed by backward()."""

    def __init__(self, cam: Camera, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                 cov3D_precomp=None, semantics=None):
        L = lib()
        self.cam = cam
        self.means3D = _f32(means3D)
        P = self.means3D.shape[0]
        self.shs, self.colors_precomp = _f32(shs), _f32(colors_precomp)
        self.scales, self.rotations, self.cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
        self.opacities = _f32(opacities)
        self.semantics = _f32(semantics) if semantics is not None else np.zeros((P, 0), np.float32)
        S = self.semantics.shape[1] if self.semantics.ndim == 2 else 0
        M = self.shs.shape[1] if self.shs is not None and self.shs.size else 0
        H, W = cam.image_height, cam.image_width
        pm = _Params(P, cam.sh_degree, M, S, W, H, cam.tanfovx, cam.tanfovy, cam.scale_modifier)
        pm.bg[:] = [float(v) for v in _f32(cam.bg).ravel()]
        pm.view[:] = [float(v) for v in _f32(cam.viewmatrix).ravel()]
        pm.proj[:] = [float(v) for v in _f32(cam.projmatrix).ravel()]
        pm.campos[:] = [float(v) for v in _f32(cam.campos).ravel()]
        self.P, self.S, self.M, self.H, self.W = P, S, M, H, W
        self.color = np.zeros((3, H, W), np.float32)
        self.depth = np.zeros((1, H, W), np.float32)
        self.alpha = np.zeros((1, H, W), np.float32)
        self.semantic = np.zeros((S, H, W), np.float32)
        self.radii = np.zeros((P,), np.int32)
        self._st = L.or_forward(C.byref(pm), _ptr(self.means3D), _ptr(self.shs), _ptr(self.colors_precomp),
                                _ptr(self.semantics), _ptr(self.opacities), _ptr(self.scales), _ptr(self.rotations),
                                _ptr(self.cov3D_precomp), _ptr(self.color), _ptr(self.depth), _ptr(self.alpha),
                                _ptr(self.semantic) if S else None, _ptr(self.radii))
        if not self._st:
            raise MemoryError("oracle forward failed")
        self.num_rendered = L.or_num_rendered(self._st)
        self.pairs_evaluated = L.or_pairs_evaluated(self._st)
        self.pairs_blended = L.or_pairs_blended(self._st)

    def geom(self):
        P = self.P
        out = dict(depth=np.zeros(P, np.float32), xy=np.zeros((P, 2), np.float32), conic_opacity=np.zeros((P, 4), np.float32),
                   rgb=np.zeros((P, 3), np.float32), clamped=np.zeros((P, 3), np.uint8), tiles=np.zeros(P, np.uint32),
                   cov3d=np.zeros((P, 6), np.float32))
        lib().or_get_geom(self._st, *[_ptr(out[k]) for k in ("depth", "xy", "conic_opacity", "rgb", "clamped", "tiles", "cov3d")])
        return out

    def n_contrib(self):
        out = np.zeros((self.H, self.W), np.uint32)
        lib().or_get_n_contrib(self._st, _ptr(out))
        return out

    def backward(self, grad_color, grad_depth, grad_alpha, grad_semantic=None):
        """Returns grads in the reference's order (DGR/diff_gaussian_rasterization/__init__.py:152-163)."""
        P, S, M = self.P, self.S, self.M
        gc, gd, ga = _f32(grad_color), _f32(grad_depth), _f32(grad_alpha)
        gs = _f32(grad_semantic) if S else None
        g = dict(means3D=np.zeros((P, 3), np.float32), means2D=np.zeros((P, 3), np.float32),
                 sh=np.zeros((P, M, 3), np.float32), colors_precomp=np.zeros((P, 3), np.float32),
                 semantics=np.zeros((P, S), np.float32), opacities=np.zeros((P, 1), np.float32),
                 scales=np.zeros((P, 3), np.float32), rotations=np.zeros((P, 4), np.float32),
                 cov3D_precomp=np.zeros((P, 6), np.float32))
        has_sh = self.shs is not None and self.shs.size
        has_sr = self.scales is not None and self.scales.size
        rc = lib().or_backward(self._st, _ptr(self.means3D), _ptr(self.shs), _ptr(self.colors_precomp), _ptr(self.semantics),
                               _ptr(self.scales), _ptr(self.rotations), _ptr(self.cov3D_precomp), _ptr(self.alpha),
                               _ptr(gc), _ptr(gd), _ptr(ga), _ptr(gs), _ptr(g["means3D"]), _ptr(g["means2D"]),
                               _ptr(g["sh"]) if has_sh else None, _ptr(g["colors_precomp"]), _ptr(g["semantics"]) if S else None,
                               _ptr(g["opacities"]), _ptr(g["scales"]) if has_sr else None, _ptr(g["rotations"]) if has_sr else None,
                               _ptr(g["cov3D_precomp"]))
        if rc:
            raise MemoryError("oracle backward failed")
        if has_sh:
            # the reference returns dL_dcolors for colors_precomp even on the SH path, but as the gradient of an
            # empty tensor it is discarded by autograd; keep it available for tests under 'colors_internal'
            g["colors_internal"] = g["colors_precomp"]
            g["colors_precomp"] = np.zeros((0,), np.float32)
        return g

    def close(self):
        if self._st:
            lib().or_free(self._st)
            self._st = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mark_visible(means3D, viewmatrix):
    m = _f32(means3D)
    out = np.zeros(m.shape[0], np.uint8)
    lib().or_mark_visible(m.shape[0], _ptr(m), _ptr(_f32(viewmatrix)), _ptr(out))
    return out.astype(bool)


def knn_mean_dist2(points):
    p = _f32(points)
    out = np.zeros(p.shape[0], np.float32)
    lib().or_knn_mean_dist2(p.shape[0], _ptr(p), _ptr(out))
    return out


def num_threads() -> int:
    return lib().or_num_threads()


def set_num_threads(n: int) -> None:
    lib().or_set_num_threads(int(n))
