"""Test infrastructure: import the UNMODIFIED reference call sites from /root/reference on a CPU-only box.

Used by tests/test_callsite_cpu.py and by tests/golden/make_callsite_golden.py /
tests/golden/make_compose_golden.py, which turn what the reference's own Python computes into committed fixtures
(the reference tree does not exist on the GPU box).  Nothing under street_gaussians_b200/ imports this.

What has to be faked to get `lib.models.street_gaussian_renderer` importable here (SURVEY.md header table):
  * six absent third-party modules that are only touched by PLY IO / slerp / sky / image saving: roma, plyfile, bidict,
    imageio, nvdiffrast, matplotlib  -> tiny sys.modules stubs;
  * argparse at import time (lib/config/config.py:150-158) -> sys.argv carries --config <example yaml> + source_path;
  * hard-coded `.cuda()` / device="cuda" in the reference (camera_utils.py:52-61, general_utils.py:130, ...): on a box
    without a GPU `Tensor.cuda()` becomes the identity and factory calls with device="cuda" are redirected to the CPU.
    On a GPU box nothing is patched.
  * `diff_gaussian_rasterization` / `simple_knn` resolve to street_gaussians_b200's shim packages (install_shims()).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_loaded = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "lib", "models"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Bidict(dict):
    @property
    def inverse(self):
        return {v: k for k, v in self.items()}


def _patch_cuda_to_cpu():
    if torch.cuda.is_available():
        return
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    for name in ("zeros", "ones", "empty", "tensor", "arange", "eye", "rand", "randn", "full", "zeros_like", "ones_like", "rand_like",
                 "empty_like", "linspace"):
        orig = getattr(torch, name)
        if getattr(orig, "_refharness", False):
            continue

        def wrap(*a, _o=orig, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k["device"] = "cpu"
            return _o(*a, **k)

        wrap._refharness = True
        setattr(torch, name, wrap)


def load(extra_opts=()):
    """Import the reference's lib.* with the fakes above.  Returns a namespace of the modules the tests use."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("/root/reference is not present on this box")
    _stub("imageio")

    def _unitquat_slerp(q0, q1, steps, shortest_arc=True):
        """Stand-in for roma.utils.unitquat_slerp (roma is not installed; x-y-z-w convention, returns [steps, ..., 4]).  Textbook
        slerp; it only produces the per-actor pose that is then fed IDENTICALLY to the reference's compose code and to ours."""
        d = (q0 * q1).sum(-1, keepdim=True)
        if shortest_arc:
            q1 = torch.where(d < 0, -q1, q1)
            d = d.abs()
        d = d.clamp(-1.0, 1.0)
        theta = torch.acos(d)
        s = torch.sin(theta)
        t = steps.view(-1, *([1] * q0.dim()))
        small = s.abs() < 1e-6
        w0 = torch.where(small, 1.0 - t, torch.sin((1.0 - t) * theta) / torch.where(small, torch.ones_like(s), s))
        w1 = torch.where(small, t, torch.sin(t * theta) / torch.where(small, torch.ones_like(s), s))
        return w0 * q0 + w1 * q1

    roma = _stub("roma")
    roma.utils = _stub("roma.utils", unitquat_slerp=_unitquat_slerp, unitquat_slerp_fast=_unitquat_slerp)
    _stub("bidict", bidict=_Bidict)
    _stub("plyfile", PlyData=object, PlyElement=object)
    nv = _stub("nvdiffrast")
    nv.torch = _stub("nvdiffrast.torch")
    mpl = _stub("matplotlib")
    mpl.__path__ = []  # a package, so that `import matplotlib.patches` resolves to the stub below
    for sub in ("pyplot", "cm", "patches", "colors"):
        setattr(mpl, sub, _stub("matplotlib." + sub))

    class _Cmap:  # img_utils.py:129-144 builds a transparent 'jet' colour map at import time
        N = 256

        def __call__(self, x):
            return np.zeros((len(x), 4))

        _lut = np.zeros((259, 4))

        def _init(self):
            pass

    mpl.pyplot.get_cmap = lambda *a, **k: _Cmap()
    _patch_cuda_to_cpu()
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import street_gaussians_b200 as sgb
    sgb.install_shims()
    if REF not in sys.path:
        sys.path.insert(1, REF)
    work = os.path.join("/tmp", "sgr_refharness")
    os.makedirs(work, exist_ok=True)
    os.environ.setdefault("PWD", work)
    saved = sys.argv
    sys.argv = ["refharness", "--config", os.path.join(REF, "configs", "example", "waymo_train_002.yaml"), "source_path", work,
                "model_path", os.path.join(work, "out")] + list(extra_opts)
    try:
        import lib.config as lc
        from lib.models import street_gaussian_renderer as rmod
        from lib.models import street_gaussian_model as mmod
        from lib.utils import camera_utils as cu
        from lib.utils import general_utils as gu
        from lib.utils import loss_utils as lu
        from lib.utils import sh_utils as shu
    finally:
        sys.argv = saved
    _loaded = types.SimpleNamespace(cfg=lc.cfg, renderer=rmod, model=mmod, camera_utils=cu, general_utils=gu, loss_utils=lu, sh_utils=shu,
                                    sgb=sgb)
    return _loaded


def make_camera(ns, width=208, height=120, frame=3, seed=0):
    """A reference `Camera` (lib/utils/camera_utils.py:18-75) looking down +z with a small ego pose."""
    g = np.random.default_rng(seed)
    R = np.eye(3, dtype=np.float64)
    T = np.zeros(3, dtype=np.float64)
    fovx = 2.0 * np.arctan(np.tan(np.deg2rad(25.0)))
    fovy = 2.0 * np.arctan(np.tan(fovx / 2) * height / width)
    ego = np.eye(4, dtype=np.float32)
    c, s = np.cos(0.05), np.sin(0.05)
    ego[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float32)
    ego[:3, 3] = np.array([0.3, -0.1, 0.5], dtype=np.float32)
    image = torch.from_numpy(g.random((3, height, width), dtype=np.float32))
    meta = dict(frame=frame, frame_idx=frame, is_val=False, timestamp=float(frame) * 0.1, cam=0, ego_pose=ego)
    return ns.camera_utils.Camera(id=0, R=R, T=T, FoVx=float(fovx), FoVy=float(fovy), K=None, image=image, image_name="synthetic",
                                  metadata=meta)


def make_street_model(ns, n_bkgd=1500, n_obj=2, per_obj=400, num_frames=8, seed=0, fourier_dim=None):
    """A real `StreetGaussianModel` (lib/models/street_gaussian_model.py:30-250) with random parameters: one background
    model + `n_obj` rigid actors with tracklets over `num_frames` frames.  Only __init__/setup_functions run from the
    reference; the parameters are then assigned directly (create_from_pcd reads point clouds from disk)."""
    g = torch.Generator().manual_seed(seed)
    cfg = ns.cfg
    C = int(fourier_dim if fourier_dim is not None else cfg.model.gaussian.get("fourier_dim", 1))
    tracklets = np.zeros((num_frames, n_obj, 8), dtype=np.float32)
    obj_meta = {}
    timestamps = np.arange(num_frames, dtype=np.float64) * 0.1
    for k in range(n_obj):
        yaw0 = 0.3 * (k + 1)
        for fidx in range(num_frames):
            yaw = yaw0 + 0.02 * fidx
            tracklets[fidx, k] = [k, -3.0 + 6.0 * k + 0.2 * fidx, 0.4, 12.0 + 5.0 * k + 0.5 * fidx, np.cos(yaw / 2), 0.0, np.sin(yaw / 2), 0.0]
        obj_meta[k] = dict(track_id=k, **{"class": "vehicle"}, class_label=0, deformable=False, start_frame=0, end_frame=num_frames - 1,
                           start_timestamp=float(timestamps[0]), end_timestamp=float(timestamps[-1]), length=4.5, width=2.0, height=1.6)
    metadata = dict(obj_tracklets=tracklets, obj_meta=obj_meta, tracklet_timestamps=timestamps,
                    camera_timestamps={0: dict(train_timestamps=list(timestamps), test_timestamps=[])},
                    scene_center=np.zeros(3, dtype=np.float32), scene_radius=20.0, sphere_center=np.zeros(3, dtype=np.float32), sphere_radius=20.0,
                    num_images=num_frames, num_cams=1, num_frames=num_frames)
    model = ns.model.StreetGaussianModel(metadata)
    deg = model.max_sh_degree
    M = (deg + 1) ** 2

    def fill(m, n, centre, spread, c_dim):
        rn = lambda *s, sd=1.0: torch.randn(*s, generator=g) * sd
        m._xyz = torch.nn.Parameter(rn(n, 3) * torch.tensor(spread) + torch.tensor(centre))
        m._features_dc = torch.nn.Parameter(rn(n, c_dim, 3))
        m._features_rest = torch.nn.Parameter(rn(n, M - 1, 3, sd=0.2))
        m._scaling = torch.nn.Parameter(torch.log(torch.tensor(0.08)) + rn(n, 3, sd=0.5))
        m._rotation = torch.nn.Parameter(rn(n, 4))
        m._opacity = torch.nn.Parameter(rn(n, 1, sd=2.0))
        m._semantic = torch.nn.Parameter(torch.zeros(n, m.num_classes))

    fill(model.background, n_bkgd, [0.0, 0.0, 20.0], [8.0, 3.0, 10.0], 1)
    for name in model.obj_list:
        fill(getattr(model, name), per_obj, [0.0, 0.0, 0.0], [1.5, 0.6, 0.5], C)
    return model
