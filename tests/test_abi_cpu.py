"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/sgr.h declares,
the size queries work without a GPU, and the Python surface mirrors the reference API."""
import ctypes as C
import inspect
import os
import re

import pytest
import torch

import street_gaussians_b200 as sgb
from street_gaussians_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sgr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgr_[a-z0-9_]+)\s*\(", src)) - {"sgr_alloc_fn"})


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"libsgr.so does not export {s}"
    assert sorted(_capi.SYMBOLS) == syms


def test_abi_version_and_sizes_no_gpu():
    L = _capi.lib()
    assert L.sgr_abi_version() == 3
    fr = _capi.SgrFrame()
    fr.P, fr.width, fr.height, fr.D, fr.M = 1000, 640, 480, 3, 16
    fr.tan_fovx, fr.tan_fovy, fr.scale_modifier = 0.5, 0.4, 1.0
    g, i = C.c_size_t(), C.c_size_t()
    assert L.sgr_state_sizes(C.byref(fr), C.byref(g), C.byref(i)) == 0
    assert g.value >= 1000 * 48 and g.value % 256 == 0
    assert i.value >= 640 * 480 * 4
    assert L.sgr_binning_bytes(0) > 0
    assert L.sgr_binning_bytes(1000) >= 1000 * 16
    # invalid frame -> error code + message, no crash
    fr.width = 0
    assert L.sgr_state_sizes(C.byref(fr), C.byref(g), C.byref(i)) == -1
    assert b"bad sizes" in L.sgr_last_error()
    fr.width = 640
    fr.row_begin, fr.row_end, fr.row_step = 5, 2, 1
    assert L.sgr_state_sizes(C.byref(fr), C.byref(g), C.byref(i)) == -1
    assert b"band" in L.sgr_last_error()
    assert L.sgr_knn_scratch_bytes(1000) > 1000 * 16


def test_settings_fields_match_reference_order():
    assert sgb.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix", "sh_degree",
        "campos", "prefiltered", "debug")


def test_rasterizer_signature_matches_reference():
    sig = inspect.signature(sgb.GaussianRasterizer.forward)
    assert list(sig.parameters)[1:] == ["means3D", "means2D", "opacities", "shs", "colors_precomp", "scales", "rotations",
                                        "cov3D_precomp", "semantics"]
    assert all(sig.parameters[k].default is None for k in ("shs", "colors_precomp", "scales", "rotations", "cov3D_precomp", "semantics"))
    sig = inspect.signature(sgb.rasterize_gaussians)
    assert list(sig.parameters) == ["means3D", "means2D", "sh", "colors_precomp", "semantics", "opacities", "scales", "rotations",
                                    "cov3Ds_precomp", "raster_settings"]
    sig = inspect.signature(sgb.GaussianRasterizer.visible_filter)
    assert list(sig.parameters)[1:] == ["means3D", "scales", "rotations", "cov3D_precomp"]
    assert hasattr(sgb.GaussianRasterizer, "markVisible")


def _settings():
    return sgb.GaussianRasterizationSettings(16, 16, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False)


def test_argument_validation_matches_reference_messages():
    r = sgb.GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, None, torch.zeros(4, 1), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, None, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=torch.zeros(4, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m, None, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m, None, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4), cov3D_precomp=torch.zeros(4, 6))


def test_cpu_tensors_fail_loudly_no_fallback():
    """The product path must not silently compute on the CPU."""
    r = sgb.GaussianRasterizer(_settings())
    with pytest.raises(_capi.SgrError, match="no CPU fallback"):
        r(torch.zeros(4, 3), None, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(_capi.SgrError):
        sgb.distCUDA2(torch.zeros(4, 3))
    with pytest.raises(RuntimeError, match=r"\(num_points, 3\)"):
        r(torch.zeros(4, 2), None, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: no module of the product package may reference it."""
    pkg = os.path.join(ROOT, "street_gaussians_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "sgr_oracle" not in txt, f


def test_shims_resolve_to_this_package():
    import importlib
    import sys
    sgb.install_shims()
    dgr = importlib.import_module("diff_gaussian_rasterization")
    assert dgr.GaussianRasterizer is sgb.GaussianRasterizer
    assert dgr.GaussianRasterizationSettings is sgb.GaussianRasterizationSettings
    knn = importlib.import_module("simple_knn._C")
    assert knn.distCUDA2 is sgb.distCUDA2
    for n in ("diff_gaussian_rasterization", "simple_knn", "simple_knn._C"):
        sys.modules.pop(n, None)


def test_band_helpers():
    from street_gaussians_b200.sharded import band_of_rows, contiguous_band, cyclic_band
    H = 1280  # 80 tile rows
    for world in (1, 2, 4, 8, 3):
        for layout, mk in (("cyclic", cyclic_band), ("contiguous", contiguous_band)):
            cover = torch.zeros(H, dtype=torch.int32)
            for r in range(world):
                cover += band_of_rows(H, r, world, layout).int()
            assert (cover == 1).all(), (world, layout)
    b = cyclic_band(100, 3, 8)  # 7 tile rows, rank 3 owns row 3 only
    assert (b.begin, b.end, b.step) == (3, 7, 8)
    assert cyclic_band(16, 5, 8).end == 0  # more ranks than rows -> empty band


def test_debug_mode_dumps_inputs_like_the_reference(tmp_path, monkeypatch):
    """debug=True: a failing forward leaves snapshot_fw.dump with the 20 arguments of the reference's
    _C.rasterize_gaussians call (DGR/diff_gaussian_rasterization/__init__.py:87-94)."""
    monkeypatch.chdir(tmp_path)
    st = _settings()._replace(debug=True)
    r = sgb.GaussianRasterizer(st)
    with pytest.raises(_capi.SgrError):
        r(torch.zeros(4, 3), None, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    dump = torch.load(tmp_path / "snapshot_fw.dump")
    assert len(dump) == 20 and dump[1].shape == (4, 3) and dump[-1] is True


def test_instance_capacity_bookkeeping():
    cap = sgb.InstanceCapacity(headroom=1.5)
    assert cap.capacity is None
    cap.observe(1000)
    assert cap.capacity == 1500 + 4096
    cap.observe(10)          # never shrinks
    assert cap.capacity == 1500 + 4096
    cap.check()              # nothing pending: no-op
