"""GPU parity tests: the CUDA path (libsgr.so through the reference-compatible API / C ABI) against
  (1) the CPU oracle on seeded inputs at sizes the oracle finishes in seconds,
  (2) the committed golden fixtures (outputs of the unmodified reference CUDA rasterizer),
  (3) the compiled reference itself (oracle/_ref) when it travelled to this box, incl. geomBuffer-level bit checks,
  (4) size-independent properties at BASELINE.json's full sizes.
Tolerances (BASELINE.json north_star): forward RGB within 1e-4, gradients within 1e-3 (max|d| / max|ref| per tensor)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest
import torch

import util
import street_gaussians_b200 as sgb
from street_gaussians_b200 import _capi, synthetic

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
FWD_TOL, GRAD_TOL = 1e-4, 1e-3
# The CPU oracle is plain C (no FMA contraction, glibc expf): the reference's hard per-pair thresholds (alpha < 1/255,
# T(1-alpha) < 1e-4) flip on a few (pixel, splat) pairs relative to ANY nvcc build, which moves per-tensor gradient
# maxima by a few 1e-3 on these tiny scenes (the compiled reference differs from the oracle by the same amount —
# tools/first_light.py).  The 1e-3 bar of BASELINE.json is enforced against the reference itself (golden fixtures and
# the live oracle/_ref build); against the oracle the bar is 5e-3.
ORACLE_GRAD_TOL = 5e-3


def oracle_maxv(fw, scene):
    """Largest blendable value per output image, for the threshold-flip bound of util.check_forward_flip_protocol."""
    g = fw.geom()
    vis = fw.radii > 0
    m = dict(color=float(g["rgb"][vis].max()) + float(scene["cam"]["bg"].abs().max()) if vis.any() else 1.0,
             depth=float(g["depth"][vis].max()) if vis.any() else 1.0, alpha=1.0)
    if "semantics" in scene:
        m["semantic"] = float(scene["semantics"].abs().max())
    return m


def assert_forward_close(res, ref, npx, allow_flips=0, maxv=None):
    assert (np.asarray(res["radii"]) == np.asarray(ref["radii"])).all(), "radii must match exactly"
    if maxv is not None:  # flip protocol: list every pixel > 1e-4, each must be explained by <= 2 threshold flips
        names = [k for k in ("color", "depth", "alpha", "semantic") if k in ref and np.asarray(ref[k]).size]
        off = util.check_forward_flip_protocol(res, ref, maxv, names=names, max_pixels=allow_flips)
        if off:
            print("threshold-flip pixels:", off[:20])
        return
    for k in ("color", "depth", "alpha", "semantic"):
        if k in ref and np.asarray(ref[k]).size:
            d = np.abs(np.asarray(res[k], np.float64) - np.asarray(ref[k], np.float64))
            scale = max(1.0, float(np.abs(ref[k]).max())) if k == "depth" else 1.0  # depth is un-normalised metres
            n_bad = int((d > FWD_TOL * scale).sum())
            assert n_bad <= allow_flips, (k, n_bad, float(d.max()))


def assert_grads_close(res, ref, tol=GRAD_TOL):
    n = 0
    for k, v in ref.items():
        if k.startswith("g_") and v is not None and res.get(k) is not None and np.asarray(v).size:
            e = util.rel_err(res[k], v)
            assert e <= tol, (k, e)
            n += 1
    assert n >= 5


def flips_allowed(npx):
    # plain-C oracle vs FMA-contracted GPU arithmetic: the reference's hard thresholds (alpha < 1/255, T(1-a) < 1e-4,
    # power > 0) can flip on isolated pixels; each flip is bounded by ~(1/255)*|c| (SURVEY.md §7 "Discontinuities")
    return max(3, npx // 2000)


SMALL = [
    ("sh3", dict(P=3000, width=208, height=120, sh_degree=3, seed=21, pose=True, scale_med=0.06), {}),
    ("sh0_odd_size", dict(P=2500, width=203, height=77, sh_degree=0, seed=22, pose=True, scale_med=0.06), {}),
    ("sh1_whitebg", dict(P=2500, width=160, height=96, sh_degree=1, seed=23, pose=True, scale_med=0.05, bg=(1.0, 1.0, 1.0)), {}),
    ("sh2_sem3", dict(P=2000, width=128, height=96, sh_degree=2, seed=24, pose=True, scale_med=0.06, semantics=3), {}),
    ("sh3_sem15", dict(P=1500, width=128, height=80, sh_degree=3, seed=25, pose=True, scale_med=0.06, semantics=15), {}),
    ("sem20", dict(P=1200, width=96, height=64, sh_degree=1, seed=26, pose=True, scale_med=0.06, semantics=20), {}),
    ("big_splats", dict(P=800, width=320, height=208, sh_degree=3, seed=27, pose=True, scale_med=0.5), {}),
]


@pytest.mark.parametrize("name,kw,opts", SMALL, ids=[s[0] for s in SMALL])
def test_cuda_vs_oracle_small(name, kw, opts):
    scene = synthetic.make_scene(**kw)
    mine = util.run_api(sgb, scene)
    orc = util.run_oracle(scene)
    maxv = oracle_maxv(orc.pop("_fw"), scene)
    npx = kw["width"] * kw["height"]
    assert_forward_close(mine, orc, npx, allow_flips=flips_allowed(npx), maxv=maxv)
    assert_grads_close(mine, orc, tol=ORACLE_GRAD_TOL)


def test_cuda_vs_oracle_colors_precomp_and_cov3d():
    scene = synthetic.make_scene(P=2000, width=160, height=96, sh_degree=0, seed=31, pose=True, scale_med=0.06)
    gen = torch.Generator().manual_seed(5)
    scene["colors_precomp"] = torch.rand(2000, 3, generator=gen)
    mine = util.run_api(sgb, scene, use_colors_precomp=True)
    orc = util.run_oracle(scene, use_colors_precomp=True)
    fw = orc.pop("_fw")
    npx = 160 * 96
    assert_forward_close(mine, orc, npx, allow_flips=flips_allowed(npx))
    assert util.rel_err(mine["g_colors_precomp"], orc["g_colors_precomp"]) < ORACLE_GRAD_TOL
    # cov3D_precomp path: feed the oracle's own cov3D back in
    cov = torch.from_numpy(fw.geom()["cov3d"])
    mine2 = util.run_api(sgb, scene, use_colors_precomp=True, use_cov3d=cov)
    orc2 = util.run_oracle(scene, use_colors_precomp=True, use_cov3d=cov)
    orc2.pop("_fw")
    assert_forward_close(mine2, orc2, npx, allow_flips=flips_allowed(npx))
    assert util.rel_err(mine2["g_cov3D_precomp"], orc2["g_cov3D_precomp"]) < ORACLE_GRAD_TOL
    assert util.rel_err(mine2["g_means3D"], orc2["g_means3D"]) < ORACLE_GRAD_TOL


def test_smoke_script_replay_vs_oracle():
    """script/test_gaussian_rasterization.py replayed (seeded): un-normalised quaternions, U[0,1) everything, with and
    without 15 semantic channels."""
    for S in (0, 15):
        scene = synthetic.smoke_script_scene(num_points=3000, width=311, height=94, seed=3, semantics=S)
        mine = util.run_api(sgb, scene)
        orc = util.run_oracle(scene)
        maxv = oracle_maxv(orc.pop("_fw"), scene)
        npx = 311 * 94
        assert_forward_close(mine, orc, npx, allow_flips=flips_allowed(npx), maxv=maxv)
        assert_grads_close(mine, orc, tol=ORACLE_GRAD_TOL)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_vs_reference_golden(path):
    """Committed outputs of the unmodified reference CUDA rasterizer: RGB within 1e-4, gradients within 1e-3."""
    from test_oracle_cpu import scene_from_npz
    z = np.load(path)
    scene = scene_from_npz(z)
    use_cp = "in_colors_precomp" in z.files and "in_shs" not in z.files
    mine = util.run_api(sgb, scene, use_colors_precomp=use_cp)
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    npx = scene["cam"]["image_height"] * scene["cam"]["image_width"]
    assert_forward_close(mine, ref, npx, allow_flips=0)
    assert_grads_close(mine, ref, tol=GRAD_TOL)


def test_golden_present():
    assert len(GOLDEN) >= 1


def test_callsite_replay_vs_reference():
    """SURVEY.md §8 a13, GPU half.  tests/golden/callsite/render_kernel.npz is the rasterizer call that the reference's UNMODIFIED
    StreetGaussianRenderer.render_kernel made for a real StreetGaussianModel (recorded on the build container by
    tests/golden/make_callsite_golden.py; tests/test_callsite_cpu.py pins the Python surface there).  Here the same call goes
    through the compiled reference (oracle/_ref) and through this library: outputs and every .grad — including all three
    columns of viewspace_points.grad — must agree to the north-star tolerances.  Without oracle/_ref the CPU oracle stands in."""
    from test_oracle_cpu import scene_from_npz
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "callsite", "render_kernel.npz")
    scene = scene_from_npz(np.load(path))
    H, W = scene["cam"]["image_height"], scene["cam"]["image_width"]
    g = torch.Generator().manual_seed(1)
    for k, c in (("grad_color", 3), ("grad_depth", 1), ("grad_alpha", 1)):
        scene[k] = torch.randn(c, H, W, generator=g) / (H * W)
    mine = util.run_api(sgb, scene)
    assert mine["g_means2D"].shape == (scene["means3D"].shape[0], 3) and np.abs(mine["g_means2D"][:, 2]).max() > 0
    if util.ref_available():
        r = util.run_api(util.load_ref(), scene)
        assert_forward_close(mine, r, H * W, allow_flips=0)
        assert_grads_close(mine, r, tol=GRAD_TOL)
    else:
        orc = util.run_oracle(scene)
        maxv = oracle_maxv(orc.pop("_fw"), scene)
        assert_forward_close(mine, orc, H * W, allow_flips=flips_allowed(H * W), maxv=maxv)
        assert_grads_close(mine, orc, tol=ORACLE_GRAD_TOL)


needs_ref = pytest.mark.skipif(not util.ref_available(), reason="oracle/_ref (compiled reference) did not travel to this box")


@needs_ref
@pytest.mark.parametrize("kw", [dict(P=200_000, width=1280, height=720, sh_degree=3, seed=41, pose=True),
                                dict(P=60_000, width=1000, height=600, sh_degree=2, seed=42, pose=True, semantics=3)],
                         ids=["200k_720p", "60k_sem3"])
def test_cuda_vs_live_reference_medium(kw):
    ref = util.load_ref()
    scene = synthetic.make_scene(**kw)
    mine = util.run_api(sgb, scene)
    r = util.run_api(ref, scene)
    assert_forward_close(mine, r, kw["width"] * kw["height"], allow_flips=0)
    assert_grads_close(mine, r)


def _parse_ref_geom(buf: torch.Tensor, P: int):
    """Decode the reference's private GeometryState layout (DGR/cuda_rasterizer/rasterizer_impl.cu:155-170): 128-B aligned
    depths f32[P], clamped bool[3P], radii i32[P], means2D float2[P], cov3D f32[6P], conic_opacity float4[P], rgb f32[3P], tiles u32[P]."""
    base = buf.data_ptr()
    raw = buf.cpu().numpy()
    off = 0

    def take(nbytes, dtype, shape):
        nonlocal off
        a = (base + off + 127) // 128 * 128 - base
        arr = raw[a:a + nbytes].view(dtype).reshape(shape)
        off = a + nbytes
        return arr

    return dict(depth=take(4 * P, np.float32, (P,)), clamped=take(3 * P, np.uint8, (P, 3)), radii=take(4 * P, np.int32, (P,)),
                xy=take(8 * P, np.float32, (P, 2)), cov3d=take(24 * P, np.float32, (P, 6)), conic_opacity=take(16 * P, np.float32, (P, 4)),
                rgb=take(12 * P, np.float32, (P, 3)), tiles=take(4 * P, np.uint32, (P,)))


@needs_ref
def test_geometry_bitwise_vs_reference():
    """The sort key is the raw float bits of the view depth and radius/rect are integer: these must be bit-equal to the
    reference's GeometryState; conic / pixel position / RGB are compared in ulps."""
    ref = util.load_ref()
    scene = synthetic.make_scene(P=100_000, width=1280, height=720, sh_degree=3, seed=43, pose=True)
    dev = "cuda"
    P = 100_000
    st = util.settings_from(ref, scene["cam"], dev)
    args = (st.bg, scene["means3D"].to(dev), torch.Tensor([]), torch.zeros(P, 0, device=dev), scene["opacities"].to(dev),
            scene["scales"].to(dev), scene["rotations"].to(dev), st.scale_modifier, torch.Tensor([]), st.viewmatrix, st.projmatrix,
            st.tanfovx, st.tanfovy, st.image_height, st.image_width, scene["shs"].to(dev), st.sh_degree, st.campos, False, False)
    n_ref, color, depth, alpha, sem, radii, geom, binning, img = ref._C.rasterize_gaussians(*args)
    g = _parse_ref_geom(geom, P)
    # candidate state through the C ABI
    from street_gaussians_b200 import rasterizer as R
    mst = util.settings_from(sgb, scene["cam"], dev)
    with torch.no_grad():
        col, rad, dep, alp, se, fst, tens = R._forward_impl(scene["means3D"].to(dev), scene["shs"].to(dev), None, None,
                                                           scene["opacities"].to(dev), scene["scales"].to(dev), scene["rotations"].to(dev),
                                                           None, mst, None)
    torch.cuda.synchronize()
    rec = fst.geom[:P * 48].cpu().numpy().view(np.float32).reshape(P, 12)
    ref_radii = radii.cpu().numpy()  # (GeometryState.internal_radii is unused when the caller passes a radii tensor)
    vis = ref_radii > 0
    assert (rad.cpu().numpy() == ref_radii).all()
    assert vis.sum() > 50_000
    # record layout: q0 = (px, py, conic.xx, conic.xy), q1 = (conic.yy, opacity, power_min, depth), q2 = (r, g, b, clamp bits)
    assert (rec[vis, 7].view(np.uint32) == g["depth"][vis].view(np.uint32)).all(), "view depth must be bit-equal (it is the sort key)"

    def ulps(a, b):
        a = a.astype(np.float32).view(np.int32).astype(np.int64); b = b.astype(np.float32).view(np.int32).astype(np.int64)
        return np.abs(a - b)

    assert ulps(rec[vis, 0:2], g["xy"][vis]).max() == 0, "pixel positions"
    assert ulps(rec[vis][:, [2, 3, 4]], g["conic_opacity"][vis, :3]).max() <= 2, "conic"
    assert (rec[vis, 5] == g["conic_opacity"][vis, 3]).all(), "opacity"
    mine_rgb = np.stack([rec[vis, 8], rec[vis, 9], rec[vis, 10]], 1)
    # SH colour is a 16-term signed sum: near a zero crossing a 1e-7 absolute difference is thousands of ulps, so bound
    # the absolute error and require ulp-level agreement for (almost) all entries
    assert np.abs(mine_rgb - g["rgb"][vis]).max() <= 1e-6, "SH colour (abs)"
    assert (ulps(mine_rgb, g["rgb"][vis]) <= 4).mean() > 0.98, "SH colour (ulps)"
    clamp = rec[vis, 11].view(np.uint32)
    assert (((clamp[:, None] >> np.arange(3)) & 1) == g["clamped"][vis]).all()
    # exact tile culling only ever REMOVES instances, and never changes the image
    assert fst.num_instances <= n_ref
    np.testing.assert_allclose(col.cpu().numpy(), color.cpu().numpy(), atol=1e-6)
    print(f"instances: reference {n_ref}, this library {fst.num_instances} ({fst.num_instances / max(n_ref, 1):.2%})")


# ---------------------------------------------------------------- properties at full size -----------------------------
def _scene_B():
    return synthetic.make_config("B", seed=0)


def test_full_size_properties_config_B():
    """BASELINE config B (500k x 1920x1280, SH3): determinism, ranges, band-sharding exactness, zero-grad upstream -> zero grads."""
    scene = _scene_B()
    dev = "cuda"
    a = util.run_api(sgb, scene, backward=False)
    b = util.run_api(sgb, scene, backward=False)
    for k in ("color", "depth", "alpha", "radii"):
        assert (a[k] == b[k]).all(), f"forward must be run-to-run deterministic ({k})"
    assert np.isfinite(a["color"]).all() and np.isfinite(a["depth"]).all()
    assert a["alpha"].min() >= 0 and a["alpha"].max() <= 1.0 + 1e-5
    assert (a["radii"] >= 0).all()
    # union of two disjoint cyclic tile-row bands == whole frame, bit for bit
    from street_gaussians_b200.sharded import cyclic_band
    parts = [util.run_api(sgb, scene, backward=False, rasterizer_kwargs=dict(band=cyclic_band(1280, r, 2))) for r in range(2)]
    for k in ("color", "depth", "alpha"):
        assert ((parts[0][k] + parts[1][k]) == a[k]).all(), k
    assert (parts[0]["radii"] == a["radii"]).all()


def test_sharded_backward_sums_to_whole():
    """grad2d partial sums over bands add up to the single-GPU result (the multi-GPU all-reduce contract)."""
    from street_gaussians_b200 import rasterizer as R
    from street_gaussians_b200.sharded import contiguous_band, cyclic_band
    scene = synthetic.make_scene(P=50_000, width=800, height=608, sh_degree=3, seed=51, pose=True)
    dev = "cuda"
    st = util.settings_from(sgb, scene["cam"], dev)
    t = {k: scene[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations", "grad_color", "grad_depth", "grad_alpha")}

    def run(band):
        with torch.no_grad():
            col, rad, dep, alp, se, fst, tens = R._forward_impl(t["means3D"], t["shs"], None, None, t["opacities"], t["scales"],
                                                               t["rotations"], None, st, band)
            g2d, _ = R._backward_blend_impl(st, band, fst, tens, alp, t["grad_color"], t["grad_depth"], t["grad_alpha"], None)
        return g2d.double().cpu().numpy()

    whole = run(None)
    for mk, world in ((cyclic_band, 3), (contiguous_band, 4)):
        parts = sum(run(mk(608, r, world)) for r in range(world))
        assert util.rel_err(parts, whole) < 1e-5


@pytest.mark.parametrize("S,world,layout", [(0, 3, "cyclic"), (3, 4, "contiguous")])
def test_gaussian_sharded_emulated_ranks_match_single_gpu(S, world, layout):
    """Gaussian-sharded mode (sgr_project / sgr_forward_records; SURVEY.md §8e variant A) with the N ranks played one after
    the other on one GPU: concatenation stands in for the all-gather, a sum + slice for the reduce-scatter.  Forward
    images must be BIT-identical to the single-GPU render, gradients equal up to float summation order."""
    from street_gaussians_b200 import rasterizer as R
    from street_gaussians_b200 import sharded as SH
    P = 50_001  # not divisible by the world size -> the last rank carries padding slots
    scene = synthetic.make_scene(P=P, width=800, height=608, sh_degree=3, seed=52, pose=True, semantics=S)
    dev = "cuda"
    st = util.settings_from(sgb, scene["cam"], dev)
    t = {k: scene[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations", "grad_color", "grad_depth", "grad_alpha")}
    sem = scene["semantics"].to(dev) if S > 0 else None
    g_sem_img = scene["grad_semantic"].to(dev) if S > 0 else None
    with torch.no_grad():
        col, rad, dep, alp, se, fst, tens = R._forward_impl(t["means3D"], t["shs"], None, sem, t["opacities"], t["scales"], t["rotations"],
                                                           None, st, None)
        g2d, gsem = R._backward_blend_impl(st, None, fst, tens, alp, t["grad_color"], t["grad_depth"], t["grad_alpha"], g_sem_img)
        ref_grads = R._backward_geom_impl(st, None, fst, tens, rad, g2d)

        chunk = (P + world - 1) // world
        P_total = chunk * world
        local, recs, radii = [], [], []
        for r in range(world):
            sl = slice(r * chunk, min(P, (r + 1) * chunk))
            lt = SH._local_tensors(t["means3D"][sl], t["shs"][sl], None, sem[sl] if S > 0 else None, t["opacities"][sl], t["scales"][sl],
                                   t["rotations"][sl], None)
            rec_r, rad_r = SH.project_records(lt, st, chunk)
            local.append(lt); recs.append(rec_r); radii.append(rad_r)
        rec_cat, radii_all = torch.cat(recs), torch.cat(radii)
        assert torch.equal(radii_all[:P], rad) and int(radii_all[P:].abs().sum()) == 0
        sem_all = torch.cat([sem, sem.new_zeros((P_total - P, S))]) if S > 0 else None
        mk = SH.cyclic_band if layout == "cyclic" else SH.contiguous_band
        imgs, g2d_sum, gsem_sum = None, 0, 0
        for r in range(world):
            band = mk(608, r, world)
            fs, rec_all, gb, ib = SH.alloc_gathered(st, P_total, S, torch.device(dev))
            rec_all.copy_(rec_cat)
            out = SH.forward_records(st, band, fs, (gb, ib), radii_all, sem_all)
            imgs = out if imgs is None else tuple(a + b for a, b in zip(imgs, out))
            part, part_sem = SH.backward_blend_records(st, band, fs, P_total, sem_all, out[2], t["grad_color"], t["grad_depth"],
                                                       t["grad_alpha"], g_sem_img)
            g2d_sum = g2d_sum + part.double()
            gsem_sum = gsem_sum + part_sem.double()
        for a, b, name in zip(imgs, (col, dep, alp, se), ("color", "depth", "alpha", "semantic")):
            assert torch.equal(a, b), f"{name} differs from the single-GPU render"
        assert util.rel_err(g2d_sum[:P].cpu().numpy(), g2d.double().cpu().numpy()) < 1e-5
        assert float(g2d_sum[P:].abs().max()) == 0.0 if P_total > P else True
        if S > 0:
            assert util.rel_err(gsem_sum[:P].cpu().numpy(), gsem.double().cpu().numpy()) < 1e-5
        parts = []
        for r in range(world):
            g_slice = g2d_sum[r * chunk:(r + 1) * chunk].float().contiguous()
            parts.append(SH.backward_geom_local(st, local[r], recs[r], radii[r], g_slice))
        for i, ref in enumerate(ref_grads):
            if ref is None:
                assert all(p[i] is None for p in parts)
                continue
            got = torch.cat([p[i] for p in parts])
            assert got.shape == ref.shape
            assert util.rel_err(got.double().cpu().numpy(), ref.double().cpu().numpy()) < 2e-5, i


def _warm_up_fused_path(dev):
    """Load every kernel of the fused Gaussian-sharded step BEFORE ranks are emulated on one GPU.  The emulation keeps the spinning
    barrier kernel of "rank 0" resident while the host enqueues "rank 1"; with CUDA's lazy module loading the FIRST launch of a kernel
    loads it at launch time, which synchronises the context — i.e. waits for the spinning kernel, which waits for work the blocked
    host thread has not enqueued yet: a deadlock that only the barrier's 2 s bound resolves (observed on B200: rank 0 then ran on
    incomplete peer data; the barrier-timeout status word reports it).  One process per GPU — the real deployment — cannot deadlock
    this way: a peer's host thread is never blocked by this process's loads."""
    from street_gaussians_b200 import sharded as SH
    scene = synthetic.make_scene(P=3000, width=160, height=96, sh_degree=3, seed=1, pose=True)
    st = util.settings_from(sgb, scene["cam"], dev)
    lt = SH._local_tensors(*(scene[k].to(dev) for k in ("means3D", "shs")), None, None, scene["opacities"].to(dev), scene["scales"].to(dev),
                           scene["rotations"].to(dev), None)
    up = [scene[k].to(dev) for k in ("grad_color", "grad_depth", "grad_alpha")]
    with torch.no_grad():
        for gcap in (3000, -1):
            ws = SH.PeerWorkspace.emulate(st, 3000, 1, dev)[0]
            col, dep, alp, _ = SH.sharded_forward_raw(st, None, ws, lt, 3000, 500_000, gcap)
            SH.sharded_backward_raw(st, None, ws, lt, 3000, 500_000, alp, *up)
        pair = SH.PeerWorkspace.emulate(st, 8, 2, dev)  # the barrier kernel itself: nothing spins yet when rank 0's launch loads it
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        torch.cuda.synchronize()
        for r in range(2):
            with torch.cuda.stream(streams[r]):
                _capi.check(_capi.lib().sgr_peer_barrier(C.byref(pair[r].peers), 1, C.c_void_p(streams[r].cuda_stream)), "sgr_peer_barrier")
        torch.cuda.synchronize()


@pytest.mark.parametrize("world,compact", [(3, True), (2, False), (4, True), (2, True), (4, False)])
def test_fused_sharded_step_emulated_ranks(world, compact):
    """sgr_sharded_forward / sgr_sharded_backward (ONE C-ABI call each: project+scatter, device barrier, compacted depth sort, bin,
    blend | blend_bwd, device barrier, chain rule with the peer gather folded in) with the N ranks' workspaces on ONE GPU.  Each
    rank runs on its own stream — the barrier kernels of the ranks must be co-resident, exactly as on N GPUs.  Two frames, each
    rendered twice: first with every depth-order slot (what the host does before it has seen a count), then with the per-rank
    count that pass reported (+16) — the compacted depth order in its steady state.  Consecutive steps exercise the epoch
    bookkeeping, the in-forward zeroing of the grad2d rows and stale data of the previous step; images bit-identical to the
    single-GPU render, gradients equal up to float summation order."""
    from street_gaussians_b200 import rasterizer as R
    from street_gaussians_b200 import sharded as SH
    P, H, W = 40_003, 608, 800
    dev = torch.device("cuda")
    _warm_up_fused_path(dev)
    frames = [synthetic.make_scene(P=P, width=W, height=H, sh_degree=3, seed=60 + i, pose=True) for i in range(2)]
    chunk = (P + world - 1) // world
    st0 = util.settings_from(sgb, frames[0]["cam"], dev)
    wss = SH.PeerWorkspace.emulate(st0, chunk, world, dev)
    for ws in wss:
        ws.buf[: ws.off_flags].fill_(0x7f)  # poison everything but the barrier pads: no array needs a particular content on entry
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    for fi, scene in enumerate(frames):
        st = util.settings_from(sgb, scene["cam"], dev)
        t = {k: scene[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations", "grad_color", "grad_depth", "grad_alpha")}
        with torch.no_grad():
            col, rad, dep, alp, se, fst, tens = R._forward_impl(t["means3D"], t["shs"], None, None, t["opacities"], t["scales"], t["rotations"],
                                                               None, st, None)
            g2d, _ = R._backward_blend_impl(st, None, fst, tens, alp, t["grad_color"], t["grad_depth"], t["grad_alpha"], None)
            ref_grads = R._backward_geom_impl(st, None, fst, tens, rad, g2d)
            capacity = int(fst.num_instances) + 1000  # every band fits
            local = []
            for r in range(world):
                sl = slice(r * chunk, min(P, (r + 1) * chunk))
                local.append(SH._local_tensors(t["means3D"][sl], t["shs"][sl], None, None, t["opacities"][sl], t["scales"][sl],
                                               t["rotations"][sl], None))
            n_sel_rank = [0] * world
            for ps in range(2 if compact else 1):
                tag = f"frame {fi} pass {ps}"
                outs, status = [], [torch.zeros(8, dtype=torch.int32).pin_memory() for _ in range(world)]
                torch.cuda.synchronize()
                for r in range(world):
                    with torch.cuda.stream(streams[r]):
                        gcap = -1 if not compact else (chunk * world if ps == 0 else n_sel_rank[r] + 16)
                        outs.append(SH.sharded_forward_raw(st, SH.cyclic_band(H, r, world), wss[r], local[r], int(local[r]["means3D"].shape[0]),
                                                           capacity, gcap, status[r]))
                torch.cuda.synchronize()
                n_sel_total = 0
                for r in range(world):
                    R_r, over, emitted, n_sel, timed_out = (int(v) for v in status[r][:5])
                    assert timed_out == 0, f"{tag}, rank {r}: the device barrier of epoch {timed_out} timed out (ranks not co-scheduled on this GPU)"
                    assert over == 0 and R_r == emitted and R_r <= fst.num_instances, (tag, r, R_r, over, emitted)
                    assert n_sel > 0
                    n_sel_total += n_sel
                    n_sel_rank[r] = n_sel
                    assert torch.equal(outs[r][3]["radii"][: int(local[r]["means3D"].shape[0])], rad[r * chunk: min(P, (r + 1) * chunk)])
                imgs = [sum(o[i] for o in outs) for i in range(3)]
                for a, b, name in zip(imgs, (col, dep, alp), ("color", "depth", "alpha")):
                    assert torch.equal(a, b), f"{tag}: {name} differs from the single-GPU render"
                if True:  # each band counts / sorts the Gaussians delivered to it, not all of them
                    assert n_sel_total <= world * int((rad > 0).sum()) and (world < 4 or n_sel_total < 0.8 * world * int((rad > 0).sum()))
                grads = []
                for r in range(world):
                    with torch.cuda.stream(streams[r]):
                        grads.append(SH.sharded_backward_raw(st, SH.cyclic_band(H, r, world), wss[r], local[r], int(local[r]["means3D"].shape[0]),
                                                             capacity, outs[r][2], t["grad_color"], t["grad_depth"], t["grad_alpha"]))
                torch.cuda.synchronize()
                for i, ref in enumerate(ref_grads):
                    if ref is None:
                        assert all(g[i] is None for g in grads)
                        continue
                    err = [util.rel_err(grads[r][i].double().cpu().numpy(), ref[r * chunk: r * chunk + grads[r][i].shape[0]].double().cpu().numpy())
                           for r in range(world)]
                    assert max(err) < 5e-5, f"{tag}, output {i}: per-rank errors vs the single-GPU gradients {err}"
    # a forward that follows a forward (no backward): the leading barrier of sgr.h is taken (epochs advance by 2)
    with torch.no_grad():
        e0 = [ws.epoch for ws in wss]
        for rep in range(2):
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    SH.sharded_forward_raw(st, SH.cyclic_band(H, r, world), wss[r], local[r], int(local[r]["means3D"].shape[0]), capacity, -1)
        torch.cuda.synchronize()
        assert [ws.epoch - e for ws, e in zip(wss, e0)] == [3] * world


def test_gaussian_capacity_overflow_is_flagged():
    """More Gaussians in the band than depth-order slots: overflow bit 2, no out-of-bounds write (the sort covers cap slots)."""
    from street_gaussians_b200 import sharded as SH
    P, H, W = 20_000, 304, 400
    dev = torch.device("cuda")
    scene = synthetic.make_scene(P=P, width=W, height=H, sh_degree=1, seed=70, pose=True)
    st = util.settings_from(sgb, scene["cam"], dev)
    ws = SH.PeerWorkspace.emulate(st, P, 1, dev)[0]
    lt = SH._local_tensors(scene["means3D"].to(dev), scene["shs"].to(dev), None, None, scene["opacities"].to(dev), scene["scales"].to(dev),
                           scene["rotations"].to(dev), None)
    hs = torch.zeros(8, dtype=torch.int32).pin_memory()
    with torch.no_grad():
        SH.sharded_forward_raw(st, None, ws, lt, P, 2_000_000, 100, hs)
    torch.cuda.synchronize()
    assert int(hs[1]) & 2 and int(hs[3]) > 100


@pytest.mark.parametrize("world", [2, 5])
def test_gaussian_sharded_peer_exchange_emulated_ranks(world):
    """The NVLink peer-memory exchange (sgr_scatter_records / sgr_gather_grad2d) with the N ranks' workspaces living on ONE
    GPU and pointing at each other: images bit-identical to the single-GPU render, every rank receives exactly the
    Gaussians whose tile rectangle meets its band, gradients equal up to float summation order."""
    from street_gaussians_b200 import rasterizer as R
    from street_gaussians_b200 import sharded as SH
    P, H, W = 50_001, 608, 800
    scene = synthetic.make_scene(P=P, width=W, height=H, sh_degree=3, seed=54, pose=True)
    dev = torch.device("cuda")
    st = util.settings_from(sgb, scene["cam"], dev)
    t = {k: scene[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations", "grad_color", "grad_depth", "grad_alpha")}
    with torch.no_grad():
        col, rad, dep, alp, se, fst, tens = R._forward_impl(t["means3D"], t["shs"], None, None, t["opacities"], t["scales"], t["rotations"],
                                                           None, st, None)
        g2d, _ = R._backward_blend_impl(st, None, fst, tens, alp, t["grad_color"], t["grad_depth"], t["grad_alpha"], None)
        ref_grads = R._backward_geom_impl(st, None, fst, tens, rad, g2d)

        chunk = (P + world - 1) // world
        wss = SH.PeerWorkspace.emulate(st, chunk, world, dev)
        for ws in wss:  # poison: stale records / radii from "earlier frames" must not leak into this one
            ws.buf.fill_(0x7f)
        local, recs, radii = [], [], []
        for r in range(world):
            sl = slice(r * chunk, min(P, (r + 1) * chunk))
            lt = SH._local_tensors(t["means3D"][sl], t["shs"][sl], None, None, t["opacities"][sl], t["scales"][sl], t["rotations"][sl], None)
            rec_r, rad_r = SH.project_records(lt, st, chunk)
            SH.scatter_records(st, wss[r], rec_r, rad_r, int(lt["means3D"].shape[0]))
            local.append(lt); recs.append(rec_r); radii.append(rad_r)
        delivered = torch.zeros(P, dtype=torch.int32, device=dev)
        imgs = None
        for r in range(world):
            ws = wss[r]
            got = ws.radii_all[:P]
            assert bool(((got == 0) | (got == rad)).all()) and int(ws.radii_all[P:].abs().sum()) == 0
            delivered += (got > 0).int()
            fs = SH.peer_forward_state(ws)
            out = SH.forward_records(st, SH.cyclic_band(H, r, world), fs, (ws.geom_bytes, ws.img_bytes), ws.radii_all, None)
            imgs = out if imgs is None else tuple(a + b for a, b in zip(imgs, out))
            SH.backward_blend_records(st, SH.cyclic_band(H, r, world), fs, chunk * world, None, out[2], t["grad_color"], t["grad_depth"],
                                      t["grad_alpha"], None, grad2d_out=ws.grad2d)
        assert bool(((delivered > 0) == (rad > 0)).all())  # every visible Gaussian reached at least one rank, no invisible one did
        assert float(delivered.float().mean()) < 0.75 * world * float((rad > 0).float().mean()) or world == 2  # sparse, not an all-gather
        for a, b, name in zip(imgs, (col, dep, alp), ("color", "depth", "alpha")):
            assert torch.equal(a, b), f"{name} differs from the single-GPU render"
        parts, g2_cat = [], []
        for r in range(world):
            P_r = int(local[r]["means3D"].shape[0])
            g2_r = SH.gather_grad2d(st, wss[r], recs[r], radii[r], P_r)
            g2_cat.append(g2_r[:P_r])
            parts.append(SH.backward_geom_local(st, local[r], recs[r], radii[r], g2_r))
        assert util.rel_err(torch.cat(g2_cat).double().cpu().numpy(), g2d.double().cpu().numpy()) < 1e-5
        for i, ref in enumerate(ref_grads):
            if ref is None:
                continue
            got = torch.cat([p[i] for p in parts])
            assert util.rel_err(got.double().cpu().numpy(), ref.double().cpu().numpy()) < 2e-5, i


def test_gaussian_sharded_module_world1_matches_plain_rasterizer():
    """GaussianShardedRasterizer without a process group (world 1) runs the project -> records -> local chain-rule path end
    to end through autograd and must reproduce GaussianRasterizer."""
    from street_gaussians_b200.sharded import GaussianShardedRasterizer
    scene = synthetic.make_scene(P=20_000, width=640, height=400, sh_degree=2, seed=53, pose=True, semantics=2)
    dev = "cuda"
    st = util.settings_from(sgb, scene["cam"], dev)
    res = {}
    for name, mod in (("plain", sgb.GaussianRasterizer(st)), ("sharded", GaussianShardedRasterizer(st)),
                      ("sharded_bounded", GaussianShardedRasterizer(st, capacity=sgb.InstanceCapacity())),
                      ("sharded_p2p", GaussianShardedRasterizer(st, exchange="p2p"))):
        for rep in range(2 if name == "sharded_bounded" else 1):  # second call of the bounded module runs sync-free
            leaves = {k: scene[k].to(dev).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations", "semantics")}
            m2d = torch.zeros(20_000, 3, device=dev, requires_grad=True)
            color, radii, depth, alpha, semantic = mod(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                                        shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"],
                                                        semantics=leaves["semantics"])
            loss = (color * scene["grad_color"].to(dev)).sum() + (depth * scene["grad_depth"].to(dev)).sum() + \
                (alpha * scene["grad_alpha"].to(dev)).sum() + (semantic * scene["grad_semantic"].to(dev)).sum()
            loss.backward()
        mod.synchronize_capacity() if hasattr(mod, "synchronize_capacity") else None
        res[name] = dict(color=color.detach(), radii=radii, depth=depth.detach(), alpha=alpha.detach(), semantic=semantic.detach(),
                         grads={k: v.grad for k, v in leaves.items()}, m2d=m2d.grad)
    for name in ("sharded", "sharded_bounded", "sharded_p2p"):
        for k in ("color", "radii", "depth", "alpha", "semantic"):
            assert torch.equal(res[name][k], res["plain"][k]), (name, k)
        for k, g in res["plain"]["grads"].items():
            assert util.rel_err(res[name]["grads"][k].double().cpu().numpy(), g.double().cpu().numpy()) < 2e-5, (name, k)
        assert util.rel_err(res[name]["m2d"].double().cpu().numpy(), res["plain"]["m2d"].double().cpu().numpy()) < 2e-5


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL all-gather / reduce-scatter)")
def test_gaussian_sharded_two_ranks_nccl():
    """tools/check_gaussian_sharded.py under torchrun on 2 GPUs: GaussianShardedRasterizer vs the single-GPU rasterizer."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "check_gaussian_sharded.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "GAUSSIAN_SHARDED_CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_edge_cases():
    dev = "cuda"
    cam = synthetic.make_camera(100, 60, sh_degree=1, bg=(0.1, 0.2, 0.3))
    st = util.settings_from(sgb, cam, dev)
    rast = sgb.GaussianRasterizer(st)
    # P == 0: the reference returns zero-filled images (DGR/rasterize_points.cu:70-86)
    z = lambda *s: torch.zeros(*s, device=dev)
    color, radii, depth, alpha, sem = rast(means3D=z(0, 3), means2D=None, opacities=z(0, 1), shs=z(0, 4, 3), scales=z(0, 3), rotations=z(0, 4))
    assert color.shape == (3, 60, 100) and radii.shape == (0,) and float(color.abs().max()) == 0.0
    # all culled (behind the camera): background everywhere, R == 0, backward gives zeros
    P = 64
    m = torch.randn(P, 3, device=dev); m[:, 2] = -abs(m[:, 2]) - 1
    m.requires_grad_(True)
    sh = torch.randn(P, 4, 3, device=dev, requires_grad=True)
    color, radii, depth, alpha, sem = rast(means3D=m, means2D=None, opacities=torch.rand(P, 1, device=dev), shs=sh,
                                           scales=torch.rand(P, 3, device=dev), rotations=torch.rand(P, 4, device=dev))
    assert (radii == 0).all() and float(alpha.max()) == 0.0
    np.testing.assert_allclose(color[:, 0, 0].detach().cpu().numpy(), [0.1, 0.2, 0.3], rtol=1e-6)
    color.sum().backward()
    assert float(m.grad.abs().max()) == 0.0 and float(sh.grad.abs().max()) == 0.0
    assert sem.shape == (0, 60, 100)
    # means2D=None in eval and non-contiguous / requires_grad=False inputs are accepted
    sc = synthetic.make_scene(P=500, width=100, height=60, sh_degree=1, seed=61, scale_med=0.05)
    with torch.no_grad():
        out = rast(means3D=sc["means3D"].to(dev), means2D=None, opacities=sc["opacities"].to(dev), shs=sc["shs"].to(dev),
                   scales=sc["scales"].to(dev).t().contiguous().t(), rotations=sc["rotations"].to(dev))
    assert out[0].shape == (3, 60, 100)


def test_mark_visible_filter_and_knn():
    from oracle import oracle as O
    dev = "cuda"
    scene = synthetic.make_scene(P=20_000, width=640, height=400, sh_degree=0, seed=71, pose=True)
    st = util.settings_from(sgb, scene["cam"], dev)
    rast = sgb.GaussianRasterizer(st)
    vis = rast.markVisible(scene["means3D"].to(dev))
    assert vis.dtype == torch.bool
    assert (vis.cpu().numpy() == O.mark_visible(scene["means3D"], scene["cam"]["viewmatrix"])).all()
    radii, m2d = rast.visible_filter(scene["means3D"].to(dev), scales=scene["scales"].to(dev), rotations=scene["rotations"].to(dev))
    full = util.run_api(sgb, scene, backward=False)
    assert (radii.cpu().numpy() == full["radii"]).all() and m2d.shape == (20_000, 2)
    orc = util.run_oracle(scene, backward=False)
    g = orc.pop("_fw").geom()
    v = full["radii"] > 0
    np.testing.assert_allclose(m2d.cpu().numpy()[v], g["xy"][v], rtol=1e-5, atol=1e-3)
    # distCUDA2: exact 3-NN mean squared distance
    pts = torch.randn(5000, 3) * torch.tensor([3.0, 1.0, 0.2])
    d = sgb.distCUDA2(pts.to(dev)).cpu().numpy()
    np.testing.assert_allclose(d, O.knn_mean_dist2(pts), rtol=1e-5, atol=1e-9)
    if util.ref_available():
        rk = util.load_ref_knn()
        big = torch.rand(300_000, 3) * torch.tensor([50.0, 5.0, 80.0])
        a = sgb.distCUDA2(big.to(dev)).cpu().numpy()
        b = rk.distCUDA2(big.to(dev)).cpu().numpy()
        assert (a == b).all(), "distCUDA2 must be bit-identical to simple-knn"
        # visible_filter / markVisible against the reference's own entry points (_C.rasterize_gaussians_filter, _C.mark_visible)
        ref = util.load_ref()
        rr = ref.GaussianRasterizer(util.settings_from(ref, scene["cam"], dev))
        r_radii, r_m2d = rr.visible_filter(scene["means3D"].to(dev), scales=scene["scales"].to(dev), rotations=scene["rotations"].to(dev))
        assert torch.equal(radii, r_radii), "visible_filter radii must equal the reference's"
        vv = r_radii > 0
        assert torch.equal(m2d[vv], r_m2d[vv]), "visible_filter means2D must be bit-equal to the reference's for visible Gaussians"
        assert torch.equal(vis, rr.markVisible(scene["means3D"].to(dev)))


def test_direct_c_abi_error_paths():
    """Straight ctypes calls: bad argument combinations return error codes and messages instead of crashing."""
    L = _capi.lib()
    fr = _capi.SgrFrame()
    fr.P, fr.width, fr.height, fr.D, fr.M, fr.S = 10, 64, 64, 0, 1, 0
    fr.tan_fovx = fr.tan_fovy = 0.5
    fr.scale_modifier = 1.0
    dev = "cuda"
    cam = [torch.zeros(3, device=dev), torch.eye(4, device=dev), torch.eye(4, device=dev), torch.zeros(3, device=dev)]
    fr.bg, fr.viewmatrix, fr.projmatrix, fr.campos = (t.data_ptr() for t in cam)
    img = torch.zeros(3, 64, 64, device=dev)
    one = torch.zeros(1, 64, 64, device=dev)
    vp = lambda t: C.c_void_p(t.data_ptr())
    m = torch.zeros(10, 3, device=dev)
    rad = torch.zeros(10, dtype=torch.int32, device=dev)
    buf = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
    cb = _capi.ALLOC_FN(lambda u, n: 0)
    binp, ninst = C.c_void_p(), C.c_int64()
    # both shs and colors_precomp NULL
    rc = L.sgr_forward(C.byref(fr), vp(m), None, None, None, vp(m), vp(m), vp(m), None, vp(img), vp(one), vp(one), None, vp(rad), vp(buf),
                       1 << 20, vp(buf), 1 << 20, cb, None, C.byref(binp), C.byref(ninst), None)
    assert rc == -1 and b"exactly one of shs" in L.sgr_last_error()
    # geom buffer too small
    rc = L.sgr_forward(C.byref(fr), vp(m), vp(m), None, None, vp(m), vp(m), vp(m), None, vp(img), vp(one), vp(one), None, vp(rad), vp(buf),
                       16, vp(buf), 1 << 20, cb, None, C.byref(binp), C.byref(ninst), None)
    assert rc == -3 and b"geom_state too small" in L.sgr_last_error()
    fr.S = 33
    rc = L.sgr_backward_blend(C.byref(fr), 0, vp(m), vp(buf), None, vp(buf), vp(one), vp(img), vp(one), vp(one), vp(img), vp(buf), vp(buf), None)
    assert rc == -4


def test_bounded_sync_free_mode():
    """sgr_forward_bounded (no host read-back): identical images/gradients to the exact mode when the capacity suffices;
    overflow is detected, reported and recoverable."""
    scene = synthetic.make_scene(P=40_000, width=640, height=416, sh_degree=3, seed=81, pose=True)
    exact = util.run_api(sgb, scene)
    cap = sgb.InstanceCapacity(headroom=1.3)
    first = util.run_api(sgb, scene, rasterizer_kwargs=dict(capacity=cap))      # exact mode, learns R
    assert cap.capacity is not None and cap.capacity > 0
    bounded = util.run_api(sgb, scene, rasterizer_kwargs=dict(capacity=cap))    # bounded mode
    cap.check(wait=True)
    for k in ("color", "depth", "alpha", "radii"):
        assert (bounded[k] == exact[k]).all() and (first[k] == exact[k]).all(), k
    for k in exact:
        if k.startswith("g_") and exact[k] is not None:
            assert util.rel_err(bounded[k], exact[k]) < 1e-5, k
    # a capacity that is far too small: the frame is truncated, the status says so, the capacity grows, the retry is exact
    small = sgb.InstanceCapacity(initial=2000)
    truncated = util.run_api(sgb, scene, backward=False, rasterizer_kwargs=dict(capacity=small))
    assert np.isfinite(truncated["color"]).all()
    with pytest.raises(_capi.SgrError, match="overflowed"):
        small.check(wait=True)
    assert small.capacity > 2000
    retry = util.run_api(sgb, scene, backward=False, rasterizer_kwargs=dict(capacity=small))
    small.check(wait=True)
    assert (retry["color"] == exact["color"]).all()


@needs_ref
def test_full_size_parity_config_C_vs_live_reference():
    """BASELINE config C (1.9 M composed Gaussians, 1920x1280, SH 3): forward RGB within 1e-4 and every gradient tensor
    within 1e-3 of the compiled reference on identical inputs — the north-star's parity bar at the headline size."""
    ref = util.load_ref()
    scene = synthetic.make_config("C", seed=0)
    mine = util.run_api(sgb, scene)
    r = util.run_api(ref, scene)
    assert_forward_close(mine, r, 1920 * 1280, allow_flips=0)
    assert_grads_close(mine, r, tol=GRAD_TOL)
    # semantics of the densification statistic: column 2 of the means2D gradient is a sum of absolute values
    assert (mine["g_means2D"][:, 2] >= 0).all()
    vis = mine["radii"] > 0
    assert np.abs(mine["g_shs"][~vis]).max() == 0 and np.abs(mine["g_means3D"][~vis]).max() == 0


def test_config_E_forward_band_union_and_determinism():
    """BASELINE config E (8 M Gaussians, 3840x2160, forward only): run-to-run determinism and exactness of tile-row
    sharding at the stress size (size-independent properties; no reference needed)."""
    from street_gaussians_b200.sharded import cyclic_band
    scene = synthetic.make_config("E", seed=0)
    a = util.run_api(sgb, scene, backward=False)
    b = util.run_api(sgb, scene, backward=False)
    for k in ("color", "depth", "alpha", "radii"):
        assert (a[k] == b[k]).all(), k
    parts = [util.run_api(sgb, scene, backward=False, rasterizer_kwargs=dict(band=cyclic_band(2160, r, 4))) for r in range(4)]
    for k in ("color", "depth", "alpha"):
        assert (sum(p[k] for p in parts) == a[k]).all(), k
    assert np.isfinite(a["color"]).all() and a["alpha"].max() <= 1.0 + 1e-5


def test_bounded_mode_with_tile_row_band():
    """sync-free binning composed with tile-row sharding (what bench.py --gpus N runs): band results identical to the exact mode."""
    from street_gaussians_b200.sharded import cyclic_band
    scene = synthetic.make_scene(P=60_000, width=800, height=608, sh_degree=2, seed=91, pose=True, scale_med=0.02)
    band = cyclic_band(608, 1, 3)
    exact = util.run_api(sgb, scene, rasterizer_kwargs=dict(band=band))
    cap = sgb.InstanceCapacity()
    util.run_api(sgb, scene, backward=False, rasterizer_kwargs=dict(band=band, capacity=cap))  # learns R
    bounded = util.run_api(sgb, scene, rasterizer_kwargs=dict(band=band, capacity=cap))
    cap.check(wait=True)
    for k in ("color", "depth", "alpha", "radii"):
        assert (bounded[k] == exact[k]).all(), k
    for k in exact:
        if k.startswith("g_") and exact[k] is not None:
            assert util.rel_err(bounded[k], exact[k]) < 1e-5, k


def test_more_than_65536_tiles_uses_32bit_tile_ids():
    """4112 x 4112 px = 257 x 257 = 66,049 tiles: the u32-key paths of emit / sort / ranges (exact and bounded)."""
    scene = synthetic.make_scene(P=3000, width=4112, height=4112, sh_degree=1, seed=92, pose=True, scale_med=0.05)
    exact = util.run_api(sgb, scene, backward=False)
    orc = util.run_oracle(scene, backward=False)
    orc.pop("_fw")
    assert_forward_close(exact, orc, 4112 * 4112, allow_flips=flips_allowed(4112 * 4112))
    cap = sgb.InstanceCapacity()
    util.run_api(sgb, scene, backward=False, rasterizer_kwargs=dict(capacity=cap))
    bounded = util.run_api(sgb, scene, backward=False, rasterizer_kwargs=dict(capacity=cap))
    cap.check(wait=True)
    for k in ("color", "depth", "alpha"):
        assert (bounded[k] == exact[k]).all(), k


def test_forward_with_more_than_32_feature_channels():
    """Forward accepts any S (channel chunks of 32); backward is capped at SGR_MAX_SEMANTIC_BWD and says so."""
    scene = synthetic.make_scene(P=1500, width=128, height=96, sh_degree=0, seed=93, pose=True, scale_med=0.06, semantics=40)
    mine = util.run_api(sgb, scene, backward=False)
    orc = util.run_oracle(scene, backward=False)
    orc.pop("_fw")
    assert mine["semantic"].shape == (40, 96, 128)
    assert_forward_close(mine, orc, 128 * 96, allow_flips=flips_allowed(128 * 96))
    with pytest.raises(_capi.SgrError, match="at most 32 semantic channels"):
        util.run_api(sgb, scene, backward=True)
