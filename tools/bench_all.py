"""All BASELINE.json configs (A, A_native, B, C, E) on one GPU: candidate vs compiled reference, fwd and fwd+bwd,
CUDA-event medians, forward AND gradient parity of the two on the same tensors.  Sweep variants of config C probe how
scene-specific the speed-up is: C_sh1 (SH degree 1, the shipped Waymo setting,
configs/example/waymo_train_002.yaml:18), C_s0.02 / C_s0.05 (splat scale median 0.02 / 0.05 m instead of 0.007: 0.05 is
SURVEY.md §8d's literal value and makes the reference instantiate ~200 tiles per Gaussian).
Writes gpurun_out/bench_all.json."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
import street_gaussians_b200 as sgb
from street_gaussians_b200 import synthetic

ref = util.load_ref() if util.ref_available() else None
dev = "cuda"
out = {}


def timeit(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return dict(median_ms=float(np.median(ts)), p10_ms=float(np.percentile(ts, 10)), p90_ms=float(np.percentile(ts, 90)), n=n)


VARIANTS = {"C_sh1": ("C", dict(sh_degree=1)), "C_s0.02": ("C", dict(scale_med=0.02)), "C_s0.05": ("C", dict(scale_med=0.05)),
            "B_sh1": ("B", dict(sh_degree=1))}

for name in sys.argv[1:] or ["A", "A_native", "B", "C", "E", "C_sh1", "C_s0.02", "C_s0.05"]:
    base, over = VARIANTS.get(name, (name, {}))
    scene = synthetic.make_config(base, seed=0, **over)
    cam = scene["cam"]
    P, W, H = scene["means3D"].shape[0], cam["image_width"], cam["image_height"]
    rec = dict(P=P, W=W, H=H, sh_degree=cam["sh_degree"])
    gc, gd, ga = (scene[k].to(dev) for k in ("grad_color", "grad_depth", "grad_alpha"))
    res = {}
    for label, mod in (("sgr", sgb), ("ref", ref)):
        if mod is None:
            continue
        rast = mod.GaussianRasterizer(util.settings_from(mod, cam, dev))
        ins = {k: scene[k].to(dev).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        m2d = torch.zeros((P, 3), device=dev, requires_grad=True)

        def fwd():
            with torch.no_grad():
                return rast(means3D=ins["means3D"], means2D=m2d, opacities=ins["opacities"], shs=ins["shs"], scales=ins["scales"],
                            rotations=ins["rotations"])

        def fwdbwd():
            for v in ins.values():
                v.grad = None
            c, r_, d, a, s = rast(means3D=ins["means3D"], means2D=m2d, opacities=ins["opacities"], shs=ins["shs"], scales=ins["scales"],
                                  rotations=ins["rotations"])
            torch.autograd.backward([c, d, a], [gc, gd, ga])

        n = 30 if P <= 2_000_000 else 10
        if name in ("C_s0.02", "C_s0.05"):
            n = 5
        rec[label + "_fwd"] = timeit(fwd, n)
        if name != "E":
            rec[label + "_fwdbwd"] = timeit(fwdbwd, n)
        o = fwd()
        res[label] = dict(color=o[0].cpu().numpy(), radii=o[1].cpu().numpy(), depth=o[2].cpu().numpy(), alpha=o[3].cpu().numpy())
        if name != "E":
            m2d.grad = None
            fwdbwd()
            res[label].update({"g_" + k: v.grad.detach().cpu().numpy() for k, v in ins.items()})
            res[label]["g_means2D"] = m2d.grad.detach().cpu().numpy()
        if label == "sgr":
            rec["num_instances_sgr"] = None
            try:
                from street_gaussians_b200 import rasterizer as R
                with torch.no_grad():
                    fs = R._forward_impl(ins["means3D"], ins["shs"], None, None, ins["opacities"], ins["scales"], ins["rotations"], None,
                                         util.settings_from(sgb, cam, dev), None)[5]
                rec["num_instances_sgr"] = int(fs.num_instances)
            except Exception as e:  # noqa: BLE001
                rec["num_instances_sgr"] = repr(e)
        rec["visible"] = int((o[1] > 0).sum())
        del rast, ins, m2d
        torch.cuda.empty_cache()
    if "ref" in res:
        d = np.abs(res["sgr"]["color"].astype(np.float64) - res["ref"]["color"])
        rec["rgb_maxabs_vs_ref"] = float(d.max()); rec["rgb_n_gt_1e-4"] = int((d > 1e-4).sum())
        rec["radii_mismatch"] = int((res["sgr"]["radii"] != res["ref"]["radii"]).sum())
        rec["alpha_maxabs_vs_ref"] = float(np.abs(res["sgr"]["alpha"] - res["ref"]["alpha"]).max())
        rec["grad_rel_vs_ref"] = {k: util.rel_err(res["sgr"][k], res["ref"][k]) for k in res["sgr"] if k.startswith("g_") and k in res["ref"]}
        if "sgr_fwdbwd" in rec:
            rec["speedup_fwdbwd"] = rec["ref_fwdbwd"]["median_ms"] / rec["sgr_fwdbwd"]["median_ms"]
    out[name] = rec
    print(name, json.dumps(rec))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_all.json"), "w"), indent=1)
