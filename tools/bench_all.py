"""All BASELINE.json configs (A, A_native, B, C, E) on one GPU: candidate vs compiled reference, fwd and fwd+bwd,
CUDA-event medians, parity of the two on the same tensors.  Writes gpurun_out/bench_all.json + a markdown table."""
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


for name in sys.argv[1:] or ["A", "A_native", "B", "C", "E"]:
    scene = synthetic.make_config(name, seed=0)
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
        rec[label + "_fwd"] = timeit(fwd, n)
        if name != "E":
            rec[label + "_fwdbwd"] = timeit(fwdbwd, n)
        o = fwd()
        res[label] = dict(color=o[0].cpu().numpy(), radii=o[1].cpu().numpy(), depth=o[2].cpu().numpy(), alpha=o[3].cpu().numpy())
        rec["visible"] = int((o[1] > 0).sum())
        del rast, ins, m2d
        torch.cuda.empty_cache()
    if "ref" in res:
        d = np.abs(res["sgr"]["color"].astype(np.float64) - res["ref"]["color"])
        rec["rgb_maxabs_vs_ref"] = float(d.max()); rec["rgb_n_gt_1e-4"] = int((d > 1e-4).sum())
        rec["radii_mismatch"] = int((res["sgr"]["radii"] != res["ref"]["radii"]).sum())
        rec["alpha_maxabs_vs_ref"] = float(np.abs(res["sgr"]["alpha"] - res["ref"]["alpha"]).max())
    out[name] = rec
    print(name, json.dumps(rec))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_all.json"), "w"), indent=1)
