"""First GPU contact: candidate vs compiled reference vs CPU oracle on small/medium seeded scenes; timing of both CUDA
implementations.  Run under gpurun; prints a report and writes gpurun_out/first_light.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa
import street_gaussians_b200 as sgb  # noqa
from street_gaussians_b200 import synthetic  # noqa

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
report = {}
ref = util.load_ref() if util.ref_available() else None
print("reference available:", ref is not None, "| device:", torch.cuda.get_device_name(0))


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


cases = [("tiny", dict(P=2000, width=160, height=96, sh_degree=3, seed=1, pose=True, scale_med=0.08)),
         ("small", dict(P=20000, width=640, height=400, sh_degree=3, seed=2, pose=True)),
         ("medium", dict(P=200000, width=1280, height=720, sh_degree=3, seed=3, pose=True))]
for name, kw in cases:
    scene = synthetic.make_scene(**kw)
    mine = util.run_api(sgb, scene)
    rep = {}
    if ref is not None:
        r = util.run_api(ref, scene)
        rep["vs_ref"] = util.compare(mine, r, tag=f"[{name}] mine vs REF-CUDA:")
    if kw["P"] <= 20000:
        o = util.run_oracle(scene)
        o.pop("_fw")
        rep["vs_oracle"] = util.compare(mine, o, tag=f"[{name}] mine vs CPU oracle:")
        if ref is not None:
            rep["oracle_vs_ref"] = util.compare(o, r, tag=f"[{name}] oracle vs REF-CUDA:")
    report[name] = rep

# smoke-script replay
scene = synthetic.smoke_script_scene(seed=0)
mine = util.run_api(sgb, scene)
if ref is not None:
    r = util.run_api(ref, scene)
    report["smoke"] = util.compare(mine, r, tag="[smoke script] mine vs REF-CUDA:")

# timing at config B scale
for name, kw in [("B_500k", dict(P=500_000, width=1920, height=1280, sh_degree=3, seed=0)),
                 ("C_1.9M", dict(P=1_500_000, width=1920, height=1280, sh_degree=3, seed=0, n_vehicles=8, per_vehicle=50_000))]:
    scene = synthetic.make_scene(**kw)
    dev = "cuda"
    for label, mod in (("mine", sgb), ("ref", ref)):
        if mod is None:
            continue
        st = util.settings_from(mod, scene["cam"], dev)
        rast = mod.GaussianRasterizer(st)
        ins = {k: scene[k].to(dev).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        m2d = torch.zeros_like(ins["means3D"], requires_grad=True)
        gc, gd, ga = (scene[k].to(dev) for k in ("grad_color", "grad_depth", "grad_alpha"))

        def fwd():
            with torch.no_grad():
                return rast(means3D=ins["means3D"], means2D=m2d, opacities=ins["opacities"], shs=ins["shs"], scales=ins["scales"],
                            rotations=ins["rotations"])

        def fwdbwd():
            for v in ins.values():
                v.grad = None
            c, r_, d, a, s = rast(means3D=ins["means3D"], means2D=m2d, opacities=ins["opacities"], shs=ins["shs"],
                                  scales=ins["scales"], rotations=ins["rotations"])
            torch.autograd.backward([c, d, a], [gc, gd, ga])

        t_f = timeit(fwd)
        t_fb = timeit(fwdbwd)
        vis = int((fwd()[1] > 0).sum())
        print(f"[{name}] {label}: fwd {t_f:.3f} ms, fwd+bwd {t_fb:.3f} ms, visible {vis}")
        report[f"time_{name}_{label}"] = dict(fwd_ms=t_f, fwdbwd_ms=t_fb, visible=vis)
    if ref is not None:
        a = util.run_api(sgb, scene)
        b = util.run_api(ref, scene)
        report[f"parity_{name}"] = util.compare(a, b, tag=f"[{name}] mine vs REF-CUDA:")

json.dump(report, open(os.path.join(ROOT, "gpurun_out", "first_light.json"), "w"), indent=1)
print("done")
