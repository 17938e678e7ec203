"""Host-side cost of one fwd+bwd step: wall-clock until each call RETURNS (launch cost) vs until the GPU is done."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
import street_gaussians_b200 as sgb
from street_gaussians_b200 import synthetic, rasterizer as R

scene = synthetic.make_config(sys.argv[1] if len(sys.argv) > 1 else "C", seed=0)
dev = "cuda"
st = util.settings_from(sgb, scene["cam"], dev)
t = {k: scene[k].to(dev) for k in ("means3D", "shs", "opacities", "scales", "rotations", "grad_color", "grad_depth", "grad_alpha")}
rast = sgb.GaussianRasterizer(st)
ins = {k: t[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
m2d = torch.zeros_like(ins["means3D"], requires_grad=True)


def wall(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ret, tot = [], []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        ret.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
    return float(np.median(ret)), float(np.median(tot))


state = {}
def f_fwd():
    with torch.no_grad():
        state["o"] = R._forward_impl(t["means3D"], t["shs"], None, None, t["opacities"], t["scales"], t["rotations"], None, st, None)
def f_bb():
    col, rad, dep, alp, sem, fst, tens = state["o"]
    state["g"] = R._backward_blend_impl(st, None, fst, tens, alp, t["grad_color"], t["grad_depth"], t["grad_alpha"], None)
def f_bg():
    col, rad, dep, alp, sem, fst, tens = state["o"]
    R._backward_geom_impl(st, None, fst, tens, rad, state["g"][0])
def f_api_fwd():
    return rast(means3D=ins["means3D"], means2D=m2d, opacities=ins["opacities"], shs=ins["shs"], scales=ins["scales"], rotations=ins["rotations"])
def f_api_fwdbwd():
    for v in ins.values(): v.grad = None
    c, r_, d, a, s = f_api_fwd()
    torch.autograd.backward([c, d, a], [t["grad_color"], t["grad_depth"], t["grad_alpha"]])

for name, fn in (("_forward_impl", f_fwd), ("_backward_blend_impl", f_bb), ("_backward_geom_impl", f_bg), ("api forward (autograd)", f_api_fwd),
                 ("api fwd+bwd", f_api_fwdbwd)):
    r, tt = wall(fn)
    print(f"{name:28s} returns after {r:7.3f} ms, GPU done after {tt:7.3f} ms")
# back-to-back steps (what bench.py 'value' measures)
for _ in range(3): f_api_fwdbwd()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): f_api_fwdbwd()
torch.cuda.synchronize(); print("back-to-back fwd+bwd: %.3f ms/step" % ((time.perf_counter() - t0) / 20 * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20): f_api_fwdbwd()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
