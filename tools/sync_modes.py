"""Distribution of the back-to-back step time under different clock-sampler periods (sync mode comes from SGR_SYNC_MODE)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import street_gaussians_b200 as sgb
from street_gaussians_b200 import synthetic

scene = synthetic.make_config("C", seed=0)
dev = torch.device("cuda", 0)
cam = scene["cam"]; P = scene["means3D"].shape[0]
params = {k: scene[k].to(dev).requires_grad_(True) for k in bench.PARAM_KEYS}
means2D = torch.zeros((P, 3), device=dev, requires_grad=True)
gc, gd, ga = (scene[k].to(dev) for k in ("grad_color", "grad_depth", "grad_alpha"))

def run(rast, n=30):
    def step():
        for v in params.values(): v.grad = None
        c, r, d, a, s = rast(means3D=params["means3D"], means2D=means2D, opacities=params["opacities"], shs=params["shs"],
                             scales=params["scales"], rotations=params["rotations"])
        torch.autograd.backward([c, d, a], [gc, gd, ga])
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for cap_mode in ("exact", "sync-free"):
    rast = sgb.GaussianRasterizer(bench.make_settings(sgb, cam, dev), capacity=sgb.InstanceCapacity() if cap_mode == "sync-free" else None)
    for period in (None, 0.05, 0.25):
        s = bench.ClockSampler(0, period_s=period) if period else None
        ts = [run(rast) for _ in range(5)]
        if s: s.stop()
        print(f"sync={os.environ.get('SGR_SYNC_MODE', 'block'):5s} binning={cap_mode:9s} sampler={'off' if not period else f'{int(period*1000)}ms':5s}: " + " ".join(f"{t:.3f}" for t in ts), flush=True)
