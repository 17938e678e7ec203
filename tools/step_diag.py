"""Where does a bench.py step spend its time? wall-clock + CUDA-event breakdown and a torch.profiler table."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import street_gaussians_b200 as sgb
from street_gaussians_b200 import synthetic
from street_gaussians_b200.sharded import ShardedGaussianRasterizer

scene = synthetic.make_config("C", seed=0)
dev = torch.device("cuda", 0)
cam = scene["cam"]; P = scene["means3D"].shape[0]
rast = ShardedGaussianRasterizer(bench.make_settings(sgb, cam, dev))
params = {k: scene[k].to(dev).requires_grad_(True) for k in bench.PARAM_KEYS}
means2D = torch.zeros((P, 3), device=dev, requires_grad=True)
gc, gd, ga = (scene[k].to(dev) for k in ("grad_color", "grad_depth", "grad_alpha"))

def step(clear=True):
    if clear:
        for v in params.values(): v.grad = None
        means2D.grad = None
    color, radii, depth, alpha, sem = rast(means3D=params["means3D"], means2D=means2D, opacities=params["opacities"], shs=params["shs"],
                                           scales=params["scales"], rotations=params["rotations"])
    torch.autograd.backward([color, depth, alpha], [gc, gd, ga])

def run(n, **kw):
    for _ in range(5): step(**kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): step(**kw)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

print("step (grad=None each step): %.3f ms" % run(30, clear=True))
print("step (accumulate into .grad): %.3f ms" % run(30, clear=False))
print("memory allocated %.1f GB reserved %.1f GB" % (torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9))
s0 = torch.cuda.memory_stats()
run(10)
s1 = torch.cuda.memory_stats()
for k in ("num_alloc_retries", "num_device_alloc", "num_device_free", "allocation.all.allocated", "segment.all.allocated"):
    print(k, s1.get(k, 0) - s0.get(k, 0))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=60))
