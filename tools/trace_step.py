"""Kernel timeline of the Gaussian-sharded fused step (bench.py's N > 1 timed region) from torch.profiler (CUPTI), one file per rank:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 tools/trace_step.py
Writes gpurun_out/trace_n{N}_rank{r}.txt: for the last profiled step every GPU kernel with its start offset, duration and the idle
gap in front of it, plus host-side totals — the evidence for where a multi-GPU step spends its time (ncu cannot wrap a multi-rank job)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import street_gaussians_b200 as sgb  # noqa: E402
from street_gaussians_b200 import synthetic  # noqa: E402
from street_gaussians_b200.sharded import GaussianShardedRasterizer, band_of_rows  # noqa: E402


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    scene = synthetic.make_config("C", seed=0)
    cam = scene["cam"]
    P, H, W = scene["means3D"].shape[0], cam["image_height"], cam["image_width"]
    chunk = (P + world - 1) // world
    lo, hi = min(P, rank * chunk), min(P, (rank + 1) * chunk)
    cap = sgb.InstanceCapacity()
    if world > 1:
        rast = GaussianShardedRasterizer(bench.make_settings(sgb, cam, dev), capacity=cap, chunk=chunk, exchange="p2p")
    else:
        rast = sgb.GaussianRasterizer(bench.make_settings(sgb, cam, dev), capacity=cap)
    params = {k: scene[k][lo:hi].to(dev).requires_grad_(True) for k in bench.PARAM_KEYS}
    m2d = torch.zeros((hi - lo, 3), device=dev, requires_grad=True)
    m = band_of_rows(H, rank, world).to(dev).view(1, H, 1).float()
    gc, gd, ga = (scene[k].to(dev) * m for k in ("grad_color", "grad_depth", "grad_alpha"))

    def step():
        for v in params.values():
            v.grad = None
        m2d.grad = None
        c, r, d, a, s = rast(means3D=params["means3D"], means2D=m2d, opacities=params["opacities"], shs=params["shs"], scales=params["scales"],
                             rotations=params["rotations"])
        torch.autograd.backward([c, d, a], [gc, gd, ga])

    for _ in range(8):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # split into steps at the first kernel of each step (the projection kernel)
    starts = [i for i, e in enumerate(evs) if "preprocess_fwd_kernel" in e.name]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", f"trace_n{world}_rank{rank}.txt")
    with open(path, "w") as f:
        if len(starts) >= 2:
            a, b = starts[-2], starts[-1]
            seg = evs[a:b]
            t0 = seg[0].time_range.start
            f.write(f"# N={world} rank {rank}: one step = {evs[b].time_range.start - t0:.1f} us (start of projection to start of the next)\n")
            f.write("# start_us  dur_us  gap_before_us  kernel\n")
            prev_end, busy = t0, 0.0
            for e in seg:
                s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
                f.write(f"{s:9.1f} {d:8.1f} {e.time_range.start - prev_end:8.1f}  {e.name[:110]}\n")
                prev_end = max(prev_end, e.time_range.end)
                busy += d
            f.write(f"# sum of kernel durations {busy:.1f} us, idle {evs[b].time_range.start - t0 - busy:.1f} us\n")
        cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith(("cudaLaunchKernel", "cuLaunchKernel"))]
        f.write(f"# host: {len(cpu)} kernel launches over 4 steps\n")
    if rank == 0:
        print(open(path).read())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
