"""EXPERIMENT, written at the end of round 1 WITHOUT GPU time left to run it: can one Gaussian-sharded fwd+bwd step of
the peer-memory path be captured in a CUDA graph, and what does a replayed step cost?

Motivation (profiles/r01_summary.md §5): at N = 8 the step takes 1.00 ms while the kernels on one rank's
critical path add up to ~0.52 ms — the Python host loop (~30 launches, 0.6 ms) and the barriers waiting for the slowest
rank's host are the difference.  The p2p path has no NCCL call and, in bounded mode, no host read-back, so the whole step
is graph-capturable in principle.  This script drives the staged C-ABI calls directly (no autograd, no InstanceCapacity:
their pinned-memory allocation and event bookkeeping are not capture-safe yet) so nothing in the product has to change
to answer the question.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 \
        tools/graph_step_probe.py [--workload C] [--steps 50]

Prints, on rank 0: eager ms/step, graph-replay ms/step (max over ranks), and where capture failed if it did.
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import street_gaussians_b200 as sgb  # noqa: E402
from street_gaussians_b200 import _capi, sharded as SH, synthetic  # noqa: E402
from street_gaussians_b200.rasterizer import _make_frame, _ptr, _stream  # noqa: E402

KEYS = ("means3D", "shs", "opacities", "scales", "rotations")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C")
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _capi.lib()

    scene = synthetic.make_config(args.workload, seed=0)
    cam = scene["cam"]
    H, W, P = cam["image_height"], cam["image_width"], scene["means3D"].shape[0]
    st = sgb.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
                                           bg=cam["bg"].to(dev), scale_modifier=cam["scale_modifier"], viewmatrix=cam["viewmatrix"].to(dev),
                                           projmatrix=cam["projmatrix"].to(dev), sh_degree=cam["sh_degree"], campos=cam["campos"].to(dev),
                                           prefiltered=False, debug=False)
    chunk = (P + world - 1) // world
    lo, hi = min(P, rank * chunk), min(P, (rank + 1) * chunk)
    P_total = chunk * world
    band = SH.cyclic_band(H, rank, world) if world > 1 else None
    rows = SH.band_of_rows(H, rank, world).to(dev).view(1, H, 1).float()
    gc, gd, ga = (scene[k].to(dev) * rows for k in ("grad_color", "grad_depth", "grad_alpha"))
    lt = SH._local_tensors(*(scene[k][lo:hi].to(dev) for k in ("means3D", "shs")), None, None,
                           *(scene[k][lo:hi].to(dev) for k in ("opacities", "scales", "rotations")), None)
    ws = SH.PeerWorkspace(st, chunk, world, rank, dev) if world > 1 else SH.PeerWorkspace.emulate(st, chunk, 1, dev)[0]

    # one exact pass to learn the instance count of this rank's band
    rec, radii_l = SH.project_records(lt, st, chunk)
    SH.scatter_records(st, ws, rec, radii_l, hi - lo)
    ws.barrier()
    fs = SH.peer_forward_state(ws)
    SH.forward_records(st, band, fs, (ws.geom_bytes, ws.img_bytes), ws.radii_all, None)
    cap = int(fs.num_instances * 1.25) + 4096
    nbytes = int(L.sgr_binning_bytes(cap))
    torch.cuda.synchronize()

    # static buffers of the captured step
    f32 = dict(device=dev, dtype=torch.float32)
    color, depth, alpha = torch.zeros((3, H, W), **f32), torch.zeros((1, H, W), **f32), torch.zeros((1, H, W), **f32)
    binning = torch.empty((nbytes,), device=dev, dtype=torch.uint8)
    fs.binning, fs.num_instances = binning, cap
    fr_total, keep = _make_frame(st, P_total, 0, 0, dev, band)

    def step():
        rec, radii_l = SH.project_records(lt, st, chunk)
        SH.scatter_records(st, ws, rec, radii_l, hi - lo)
        ws.barrier()
        rc = L.sgr_forward_records(C.byref(fr_total), _ptr(ws.radii_all), None, _ptr(color), _ptr(depth), _ptr(alpha), None, _ptr(fs.geom),
                                   ws.geom_bytes, _ptr(fs.img), ws.img_bytes, _capi.ALLOC_FN(), None, None, None, _ptr(binning), nbytes, cap,
                                   _stream(dev))
        _capi.check(rc, "sgr_forward_records")
        SH.backward_blend_records(st, band, fs, P_total, None, alpha, gc, gd, ga, None, grad2d_out=ws.grad2d)
        ws.barrier()
        g2 = SH.gather_grad2d(st, ws, rec, radii_l, hi - lo)
        return SH.backward_geom_local(st, lt, rec, radii_l, g2)

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / n], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        for _ in range(5):
            step()
        eager_ms = timed(step, args.steps)
        graph_ms, err = None, None
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):  # torch's documented warm-up on a side stream before capture
                for _ in range(3):
                    step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                grads = step()
            for _ in range(5):
                g.replay()
            graph_ms = timed(g.replay, args.steps)
            del grads
        except Exception as e:  # noqa: BLE001
            err = repr(e)
    if rank == 0:
        print(json.dumps(dict(experiment="cuda-graph replay of the Gaussian-sharded p2p step", n_gpus=world, workload=args.workload,
                              eager_ms_per_step=eager_ms, graph_ms_per_step=graph_ms, capture_error=err, capacity=cap)))
    del keep
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
