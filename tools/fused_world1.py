"""The fused Gaussian-sharded step (sgr_sharded_forward / sgr_sharded_backward) at world = 1 on config C — the single-process
stand-in used to put the multi-GPU-only kernels (projection + record scatter, compact-first tile count, chain rule + grad2d gather,
peer barrier) under `ncu`, which cannot wrap a multi-rank job.  All records are self-deliveries here, so the NVLink share of their
time is NOT in these captures (the N = 2 / N = 8 timelines under profiles/ carry that).

  ncu --set full -k regex:"preprocess_fwd_kernel|count_compact_kernel|preprocess_bwd_tma_kernel" -s 8 -c 3 \
      python tools/fused_world1.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, util
import street_gaussians_b200 as sgb
from street_gaussians_b200 import synthetic
from street_gaussians_b200 import sharded as SH

dev = torch.device("cuda")
sc = synthetic.make_config(os.environ.get("SGR_WORKLOAD", "C"))
P = sc["means3D"].shape[0]
st = util.settings_from(sgb, sc["cam"], dev)
lt = SH._local_tensors(sc["means3D"].to(dev), sc["shs"].to(dev), None, None, sc["opacities"].to(dev), sc["scales"].to(dev), sc["rotations"].to(dev), None)
up = [sc[k].to(dev) for k in ("grad_color", "grad_depth", "grad_alpha")]
ws = SH.PeerWorkspace.emulate(st, P, 1, dev)[0]
steps = int(os.environ.get("SGR_STEPS", "4"))
with torch.no_grad():
    for i in range(steps):
        gcap = -1 if i == 0 else int(1.02 * P)  # first step uncompacted, then the compacted depth order
        col, dep, alp, _ = SH.sharded_forward_raw(st, None, ws, lt, P, 16_000_000, gcap)
        g = SH.sharded_backward_raw(st, None, ws, lt, P, 16_000_000, alp, *up)
torch.cuda.synchronize()
print("fused world=1 ok", float(col.abs().sum()), float(g[0].abs().sum()))
