"""distCUDA2 timing: this library vs the compiled reference simple-knn (oracle/_ref/ref_knn) at init-time sizes
(VERDICT r1 next #9).  CUDA events, median of 10.  Writes gpurun_out/knn_bench.json."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
import street_gaussians_b200 as sgb

out = {}
rk = util.load_ref_knn() if util.ref_available() else None
for P in (100_000, 1_000_000, 4_000_000):
    g = torch.Generator().manual_seed(P)
    pts = (torch.rand(P, 3, generator=g) * torch.tensor([120.0, 8.0, 200.0])).cuda()
    rec = {}
    for label, fn in (("sgr", sgb.distCUDA2), ("ref", rk.distCUDA2 if rk else None)):
        if fn is None:
            continue
        for _ in range(2):
            r = fn(pts)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = fn(pts); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        rec[label + "_ms"] = float(np.median(ts))
        rec[label + "_sum"] = float(r.double().sum())
    if "ref_ms" in rec:
        rec["bit_identical"] = bool(torch.equal(sgb.distCUDA2(pts), rk.distCUDA2(pts)))
    out[str(P)] = rec
    print(P, json.dumps(rec))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "knn_bench.json"), "w"), indent=1)
