#!/usr/bin/env bash
# compute-sanitizer (memcheck + racecheck + synccheck) over one small forward+backward of every kernel family.
# Run under gpurun; logs land in gpurun_out/sanitizer_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, util
import street_gaussians_b200 as sgb
from street_gaussians_b200 import synthetic
from street_gaussians_b200.sharded import cyclic_band
for kw, extra in ((dict(P=4000, width=208, height=120, sh_degree=3, seed=21, pose=True, scale_med=0.06), {}),
                  (dict(P=1500, width=128, height=80, sh_degree=2, seed=25, pose=True, scale_med=0.3, semantics=5), {}),
                  (dict(P=3000, width=160, height=112, sh_degree=1, seed=26, pose=True, scale_med=0.06), dict(band=cyclic_band(112, 1, 3)))):
    sc = synthetic.make_scene(**kw)
    r = util.run_api(sgb, sc, rasterizer_kwargs=extra)
    print("ok", kw["P"], float(abs(r["color"]).sum()), float(abs(r["g_means3D"]).sum()))
print(float(sgb.distCUDA2(torch.rand(3000, 3, device="cuda")).sum()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$? =="; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok |Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
