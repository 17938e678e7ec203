"""How many host cores does this box really give us? (affinity mask, cgroup quota, OpenMP scaling of the oracle)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
print("os.cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try:
        print(p, open(p).read().strip())
    except Exception as e:
        print(p, "n/a")
import util
from oracle import oracle as O
from street_gaussians_b200 import synthetic
sc = synthetic.make_scene(P=60000, width=960, height=640, sh_degree=3, seed=1)
for n in (1, 2, 4, 8, 16, 32, 64):
    if n > os.cpu_count():
        break
    O.set_num_threads(n)
    t = time.perf_counter()
    r = util.run_oracle(sc, backward=True); r.pop("_fw")
    print(f"oracle fwd+bwd with {n:2d} threads: {time.perf_counter() - t:.2f} s")
