"""Per-iteration wall and device times of vs_kmeans at the reference's own shape (50 000 sampled rows, k = 5, superset 25)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_pkg
from _util import unit_rows
pkg = load_pkg(); pkg._lib.init(0)
cp = pkg.compute
n, d, k = 50000, 768, 5
data = cp.NewMatrix(cp.QuantizeMatrixFloat32(unit_rows(n, d, 1)))
rows = np.random.default_rng(1).choice(n, 5 * k, replace=False)
ctx = cp.Context()
for lim in (50, 200):
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    _, st = pkg.dnc.KMeans(data, k, superset_rows=rows, iter_limit=lim, ctx=ctx, want_stats=True)
    dt = time.perf_counter() - t0
    its = st["superset_iterations"] + st["set_iterations"]
    print(f"limit {lim}: {its} iterations in {dt*1e3:.1f} ms wall = {dt/its*1e6:.0f} us/iter; events: assign {st['assign_us']/its:.0f} us/iter, update {st['update_us']/its:.0f} us/iter; launches/iter {(ctx.launch_count()-l0)/its:.1f}")
