import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_pkg
from _util import unit_rows
pkg = load_pkg(); pkg._lib.init(0)
cp = pkg.compute
d, C, n = 768, 4096, 20000
cent = cp.QuantizeMatrixFloat32(unit_rows(C, d, 1))
rows = cp.QuantizeMatrixFloat32(unit_rows(n, d, 2))
lists = (np.arange(n) % C).astype(np.uint32)
ix = pkg.ivf.Index.build_assigned(rows, np.arange(n, dtype=np.uint64), lists, cent)
qs = cp.QuantizeMatrixFloat32(unit_rows(640, d, 3))
ctx = cp.Context()
s0 = ctx.slowpath_count()
for i in range(0, 640, 64): ix.SelectProbes(qs[i:i+64], 32, ctx=ctx)
s1 = ctx.slowpath_count()
for i in range(0, 640, 4): ix.SelectProbes(qs[i:i+4], 32, ctx=ctx)
s2 = ctx.slowpath_count()
print("batched path rescored", s1 - s0, "streaming path rescored", s2 - s1, "per 640 queries")
