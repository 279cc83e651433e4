"""Debug aid: run the tcgen05 batch path on a small problem and compare with the scan path."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_pkg
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _util import unit_rows, noop_rows

vs = load_pkg()
vs._lib.init(0)
n, d, nq, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
x = unit_rows(n, d, 1)
rows = vs.compute.QuantizeMatrixFloat32(x)
qs = vs.compute.QuantizeMatrixFloat32(unit_rows(nq, d, 2))
m = vs.compute.NewMatrix(rows)
t = time.time()
ids, sims, counts = vs.ivf.SearchFlatBatch(m, qs, k)
print("gemm path %.3fs" % (time.time() - t), flush=True)
ids2, sims2, counts2 = vs.ivf.SearchFlat(m, qs, k)
bad = [i for i in range(nq) if counts[i] != counts2[i] or (ids[i] != ids2[i]).any() or (sims[i].view(np.uint32) != sims2[i].view(np.uint32)).any()]
print("mismatching queries:", len(bad), bad[:10])
if bad:
    i = bad[0]
    print(counts[i], counts2[i]); print(ids[i]); print(ids2[i]); print(sims[i]); print(sims2[i])
import torch
qm = vs.compute.NewMatrix(qs)
d_ids = torch.zeros((nq, k), dtype=torch.int64, device="cuda"); d_sims = torch.zeros((nq, k), dtype=torch.float32, device="cuda"); d_counts = torch.zeros(nq, dtype=torch.int32, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    st = vs.ivf.SearchBatchDev(m, qm, k, d_ids.data_ptr(), d_sims.data_ptr(), d_counts.data_ptr())
    vs.compute.default_context().sync(); dt = time.time() - t
    print("dev path %.3f ms stats(cand, fallback, tiles, sample)=%s  %.1f TOPS" % (dt * 1e3, st, 2.0 * n * nq * d / dt / 1e12), flush=True)
sys.exit(1 if bad else 0)
