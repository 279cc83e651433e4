"""Debug aid: one query through the fused kernel without the resolve step; prints status / counts / first ids."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from __graft_entry__ import load_pkg
from _util import unit_rows
import oracle
pkg = load_pkg(); pkg._lib.init(0)
cp = pkg.compute
n, d, C = 30000, 768, 96
rows = oracle.quantize_matrix_f32(unit_rows(n, d, 5)); cent = oracle.quantize_matrix_f32(unit_rows(C, d, 6))
_, lists = oracle.argmax_MxN(cent, rows)
for with_ids in (True, False):
    doc = (np.arange(n, dtype=np.uint64) + 1000) if with_ids else None
    ix = pkg.ivf.Index.build_assigned(rows, doc, lists.astype(np.uint32), cent)
    qs = oracle.quantize_matrix_f32(unit_rows(4, d, 99))
    dev = torch.device("cuda", 0)
    k = 10
    d_ids = torch.zeros((1, k), device=dev, dtype=torch.int64); d_sims = torch.zeros((1, k), device=dev, dtype=torch.float32)
    d_counts = torch.zeros(1, device=dev, dtype=torch.int32); d_status = torch.full((1,), 77, device=dev, dtype=torch.int32)
    ctx = cp.default_context()
    for nprobe in (8, C):
        for i in range(2):
            one = cp.NewMatrix(qs[i:i + 1])
            for fused in (True, False):
                cp.debug_set_fused(fused)
                d_status.fill_(77); torch.cuda.synchronize()
                ix.SearchDev(one, nprobe, k, d_ids.data_ptr(), d_sims.data_ptr(), d_counts.data_ptr(), d_status.data_ptr(), ctx=ctx)
                ctx.sync()
                print(f"ids={with_ids} nprobe={nprobe} q={i} fused={fused}: status={d_status.item()} count={d_counts.item()} ids={d_ids.cpu().numpy()[0,:4].tolist()} sims={d_sims.cpu().numpy()[0,:2].tolist()}")
            cp.debug_set_fused(True)
