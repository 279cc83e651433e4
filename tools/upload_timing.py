"""Upload (vs_index_upload) into a 10M-row index, repeated: seconds per call for a few batch sizes.  Prints one JSON line.
usage: python tools/upload_timing.py [--rows N] [--centroids C]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--centroids", type=int, default=4096)
    a = ap.parse_args()
    import torch
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    pkg._lib.init(0)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ctx = pkg.compute.Context()
    ix, _, _ = bench.build_index(pkg, torch, ctx, a, 0, 1, dev)
    x = bench.gen_unit_rows(torch, bench.SEED_DATA, 7777, 100_000, dev)
    torch.cuda.synchronize()
    um = pkg.compute.EmptyMatrix(100_000, bench.D, ctx=ctx)
    um.FillFloat32Dev(0, x.data_ptr(), 100_000, ctx=ctx)
    ctx.sync()
    new_rows = um.ReadRows()
    out = {}
    for nu in (100_000, 1000, 100_000, 1000, 100_000):
        ids = np.arange(a.rows, a.rows + nu, dtype=np.uint64)
        t0 = time.perf_counter()
        ix2, _ = ix.Upload(new_rows[:nu], ids, ctx=ctx)
        out.setdefault(str(nu), []).append(round(time.perf_counter() - t0, 4))
        del ix2
    print(json.dumps({"workload": f"vs_index_upload into a {a.rows}-row index, host rows, seconds per call in call order", "seconds": out}))


if __name__ == "__main__":
    main()
