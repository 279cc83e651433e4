"""One query against every row of the store (nprobe = all lists): the single-query streaming scan at a size where its
bandwidth shows.  Prints one JSON line; short enough to run under `ncu --set full -k regex:stage_kernel`.
usage: python tools/single_query_scan.py [--rows N] [--centroids C] [--reps R]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--centroids", type=int, default=4096)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--reps", type=int, default=8)
    a = ap.parse_args()
    import torch
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    pkg._lib.init(0)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ctx = pkg.compute.Context()
    ix, _, _ = bench.build_index(pkg, torch, ctx, a, 0, 1, dev)
    k = a.k
    d_ids = torch.zeros((1, k), device=dev, dtype=torch.int64)
    d_sims = torch.zeros((1, k), device=dev, dtype=torch.float32)
    d_counts = torch.zeros(1, device=dev, dtype=torch.int32)
    d_status = torch.zeros(1, device=dev, dtype=torch.int32)
    ms = []
    for rep in range(a.reps):
        x = bench.gen_unit_rows(torch, bench.SEED_QUERY, 2000 + rep, 1, dev)
        torch.cuda.synchronize()
        q = pkg.compute.EmptyMatrix(1, bench.D, ctx=ctx)
        q.FillFloat32Dev(0, x.data_ptr(), 1, ctx=ctx)
        ctx.sync()
        ctx.timer_start()
        ix.SearchDev(q, a.centroids, k, d_ids.data_ptr(), d_sims.data_ptr(), d_counts.data_ptr(), d_status.data_ptr(), ctx=ctx)
        ms.append(ctx.timer_stop())
    best = sorted(ms[2:])[len(ms[2:]) // 2]
    print(json.dumps({"workload": f"1 query x {a.rows} rows x {bench.D}-d uint8, top-{k}, flat scan through the index",
                      "ms_per_query_p50": round(best, 4), "ms_all": [round(v, 4) for v in ms],
                      "algorithmic_gbs": round(a.rows * bench.ROW_BYTES / (best * 1e-3) / 1e9, 1)}))


if __name__ == "__main__":
    main()
