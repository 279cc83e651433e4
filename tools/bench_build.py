"""The IVF-Flat centroid build end to end on one GPU (dnc/dnc.go KMeansDivideAndConquer with the database demoted to a
loader): divide and conquer down to CENTROID_SIZE-row leaves, reassign every row to its nearest leaf centroid, re-centre,
group into posting lists, answer a few queries.  Prints one JSON line with the seconds of each phase."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--target", type=int, default=10_000)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--workers", type=int, default=8, help="host threads (each with its own CUDA stream) walking the tree")
    a = ap.parse_args()
    import torch
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    pkg._lib.init(0)
    cp, dnc = pkg.compute, pkg.dnc
    device = torch.device("cuda", 0)
    ctx = cp.default_context()
    t0 = time.time()
    data = cp.EmptyMatrix(a.rows, B.D, ctx=ctx)
    done = 0
    for ci, r0 in enumerate(range(0, a.rows, B.CHUNK)):
        cnt = min(B.CHUNK, a.rows - r0)
        x = B.gen_unit_rows(torch, B.SEED_DATA, ci, cnt, device)
        torch.cuda.synchronize()
        data.FillFloat32Dev(done, x.data_ptr(), cnt, ctx=ctx)
        ctx.sync()
        done += cnt
        del x
    torch.cuda.empty_cache()
    t1 = time.time()
    stats = {}
    cents = dnc.DivideAndConquer(data, target_size=a.target, rng=np.random.default_rng(a.seed), ctx=ctx, workers=a.workers, stats=stats)
    ctx.sync()
    t2 = time.time()
    assign = torch.empty(a.rows, dtype=torch.int32, device=device)
    _, cents2, counts = dnc.ReassignRecenter(data, cents, d_assign=assign.data_ptr(), want_assign=False, ctx=ctx)
    t3 = time.time()
    ids = torch.arange(a.rows, device=device, dtype=torch.int64)
    ix = pkg.ivf.Index.build_dev(data, assign.data_ptr(), ids.data_ptr(), cp.NewMatrix(cents2, ctx=ctx), ctx=ctx)
    ctx.sync()
    t4 = time.time()
    qs = data.ReadRows(0, 8)
    hit_ids, hit_sims, _ = ix.Search(qs, 32, 10, ctx=ctx)
    self_hit = bool((hit_ids[:, 0] == np.arange(8)).all())
    print(json.dumps({
        "workload": f"IVF-Flat centroid build, {a.rows} x {B.D}-d uint8 rows, leaves of at most {a.target} rows",
        "generate_s": round(t1 - t0, 2), "divide_and_conquer_s": round(t2 - t1, 2), "reassign_recenter_s": round(t3 - t2, 2),
        "group_into_lists_s": round(t4 - t3, 2), "total_build_s": round(t4 - t1, 2),
        "centroids": int(cents.shape[0]), "list_rows_min_median_max": [int(counts.min()), int(np.median(counts)), int(counts.max())],
        "workers": a.workers, "host_cores": os.cpu_count(),
        "where_the_time_went": {k: (round(v, 2) if isinstance(v, float) else int(v)) for k, v in sorted(stats.items())},
        "query_rows_find_themselves": self_hit}), flush=True)


if __name__ == "__main__":
    main()
