"""BASELINE config 4: dnc k-means centroid build (assign + update) over n x 768-d uint8 rows with k centroids
(dnc/k_means.go:67-117), every iteration device-resident (vs_kmeans).  Runs `--iters` iterations of each phase with
ks = k (the set-phase shape) and prints one JSON line: seconds per iteration split into assign (tensor cores:
2*n*k*768 integer ops) and update (HBM: 776 B per row), achieved TOPS / GB/s."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=20_000_000)
    ap.add_argument("--k", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=1)
    a = ap.parse_args()
    import torch
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    pkg._lib.init(0)
    cp = pkg.compute
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
    build_s = time.time() - t0
    rows = np.random.default_rng(B.SEED_CENT).choice(a.rows, a.k, replace=False)
    sampler = B.ClockSampler(0)
    sampler.start()
    t0 = time.time()
    cent, st = pkg.dnc.KMeans(data, a.k, superset_rows=rows, iter_limit=a.iters, ctx=ctx, want_stats=True) \
        if a.k * 5 <= 0 else _kmeans_ks_equals_k(pkg, data, a.k, rows, a.iters, ctx)
    wall = time.time() - t0
    clocks = sampler.stop()
    its = st["superset_iterations"] + st["set_iterations"]
    assign_s = st["assign_us"] / 1e6 / its
    update_s = st["update_us"] / 1e6 / its
    print(json.dumps({
        "workload": f"k-means iteration (assign + update), {a.rows} x {B.D}-d uint8 rows, {a.k} centroids",
        "iterations_timed": its, "s_per_iteration": round(assign_s + update_s, 4),
        "assign": {"s": round(assign_s, 4), "int_ops": 2.0 * a.rows * a.k * B.D,
                   "achieved_tops": round(2.0 * a.rows * a.k * B.D / assign_s / 1e12, 1)},
        "update": {"s": round(update_s, 4), "bytes": a.rows * B.ROW_BYTES,
                   "achieved_gbs": round(a.rows * B.ROW_BYTES / update_s / 1e9, 1)},
        "wall_s": round(wall, 2), "build_s": round(build_s, 2), "clocks": clocks,
        "centroid_checksum": int(np.frombuffer(cent.tobytes(), np.uint8).astype(np.uint64).sum())}), flush=True)


def _kmeans_ks_equals_k(pkg, data, k, rows, iters, ctx):
    """vs_kmeans with a superset of exactly k rows: both phases run the k-centroid shape."""
    import ctypes as C
    from go_vectorsearch_b200.compute import _check, _p
    out = np.empty((k, 8 + data.cols), np.uint8)
    stats = np.zeros(4, np.int64)
    r = np.ascontiguousarray(rows, dtype=np.uint64)
    _check(data._L.vs_kmeans(ctx.handle, data.handle, int(k), _p(r), int(k), int(iters), _p(out), _p(stats)))
    return out, {"superset_iterations": int(stats[0]), "set_iterations": int(stats[1]),
                 "assign_us": int(stats[2]), "update_us": int(stats[3])}


if __name__ == "__main__":
    main()
