"""BASELINE config 4 (assign half): nearest centroid of n x 768-d uint8 rows among m centroids (compute/cosine.go:70-125 as
called from dnc/k_means.go:75) through the tensor-core path.  Prints one JSON line: rows/s, integer TOPS of the whole
call (2*n*m*768 ops / CUDA-event time), rows that took the literal path, and parity with the dp4a scan form on a sample."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4_000_000)
    ap.add_argument("--centroids", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--check", type=int, default=4096, help="rows re-assigned by the scan form for parity")
    a = ap.parse_args()
    import torch
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    pkg._lib.init(0)
    cp = pkg.compute
    device = torch.device("cuda", 0)
    ctx = cp.default_context()

    def fill(n, seed):
        m = cp.EmptyMatrix(n, B.D, ctx=ctx)
        done = 0
        for ci, r0 in enumerate(range(0, n, B.CHUNK)):
            cnt = min(B.CHUNK, n - r0)
            x = B.gen_unit_rows(torch, seed, ci, cnt, device)
            torch.cuda.synchronize()
            m.FillFloat32Dev(done, x.data_ptr(), cnt, ctx=ctx)
            ctx.sync()
            done += cnt
            del x
        return m

    data = fill(a.rows, B.SEED_DATA)
    cent = fill(a.centroids, B.SEED_CENT)
    assign = torch.empty(a.rows, device=device, dtype=torch.int32)
    cent.ArgmaxDev(data, assign.data_ptr(), ctx=ctx)   # warm-up (allocates the scratch)
    ctx.sync()
    sampler = B.ClockSampler(0)
    sampler.start()
    slow0 = ctx.slowpath_count()
    times = []
    for _ in range(a.reps):
        ctx.timer_start()
        cent.ArgmaxDev(data, assign.data_ptr(), ctx=ctx)
        times.append(ctx.timer_stop())
    clocks = sampler.stop()
    slow = (ctx.slowpath_count() - slow0) // a.reps
    ms = min(times)
    ops = 2.0 * a.rows * a.centroids * B.D
    # parity of a sample against the scan form
    nchk = min(a.check, a.rows)
    sample = cp.NewMatrix(data.ReadRows(0, nchk))
    cp.debug_set_argmax_gemm_min(1 << 30)
    _, want = cent.MatrixCosineSimilarity(sample, ctx=ctx, want_sims=False)
    cp.debug_set_argmax_gemm_min(256)
    got = assign[:nchk].cpu().numpy()
    print(json.dumps({
        "workload": f"nearest centroid, {a.rows} x {B.D}-d uint8 rows, {a.centroids} centroids",
        "rows_per_s": round(a.rows / (ms * 1e-3), 1), "ms": round(ms, 3), "ms_all": [round(t, 3) for t in times],
        "int_ops": ops, "achieved_tops_whole_call": round(ops / (ms * 1e-3) / 1e12, 1),
        "rows_literal_path": int(slow), "parity_vs_scan_form": {"rows_checked": nchk, "match": bool((got == want).all())},
        "clocks": clocks}), flush=True)


if __name__ == "__main__":
    main()
