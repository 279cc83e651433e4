"""Throughput of the single-process multi-device path (vs_sharded_*, csrc/sharded.cu): one host process, G devices,
T host threads each holding its own vs_sharded_ctx (one goroutine + closure each, server/search.go:230) and making
synchronous host-buffer search calls.  Also checks the hits against a single-device index.
usage: python tools/sharded_single_process.py [--rows N] [--devices 0,1] [--threads 4] [--out profiles/xxx.json]"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4_000_000)
    ap.add_argument("--centroids", type=int, default=4096)
    ap.add_argument("--nprobe", type=int, default=32)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    pkg._lib.init(0)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    cp = pkg.compute
    ctx = cp.Context()
    D = bench.D
    # rows and their nearest centroid, generated and assigned on device 0, read back to host buffers (what a loader has)
    x = bench.gen_unit_rows(torch, bench.SEED_CENT, 0, a.centroids, dev)
    torch.cuda.synchronize()
    cent = cp.EmptyMatrix(a.centroids, D, ctx=ctx)
    cent.FillFloat32Dev(0, x.data_ptr(), a.centroids, ctx=ctx)
    ctx.sync()
    cent_rows = cent.ReadRows()
    rows = np.empty((a.rows, bench.ROW_BYTES), np.uint8)
    lists = np.empty(a.rows, np.uint32)
    for ci, r0 in enumerate(range(0, a.rows, bench.CHUNK)):
        cnt = min(bench.CHUNK, a.rows - r0)
        xs = bench.gen_unit_rows(torch, bench.SEED_DATA, ci, cnt, dev)
        torch.cuda.synchronize()
        m = cp.EmptyMatrix(cnt, D, ctx=ctx)
        m.FillFloat32Dev(0, xs.data_ptr(), cnt, ctx=ctx)
        ctx.sync()
        asg = torch.empty(cnt, device=dev, dtype=torch.int32)
        cent.ArgmaxDev(m, asg.data_ptr(), ctx=ctx)
        ctx.sync()
        rows[r0:r0 + cnt] = m.ReadRows()
        lists[r0:r0 + cnt] = asg.cpu().numpy().astype(np.uint32)
        del xs, m, asg
    nsteps = a.steps + 4
    qs = []
    for s in range(nsteps * a.threads):
        xq = bench.gen_unit_rows(torch, bench.SEED_QUERY, s, a.batch, dev)
        torch.cuda.synchronize()
        qm = cp.EmptyMatrix(a.batch, D, ctx=ctx)
        qm.FillFloat32Dev(0, xq.data_ptr(), a.batch, ctx=ctx)
        ctx.sync()
        qs.append(qm.ReadRows())
        del qm, xq
    out = {"rows": a.rows, "centroids": a.centroids, "nprobe": a.nprobe, "k": a.k, "batch": a.batch, "threads": a.threads, "runs": []}
    want = None
    for devs in [[0]] + ([[int(v) for v in a.devices.split(",")]] if a.devices != "0" else []):
        t0 = time.time()
        sh = pkg.ivf.ShardedIndex(devs).build_assigned(rows, None, lists, cent_rows)
        build_s = time.time() - t0
        got = sh.Search(qs[0], a.nprobe, a.k)
        if want is None:
            want = got
        same = bool((got[0] == want[0]).all() and (got[1].view(np.uint32) == want[1].view(np.uint32)).all() and (got[2] == want[2]).all())
        sctx = [sh.NewSearchContext() for _ in range(a.threads)]
        for t in range(a.threads):       # warm up every context (scratch growth, first launches)
            for s in range(2):
                sh.Search(qs[t * nsteps + s], a.nprobe, a.k, sctx=sctx[t])
        barrier = threading.Barrier(a.threads + 1)

        def worker(t):
            barrier.wait()
            for s in range(4, 4 + a.steps):
                sh.Search(qs[t * nsteps + s], a.nprobe, a.k, sctx=sctx[t])
            barrier.wait()

        th = [threading.Thread(target=worker, args=(t,)) for t in range(a.threads)]
        for t_ in th:
            t_.start()
        barrier.wait()
        t0 = time.perf_counter()
        barrier.wait()
        dt = time.perf_counter() - t0
        for t_ in th:
            t_.join()
        # one caller, one call at a time (latency of a batch through the handle's own context)
        t0 = time.perf_counter()
        for s in range(4, 4 + a.steps):
            sh.Search(qs[s], a.nprobe, a.k)
        dt1 = time.perf_counter() - t0
        out["runs"].append({"devices": devs, "queries_per_s": round(a.batch * a.steps * a.threads / dt, 1),
                            "queries_per_s_one_caller": round(a.batch * a.steps / dt1, 1), "ms_per_call_one_caller": round(dt1 / a.steps * 1e3, 3),
                            "build_s": round(build_s, 2), "same_hits_as_one_device": same})
        for c_ in sctx:
            sh.CloseSearchContext(c_)
        sh.close()
    print(json.dumps(out, indent=1))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
