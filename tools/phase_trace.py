"""Phase timing of the two search stages for one launch (profiling aid; writes a text summary).
usage: python tools/phase_trace.py [--rows N] [--batch B] [--out profiles/xxx.txt]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--centroids", type=int, default=4096)
    ap.add_argument("--nprobe", type=int, default=32)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    pkg._lib.init(0)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ctx = pkg.compute.Context()
    ix, cent, _ = bench.build_index(pkg, torch, ctx, a, 0, 1, dev)
    B, k = a.batch, a.k
    d_ids = torch.zeros((B, k), device=dev, dtype=torch.int64)
    d_sims = torch.zeros((B, k), device=dev, dtype=torch.float32)
    d_counts = torch.zeros(B, device=dev, dtype=torch.int32)
    d_status = torch.zeros(B, device=dev, dtype=torch.int32)
    lines = []
    for rep in range(6):
        x = bench.gen_unit_rows(torch, bench.SEED_QUERY, 1000 + rep, B, dev)
        torch.cuda.synchronize()
        q = pkg.compute.EmptyMatrix(B, bench.D, ctx=ctx)
        q.FillFloat32Dev(0, x.data_ptr(), B, ctx=ctx)
        ctx.sync()
        ctx.trace_enable(rep >= 3)
        ctx.timer_start()
        ix.SearchDev(q, a.nprobe, k, d_ids.data_ptr(), d_sims.data_ptr(), d_counts.data_ptr(), d_status.data_ptr(), ctx=ctx)
        ms = ctx.timer_stop()
        if rep < 3:
            continue
        lines.append(f"rep {rep}: search (both stages) {ms * 1e3:.1f} us by CUDA events (trace on)")
        for stage in (1, 2):
            t = ctx.trace_read(stage).astype(np.int64)
            act = t[t[:, 0] > 0]
            if act.shape[0] == 0:
                continue
            t0 = act[:, 0].min()
            def rel(col, f):
                v = act[:, col]
                v = v[v > 0]
                return f(v - t0) / 1e3 if v.size else float("nan")
            lines.append(f"  stage {stage} detail: sorted {rel(8, np.max):.1f} | written {rel(9, np.max):.1f} | fenced {rel(10, np.max):.1f} | ticket {rel(3, np.max):.1f}"
                         f" || gathered {rel(11, np.max):.1f} | sorted {rel(12, np.max):.1f} | walked {rel(13, np.max):.1f} | merged {rel(4, np.max):.1f}")
            lines.append(f"  stage {stage}: blocks={act.shape[0]}  start spread {rel(0, np.max):.1f} us | prologue done max {rel(1, np.max):.1f}"
                         f" | scan done median {rel(2, np.median):.1f} max {rel(2, np.max):.1f} | partial published max {rel(3, np.max):.1f}"
                         f" | last: slots merged {rel(4, np.max):.1f} | final list {rel(5, np.max):.1f} | emitted {max(rel(6, np.max), rel(7, np.max) if (act[:,7]>0).any() else 0):.1f} us")
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(f"# phase trace, rows={a.rows} centroids={a.centroids} nprobe={a.nprobe} k={k} batch={B}\n" + txt + "\n")


if __name__ == "__main__":
    main()
