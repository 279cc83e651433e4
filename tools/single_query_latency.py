"""Single-query latency of the fused one-launch search (csrc/fused.cu) beside the two-launch path, device-timed
(CUDA events on the context's stream), with the kernel's own phase stamps.  BASELINE configs 1 and 2.
usage: python tools/single_query_latency.py [--rows N] [--centroids C] [--nprobe P] [--out profiles/xxx.json]"""
import argparse
import json
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
    ap.add_argument("--queries", type=int, default=200)
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
    ix, cent, _ = bench.build_index(pkg, torch, ctx, a, 0, 1, dev)
    k, nq = a.k, a.queries
    x = bench.gen_unit_rows(torch, bench.SEED_QUERY, 4242, nq, dev)
    torch.cuda.synchronize()
    qm = cp.EmptyMatrix(nq, bench.D, ctx=ctx)
    qm.FillFloat32Dev(0, x.data_ptr(), nq, ctx=ctx)
    ctx.sync()
    qh = qm.ReadRows()
    d_ids = torch.zeros((1, k), device=dev, dtype=torch.int64)
    d_sims = torch.zeros((1, k), device=dev, dtype=torch.float32)
    d_counts = torch.zeros(1, device=dev, dtype=torch.int32)
    d_status = torch.zeros(1, device=dev, dtype=torch.int32)
    one = cp.EmptyMatrix(1, bench.D, ctx=ctx)
    offsets = ix.ListOffsets(ctx=ctx).astype(np.int64)
    out = {"rows": a.rows, "centroids": a.centroids, "nprobe": a.nprobe, "k": k, "queries": nq}

    def run(label, nprobe, fused):
        cp.debug_set_fused(fused)
        lat, res, flagged = [], [], 0
        for i in range(nq):
            one.LoadRows(0, qh[i:i + 1], ctx=ctx)
            ctx.sync()
            ctx.timer_start()
            ix.SearchDev(one, nprobe, k, d_ids.data_ptr(), d_sims.data_ptr(), d_counts.data_ptr(), d_status.data_ptr(), ctx=ctx)
            lat.append(ctx.timer_stop() * 1e3)
            flagged += int(d_status.item() != 0)
            ix.Resolve(one, nprobe, k, d_ids.data_ptr(), d_sims.data_ptr(), d_counts.data_ptr(), d_status.data_ptr(), ctx=ctx)
            res.append((d_ids.cpu().numpy().copy(), d_sims.cpu().numpy().copy()))
        cp.debug_set_fused(True)
        lat = np.sort(np.array(lat[8:]))
        out[label] = {"p50_us": round(float(lat[len(lat) // 2]), 2), "p10_us": round(float(lat[len(lat) // 10]), 2),
                      "p99_us": round(float(lat[int(len(lat) * 0.99)]), 2), "min_us": round(float(lat[0]), 2),
                      "queries_sent_to_the_literal_path": flagged}
        return res

    def run_back_to_back(label, nprobe, fused, n=50):
        """n searches of the same query enqueued without a host sync in between: kernel time + launch gap."""
        cp.debug_set_fused(fused)
        one.LoadRows(0, qh[0:1], ctx=ctx)
        for rep in range(2):
            ctx.sync()
            ctx.timer_start()
            for _ in range(n):
                ix.SearchDev(one, nprobe, k, d_ids.data_ptr(), d_sims.data_ptr(), d_counts.data_ptr(), d_status.data_ptr(), ctx=ctx)
            ms = ctx.timer_stop()
        cp.debug_set_fused(True)
        out[label] = {"us_per_search_back_to_back": round(ms * 1e3 / n, 2)}

    a_res = run("ivf_fused", a.nprobe, True)
    run_back_to_back("ivf_fused_b2b", a.nprobe, True)
    run_back_to_back("ivf_two_launch_b2b", a.nprobe, False)
    b_res = run("ivf_two_launch", a.nprobe, False)
    out["ivf_results_equal"] = bool(all((x[0] == y[0]).all() and (x[1].view(np.uint32) == y[1].view(np.uint32)).all()
                                        for x, y in zip(a_res, b_res)))
    probes, _ = ix.SelectProbes(qh[:nq], a.nprobe, ctx=ctx)
    rows_scored = np.diff(offsets)[probes.astype(np.int64)].sum(axis=1)
    bytes_q = (float(rows_scored.mean()) + a.centroids) * bench.ROW_BYTES
    out["ivf_bytes_per_query"] = round(bytes_q)
    for lab in ("ivf_fused", "ivf_two_launch"):
        out[lab]["gbs_at_p50"] = round(bytes_q / (out[lab]["p50_us"] * 1e-6) / 1e9, 1)
    # phase stamps of the fused kernel (%globaltimer, thread 0 of every block; trace on: a few more stores per block)
    ctx.trace_enable(True)
    ph = []
    for i in range(8, 40):
        one.LoadRows(0, qh[i:i + 1], ctx=ctx)
        ctx.sync()
        ix.SearchDev(one, a.nprobe, k, d_ids.data_ptr(), d_sims.data_ptr(), d_counts.data_ptr(), d_status.data_ptr(), ctx=ctx)
        ctx.sync()
        act = ctx.trace_read(1).astype(np.int64)[:148]
        t0 = act[:, 0].min()
        row = []
        for col in list(range(9)) + [9, 12]:
            v = act[:, col]
            v = v[v > 0]
            row.append((float((v - t0).max()) / 1e3) if v.size else float("nan"))
        ok = (act[:, 5] > act[:, 0]) & (act[:, 13] > 0)
        row.append(float(np.median(act[ok, 13] / ((act[ok, 5] - act[ok, 0]) / 1e3))) if ok.any() else float("nan"))   # MHz
        ph.append(row)
    ctx.trace_enable(False)
    ph = np.array(ph)
    names = ["start_spread", "probe_scored", "grid_barrier_passed", "probes_selected_and_chunk_tables", "scan_done_warp0", "published",
             "last_block_starts", "collected_and_ranked", "emitted", "scan_done_all_warps", "block_selected", "sm_mhz_in_kernel"]
    out["fused_phase_us_max_over_blocks_median_over_queries"] = {n: round(float(np.nanmedian(ph[:, i])), 2) for i, n in enumerate(names)}
    # config 1: flat scan over 100k rows (and rotating over 4 stores so that rows come from HBM, not L2)
    n1 = 100_000
    stores = []
    for j in range(5):
        xj = bench.gen_unit_rows(torch, bench.SEED_DATA, 9200 + j, n1, dev)
        torch.cuda.synchronize()
        mj = cp.EmptyMatrix(n1, bench.D, ctx=ctx)
        mj.FillFloat32Dev(0, xj.data_ptr(), n1, ctx=ctx)
        ctx.sync()
        stores.append(pkg.ivf.Index.build_dev(mj, torch.zeros(n1, dtype=torch.int32, device=dev).data_ptr(), None, cp.NewMatrix(qh[:1], ctx=ctx),
                                              ctx=ctx))
        del xj
    for label, fused in (("flat100k_fused", True), ("flat100k_two_launch", False)):
        cp.debug_set_fused(fused)
        for mode in ("l2_warm", "rotating"):
            lat = []
            for i in range(nq):
                st = stores[0] if mode == "l2_warm" else stores[1 + i % 4]
                one.LoadRows(0, qh[i:i + 1], ctx=ctx)
                ctx.sync()
                ctx.timer_start()
                st.SearchDev(one, 1, k, d_ids.data_ptr(), d_sims.data_ptr(), d_counts.data_ptr(), d_status.data_ptr(), ctx=ctx)
                lat.append(ctx.timer_stop() * 1e3)
            lat = np.sort(np.array(lat[8:]))
            p50 = float(lat[len(lat) // 2])
            out[f"{label}_{mode}"] = {"p50_us": round(p50, 2), "p99_us": round(float(lat[int(len(lat) * 0.99)]), 2),
                                      "gbs_at_p50": round(n1 * bench.ROW_BYTES / (p50 * 1e-6) / 1e9, 1)}
    cp.debug_set_fused(True)
    print(json.dumps(out, indent=1))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
