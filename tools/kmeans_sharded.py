"""k-means iterations over a store cut into contiguous row blocks, one block per GPU (torchrun, NCCL relay of the running
float32 sums; shard.kmeans_step_relay).  Rank 0 prints one JSON line: seconds per iteration (max over ranks), and --
with --check -- whether the centroid bytes equal those of the same iterations on ONE device (rank 0 re-runs them on the
whole store with vs_kmeans_step's kernels; needs the store to fit one GPU).

  torchrun --nproc-per-node N tools/kmeans_sharded.py --rows 20000000 --k 65536 --iters 2 --check
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B


def fill(cp, torch, ctx, device, lo, hi, seed):
    """Rows [lo, hi) of the synthetic store (chunk-keyed generator: any rank regenerates any chunk)."""
    m = cp.EmptyMatrix(hi - lo, B.D, ctx=ctx)
    done = 0
    for ci in range(lo // B.CHUNK, (hi + B.CHUNK - 1) // B.CHUNK):
        c0 = ci * B.CHUNK
        x = B.gen_unit_rows(torch, seed, ci, B.CHUNK, device)
        a, b = max(lo, c0) - c0, min(hi, c0 + B.CHUNK) - c0
        xs = x[a:b].contiguous()
        torch.cuda.synchronize()
        m.FillFloat32Dev(done, xs.data_ptr(), b - a, ctx=ctx)
        ctx.sync()
        done += b - a
        del x, xs
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=20_000_000)
    ap.add_argument("--k", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_pkg
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    pkg = load_pkg()
    pkg._lib.init(local_rank)
    cp, dnc, shard = pkg.compute, pkg.dnc, pkg.shard
    stream = torch.cuda.Stream(device=device)
    ctx = cp.Context(cuda_stream=stream.cuda_stream)   # libvscuda kernels and NCCL share one stream
    lo, hi = shard.block_range(a.rows, rank, world)
    k, d = a.k, B.D
    with torch.cuda.stream(stream):
        data = fill(cp, torch, ctx, device, lo, hi, B.SEED_DATA)
        cent0 = fill(cp, torch, ctx, device, 0, k, B.SEED_CENT)      # the same initial centroids on every rank
        cent_rows = torch.from_numpy(cent0.ReadRows()).to(device)
        cmat = cent0
        sums = torch.zeros(k * d, dtype=torch.float32, device=device)
        counts = torch.zeros(k, dtype=torch.int64, device=device)
        means = torch.zeros(k * d, dtype=torch.float32, device=device)
        assign = torch.empty(hi - lo, dtype=torch.int32, device=device)
        times, convs = [], []
        for it in range(a.iters):
            state = {}

            def do_assign():
                state["cmat"].ArgmaxDev(data, assign.data_ptr(), ctx=ctx)

            def do_accumulate(s, c):
                dnc.KMeansAccumulateDev(data, k, assign.data_ptr(), s.data_ptr(), c.data_ptr(), ctx=ctx)

            def do_finish(s, c):
                newm, conv = dnc.KMeansFinishDev(state["cmat"], s.data_ptr(), c.data_ptr(), means.data_ptr(), ctx=ctx)
                return torch.from_numpy(newm.ReadRows()).to(device), conv

            state["cmat"] = cmat
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            conv = shard.kmeans_step_relay(do_assign, do_accumulate, do_finish, sums, counts, cent_rows)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            times.append(float(dt.item()))
            convs.append(conv)
            cmat = cp.NewMatrix(cent_rows.cpu().numpy(), ctx=ctx)       # every rank: the broadcast centroids
        final_rows = cent_rows.cpu().numpy()
    out = {"workload": f"k-means iteration over {world} contiguous row blocks, {a.rows} x {d}-d uint8 rows, {k} centroids",
           "n_gpus": world, "s_per_iteration": [round(t, 4) for t in times], "converged": convs,
           "assign_int_ops_per_iteration": 2.0 * a.rows * k * d,
           "tops_whole_iteration": round(2.0 * a.rows * k * d / min(times) / 1e12, 1)}
    if a.check and rank == 0:
        # the same iterations on one device: whole store, vs_kmeans_step's kernels, state kept on the device
        del data
        torch.cuda.empty_cache()
        with torch.cuda.stream(stream):
            full = fill(cp, torch, ctx, device, 0, a.rows, B.SEED_DATA)
            c1 = fill(cp, torch, ctx, device, 0, k, B.SEED_CENT)
            s1 = torch.zeros(k * d, dtype=torch.float32, device=device)
            n1 = torch.zeros(k, dtype=torch.int64, device=device)
            m1 = torch.zeros(k * d, dtype=torch.float32, device=device)
            a1 = torch.empty(a.rows, dtype=torch.int32, device=device)
            t1 = []
            for it in range(a.iters):
                s1.zero_(); n1.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                c1.ArgmaxDev(full, a1.data_ptr(), ctx=ctx)
                dnc.KMeansAccumulateDev(full, k, a1.data_ptr(), s1.data_ptr(), n1.data_ptr(), ctx=ctx)
                c1, _ = dnc.KMeansFinishDev(c1, s1.data_ptr(), n1.data_ptr(), m1.data_ptr(), ctx=ctx)
                torch.cuda.synchronize()
                t1.append(time.perf_counter() - t0)
            out["one_device_s_per_iteration"] = [round(t, 4) for t in t1]
            out["centroid_bytes_equal_one_device"] = bool((c1.ReadRows() == final_rows).all())
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
