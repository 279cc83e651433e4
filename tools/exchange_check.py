"""Two or more ranks (torchrun): the striped index answered with the exchange inside the merge kernel (shard.PeerHits,
vs_exchange_*) and with the NCCL all-gather (shard.PackedHits) -- both must equal the CPU oracle over all the rows.
Prints one JSON line on rank 0; exit code 1 on a mismatch.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/exchange_check.py
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_pkg
    import oracle
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    pkg = load_pkg()
    pkg._lib.init(local)
    n, d, C, nq, nprobe, k, steps = 24000, 768, 48, 64, 8, 10, 7
    rng = np.random.default_rng(5)
    x = rng.standard_normal((n + C + nq * steps, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    rows_all = oracle.quantize_matrix_f32(x)
    rows, cent, qs = rows_all[:n], rows_all[n:n + C], rows_all[n + C:].reshape(steps, nq, -1)
    doc = (np.arange(n, dtype=np.uint64) // 3) * 7 + 11          # three embeddings per document: the merge must de-duplicate
    _, lists = oracle.argmax_MxN(cent, rows)
    mine = pkg.shard.stripe(n, rank, world)
    ix = pkg.ivf.Index.build_assigned(rows[mine], doc[mine], lists[mine].astype(np.uint32), cent)
    ctx = pkg.compute.default_context()
    status = torch.zeros(nq, dtype=torch.int32, device=device)
    bad = 0
    outs = {}
    for name, hits in (("peer", pkg.shard.PeerHits(nq, k, device, world, rank)), ("nccl", pkg.shard.PackedHits(nq, k, device, world))):
        got = []
        for s in range(steps):
            qm = pkg.compute.NewMatrix(qs[s], ctx=ctx)
            ix.SearchDev(qm, nprobe, k, hits.ids.data_ptr(), hits.sims.data_ptr(), hits.counts.data_ptr(), status.data_ptr(), ctx=ctx)
            ix.Resolve(qm, nprobe, k, hits.ids.data_ptr(), hits.sims.data_ptr(), hits.counts.data_ptr(), status.data_ptr(), ctx=ctx)
            ctx.sync()
            torch.cuda.synchronize()
            hits.gather_and_merge(ctx=ctx)
            ctx.sync()
            torch.cuda.synchronize()
            got.append((hits.out_ids.cpu().numpy().view(np.uint64).copy(), hits.out_sims.cpu().numpy().copy(),
                        hits.out_counts.cpu().numpy().copy()))
        outs[name] = got
    if rank == 0:
        for s in range(steps):
            for i in range(0, nq, 5):
                wi, ws = oracle.search(qs[s][i], cent, rows, lists.astype(np.uint32), doc, nprobe, k)
                for name in ("peer", "nccl"):
                    ids, sims, cnt = outs[name][s]
                    ok = cnt[i] == len(wi) and ids[i, :cnt[i]].tolist() == wi.tolist() and \
                        (sims[i, :cnt[i]].view(np.uint32) == ws.view(np.uint32)).all()
                    bad += 0 if ok else 1
        print(json.dumps({"world": world, "steps": steps, "queries_per_step": nq, "mismatches_vs_oracle": bad}), flush=True)
    flag = torch.tensor([bad], device=device)
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
