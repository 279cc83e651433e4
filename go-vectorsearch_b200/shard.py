"""Multi-GPU plumbing for the row-striped index: one process per GPU, rows of the store
striped across ranks by global row id (so every posting list is split evenly for any probe set), the
centroid table replicated, each rank producing a shard-local top-k that is all-gathered and merged.

torch.distributed is plumbing only; the local search and the merge are libvscuda kernels.
"""
import numpy as np


def stripe(n_total, rank, world):
    """Global row ids owned by `rank`: rank, rank+world, ... (ascending, so primary-key order is kept)."""
    return np.arange(rank, n_total, world, dtype=np.int64)


def local_count(n_total, rank, world):
    return (n_total - rank + world - 1) // world if rank < n_total else 0


def gather_hits(ids, sims, counts, group=None):
    """all_gather the shard-local results. ids [nq,k] int64 (uint64 bit pattern), sims [nq,k] float32,
    counts [nq] int32 torch tensors (CPU/gloo or CUDA/nccl). Returns ([G,nq,k], [G,nq,k], [G,nq])."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)

    def gather(t):
        t = t.contiguous()
        out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)  # rank-major concat
        dist.all_gather_into_tensor(out, t, group=group)
        return out.view((world,) + tuple(t.shape))

    g_ids, g_sims, g_counts = gather(ids), gather(sims), gather(counts)
    return g_ids, g_sims, g_counts


def merge_hits_dev(g_ids, g_sims, g_counts, k, out_ids, out_sims, out_counts, ctx=None):
    """Device merge (vs_topk_merge_dev) of gathered CUDA tensors into the global top-k, same order and
    dedup rule as the single-GPU path (similarity desc as float32, then id asc; one hit per document)."""
    from . import ivf
    G, nq = int(g_ids.shape[0]), int(g_ids.shape[1])
    ivf.TopKMergeDev(g_ids.data_ptr(), g_sims.data_ptr(), g_counts.data_ptr(), G, nq, k, out_ids.data_ptr(),
                     out_sims.data_ptr(), out_counts.data_ptr(), ctx=ctx)
