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


def upload_stripe(first_key, n_new, rank, world):
    """Upload on a striped store (server/upload.go:239-279): the new embeddings get the primary keys first_key ..
    first_key + n_new - 1 and the stripe rule is unchanged -- rank r owns the keys with key % world == r.  Returns the
    indices (into the uploaded batch, ascending) of the rows this rank passes to its own Index.Upload; the ranks need no
    exchange because the centroid table is replicated, and the next search merges the shards as before."""
    start = (rank - first_key) % world
    return np.arange(start, n_new, world, dtype=np.int64)


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


class PackedHits:
    """Shard-local hit buffers laid out for ONE all-gather per step: [ids nq*k u64 | sims nq*k f32 | counts nq i32]
    in a single device allocation per rank, plus the gathered [world] copy and the merged outputs."""

    def __init__(self, nq, k, device, world):
        import torch
        self.nq, self.k, self.world = nq, k, world
        self.ids_off = 0
        self.sims_off = nq * k * 8
        self.counts_off = self.sims_off + ((nq * k * 4 + 7) // 8) * 8
        self.nbytes = self.counts_off + ((nq * 4 + 15) // 16) * 16
        self.local = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.gathered = torch.zeros(self.nbytes * world, dtype=torch.uint8, device=device)
        self.ids = self.local[self.ids_off:self.ids_off + nq * k * 8].view(torch.int64).view(nq, k)
        self.sims = self.local[self.sims_off:self.sims_off + nq * k * 4].view(torch.float32).view(nq, k)
        self.counts = self.local[self.counts_off:self.counts_off + nq * 4].view(torch.int32)
        self.out_ids = torch.zeros((nq, k), dtype=torch.int64, device=device)
        self.out_sims = torch.zeros((nq, k), dtype=torch.float32, device=device)
        self.out_counts = torch.zeros(nq, dtype=torch.int32, device=device)

    def gather_and_merge(self, ctx=None, group=None):
        """One NCCL all-gather of the packed buffer, then the device merge (asynchronous on the shared stream)."""
        import ctypes as C
        import torch.distributed as dist
        from . import _lib
        from .compute import _check, default_context
        dist.all_gather_into_tensor(self.gathered, self.local, group=group)
        ctx = ctx or default_context()
        L = _lib.init()
        vp = lambda t: C.c_void_p(t.data_ptr())
        _check(L.vs_topk_merge_packed_dev(ctx.handle, vp(self.gathered), self.nbytes, self.ids_off, self.sims_off,
                                          self.counts_off, self.world, self.nq, self.k, vp(self.out_ids), vp(self.out_sims),
                                          vp(self.out_counts)))


class _DevArray:
    """A raw device pointer with the CUDA array interface, so that torch can wrap memory libvscuda owns."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerHits:
    """PackedHits without the collective: the shard-local hits live in a libvscuda allocation that every other rank has
    mapped through CUDA IPC (vs_exchange_*): per step one warp signals the peers over NVLink and waits for their signals
    (flag words in peer memory), and the merge kernel reads the peers' hits in place.  Same attributes as PackedHits; `ids`, `sims`, `counts`
    are the views of the slot the NEXT gather_and_merge will exchange.  Construction is collective (one all-gather of the
    64-byte handles and two barriers); every rank must make the same sequence of gather_and_merge calls."""

    def __init__(self, nq, k, device, world, rank, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        from .compute import _check
        self.nq, self.k, self.world, self.rank = nq, k, world, rank
        self.ids_off = 0
        self.sims_off = nq * k * 8
        self.counts_off = self.sims_off + ((nq * k * 4 + 7) // 8) * 8
        self.nbytes = self.counts_off + ((nq * 4 + 15) // 16) * 16
        L = self._L = _lib.init()
        h = C.c_void_p()
        _check(L.vs_exchange_create(rank, world, self.nbytes, C.byref(h)))
        self._h = h
        mine = np.zeros(64, np.uint8)
        _check(L.vs_exchange_handle(h, C.c_void_p(mine.ctypes.data), 64))
        every = torch.zeros(64 * world, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(every, torch.from_numpy(mine).to(device), group=group)   # (also: every rank has created)
        handles = every.cpu().numpy().copy()
        _check(L.vs_exchange_connect(h, C.c_void_p(handles.ctypes.data)))
        dist.barrier(group=group)
        self._views = []
        for slot in (0, 1):
            base = int(L.vs_exchange_slot(h, slot))
            self._views.append((
                torch.as_tensor(_DevArray(base + self.ids_off, (nq, k), "<i8"), device=device),
                torch.as_tensor(_DevArray(base + self.sims_off, (nq, k), "<f4"), device=device),
                torch.as_tensor(_DevArray(base + self.counts_off, (nq,), "<i4"), device=device)))
        self.step = 1
        self.out_ids = torch.zeros((nq, k), dtype=torch.int64, device=device)
        self.out_sims = torch.zeros((nq, k), dtype=torch.float32, device=device)
        self.out_counts = torch.zeros(nq, dtype=torch.int32, device=device)

    ids = property(lambda self: self._views[self.step & 1][0])
    sims = property(lambda self: self._views[self.step & 1][1])
    counts = property(lambda self: self._views[self.step & 1][2])

    def gather_and_merge(self, ctx=None, group=None):
        """The exchange and the merge in one kernel (asynchronous on ctx's stream)."""
        import ctypes as C
        from .compute import _check, default_context
        ctx = ctx or default_context()
        vp = lambda t: C.c_void_p(t.data_ptr())
        _check(self._L.vs_exchange_merge(ctx.handle, self._h, self.step & 0xFFFFFFFF, self.ids_off, self.sims_off, self.counts_off,
                                         self.nq, self.k, vp(self.out_ids), vp(self.out_sims), vp(self.out_counts)))
        self.step += 1

    def close(self):
        if self._h is not None:
            self._L.vs_exchange_release(self._h)
            self._h = None


class SharedProbes:
    """The probe stage of a batch shared between the ranks of a striped index (every rank holds the whole centroid table):
    rank r selects the probe lists of queries [r * nq / world, (r + 1) * nq / world) only (Index.ProbeDev), and two small
    all-gathers (list numbers, status words) hand every rank the lists of the whole batch for its list stage
    (Index.SearchDevProbed).  The replicated form costs every rank the centroid scoring of the WHOLE batch, a per-step
    constant that does not shrink with the shard.  nq must be a multiple of world."""

    def __init__(self, nq, nprobe, device, world, rank):
        import torch
        if nq % world:
            raise ValueError("SharedProbes: the batch must divide evenly among the ranks")
        self.nq, self.nprobe, self.world, self.rank, self.per = nq, nprobe, world, rank, nq // world
        self.local_probe = torch.zeros(self.per * nprobe, dtype=torch.int32, device=device)
        self.local_status = torch.zeros(self.per, dtype=torch.int32, device=device)
        self.probe = torch.zeros(nq * nprobe, dtype=torch.int32, device=device)

    def rows_of(self, rank=None):
        """Query rows of the batch whose probe lists `rank` (default: this rank) selects."""
        r = self.rank if rank is None else rank
        return slice(r * self.per, (r + 1) * self.per)

    def select_and_gather(self, ix, q_slice, d_status, ctx, group=None):
        """q_slice: device matrix of this rank's share of the batch; d_status: [nq] int32 device tensor that receives the
        probe stage's status words of the WHOLE batch.  Asynchronous on the CURRENT torch stream, which must be ctx's."""
        import torch.distributed as dist
        ix.ProbeDev(q_slice, self.nprobe, self.local_probe.data_ptr(), self.local_status.data_ptr(), ctx=ctx)
        dist.all_gather_into_tensor(self.probe, self.local_probe, group=group)
        dist.all_gather_into_tensor(d_status, self.local_status, group=group)
        return self.probe


# ---- k-means over a store cut into contiguous row blocks -------------------------------------------
def block_range(n_total, rank, world):
    """Rows [lo, hi) of rank `rank` when the store is cut into `world` contiguous blocks (row order = rank order, which is
    what lets the float32 sums of k_means.go:80-86 be continued from rank to rank in the reference's order)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def kmeans_step_relay(assign, accumulate, finish, sums, counts, centroid_rows, group=None):
    """One Lloyd iteration (dnc/k_means.go:67-117) across ranks, bit-identical to one device.

    Every rank first assigns its own block of rows (`assign()`; at many centroids this is the whole cost and it runs on
    all ranks at once).  The running float32 sums / int64 counts (torch tensors) are then relayed rank 0 -> 1 -> ... in
    row order: a rank receives them (zeros on rank 0), continues them over its rows (`accumulate(sums, counts)`) and
    sends them on; the last rank calls `finish(sums, counts) -> (new centroid rows uint8 tensor [k, 8+d], converged)`
    and broadcasts both.  An all-reduce would be one collective instead of a relay, but float32 addition is not
    associative: the reference's bytes need its order.  `centroid_rows`: uint8 tensor [k, 8+d] that receives the
    broadcast.  Returns converged (bool).
    """
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    assign()
    if rank > 0:
        dist.recv(sums, src=rank - 1, group=group)
        dist.recv(counts, src=rank - 1, group=group)
    else:
        sums.zero_()
        counts.zero_()
    accumulate(sums, counts)
    flag = torch.zeros(1, dtype=torch.int32, device=centroid_rows.device)
    if rank < world - 1:
        dist.send(sums, dst=rank + 1, group=group)
        dist.send(counts, dst=rank + 1, group=group)
    else:
        rows, conv = finish(sums, counts)
        centroid_rows.copy_(rows)
        flag.fill_(1 if conv else 0)
    dist.broadcast(centroid_rows, src=world - 1, group=group)
    dist.broadcast(flag, src=world - 1, group=group)
    return bool(int(flag.item()))
