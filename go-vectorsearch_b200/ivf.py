"""Device-resident IVF-Flat probe-and-scan: server/search.go:202-273 behind the C ABI.

The database layer is demoted to a loader: rows (Embedding.Vector, database/model.go:11), their
DocumentID (:14) and CentroidID (:16), and the Centroid.Vector table (:37) are streamed into HBM once
(Index.build*); Search then runs centroid scoring, the nprobe cut, the posting-list scan, the running
top-k with dedup-by-document and the truncation entirely on device.
"""
import ctypes as C

import numpy as np

from . import _lib
from .compute import Context, _check, _p, _rows_array, default_context


class Index:
    def __init__(self, handle, L):
        self._h = handle
        self._L = L

    @property
    def handle(self):
        return self._h

    @classmethod
    def build(cls, rows, doc_ids, list_offsets, centroids, ctx=None):
        """Rows already grouped by list; list_offsets[C+1] is the CSR of the lists."""
        L = _lib.init()
        ctx = ctx or default_context()
        rows = _rows_array(rows)
        centroids = _rows_array(centroids)
        off = np.ascontiguousarray(list_offsets, dtype=np.uint64)
        ids = None if doc_ids is None else np.ascontiguousarray(doc_ids, dtype=np.uint64)
        h = C.c_void_p()
        _check(L.vs_index_build(ctx.handle, _p(rows), rows.shape[0], rows.shape[1], _p(ids) if ids is not None else None,
                                _p(off), _p(centroids), centroids.shape[0], C.byref(h)))
        return cls(h, L)

    @classmethod
    def build_assigned(cls, rows, doc_ids, list_of_row, centroids, ctx=None):
        """Rows in primary-key order + the centroid index of each (Embedding.CentroidID)."""
        L = _lib.init()
        ctx = ctx or default_context()
        rows = _rows_array(rows)
        centroids = _rows_array(centroids)
        lor = np.ascontiguousarray(list_of_row, dtype=np.uint32)
        ids = None if doc_ids is None else np.ascontiguousarray(doc_ids, dtype=np.uint64)
        h = C.c_void_p()
        _check(L.vs_index_build_assigned(ctx.handle, _p(rows), rows.shape[0], rows.shape[1],
                                         _p(ids) if ids is not None else None, _p(lor), _p(centroids),
                                         centroids.shape[0], C.byref(h)))
        return cls(h, L)

    @classmethod
    def build_dev(cls, data, d_list_of_row_ptr, d_doc_ids_ptr, centroids, id_base=0, ctx=None):
        """Device-resident build: data/centroids are compute.Matrix, the two pointers are device arrays
        (int32 list index per row; uint64 ids or 0/None for id_base + row)."""
        L = _lib.init()
        ctx = ctx or default_context()
        h = C.c_void_p()
        _check(L.vs_index_build_dev(ctx.handle, data.handle, C.c_void_p(int(d_list_of_row_ptr)),
                                    C.c_void_p(int(d_doc_ids_ptr)) if d_doc_ids_ptr else None, int(id_base),
                                    centroids.handle, C.byref(h)))
        return cls(h, L)

    @classmethod
    def create_empty(cls, centroids, list_counts, ctx=None):
        """Streaming loader, step 1: reserve the grouped store from the per-list row counts (centroids: compute.Matrix).
        Fill it with Fill / FillDev in primary-key order; searches answer from the rows placed so far."""
        L = _lib.init()
        ctx = ctx or default_context()
        counts = np.ascontiguousarray(list_counts, dtype=np.uint64)
        assert counts.shape == (centroids.rows,)
        h = C.c_void_p()
        _check(L.vs_index_create_empty(ctx.handle, centroids.handle, _p(counts), C.byref(h)))
        return cls(h, L)

    def FillDev(self, chunk, d_list_of_row_ptr, d_doc_ids_ptr=None, id_base=0, ctx=None):
        """Place a device chunk (compute.Matrix; int32 list index per row and optional uint64 ids as raw device pointers)."""
        ctx = ctx or default_context()
        _check(self._L.vs_index_fill_dev(ctx.handle, self._h, chunk.handle, C.c_void_p(int(d_list_of_row_ptr)),
                                         C.c_void_p(int(d_doc_ids_ptr)) if d_doc_ids_ptr else None, int(id_base)))

    def Fill(self, rows, list_of_row, doc_ids=None, id_base=0, ctx=None):
        """Place a chunk of host rows (row776) with their list index and ids (None: id_base + row index in the chunk)."""
        ctx = ctx or default_context()
        rows = _rows_array(rows)
        lor = np.ascontiguousarray(list_of_row, dtype=np.uint32)
        ids = None if doc_ids is None else np.ascontiguousarray(doc_ids, dtype=np.uint64)
        assert lor.shape == (rows.shape[0],) and (ids is None or ids.shape == lor.shape)
        _check(self._L.vs_index_fill(ctx.handle, self._h, _p(rows), rows.shape[0], rows.shape[1], _p(lor),
                                     _p(ids) if ids is not None else None, int(id_base)))

    def Upload(self, rows, doc_ids=None, ctx=None):
        """Upload's assignment and insert (server/upload.go:239-279): every new row joins the list of its nearest centroid
        (upload.go:245) behind the rows already there.  Returns (new Index, assign int64[n]); this Index stays valid."""
        ctx = ctx or default_context()
        rows = _rows_array(rows)
        ids = None if doc_ids is None else np.ascontiguousarray(doc_ids, dtype=np.uint64)
        assert ids is None or ids.shape == (rows.shape[0],)
        assign = np.empty(rows.shape[0], np.int64)
        h = C.c_void_p()
        _check(self._L.vs_index_upload(ctx.handle, self._h, _p(rows), rows.shape[0], rows.shape[1],
                                       _p(ids) if ids is not None else None, _p(assign), C.byref(h)))
        return Index(h, self._L), assign

    def WithRoom(self, percent=25, min_rows=64, ctx=None):
        """A copy of this index in which every list can grow by max(percent % of its rows, min_rows) before anything has to
        be copied again (vs_index_with_room)."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        _check(self._L.vs_index_with_room(ctx.handle, self._h, int(percent), int(min_rows), C.byref(h)))
        return Index(h, self._L)

    def Append(self, rows, doc_ids=None, ctx=None):
        """Upload's assignment and insert (server/upload.go:239-279) in place: every new row is written behind the last row of
        its nearest centroid's list.  Returns assign int64[n]; raises compute.IndexFull, with the index unchanged, when a
        list has no room (UploadInPlace handles that)."""
        ctx = ctx or default_context()
        rows = _rows_array(rows)
        ids = None if doc_ids is None else np.ascontiguousarray(doc_ids, dtype=np.uint64)
        assert ids is None or ids.shape == (rows.shape[0],)
        assign = np.empty(rows.shape[0], np.int64)
        _check(self._L.vs_index_append(ctx.handle, self._h, _p(rows), rows.shape[0], rows.shape[1],
                                       _p(ids) if ids is not None else None, _p(assign)))
        return assign

    def UploadInPlace(self, rows, doc_ids=None, percent=25, min_rows=64, ctx=None):
        """Append, growing when needed: when a list is full (or this index was built without room) the store is copied once
        into a roomier one -- amortised over the appends that follow, like a growing slice -- and this object continues as
        the new index (searches in flight on the old store must have been waited for).  Returns (assign, copied: bool)."""
        from .compute import IndexFull
        rows = _rows_array(rows)
        try:
            return self.Append(rows, doc_ids, ctx=ctx), False
        except IndexFull:
            pass
        grown = self.WithRoom(percent, max(min_rows, rows.shape[0]), ctx=ctx)      # every list can take all the new rows
        old, self._h, grown._h = self._h, grown._h, None
        self._L.vs_index_release(old)
        return self.Append(rows, doc_ids, ctx=ctx), True

    @property
    def capacity(self):
        return int(self._L.vs_index_capacity(self._h))

    def ListLengths(self, ctx=None):
        ctx = ctx or default_context()
        out = np.empty(self.lists, np.uint64)
        _check(self._L.vs_index_list_lengths(ctx.handle, self._h, _p(out)))
        return out

    def ListOffsets(self, ctx=None):
        ctx = ctx or default_context()
        out = np.empty(self.lists + 1, np.uint64)
        _check(self._L.vs_index_list_offsets(ctx.handle, self._h, _p(out)))
        return out

    def ReadRows(self, first, count, ctx=None):
        """(rows776, ids) of the grouped store."""
        ctx = ctx or default_context()
        width = 8 + self.cols
        rows = np.empty((count, width), np.uint8)
        ids = np.empty(count, np.uint64)
        _check(self._L.vs_index_read_rows(ctx.handle, self._h, int(first), int(count), _p(rows), _p(ids)))
        return rows, ids

    def SearchDev(self, queries, nprobe, k, d_ids, d_sims, d_counts, d_status, ctx=None):
        """Asynchronous device-resident search: queries is a compute.Matrix, outputs are raw device pointers."""
        ctx = ctx or default_context()
        _check(self._L.vs_search_dev(ctx.handle, self._h, queries.handle, int(nprobe), int(k), C.c_void_p(int(d_ids)),
                                     C.c_void_p(int(d_sims)), C.c_void_p(int(d_counts)), C.c_void_p(int(d_status))))

    def ProbeDev(self, queries, nprobe, d_probe, d_status, ctx=None):
        """The centroid stage alone (search.go:205-223), device to device and asynchronous: d_probe[nq*nprobe] list numbers
        in rank order, d_status[nq] the stage's status bits.  For an index striped over several devices, each of which
        selects the probe lists of its share of the batch (see SearchDevProbed)."""
        ctx = ctx or default_context()
        _check(self._L.vs_probe_dev(ctx.handle, self._h, queries.handle, int(nprobe), C.c_void_p(int(d_probe)), C.c_void_p(int(d_status))))

    def SearchDevProbed(self, queries, nprobe, k, d_probe, d_ids, d_sims, d_counts, d_status, ctx=None):
        """The posting-list stage alone (search.go:239-273) for probe lists selected elsewhere (ProbeDev on this or another
        device); d_status holds the probe stage's words on entry.  Same hits as SearchDev."""
        ctx = ctx or default_context()
        _check(self._L.vs_search_dev_probed(ctx.handle, self._h, queries.handle, int(nprobe), int(k), C.c_void_p(int(d_probe)),
                                            C.c_void_p(int(d_ids)), C.c_void_p(int(d_sims)), C.c_void_p(int(d_counts)),
                                            C.c_void_p(int(d_status))))

    def SearchBatchDev(self, queries, k, d_ids, d_sims, d_counts, ctx=None):
        """Query batch over the whole store (all lists) as a tensor-core GEMM; device-resident outputs (raw pointers).
        Returns (candidates, queries finished by the scan, store tiles, sampled tiles, us pre-pass, us GEMM, us resolve, 0)."""
        ctx = ctx or default_context()
        stats = np.zeros(8, np.uint64)
        _check(self._L.vs_index_search_batch_dev(ctx.handle, self._h, queries.handle, int(k), C.c_void_p(int(d_ids)),
                                                 C.c_void_p(int(d_sims)), C.c_void_p(int(d_counts)), _p(stats)))
        return tuple(int(x) for x in stats)

    def Resolve(self, queries, nprobe, k, d_ids, d_sims, d_counts, d_status, ctx=None):
        """Finish queries whose float32 roundings could not be certified (reads the status: synchronizes)."""
        ctx = ctx or default_context()
        n = C.c_int(0)
        _check(self._L.vs_search_resolve(ctx.handle, self._h, queries.handle, int(nprobe), int(k),
                                         C.c_void_p(int(d_ids)), C.c_void_p(int(d_sims)), C.c_void_p(int(d_counts)),
                                         C.c_void_p(int(d_status)), C.byref(n)))
        return n.value

    @property
    def rows(self):
        return int(self._L.vs_index_rows(self._h))

    @property
    def cols(self):
        return int(self._L.vs_index_cols(self._h))

    @property
    def lists(self):
        return int(self._L.vs_index_lists(self._h))

    def Search(self, queries, nprobe, k, ctx=None):
        """search.go:115-273 for a batch: returns (ids[nq,k] uint64, sims[nq,k] float32, counts[nq])."""
        ctx = ctx or default_context()
        q = _rows_array(queries if np.asarray(queries).ndim == 2 else np.asarray(queries, np.uint8)[None, :])
        nq = q.shape[0]
        ids = np.zeros((nq, k), np.uint64)
        sims = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.int32)
        _check(self._L.vs_search(ctx.handle, self._h, _p(q), nq, int(nprobe), int(k), _p(ids), _p(sims), _p(counts)))
        return ids, sims, counts

    def SelectProbes(self, queries, nprobe, ctx=None):
        """search.go:202-227: ranked list indices (and their float32 similarities) per query."""
        ctx = ctx or default_context()
        q = _rows_array(queries if np.asarray(queries).ndim == 2 else np.asarray(queries, np.uint8)[None, :])
        nq = q.shape[0]
        keep = min(int(nprobe) if nprobe > 0 else 1, self.lists)
        probes = np.zeros((nq, keep), np.uint32)
        sims = np.zeros((nq, keep), np.float32)
        _check(self._L.vs_select_probes(ctx.handle, self._h, _p(q), nq, int(nprobe), _p(probes), _p(sims)))
        return probes, sims

    def __del__(self):
        try:
            if self._h is not None:
                self._L.vs_index_release(self._h)
                self._h = None
        except Exception:
            pass


def SearchFlat(matrix, queries, k, ctx=None):
    """Brute force over a compute.Matrix (BASELINE config 1): ids are row indices."""
    ctx = ctx or default_context()
    q = _rows_array(queries if np.asarray(queries).ndim == 2 else np.asarray(queries, np.uint8)[None, :])
    nq = q.shape[0]
    ids = np.zeros((nq, k), np.uint64)
    sims = np.zeros((nq, k), np.float32)
    counts = np.zeros(nq, np.int32)
    _check(matrix._L.vs_search_flat(ctx.handle, matrix.handle, None, _p(q), nq, int(k), _p(ids), _p(sims), _p(counts)))
    return ids, sims, counts


def SearchFlatBatch(matrix, queries, k, doc_ids_dev=None, ctx=None):
    """Query batch over the whole store as a tensor-core GEMM (BASELINE config 3); same contract and bits as
    SearchFlat. doc_ids_dev: optional device pointer (int) to uint64 document ids per row."""
    ctx = ctx or default_context()
    q = _rows_array(queries if np.asarray(queries).ndim == 2 else np.asarray(queries, np.uint8)[None, :])
    nq = q.shape[0]
    ids = np.zeros((nq, k), np.uint64)
    sims = np.zeros((nq, k), np.float32)
    counts = np.zeros(nq, np.int32)
    dp = C.c_void_p(int(doc_ids_dev)) if doc_ids_dev else None
    _check(matrix._L.vs_search_flat_gemm(ctx.handle, matrix.handle, dp, _p(q), nq, int(k), _p(ids), _p(sims), _p(counts)))
    return ids, sims, counts


def SearchBatchDev(matrix, queries_matrix, k, d_ids, d_sims, d_counts, doc_ids_dev=None, id_base=0, ctx=None):
    """Device-resident form: results stay in device buffers (raw pointers). Returns the stats tuple
    (candidates, queries finished by the scan, store tiles, sampled tiles, us pre-pass, us GEMM, us resolution, 0)."""
    ctx = ctx or default_context()
    stats = np.zeros(8, np.uint64)
    vp = lambda x: C.c_void_p(int(x)) if x else None
    _check(matrix._L.vs_search_batch_dev(ctx.handle, matrix.handle, vp(doc_ids_dev), int(id_base), queries_matrix.handle, int(k),
                                         vp(d_ids), vp(d_sims), vp(d_counts), _p(stats)))
    return tuple(int(x) for x in stats)


def TopKMergeDev(d_ids_in, d_sims_in, d_counts_in, G, nq, k, d_ids_out, d_sims_out, d_counts_out, ctx=None):
    """Merge G gathered shard-local hit lists per query ([G][nq][k] device arrays) into the global top-k."""
    L = _lib.init()
    ctx = ctx or default_context()
    vp = lambda x: C.c_void_p(int(x))
    _check(L.vs_topk_merge_dev(ctx.handle, vp(d_ids_in), vp(d_sims_in), vp(d_counts_in), int(G), int(nq), int(k),
                               vp(d_ids_out), vp(d_sims_out), vp(d_counts_out)))


class ShardedIndex:
    """One host process driving several GPUs (vs_sharded_*, csrc/sharded.cu): rows striped over the devices by primary
    key, centroids replicated, shard-local top-k merged on device 0 over NVLink.  Mirrors what a `cuda`-tagged
    server.Search would hold instead of a database handle."""

    def __init__(self, devices):
        self._L = _lib.init()
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        _check(self._L.vs_sharded_create(_p(dev), int(dev.shape[0]), C.byref(h)))
        self._h = h
        self.cols = 0

    def close(self):
        if self._h is not None:
            self._L.vs_sharded_release(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @property
    def rows(self):
        return int(self._L.vs_sharded_rows(self._h))

    @property
    def shards(self):
        return int(self._L.vs_sharded_shards(self._h))

    def shard_rows(self, g):
        return int(self._L.vs_sharded_shard_rows(self._h, int(g)))

    def build_assigned(self, rows, doc_ids, list_of_row, centroids):
        rows = _rows_array(rows)
        centroids = _rows_array(centroids)
        lor = np.ascontiguousarray(list_of_row, dtype=np.uint32)
        ids = None if doc_ids is None else np.ascontiguousarray(doc_ids, dtype=np.uint64)
        _check(self._L.vs_sharded_build_assigned(self._h, _p(rows), rows.shape[0], rows.shape[1], _p(ids) if ids is not None else None,
                                                 _p(lor), _p(centroids), centroids.shape[0]))
        self.cols = rows.shape[1] - 8
        return self

    def Upload(self, rows, doc_ids=None):
        rows = _rows_array(rows)
        ids = None if doc_ids is None else np.ascontiguousarray(doc_ids, dtype=np.uint64)
        assign = np.empty(rows.shape[0], np.int64)
        _check(self._L.vs_sharded_upload(self._h, _p(rows), rows.shape[0], rows.shape[1], _p(ids) if ids is not None else None,
                                         _p(assign)))
        return assign

    def Search(self, queries, nprobe, k, sctx=None):
        """sctx: a handle from NewSearchContext() (one per concurrent caller); None = the index's own, one search at a time."""
        q = _rows_array(queries)
        nq = q.shape[0]
        ids = np.zeros((nq, k), np.uint64)
        sims = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.int32)
        if sctx is None:
            _check(self._L.vs_sharded_search(self._h, _p(q), nq, int(nprobe), int(k), _p(ids), _p(sims), _p(counts)))
        else:
            _check(self._L.vs_sharded_search_ctx(sctx, _p(q), nq, int(nprobe), int(k), _p(ids), _p(sims), _p(counts)))
        return ids, sims, counts

    def NewSearchContext(self):
        h = C.c_void_p()
        _check(self._L.vs_sharded_ctx_create(self._h, C.byref(h)))
        return h

    def CloseSearchContext(self, sctx):
        self._L.vs_sharded_ctx_destroy(sctx)
