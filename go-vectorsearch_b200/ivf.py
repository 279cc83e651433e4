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

    @property
    def rows(self):
        return int(self._L.vs_index_rows(self._h))

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
