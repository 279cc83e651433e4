"""dnc/k_means.go:67-117 (one Lloyd iteration) and dnc/dnc.go:417-449 (recenter) on device."""
import ctypes as C

import numpy as np

from .compute import _check, _p, _rows_array, default_context


def KMeansStep(data_matrix, centroids, means, ctx=None, want_assign=True):
    """One iteration of the loop at k_means.go:67-117.

    data_matrix: compute.Matrix (device); centroids: (k, 8+d) uint8; means: (k, d) float32 state,
    updated in place (an empty cluster keeps its previous mean, k_means.go:90-92).
    Returns (assign[n] int64 or None, counts[k] int64, new_centroids (k, 8+d) uint8, converged bool).
    """
    ctx = ctx or default_context()
    cent = _rows_array(centroids)
    k, rb = cent.shape
    assert means.dtype == np.float32 and means.flags.c_contiguous and means.shape == (k, rb - 8)
    n = data_matrix.rows
    assign = np.empty(n, np.int64) if want_assign else None
    counts = np.empty(k, np.int64)
    newc = np.empty((k, rb), np.uint8)
    conv = C.c_int(0)
    _check(data_matrix._L.vs_kmeans_step(ctx.handle, data_matrix.handle, _p(cent), k, _p(means),
                                         _p(assign) if want_assign else None, _p(counts), _p(newc), C.byref(conv)))
    return assign, counts, newc, bool(conv.value)


def Recenter(matrix, ctx=None):
    """recenterDbCentroid's arithmetic (dnc.go:417-449) over all rows of `matrix` -> row776."""
    ctx = ctx or default_context()
    out = np.empty(8 + matrix.cols, np.uint8)
    _check(matrix._L.vs_recenter(ctx.handle, matrix.handle, _p(out)))
    return out
