"""ctypes binding of oracle/_build/liboracle.so (built by oracle/Makefile)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
# the same source with the float64 sums free to be reordered and vectorized: a TIMING proxy for the reference's BLAS
# (gonum) backend, never a parity anchor (see the note at the top of oracle.c)
_SO_PROXY = os.path.join(_HERE, "_build", "liboracle_blas_proxy.so")


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    stale = any(not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src) for so in (_SO, _SO_PROXY))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None
_lib_proxy = None


def lib_blas_proxy():
    global _lib_proxy
    if _lib_proxy is None:
        if not os.path.exists(_SO_PROXY):
            build()
        _lib_proxy = C.CDLL(_SO_PROXY)
    return _lib_proxy


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.ora_quantize_f32.restype = C.c_uint8
        _lib.ora_quantize_f32.argtypes = [C.c_float] * 3
        _lib.ora_quantize_f64.restype = C.c_uint8
        _lib.ora_quantize_f64.argtypes = [C.c_double] * 3
        _lib.ora_dequantize_f32.restype = C.c_float
        _lib.ora_dequantize_f32.argtypes = [C.c_uint8, C.c_float, C.c_float]
        _lib.ora_dequantize_f64.restype = C.c_double
        _lib.ora_dequantize_f64.argtypes = [C.c_uint8, C.c_double, C.c_double]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def quantize_vector_f32(v):
    v = np.ascontiguousarray(v, dtype=np.float32)
    out = np.empty(8 + v.shape[0], np.uint8)
    lib().ora_quantize_vector_f32(_p(v), C.c_size_t(v.shape[0]), _p(out))
    return out


def quantize_vector_f64(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.empty(8 + v.shape[0], np.uint8)
    lib().ora_quantize_vector_f64(_p(v), C.c_size_t(v.shape[0]), _p(out))
    return out


def quantize_matrix_f32(m):
    m = np.ascontiguousarray(m, dtype=np.float32)
    n, d = m.shape
    out = np.empty((n, 8 + d), np.uint8)
    lib().ora_quantize_matrix_f32(_p(m), C.c_size_t(n), C.c_size_t(d), _p(out))
    return out


def quantize_matrix_f64(m):
    m = np.ascontiguousarray(m, dtype=np.float64)
    n, d = m.shape
    out = np.empty((n, 8 + d), np.uint8)
    lib().ora_quantize_matrix_f64(_p(m), C.c_size_t(n), C.c_size_t(d), _p(out))
    return out


def dequantize_matrix_f32(rows):
    rows = _u8(rows)
    n, rb = rows.shape
    out = np.empty((n, rb - 8), np.float32)
    lib().ora_dequantize_matrix_f32(_p(rows), C.c_size_t(n), C.c_size_t(rb), _p(out))
    return out


def dequantize_matrix_f64(rows):
    rows = _u8(rows)
    n, rb = rows.shape
    out = np.empty((n, rb - 8), np.float64)
    lib().ora_dequantize_matrix_f64(_p(rows), C.c_size_t(n), C.c_size_t(rb), _p(out))
    return out


def dot_u8_1xN(q, rows):
    q = _u8(q)
    rows = _u8(rows)
    n, rb = rows.shape
    out = np.empty(n, np.uint32)
    lib().ora_dot_u8_1xN(_p(q), _p(rows), C.c_size_t(n), C.c_size_t(rb), _p(out))
    return out


class OraclePanic(Exception):
    """The reference panics / Fatalf's here (compute.go:13,26,30; cosine.go:19-21,77-79)."""


def _check(rc):
    if rc == -1:
        raise OraclePanic("empty vector/matrix")
    if rc == -2:
        raise OraclePanic("column size does not match")


def cosine_1xN(q, rows, blas_proxy=False):
    q = _u8(q)
    rows = _u8(rows)
    n, rb = rows.shape if rows.ndim == 2 else (0, 0)
    out = np.empty(n, np.float32)
    rc = (lib_blas_proxy() if blas_proxy else lib()).ora_cosine_1xN(_p(q), C.c_size_t(q.shape[0]), _p(rows), C.c_size_t(n), C.c_size_t(rb), _p(out))
    _check(rc)
    return out


def argmax_MxN(cent, rows):
    cent = _u8(cent)
    rows = _u8(rows)
    m, cb = cent.shape if cent.ndim == 2 else (0, 0)
    n, rb = rows.shape if rows.ndim == 2 else (0, 0)
    sims = np.empty(n, np.float32)
    idx = np.empty(n, np.int64)
    rc = lib().ora_argmax_MxN(_p(cent), C.c_size_t(m), C.c_size_t(cb), _p(rows), C.c_size_t(n), C.c_size_t(rb),
                              _p(sims), _p(idx))
    _check(rc)
    return sims, idx


def upload(centroids, new_rows, list_of_row, rows=None, doc_ids=None, new_doc_ids=None):
    """The state of the embeddings table after Upload (server/upload.go:239-279): the new rows are appended (larger
    primary keys, upload.go:284-287) with CentroidID = the nearest centroid (centroids.MatrixCosineSimilarity(new rows),
    upload.go:245).  Returns (assign of the new rows, list_of_row of old + new, rows, doc_ids); the last two are None
    when not given."""
    _, assign = argmax_MxN(centroids, new_rows)
    lists = np.concatenate([np.asarray(list_of_row, dtype=np.uint32), assign.astype(np.uint32)])
    all_rows = None if rows is None else np.concatenate([_u8(rows), _u8(new_rows)])
    all_ids = None if doc_ids is None else np.concatenate([np.asarray(doc_ids, np.uint64), np.asarray(new_doc_ids, np.uint64)])
    return assign, lists, all_rows, all_ids


def select_probes(q, centroids, nprobe):
    q = _u8(q)
    centroids = _u8(centroids)
    Cn, rb = centroids.shape
    keep = min(nprobe, Cn)
    probes = np.empty(max(keep, 1), np.uint32)
    sims = np.empty(max(keep, 1), np.float32)
    rc = lib().ora_select_probes(_p(q), _p(centroids), C.c_size_t(Cn), C.c_size_t(rb), C.c_size_t(nprobe),
                                 _p(probes), _p(sims))
    _check(rc)
    return probes[:rc], sims[:rc]


def search(q, centroids, rows, list_of_row, doc_ids, nprobe, k, blas_proxy=False):
    q = _u8(q)
    centroids = _u8(centroids)
    rows = _u8(rows)
    list_of_row = np.ascontiguousarray(list_of_row, dtype=np.uint32)
    doc_ids = np.ascontiguousarray(doc_ids, dtype=np.uint64)
    Cn, rb = centroids.shape
    n = rows.shape[0]
    ids = np.empty(k, np.uint64)
    sims = np.empty(k, np.float32)
    rc = (lib_blas_proxy() if blas_proxy else lib()).ora_search(_p(q), _p(centroids), C.c_size_t(Cn), _p(rows), C.c_size_t(n), C.c_size_t(rb),
                          _p(list_of_row), _p(doc_ids), C.c_size_t(nprobe), C.c_size_t(k), _p(ids), _p(sims))
    _check(rc)
    return ids[:rc], sims[:rc]


def search_many(qs, centroids, rows, list_of_row, doc_ids, nprobe, k, threads=1, blas_proxy=False):
    qs = _u8(qs)
    centroids = _u8(centroids)
    rows = _u8(rows)
    list_of_row = np.ascontiguousarray(list_of_row, dtype=np.uint32)
    doc_ids = np.ascontiguousarray(doc_ids, dtype=np.uint64)
    nq = qs.shape[0]
    Cn, rb = centroids.shape
    n = rows.shape[0]
    ids = np.zeros((nq, k), np.uint64)
    sims = np.zeros((nq, k), np.float32)
    counts = np.zeros(nq, np.int32)
    rc = (lib_blas_proxy() if blas_proxy else lib()).ora_search_many(_p(qs), C.c_size_t(nq), _p(centroids), C.c_size_t(Cn), _p(rows), C.c_size_t(n),
                               C.c_size_t(rb), _p(list_of_row), _p(doc_ids), C.c_size_t(nprobe), C.c_size_t(k),
                               _p(ids), _p(sims), _p(counts), C.c_int(threads))
    _check(rc)
    return ids, sims, counts


def search_flat(q, rows, doc_ids, k):
    q = _u8(q)
    rows = _u8(rows)
    n, rb = rows.shape
    ids = np.empty(k, np.uint64)
    sims = np.empty(k, np.float32)
    dp = None
    if doc_ids is not None:
        doc_ids = np.ascontiguousarray(doc_ids, dtype=np.uint64)
        dp = _p(doc_ids)
    rc = lib().ora_search_flat(_p(q), _p(rows), C.c_size_t(n), C.c_size_t(rb), dp, C.c_size_t(k), _p(ids), _p(sims))
    _check(rc)
    return ids[:rc], sims[:rc]


def kmeans_step(data, centroids, means):
    """One Lloyd iteration (dnc/k_means.go:67-117). `means` ([k][d] float32) is updated in place."""
    data = _u8(data)
    centroids = _u8(centroids)
    assert means.dtype == np.float32 and means.flags.c_contiguous
    n, rb = data.shape
    k = centroids.shape[0]
    assign = np.empty(n, np.int64)
    counts = np.empty(k, np.int64)
    newc = np.empty((k, rb), np.uint8)
    rc = lib().ora_kmeans_step(_p(data), C.c_size_t(n), _p(centroids), C.c_size_t(k), C.c_size_t(rb), _p(means),
                               _p(assign), _p(counts), _p(newc))
    _check(rc)
    return assign, counts, newc, bool(rc)


def recenter(rows):
    rows = _u8(rows)
    n, rb = rows.shape
    out = np.empty(rb, np.uint8)
    lib().ora_recenter(_p(rows), C.c_size_t(n), C.c_size_t(rb), _p(out))
    return out


def kmeans(data, k, superset_rows, limit=1000):
    """kMeans (dnc/k_means.go:19-212) with the random superset draw (:35-44) supplied by the caller: iterate the
    superset until the code bytes stop changing (:67-117), keep the first k (the sort at :132 compares counts that
    :111-113 have already zeroed, so Go's pattern-defeating quicksort leaves the all-equal slice as it is), iterate the
    set (:157-207).  The float32 means survive from iteration to iteration and from the superset to the set (:153-154).
    Returns (centroids, iterations of the superset phase, iterations of the set phase)."""
    data = _u8(data)
    if k <= 0:
        return None, 0, 0                      # :20-22
    if data.shape[0] == 0 or data.shape[0] <= k:
        return data, 0, 0                      # :24-26
    cent = data[np.asarray(superset_rows, np.int64)].copy()
    means = np.zeros((cent.shape[0], data.shape[1] - 8), np.float32)
    iters = []
    for phase in range(2):
        if phase == 1:
            cent = np.ascontiguousarray(cent[:k])
            means = np.ascontiguousarray(means[:k])
        n, conv = 0, False
        while n < limit and not conv:
            _, _, cent, conv = kmeans_step(data, cent, means)
            n += 1
        iters.append(n)
    return cent, iters[0], iters[1]


def divide_and_conquer(rows, target_size, sample_size, split_size, rng, limit=1000):
    """divideNconquer (dnc/dnc.go:300-400) on host arrays.  The reference seeds each draw from the clock and runs children
    concurrently, so only a restatement with agreed draws can be compared: every node owns a generator (the root's is
    `rng`); it draws its sample (sampling.go:12-74: all rows or sample_size sorted distinct rows), then its k-means
    superset, then hands its i-th non-empty child the i-th generator of `spawn()`.  A set of at most target_size rows
    yields kMeans(sample, 1)[0] (dataset.go:93-98); a larger one is split by nearest centroid of
    kMeans(sample, min(split, max(2, rows/target))) (dnc.go:330-389).  Leaves in depth-first order, children in index order."""
    rows = _u8(rows)

    def sample(x, g):
        if x.shape[0] <= sample_size:
            return x
        return x[np.sort(g.choice(x.shape[0], sample_size, replace=False))]

    def km(x, k, g):
        if x.shape[0] == 0 or x.shape[0] <= k:
            return x
        ks = min(x.shape[0], 5 * k)
        return kmeans(x, k, g.choice(x.shape[0], ks, replace=False), limit)[0]

    out, stack = [], [(rows, rng)]
    while stack:
        x, g = stack.pop()
        s = sample(x, g)
        if x.shape[0] <= target_size:
            out.append(km(s, 1, g)[0])
            continue
        cents = km(s, min(split_size, max(2, x.shape[0] // target_size)), g)
        _, idx = argmax_MxN(cents, x)
        children = [c for c in (x[idx == j] for j in range(cents.shape[0])) if c.shape[0]]
        stack.extend(reversed(list(zip(children, g.spawn(len(children))))))
    return np.stack(out)


def reassign_recenter(rows, centroids):
    """The tail of KMeansDivideAndConquer: nearest new centroid for every row (dnc.go:194-208), dropSmallCentroids (a no-op
    upstream, see dnc.py), recenterDbCentroid for every centroid (dnc.go:402-456; a cluster without members divides by
    zero: NaN -> all-zero codes)."""
    rows = _u8(rows)
    centroids = _u8(centroids)
    _, idx = argmax_MxN(centroids, rows)
    out = np.empty_like(centroids)
    counts = np.zeros(centroids.shape[0], np.int64)
    for j in range(centroids.shape[0]):
        members = rows[idx == j]
        counts[j] = members.shape[0]
        if members.shape[0]:
            out[j] = recenter(members)
        else:
            d = rows.shape[1] - 8
            with np.errstate(invalid="ignore"):
                out[j] = quantize_vector_f64(np.full(d, np.nan))
    return idx.astype(np.int32), out, counts
