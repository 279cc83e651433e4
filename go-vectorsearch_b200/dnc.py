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


CENTROID_SIZE = 10000            # config/constants.go:8
SAMPLE_SIZE = 50000              # config/constants.go:9
SPLIT_SIZE = 5                   # config/constants.go:10
SUPERSET_MUL = 5                 # config/constants.go:11
KMEANS_ITTERATION_LIMIT = 1000   # config/constants.go:12


def KMeans(data_matrix, k, superset_rows=None, rng=None, iter_limit=KMEANS_ITTERATION_LIMIT, ctx=None, want_stats=False):
    """kMeans (dnc/k_means.go:19-212): the k centroids ((k, 8+d) uint8) of the rows of `data_matrix` (compute.Matrix).

    k <= 0 returns None (:20-22); a matrix with at most k rows is returned unchanged as rows (:24-26).  The reference
    seeds its superset from the wall clock (:29); pass `superset_rows` (min(n, 5k) distinct row indices) or an `rng`
    (numpy Generator) to make a build reproducible.
    """
    ctx = ctx or default_context()
    if k <= 0:
        return None
    n = data_matrix.rows
    if n == 0 or n <= k:
        return data_matrix.ReadRows()
    ks = min(n, k * SUPERSET_MUL)
    if superset_rows is None:
        rng = rng or np.random.default_rng()
        superset_rows = rng.choice(n, ks, replace=False)
    rows = np.ascontiguousarray(superset_rows, dtype=np.uint64)
    assert rows.shape == (ks,) and len(set(rows.tolist())) == ks, "superset rows must be min(n, 5k) distinct indices"
    out = np.empty((k, 8 + data_matrix.cols), np.uint8)
    stats = np.zeros(4, np.int64)
    _check(data_matrix._L.vs_kmeans(ctx.handle, data_matrix.handle, int(k), _p(rows), ks, int(iter_limit), _p(out), _p(stats)))
    if want_stats:
        return out, {"superset_iterations": int(stats[0]), "set_iterations": int(stats[1]),
                     "assign_us": int(stats[2]), "update_us": int(stats[3])}
    return out


def KMeansAccumulateDev(data_matrix, k, d_assign, d_sums, d_counts, ctx=None):
    """One block of rows of a Lloyd iteration (vs_kmeans_accumulate_dev): given the rows' nearest-centroid indices (int32,
    raw device pointer d_assign, from Matrix.ArgmaxDev) continue the float32 sums [k][d] / int64 counts [k] (raw device
    pointers) that the previous block of rows left -- zeros for the first block.  Asynchronous on the ctx stream."""
    ctx = ctx or default_context()
    _check(data_matrix._L.vs_kmeans_accumulate_dev(ctx.handle, data_matrix.handle, int(k), C.c_void_p(int(d_assign)),
                                                   C.c_void_p(int(d_sums)), C.c_void_p(int(d_counts))))


def KMeansFinishDev(centroid_matrix, d_sums, d_counts, d_means, ctx=None):
    """After the last block (vs_kmeans_finish_dev): new means (raw device pointer d_means, [k][d] float32, in/out), the new
    centroid matrix and the convergence flag of k_means.go:102-108.  Returns (compute.Matrix, converged)."""
    from .compute import Matrix
    ctx = ctx or default_context()
    h = C.c_void_p()
    conv = C.c_int(0)
    _check(centroid_matrix._L.vs_kmeans_finish_dev(ctx.handle, centroid_matrix.handle, C.c_void_p(int(d_sums)),
                                                   C.c_void_p(int(d_counts)), C.c_void_p(int(d_means)), C.byref(h), C.byref(conv)))
    return Matrix(h, centroid_matrix._L), bool(conv.value)


def Recenter(matrix, ctx=None):
    """recenterDbCentroid's arithmetic (dnc.go:417-449) over all rows of `matrix` -> row776."""
    ctx = ctx or default_context()
    out = np.empty(8 + matrix.cols, np.uint8)
    _check(matrix._L.vs_recenter(ctx.handle, matrix.handle, _p(out)))
    return out


def _sample(X, sample_size, rng, ctx):
    """sample() (dnc/sampling.go:12-74): all rows when there are at most sample_size, else sample_size distinct random rows
    in ascending order."""
    if X.rows <= sample_size:
        return X
    rows = np.sort(rng.choice(X.rows, sample_size, replace=False)).astype(np.uint64)
    from .compute import Matrix
    h = C.c_void_p()
    _check(X._L.vs_matrix_gather(ctx.handle, X.handle, _p(rows), rows.shape[0], C.byref(h)))
    return Matrix(h, X._L)


def Split(X, centroids, ctx=None):
    """The split loop of divideNconquer (dnc.go:363-389): child j = the rows of X nearest to centroid j, in X's order.
    Returns a list of compute.Matrix (None for a child without rows)."""
    from .compute import Matrix
    ctx = ctx or default_context()
    cent = _rows_array(centroids)
    k = cent.shape[0]
    handles = (C.c_void_p * k)()
    counts = np.zeros(k, np.uint64)
    _check(X._L.vs_matrix_split(ctx.handle, X.handle, _p(cent), k, C.cast(handles, C.c_void_p), _p(counts)))
    return [Matrix(C.c_void_p(handles[j]), X._L) if handles[j] else None for j in range(k)]


def DivideAndConquer(data_matrix, target_size=CENTROID_SIZE, sample_size=SAMPLE_SIZE, split_size=SPLIT_SIZE, rng=None,
                     iter_limit=KMEANS_ITTERATION_LIMIT, ctx=None, workers=8, stats=None):
    """divideNconquer (dnc/dnc.go:300-400) with every dataset resident in HBM instead of a temp file: a set of at most
    target_size rows yields one centroid, kMeans(sample, 1)[0] (dataset.go:93-98); a larger one is split by the
    min(split_size, max(2, rows / target_size)) centroids of kMeans(sample) and its children are treated the same way.

    Like the reference (one goroutine per child behind a semaphore, dnc.go:30-33,346-358) the subproblems run concurrently:
    `workers` host threads, each with its own compute.Context (a CUDA stream + scratch), take nodes from one queue, so the
    latency-bound Lloyd iterations of different nodes overlap on the device.  The reference seeds every draw from the
    clock; here every node owns a generator -- the root's is `rng`, a node hands its i-th non-empty child the i-th
    generator of `rng.spawn()` after its own draws (sample, then the k-means superset) -- so the result does not depend on
    the schedule or on `workers`, and the leaves are returned in depth-first order, children in index order.  A child that
    receives no row is skipped (the reference would index an empty slice).  Returns the leaf centroids, (m, 8+d) uint8.
    `stats`: optional dict that receives where the time went (host seconds per phase summed over the workers, Lloyd
    iterations, device microseconds of the assign / update halves of the iterations).
    """
    import queue
    import threading
    import time
    from .compute import Context
    ctx = ctx or default_context()
    rng = rng or np.random.default_rng()
    tally = threading.Lock()

    def note(**kw):
        if stats is not None:
            with tally:
                for key, v in kw.items():
                    stats[key] = stats.get(key, 0) + v

    def km(S, k, g, c):
        if stats is None or S.rows <= k:
            return KMeans(S, k, rng=g, iter_limit=iter_limit, ctx=c)
        out, st = KMeans(S, k, rng=g, iter_limit=iter_limit, ctx=c, want_stats=True)
        note(kmeans_calls=1, iterations=st["superset_iterations"] + st["set_iterations"], assign_us=st["assign_us"],
             update_us=st["update_us"])
        return out

    def node(X, g, c):
        """-> (centroid, None) for a leaf, (None, [(child matrix, child generator)]) for a split."""
        t0 = time.perf_counter()
        S = _sample(X, sample_size, g, c)
        t1 = time.perf_counter()
        if X.rows <= target_size:                                          # dnc.go:316-319
            cent = km(S, 1, g, c)[0]
            note(sample_s=t1 - t0, leaf_kmeans_s=time.perf_counter() - t1, leaves=1)
            return cent, None
        k = min(split_size, max(2, X.rows // target_size))                 # dnc.go:330-339
        cents = km(S, k, g, c)
        del S
        t2 = time.perf_counter()
        kids = [ch for ch in Split(X, cents, ctx=c) if ch is not None]
        note(sample_s=t1 - t0, split_kmeans_s=t2 - t1, split_s=time.perf_counter() - t2, splits=1)
        return None, list(zip(kids, g.spawn(len(kids))))

    leaves = {}                                                            # path in the tree -> leaf centroid
    if workers <= 1:
        stack = [(data_matrix, rng, ())]
        while stack:
            X, g, path = stack.pop()
            cent, kids = node(X, g, ctx)
            t0 = time.perf_counter()
            del X
            note(release_s=time.perf_counter() - t0)
            if kids is None:
                leaves[path] = cent
            else:
                stack.extend((ch, cg, path + (j,)) for j, (ch, cg) in reversed(list(enumerate(kids))))
        return np.stack([leaves[p] for p in sorted(leaves)])

    ctx.sync()                                                             # the rows may still be on their way on ctx's stream
    todo = queue.Queue()
    lock = threading.Lock()
    state = {"open": 1, "error": None}
    todo.put((data_matrix, rng, ()))

    def work():
        c = Context()
        try:
            while True:
                item = todo.get()
                if item is None:
                    return
                X, g, path = item
                del item
                try:
                    if state["error"] is not None:
                        cent, kids = None, []
                    else:
                        cent, kids = node(X, g, c)
                except BaseException as e:                                  # noqa: BLE001 -- handed to the caller below
                    with lock:
                        state["error"] = state["error"] or e
                    cent, kids = None, []
                t0 = time.perf_counter()
                del X
                note(release_s=time.perf_counter() - t0)
                with lock:
                    if kids is None:
                        leaves[path] = cent
                    else:
                        for j, (ch, cg) in enumerate(kids):
                            todo.put((ch, cg, path + (j,)))
                        state["open"] += len(kids)
                    state["open"] -= 1
                    if state["open"] == 0:
                        for _ in range(workers):
                            todo.put(None)
        finally:
            c.close()

    threads = [threading.Thread(target=work, name=f"dnc-{i}") for i in range(workers)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if state["error"] is not None:
        raise state["error"]
    return np.stack([leaves[p] for p in sorted(leaves)])


def ReassignRecenter(data_matrix, centroids, d_assign=None, want_assign=True, ctx=None):
    """The tail of KMeansDivideAndConquer (dnc.go:177-291): every row to its nearest new centroid, then every centroid
    re-centred on its members (float64 mean in row order, QuantizeVectorFloat64).  dropSmallCentroids (dnc.go:458-574)
    sits between the two upstream and never drops anything: its list is sorted ascending and its loop stops at the
    first cluster below CENTROID_SIZE/10, which can only be index 0, so `oldCentroids = results[:0]` is always empty.
    Returns (assign int32[n] or None, recentred centroids (k, 8+d) uint8, counts int64[k]); d_assign: optional raw device
    pointer that also receives the assignment (for vs_index_build_dev)."""
    ctx = ctx or default_context()
    cent = _rows_array(centroids)
    k, rb = cent.shape
    n = data_matrix.rows
    assign = np.empty(n, np.int32) if want_assign else None
    out = np.empty((k, rb), np.uint8)
    counts = np.empty(k, np.int64)
    _check(data_matrix._L.vs_reassign_recenter(ctx.handle, data_matrix.handle, _p(cent), k, _p(assign) if want_assign else None,
                                               C.c_void_p(int(d_assign)) if d_assign else None, _p(out), _p(counts)))
    return assign, out, counts
