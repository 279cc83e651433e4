"""Timing PROXY for the reference's `gonum` build-tag backend -- TEST / BENCH INFRASTRUCTURE ONLY.

compute/cosine_gonum.go scores with gonum's BLAS (gonum.org/v1/gonum v0.16.0, go.mod:11: Dnrm2, Dscal, Ddot in amd64
assembly).  That library is not under /root/reference and Go is not in the image, so its summation order cannot be
restated bit for bit; this module restates the same sequence of BLAS calls through numpy's OpenBLAS instead:

  compute_gonum.go:10-44    NewVector / NewMatrix     dequantize every row to float64
  cosine_gonum.go:129-149   normalizeMatrixRows/Vector  Dnrm2, then Dscal(1/norm) when norm != 0
  cosine_gonum.go:36-38     scores[i] = Ddot(B[i], A)   (one matrix-vector product here)
  server/search.go:214-273  probe cut, 1000-row batches, sort + dedup + truncate

It answers "how fast is a vectorised float64 CPU path on these host cores", next to the scalar default backend that
oracle.c restates exactly; its float32 similarities stay within north_star's 1e-6 relative of the default backend's
(tests/test_oracle_golden.py), so its top-k can differ only where two scores tie or straddle a float32 rounding
boundary.  Parity claims are never anchored on this module.
"""
import numpy as np

BATCH_SIZE_DATABASE = 1000  # config/constants.go:6


def new_matrix(rows):
    """compute_gonum.go:23-44: [][]uint8 -> row-major float64."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    hdr = rows[:, :8].copy().view(np.float32).astype(np.float64)   # quantization.go:124-132: float32 header widened
    mn, mx = hdr[:, 0:1], hdr[:, 1:2]
    with np.errstate(all="ignore"):
        return mn + (rows[:, 8:].astype(np.float64) / 255.0) * (mx - mn)


def normalize_rows(B):
    """cosine_gonum.go:129-141: Dnrm2 per row, Dscal(1/norm) unless the norm is zero."""
    norm = np.sqrt(np.einsum("ij,ij->i", B, B))
    with np.errstate(all="ignore"):
        inv = np.where(norm != 0, 1.0 / norm, 1.0)
    B *= inv[:, None]
    return B


def cosine_1xN(A_normalized, rows):
    """cosine_gonum.go:13-48 for a query already normalized (the reference re-normalizes its clone every call)."""
    B = normalize_rows(new_matrix(rows))
    return (B @ A_normalized).astype(np.float32)


def _order_desc(sims):
    """slices.SortFunc(cmp.Compare(b, a)): descending, NaN last; stable here (upstream's sort is not)."""
    key = np.where(np.isnan(sims), -np.inf, sims.astype(np.float64))
    return np.argsort(-key, kind="stable")


def search(q, centroids, rows, list_of_row, doc_ids, nprobe, k):
    """server/search.go:202-273 with the gonum backend's arithmetic; rows in primary-key order."""
    A = normalize_rows(new_matrix(np.asarray(q, np.uint8)[None, :]))[0]
    csims = cosine_1xN(A, centroids)
    probes = _order_desc(csims)[:min(nprobe, len(csims))]
    sel = np.flatnonzero(np.isin(list_of_row, probes))
    best_ids = np.empty(0, np.uint64)
    best_sims = np.empty(0, np.float32)
    for b0 in range(0, sel.shape[0], BATCH_SIZE_DATABASE):
        idx = sel[b0:b0 + BATCH_SIZE_DATABASE]
        sims = cosine_1xN(A, rows[idx])
        ids = doc_ids[idx] if doc_ids is not None else idx.astype(np.uint64)
        best_ids = np.concatenate([best_ids, ids])
        best_sims = np.concatenate([best_sims, sims])
        order = _order_desc(best_sims)                       # sort by nearest
        best_ids, best_sims = best_ids[order], best_sims[order]
        _, first = np.unique(best_ids, return_index=True)    # dedup keeping the best-ranked entry per document
        keep = np.sort(first)[:k]                            # truncate to Count+Offset
        best_ids, best_sims = best_ids[keep], best_sims[keep]
    return best_ids, best_sims
