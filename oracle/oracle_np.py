"""Second, independent restatement of the reference arithmetic in numpy -- TEST INFRASTRUCTURE ONLY.

Written from the cited reference lines without looking at oracle.c's control flow, so the two
can pin each other bit-for-bit (the reference itself ships no golden vectors: PARITY UNPINNED).
numpy elementwise ops are IEEE-754 round-to-nearest, never fused; sequential float64
accumulation is obtained from np.cumsum (a left-to-right running sum), not np.sum (pairwise).
"""
import numpy as np


def _go_u8(x):
    """Go uint8(float) on amd64: truncate via CVTTSx2SI, keep the low byte; NaN -> 0."""
    x = np.asarray(x)
    bad = ~np.isfinite(x)
    t = np.where(bad, 0, np.trunc(np.where(bad, 0, x)))
    return (t.astype(np.int64) & 0xFF).astype(np.uint8)


def range0(v):
    """compute/quantization.go:182-216: min/max seeded at 0; NaN ignored (comparisons false)."""
    mn = v.dtype.type(0)
    mx = v.dtype.type(0)
    for x in v:
        if x < mn:
            mn = x
        if x > mx:
            mx = x
    return mn, mx


def quantize_vector(v, dtype):
    """compute/quantization.go:82-102 (dtype float32 or float64)."""
    v = np.asarray(v, dtype=dtype)
    mn, mx = range0(v)
    out = np.empty(8 + v.shape[0], np.uint8)
    out[0:4] = np.frombuffer(np.float32(mn).tobytes(), np.uint8)
    out[4:8] = np.frombuffer(np.float32(mx).tobytes(), np.uint8)
    with np.errstate(all="ignore"):
        c = np.clip(v, mn, mx) if v.size else v  # :22-26 (NaN passes through np.clip unchanged)
        normalized = (c - mn) / (mx - mn)        # :28
        out[8:] = _go_u8(normalized * dtype(255))  # :30
    return out


def dequantize_vector(row, dtype):
    """compute/quantization.go:114-132."""
    row = np.asarray(row, np.uint8)
    mn = dtype(np.frombuffer(row[0:4].tobytes(), np.float32)[0])
    mx = dtype(np.frombuffer(row[4:8].tobytes(), np.float32)[0])
    with np.errstate(all="ignore"):
        normalized = row[8:].astype(dtype) / dtype(255.0)  # :57 / :65
        scaled = normalized * (mx - mn)
        return mn + scaled                                  # :59 / :67


def _seqsum(x):
    # `var dot float64; dot += ...` starts from +0.0, so (+0.0) + (-0.0) = +0.0: seed the running sum.
    return np.cumsum(np.concatenate(([0.0], x)), dtype=np.float64)[-1]


def normalize(v):
    """compute/cosine.go:138-149."""
    norm = np.sqrt(_seqsum(v * v))
    return v / norm if norm != 0 else v


def cosine_1xN(q, rows):
    """compute/cosine.go:13-57 after compute.go:10-44."""
    A = normalize(dequantize_vector(q, np.float64))
    out = np.empty(len(rows), np.float32)
    with np.errstate(all="ignore"):
        for i, r in enumerate(rows):
            B = normalize(dequantize_vector(r, np.float64))
            out[i] = np.float32(_seqsum(A * B))
    return out


def argmax_MxN(cent, rows):
    """compute/cosine.go:70-125: strict '>' from maxVal=-1.0, maxIdx=0."""
    A = [normalize(dequantize_vector(c, np.float64)) for c in cent]
    idx = np.zeros(len(rows), np.int64)
    sims = np.zeros(len(rows), np.float32)
    with np.errstate(all="ignore"):
        for i, r in enumerate(rows):
            B = normalize(dequantize_vector(r, np.float64))
            best, bi = -1.0, 0
            for j, a in enumerate(A):
                d = _seqsum(a * B)
                if d > best:
                    best, bi = d, j
            idx[i] = bi
            sims[i] = np.float32(best)
    return sims, idx


def kmeans_update(data, assign, k, means):
    """dnc/k_means.go:80-99: float32 sums in row order, mean, requantize; empty keeps previous mean."""
    d = data.shape[1] - 8
    sums = np.zeros((k, d), np.float32)
    counts = np.zeros(k, np.int64)
    for i, c in enumerate(assign):
        sums[c] = sums[c] + dequantize_vector(data[i], np.float32)
        counts[c] += 1
    means = means.copy()
    for c in range(k):
        if counts[c] > 0:
            means[c] = sums[c] / np.float32(counts[c])
    newc = np.stack([quantize_vector(means[c], np.float32) for c in range(k)])
    return counts, means, newc


def recenter(rows):
    """dnc/dnc.go:417-449."""
    d = rows.shape[1] - 8
    s = np.zeros(d, np.float64)
    for r in rows:
        s = s + dequantize_vector(r, np.float64)
    with np.errstate(all="ignore"):
        s = s / np.float64(len(rows))
    return quantize_vector(s, np.float64)
