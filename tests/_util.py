"""Seeded synthetic inputs shared by the parity tests (unit-norm N(0,I) rows)."""
import numpy as np


def unit_rows(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


def noop_rows(n, d, seed):
    """noop/ai.go:47-64: header -1/+1, uniform random code bytes."""
    rng = np.random.default_rng(seed)
    rows = np.empty((n, 8 + d), np.uint8)
    rows[:, 0:4] = np.frombuffer(np.float32(-1).tobytes(), np.uint8)
    rows[:, 4:8] = np.frombuffer(np.float32(1).tobytes(), np.uint8)
    rows[:, 8:] = rng.integers(0, 256, (n, d), dtype=np.uint8)
    return rows


def clustered_rows(n, d, k, seed, spread=0.35):
    """Rows around k random directions, so IVF lists and k-means clusters are meaningful."""
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((k, d))
    centers /= np.linalg.norm(centers, axis=1, keepdims=True)
    which = rng.integers(0, k, n)
    x = centers[which] + spread * rng.standard_normal((n, d)) / np.sqrt(d) * np.sqrt(d) * 0.05
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32), which


def f32_bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_float(a, b):
    """Bit-equal, except that any NaN matches any NaN (x86 and CUDA produce different NaN payloads)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape
    u = np.uint32 if a.dtype == np.float32 else np.uint64
    return bool(((a.view(u) == b.view(u)) | (np.isnan(a) & np.isnan(b))).all())
