"""GPU parity: compute package (quantization, NewMatrix, 1xN cosine, MxN argmax) through the C ABI,
bit-exact against the CPU oracle on the same seeded inputs."""
import json
import os
import struct
import threading

import numpy as np
import pytest

from _util import f32_bits, noop_rows, same_float, unit_rows

pytestmark = pytest.mark.gpu
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


def _inp(v):
    return [float("nan") if x == "nan" else x for x in v]


def test_quantize_kat(vs):
    for case in KAT["quantize_f32"]:
        got = vs.compute.QuantizeVectorFloat32(np.array(_inp(case["input"]), np.float32))
        assert got.tolist() == case["expect"], case["name"]
    for case in KAT["quantize_f64"]:
        got = vs.compute.QuantizeVectorFloat64(np.array(_inp(case["input"]), np.float64))
        assert got.tolist() == case["expect"], case["name"]


def test_dequantize_kat(vs):
    for case in KAT["dequantize"]:
        row = np.array(case["row"], np.uint8)
        assert [struct.pack(">d", x).hex() for x in vs.compute.DequantizeVectorFloat64(row)] == case["f64"]
        assert [struct.pack(">f", x).hex() for x in vs.compute.DequantizeVectorFloat32(row)] == case["f32"]


@pytest.mark.parametrize("d", [1, 3, 5, 64, 100, 512, 768, 1024])
def test_quantize_random_parity(vs, oracle, d):
    rng = np.random.default_rng(d)
    x = (rng.standard_normal((257, d)) * rng.uniform(1e-3, 1e3, (257, 1))).astype(np.float32)
    x[3] = np.abs(x[3])
    x[4] = 0
    x[5] = -np.abs(x[5])
    if d > 2:
        x[6, 1] = np.nan
        x[7, 0] = np.inf
    assert (vs.compute.QuantizeMatrixFloat32(x) == oracle.quantize_matrix_f32(x)).all()
    x64 = rng.standard_normal((129, d)) * 1e-2
    x64[2] = 0
    assert (vs.compute.QuantizeMatrixFloat64(x64) == oracle.quantize_matrix_f64(x64)).all()
    rows = oracle.quantize_matrix_f32(x)
    assert same_float(vs.compute.DequantizeMatrixFloat32(rows), oracle.dequantize_matrix_f32(rows))
    assert same_float(vs.compute.DequantizeMatrixFloat64(rows), oracle.dequantize_matrix_f64(rows))


def test_quantize_unit_rows_768(vs, oracle):
    x = unit_rows(5000, 768, 11)
    assert (vs.compute.QuantizeMatrixFloat32(x) == oracle.quantize_matrix_f32(x)).all()


@pytest.mark.parametrize("d", [5, 100, 512, 768])
def test_matrix_roundtrip(vs, d):
    rows = noop_rows(1000, d, d)
    m = vs.compute.NewMatrix(rows)
    assert m.rows == 1000 and m.cols == d
    assert (m.ReadRows() == rows).all()
    assert (m.Clone().ReadRows(10, 5) == rows[10:15]).all()


def test_cosine_kat(vs):
    c = KAT["cosine"]
    q = vs.compute.NewVector(np.array(c["query"], np.uint8))
    m = vs.compute.NewMatrix(np.array(c["rows"], np.uint8))
    sims = q.Clone().MatrixCosineSimilarity(m)
    assert [struct.pack(">f", x).hex() for x in sims] == c["sims_f32"], c["why"]
    assert q.IntegerDots(m).tolist() == c["dots"]


def _adversarial(n, d, seed):
    """noop header with codes near mid-scale: tiny norms, heavy cancellation -> the certified path must
    hand these to the literal-arithmetic kernel."""
    rng = np.random.default_rng(seed)
    rows = noop_rows(n, d, seed)
    rows[:, 8:] = rng.integers(126, 130, (n, d), dtype=np.uint8)
    return rows


@pytest.mark.parametrize("d,n", [(768, 20000), (512, 5000), (1024, 3000), (384, 3000), (1536, 2000), (100, 3000), (5, 500)])
def test_cosine_1xN_bit_exact(vs, oracle, d, n):
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 100 + d))
    rows[7, :8] = 0                     # zero vector: header 0/0
    rows[9] = rows[11]                  # duplicate
    m = vs.compute.NewMatrix(rows)
    for qi in range(3):
        qrow = oracle.quantize_vector_f32(unit_rows(1, d, 7 + qi)[0]) if qi else rows[11]
        q = vs.compute.NewVector(qrow)
        assert (q.IntegerDots(m) == oracle.dot_u8_1xN(qrow, rows)).all()          # integer dots: bit-exact
        got, want = q.MatrixCosineSimilarity(m), oracle.cosine_1xN(qrow, rows)
        assert (f32_bits(got) == f32_bits(want)).all()                            # float32 sims: bit-exact
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9)                # north_star's stated tolerance


def test_cosine_noop_and_adversarial_rows(vs, oracle):
    ctx = vs.compute.Context()
    for rows in (noop_rows(4000, 512, 5), _adversarial(2000, 768, 6)):
        m = vs.compute.NewMatrix(rows)
        q = vs.compute.NewVector(rows[3])
        before = ctx.slowpath_count()
        got = q.MatrixCosineSimilarity(m, ctx=ctx)
        assert (f32_bits(got) == f32_bits(oracle.cosine_1xN(rows[3], rows))).all()
    assert ctx.slowpath_count() > before   # the adversarial rows went through the literal kernel
    ctx.close()


def test_cosine_errors(vs):
    m = vs.compute.NewMatrix(noop_rows(4, 16, 0))
    with pytest.raises(vs.compute.ComputeFatal, match="column size does not match"):
        vs.compute.NewVector(noop_rows(1, 12, 0)[0]).MatrixCosineSimilarity(m)   # cosine.go:19-21
    m2 = vs.compute.NewMatrix(noop_rows(4, 12, 0))
    with pytest.raises(vs.compute.ComputeFatal, match="column size does not match"):
        m.MatrixCosineSimilarity(m2)                                             # cosine.go:77-79
    with pytest.raises(vs.compute.ComputePanic):
        vs.compute.NewMatrix(np.zeros((3, 8), np.uint8))                         # compute.go:29-31


def test_argmax_kat(vs):
    c = KAT["argmax"]
    cent = vs.compute.NewMatrix(np.array(c["centroids"], np.uint8))
    data = vs.compute.NewMatrix(np.array(c["data"], np.uint8))
    _, idx = cent.Clone().MatrixCosineSimilarity(data.Clone())
    assert idx.tolist() == c["argmax"], c["why"]


@pytest.mark.parametrize("d,m,n", [(768, 5, 10000), (768, 25, 10000), (768, 40, 3000), (768, 100, 2000),
                                   (512, 25, 3000), (64, 7, 3000), (100, 3, 1000), (1024, 4, 1000)])
def test_argmax_parity(vs, oracle, d, m, n):
    data = oracle.quantize_matrix_f32(unit_rows(n, d, 31 + d + m))
    cent = data[np.random.default_rng(m).choice(n, m, replace=False)].copy()
    if m >= 5:
        cent[3] = cent[1]            # duplicate centroid: lowest index must win
        cent[4, :] = 0               # zero centroid (an empty cluster's mean)
    want_s, want_i = oracle.argmax_MxN(cent, data)
    calc, done = vs.compute.MatrixCosineSimilarity()
    got_s, got_i = calc(vs.compute.NewMatrix(cent).Clone(), vs.compute.NewMatrix(data).Clone())
    done()
    assert (got_i == want_i).all()
    assert (f32_bits(got_s) == f32_bits(want_s)).all()


def test_argmax_adversarial(vs, oracle):
    data = _adversarial(1500, 768, 1)
    cent = _adversarial(6, 768, 2)
    want_s, want_i = oracle.argmax_MxN(cent, data)
    got_s, got_i = vs.compute.NewMatrix(cent).MatrixCosineSimilarity(vs.compute.NewMatrix(data))
    assert (got_i == want_i).all() and (f32_bits(got_s) == f32_bits(want_s)).all()


def test_closures_concurrent(vs, oracle):
    """calculate closures are per goroutine and run concurrently (dnc/dnc.go:30-33, search.go:230)."""
    rows = oracle.quantize_matrix_f32(unit_rows(4000, 768, 77))
    m = vs.compute.NewMatrix(rows)
    want = oracle.cosine_1xN(rows[0], rows)
    errs = []

    def worker():
        calc, done = vs.compute.VectorMatrixCosineSimilarity()
        try:
            for _ in range(5):
                got = calc(vs.compute.NewVector(rows[0]).Clone(), m.Clone())
                if not (f32_bits(got) == f32_bits(want)).all():
                    errs.append("mismatch")
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))
        finally:
            done()

    ts = [threading.Thread(target=worker) for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


def test_row_spool_load_and_save(vs, oracle, tmp_path):
    """dnc/dataset.go's cache file: a flat file of 8+D-byte rows.  Load (whole file, ranges), save, append."""
    d, n = 768, 200_000                     # 155 MB: several 64 MB staging chunks
    rows = noop_rows(n, d, 9)
    rows[:, 0:8] = np.random.default_rng(1).integers(0, 256, (n, 8), dtype=np.uint8)   # arbitrary header bytes survive
    path = tmp_path / "1234.cache"
    rows.tofile(path)
    m = vs.compute.LoadSpool(path, d)
    assert m.rows == n and m.cols == d
    assert (m.ReadRows() == rows).all()
    part = vs.compute.LoadSpool(path, d, first_row=1000, count=5000)
    assert (part.ReadRows() == rows[1000:6000]).all()
    tail = vs.compute.LoadSpool(path, d, first_row=n - 10)
    assert (tail.ReadRows() == rows[n - 10:]).all()
    # the loaded matrix scores like one built from the rows
    q = vs.compute.NewVector(rows[7])
    assert (f32_bits(q.MatrixCosineSimilarity(part)) == f32_bits(oracle.cosine_1xN(rows[7], rows[1000:6000]))).all()
    out = tmp_path / "child.cache"
    m.SaveSpool(out, first=0, count=90_000)
    m.SaveSpool(out, first=90_000, append=True)
    assert (np.fromfile(out, np.uint8).reshape(n, 8 + d) == rows).all()
    with open(path, "ab") as f:
        f.write(b"\x00" * 5)                # a torn last row: dataset.ReadRow's io.ReadFull would fail
    with pytest.raises(vs.compute.ComputeError):
        vs.compute.LoadSpool(path, d)
    with pytest.raises(vs.compute.ComputePanic):
        vs.compute.LoadSpool(out, d, first_row=n)   # no rows: compute.go:25-27


def _near_midpoint_rows(oracle, qrow, d, n_total, keep, seed):
    """Adversarial search for rows the certified scorer could get wrong: out of n_total random rows, the `keep` whose
    float64 cosine with the query lies closest to a float32 rounding boundary (the midpoint of two adjacent float32
    values) -- there a float64 error of a few 1e-16 flips the float32 result, so the kernel must either have certified a
    half-width that really is smaller than the distance or have sent the row through the literal arithmetic."""
    def deq(rows):
        h = np.ascontiguousarray(rows[:, :8]).view(np.float32).astype(np.float64)
        return h[:, :1] + (h[:, 1:2] - h[:, :1]) * rows[:, 8:].astype(np.float64) / 255.0
    q = deq(qrow[None, :])[0]
    q /= np.linalg.norm(q)
    best_rows, best_dist = [], []
    for c in range(0, n_total, 20000):
        rows = oracle.quantize_matrix_f32(unit_rows(min(20000, n_total - c), d, seed + c))
        x = deq(rows)
        cos = (x @ q) / np.linalg.norm(x, axis=1)
        f = cos.astype(np.float32)
        up = np.nextafter(f, np.float32(2)).astype(np.float64)
        dn = np.nextafter(f, np.float32(-2)).astype(np.float64)
        f64 = f.astype(np.float64)
        dist = np.minimum(np.abs(cos - (f64 + up) / 2), np.abs(cos - (f64 + dn) / 2)) / (up - f64)   # in float32 steps
        idx = np.argsort(dist)[:keep]
        best_rows.append(rows[idx])
        best_dist.append(dist[idx])
    rows, dist = np.concatenate(best_rows), np.concatenate(best_dist)
    order = np.argsort(dist)[:keep]
    return rows[order], dist[order]


@pytest.mark.parametrize("d", [768, 384])
def test_cosine_near_float32_rounding_boundaries(vs, oracle, d):
    qrow = oracle.quantize_vector_f32(unit_rows(1, d, 4242)[0])
    rows, dist = _near_midpoint_rows(oracle, qrow, d, 200000, 2000, 77)
    assert dist[0] < 1e-4 and dist[-1] < 2e-2          # the set really hugs the boundaries (distances in float32 steps)
    ctx = vs.compute.Context()
    m = vs.compute.NewMatrix(rows, ctx=ctx)
    q = vs.compute.NewVector(qrow)
    got, want = q.MatrixCosineSimilarity(m, ctx=ctx), oracle.cosine_1xN(qrow, rows)
    bad = np.flatnonzero(f32_bits(got) != f32_bits(want))
    assert bad.size == 0, f"{bad.size} rows differ, e.g. row {bad[:3]}: distance {dist[bad[:3]]} float32 steps"
    # the same rows through the search path (top-k over a flat store: sort order by float32 similarity, ties by id)
    ids, sims, counts = vs.ivf.SearchFlat(m, qrow[None, :], 100, ctx=ctx)
    w_ids, w_sims = oracle.search_flat(qrow, rows, np.arange(len(rows), dtype=np.uint64), 100)
    assert ids[0, :counts[0]].tolist() == w_ids.tolist()
    assert (f32_bits(sims[0, :counts[0]]) == f32_bits(w_sims)).all()
    ctx.close()
