"""CPU: both oracles against the hand-derived golden vectors, and against each other.

The reference has no tests, so these known-answer vectors (tests/golden/kat.json, built by
tests/golden/make_golden.py without the oracle) are what pins oracle.c; oracle_np.py is a second
restatement that must agree bit-for-bit on random inputs.
"""
import json
import os
import struct

import numpy as np
import pytest

from _util import f32_bits, noop_rows, unit_rows

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


def _inp(v):
    return [float("nan") if x == "nan" else x for x in v]


@pytest.mark.parametrize("case", KAT["quantize_f32"], ids=lambda c: c["name"])
def test_quantize_f32_kat(oracle, case):
    from oracle import oracle_np as onp
    v = np.array(_inp(case["input"]), np.float32)
    want = np.array(case["expect"], np.uint8)
    assert (oracle.quantize_vector_f32(v) == want).all(), case["why"]
    assert (onp.quantize_vector(v, np.float32) == want).all(), case["why"]


@pytest.mark.parametrize("case", KAT["quantize_f64"], ids=lambda c: c["name"])
def test_quantize_f64_kat(oracle, case):
    from oracle import oracle_np as onp
    v = np.array(_inp(case["input"]), np.float64)
    want = np.array(case["expect"], np.uint8)
    assert (oracle.quantize_vector_f64(v) == want).all(), case["why"]
    assert (onp.quantize_vector(v, np.float64) == want).all(), case["why"]


def test_dequantize_kat(oracle):
    for case in KAT["dequantize"]:
        row = np.array(case["row"], np.uint8)[None, :]
        f64 = oracle.dequantize_matrix_f64(row)[0]
        f32 = oracle.dequantize_matrix_f32(row)[0]
        assert [struct.pack(">d", x).hex() for x in f64] == case["f64"]
        assert [struct.pack(">f", x).hex() for x in f32] == case["f32"]


def test_cosine_kat(oracle):
    from oracle import oracle_np as onp
    c = KAT["cosine"]
    q = np.array(c["query"], np.uint8)
    rows = np.array(c["rows"], np.uint8)
    sims = oracle.cosine_1xN(q, rows)
    assert [struct.pack(">f", x).hex() for x in sims] == c["sims_f32"], c["why"]
    assert [struct.pack(">f", x).hex() for x in onp.cosine_1xN(q, rows)] == c["sims_f32"]
    assert oracle.dot_u8_1xN(q, rows).tolist() == c["dots"]


def test_argmax_kat(oracle):
    from oracle import oracle_np as onp
    c = KAT["argmax"]
    cent = np.array(c["centroids"], np.uint8)
    data = np.array(c["data"], np.uint8)
    assert oracle.argmax_MxN(cent, data)[1].tolist() == c["argmax"], c["why"]
    assert onp.argmax_MxN(cent, data)[1].tolist() == c["argmax"]


def test_search_and_upload_kat(oracle):
    """server/search.go:202-273 and server/upload.go:239-279 on the hand-derived five-row store."""
    c = KAT["search"]
    cent = np.array(c["centroids"], np.uint8)
    rows = np.array(c["rows"], np.uint8)
    doc = np.array(c["doc_ids"], np.uint64)
    lists = np.array(c["lists"], np.uint32)
    q = np.array(c["query"], np.uint8)
    assert oracle.argmax_MxN(cent, rows)[1].tolist() == c["lists"]
    for s in c["searches"]:
        ids, sims = oracle.search(q, cent, rows, lists, doc, s["nprobe"], s["k"])
        assert ids.tolist() == s["ids"], s["why"]
        assert [struct.pack(">f", x).hex() for x in sims] == s["sims_f32"], s["why"]
    dd = c["dedup"]
    ids, sims = oracle.search(q, cent, rows, lists, np.array(dd["doc_ids"], np.uint64), dd["nprobe"], dd["k"])
    assert ids.tolist() == dd["ids"] and [struct.pack(">f", x).hex() for x in sims] == dd["sims_f32"], dd["why"]
    u = c["upload"]
    n0 = u["first"]
    assign, all_lists, all_rows, all_doc = oracle.upload(cent, rows[n0:], lists[:n0], rows[:n0], doc[:n0], doc[n0:])
    assert assign.tolist() == u["assign"], u["why"]
    assert all_lists.tolist() == c["lists"] and (all_rows == rows).all() and (all_doc == doc).all()


def test_kmeans_step_and_recenter_kat(oracle):
    """dnc/k_means.go:67-117 and dnc/dnc.go:417-449 on the hand-derived cases (plain Python arithmetic in make_golden.py)."""
    from oracle import oracle_np as onp
    c = KAT["kmeans_step"]
    cent = np.array(c["centroids"], np.uint8)
    data = np.array(c["data"], np.uint8)
    means = np.array(c["prev_means"], np.float32)
    assign, counts, newc, conv = oracle.kmeans_step(data, cent, means)
    assert assign.tolist() == c["assign"] and counts.tolist() == c["counts"], c["why"]
    assert [[struct.pack(">f", v).hex() for v in m] for m in means] == c["means_f32"], c["why"]
    assert newc.tolist() == c["new_centroids"] and conv == c["converged"], c["why"]
    counts2, means2, newc2 = onp.kmeans_update(data, np.array(c["assign"]), len(cent), np.array(c["prev_means"], np.float32))
    assert counts2.tolist() == c["counts"] and newc2.tolist() == c["new_centroids"]
    assert [[struct.pack(">f", v).hex() for v in m] for m in means2] == c["means_f32"]
    r = KAT["recenter"]
    rows = np.array(r["rows"], np.uint8)
    assert oracle.recenter(rows).tolist() == r["expect"], r["why"]
    assert onp.recenter(rows).tolist() == r["expect"]


def test_reference_panics(oracle):
    q = np.zeros(8, np.uint8)
    rows = np.zeros((2, 10), np.uint8)
    with pytest.raises(oracle.OraclePanic):
        oracle.cosine_1xN(q, rows)                      # compute.go:12-14 empty vector
    with pytest.raises(oracle.OraclePanic):
        oracle.cosine_1xN(np.zeros(10, np.uint8), np.zeros((0, 10), np.uint8))  # compute.go:25-27
    with pytest.raises(oracle.OraclePanic):
        oracle.cosine_1xN(np.zeros(12, np.uint8), rows)  # cosine.go:19-21 dimension mismatch


@pytest.mark.parametrize("d", [1, 3, 64, 768])
def test_two_restatements_agree(oracle, d):
    from oracle import oracle_np as onp
    rng = np.random.default_rng(d)
    x = rng.standard_normal((24, d)).astype(np.float32)
    x[3] = np.abs(x[3])
    x[4] = 0
    if d > 2:
        x[5, 1] = np.nan
    rows = oracle.quantize_matrix_f32(x)
    assert (rows == np.stack([onp.quantize_vector(r, np.float32) for r in x])).all()
    x64 = rng.standard_normal((8, d))
    assert (oracle.quantize_matrix_f64(x64) == np.stack([onp.quantize_vector(r, np.float64) for r in x64])).all()
    assert (f32_bits(oracle.dequantize_matrix_f32(rows)) ==
            f32_bits(np.stack([onp.dequantize_vector(r, np.float32) for r in rows]))).all()
    assert (oracle.dequantize_matrix_f64(rows).view(np.uint64) ==
            np.stack([onp.dequantize_vector(r, np.float64) for r in rows]).view(np.uint64)).all()
    assert (f32_bits(oracle.cosine_1xN(rows[0], rows)) == f32_bits(onp.cosine_1xN(rows[0], rows))).all()
    s1, i1 = oracle.argmax_MxN(rows[:5], rows)
    s2, i2 = onp.argmax_MxN(rows[:5], rows)
    assert (i1 == i2).all() and (f32_bits(s1) == f32_bits(s2)).all()
    means = np.zeros((5, d), np.float32)
    a, c, nc, _ = oracle.kmeans_step(rows, rows[:5], means)
    c2, m2, nc2 = onp.kmeans_update(rows, a, 5, np.zeros((5, d), np.float32))
    assert (c == c2).all() and (nc == nc2).all() and (f32_bits(means) == f32_bits(m2)).all()
    assert (oracle.recenter(rows) == onp.recenter(rows)).all()


def test_search_oracle_properties(oracle):
    """ora_search (batched running sort+dedup+truncate, search.go:239-273) == global sort of the probed rows."""
    d, n, C = 32, 3000, 12
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 1))
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, 2))
    _, lists = oracle.argmax_MxN(cent, rows)
    rng = np.random.default_rng(3)
    doc = rng.integers(0, n // 2, n).astype(np.uint64)   # several embeddings per document
    q = oracle.quantize_vector_f32(unit_rows(1, d, 4)[0])
    ids, sims = oracle.search(q, cent, rows, lists, doc, nprobe=4, k=15)
    probes, _ = oracle.select_probes(q, cent, 4)
    mask = np.isin(lists, probes)
    s = oracle.cosine_1xN(q, rows[mask])
    order = sorted(zip(-s.astype(np.float64), doc[mask].tolist()))
    seen, want = set(), []
    for negs, dd in order:
        if dd not in seen:
            seen.add(dd)
            want.append((dd, np.float32(-negs)))
        if len(want) == 15:
            break
    assert ids.tolist() == [w[0] for w in want]
    assert (f32_bits(sims) == f32_bits(np.array([w[1] for w in want], np.float32))).all()
    assert len(set(ids.tolist())) == len(ids)


def test_blas_proxy_build_within_tolerance(oracle):
    """The oracle's second build (reorderable, vectorized float64 sums: the timing proxy for the reference's gonum/BLAS
    backend) stays within north_star's 1e-6 relative of the exact build and returns the same hits on a store without
    near-ties.  It is never used as a parity anchor."""
    d, n, C = 768, 6000, 24
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 11))
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, 12))
    _, lists = oracle.argmax_MxN(cent, rows)
    lists = lists.astype(np.uint32)
    doc = np.random.default_rng(13).integers(0, n // 2, n).astype(np.uint64)
    qs = oracle.quantize_matrix_f32(unit_rows(3, d, 14))
    for q in qs:
        got, want = oracle.cosine_1xN(q, rows[:2000], blas_proxy=True), oracle.cosine_1xN(q, rows[:2000])
        assert np.max(np.abs(got.astype(np.float64) - want) / np.maximum(np.abs(want), 1e-30)) <= 1e-6
    ids, sims, counts = oracle.search_many(qs, cent, rows, lists, doc, 6, 10, threads=2, blas_proxy=True)
    w_ids, w_sims, w_counts = oracle.search_many(qs, cent, rows, lists, doc, 6, 10, threads=2)
    assert (counts == w_counts).all() and (ids == w_ids).all()
    assert np.allclose(sims, w_sims, rtol=1e-6, atol=0)


def test_noop_fixture_shape():
    r = noop_rows(4, 512, 0)
    assert r.shape == (4, 520) and struct.unpack("<ff", r[0, :8].tobytes()) == (-1.0, 1.0)
