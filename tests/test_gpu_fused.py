"""GPU parity of the one-launch single-query search (csrc/fused.cu: probe stage, grid barrier, selection, TMA-ring list
scan, top-k) against the oracle's restatement of server/search.go:202-273, one query per call."""
import numpy as np
import pytest

from _util import f32_bits, noop_rows, unit_rows
from test_gpu_search import _crowded_inputs, _index_inputs

pytestmark = pytest.mark.gpu


def _check_one_by_one(oracle, ix, qs, cent, rows, lists, doc, nprobe, k, ctx=None):
    for i, q in enumerate(qs):
        ids, sims, counts = ix.Search(q[None, :], nprobe, k, ctx=ctx)
        want_ids, want_sims = oracle.search(q, cent, rows, lists, doc, nprobe, k)
        c = counts[0]
        assert c == len(want_ids), (i, c, len(want_ids))
        assert ids[0, :c].tolist() == want_ids.tolist(), f"query {i}"
        assert (f32_bits(sims[0, :c]) == f32_bits(want_sims)).all(), f"query {i}"


def test_fused_path_is_taken(vs, oracle):
    """One query = one kernel launch (plus the query's ingest); the two-launch path is what the test hook restores."""
    n, d, C = 20000, 768, 64
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 5)
    ctx = vs.compute.Context()
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent, ctx=ctx)
    q = oracle.quantize_matrix_f32(unit_rows(1, d, 9))
    ix.Search(q, 8, 10, ctx=ctx)
    l0 = ctx.launch_count()
    a = ix.Search(q, 8, 10, ctx=ctx)
    fused_launches = ctx.launch_count() - l0
    vs.compute.debug_set_fused(False)
    try:
        l0 = ctx.launch_count()
        b = ix.Search(q, 8, 10, ctx=ctx)
        two_stage_launches = ctx.launch_count() - l0
    finally:
        vs.compute.debug_set_fused(True)
    assert fused_launches < two_stage_launches
    assert (a[0] == b[0]).all() and (f32_bits(a[1]) == f32_bits(b[1])).all() and (a[2] == b[2]).all()
    ctx.close()


@pytest.mark.parametrize("nprobe,k", [(8, 10), (1, 10), (32, 20), (3, 40), (64, 100), (95, 128), (200, 10)])
def test_fused_ivf_parity(vs, oracle, nprobe, k):
    n, d, C = 30000, 768, 96
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 5)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = oracle.quantize_matrix_f32(unit_rows(5, d, 99))
    _check_one_by_one(oracle, ix, qs, cent, rows, lists, doc, nprobe, k)


@pytest.mark.parametrize("d", [384, 512, 1024, 1536, 100])
def test_fused_other_widths(vs, oracle, d):
    """Row widths with their own lane layout (384, 512, 1024, 1536) and one the fused kernel does not take (100)."""
    n, C = 9000, 40
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 50 + d)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = oracle.quantize_matrix_f32(unit_rows(3, d, 7))
    _check_one_by_one(oracle, ix, qs, cent, rows, lists, doc, 6, 10)
    _check_one_by_one(oracle, ix, qs, cent, rows, lists, doc, C, 10)


def test_fused_many_centroids(vs, oracle):
    """More centroids than one pass of the selection holds per thread, runs of identical centroids (ties by index)."""
    n, d, C = 6000, 384, 20000
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 1))
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, 2))
    cent[8000:8100] = cent[100]
    cent[7, :] = 0
    lists = (np.arange(n) * 7919 % C).astype(np.uint32)
    doc = np.arange(n, dtype=np.uint64)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = oracle.quantize_matrix_f32(unit_rows(3, d, 3))
    qs[0] = cent[100]
    _check_one_by_one(oracle, ix, qs, cent, rows, lists, doc, 48, 10)
    _check_one_by_one(oracle, ix, qs, cent, rows, lists, doc, 128, 10)


def test_fused_ragged_empty_and_tiny(vs, oracle):
    d, C = 768, 10
    rows = oracle.quantize_matrix_f32(unit_rows(1000, d, 2))
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, 3))
    sizes = [0, 1, 31, 32, 33, 0, 500, 7, 396, 0]
    lists = np.repeat(np.arange(C), sizes).astype(np.uint32)
    doc = np.arange(1000, dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    ix = vs.ivf.Index.build(rows, doc, offs, cent)
    qs = oracle.quantize_matrix_f32(unit_rows(4, d, 5))
    for nprobe, k in ((3, 10), (9, 10), (1, 5), (10, 10)):
        _check_one_by_one(oracle, ix, qs, cent, rows, lists, doc, nprobe, k)
    # fewer rows than k, flat
    few = oracle.quantize_matrix_f32(unit_rows(7, d, 2))
    m = vs.compute.NewMatrix(few)
    ids, sims, counts = vs.ivf.SearchFlat(m, qs[0], 10)
    want_ids, want_sims = oracle.search_flat(qs[0], few, None, 10)
    assert counts[0] == 7 and ids[0, :7].tolist() == want_ids.tolist()
    assert (f32_bits(sims[0, :7]) == f32_bits(want_sims)).all()


def test_fused_ties_and_zero_vectors(vs, oracle):
    n, d, C = 4000, 768, 8
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 33)
    rows[100:140] = rows[100]      # 40 identical rows with distinct ids
    rows[200:204, 8:] = 0          # all-zero codes
    rows[300, :8] = 0              # min = max = 0
    _, lists = oracle.argmax_MxN(cent, rows)
    lists = lists.astype(np.uint32)
    doc = np.random.default_rng(0).permutation(n).astype(np.uint64)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    zq = oracle.quantize_vector_f32(np.zeros(d, np.float32))
    qs = np.stack([rows[100], oracle.quantize_vector_f32(unit_rows(1, d, 1)[0]), zq])
    _check_one_by_one(oracle, ix, qs, cent, rows, lists, doc, 7, 25)
    _check_one_by_one(oracle, ix, qs, cent, rows, lists, doc, C, 25)


@pytest.mark.parametrize("k,crowd,ndocs", [(10, 40, 1), (32, 300, 3), (100, 500, 5)])
def test_fused_document_crowding(vs, oracle, k, crowd, ndocs):
    n, d, C = 12000, 256 + 128, 12
    rows, cent, lists, doc, q = _crowded_inputs(oracle, n, d, C, 300 + k, crowd, ndocs)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = np.stack([q, oracle.quantize_vector_f32(unit_rows(1, d, 5)[0])])
    for nprobe in (C, 5):
        _check_one_by_one(oracle, ix, qs, cent, rows, lists, doc, nprobe, k)


def test_fused_flat_100k(vs, oracle):
    """BASELINE config 1: brute force over 100k x 768, 1 query per call, top-10."""
    n, d = 100000, 768
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 1234))
    m = vs.compute.NewMatrix(rows)
    qs = oracle.quantize_matrix_f32(unit_rows(3, d, 4321))
    for q in qs:
        ids, sims, counts = vs.ivf.SearchFlat(m, q, 10)
        want_ids, want_sims = oracle.search_flat(q, rows, None, 10)
        assert ids[0].tolist() == want_ids.tolist()
        assert (f32_bits(sims[0]) == f32_bits(want_sims)).all()


def test_fused_uncertified_scores_go_to_the_literal_path(vs, oracle):
    """With the certification disabled (test hook) every window holds uncertified scores: the fused kernel must say so
    and the resolve path must return the oracle's bits."""
    n, d, C = 6000, 768, 24
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 15, docs_per=2)
    ctx = vs.compute.Context()
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent, ctx=ctx)
    qs = oracle.quantize_matrix_f32(unit_rows(3, d, 77))
    vs.compute.debug_set_certify_scale(1.0e7)
    try:
        _check_one_by_one(oracle, ix, qs, cent, rows, lists, doc, 5, 12, ctx=ctx)
        _check_one_by_one(oracle, ix, qs, cent, rows, lists, doc, C, 12, ctx=ctx)
    finally:
        vs.compute.debug_set_certify_scale(1.0)
    assert ctx.slowpath_count() > 0
    ctx.close()


def test_fused_repeated_launches_rearm(vs, oracle):
    """The grid barrier, ticket and uncertified-pair counters are re-armed by every launch: 200 back-to-back queries on
    one context, alternating probed and flat searches, all equal to the two-launch path."""
    n, d, C = 50000, 768, 128
    rows = noop_rows(n, d, 3)
    lists = (np.arange(n) % C).astype(np.uint32)
    ctx = vs.compute.Context()
    ix = vs.ivf.Index.build_assigned(rows, None, lists, rows[:C], ctx=ctx)
    qs = noop_rows(100, d, 4)
    got = [ix.Search(qs[i:i + 1], 16 if i % 2 else C, 10, ctx=ctx) for i in range(100)]
    vs.compute.debug_set_fused(False)
    try:
        want = [ix.Search(qs[i:i + 1], 16 if i % 2 else C, 10, ctx=ctx) for i in range(100)]
    finally:
        vs.compute.debug_set_fused(True)
    for g, w in zip(got, want):
        assert (g[0] == w[0]).all() and (f32_bits(g[1]) == f32_bits(w[1])).all() and (g[2] == w[2]).all()
    ctx.close()
