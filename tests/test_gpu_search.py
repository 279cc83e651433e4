"""GPU parity: IVF-Flat probe-and-scan (server/search.go:202-273) and flat scan, top-k ids bit-exact."""
import numpy as np
import pytest

from _util import f32_bits, noop_rows, unit_rows

pytestmark = pytest.mark.gpu


def _index_inputs(oracle, n, d, C, seed, docs_per=1):
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, seed))
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, seed + 1))
    _, lists = oracle.argmax_MxN(cent, rows)
    if docs_per == 1:
        doc = np.arange(n, dtype=np.uint64) + 1000
    else:
        doc = np.random.default_rng(seed + 2).integers(0, max(1, n // docs_per), n).astype(np.uint64)
    return rows, cent, lists.astype(np.uint32), doc


def _check(oracle, ix, qs, cent, rows, lists, doc, nprobe, k):
    ids, sims, counts = ix.Search(qs, nprobe, k)
    for i, q in enumerate(qs):
        want_ids, want_sims = oracle.search(q, cent, rows, lists, doc, nprobe, k)
        c = counts[i]
        assert c == len(want_ids), (i, c, len(want_ids))
        assert ids[i, :c].tolist() == want_ids.tolist(), f"query {i}"
        assert (f32_bits(sims[i, :c]) == f32_bits(want_sims)).all(), f"query {i}"


@pytest.mark.parametrize("nprobe,k", [(8, 10), (1, 10), (32, 20), (3, 40), (64, 100), (200, 10)])
def test_ivf_search_parity(vs, oracle, nprobe, k):
    n, d, C = 30000, 768, 96
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 5)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    assert ix.rows == n and ix.lists == C
    qs = oracle.quantize_matrix_f32(unit_rows(6, d, 99))
    _check(oracle, ix, qs, cent, rows, lists, doc, nprobe, k)


def test_ivf_probe_selection_parity(vs, oracle):
    n, d, C = 5000, 768, 300
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 8)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = oracle.quantize_matrix_f32(unit_rows(9, d, 3))
    for nprobe in (1, 32, 64, 100):
        probes, sims = ix.SelectProbes(qs, nprobe)
        for i, q in enumerate(qs):
            wp, ws = oracle.select_probes(q, cent, nprobe)
            assert probes[i].tolist() == wp.tolist()
            assert (f32_bits(sims[i]) == f32_bits(ws)).all()


def test_ivf_dedup_by_document(vs, oracle):
    """search.go:260-268: one hit per document, keeping its best embedding."""
    n, d, C = 8000, 256, 16
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 21, docs_per=3)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = oracle.quantize_matrix_f32(unit_rows(5, d, 4))
    _check(oracle, ix, qs, cent, rows, lists, doc, nprobe=6, k=15)


def test_ivf_ties_broken_by_id(vs, oracle):
    n, d, C = 4000, 768, 8
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 33)
    rows[100:140] = rows[100]      # 40 identical rows with distinct ids
    _, lists = oracle.argmax_MxN(cent, rows)
    lists = lists.astype(np.uint32)
    doc = np.random.default_rng(0).permutation(n).astype(np.uint64)   # ids not in row order
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = np.stack([rows[100], oracle.quantize_vector_f32(unit_rows(1, d, 1)[0])])
    _check(oracle, ix, qs, cent, rows, lists, doc, nprobe=8, k=25)


def test_ivf_ragged_and_empty_lists(vs, oracle):
    d, C = 768, 10
    rows = oracle.quantize_matrix_f32(unit_rows(1000, d, 2))
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, 3))
    sizes = [0, 1, 31, 32, 33, 0, 500, 7, 396, 0]
    lists = np.repeat(np.arange(C), sizes).astype(np.uint32)
    doc = np.arange(1000, dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    ix = vs.ivf.Index.build(rows, doc, offs, cent)     # rows already grouped
    qs = oracle.quantize_matrix_f32(unit_rows(4, d, 5))
    for nprobe, k in ((3, 10), (10, 10), (1, 5)):
        _check(oracle, ix, qs, cent, rows, lists, doc, nprobe, k)
    ix2 = vs.ivf.Index.build_assigned(rows[::-1].copy(), doc[::-1].copy(), lists[::-1].copy(), cent)
    _check(oracle, ix2, qs, cent, rows[::-1], lists[::-1], doc[::-1], 4, 10)


def test_fewer_rows_than_k(vs, oracle):
    d = 768
    rows = oracle.quantize_matrix_f32(unit_rows(7, d, 2))
    m = vs.compute.NewMatrix(rows)
    q = oracle.quantize_vector_f32(unit_rows(1, d, 9)[0])
    ids, sims, counts = vs.ivf.SearchFlat(m, q, 10)
    want_ids, want_sims = oracle.search_flat(q, rows, None, 10)
    assert counts[0] == 7 and ids[0, :7].tolist() == want_ids.tolist()
    assert (f32_bits(sims[0, :7]) == f32_bits(want_sims)).all()


@pytest.mark.parametrize("d,n", [(768, 100000), (512, 20000), (100, 5000)])
def test_flat_search_parity(vs, oracle, d, n):
    """BASELINE config 1: brute force over 100k x 768, 1 query, top-10."""
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 1234))
    m = vs.compute.NewMatrix(rows)
    qs = oracle.quantize_matrix_f32(unit_rows(3, d, 4321))
    ids, sims, counts = vs.ivf.SearchFlat(m, qs, 10)
    for i, q in enumerate(qs):
        want_ids, want_sims = oracle.search_flat(q, rows, None, 10)
        assert ids[i].tolist() == want_ids.tolist()
        assert (f32_bits(sims[i]) == f32_bits(want_sims)).all()


@pytest.fixture
def force_literal(vs):
    """Every certification fails -> all scores are resolved by the literal reference-arithmetic paths."""
    vs.compute.debug_set_certify_scale(1.0e7)
    yield
    vs.compute.debug_set_certify_scale(1.0)


def test_search_literal_path(vs, oracle, force_literal):
    """Rows whose float32 rounding cannot be certified must be resolved with reference arithmetic: with the
    certification disabled (test hook) the search must still return the oracle's bits."""
    d, n, C = 768, 3000, 6
    rng = np.random.default_rng(5)
    rows = noop_rows(n, d, 5)
    rows[:, 8:] = rng.integers(126, 130, (n, d), dtype=np.uint8)
    cent = rows[:C].copy()
    _, lists = oracle.argmax_MxN(cent, rows)
    lists = lists.astype(np.uint32)
    doc = np.arange(n, dtype=np.uint64)
    ctx = vs.compute.Context()
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent, ctx=ctx)
    qs = rows[10:13]
    ids, sims, counts = ix.Search(qs, 3, 10, ctx=ctx)
    assert ctx.slowpath_count() > 0
    for i, q in enumerate(qs):
        want_ids, want_sims = oracle.search(q, cent, rows, lists, doc, 3, 10)
        assert ids[i, :counts[i]].tolist() == want_ids.tolist()
        assert (f32_bits(sims[i, :counts[i]]) == f32_bits(want_sims)).all()
    ctx.close()


def test_search_literal_path_unit_rows(vs, oracle, force_literal):
    n, d, C = 6000, 768, 24
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 15, docs_per=2)
    ctx = vs.compute.Context()
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent, ctx=ctx)
    qs = oracle.quantize_matrix_f32(unit_rows(4, d, 77))
    ids, sims, counts = ix.Search(qs, 5, 12, ctx=ctx)
    assert ctx.slowpath_count() > 0
    for i, q in enumerate(qs):
        want_ids, want_sims = oracle.search(q, cent, rows, lists, doc, 5, 12)
        assert ids[i, :counts[i]].tolist() == want_ids.tolist()
        assert (f32_bits(sims[i, :counts[i]]) == f32_bits(want_sims)).all()
    ctx.close()


def test_cosine_and_argmax_literal_path(vs, oracle, force_literal):
    d, n, m = 768, 1500, 7
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 3))
    cent = oracle.quantize_matrix_f32(unit_rows(m, d, 4))
    ctx = vs.compute.Context()
    M = vs.compute.NewMatrix(rows, ctx=ctx)
    got = vs.compute.NewVector(cent[0]).MatrixCosineSimilarity(M, ctx=ctx)
    assert (f32_bits(got) == f32_bits(oracle.cosine_1xN(cent[0], rows))).all()
    sims, idx = vs.compute.NewMatrix(cent, ctx=ctx).MatrixCosineSimilarity(M, ctx=ctx)
    want_sims, want_idx = oracle.argmax_MxN(cent, rows)
    assert (idx == want_idx).all()
    assert ctx.slowpath_count() >= n
    ctx.close()


def test_full_size_properties(vs):
    """Size-independent properties at a larger size than the oracle handles quickly: the flat top-k
    must be sorted, unique, idempotent, and equal to the IVF search that probes every list."""
    d, n, C = 768, 400000, 64
    rows = noop_rows(n, d, 17)
    m = vs.compute.NewMatrix(rows)
    q = noop_rows(4, d, 18)
    ids, sims, counts = vs.ivf.SearchFlat(m, q, 10)
    ids2, sims2, _ = vs.ivf.SearchFlat(m, q, 10)
    assert (ids == ids2).all() and (f32_bits(sims) == f32_bits(sims2)).all()
    assert (counts == 10).all()
    for i in range(4):
        assert all(sims[i, j] > sims[i, j + 1] or (sims[i, j] == sims[i, j + 1] and ids[i, j] < ids[i, j + 1])
                   for j in range(9))
        full = vs.compute.NewVector(q[i]).MatrixCosineSimilarity(m)
        order = np.lexsort((np.arange(n), -full.astype(np.float64)))[:10]
        assert ids[i].tolist() == order.tolist()
        assert (f32_bits(sims[i]) == f32_bits(full[order])).all()
    lists = (np.arange(n) % C).astype(np.uint32)
    ix = vs.ivf.Index.build_assigned(rows, None, lists, rows[:C])
    ids3, sims3, _ = ix.Search(q, C, 10)
    assert (ids3 == ids).all() and (f32_bits(sims3) == f32_bits(sims)).all()


@pytest.mark.parametrize("G", [2, 4])
def test_striped_shards_merge_matches_full(vs, oracle, G):
    """Multi-GPU path emulated on one GPU: G row-striped shard indexes, shard-local top-k, device merge
    (vs_topk_merge_dev) == the oracle's search over the whole store (the ranks of an N-GPU run do exactly this,
    with an NCCL all-gather between the two steps)."""
    import torch
    n, d, C, nprobe, k, nq = 20000, 768, 40, 6, 10, 7
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 71, docs_per=2)
    qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 72))
    dev = torch.device("cuda", 0)
    g_ids = torch.zeros((G, nq, k), dtype=torch.int64, device=dev)
    g_sims = torch.zeros((G, nq, k), dtype=torch.float32, device=dev)
    g_counts = torch.zeros((G, nq), dtype=torch.int32, device=dev)
    for r in range(G):
        mine = vs.shard.stripe(n, r, G)
        ix = vs.ivf.Index.build_assigned(rows[mine], doc[mine], lists[mine], cent)
        ids, sims, counts = ix.Search(qs, nprobe, k)
        g_ids[r] = torch.from_numpy(ids.view(np.int64)).to(dev)
        g_sims[r] = torch.from_numpy(sims).to(dev)
        g_counts[r] = torch.from_numpy(counts).to(dev)
    out_ids = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    out_sims = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    out_counts = torch.zeros(nq, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx = vs.compute.default_context()
    vs.shard.merge_hits_dev(g_ids, g_sims, g_counts, k, out_ids, out_sims, out_counts, ctx=ctx)
    ctx.sync()
    ids = out_ids.cpu().numpy().view(np.uint64)
    sims = out_sims.cpu().numpy()
    counts = out_counts.cpu().numpy()
    for i in range(nq):
        wi, ws = oracle.search(qs[i], cent, rows, lists, doc, nprobe, k)
        assert counts[i] == len(wi)
        assert ids[i, :counts[i]].tolist() == wi.tolist()
        assert (f32_bits(sims[i, :counts[i]]) == f32_bits(ws)).all()


# ---- batches of >= 8 queries: the probe stage reads the centroid table once (probe.cu) -------------------------
@pytest.mark.parametrize("d,C,nq,nprobe", [(768, 300, 40, 32), (768, 96, 9, 1), (512, 1000, 70, 100), (256, 64, 33, 64),
                                           (1024, 40, 8, 7)])
def test_batched_probe_selection_parity(vs, oracle, d, C, nq, nprobe):
    n = 4000
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 40 + C)
    cent[5] = cent[2]                       # identical centroids: equal similarities, lower index first
    cent[7, :] = 0                          # zero centroid
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    q = unit_rows(nq, d, 7)
    q[3] = 0
    qs = oracle.quantize_matrix_f32(q)
    qs[4] = cent[2]
    probes, sims = ix.SelectProbes(qs, nprobe)
    for i in range(nq):
        wp, ws = oracle.select_probes(qs[i], cent, nprobe)
        assert probes[i].tolist() == wp.tolist(), f"query {i}"
        assert (f32_bits(sims[i]) == f32_bits(ws)).all(), f"query {i}"


def test_batched_probe_many_centroids(vs, oracle):
    """More than 8192 centroids: the select kernel reads its key slice from memory instead of registers."""
    n, d, C, nq, nprobe = 2000, 256, 9000, 12, 48
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 1))
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, 2))
    cent[8000:8100] = cent[100]             # a run of ties far apart in index
    lists = (np.arange(n) % C).astype(np.uint32)
    doc = np.arange(n, dtype=np.uint64)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 3))
    qs[0] = cent[100]
    probes, sims = ix.SelectProbes(qs, nprobe)
    for i in range(nq):
        wp, ws = oracle.select_probes(qs[i], cent, nprobe)
        assert probes[i].tolist() == wp.tolist(), f"query {i}"
        assert (f32_bits(sims[i]) == f32_bits(ws)).all()


def test_batched_search_parity_and_small_call_agreement(vs, oracle):
    n, d, C, nq, nprobe, k = 30000, 768, 128, 24, 16, 10
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 61, docs_per=2)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 62))
    _check(oracle, ix, qs, cent, rows, lists, doc, nprobe, k)
    ids, sims, counts = ix.Search(qs, nprobe, k)                           # batched probe stage
    parts = [ix.Search(qs[i:i + 4], nprobe, k) for i in range(0, nq, 4)]   # per-query streaming probe stage
    assert (ids == np.concatenate([p[0] for p in parts])).all()
    assert (f32_bits(sims) == f32_bits(np.concatenate([p[1] for p in parts]))).all()


def test_batched_probe_literal_paths(vs, oracle):
    """Adversarial centroids (tiny norms: many uncertified pairs, re-scored inside the select kernel) and the
    certification disabled altogether (more uncertified pairs than the per-query list holds: the caller's path)."""
    d, n, C, nq = 768, 3000, 200, 16
    rng = np.random.default_rng(5)
    rows = noop_rows(n, d, 5)
    rows[:, 8:] = rng.integers(126, 130, (n, d), dtype=np.uint8)
    cent = rows[:C].copy()
    cent[::3] = noop_rows((C + 2) // 3, d, 6)      # a mix of well- and ill-conditioned centroids
    _, lists = oracle.argmax_MxN(cent, rows)
    lists = lists.astype(np.uint32)
    doc = np.arange(n, dtype=np.uint64)
    ctx = vs.compute.Context()
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent, ctx=ctx)
    qs = np.concatenate([rows[10:10 + nq // 2], noop_rows(nq // 2, d, 8)])
    for scale in (1.0, 1.0e7):
        vs.compute.debug_set_certify_scale(scale)
        try:
            probes, sims = ix.SelectProbes(qs, 20, ctx=ctx)
            ids, hs, counts = ix.Search(qs, 20, 10, ctx=ctx)
        finally:
            vs.compute.debug_set_certify_scale(1.0)
        for i in range(nq):
            wp, ws = oracle.select_probes(qs[i], cent, 20)
            assert probes[i].tolist() == wp.tolist(), f"scale {scale} query {i}"
            assert (f32_bits(sims[i]) == f32_bits(ws)).all()
            want_ids, want_sims = oracle.search(qs[i], cent, rows, lists, doc, 20, 10)
            assert ids[i, :counts[i]].tolist() == want_ids.tolist()
            assert (f32_bits(hs[i, :counts[i]]) == f32_bits(want_sims)).all()
    assert ctx.slowpath_count() > 0
    ctx.close()


def test_probe_selection_on_tensor_cores(vs, oracle):
    """>= 16384 centroids and >= 64 queries: the probe stage runs through the int8 GEMM pipeline with k = nprobe
    (sampled thresholds, fused filter, certified selection); same probe lists, similarities and hits as the oracle."""
    n, d, C, nq, nprobe, k = 30000, 128, 16500, 70, 40, 10
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 71))
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, 72))
    cent[9000:9040] = cent[77]               # a run of identical centroids: ties by index
    cent[123, :] = 0                          # zero centroid
    lists = (np.arange(n) % C).astype(np.uint32)
    doc = np.arange(n, dtype=np.uint64)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    q = unit_rows(nq, d, 73)
    q[5] = 0                                  # zero query: unusable for the filter -> the literal path
    qs = oracle.quantize_matrix_f32(q)
    qs[6] = cent[77]
    probes, sims = ix.SelectProbes(qs, nprobe)
    for i in range(nq):
        wp, ws = oracle.select_probes(qs[i], cent, nprobe)
        assert probes[i].tolist() == wp.tolist(), f"query {i}"
        assert (f32_bits(sims[i]) == f32_bits(ws)).all(), f"query {i}"
    ids, hs, counts = ix.Search(qs, nprobe, k)
    for i in (0, 5, 6, 33, 69):
        want_ids, want_sims = oracle.search(qs[i], cent, rows, lists, doc, nprobe, k)
        assert ids[i, :counts[i]].tolist() == want_ids.tolist(), f"query {i}"
        assert (f32_bits(hs[i, :counts[i]]) == f32_bits(want_sims)).all()


# ---- one hit per document BEFORE the cut (search.go:259-271): a document whose embeddings crowd the top of the list ----
def _crowded_inputs(oracle, n, d, C, seed, crowd, ndocs_crowd=1):
    """`crowd` rows close to the query direction share `ndocs_crowd` document ids; every other row is its own document."""
    rng = np.random.default_rng(seed)
    x = unit_rows(n, d, seed)
    qdir = unit_rows(1, d, seed + 7)[0]
    hot = rng.choice(n, crowd, replace=False)
    x[hot] = qdir + 0.05 * rng.standard_normal((crowd, d)).astype(np.float32) / np.sqrt(d) * 4.0
    x[hot] /= np.linalg.norm(x[hot], axis=1, keepdims=True)
    rows = oracle.quantize_matrix_f32(x)
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, seed + 1))
    _, lists = oracle.argmax_MxN(cent, rows)
    doc = np.arange(n, dtype=np.uint64) + 10_000
    doc[hot] = 7 + (np.arange(crowd) % ndocs_crowd)
    q = oracle.quantize_vector_f32(qdir)
    return rows, cent, lists.astype(np.uint32), doc, q


@pytest.mark.parametrize("k,crowd,ndocs", [(10, 40, 1), (32, 40, 1), (20, 300, 3), (64, 200, 2), (100, 500, 5)])
def test_document_crowding_the_top(vs, oracle, k, crowd, ndocs):
    """One document owns more rows at the top than any kept list holds (32 / 64 / 128 entries): the reference still
    returns k distinct documents (it de-duplicates the whole running list before it truncates)."""
    n, d, C = 12000, 256, 12
    rows, cent, lists, doc, q = _crowded_inputs(oracle, n, d, C, 300 + k, crowd, ndocs)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = np.stack([q, oracle.quantize_vector_f32(unit_rows(1, d, 5)[0])])
    for nprobe in (C, 5):      # every list (flat form of the scan) and a probed subset
        ids, sims, counts = ix.Search(qs, nprobe, k)
        for i in range(2):
            want_ids, want_sims = oracle.search(qs[i], cent, rows, lists, doc, nprobe, k)
            assert counts[i] == len(want_ids) == k, (nprobe, i, counts[i], len(want_ids))
            assert ids[i, :k].tolist() == want_ids.tolist(), (nprobe, i)
            assert (f32_bits(sims[i, :k]) == f32_bits(want_sims)).all()
            assert len(set(ids[i, :k].tolist())) == k


def test_merge_removes_cross_shard_duplicates_before_the_cut(vs, oracle):
    """vs_topk_merge_dev: 8 shards x 20 hits where the same few documents lead every shard's list (a document's
    embeddings are striped over the ranks): the merged list still has k distinct documents."""
    import torch
    G, nq, k = 8, 3, 20
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    ids = np.zeros((G, nq, k), np.uint64)
    sims = np.zeros((G, nq, k), np.float32)
    for g in range(G):
        for q in range(nq):
            s = np.sort(rng.random(k).astype(np.float32))[::-1]
            i = rng.permutation(1000)[:k].astype(np.uint64) + 100 * (g + 1) * 1000
            i[:12] = np.arange(12) + q          # the same 12 documents lead on every shard
            s[:12] = s[:12] + 1.0
            ids[g, q], sims[g, q] = i, s
    counts = np.full((G, nq), k, np.int32)
    t = lambda a, dt: torch.from_numpy(a.view(dt) if a.dtype == np.uint64 else a).to(dev)
    g_ids, g_sims, g_counts = t(ids, np.int64), t(sims, None), t(counts, None)
    out_ids = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    out_sims = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    out_counts = torch.zeros(nq, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx = vs.compute.default_context()
    vs.shard.merge_hits_dev(g_ids, g_sims, g_counts, k, out_ids, out_sims, out_counts, ctx=ctx)
    ctx.sync()
    got_ids = out_ids.cpu().numpy().view(np.uint64)
    got_sims = out_sims.cpu().numpy()
    for q in range(nq):
        best = {}
        for g in range(G):
            for j in range(k):
                d_, s_ = int(ids[g, q, j]), float(sims[g, q, j])
                if d_ not in best or s_ > best[d_]:
                    best[d_] = s_
        want = sorted(best.items(), key=lambda kv: (-kv[1], kv[0]))[:k]
        assert int(out_counts[q]) == k
        assert got_ids[q].tolist() == [w[0] for w in want]
        assert got_sims[q].tolist() == [np.float32(w[1]) for w in want]


# ---- requests beyond the fused kernels' capacities (server/search.go:116-122: any Centroids value, unbounded Offset) ----
@pytest.mark.parametrize("nprobe,k", [(200, 10), (200, 200), (8, 300), (1000, 150)])
def test_wide_requests(vs, oracle, nprobe, k):
    """More than 128 probed lists (without probing all of them) and more than 128 hits per query."""
    n, d, C = 20000, 256, 1000
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 91, docs_per=2)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = oracle.quantize_matrix_f32(unit_rows(3, d, 92))
    _check(oracle, ix, qs, cent, rows, lists, doc, nprobe, k)


def test_more_queries_than_one_launch_takes(vs, oracle):
    n, d, C, nq = 6000, 128, 16, 4100
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 93)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 94))
    ids, sims, counts = ix.Search(qs, 4, 10)
    for i in (0, 4095, 4096, 4099):
        want_ids, want_sims = oracle.search(qs[i], cent, rows, lists, doc, 4, 10)
        assert ids[i, :counts[i]].tolist() == want_ids.tolist()
        assert (f32_bits(sims[i, :counts[i]]) == f32_bits(want_sims)).all()


@pytest.mark.parametrize("G,nq,nprobe,C", [(2, 64, 8, 96), (4, 16, 3, 40), (4, 8, 2, 40), (8, 1024, 32, 256)])
def test_split_probe_and_list_stage_match_oracle(vs, oracle, G, nq, nprobe, C):
    """The sharded probe stage emulated on one GPU: shard g selects the probe lists of ITS 1/G of the batch
    (vs_probe_dev), the lists and status words are put side by side (the all-gather of an N-GPU run), every shard runs
    the list stage on the WHOLE batch with them (vs_search_dev_probed) and the shard-local hits are merged -- equal to
    the oracle's search over the whole store, and the gathered probe lists equal to what one device selects for all."""
    import torch
    n, d, k = 24000, 768, 10
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 171, docs_per=2)
    qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 172))
    dev = torch.device("cuda", 0)
    ctx = vs.compute.default_context()
    shards = []
    for r in range(G):
        mine = vs.shard.stripe(n, r, G)
        shards.append(vs.ivf.Index.build_assigned(rows[mine], doc[mine], lists[mine], cent))
    per = nq // G
    q_all = vs.compute.NewMatrix(qs, ctx=ctx)
    d_probe = torch.zeros((nq, nprobe), dtype=torch.int32, device=dev)
    d_pstat = torch.zeros(nq, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    slices = []
    for r in range(G):
        q_r = vs.compute.NewMatrix(qs[r * per:(r + 1) * per], ctx=ctx)
        slices.append(q_r)
        shards[r].ProbeDev(q_r, nprobe, d_probe[r * per:].data_ptr(), d_pstat[r * per:].data_ptr(), ctx=ctx)
    ctx.sync()
    want_probe, _ = shards[0].SelectProbes(qs, nprobe, ctx=ctx)
    assert (d_probe.cpu().numpy().astype(np.uint32) == want_probe).all()
    g_ids = torch.zeros((G, nq, k), dtype=torch.int64, device=dev)
    g_sims = torch.zeros((G, nq, k), dtype=torch.float32, device=dev)
    g_counts = torch.zeros((G, nq), dtype=torch.int32, device=dev)
    g_stat = d_pstat.repeat(G, 1).contiguous()
    torch.cuda.synchronize()
    for r in range(G):
        shards[r].SearchDevProbed(q_all, nprobe, k, d_probe.data_ptr(), g_ids[r].data_ptr(), g_sims[r].data_ptr(), g_counts[r].data_ptr(),
                                  g_stat[r].data_ptr(), ctx=ctx)
        shards[r].Resolve(q_all, nprobe, k, g_ids[r].data_ptr(), g_sims[r].data_ptr(), g_counts[r].data_ptr(), g_stat[r].data_ptr(), ctx=ctx)
    out_ids = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    out_sims = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    out_counts = torch.zeros(nq, dtype=torch.int32, device=dev)
    ctx.sync()
    vs.shard.merge_hits_dev(g_ids, g_sims, g_counts, k, out_ids, out_sims, out_counts, ctx=ctx)
    ctx.sync()
    ids = out_ids.cpu().numpy().view(np.uint64)
    sims = out_sims.cpu().numpy()
    counts = out_counts.cpu().numpy()
    for i in range(0, nq, max(1, nq // 16)):
        wi, ws = oracle.search(qs[i], cent, rows, lists, doc, nprobe, k)
        assert counts[i] == len(wi)
        assert ids[i, :counts[i]].tolist() == wi.tolist()
        assert (f32_bits(sims[i, :counts[i]]) == f32_bits(ws)).all()
