"""GPU parity: query batches as a tcgen05 int8 GEMM with fused filter (BASELINE config 3) -- same contract and the
same bits as the per-query scan (server/search.go:241-273 over the whole store)."""
import numpy as np
import pytest

from _util import f32_bits, noop_rows, unit_rows

pytestmark = pytest.mark.gpu


def _check_vs_oracle(oracle, rows, qs, ids, sims, counts, k, doc=None):
    for i, q in enumerate(qs):
        want_ids, want_sims = oracle.search_flat(q, rows, doc, k)
        c = counts[i]
        assert c == len(want_ids), (i, c, len(want_ids))
        assert ids[i, :c].tolist() == want_ids.tolist(), f"query {i}"
        assert (f32_bits(sims[i, :c]) == f32_bits(want_sims)).all(), f"query {i}"


@pytest.mark.parametrize("n,d,nq,k", [(12345, 768, 130, 10), (5000, 512, 7, 20), (3000, 100, 64, 5)])
def test_batch_gemm_parity_oracle(vs, oracle, n, d, nq, k):
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 7))
    qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 8))
    m = vs.compute.NewMatrix(rows)
    ids, sims, counts = vs.ivf.SearchFlatBatch(m, qs, k)
    _check_vs_oracle(oracle, rows, qs, ids, sims, counts, k)


def test_batch_gemm_matches_scan_large(vs):
    """Larger than the oracle handles quickly: the GEMM path must return the scan path's bits (the scan path is
    pinned to the oracle by test_gpu_search.py)."""
    n, d, nq, k = 300000, 768, 300, 10
    rows = noop_rows(n, d, 31)
    m = vs.compute.NewMatrix(rows)
    qs = noop_rows(nq, d, 32)
    ids, sims, counts = vs.ivf.SearchFlatBatch(m, qs, k)
    ids2, sims2, counts2 = vs.ivf.SearchFlat(m, qs, k)
    assert (counts == counts2).all() and (ids == ids2).all() and (f32_bits(sims) == f32_bits(sims2)).all()


def test_batch_gemm_unit_rows_matches_scan(vs, oracle):
    n, d, nq, k = 200000, 768, 256, 10
    x = unit_rows(n, d, 41)
    m = vs.compute.NewMatrix(vs.compute.QuantizeMatrixFloat32(x))
    qs = vs.compute.QuantizeMatrixFloat32(unit_rows(nq, d, 42))
    ids, sims, counts = vs.ivf.SearchFlatBatch(m, qs, k)
    ids2, sims2, counts2 = vs.ivf.SearchFlat(m, qs, k)
    assert (counts == counts2).all() and (ids == ids2).all() and (f32_bits(sims) == f32_bits(sims2)).all()


def test_batch_gemm_dedup_by_document(vs, oracle):
    """One hit per document (search.go:260-268) with several embeddings per document."""
    import torch
    n, d, nq, k = 9000, 768, 40, 12
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 51))
    qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 52))
    doc = np.random.default_rng(3).integers(0, n // 3, n).astype(np.uint64)
    m = vs.compute.NewMatrix(rows)
    d_doc = torch.from_numpy(doc.view(np.int64)).cuda()
    ids, sims, counts = vs.ivf.SearchFlatBatch(m, qs, k, doc_ids_dev=d_doc.data_ptr())
    _check_vs_oracle(oracle, rows, qs, ids, sims, counts, k, doc)


def test_batch_gemm_degenerate_rows_and_queries(vs, oracle):
    """Zero rows, constant rows, all-positive rows, identical rows (ties by id), a zero query: the filter must hand
    them to the certified / literal paths or the scan."""
    n, d, nq, k = 4000, 768, 20, 10
    x = unit_rows(n, d, 61)
    x[5] = 0
    x[6] = np.abs(x[6])
    x[7] = 0.25
    x[100:130] = x[100]
    rows = oracle.quantize_matrix_f32(x)
    q = unit_rows(nq, d, 62)
    q[3] = 0
    q[4] = np.abs(q[4])
    q[5] = x[100]
    qs = oracle.quantize_matrix_f32(q)
    m = vs.compute.NewMatrix(rows)
    ids, sims, counts = vs.ivf.SearchFlatBatch(m, qs, k)
    _check_vs_oracle(oracle, rows, qs, ids, sims, counts, k)


def test_batch_gemm_fewer_rows_than_k(vs, oracle):
    d = 768
    rows = oracle.quantize_matrix_f32(unit_rows(7, d, 2))
    qs = oracle.quantize_matrix_f32(unit_rows(3, d, 9))
    m = vs.compute.NewMatrix(rows)
    ids, sims, counts = vs.ivf.SearchFlatBatch(m, qs, 10)
    _check_vs_oracle(oracle, rows, qs, ids, sims, counts, 10)


def test_batch_gemm_literal_path(vs, oracle):
    n, d, nq, k = 6000, 768, 33, 10
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 71))
    qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 72))
    ctx = vs.compute.Context()
    m = vs.compute.NewMatrix(rows, ctx=ctx)
    vs.compute.debug_set_certify_scale(1.0e7)
    try:
        ids, sims, counts = vs.ivf.SearchFlatBatch(m, qs, k, ctx=ctx)
    finally:
        vs.compute.debug_set_certify_scale(1.0)
    assert ctx.slowpath_count() > 0
    _check_vs_oracle(oracle, rows, qs, ids, sims, counts, k)
    ctx.close()


def test_index_all_lists_batch_dispatch_matches_scan(vs):
    """vs_search with nprobe >= lists and a large batch is handed to the GEMM path (grouped store, document ids with
    duplicates); it must return what the per-query scans return."""
    import torch
    n, d, C, nq, k = 70000, 768, 32, 70, 10
    rows = noop_rows(n, d, 81)
    lists = (np.arange(n) % C).astype(np.uint32)
    doc = np.random.default_rng(5).integers(0, n // 2, n).astype(np.uint64)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, rows[:C])
    qs = noop_rows(nq, d, 82)
    ids, sims, counts = ix.Search(qs, C, k)                      # >= 64 queries, all lists: GEMM path
    parts = [ix.Search(qs[i:i + 16], C, k) for i in range(0, nq, 16)]   # small calls: streaming scans
    ids2 = np.concatenate([p[0] for p in parts]); sims2 = np.concatenate([p[1] for p in parts])
    counts2 = np.concatenate([p[2] for p in parts])
    assert (counts == counts2).all() and (ids == ids2).all() and (f32_bits(sims) == f32_bits(sims2)).all()
    # device-resident form over the index
    qm = vs.compute.NewMatrix(qs)
    dev = torch.device("cuda", 0)
    d_ids = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    d_sims = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    d_counts = torch.zeros(nq, dtype=torch.int32, device=dev)
    stats = ix.SearchBatchDev(qm, k, d_ids.data_ptr(), d_sims.data_ptr(), d_counts.data_ptr())
    vs.compute.default_context().sync()
    assert stats[2] == (n + 127) // 128
    assert (d_ids.cpu().numpy().view(np.uint64) == ids2).all()
    assert (f32_bits(d_sims.cpu().numpy()) == f32_bits(sims2)).all()
