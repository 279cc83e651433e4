"""GPU parity of the list-major batch scan (csrc/listmajor.cu: every probed list read once for all the queries that probe
it) against the oracle's restatement of server/search.go:202-273 and against the query-major scan."""
import numpy as np
import pytest

from _util import f32_bits, noop_rows, unit_rows
from test_gpu_search import _check, _crowded_inputs, _index_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["dp4a", "dense"], autouse=True)
def scan_form(request, vs):
    """Every test runs twice: with the dp4a form of the scan (lm_scan_kernel) and with the tensor-core form
    (lm_dense_kernel) wherever the row width allows it -- the library picks by queries per list on its own."""
    vs.compute.debug_set_lm_dense_min(0 if request.param == "dp4a" else 1)
    yield request.param
    vs.compute.debug_set_lm_dense_min(4)


def _same_as_query_major(vs, ix, qs, nprobe, k, ctx=None):
    a = ix.Search(qs, nprobe, k, ctx=ctx)
    vs.compute.debug_set_list_major(False)
    try:
        b = ix.Search(qs, nprobe, k, ctx=ctx)
    finally:
        vs.compute.debug_set_list_major(True)
    assert (a[2] == b[2]).all()
    assert (a[0] == b[0]).all()
    assert (f32_bits(a[1]) == f32_bits(b[1])).all()
    return a


def test_list_major_path_is_taken(vs, oracle):
    n, d, C = 20000, 768, 64
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 5)
    ctx = vs.compute.Context()
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent, ctx=ctx)
    qs = oracle.quantize_matrix_f32(unit_rows(32, d, 9))
    ix.Search(qs, 8, 10, ctx=ctx)
    l0 = ctx.launch_count()
    ix.Search(qs, 8, 10, ctx=ctx)
    lm = ctx.launch_count() - l0
    vs.compute.debug_set_list_major(False)
    try:
        l0 = ctx.launch_count()
        ix.Search(qs, 8, 10, ctx=ctx)
        qm = ctx.launch_count() - l0
    finally:
        vs.compute.debug_set_list_major(True)
    assert lm != qm       # (the list-major pipeline has its inversion and final kernels)
    ctx.close()


@pytest.mark.parametrize("d,C,nq,nprobe,k", [(768, 96, 40, 8, 10), (768, 96, 64, 32, 20), (512, 300, 200, 6, 10), (384, 40, 17, 39, 32),
                                             (1024, 24, 48, 3, 10), (1536, 16, 32, 5, 7)])
def test_list_major_parity(vs, oracle, d, C, nq, nprobe, k):
    n = 30000 if d == 768 else 9000
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 50 + C, docs_per=2)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 99))
    _same_as_query_major(vs, ix, qs, nprobe, k)
    sel = np.arange(0, nq, max(1, nq // 12))
    ids, sims, counts = ix.Search(qs, nprobe, k)
    for i in sel:
        want_ids, want_sims = oracle.search(qs[i], cent, rows, lists, doc, nprobe, k)
        c = counts[i]
        assert c == len(want_ids)
        assert ids[i, :c].tolist() == want_ids.tolist(), f"query {i}"
        assert (f32_bits(sims[i, :c]) == f32_bits(want_sims)).all(), f"query {i}"


def test_list_major_many_queries_per_list(vs, oracle):
    """More queries on a list than a block has warps (several passes over the item), long lists cut into sub-items."""
    n, d, C = 24000, 768, 6
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 77)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = oracle.quantize_matrix_f32(unit_rows(50, d, 3))
    ids, sims, counts = _same_as_query_major(vs, ix, qs, 4, 10)
    for i in (0, 13, 49):
        want_ids, want_sims = oracle.search(qs[i], cent, rows, lists, doc, 4, 10)
        assert ids[i, :counts[i]].tolist() == want_ids.tolist()
        assert (f32_bits(sims[i, :counts[i]]) == f32_bits(want_sims)).all()


def test_list_major_ragged_empty_ties_and_zero_vectors(vs, oracle):
    d, C = 768, 10
    rows = oracle.quantize_matrix_f32(unit_rows(3000, d, 2))
    rows[100:140] = rows[100]
    rows[200:204, 8:] = 0
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, 3))
    sizes = [0, 1, 31, 32, 33, 0, 1500, 7, 1396, 0]
    lists = np.repeat(np.arange(C), sizes).astype(np.uint32)
    doc = np.random.default_rng(0).permutation(3000).astype(np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    ix = vs.ivf.Index.build(rows, doc, offs, cent)
    q = unit_rows(20, d, 5)
    q[3] = 0
    qs = oracle.quantize_matrix_f32(q)
    qs[4] = rows[100]
    for nprobe, k in ((3, 10), (9, 25), (1, 5)):
        _check(oracle, ix, qs, cent, rows, lists, doc, nprobe, k)


@pytest.mark.parametrize("k,crowd,ndocs", [(10, 40, 1), (32, 300, 3)])
def test_list_major_document_crowding(vs, oracle, k, crowd, ndocs):
    n, d, C = 12000, 384, 12
    rows, cent, lists, doc, q = _crowded_inputs(oracle, n, d, C, 300 + k, crowd, ndocs)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    qs = np.concatenate([q[None, :], oracle.quantize_matrix_f32(unit_rows(19, d, 5))])
    _check(oracle, ix, qs, cent, rows, lists, doc, 5, k)


def test_list_major_uncertified_scores_go_to_the_literal_path(vs, oracle):
    n, d, C = 6000, 768, 24
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 15, docs_per=2)
    ctx = vs.compute.Context()
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent, ctx=ctx)
    qs = oracle.quantize_matrix_f32(unit_rows(16, d, 77))
    vs.compute.debug_set_certify_scale(1.0e7)
    try:
        ids, sims, counts = ix.Search(qs, 5, 12, ctx=ctx)
    finally:
        vs.compute.debug_set_certify_scale(1.0)
    assert ctx.slowpath_count() > 0
    for i in (0, 7, 15):
        want_ids, want_sims = oracle.search(qs[i], cent, rows, lists, doc, 5, 12)
        assert ids[i, :counts[i]].tolist() == want_ids.tolist()
        assert (f32_bits(sims[i, :counts[i]]) == f32_bits(want_sims)).all()
    ctx.close()


def test_list_major_large_batch_equals_query_major(vs):
    """256 queries x 16 probes over 512 lists of noop rows (no oracle at this size): same hits as the query-major scan,
    three times in a row on one context (work queue, running bounds and candidate lists are re-armed every call)."""
    n, d, C = 200000, 768, 512
    rows = noop_rows(n, d, 3)
    lists = (np.arange(n) * 2654435761 % C).astype(np.uint32)
    ctx = vs.compute.Context()
    ix = vs.ivf.Index.build_assigned(rows, None, lists, rows[:C], ctx=ctx)
    for rep in range(3):
        qs = noop_rows(256, d, 40 + rep)
        ids, sims, counts = _same_as_query_major(vs, ix, qs, 16, 10, ctx=ctx)
        assert (counts == 10).all()
    ctx.close()


def test_dense_form_passes_and_partial_groups(vs, oracle, scan_form):
    """Queries per list around the pass and group sizes of the tensor-core form (8 per instruction group, 16 per pass):
    1, 7, 8, 9, 16, 17 and 33 queries on the lists of one small index; list lengths around the 16-row tiles and 32-row stages."""
    d, C = 768, 7
    sizes = [15, 16, 17, 33, 1100, 64, 2049]
    n = sum(sizes)
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 21))
    lists = np.repeat(np.arange(C), sizes).astype(np.uint32)
    doc = np.random.default_rng(4).permutation(n).astype(np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, 22))
    ix = vs.ivf.Index.build(rows, doc, offs, cent)
    for nq in (16, 17, 33, 91):
        qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 23 + nq))
        for nprobe in (1, 3, 7):
            _check(oracle, ix, qs, cent, rows, lists, doc, nprobe, 10)


def test_default_form_follows_queries_per_list(vs, oracle):
    """With the default setting a sparse batch runs the dp4a kernel and a dense one the tensor-core kernel; both give the
    query-major hits (launch counts are equal, so the forms are told apart only by their results being checked here)."""
    vs.compute.debug_set_lm_dense_min(4)
    n, d, C = 40000, 768, 128
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 31)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    for nq, nprobe in ((32, 8), (128, 8), (512, 16)):   # 2, 8 and 64 queries per list
        qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 100 + nq))
        _same_as_query_major(vs, ix, qs, nprobe, 10)


def _hard_rows(n, d, seed):
    """Rows that stress the score identity and the float32 screen of the dense form: large common offsets (strong
    cancellation inside the identity), tiny and huge ranges, all-positive data, constant rows, all-zero rows, a few
    near-duplicates (scores a float32 step apart)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float64)
    fam = rng.integers(0, 8, n)
    x[fam == 0] += 40.0                                   # offset >> spread
    x[fam == 1] = np.abs(x[fam == 1]) * 1e-6              # tiny, all positive
    x[fam == 2] *= 1e6                                    # huge
    x[fam == 3] = x[fam == 3] * 1e-3 + 7.0                # nearly constant
    x[fam == 4] = np.round(x[fam == 4] * 3)               # few distinct values
    x[fam == 5, : d // 2] = 0                             # half zero
    const = np.flatnonzero(fam == 6)[:5]
    x[const] = 3.25                                       # constant rows (range 0)
    zero = np.flatnonzero(fam == 6)[5:9]
    x[zero] = 0.0                                         # all-zero rows
    base = np.flatnonzero(fam == 7)
    for i in range(0, len(base) - 1, 2):                  # near-duplicates
        x[base[i + 1]] = x[base[i]] * (1.0 + 1e-7 * rng.standard_normal(d))
    return x.astype(np.float32)


def test_screen_never_drops_a_hit_on_hard_data(vs, oracle):
    """The dense form's float32 screen against data families chosen to break it; every query against the oracle."""
    d, C, n = 768, 6, 4200
    rows = oracle.quantize_matrix_f32(_hard_rows(n, d, 5))
    cent = oracle.quantize_matrix_f32(_hard_rows(C, d, 6))
    lists = (np.arange(n) % C).astype(np.uint32)
    order = np.argsort(lists, kind="stable")
    rows, lists = rows[order], lists[order]
    doc = np.random.default_rng(1).permutation(n).astype(np.uint64)
    offs = np.concatenate([[0], np.cumsum(np.bincount(lists, minlength=C))]).astype(np.uint64)
    ix = vs.ivf.Index.build(rows, doc, offs, cent)
    qs = oracle.quantize_matrix_f32(_hard_rows(40, d, 7))
    for nprobe, k in ((2, 10), (6, 32)):
        _check(oracle, ix, qs, cent, rows, lists, doc, nprobe, k)


def test_weak_seed_is_repaired_by_the_running_bound(vs, oracle, scan_form):
    """Most queries' nearest list is shorter than k, so the seed gives no bound, while one list holds more rows than a
    query's candidate list can: the dp4a form limits itself through its per-warp lists, the tensor-core form by recomputing
    the query's bound from its own candidate list -- neither may overflow into the literal path, and the hits are the
    oracle's."""
    d, C, k, nprobe, nq = 768, 6, 32, 5, 32
    sizes = [20, 9000, 20, 20, 20, 20]
    n = sum(sizes)
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 61))
    lists = np.repeat(np.arange(C), sizes).astype(np.uint32)
    doc = np.random.default_rng(6).permutation(n).astype(np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, 62))
    import torch
    ctx = vs.compute.Context()
    ix = vs.ivf.Index.build(rows, doc, offs, cent, ctx=ctx)
    qs = oracle.quantize_matrix_f32(unit_rows(nq, d, 63))
    dev = torch.device("cuda", 0)
    d_ids = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    d_sims = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    d_counts = torch.zeros(nq, dtype=torch.int32, device=dev)
    d_status = torch.zeros(nq, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    qm = vs.compute.NewMatrix(qs, ctx=ctx)
    ix.SearchDev(qm, nprobe, k, d_ids.data_ptr(), d_sims.data_ptr(), d_counts.data_ptr(), d_status.data_ptr(), ctx=ctx)
    ctx.sync()
    status = d_status.cpu().numpy()
    assert (status & 2).sum() == 0, f"candidate lists overflowed: {status.tolist()}"
    ids, sims, counts = d_ids.cpu().numpy().view(np.uint64), d_sims.cpu().numpy(), d_counts.cpu().numpy()
    for i in range(0, nq, 3):
        want_ids, want_sims = oracle.search(qs[i], cent, rows, lists, doc, nprobe, k)
        assert ids[i, :counts[i]].tolist() == want_ids.tolist()
        assert (f32_bits(sims[i, :counts[i]]) == f32_bits(want_sims)).all()
    ctx.close()
