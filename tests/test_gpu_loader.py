"""GPU parity: the streaming loader (database -> HBM with centroid_id grouping; database/model.go:9-18, search.go:241-243).
An index filled chunk by chunk in primary-key order is the index built in one piece from the same rows."""
import numpy as np
import pytest
import torch

from _util import f32_bits, unit_rows

pytestmark = pytest.mark.gpu


def _table(oracle, n, d, C, seed):
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, seed))
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, seed + 1))
    _, lists = oracle.argmax_MxN(cent, rows)
    doc = np.random.default_rng(seed + 2).integers(0, n // 2, n).astype(np.uint64)   # several rows per document
    return rows, cent, lists.astype(np.uint32), doc


def _same_store(a, b):
    assert a.rows == b.rows and a.lists == b.lists and a.cols == b.cols
    assert a.ListOffsets().tolist() == b.ListOffsets().tolist()
    ra, ia = a.ReadRows(0, a.rows)
    rb, ib = b.ReadRows(0, b.rows)
    assert (ia == ib).all() and (ra == rb).all()


def _search_parity(oracle, ix, qs, cent, rows, lists, doc, nprobe, k):
    ids, sims, counts = ix.Search(qs, nprobe, k)
    for i, q in enumerate(qs):
        want_ids, want_sims = oracle.search(q, cent, rows, lists, doc, nprobe, k)
        c = counts[i]
        assert c == len(want_ids)
        assert ids[i, :c].tolist() == want_ids.tolist(), f"query {i}"
        assert (f32_bits(sims[i, :c]) == f32_bits(want_sims)).all(), f"query {i}"


@pytest.mark.parametrize("chunks", [[12000], [1, 4999, 7000], [3000, 3000, 3000, 3000], [11999, 1]])
def test_filled_equals_built(vs, oracle, chunks):
    n, d, C = 12000, 768, 40
    rows, cent, lists, doc = _table(oracle, n, d, C, 17)
    lists[lists == 5] = 6                                     # an empty list
    cm = vs.compute.NewMatrix(cent)
    ix = vs.ivf.Index.create_empty(cm, np.bincount(lists, minlength=C))
    qs = oracle.quantize_matrix_f32(unit_rows(4, d, 8))
    at = 0
    for m in chunks:
        if 0 < at < n:                                           # half loaded: answers from the rows placed so far
            assert ix.rows == at
            _search_parity(oracle, ix, qs[:1], cent, rows[:at], lists[:at], doc[:at], nprobe=4, k=10)
        ix.Fill(rows[at:at + m], lists[at:at + m], doc[at:at + m])
        at += m
    _same_store(ix, vs.ivf.Index.build_assigned(rows, doc, lists, cent))
    _search_parity(oracle, ix, qs, cent, rows, lists, doc, nprobe=6, k=10)
    ix2, assign = ix.Upload(rows[:50], doc[:50])              # a loaded index takes uploads like any other
    assert ix2.rows == n + 50


def test_fill_from_device_chunks_with_implicit_ids(vs, oracle):
    """The device form, as bench.py uses it: chunk matrix + int32 assignment from ArgmaxDev, ids = primary key."""
    n, d, C, step = 9000, 256, 24, 2500
    rows, cent, lists, _ = _table(oracle, n, d, C, 23)
    cm = vs.compute.NewMatrix(cent)
    ix = vs.ivf.Index.create_empty(cm, np.bincount(lists, minlength=C))
    for at in range(0, n, step):
        chunk = vs.compute.NewMatrix(rows[at:at + step])
        assign = torch.empty(chunk.rows, device="cuda", dtype=torch.int32)
        cm.ArgmaxDev(chunk, assign.data_ptr())
        ix.FillDev(chunk, assign.data_ptr(), None, id_base=at)
        assert (assign.cpu().numpy() == lists[at:at + step]).all()
    doc = np.arange(n, dtype=np.uint64)
    _same_store(ix, vs.ivf.Index.build_assigned(rows, doc, lists, cent))
    qs = oracle.quantize_matrix_f32(unit_rows(3, d, 2))
    _search_parity(oracle, ix, qs, cent, rows, lists, doc, nprobe=5, k=20)


def test_loader_refuses_bad_chunks(vs, oracle):
    n, d, C = 600, 64, 6
    rows, cent, lists, doc = _table(oracle, n, d, C, 31)
    cm = vs.compute.NewMatrix(cent)
    counts = np.bincount(lists, minlength=C)
    c = vs.compute
    ix = vs.ivf.Index.create_empty(cm, counts)
    with pytest.raises(c.ComputeError, match="name a list"):
        ix.Fill(rows[:10], np.full(10, C, np.uint32), doc[:10])
    with pytest.raises(c.ComputeFatal, match="column size does not match"):
        ix.Fill(np.zeros((4, 8 + d + 16), np.uint8), np.zeros(4, np.uint32), doc[:4])
    with pytest.raises(c.ComputeError, match="do not fit"):
        ix.Fill(np.concatenate([rows, rows[:1]]), np.concatenate([lists, lists[:1]]), None)
    short = counts.copy()
    short[lists[0]] -= 1
    short[(lists[0] + 1) % C] += 1
    ix = vs.ivf.Index.create_empty(cm, short)
    with pytest.raises(c.ComputeError, match="more rows than"):
        ix.Fill(rows, lists, doc)
    built = vs.ivf.Index.build_assigned(rows, doc, lists, cent)
    with pytest.raises(c.ComputeError, match="not created by"):
        built.Fill(rows[:3], lists[:3], doc[:3])
