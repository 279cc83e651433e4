"""GPU parity: Upload's assignment and insert (server/upload.go:239-279) -- an index that took uploads is the index
built from the embeddings table after those uploads, and answers searches like the oracle over that table."""
import numpy as np
import pytest

from _util import f32_bits, unit_rows

pytestmark = pytest.mark.gpu


def _table(oracle, n, d, C, seed):
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, seed))
    cent = oracle.quantize_matrix_f32(unit_rows(C, d, seed + 1))
    doc = np.random.default_rng(seed + 2).permutation(n).astype(np.uint64) + 7   # ids not in row order
    return rows, cent, doc


def _same_store(a, b):
    assert a.rows == b.rows and a.lists == b.lists
    assert a.ListOffsets().tolist() == b.ListOffsets().tolist()
    ra, ia = a.ReadRows(0, a.rows)
    rb, ib = b.ReadRows(0, b.rows)
    assert (ia == ib).all()
    assert (ra == rb).all()


def _search_parity(oracle, ix, qs, cent, rows, lists, doc, nprobe, k):
    ids, sims, counts = ix.Search(qs, nprobe, k)
    for i, q in enumerate(qs):
        want_ids, want_sims = oracle.search(q, cent, rows, lists, doc, nprobe, k)
        c = counts[i]
        assert c == len(want_ids)
        assert ids[i, :c].tolist() == want_ids.tolist(), f"query {i}"
        assert (f32_bits(sims[i, :c]) == f32_bits(want_sims)).all(), f"query {i}"


@pytest.mark.parametrize("n0,n1,C", [(20000, 3000, 96), (6000, 1500, 300), (500, 1, 7), (64, 4000, 20)])
def test_upload_equals_rebuild(vs, oracle, n0, n1, C):
    """C=300 with 1500 new rows takes the tensor-core assignment, the others the scan form."""
    d = 768
    rows, cent, doc = _table(oracle, n0 + n1, d, C, 11)
    _, lists0 = oracle.argmax_MxN(cent, rows[:n0])
    ix0 = vs.ivf.Index.build_assigned(rows[:n0], doc[:n0], lists0.astype(np.uint32), cent)
    ix1, assign = ix0.Upload(rows[n0:], doc[n0:])
    want_assign, lists, all_rows, all_doc = oracle.upload(cent, rows[n0:], lists0, rows[:n0], doc[:n0], doc[n0:])
    assert assign.tolist() == want_assign.tolist()
    _same_store(ix1, vs.ivf.Index.build_assigned(all_rows, all_doc, lists, cent))
    qs = oracle.quantize_matrix_f32(unit_rows(5, d, 77))
    _search_parity(oracle, ix1, qs, cent, all_rows, lists, all_doc, nprobe=8, k=10)
    # the index the upload started from is untouched and still answers
    assert ix0.rows == n0
    _search_parity(oracle, ix0, qs[:2], cent, rows[:n0], lists0.astype(np.uint32), doc[:n0], nprobe=8, k=10)


def test_uploads_in_sequence_and_empty_lists(vs, oracle):
    """Lists that were empty receive rows; a row identical to two centroids goes to the lower index (cosine.go:114)."""
    d, C = 256, 12
    rows, cent, doc = _table(oracle, 3000, d, C, 5)
    cent[9] = cent[4]                                        # duplicate centroid: never wins
    _, lists_all = oracle.argmax_MxN(cent, rows)
    keep = np.flatnonzero((lists_all != 2) & (lists_all != 7))[:1500]   # lists 2 and 7 start empty
    lists0 = lists_all[keep].astype(np.uint32)
    ix = vs.ivf.Index.build_assigned(rows[keep], doc[keep], lists0, cent)
    cur_rows, cur_doc, cur_lists = rows[keep], doc[keep], lists0
    rest = np.setdiff1d(np.arange(3000), keep)
    for part in np.array_split(rest, 3):
        new_rows = rows[part].copy()
        new_rows[0] = cent[4]                                # equally near centroid 4 and its copy 9
        ix, assign = ix.Upload(new_rows, doc[part])
        want, cur_lists, cur_rows, cur_doc = oracle.upload(cent, new_rows, cur_lists, cur_rows, cur_doc, doc[part])
        assert assign.tolist() == want.tolist() and assign[0] == 4
    off = ix.ListOffsets()
    assert off[3] > off[2] and off[8] > off[7] and off[10] == off[9]
    _same_store(ix, vs.ivf.Index.build_assigned(cur_rows, cur_doc, cur_lists, cent))
    qs = oracle.quantize_matrix_f32(unit_rows(4, d, 3))
    _search_parity(oracle, ix, qs, cent, cur_rows, cur_lists, cur_doc, nprobe=5, k=20)


def test_search_and_upload_kat(vs):
    """The hand-derived five-row store of tests/golden/kat.json (no oracle involved): search, dedup, upload, loader."""
    import json
    import os
    import struct
    c = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat.json")))["search"]
    cent = np.array(c["centroids"], np.uint8)
    rows = np.array(c["rows"], np.uint8)
    doc = np.array(c["doc_ids"], np.uint64)
    lists = np.array(c["lists"], np.uint32)
    q = np.array(c["query"], np.uint8)[None, :]

    def check(ix, cases):
        for s in cases:
            ids, sims, counts = ix.Search(q, s["nprobe"], s["k"])
            assert ids[0, :counts[0]].tolist() == s["ids"], s["why"]
            assert [struct.pack(">f", x).hex() for x in sims[0, :counts[0]]] == s["sims_f32"], s["why"]

    check(vs.ivf.Index.build_assigned(rows, doc, lists, cent), c["searches"])
    check(vs.ivf.Index.build_assigned(rows, np.array(c["dedup"]["doc_ids"], np.uint64), lists, cent), [c["dedup"]])
    n0 = c["upload"]["first"]
    ix1, assign = vs.ivf.Index.build_assigned(rows[:n0], doc[:n0], lists[:n0], cent).Upload(rows[n0:], doc[n0:])
    assert assign.tolist() == c["upload"]["assign"], c["upload"]["why"]
    check(ix1, c["searches"])
    ld = vs.ivf.Index.create_empty(vs.compute.NewMatrix(cent), np.bincount(lists, minlength=2))
    ld.Fill(rows[:2], lists[:2], doc[:2])
    ld.Fill(rows[2:], lists[2:], doc[2:])
    check(ld, c["searches"])


def test_upload_implicit_ids(vs, oracle):
    """An index built without document ids numbers its rows; uploaded rows continue the numbering."""
    d, C, n0, n1 = 128, 5, 700, 90
    rows, cent, _ = _table(oracle, n0 + n1, d, C, 9)
    _, lists0 = oracle.argmax_MxN(cent, rows[:n0])
    ix0 = vs.ivf.Index.build_assigned(rows[:n0], None, lists0.astype(np.uint32), cent)
    ix1, assign = ix0.Upload(rows[n0:])
    _, lists, _, _ = oracle.upload(cent, rows[n0:], lists0)
    qs = oracle.quantize_matrix_f32(unit_rows(3, d, 1))
    _search_parity(oracle, ix1, qs, cent, rows, lists, np.arange(n0 + n1, dtype=np.uint64), nprobe=C, k=10)


def test_upload_errors_like_reference(vs, oracle):
    d, C = 64, 4
    rows, cent, doc = _table(oracle, 100, d, C, 2)
    _, lists0 = oracle.argmax_MxN(cent, rows)
    ix = vs.ivf.Index.build_assigned(rows, doc, lists0.astype(np.uint32), cent)
    c = vs.compute
    with pytest.raises(c.ComputeFatal, match="column size does not match"):      # cosine.go:77-79
        ix.Upload(np.zeros((3, 8 + d + 1), np.uint8), np.arange(3, dtype=np.uint64))
    with pytest.raises(c.ComputePanic):                                          # compute.go:25-27
        ix.Upload(np.zeros((0, 8 + d), np.uint8), np.zeros(0, np.uint64))
    with pytest.raises(c.ComputeError, match="doc_ids is required"):
        ix.Upload(rows[:2])
    assert ix.rows == 100


def _lists_equal_table(ix, all_rows, all_doc, lists, C):
    """Every list of the (possibly holey) store holds exactly the table's rows of that list, in primary-key order."""
    off, ln = ix.ListOffsets(), ix.ListLengths()
    assert int(ln.sum()) == ix.rows == all_rows.shape[0]
    for l in range(C):
        members = np.flatnonzero(lists == l)
        assert ln[l] == members.shape[0] and off[l] + ln[l] <= off[l + 1]
        if ln[l]:
            r, i = ix.ReadRows(int(off[l]), int(ln[l]))
            assert (i == all_doc[members]).all() and (r == all_rows[members]).all()


@pytest.mark.parametrize("n0,C,parts", [(20000, 96, 4), (3000, 300, 3), (64, 20, 5)])
def test_upload_in_place_equals_rebuild(vs, oracle, n0, C, parts):
    """vs_index_with_room + vs_index_append: after every append the lists are the embeddings table's lists and searches
    (one query: fused launch; a batch: list-major; few queries: streaming scan; every list probed; wide requests) answer
    like the oracle over that table.  A list that fills up raises IndexFull and leaves the index as it was."""
    d, n1 = 768, 2400
    rows, cent, doc = _table(oracle, n0 + n1, d, C, 21)
    _, lists0 = oracle.argmax_MxN(cent, rows[:n0])
    packed = vs.ivf.Index.build_assigned(rows[:n0], doc[:n0], lists0.astype(np.uint32), cent)
    with pytest.raises(vs.compute.IndexFull):
        packed.Append(rows[n0:n0 + 1], doc[n0:n0 + 1])
    ix = packed.WithRoom(percent=10, min_rows=n1)                # enough for everything: no list can overflow
    assert ix.rows == n0 and ix.capacity >= n0 + C * n1
    cur_rows, cur_doc, cur_lists = rows[:n0], doc[:n0], lists0.astype(np.uint32)
    qs = oracle.quantize_matrix_f32(unit_rows(24, d, 78))
    for part in np.array_split(np.arange(n0, n0 + n1), parts):
        assign = ix.Append(rows[part], doc[part])
        want, cur_lists, cur_rows, cur_doc = oracle.upload(cent, rows[part], cur_lists, cur_rows, cur_doc, doc[part])
        assert assign.tolist() == want.tolist()
        assert ix.rows == cur_rows.shape[0]
        _search_parity(oracle, ix, qs[:1], cent, cur_rows, cur_lists, cur_doc, nprobe=8, k=10)      # fused
    _lists_equal_table(ix, cur_rows, cur_doc, cur_lists, C)
    _search_parity(oracle, ix, qs[:3], cent, cur_rows, cur_lists, cur_doc, nprobe=8, k=10)           # streaming scan
    _search_parity(oracle, ix, qs, cent, cur_rows, cur_lists, cur_doc, nprobe=min(C, 16), k=10)      # batch
    _search_parity(oracle, ix, qs[:2], cent, cur_rows, cur_lists, cur_doc, nprobe=C, k=10)           # every list (holes: by lists)
    _search_parity(oracle, ix, qs[:1], cent, cur_rows, cur_lists, cur_doc, nprobe=C + 5, k=200)      # wide request
    # the packed index it was copied from is untouched
    assert packed.rows == n0


def test_append_refuses_what_does_not_fit(vs, oracle):
    d, C, n0 = 256, 8, 800
    rows, cent, doc = _table(oracle, n0 + 600, d, C, 31)
    _, lists0 = oracle.argmax_MxN(cent, rows[:n0])
    ix = vs.ivf.Index.build_assigned(rows[:n0], doc[:n0], lists0.astype(np.uint32), cent).WithRoom(percent=0, min_rows=10)
    before = ix.ListLengths().copy()
    with pytest.raises(vs.compute.IndexFull):
        ix.Append(rows[n0:], doc[n0:])                           # 600 rows over 8 lists with 10 free places each
    assert (ix.ListLengths() == before).all() and ix.rows == n0
    qs = oracle.quantize_matrix_f32(unit_rows(2, d, 5))
    _search_parity(oracle, ix, qs, cent, rows[:n0], lists0.astype(np.uint32), doc[:n0], nprobe=3, k=10)
    # UploadInPlace copies once into a roomier store, then appends; later small uploads need no copy
    assign, copied = ix.UploadInPlace(rows[n0:n0 + 500], doc[n0:n0 + 500])
    assert copied
    want, lists, all_rows, all_doc = oracle.upload(cent, rows[n0:n0 + 500], lists0, rows[:n0], doc[:n0], doc[n0:n0 + 500])
    assert assign.tolist() == want.tolist()
    assign2, copied2 = ix.UploadInPlace(rows[n0 + 500:], doc[n0 + 500:])
    assert not copied2
    want2, lists, all_rows, all_doc = oracle.upload(cent, rows[n0 + 500:], lists, all_rows, all_doc, doc[n0 + 500:])
    assert assign2.tolist() == want2.tolist()
    _lists_equal_table(ix, all_rows, all_doc, lists, C)
    _search_parity(oracle, ix, qs, cent, all_rows, lists, all_doc, nprobe=3, k=10)


def test_append_continues_implicit_numbering(vs, oracle):
    """An index built without document ids numbers its rows in primary-key order; appended rows continue the count."""
    d, C, n0, n1 = 128, 6, 500, 120
    rows, cent, _ = _table(oracle, n0 + n1, d, C, 41)
    _, lists0 = oracle.argmax_MxN(cent, rows[:n0])
    ix = vs.ivf.Index.build_assigned(rows[:n0], None, lists0.astype(np.uint32), cent).WithRoom(percent=50, min_rows=n1)
    ix.Append(rows[n0:])
    ids = np.arange(n0 + n1, dtype=np.uint64)
    _, lists, all_rows, all_doc = oracle.upload(cent, rows[n0:], lists0, rows[:n0], ids[:n0], ids[n0:])
    qs = oracle.quantize_matrix_f32(unit_rows(3, d, 6))
    _search_parity(oracle, ix, qs, cent, all_rows, lists, all_doc, nprobe=C, k=15)
