"""GPU parity of the single-process multi-device index (vs_sharded_*, csrc/sharded.cu): rows striped over G shards by
primary key, shard-local top-k merged on device 0.  Runs with every stripe on device 0 on a one-GPU box and on real peer
devices when the box has them; results must equal the oracle (server/search.go:202-273) over ALL the rows."""
import numpy as np
import pytest

from _util import f32_bits, unit_rows
from test_gpu_search import _crowded_inputs, _index_inputs

pytestmark = pytest.mark.gpu


def _device_sets():
    import torch
    n = torch.cuda.device_count()
    sets = [[0], [0, 0, 0]]
    if n >= 2:
        sets.append([0, 1])
    if n >= 4:
        sets.append([0, 1, 2, 3])
    return sets


def _check(oracle, sh, qs, cent, rows, lists, doc, nprobe, k, sel=None):
    ids, sims, counts = sh.Search(qs, nprobe, k)
    for i in (range(len(qs)) if sel is None else sel):
        want_ids, want_sims = oracle.search(qs[i], cent, rows, lists, doc, nprobe, k)
        c = counts[i]
        assert c == len(want_ids), (i, c, len(want_ids))
        assert ids[i, :c].tolist() == want_ids.tolist(), f"query {i}"
        assert (f32_bits(sims[i, :c]) == f32_bits(want_sims)).all(), f"query {i}"


@pytest.mark.parametrize("docs_per", [1, 3])
def test_sharded_search_parity(vs, oracle, docs_per):
    n, d, C = 30000, 768, 96
    rows, cent, lists, doc = _index_inputs(oracle, n, d, C, 5, docs_per=docs_per)
    qs = oracle.quantize_matrix_f32(unit_rows(40, d, 99))
    for devices in _device_sets():
        sh = vs.ivf.ShardedIndex(devices).build_assigned(rows, doc, lists, cent)
        assert sh.rows == n and sh.shards == len(devices)
        assert sum(sh.shard_rows(g) for g in range(sh.shards)) == n
        _check(oracle, sh, qs[:3], cent, rows, lists, doc, 8, 10)                    # few queries: per-query kernels
        _check(oracle, sh, qs[:1], cent, rows, lists, doc, 8, 10)                    # one query: the fused kernel per shard
        _check(oracle, sh, qs, cent, rows, lists, doc, 8, 20, sel=range(0, 40, 7))   # a batch: list-major per shard
        _check(oracle, sh, qs[:2], cent, rows, lists, doc, C, 10)                    # every list
        sh.close()


def test_sharded_implicit_ids_and_crowding(vs, oracle):
    """doc_ids = None: the primary key is the document id; and a document whose embeddings sit on different shards."""
    n, d, C = 12000, 384, 12
    rows, cent, lists, doc, q = _crowded_inputs(oracle, n, d, C, 310, 60, 2)
    qs = np.stack([q, oracle.quantize_vector_f32(unit_rows(1, d, 5)[0])])
    for devices in _device_sets()[1:]:
        sh = vs.ivf.ShardedIndex(devices).build_assigned(rows, doc, lists, cent)
        _check(oracle, sh, qs, cent, rows, lists, doc, 5, 10)
        _check(oracle, sh, qs, cent, rows, lists, doc, 5, 32)
        sh.close()
        sh = vs.ivf.ShardedIndex(devices).build_assigned(rows, None, lists, cent)
        _check(oracle, sh, qs, cent, rows, lists, np.arange(n, dtype=np.uint64), 5, 10)
        sh.close()


def test_sharded_upload_equals_rebuild(vs, oracle):
    """Upload (server/upload.go:239-279) on the striped store: same hits as an index built from the table after the upload."""
    n, d, C, n_new = 9000, 768, 24, 1000
    rows, cent, lists, doc = _index_inputs(oracle, n + n_new, d, C, 41)
    doc = np.arange(n + n_new, dtype=np.uint64) + 500
    qs = oracle.quantize_matrix_f32(unit_rows(4, d, 43))
    qs[0] = rows[n + 17]
    for devices in _device_sets()[1:]:
        sh = vs.ivf.ShardedIndex(devices).build_assigned(rows[:n], doc[:n], lists[:n], cent)
        assign = sh.Upload(rows[n:], doc[n:])
        assert (assign == lists[n:].astype(np.int64)).all()
        assert sh.rows == n + n_new
        _check(oracle, sh, qs, cent, rows, lists, doc, 6, 10)
        sh.close()
