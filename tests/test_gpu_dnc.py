"""GPU parity: the divide-and-conquer centroid build (dnc/dnc.go:300-400 and the tail of KMeansDivideAndConquer, :177-291)
with every dataset resident in HBM.  Upstream the build is seeded from the clock and runs its children concurrently; the
device driver (concurrent too: one host thread + CUDA stream per worker) and the oracle's restatement give every node of the
tree its own generator, spawned from its parent's, so their outputs can be compared bit for bit whatever the schedule."""
import numpy as np
import pytest

from _util import clustered_rows, noop_rows, unit_rows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("workers", [1, 6])
@pytest.mark.parametrize("d,n,target,sample,seed", [(256, 6000, 500, 2000, 1), (768, 4000, 900, 1500, 2), (64, 3000, 200, 5000, 3)])
def test_divide_and_conquer_matches_oracle(vs, oracle, d, n, target, sample, seed, workers):
    x, _ = clustered_rows(n, d, 7, seed)
    rows = oracle.quantize_matrix_f32(x)
    want = oracle.divide_and_conquer(rows, target, sample, 5, np.random.default_rng(seed), limit=6)
    got = vs.dnc.DivideAndConquer(vs.compute.NewMatrix(rows), target_size=target, sample_size=sample, split_size=5,
                                  rng=np.random.default_rng(seed), iter_limit=6, workers=workers)
    assert got.shape == want.shape and (got == want).all()
    assert got.shape[0] >= n // target      # enough leaves that none can exceed the target on average


def test_split_is_a_stable_partition(vs, oracle):
    d, n, k = 768, 5000, 5
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 4))
    cent = rows[[10, 20, 30, 40, 50]].copy()
    cent[3] = cent[1]                        # a duplicate centroid: its child receives nothing
    _, idx = oracle.argmax_MxN(cent, rows)
    children = vs.dnc.Split(vs.compute.NewMatrix(rows), cent)
    assert len(children) == k and children[3] is None
    for j, ch in enumerate(children):
        want = rows[idx == j]
        if ch is None:
            assert want.shape[0] == 0
        else:
            assert (ch.ReadRows() == want).all()


def test_reassign_recenter_matches_oracle(vs, oracle):
    d, n, k = 768, 6000, 40
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 5))
    cent = oracle.quantize_matrix_f32(unit_rows(k, d, 6))
    cent[7] = cent[2]                        # never nearest: an empty cluster (0/0 -> NaN -> zero codes, dnc.go:437-441)
    a_o, c_o, n_o = oracle.reassign_recenter(rows, cent)
    a_g, c_g, n_g = vs.dnc.ReassignRecenter(vs.compute.NewMatrix(rows), cent)
    assert (a_g == a_o).all() and (n_g == n_o).all()
    assert n_o[7] == 0
    assert (c_g == c_o).all()


def test_recenter_noop_rows_many_clusters(vs, oracle):
    """Arbitrary headers, more clusters than the tensor-core assignment threshold."""
    d, n, k = 256, 5000, 300
    rows = noop_rows(n, d, 8)
    cent = rows[np.random.default_rng(1).choice(n, k, replace=False)].copy()
    a_o, c_o, n_o = oracle.reassign_recenter(rows, cent)
    a_g, c_g, n_g = vs.dnc.ReassignRecenter(vs.compute.NewMatrix(rows), cent)
    assert (a_g == a_o).all() and (n_g == n_o).all() and (c_g == c_o).all()
