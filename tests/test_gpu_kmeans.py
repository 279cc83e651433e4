"""GPU parity: one Lloyd step (dnc/k_means.go:67-117) and recenter (dnc/dnc.go:417-449), byte-exact."""
import numpy as np
import pytest

from _util import f32_bits, unit_rows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,n,k", [(768, 20000, 5), (768, 5000, 25), (512, 3000, 2), (64, 2000, 10), (768, 300, 1)])
def test_kmeans_steps_parity(vs, oracle, d, n, k):
    data = oracle.quantize_matrix_f32(unit_rows(n, d, 50 + k))
    rng = np.random.default_rng(k)
    cent = data[rng.choice(n, k, replace=False)].copy()
    if k >= 10:
        cent[2] = cent[1]          # duplicate -> cluster 2 stays empty and keeps its previous (zero) mean
    m = vs.compute.NewMatrix(data)
    means_o = np.zeros((k, d), np.float32)
    means_g = np.zeros((k, d), np.float32)
    cent_o, cent_g = cent.copy(), cent.copy()
    for it in range(3):
        a_o, c_o, new_o, conv_o = oracle.kmeans_step(data, cent_o, means_o)
        a_g, c_g, new_g, conv_g = vs.dnc.KMeansStep(m, cent_g, means_g)
        assert (a_g == a_o).all(), f"assign differs at iteration {it}"
        assert (c_g == c_o).all()
        assert (f32_bits(means_g) == f32_bits(means_o)).all()
        assert (new_g == new_o).all(), f"new centroid bytes differ at iteration {it}"
        assert conv_g == conv_o
        cent_o, cent_g = new_o, new_g


@pytest.mark.parametrize("d,n", [(768, 4000), (100, 999), (768, 1)])
def test_recenter_parity(vs, oracle, d, n):
    rows = oracle.quantize_matrix_f32(unit_rows(n, d, 5))
    assert (vs.dnc.Recenter(vs.compute.NewMatrix(rows)) == oracle.recenter(rows)).all()


@pytest.mark.parametrize("d,n,k", [(768, 4000, 5), (256, 3000, 3), (64, 1500, 8)])
def test_kmeans_full_loop_parity(vs, oracle, d, n, k):
    """The whole kMeans (k_means.go:19-212) given the superset draw: superset phase, truncation to the first k, set phase."""
    data = oracle.quantize_matrix_f32(unit_rows(n, d, 70 + k))
    rows = np.random.default_rng(k).choice(n, min(n, 5 * k), replace=False)
    want, it1, it2 = oracle.kmeans(data, k, rows, limit=12)
    got, st = vs.dnc.KMeans(vs.compute.NewMatrix(data), k, superset_rows=rows, iter_limit=12, want_stats=True)
    assert (st["superset_iterations"], st["set_iterations"]) == (it1, it2)
    assert (got == want).all()


def test_kmeans_full_loop_tensor_core_assign(vs, oracle):
    """Same, with the centroid counts of both phases above the tensor-core assignment threshold."""
    d, n, k = 512, 5000, 20
    data = oracle.quantize_matrix_f32(unit_rows(n, d, 91))
    rows = np.random.default_rng(9).choice(n, 5 * k, replace=False)
    want, it1, it2 = oracle.kmeans(data, k, rows, limit=4)
    vs.compute.debug_set_argmax_gemm_min(16)
    try:
        got, st = vs.dnc.KMeans(vs.compute.NewMatrix(data), k, superset_rows=rows, iter_limit=4, want_stats=True)
    finally:
        vs.compute.debug_set_argmax_gemm_min(256)
    assert (st["superset_iterations"], st["set_iterations"]) == (it1, it2)
    assert (got == want).all()


def test_kmeans_early_returns(vs):
    """k <= 0 -> nil; len(data) <= k -> data unchanged (k_means.go:20-26)."""
    data = vs.compute.QuantizeMatrixFloat32(unit_rows(4, 64, 1))
    m = vs.compute.NewMatrix(data)
    assert vs.dnc.KMeans(m, 0) is None
    assert (vs.dnc.KMeans(m, 4) == data).all() and (vs.dnc.KMeans(m, 9) == data).all()


@pytest.mark.parametrize("d,n,k,blocks", [(768, 9000, 7, 3), (256, 5000, 40, 4), (768, 2500, 300, 2)])
def test_kmeans_row_blocks_relay_matches_one_device(vs, oracle, d, n, k, blocks):
    """The store cut into contiguous row blocks (one per GPU in production, shard.kmeans_step_relay): assigning block by
    block and continuing the float32 sums in row order gives the oracle's single-device bits (here the blocks are
    separate matrices on one GPU, processed in order)."""
    import torch
    data = oracle.quantize_matrix_f32(unit_rows(n, d, 30 + k))
    cent = data[np.random.default_rng(k).choice(n, k, replace=False)].copy()
    cent[1] = cent[0]                                  # an empty cluster keeps its previous mean
    means_o = np.full((k, d), 0.25, np.float32)
    dev = torch.device("cuda", 0)
    ctx = vs.compute.default_context()
    d_means = torch.from_numpy(means_o.copy()).to(dev)
    cmat = vs.compute.NewMatrix(cent)
    cent_o = cent
    bounds = [vs.shard.block_range(n, r, blocks) for r in range(blocks)]
    mats = [vs.compute.NewMatrix(data[lo:hi]) for lo, hi in bounds]
    for it in range(2):
        a_o, c_o, new_o, conv_o = oracle.kmeans_step(data, cent_o, means_o)
        sums = torch.zeros(k * d, dtype=torch.float32, device=dev)
        counts = torch.zeros(k, dtype=torch.int64, device=dev)
        assigns = []
        for (lo, hi), m in zip(bounds, mats):
            a = torch.empty(hi - lo, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            cmat.ArgmaxDev(m, a.data_ptr(), ctx=ctx)
            vs.dnc.KMeansAccumulateDev(m, k, a.data_ptr(), sums.data_ptr(), counts.data_ptr(), ctx=ctx)
            ctx.sync()
            assigns.append(a.cpu().numpy())
        cmat, conv = vs.dnc.KMeansFinishDev(cmat, sums.data_ptr(), counts.data_ptr(), d_means.data_ptr(), ctx=ctx)
        assert (np.concatenate(assigns) == a_o).all(), f"assign differs at iteration {it}"
        assert (counts.cpu().numpy() == c_o).all()
        assert (f32_bits(d_means.cpu().numpy()) == f32_bits(means_o)).all()
        assert (cmat.ReadRows() == new_o).all() and conv == conv_o
        cent_o = new_o


def test_kmeans_step_and_recenter_known_answers(vs):
    """The hand-derived known answers of tests/golden/kat.json (plain Python IEEE arithmetic following the Go lines,
    tests/golden/make_golden.py -- no oracle involved) through the CUDA path: one Lloyd iteration
    (dnc/k_means.go:67-117) and recenterDbCentroid (dnc/dnc.go:417-449)."""
    import json
    import os
    import struct
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))
    c = kat["kmeans_step"]
    cent = np.array(c["centroids"], np.uint8)
    data = np.array(c["data"], np.uint8)
    means = np.array(c["prev_means"], np.float32)
    assign, counts, newc, conv = vs.dnc.KMeansStep(vs.compute.NewMatrix(data), cent, means)
    assert assign.tolist() == c["assign"] and counts.tolist() == c["counts"], c["why"]
    assert [[struct.pack(">f", v).hex() for v in m] for m in means] == c["means_f32"], c["why"]
    assert newc.tolist() == c["new_centroids"] and conv == c["converged"], c["why"]
    r = kat["recenter"]
    rows = np.array(r["rows"], np.uint8)
    assert vs.dnc.Recenter(vs.compute.NewMatrix(rows)).tolist() == r["expect"], r["why"]
