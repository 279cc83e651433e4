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
