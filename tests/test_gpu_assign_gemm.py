"""GPU parity: nearest-centroid assignment for many centroids through the tensor cores (BASELINE config 4) --
compute/cosine.go:70-125 as called from dnc/k_means.go:75: strict '>' from -1.0, lowest index wins ties, float64
comparison.  Same bits as the oracle, and as the dp4a scan form that serves small centroid counts."""
import numpy as np
import pytest

from _util import f32_bits, noop_rows, unit_rows

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gemm_min(vs):
    """Lower / raise the centroid count from which assignment goes to the tensor cores; restore the default."""
    yield vs.compute.debug_set_argmax_gemm_min
    vs.compute.debug_set_argmax_gemm_min(256)


def _argmax(vs, cent, data, ctx=None):
    return vs.compute.NewMatrix(cent).MatrixCosineSimilarity(vs.compute.NewMatrix(data), ctx=ctx)


@pytest.mark.parametrize("d,m,n", [(768, 300, 3000), (512, 700, 2500), (100, 260, 4000), (1024, 256, 1100)])
def test_assign_gemm_parity_oracle(vs, oracle, d, m, n):
    data = oracle.quantize_matrix_f32(unit_rows(n, d, 3 + d + m))
    cent = oracle.quantize_matrix_f32(unit_rows(m, d, 5 + d))
    cent[:40] = data[np.random.default_rng(m).choice(n, 40, replace=False)]   # rows that coincide with a centroid
    cent[3] = cent[1]            # duplicate centroid: the lowest index must win
    cent[200] = cent[1]
    cent[4, :] = 0               # zero centroid (an empty cluster's mean): scores exactly 0
    data[17, :] = 0              # zero row: every score is 0 -> index 0
    data[18, 8:] = 77            # constant codes
    want_s, want_i = oracle.argmax_MxN(cent, data)
    got_s, got_i = _argmax(vs, cent, data)
    assert (got_i == want_i).all()
    assert (f32_bits(got_s) == f32_bits(want_s)).all()


def test_assign_gemm_noop_rows_and_ties(vs, oracle):
    """noop/ai.go-shaped rows (header -1/+1, uniform bytes); data rows repeated so that equal scores are common."""
    d, m, n = 512, 300, 3000
    cent = noop_rows(m, d, 11)
    data = noop_rows(n, d, 12)
    data[100:200] = data[100]
    cent[50] = cent[20]
    cent[51, 8:] = cent[20, 8:]
    cent[51, 0:8] = np.frombuffer(np.float32([-2, 2]).tobytes(), np.uint8)   # same direction, other header: exact tie in R
    want_s, want_i = oracle.argmax_MxN(cent, data)
    got_s, got_i = _argmax(vs, cent, data)
    assert (got_i == want_i).all()
    assert (f32_bits(got_s) == f32_bits(want_s)).all()


def test_assign_gemm_literal_path(vs, oracle):
    """Certification half-widths inflated: every candidate is re-scored with literal reference arithmetic."""
    d, m, n = 768, 280, 1500
    data = oracle.quantize_matrix_f32(unit_rows(n, d, 21))
    cent = oracle.quantize_matrix_f32(unit_rows(m, d, 22))
    want_s, want_i = oracle.argmax_MxN(cent, data)
    ctx = vs.compute.Context()
    vs.compute.debug_set_certify_scale(1.0e9)
    try:
        got_s, got_i = _argmax(vs, cent, data, ctx=ctx)
    finally:
        vs.compute.debug_set_certify_scale(1.0)
    assert ctx.slowpath_count() > 0
    assert (got_i == want_i).all() and (f32_bits(got_s) == f32_bits(want_s)).all()
    ctx.close()


def test_assign_gemm_matches_scan_form_large(vs, gemm_min):
    """Larger than the oracle handles quickly, several GEMM batches, a store small enough that the work is cut into
    (store tile, query range) items: the tensor-core path must return the scan form's indices and similarities (the
    scan form is pinned to the oracle by test_gpu_compute.py)."""
    d, m, n = 768, 2048, 100000
    data = vs.compute.NewMatrix(vs.compute.QuantizeMatrixFloat32(unit_rows(n, d, 31)))
    cent = vs.compute.NewMatrix(vs.compute.QuantizeMatrixFloat32(unit_rows(m, d, 32)))
    ctx = vs.compute.Context()
    l0 = ctx.launch_count()
    s1, i1 = cent.MatrixCosineSimilarity(data, ctx=ctx)
    gemm_launches = ctx.launch_count() - l0
    gemm_min(1 << 30)
    l0 = ctx.launch_count()
    s2, i2 = cent.MatrixCosineSimilarity(data, ctx=ctx)
    assert ctx.launch_count() - l0 < gemm_launches      # the scan form is two launches; the GEMM path several per batch
    assert (i1 == i2).all() and (f32_bits(s1) == f32_bits(s2)).all()
    ctx.close()


def test_kmeans_step_many_centroids(vs, oracle):
    """One Lloyd iteration (dnc/k_means.go:67-117) with enough centroids for the tensor-core assignment."""
    d, n, k = 768, 6000, 300
    data = oracle.quantize_matrix_f32(unit_rows(n, d, 41))
    cent = data[np.random.default_rng(4).choice(n, k, replace=False)].copy()
    cent[2] = cent[1]
    m = vs.compute.NewMatrix(data)
    means_o = np.zeros((k, d), np.float32)
    means_g = np.zeros((k, d), np.float32)
    cent_o, cent_g = cent.copy(), cent.copy()
    for it in range(2):
        a_o, c_o, new_o, conv_o = oracle.kmeans_step(data, cent_o, means_o)
        a_g, c_g, new_g, conv_g = vs.dnc.KMeansStep(m, cent_g, means_g)
        assert (a_g == a_o).all(), f"assign differs at iteration {it}"
        assert (c_g == c_o).all() and (new_g == new_o).all() and conv_g == conv_o
        assert (f32_bits(means_g) == f32_bits(means_o)).all()
        cent_o, cent_g = new_o, new_g
