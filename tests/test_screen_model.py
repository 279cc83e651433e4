"""CPU check of the error model behind the float32 screen of the tensor-core list scan (csrc/listmajor.cu: dn_row_side,
dn_query_side, dn_skip).  The kernel skips a (query, row) pair only if  t1 - t2 + t3 + E < bound  in float32; that is safe
iff E bounds the distance between the float32 value and the reference's float64 cosine.  Here the same float32 arithmetic is
restated in numpy (one rounding per operation, like the kernel's; the kernel's fused multiply-adds only round less) and
compared with the float64 cosine of the dequantized vectors over the data families the parity tests use, including the
ones built to break the identity (large offsets, tiny and huge ranges, constant and all-zero rows).  Not a test of the
kernel itself -- tests/test_gpu_listmajor.py does that against the oracle -- but of the bound it relies on."""
import numpy as np
import pytest

import oracle
from _util import noop_rows, unit_rows
from test_gpu_listmajor import _hard_rows

F = np.float32


def _sides(rows):
    """Per-row quantities of common.cuh make_side / dn_row_side, float64 where the kernel uses float64."""
    d = rows.shape[1] - 8
    h = np.ascontiguousarray(rows[:, :8]).view(np.float32).astype(np.float64)
    a, R = h[:, 0], h[:, 1] - h[:, 0]
    codes = rows[:, 8:].astype(np.int64)
    s1, s2 = codes.sum(1), (codes * codes).sum(1)
    DA, rs = d * (255.0 * a), R * s1
    ux, Md = DA + rs, np.abs(DA) + np.abs(rs)
    r2i = R * R * (d * s2 - s1 * s1).astype(np.float64)
    u2 = ux * ux
    P = r2i + u2
    return dict(d=d, a=a, R=R, s1=s1, ux=ux, Md=Md, r2i=r2i, u2=u2, P=P)


def _row_side32(s):
    with np.errstate(all="ignore"):
        rP = F(1.0) / np.sqrt(s["P"].astype(F))                      # rsqrtf((float)P)
        eP = F(2) * np.abs(s["r2i"].astype(F)) + F(4) * np.abs(s["ux"].astype(F)) * s["Md"].astype(F) + s["u2"].astype(F) + np.abs(s["P"].astype(F))
        Ay = s["R"].astype(F) * rP
        AyS = Ay * s["s1"].astype(F)
        By = s["ux"].astype(F) * rP
        Myp = s["Md"].astype(F) * rP * F(1.0001)
        mgs = F(255) * (np.abs(s["a"].astype(F)) + np.abs(s["R"].astype(F))) * rP
        Ey = F(1.5e-16) * (eP * rP * rP + F(3.5) * F(s["d"]) * mgs + F(3) * F(s["d"]) + F(16)) + F(4e-6)
    return Ay, AyS, By, Myp, Ey


def _query_side32(s, i):
    with np.errstate(all="ignore"):
        rx = 1.0 / np.sqrt(s["P"][i])
        eP = 2.0 * s["r2i"][i] + 4.0 * abs(s["ux"][i]) * s["Md"][i] + s["u2"][i] + s["P"][i]
        mgs = 255.0 * (abs(s["a"][i]) + abs(s["R"][i])) / np.sqrt(s["P"][i])
        return (F(s["R"][i] * rx * s["d"]), F(s["R"][i] * rx * s["s1"][i]), F(s["ux"][i] * rx), F(F(s["Md"][i] * rx) * F(1.0001)),
                F(1.5e-16) * (F(eP / s["P"][i]) + F(3.5) * F(s["d"]) * F(mgs) + F(3) * F(s["d"]) + F(16)))


def _reference_cosine(qrow, rows):
    def deq(r):
        h = np.ascontiguousarray(r[:, :8]).view(np.float32).astype(np.float64)
        return h[:, :1] + (h[:, 1:2] - h[:, :1]) * r[:, 8:].astype(np.float64) / 255.0
    x, q = deq(rows), deq(qrow[None, :])[0]
    with np.errstate(all="ignore"):
        return (x @ q) / (np.linalg.norm(x, axis=1) * np.linalg.norm(q))


@pytest.mark.parametrize("family", ["unit", "hard", "noop"])
def test_screen_error_term_covers_the_float32_evaluation(family):
    oracle.binding.build()
    d, n = 768, 6000
    if family == "unit":
        rows, qs = oracle.quantize_matrix_f32(unit_rows(n, d, 1)), oracle.quantize_matrix_f32(unit_rows(6, d, 2))
    elif family == "hard":
        rows, qs = oracle.quantize_matrix_f32(_hard_rows(n, d, 5)), oracle.quantize_matrix_f32(_hard_rows(24, d, 7))
    else:
        rows, qs = noop_rows(n, d, 3), noop_rows(6, d, 4)
    rs, qs_s = _sides(rows), _sides(qs)
    Ay, AyS, By, Myp, Ey = _row_side32(rs)
    worst, checked = 0.0, 0
    for i, qrow in enumerate(qs):
        Ax, Cx, Bx, Mxp, Ex = _query_side32(qs_s, i)
        dot = (rows[:, 8:].astype(np.int64) @ qrow[8:].astype(np.int64))
        with np.errstate(all="ignore"):
            t1 = (Ax * Ay) * dot.astype(F)
            t2, t3 = Cx * AyS, Bx * By
            mag = np.abs(t1) + np.abs(t2) + np.abs(t3) + (np.abs(By) * Mxp + np.abs(Bx) * Myp)
            E = F(9.5367431640625e-7) * mag + (Ex + Ey)
            c32 = (t1 - t2) + t3
        ref = _reference_cosine(qrow, rows)
        ok = np.isfinite(c32) & np.isfinite(E) & np.isfinite(ref)        # anything else never skips in the kernel (NaN compare)
        err = np.abs(c32[ok].astype(np.float64) - ref[ok])
        assert (err <= E[ok].astype(np.float64)).all(), f"{family}: query {i}: the screen's error term does not cover the evaluation"
        worst = max(worst, float((err / E[ok].astype(np.float64)).max())) if ok.any() else worst
        checked += int(ok.sum())
    assert checked > n                       # the families are not all degenerate
    assert worst < 0.5, worst                # and the bound has room: the float32 evaluation uses under half of E
