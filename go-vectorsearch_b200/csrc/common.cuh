// common.cuh -- shared device helpers for libvscuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vs {

constexpr int kWarp = 32;
constexpr double kU = 1.1102230246251565e-16;  // 2^-53, float64 unit roundoff

// Device-resident quantized matrix: codes[n][d_pad] (zero padded to a 16-byte multiple), the
// reference's 8-byte header per row (float min, float max; compute/quantization.go:82-91) and the
// two integer sums the cosine identity needs (sum v, sum v^2).
struct MatView {
    const uint8_t *codes;
    const float2 *hdr;
    const uint2 *sums;
    size_t n;
    int d;
    int d_pad;
};

// A top-k candidate. skey = order-preserving image of the float32 similarity (0 = empty slot,
// 1 = NaN which cmp.Compare sorts below every number); meta bit31 = "float32 rounding of this score
// could not be certified against the reference's float64 arithmetic".
struct Cand {
    uint32_t skey;
    uint32_t meta;
    uint64_t id;
};
constexpr uint32_t kFlagBit = 0x80000000u;
// meta bit30 (with explicit document ids only): another row of the same document with the same float32 score, equally
// uncertain, was dropped in favour of this one (one hit per document, server/search.go:260-268).  If the literal
// re-score moves this row down, the dropped one may have been the document's best: the query is then finished by the
// literal-arithmetic kernel.  Rows per device store are < 2^30 (a 180 GB GPU holds ~2.3e8 rows of 784 B).
constexpr uint32_t kSibBit = 0x40000000u;
constexpr uint32_t kMetaRowMask = 0x3FFFFFFFu;
constexpr uint64_t kEmptyId = 0xFFFFFFFFFFFFFFFFull;

__host__ __device__ inline uint32_t f32_to_key(float f) {
    if (f != f) return 1u;
    if (f == 0.0f) f = 0.0f;  // -0 == +0 under cmp.Compare
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float key_to_f32(uint32_t k) {
    uint32_t u;
    if (k == 1u) u = 0x7FC00000u;
    else u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    float f;
#ifdef __CUDA_ARCH__
    f = __uint_as_float(u);
#else
    memcpy(&f, &u, 4);
#endif
    return f;
}

// a ranks strictly before b: similarity desc, then id asc (the tie rule of this repo's contract).
__device__ __forceinline__ bool cand_better(uint32_t ak, uint64_t aid, uint32_t bk, uint64_t bid) {
    return ak > bk || (ak == bk && aid < bid);
}

// 128-bit streaming load: read-only path, do not allocate in L1 (each row byte is used once).
__device__ __forceinline__ uint4 ld_stream_u4(const void *p) {
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// TMA bulk prefetch of a contiguous, 16-byte aligned span into L2 (no shared memory, no completion to wait on).
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c) { return __dp4a(a, b, c); }

__device__ __forceinline__ uint32_t dot16(const uint4 &a, const uint4 &b, uint32_t acc) {
    acc = dp4a_u(a.x, b.x, acc);
    acc = dp4a_u(a.y, b.y, acc);
    acc = dp4a_u(a.z, b.z, acc);
    acc = dp4a_u(a.w, b.w, acc);
    return acc;
}

// ---- literal reference arithmetic (never contracted, IEEE rn), operation for operation ----
// compute/quantization.go:63-69 DequantizeFloat64
__device__ __forceinline__ double ref_dequant_f64(uint32_t q, double mn, double range) {
    double normalized = __ddiv_rn((double)q, 255.0);
    return __dadd_rn(mn, __dmul_rn(normalized, range));
}
// compute/quantization.go:55-61 DequantizeFloat32
__device__ __forceinline__ float ref_dequant_f32(uint32_t q, float mn, float range) {
    float normalized = __fdiv_rn((float)q, 255.0f);
    return __fadd_rn(mn, __fmul_rn(normalized, range));
}
// Go uint8(float32) on amd64 (CVTTSS2SL, low byte; NaN / out of range -> 0x80000000 -> 0)
__device__ __forceinline__ uint32_t go_u8_from_f32(float x) {
    int i;
    if (x != x || x >= 2147483648.0f || x < -2147483648.0f) i = (int)0x80000000;
    else i = __float2int_rz(x);
    return (uint32_t)i & 0xFFu;
}
__device__ __forceinline__ uint32_t go_u8_from_f64(double x) {
    long long i;
    if (x != x || x >= 9223372036854775808.0 || x < -9223372036854775808.0) i = (long long)0x8000000000000000ull;
    else i = __double2ll_rz(x);
    return (uint32_t)((unsigned long long)i & 0xFFull);
}
// compute/quantization.go:21-45 QuantizeFloat32 / QuantizeFloat64
__device__ __forceinline__ uint32_t ref_quant_f32(float v, float mn, float mx) {
    if (v < mn) v = mn;
    else if (v > mx) v = mx;
    float normalized = __fdiv_rn(__fsub_rn(v, mn), __fsub_rn(mx, mn));
    return go_u8_from_f32(__fmul_rn(normalized, 255.0f));
}
__device__ __forceinline__ uint32_t ref_quant_f64(double v, double mn, double mx) {
    if (v < mn) v = mn;
    else if (v > mx) v = mx;
    double normalized = __ddiv_rn(__dsub_rn(v, mn), __dsub_rn(mx, mn));
    return go_u8_from_f64(__dmul_rn(normalized, 255.0));
}

// ---- the certified integer-identity score ------------------------------------------------------
// With x_i = a + q_i*R/255 and y_i = b + v_i*S/255 (the values the reference dequantizes to; R = a'-a,
// S = b'-b in float64 exactly as the reference widens its float32 header) and A = 255a, B = 255b:
//   ux = D*A + R*sum(q)                      ( = 255*D*mean(x) )
//   Ix = D*sum(q^2) - sum(q)^2               ( exact int64: D^2 * variance of the codes )
//   I1 = D*sum(q*v) - sum(q)*sum(v)          ( exact int64: D^2 * covariance of the codes )
//   255^2*D * x.y   = R*S*I1 + ux*uy
//   255^2*D * |x|^2 = R^2*Ix + ux^2          ( two non-negative terms: no cancellation )
//   cos = (R*S*I1 + ux*uy) * rsqrt((R^2*Ix + ux^2)(S^2*Iy + uy^2))
// The cancellation between the large "offset" terms of the naive expansion happens in integers, exactly.
// Only sum(q*v) depends on both vectors; it is an exact uint8 dot product.
struct SideConst {
    double R;      // max - min
    double ux;     // D*255*min + R*sum
    double P;      // R^2*I + ux^2  = 255^2 * D * |x|^2
    double epP;    // (rounding-error coefficient of P) / P, units of u
    double mgs;    // 255*(|min| + |R|) / sqrt(P): reference dequantization error relative to |x| (times 1/sqrt(D))
    float M;       // |D*A| + |R*sum|, rounded up (error scale of ux)
    long long s1;  // sum of codes
    int trivial_zero;  // every dequantized value is exactly 0 (min == 0 and (R == 0 or all codes 0))
};

__device__ __forceinline__ SideConst make_side(float mn, float mx, uint32_t s1, uint32_t s2, int D) {
    SideConst c;
    const double a = (double)mn;
    c.R = (double)mx - a;
    const double DA = (double)D * (255.0 * a);  // exact: 12 + 32 significant bits
    const double rs = c.R * (double)s1;
    c.ux = DA + rs;
    const double Md = fabs(DA) + fabs(rs);
    const long long I = (long long)D * (long long)s2 - (long long)s1 * (long long)s1;
    const double r2i = c.R * c.R * (double)I;
    const double u2 = c.ux * c.ux;
    c.P = r2i + u2;
    const double eP = 2.0 * r2i + 4.0 * fabs(c.ux) * Md + u2 + c.P;
    c.epP = eP / c.P;
    c.mgs = 255.0 * (fabs(a) + fabs(c.R)) / sqrt(c.P);
    c.M = __double2float_ru(Md);
    c.s1 = (long long)s1;
    c.trivial_zero = (mn == 0.0f) && (c.R == c.R) && (fabs(c.R) < 1.0e300) && (c.R == 0.0 || s2 == 0u);
    return c;
}

// Test hook (vs_debug_set_certify_scale): inflates every certification half-width so that the parity tests can
// drive all rows through the literal-arithmetic paths.  1.0 in production.  One copy per translation unit.
static __constant__ float c_certify_scale = 1.0f;

// Error model (DESIGN.md "certified scores"), all in units of u = 2^-53:
//   ours      eN/den + 0.5|c| (ePx/Px + ePy/Py) + 4       (products / sums of the identity, rsqrt, final product)
//   reference (D+4) + |c| (D+2)                            (sequential dot; sequential norms, sqrt, divides)
//             3.5 * D * (mag_x/sqrt(Px) + mag_y/sqrt(Py))  (three roundings per dequantized element)
//   delta = 1.25 * u * (ours + reference)
// The delta terms only need to be upper bounds: they are evaluated in float32 (with a 1.001 inflation for the
// float32 roundings).  Overflow, underflow or NaN anywhere makes delta non-finite or huge: flagged, never
// mis-certified.  Measured |ours - reference| / delta <= 0.11 over seven data families.
__device__ __forceinline__ bool score_core(const SideConst &x, double yR, double yux, double yP, float yM, float yepP,
                                           float ymgs, long long ys1, int ytz, uint32_t dot_qv, int D, double *c_out,
                                           double *delta_out) {
    if (x.trivial_zero || ytz) {  // the reference leaves an all-zero vector unnormalized: dot = +0
        *c_out = 0.0;
        *delta_out = 0.0;
        // the other side must still be finite (NaN * 0 = NaN in the reference)
        return (x.P == x.P) && (yP == yP) && fabs(x.P) < 1.0e300 && fabs(yP) < 1.0e300;
    }
    const long long I1 = (long long)D * (long long)dot_qv - x.s1 * ys1;
    const double n1 = x.R * yR * (double)I1;
    const double n2 = x.ux * yux;
    const double N = n1 + n2;
    const double pp = x.P * yP;
    const double r = rsqrt(pp);
    const double c = N * r;
    const float rf = __double2float_ru(r);
    const float eN = 2.0f * fabsf((float)n1) + 2.0f * (fabsf((float)yux) * x.M + fabsf((float)x.ux) * yM) +
                     fabsf((float)n2) + fabsf((float)N);
    const float ac = fabsf((float)c);
    const float e = eN * rf + 0.5f * ac * ((float)x.epP + yepP) + 4.0f + (float)(D + 4) + ac * (float)(D + 2) +
                    3.5f * (float)D * ((float)x.mgs + ymgs);
    const double delta = (1.25 * 1.001 * kU) * (double)(e * c_certify_scale);
    *c_out = c;
    *delta_out = delta;
    return (x.P > 0.0) && (yP > 0.0) && (delta < 1.0e-6);  // false for NaN / Inf as well
}

// Row side evaluated per scanned row (no float64 divide or sqrt: float32 reciprocal sqrt for the bound terms).
__device__ __forceinline__ float score_fast(const SideConst &x, float mn, float mx, uint32_t s1, uint32_t s2, uint32_t dot_qv,
                                            int D, bool *flag) {
    const double a = (double)mn;
    const double R = (double)mx - a;
    const double DA = (double)D * (255.0 * a);
    const double rs = R * (double)s1;
    const double ux = DA + rs;
    const double Md = fabs(DA) + fabs(rs);
    const long long I = (long long)D * (long long)s2 - (long long)s1 * (long long)s1;
    const double r2i = R * R * (double)I;
    const double u2 = ux * ux;
    const double P = r2i + u2;
    const float rP = rsqrtf(__double2float_rd(P));  // ~ 1/sqrt(P), rounded toward larger
    const float eP = 2.0f * (float)r2i + 4.0f * fabsf((float)ux) * __double2float_ru(Md) + (float)u2 + (float)P;
    const int tz = (mn == 0.0f) && (R == R) && (fabs(R) < 1.0e300) && (R == 0.0 || s2 == 0u);
    double c, delta;
    const bool ok = score_core(x, R, ux, P, __double2float_ru(Md), eP * rP * rP, 255.0f * (fabsf(mn) + fabsf((float)R)) * rP,
                               (long long)s1, tz, dot_qv, D, &c, &delta);
    if (!ok) {
        *flag = true;
        return 2.0f;
    }
    const float lo = __double2float_rn(c - delta);
    const float hi = __double2float_rn(c + delta);
    *flag = (lo != hi);
    return hi;
}

// Value and rigorous half-width for comparisons done in float64 (compute/cosine.go:114 `dot > maxVal`).
// ok=false -> needs the literal path.
__device__ __forceinline__ bool score_interval(const SideConst &x, const SideConst &y, uint32_t dot_qv, int D, double sqrtD,
                                               double *c_out, double *delta_out) {
    (void)sqrtD;
    return score_core(x, y.R, y.ux, y.P, y.M, (float)y.epP, (float)y.mgs, y.s1, y.trivial_zero, dot_qv, D, c_out, delta_out);
}

// Ordered sum of D terms computed in parallel by the warp: lane L owns elements base+8L .. base+8L+7 of every
// 256-element chunk; the additions run in element order (lane after lane), so the result equals a sequential
// `for i { acc += term(i) }` bit for bit while the terms themselves are evaluated 32-wide.
template <typename F>
__device__ __forceinline__ double warp_ordered_sum(int D, int lane, F term) {
    double acc = 0.0;
    for (int base = 0; base < D; base += 256) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int i = base + lane * 8 + j;
            v[j] = i < D ? term(i) : 0.0;
        }
        const int lanes = min(32, (D - base + 7) >> 3);
        for (int L = 0; L < lanes; L++) {
            if (lane == L) {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (base + L * 8 + j < D) acc = __dadd_rn(acc, v[j]);
            }
            acc = __shfl_sync(0xFFFFFFFFu, acc, L);
        }
    }
    return acc;
}

// Literal reference cosine of one row (compute/cosine.go:29-33,43-50,138-149) by a whole warp; qn = the query
// after normalizeVector. Every lane returns the same value.
__device__ __forceinline__ double warp_ref_cosine_row_f64(const uint8_t *codes, float mn_f, float mx_f, const double *qn, int D,
                                                          int lane) {
    const double mn = (double)mn_f, range = __dsub_rn((double)mx_f, (double)mn_f);
    const double normsq = warp_ordered_sum(D, lane, [&](int i) {
        const double x = ref_dequant_f64(codes[i], mn, range);
        return __dmul_rn(x, x);
    });
    const double norm = __dsqrt_rn(normsq);
    return warp_ordered_sum(D, lane, [&](int i) {
        double x = ref_dequant_f64(codes[i], mn, range);
        if (norm != 0.0) x = __ddiv_rn(x, norm);
        return __dmul_rn(qn[i], x);
    });
}

// Literal reference cosine of one row against a pre-normalized query (compute/cosine.go:29-33,43-50,
// 138-149).  qn = the query after normalizeVector (float64[d]).  Sequential, uncontracted.
__device__ inline float ref_cosine_row(const uint8_t *codes, float mn_f, float mx_f, const double *qn, int d) {
    double mn = (double)mn_f, mx = (double)mx_f;
    double range = __dsub_rn(mx, mn);
    double norm = 0.0;
    for (int i = 0; i < d; i++) {
        double x = ref_dequant_f64(codes[i], mn, range);
        norm = __dadd_rn(norm, __dmul_rn(x, x));
    }
    norm = __dsqrt_rn(norm);
    double dot = 0.0;
    for (int i = 0; i < d; i++) {
        double x = ref_dequant_f64(codes[i], mn, range);
        if (norm != 0.0) x = __ddiv_rn(x, norm);
        dot = __dadd_rn(dot, __dmul_rn(qn[i], x));
    }
    return __double2float_rn(dot);
}
__device__ inline double ref_cosine_row_f64(const uint8_t *codes, float mn_f, float mx_f, const double *qn, int d) {
    double mn = (double)mn_f, mx = (double)mx_f;
    double range = __dsub_rn(mx, mn);
    double norm = 0.0;
    for (int i = 0; i < d; i++) {
        double x = ref_dequant_f64(codes[i], mn, range);
        norm = __dadd_rn(norm, __dmul_rn(x, x));
    }
    norm = __dsqrt_rn(norm);
    double dot = 0.0;
    for (int i = 0; i < d; i++) {
        double x = ref_dequant_f64(codes[i], mn, range);
        if (norm != 0.0) x = __ddiv_rn(x, norm);
        dot = __dadd_rn(dot, __dmul_rn(qn[i], x));
    }
    return dot;
}

}  // namespace vs
