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
// With x_i = a + q_i*(a'-a)/255 and y_i = b + v_i*(b'-b)/255 (the values the reference dequantizes to),
//   255^2 * x.y   = D*A*B + A*S*sum(v) + B*R*sum(q) + R*S*sum(q*v),  A=255a, R=a'-a, B=255b, S=b'-b
//   255^2 * |y|^2 = D*B^2 + 2*B*S*sum(v) + S^2*sum(v^2)
// Only sum(q*v) depends on both vectors; it is an exact uint8 dot product.
struct SideConst {   // per vector: A, R and the derived norm terms
    double A, R;     // 255*min, max-min (float64 of the float32 header, as the reference widens it)
    double sum, sumsq;  // sum of codes, sum of squared codes
    double P;        // 255^2 * |x|^2
    double T;        // sum of |terms| of P (cancellation measure)
    double sqrtP;
    double tp;       // 8*T/P: relative rounding error of P in units of u
    double mg;       // 8*255*(|a|+|a'-a|)/sqrt(P): reference dequantization error relative to |x|, per sqrt(D)
};

__device__ __forceinline__ SideConst make_side(float mn, float mx, uint32_t s1, uint32_t s2, int D) {
    SideConst c;
    double a = (double)mn, ap = (double)mx;
    c.A = 255.0 * a;
    c.R = ap - a;
    c.sum = (double)s1;
    c.sumsq = (double)s2;
    double p1 = (double)D * c.A * c.A;
    double p2 = 2.0 * c.A * c.R * c.sum;
    double p3 = c.R * c.R * c.sumsq;
    c.P = p1 + p2 + p3;
    c.T = p1 + fabs(p2) + p3;
    c.sqrtP = sqrt(c.P);
    // magnitude bound of the reference's dequantization intermediates relative to |x|:
    // (|a| + |a'-a|) / |x| = 255*(|a|+|R|)/sqrt(P)
    double mag = 255.0 * (fabs(a) + fabs(c.R));
    c.tp = 8.0 * (c.T / c.P);
    c.mg = 8.0 * mag / c.sqrtP;
    return c;
}

// Per-row side of the identity, without the slow float64 divide / sqrt: only what score_fast needs.
struct RowSide {
    double A, R, sum;  // 255*min, max-min, sum of codes
    double P;          // 255^2 * |y|^2
    float Tf;          // sum of |terms| of P, rounded up
    float magf;        // 255*(|min| + |max-min|), rounded up
};

__device__ __forceinline__ RowSide make_row_side(float mn, float mx, uint32_t s1, uint32_t s2, int D) {
    RowSide c;
    double a = (double)mn, ap = (double)mx;
    c.A = 255.0 * a;
    c.R = ap - a;
    c.sum = (double)s1;
    double p1 = (double)D * c.A * c.A;
    double p2 = 2.0 * c.A * c.R * c.sum;
    double p3 = c.R * c.R * (double)s2;
    c.P = p1 + p2 + p3;
    c.Tf = __double2float_ru(p1 + fabs(p2) + p3);
    c.magf = __double2float_ru(255.0 * (fabs(a) + fabs(c.R)));
    return c;
}

// Returns the float32 similarity; *flag = the float32 rounding (or the value) cannot be certified
// to equal the reference's float32(dot) -- such rows are recomputed with literal arithmetic.
// Error model (DESIGN.md "certified scores"): |ours - reference| <= delta with
//   delta = 2u * ( 8*T/den + 8*Tx/Px + 8*Ty/Py            (our own roundings, cancellation-aware)
//                + (2D+16)                                  (reference: sequential norm, divide, dot)
//                + 8*sqrt(D)*(mag_x/|x| + mag_y/|y|) )      (reference: dequantization roundings)
// The score itself is N * rsqrt(Px*Py) in float64 (one rsqrt, no divide); the delta terms only need to be
// upper bounds, so they are evaluated in float32 with directed rounding / a 1.001 inflation.  Overflow,
// underflow or NaN anywhere makes delta non-finite or huge and the row is flagged: never mis-certified.
__device__ __forceinline__ float score_fast(const SideConst &x, const RowSide &y, uint32_t dot_qv, int D, float sqrtDf,
                                            bool *flag) {
    double t1 = (double)D * x.A * y.A;
    double t2 = x.A * y.R * y.sum;
    double t3 = y.A * x.R * x.sum;
    double t4 = x.R * y.R * (double)dot_qv;
    double N = (t1 + t2) + (t3 + t4);
    float Tf = __double2float_ru(fabs(t1) + fabs(t2) + fabs(t3) + fabs(t4));
    if (x.T == 0.0 || y.Tf == 0.0f) {  // an exactly-zero vector (header 0/0 or all terms 0): the reference leaves it
        bool finite = (x.T == x.T) && (y.Tf == y.Tf) && (Tf == Tf) && Tf < 3.0e38f;  // unnormalized, dot = +0
        *flag = !finite;
        return finite ? 0.0f : 2.0f;
    }
    double pp = x.P * y.P;
    double c = N * rsqrt(pp);
    float ppf = __double2float_rd(pp);
    float rden = rsqrtf(ppf);                                  // ~ 1/den
    float ry = rsqrtf(__double2float_rd(y.P));                 // ~ 1/sqrt(Py)
    float e = 8.0f * (Tf * rden) + (float)x.tp + 8.0f * (y.Tf * ry * ry) + (float)(2 * D + 16) +
              sqrtDf * ((float)x.mg + 8.0f * (y.magf * ry));
    double delta = (2.0 * 1.001 * kU) * (double)e;
    bool ok = (x.P > 0.0) && (y.P > 0.0) && (delta < 1.0e-6);   // false for NaN / Inf as well
    float lo = __double2float_rn(c - delta);
    float hi = __double2float_rn(c + delta);
    if (!ok) {
        *flag = true;
        return 2.0f;
    }
    *flag = (lo != hi);
    return hi;
}

// Same model without the float32 rounding: value and rigorous half-width, for argmax comparisons
// done in float64 (compute/cosine.go:114 `dot > maxVal`). ok=false -> needs the literal path.
__device__ __forceinline__ bool score_interval(const SideConst &x, const SideConst &y, uint32_t dot_qv, int D,
                                               double sqrtD, double *c_out, double *delta_out) {
    double t1 = (double)D * x.A * y.A;
    double t2 = x.A * y.R * y.sum;
    double t3 = y.A * x.R * x.sum;
    double t4 = x.R * y.R * (double)dot_qv;
    double N = (t1 + t2) + (t3 + t4);
    double T = fabs(t1) + fabs(t2) + fabs(t3) + fabs(t4);
    double fin = T + x.T + y.T;
    if (!(fin < 1.0e300)) return false;
    if (x.T == 0.0 || y.T == 0.0) {
        *c_out = 0.0;
        *delta_out = 0.0;
        return true;
    }
    double den = x.sqrtP * y.sqrtP;
    double c = N / den;
    double delta = 2.0 * kU * (8.0 * (T / den) + x.tp + y.tp + (double)(2 * D + 16) + sqrtD * (x.mg + y.mg));
    *c_out = c;
    *delta_out = delta;
    return (x.P > 0.0) && (y.P > 0.0) && (delta < 1.0e-6);
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
