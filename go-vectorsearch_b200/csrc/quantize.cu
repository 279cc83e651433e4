// quantize.cu -- the 1-byte min/max codec (K1/K2 of SURVEY.md 2b) and the row-format movers.
//
// Replaces compute/quantization.go: QuantizeVectorFloat32/64 (:82-102), QuantizeMatrixFloat32/64
// (:142-156), DequantizeVector/MatrixFloat32/64 (:114-132,:166-180), rangeFloat32/64 (:194-216).
// All arithmetic is IEEE round-to-nearest, uncontracted (see common.cuh ref_* helpers), so codes and
// headers are byte-identical to the reference's.
//
// Row formats: "row776" AoS = [f32 min][f32 max][d codes] (8+d bytes, the reference's stored vector);
// device SoA = codes[n][d_pad] + hdr[n] + sums[n] (sum v, sum v^2), d_pad = d rounded up to 16.
// All kernels are one warp per row, HBM-bound: quantize reads 4d (8d) bytes and writes 8+d.
#include "internal.h"

namespace vs {

#define FULL 0xFFFFFFFFu
constexpr int kRowWarps = 8;

template <typename T>
__device__ __forceinline__ T shfl_xor_t(T v, int o) { return __shfl_xor_sync(FULL, v, o); }

template <typename T>
__device__ __forceinline__ uint32_t quant_one(T v, T mn, T mx);
template <>
__device__ __forceinline__ uint32_t quant_one<float>(float v, float mn, float mx) { return ref_quant_f32(v, mn, mx); }
template <>
__device__ __forceinline__ uint32_t quant_one<double>(double v, double mn, double mx) { return ref_quant_f64(v, mn, mx); }

// SOA=false: out = AoS rows of (8+d) bytes.  SOA=true: out = codes (stride d_pad) + hdr + sums.
template <typename T, bool SOA>
__global__ void __launch_bounds__(kRowWarps * 32) quantize_kernel(const T *__restrict__ in, size_t n, int d,
                                                                  uint8_t *__restrict__ out, int d_pad,
                                                                  float2 *__restrict__ hdr, uint2 *__restrict__ sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool vec4 = (d & 3) == 0;
    for (size_t row = (size_t)blockIdx.x * kRowWarps + warp; row < n; row += (size_t)gridDim.x * kRowWarps) {
        const T *v = in + row * (size_t)d;
        // rangeFloat32/64 (quantization.go:194-216): both seeded at 0, strict comparisons (NaN ignored)
        T mn = 0, mx = 0;
        for (int i = lane; i < d; i += 32) {
            T x = v[i];
            if (x < mn) mn = x;
            if (x > mx) mx = x;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            T a = shfl_xor_t(mn, o), b = shfl_xor_t(mx, o);
            if (a < mn) mn = a;
            if (b > mx) mx = b;
        }
        uint8_t *dst = SOA ? out + row * (size_t)d_pad : out + row * (size_t)(8 + d) + 8;
        uint32_t s1 = 0, s2 = 0;
        if (vec4) {
            for (int i = lane * 4; i < d; i += 128) {
                uint32_t c0 = quant_one<T>(v[i], mn, mx), c1 = quant_one<T>(v[i + 1], mn, mx);
                uint32_t c2 = quant_one<T>(v[i + 2], mn, mx), c3 = quant_one<T>(v[i + 3], mn, mx);
                uint32_t w = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
                *reinterpret_cast<uint32_t *>(dst + i) = w;
                if (SOA) {
                    s1 = dp4a_u(w, 0x01010101u, s1);
                    s2 = dp4a_u(w, w, s2);
                }
            }
        } else {
            for (int i = lane; i < d; i += 32) {
                uint32_t c = quant_one<T>(v[i], mn, mx);
                dst[i] = (uint8_t)c;
                s1 += c;
                s2 += c * c;
            }
        }
        if (SOA) {
            for (int i = d + lane; i < d_pad; i += 32) dst[i] = 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s1 += __shfl_xor_sync(FULL, s1, o);
                s2 += __shfl_xor_sync(FULL, s2, o);
            }
            if (lane == 0) {
                hdr[row] = make_float2((float)mn, (float)mx);  // float32(min), float32(max) (quantization.go:96-97)
                sums[row] = make_uint2(s1, s2);
            }
        } else if (lane == 0) {
            float *h = reinterpret_cast<float *>(out + row * (size_t)(8 + d));  // 8+d is a multiple of 4 when vec4
            if (vec4) {
                h[0] = (float)mn;
                h[1] = (float)mx;
            } else {
                uint32_t a = __float_as_uint((float)mn), b = __float_as_uint((float)mx);
                uint8_t *hb = out + row * (size_t)(8 + d);
                for (int k = 0; k < 4; k++) {
                    hb[k] = (uint8_t)(a >> (8 * k));
                    hb[4 + k] = (uint8_t)(b >> (8 * k));
                }
            }
        }
    }
}

static unsigned row_grid(size_t n) {
    size_t b = (n + kRowWarps - 1) / kRowWarps;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (unsigned)b;
}

cudaError_t launch_quantize_f32(const float *in, size_t n, int d, uint8_t *out_rows, cudaStream_t st) {
    quantize_kernel<float, false><<<row_grid(n), kRowWarps * 32, 0, st>>>(in, n, d, out_rows, 0, nullptr, nullptr);
    return cudaGetLastError();
}
cudaError_t launch_quantize_f64(const double *in, size_t n, int d, uint8_t *out_rows, cudaStream_t st) {
    quantize_kernel<double, false><<<row_grid(n), kRowWarps * 32, 0, st>>>(in, n, d, out_rows, 0, nullptr, nullptr);
    return cudaGetLastError();
}
cudaError_t launch_quantize_f32_soa(const float *in, size_t n, int d, uint8_t *codes, int d_pad, float2 *hdr,
                                    uint2 *sums, cudaStream_t st) {
    quantize_kernel<float, true><<<row_grid(n), kRowWarps * 32, 0, st>>>(in, n, d, codes, d_pad, hdr, sums);
    return cudaGetLastError();
}

// DequantizeMatrixFloat32/64 (quantization.go:114-132,166-180) from AoS rows.
template <typename T>
__global__ void __launch_bounds__(kRowWarps * 32) dequantize_kernel(const uint8_t *__restrict__ rows, size_t n,
                                                                    int row_bytes, T *__restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = row_bytes - 8;
    for (size_t row = (size_t)blockIdx.x * kRowWarps + warp; row < n; row += (size_t)gridDim.x * kRowWarps) {
        const uint8_t *r = rows + row * (size_t)row_bytes;
        uint32_t a = 0, b = 0;
        for (int k = 0; k < 4; k++) {
            a |= (uint32_t)r[k] << (8 * k);
            b |= (uint32_t)r[4 + k] << (8 * k);
        }
        const float mnf = __uint_as_float(a), mxf = __uint_as_float(b);
        T *o = out + row * (size_t)d;
        if (sizeof(T) == 4) {
            const float range = __fsub_rn(mxf, mnf);
            for (int i = lane; i < d; i += 32) o[i] = (T)ref_dequant_f32(r[8 + i], mnf, range);
        } else {
            const double mn = (double)mnf, range = __dsub_rn((double)mxf, (double)mnf);
            for (int i = lane; i < d; i += 32) o[i] = (T)ref_dequant_f64(r[8 + i], mn, range);
        }
    }
}

cudaError_t launch_dequantize_f32(const uint8_t *rows, size_t n, int row_bytes, float *out, cudaStream_t st) {
    dequantize_kernel<float><<<row_grid(n), kRowWarps * 32, 0, st>>>(rows, n, row_bytes, out);
    return cudaGetLastError();
}
cudaError_t launch_dequantize_f64(const uint8_t *rows, size_t n, int row_bytes, double *out, cudaStream_t st) {
    dequantize_kernel<double><<<row_grid(n), kRowWarps * 32, 0, st>>>(rows, n, row_bytes, out);
    return cudaGetLastError();
}

// NewMatrix (compute/compute.go:23-44) without the dequantize: AoS row776 -> SoA + integer sums.
__global__ void __launch_bounds__(kRowWarps * 32) ingest_kernel(const uint8_t *__restrict__ rows, size_t n,
                                                                int row_bytes, uint8_t *__restrict__ codes, int d_pad,
                                                                float2 *__restrict__ hdr, uint2 *__restrict__ sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = row_bytes - 8;
    const bool w8 = (row_bytes & 7) == 0 && ((uintptr_t)rows & 7) == 0;
    for (size_t row = (size_t)blockIdx.x * kRowWarps + warp; row < n; row += (size_t)gridDim.x * kRowWarps) {
        const uint8_t *r = rows + row * (size_t)row_bytes;
        uint8_t *dst = codes + row * (size_t)d_pad;
        uint32_t s1 = 0, s2 = 0;
        if (w8) {
            const uint2 *r8 = reinterpret_cast<const uint2 *>(r);
            const int W = row_bytes >> 3;
            for (int w = lane; w < W; w += 32) {
                uint2 v = r8[w];
                if (w == 0) {
                    hdr[row] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
                } else {
                    *reinterpret_cast<uint2 *>(dst + (size_t)(w - 1) * 8) = v;
                    s1 = dp4a_u(v.x, 0x01010101u, dp4a_u(v.y, 0x01010101u, s1));
                    s2 = dp4a_u(v.x, v.x, dp4a_u(v.y, v.y, s2));
                }
            }
        } else {
            if (lane == 0) {
                uint32_t a = 0, b = 0;
                for (int k = 0; k < 4; k++) {
                    a |= (uint32_t)r[k] << (8 * k);
                    b |= (uint32_t)r[4 + k] << (8 * k);
                }
                hdr[row] = make_float2(__uint_as_float(a), __uint_as_float(b));
            }
            for (int i = lane; i < d; i += 32) {
                uint32_t c = r[8 + i];
                dst[i] = (uint8_t)c;
                s1 += c;
                s2 += c * c;
            }
        }
        for (int i = d + lane; i < d_pad; i += 32) dst[i] = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(FULL, s1, o);
            s2 += __shfl_xor_sync(FULL, s2, o);
        }
        if (lane == 0) sums[row] = make_uint2(s1, s2);
    }
}

cudaError_t launch_ingest(const uint8_t *rows, size_t n, int row_bytes, uint8_t *codes, int d_pad, float2 *hdr,
                          uint2 *sums, cudaStream_t st) {
    ingest_kernel<<<row_grid(n), kRowWarps * 32, 0, st>>>(rows, n, row_bytes, codes, d_pad, hdr, sums);
    return cudaGetLastError();
}

// SoA -> AoS row776 (reading rows back, e.g. new centroids).
__global__ void __launch_bounds__(kRowWarps * 32) export_kernel(MatView m, size_t first, size_t count,
                                                                uint8_t *__restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rb = 8 + m.d;
    for (size_t i = (size_t)blockIdx.x * kRowWarps + warp; i < count; i += (size_t)gridDim.x * kRowWarps) {
        const size_t row = first + i;
        uint8_t *o = out + i * (size_t)rb;
        if (lane == 0) {
            float2 h = m.hdr[row];
            uint32_t a = __float_as_uint(h.x), b = __float_as_uint(h.y);
            for (int k = 0; k < 4; k++) {
                o[k] = (uint8_t)(a >> (8 * k));
                o[4 + k] = (uint8_t)(b >> (8 * k));
            }
        }
        const uint8_t *src = m.codes + row * (size_t)m.d_pad;
        for (int j = lane; j < m.d; j += 32) o[8 + j] = src[j];
    }
}

cudaError_t launch_export(const MatView &m, size_t first, size_t count, uint8_t *rows_out, cudaStream_t st) {
    export_kernel<<<row_grid(count), kRowWarps * 32, 0, st>>>(m, first, count, rows_out);
    return cudaGetLastError();
}

// dst row i <- src row order[i] (grouping rows by list when an index is built).
__global__ void __launch_bounds__(kRowWarps * 32) gather_rows_kernel(MatView src, const uint32_t *__restrict__ order,
                                                                     size_t n, uint8_t *__restrict__ codes,
                                                                     float2 *__restrict__ hdr, uint2 *__restrict__ sums,
                                                                     const uint64_t *__restrict__ ids_in,
                                                                     uint64_t id_base, uint64_t *__restrict__ ids_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int CH = src.d_pad >> 4;
    for (size_t i = (size_t)blockIdx.x * kRowWarps + warp; i < n; i += (size_t)gridDim.x * kRowWarps) {
        const size_t s = order[i];
        const uint4 *sp = reinterpret_cast<const uint4 *>(src.codes + s * (size_t)src.d_pad);
        uint4 *dp = reinterpret_cast<uint4 *>(codes + i * (size_t)src.d_pad);
        for (int c = lane; c < CH; c += 32) dp[c] = sp[c];
        if (lane == 0) {
            hdr[i] = src.hdr[s];
            sums[i] = src.sums[s];
            if (ids_out) ids_out[i] = ids_in ? ids_in[s] : id_base + s;
        }
    }
}

cudaError_t launch_gather_rows(const MatView &src, const uint32_t *order, size_t n, uint8_t *codes, float2 *hdr,
                               uint2 *sums, const uint64_t *ids_in, uint64_t id_base, uint64_t *ids_out,
                               cudaStream_t st) {
    gather_rows_kernel<<<row_grid(n), kRowWarps * 32, 0, st>>>(src, order, n, codes, hdr, sums, ids_in, id_base, ids_out);
    return cudaGetLastError();
}

// dst row i <- row (order[i] & 0x7fffffff) of `a`, or of `b` when bit 31 of order[i] is set: the rows of an index and a batch
// of uploaded rows interleaved into the new list order (vs_index_upload).  Same copy shape as gather_rows_kernel: a warp per
// row, 128-bit loads and stores, 776+ B read and written once per row.
__global__ void __launch_bounds__(kRowWarps * 32) merge_rows_kernel(MatView a, const uint64_t *__restrict__ a_ids, uint64_t a_id_base,
                                                                    MatView b, const uint64_t *__restrict__ b_ids, uint64_t b_id_base,
                                                                    const uint32_t *__restrict__ order, size_t n,
                                                                    uint8_t *__restrict__ codes, float2 *__restrict__ hdr,
                                                                    uint2 *__restrict__ sums, uint64_t *__restrict__ ids_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int CH = a.d_pad >> 4;
    for (size_t i = (size_t)blockIdx.x * kRowWarps + warp; i < n; i += (size_t)gridDim.x * kRowWarps) {
        const uint32_t o = order[i];
        const bool from_b = (o >> 31) != 0;
        const size_t s = o & 0x7fffffffu;
        const uint8_t *sc = from_b ? b.codes : a.codes;   // both matrices have the same d_pad
        const uint4 *sp = reinterpret_cast<const uint4 *>(sc + s * (size_t)a.d_pad);
        uint4 *dp = reinterpret_cast<uint4 *>(codes + i * (size_t)a.d_pad);
        for (int c = lane; c < CH; c += 32) dp[c] = sp[c];
        if (lane == 0) {
            hdr[i] = (from_b ? b.hdr : a.hdr)[s];
            sums[i] = (from_b ? b.sums : a.sums)[s];
            const uint64_t *ids = from_b ? b_ids : a_ids;
            ids_out[i] = ids ? ids[s] : (from_b ? b_id_base : a_id_base) + s;
        }
    }
}

cudaError_t launch_merge_rows(const MatView &a, const uint64_t *a_ids, uint64_t a_id_base, const MatView &b, const uint64_t *b_ids,
                              uint64_t b_id_base, const uint32_t *order, size_t n, uint8_t *codes, float2 *hdr, uint2 *sums,
                              uint64_t *ids_out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    merge_rows_kernel<<<row_grid(n), kRowWarps * 32, 0, st>>>(a, a_ids, a_id_base, b, b_ids, b_id_base, order, n, codes, hdr, sums,
                                                             ids_out);
    return cudaGetLastError();
}

// Streaming loader (vs_index_fill*): the rows of one chunk, already ordered by list (order[j] = chunk row of sorted position j,
// keys[j] = its list), go straight to their final place in the grouped store: list_off[l] + rows the list already holds
// (cursor[l]) + position among this chunk's rows of the list.  A list that would overflow its reserved length raises *overflow
// and its surplus rows are dropped.  Same copy shape as gather_rows_kernel.
__global__ void __launch_bounds__(kRowWarps * 32) scatter_rows_kernel(MatView src, const uint32_t *__restrict__ order,
                                                                      const uint32_t *__restrict__ keys,
                                                                      const uint32_t *__restrict__ chunk_off,
                                                                      const uint64_t *__restrict__ list_off,
                                                                      const uint64_t *__restrict__ cursor, size_t n,
                                                                      uint8_t *__restrict__ codes, float2 *__restrict__ hdr,
                                                                      uint2 *__restrict__ sums, const uint64_t *__restrict__ ids_in,
                                                                      uint64_t id_base, uint64_t *__restrict__ ids_out,
                                                                      unsigned int *__restrict__ overflow) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int CH = src.d_pad >> 4;
    for (size_t j = (size_t)blockIdx.x * kRowWarps + warp; j < n; j += (size_t)gridDim.x * kRowWarps) {
        const uint32_t l = keys[j];
        const uint64_t at = cursor[l] + (j - chunk_off[l]);
        if (at >= list_off[l + 1] - list_off[l]) {
            if (lane == 0) *overflow = 1u;
            continue;
        }
        const size_t s = order[j], i = list_off[l] + at;
        const uint4 *sp = reinterpret_cast<const uint4 *>(src.codes + s * (size_t)src.d_pad);
        uint4 *dp = reinterpret_cast<uint4 *>(codes + i * (size_t)src.d_pad);
        for (int c = lane; c < CH; c += 32) dp[c] = sp[c];
        if (lane == 0) {
            hdr[i] = src.hdr[s];
            sums[i] = src.sums[s];
            ids_out[i] = ids_in ? ids_in[s] : id_base + s;
        }
    }
}

cudaError_t launch_scatter_rows(const MatView &src, const uint32_t *order, const uint32_t *keys, const uint32_t *chunk_off,
                                const uint64_t *list_off, const uint64_t *cursor, uint8_t *codes, float2 *hdr, uint2 *sums,
                                const uint64_t *ids_in, uint64_t id_base, uint64_t *ids_out, unsigned int *overflow, cudaStream_t st) {
    if (src.n == 0) return cudaSuccess;
    scatter_rows_kernel<<<row_grid(src.n), kRowWarps * 32, 0, st>>>(src, order, keys, chunk_off, list_off, cursor, src.n, codes, hdr, sums,
                                                                  ids_in, id_base, ids_out, overflow);
    return cudaGetLastError();
}

}  // namespace vs
