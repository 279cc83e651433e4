// kmeans.cu -- centroid update (K6) and recenter (K7) of SURVEY.md 2b.
//
// Replaces dnc/k_means.go:80-96 (float32 sum of dequantized member rows per centroid, in row order;
// mean = sum / float32(count); an empty cluster keeps its previous mean) and dnc/dnc.go:417-449
// (float64 sum of a cluster's rows in row order, divided by the count).  Float addition is not
// associative, so to reproduce the reference's bytes each (centroid, dimension) pair is one
// sequential chain over the member rows in ascending row order; parallelism comes from the
// k x D independent chains (four per thread, a 32-bit word of codes per row) and from the unrolled
// row loop, whose loads do not depend on the additions.  HBM-bound: every member row is read once.
#include "internal.h"

namespace vs {

constexpr int kAccThreads = 192;  // four adjacent dimensions per thread: one block covers 768 columns of a centroid
constexpr int kAccChunk = 64;     // member rows staged per round (row index + header in shared memory)

// float32(q) / 255.0f, correctly rounded like the reference's `float32(quantized) / 255.0` (quantization.go:56), without
// the division sequence: q as a float by a byte permute into 2^23's mantissa, one multiply by float32(1/255) and one
// Newton correction (exact for every q in 0..255: checked exhaustively against IEEE division).
__device__ __forceinline__ float byte_over_255(uint32_t word, int k) {
    const float x = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650u + (uint32_t)k)) - 8388608.0f;
    const float rcp = 0x1.010102p-8f;
    const float r = __fmul_rn(x, rcp);
    const float e = __fmaf_rn(-r, 255.0f, x);
    return __fmaf_rn(e, rcp, r);
}

// order[seg_off[c] .. seg_off[c+1]) = rows assigned to centroid c, ascending.
// RELAY = false: sums start at zero and the mean is written (one device holds every row).
// RELAY = true: the chains continue from sums[] and the running sums are written back, counts are added to: the rows
// of a store cut into contiguous blocks are accumulated block after block (device after device) in row order, which
// is the reference's single left-to-right chain, bit for bit; kmeans_finalize_kernel divides at the end.
template <bool RELAY>
__global__ void __launch_bounds__(kAccThreads)
kmeans_accumulate_kernel(MatView data, const uint32_t *__restrict__ order, const uint32_t *__restrict__ seg_off,
                         float *__restrict__ means, int64_t *__restrict__ counts) {
    __shared__ uint32_t s_row[kAccChunk];
    __shared__ float2 s_hdr[kAccChunk];
    const int c = blockIdx.x;
    const int j0 = (blockIdx.y * kAccThreads + threadIdx.x) * 4;
    const uint32_t beg = seg_off[c], end = seg_off[c + 1];
    if (blockIdx.y == 0 && threadIdx.x == 0) counts[c] = (RELAY ? counts[c] : 0) + (int64_t)(end - beg);
    const bool live = j0 < data.d_pad;  // d_pad is a multiple of 16: the 4-byte word stays inside the padded row
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;  // k_means.go:60-65 zero-initialised sumVectors
    if (RELAY && live) {
        const float *m = means + (size_t)c * data.d;
        if (j0 + 0 < data.d) s0 = m[j0 + 0];
        if (j0 + 1 < data.d) s1 = m[j0 + 1];
        if (j0 + 2 < data.d) s2 = m[j0 + 2];
        if (j0 + 3 < data.d) s3 = m[j0 + 3];
    }
    for (uint32_t base = beg; base < end; base += kAccChunk) {
        const int cnt = (int)min((uint32_t)kAccChunk, end - base);
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            const uint32_t r = order[base + threadIdx.x];
            s_row[threadIdx.x] = r;
            s_hdr[threadIdx.x] = data.hdr[r];
        }
        __syncthreads();
        if (live) {
#pragma unroll 8
            for (int i = 0; i < cnt; i++) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(data.codes + (size_t)s_row[i] * data.d_pad + j0);
                const float2 h = s_hdr[i];
                const float range = __fsub_rn(h.y, h.x);
                // compute.DequantizeVectorFloat32(data[i]) then sumVectors[c][j] += val  (k_means.go:81-84)
                s0 = __fadd_rn(s0, __fadd_rn(h.x, __fmul_rn(byte_over_255(w, 0), range)));
                s1 = __fadd_rn(s1, __fadd_rn(h.x, __fmul_rn(byte_over_255(w, 1), range)));
                s2 = __fadd_rn(s2, __fadd_rn(h.x, __fmul_rn(byte_over_255(w, 2), range)));
                s3 = __fadd_rn(s3, __fadd_rn(h.x, __fmul_rn(byte_over_255(w, 3), range)));
            }
        }
    }
    if (end > beg && live) {  // :89-96 (an empty cluster keeps its previous mean)
        const float n = RELAY ? 1.0f : (float)(int64_t)(end - beg);
        float *m = means + (size_t)c * data.d;
        if (j0 + 0 < data.d) m[j0 + 0] = RELAY ? s0 : __fdiv_rn(s0, n);
        if (j0 + 1 < data.d) m[j0 + 1] = RELAY ? s1 : __fdiv_rn(s1, n);
        if (j0 + 2 < data.d) m[j0 + 2] = RELAY ? s2 : __fdiv_rn(s2, n);
        if (j0 + 3 < data.d) m[j0 + 3] = RELAY ? s3 : __fdiv_rn(s3, n);
    }
}

// means[c][j] = sums[c][j] / float32(counts[c]) where the cluster has members (k_means.go:89-96)
__global__ void kmeans_finalize_kernel(const float *__restrict__ sums, const int64_t *__restrict__ counts, size_t k, int d,
                                       float *__restrict__ means) {
    const size_t total = k * (size_t)d;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int64_t n = counts[i / d];
        if (n > 0) means[i] = __fdiv_rn(sums[i], (float)n);
    }
}

cudaError_t launch_kmeans_accumulate(const MatView &data, const uint32_t *order, const uint32_t *seg_off, int k,
                                     float *means, int64_t *counts, cudaStream_t st) {
    dim3 grid(k, (data.d_pad + kAccThreads * 4 - 1) / (kAccThreads * 4));
    kmeans_accumulate_kernel<false><<<grid, kAccThreads, 0, st>>>(data, order, seg_off, means, counts);
    return cudaGetLastError();
}

cudaError_t launch_kmeans_accumulate_relay(const MatView &data, const uint32_t *order, const uint32_t *seg_off, int k,
                                           float *sums, int64_t *counts, cudaStream_t st) {
    dim3 grid(k, (data.d_pad + kAccThreads * 4 - 1) / (kAccThreads * 4));
    kmeans_accumulate_kernel<true><<<grid, kAccThreads, 0, st>>>(data, order, seg_off, sums, counts);
    return cudaGetLastError();
}

cudaError_t launch_kmeans_finalize(const float *sums, const int64_t *counts, size_t k, int d, float *means, cudaStream_t st) {
    kmeans_finalize_kernel<<<1024, 256, 0, st>>>(sums, counts, k, d, means);
    return cudaGetLastError();
}

// dnc/dnc.go:417-449: dataSum[idx] += DequantizeVectorFloat64(row)[idx] over all rows in order, then
// dataSum[idx] / float64(count).  mean_out: float64[d] (quantized afterwards by QuantizeVectorFloat64).
__global__ void __launch_bounds__(32) recenter_kernel(MatView data, double *__restrict__ mean_out) {
    const int j = blockIdx.x * 32 + threadIdx.x;
    if (j >= data.d) return;
    double sum = 0.0;
#pragma unroll 4
    for (size_t row = 0; row < data.n; row++) {
        const float2 h = data.hdr[row];
        const uint32_t q = data.codes[row * (size_t)data.d_pad + j];
        const double mn = (double)h.x;
        sum = __dadd_rn(sum, ref_dequant_f64(q, mn, __dsub_rn((double)h.y, mn)));
    }
    mean_out[j] = __ddiv_rn(sum, (double)(unsigned long long)data.n);
}

// recenterDbCentroid for every cluster at once: order[seg_off[c] .. seg_off[c+1]) = member rows of cluster c, ascending
// (the reference reads a cluster's embeddings in primary-key order); one float64 chain per (cluster, dimension).
// float64(q) / 255.0 comes from a 256-entry shared-memory table built with IEEE division (quantization.go:64).
// A cluster without members divides 0 by 0: NaN, which QuantizeVectorFloat64 turns into all-zero codes (dnc.go:437-441).
constexpr int kRecThreads = 128;
__global__ void __launch_bounds__(kRecThreads)
recenter_clusters_kernel(MatView data, const uint32_t *__restrict__ order, const uint32_t *__restrict__ seg_off,
                         double *__restrict__ means, int64_t *__restrict__ counts) {
    __shared__ double s_div[256];
    __shared__ uint32_t s_row[kAccChunk];
    __shared__ float2 s_hdr[kAccChunk];
    for (int q = threadIdx.x; q < 256; q += kRecThreads) s_div[q] = __ddiv_rn((double)q, 255.0);
    const int c = blockIdx.x;
    const int j = blockIdx.y * kRecThreads + threadIdx.x;
    const uint32_t beg = seg_off[c], end = seg_off[c + 1];
    if (blockIdx.y == 0 && threadIdx.x == 0) counts[c] = (int64_t)(end - beg);
    const bool live = j < data.d;
    double sum = 0.0;
    for (uint32_t base = beg; base < end; base += kAccChunk) {
        const int cnt = (int)min((uint32_t)kAccChunk, end - base);
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
            const uint32_t r = order[base + threadIdx.x];
            s_row[threadIdx.x] = r;
            s_hdr[threadIdx.x] = data.hdr[r];
        }
        __syncthreads();
        if (live) {
#pragma unroll 8
            for (int i = 0; i < cnt; i++) {
                const uint32_t q = data.codes[(size_t)s_row[i] * data.d_pad + j];
                const float2 h = s_hdr[i];
                const double mn = (double)h.x;
                // DequantizeVectorFloat64 (quantization.go:63-69,124-132), then dataSum[idx] += val (dnc.go:431-433)
                sum = __dadd_rn(sum, __dadd_rn(mn, __dmul_rn(s_div[q], __dsub_rn((double)h.y, mn))));
            }
        }
    }
    __syncthreads();
    if (live) means[(size_t)c * data.d + j] = __ddiv_rn(sum, (double)(unsigned long long)(end - beg));  // dnc.go:437-439
}

cudaError_t launch_recenter_clusters(const MatView &data, const uint32_t *order, const uint32_t *seg_off, int k, double *means,
                                     int64_t *counts, cudaStream_t st) {
    dim3 grid(k, (data.d + kRecThreads - 1) / kRecThreads);
    recenter_clusters_kernel<<<grid, kRecThreads, 0, st>>>(data, order, seg_off, means, counts);
    return cudaGetLastError();
}

cudaError_t launch_recenter(const MatView &data, double *mean_out, cudaStream_t st) {
    recenter_kernel<<<(data.d + 31) / 32, 32, 0, st>>>(data, mean_out);
    return cudaGetLastError();
}

}  // namespace vs
