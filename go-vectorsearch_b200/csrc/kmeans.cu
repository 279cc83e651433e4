// kmeans.cu -- centroid update (K6) and recenter (K7) of SURVEY.md 2b.
//
// Replaces dnc/k_means.go:80-96 (float32 sum of dequantized member rows per centroid, in row order;
// mean = sum / float32(count); an empty cluster keeps its previous mean) and dnc/dnc.go:417-449
// (float64 sum of a cluster's rows in row order, divided by the count).  Float addition is not
// associative, so to reproduce the reference's bytes each (centroid, dimension) pair is one
// sequential chain over the member rows in ascending row order; parallelism comes from the
// k x D independent chains.  HBM-bound: every member row is read once (776 B per row).
#include "internal.h"

namespace vs {

constexpr int kDimsPerBlock = 64;

// order[seg_off[c] .. seg_off[c+1]) = rows assigned to centroid c, ascending.
__global__ void __launch_bounds__(kDimsPerBlock)
kmeans_accumulate_kernel(MatView data, const uint32_t *__restrict__ order, const uint32_t *__restrict__ seg_off,
                         float *__restrict__ means, int64_t *__restrict__ counts) {
    const int c = blockIdx.x;
    const int j = blockIdx.y * kDimsPerBlock + threadIdx.x;
    const uint32_t beg = seg_off[c], end = seg_off[c + 1];
    if (blockIdx.y == 0 && threadIdx.x == 0) counts[c] = (int64_t)(end - beg);
    if (j >= data.d) return;
    float sum = 0.0f;  // k_means.go:60-65 zero-initialised sumVectors
#pragma unroll 4
    for (uint32_t i = beg; i < end; i++) {
        const uint32_t row = order[i];
        const float2 h = data.hdr[row];
        const uint32_t q = data.codes[(size_t)row * data.d_pad + j];
        // compute.DequantizeVectorFloat32(data[i]) then sumVectors[c][j] += val  (k_means.go:81-84)
        sum = __fadd_rn(sum, ref_dequant_f32(q, h.x, __fsub_rn(h.y, h.x)));
    }
    if (end > beg) means[(size_t)c * data.d + j] = __fdiv_rn(sum, (float)(int64_t)(end - beg));  // :89-96
}

cudaError_t launch_kmeans_accumulate(const MatView &data, const uint32_t *order, const uint32_t *seg_off, int k,
                                     float *means, int64_t *counts, cudaStream_t st) {
    dim3 grid(k, (data.d + kDimsPerBlock - 1) / kDimsPerBlock);
    kmeans_accumulate_kernel<<<grid, kDimsPerBlock, 0, st>>>(data, order, seg_off, means, counts);
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

cudaError_t launch_recenter(const MatView &data, double *mean_out, cudaStream_t st) {
    recenter_kernel<<<(data.d + 31) / 32, 32, 0, st>>>(data, mean_out);
    return cudaGetLastError();
}

}  // namespace vs
