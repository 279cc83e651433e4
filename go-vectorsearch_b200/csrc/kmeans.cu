// kmeans.cu -- centroid update (K6) and recenter (K7) of SURVEY.md 2b.
//
// Replaces dnc/k_means.go:80-96 (float32 sum of dequantized member rows per centroid, in row order;
// mean = sum / float32(count); an empty cluster keeps its previous mean) and dnc/dnc.go:417-449
// (float64 sum of a cluster's rows in row order, divided by the count).  Float addition is not
// associative, so to reproduce the reference's bytes each (centroid, dimension) pair is one
// sequential chain over the member rows in ascending row order; parallelism comes from the
// k x D independent chains (four per thread, a 32-bit word of codes per row) and from the unrolled
// row loop, whose loads do not depend on the additions.  HBM-bound: every member row is read once.
#include <cuda.h>
#include <stdlib.h>

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

// Two codes at a time with Blackwell's packed float32 pairs (add/mul/fma .rn.f32x2: IEEE per half, so the bits are those
// of the scalar sequence): (lo, hi) = dequantized values of bytes k and k+1 of `word`, added onto the running pair.
__device__ __forceinline__ uint64_t km_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t km_add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t km_mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t km_fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// sum2 += mn + (float32(q)/255) * range for the byte pair (k, k+1) of word; constants pre-packed by the caller
__device__ __forceinline__ uint64_t km_accumulate_pair(uint64_t sum2, uint32_t word, int k, uint64_t mn2, uint64_t range2,
                                                       uint64_t neg2p23, uint64_t rcp2, uint64_t neg255, uint64_t one2) {
    const uint64_t bits = ((uint64_t)__byte_perm(word, 0x4B000000u, 0x7650u + (uint32_t)k + 1u) << 32) |
                          (uint64_t)__byte_perm(word, 0x4B000000u, 0x7650u + (uint32_t)k);
    const uint64_t x = km_add2(bits, neg2p23);       // float32(q), exact
    const uint64_t r = km_mul2(x, rcp2);
    const uint64_t e = km_fma2(r, neg255, x);        // fma(-r, 255, x) = fma(r, -255, x)
    const uint64_t q = km_fma2(e, rcp2, r);          // float32(q) / 255.0f, correctly rounded
    // mn + q*range with the product rounded on its own, as the reference does.  In packed form ptxas contracts the
    // multiply into the addition (mul.rn.f32x2 + add.rn.f32x2 -> FFMA2, even through fma(product, 1, mn): seen in the
    // SASS), which changes the bits, so this step stays scalar: mul.rn.f32 / add.rn.f32 are never contracted.
    float q_lo, q_hi, mn, range, mn_hi, range_hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(q_lo), "=f"(q_hi) : "l"(q));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(mn), "=f"(mn_hi) : "l"(mn2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(range), "=f"(range_hi) : "l"(range2));
    const uint64_t val = km_pack(__fadd_rn(mn, __fmul_rn(q_lo, range)), __fadd_rn(mn_hi, __fmul_rn(q_hi, range_hi)));
    return km_add2(sum2, val);
}

// order[seg_off[c] .. seg_off[c+1]) = rows assigned to centroid c, ascending.
// RELAY = false: sums start at zero and the mean is written (one device holds every row).
// RELAY = true: the chains continue from sums[] and the running sums are written back, counts are added to: the rows
// of a store cut into contiguous blocks are accumulated block after block (device after device) in row order, which
// is the reference's single left-to-right chain, bit for bit; kmeans_finalize_kernel divides at the end.
template <bool RELAY>
__global__ void __launch_bounds__(kAccThreads)
kmeans_accumulate_kernel(MatView data, const uint32_t *__restrict__ order, const uint32_t *__restrict__ seg_off,
                         float *means, int64_t *__restrict__ counts, const float *means_prev) {
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
    } else if (!RELAY && live && means_prev != means) {  // the previous mean carries over into the other buffer
        float *m = means + (size_t)c * data.d;
        const float *mp = means_prev + (size_t)c * data.d;
        for (int t = 0; t < 4; t++)
            if (j0 + t < data.d) m[j0 + t] = mp[j0 + t];
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

// Few centroids (the reference's own shapes: k <= 25): no sort at all.  Block (c, slice) walks the assignment in row
// order, compacts the rows of centroid c of every 768-row window into shared memory (order kept: a thread owns four
// consecutive rows, positions come from a block prefix sum) and continues its chains over them.  Reading the whole
// assignment once per centroid (k x n x 4 B) is nothing next to what a radix sort's launches cost at these sizes.
constexpr int kScanRowsPerThread = 4;
constexpr int kScanWindow = kAccThreads * kScanRowsPerThread;
__global__ void __launch_bounds__(kAccThreads)
kmeans_accumulate_scan_kernel(MatView data, const int32_t *__restrict__ assign, float *means, int64_t *__restrict__ counts,
                              const float *means_prev) {
    __shared__ uint32_t s_row[kScanWindow];
    __shared__ float2 s_hdr[kScanWindow];
    __shared__ uint32_t s_wsum[kAccThreads / 32];
    const int c = blockIdx.x;
    const int j0 = (blockIdx.y * kAccThreads + threadIdx.x) * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool live = j0 < data.d_pad;
    const size_t n = data.n;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    uint64_t total = 0;
    for (size_t base = 0; base < n; base += kScanWindow) {
        const size_t r0 = base + (size_t)threadIdx.x * kScanRowsPerThread;
        bool hit[kScanRowsPerThread];
        uint32_t mine = 0;
#pragma unroll
        for (int t = 0; t < kScanRowsPerThread; t++) {
            hit[t] = r0 + t < n && assign[r0 + t] == c;
            mine += hit[t] ? 1u : 0u;
        }
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += x;
        }
        __syncthreads();  // the previous window's rows are consumed
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t pos = incl - mine, cnt = 0;
#pragma unroll
        for (int w = 0; w < kAccThreads / 32; w++) {
            if (w < warp) pos += s_wsum[w];
            cnt += s_wsum[w];
        }
#pragma unroll
        for (int t = 0; t < kScanRowsPerThread; t++) {
            if (hit[t]) {
                s_row[pos] = (uint32_t)(r0 + t);
                s_hdr[pos] = data.hdr[r0 + t];
                pos++;
            }
        }
        __syncthreads();
        total += cnt;
        if (live) {
#pragma unroll 8
            for (uint32_t i = 0; i < cnt; i++) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(data.codes + (size_t)s_row[i] * data.d_pad + j0);
                const float2 h = s_hdr[i];
                const float range = __fsub_rn(h.y, h.x);
                s0 = __fadd_rn(s0, __fadd_rn(h.x, __fmul_rn(byte_over_255(w, 0), range)));
                s1 = __fadd_rn(s1, __fadd_rn(h.x, __fmul_rn(byte_over_255(w, 1), range)));
                s2 = __fadd_rn(s2, __fadd_rn(h.x, __fmul_rn(byte_over_255(w, 2), range)));
                s3 = __fadd_rn(s3, __fadd_rn(h.x, __fmul_rn(byte_over_255(w, 3), range)));
            }
        }
    }
    if (blockIdx.y == 0 && threadIdx.x == 0) counts[c] = (int64_t)total;
    if (!live) return;
    float *m = means + (size_t)c * data.d;
    if (total > 0) {  // :89-96
        const float nf = (float)(int64_t)total;
        if (j0 + 0 < data.d) m[j0 + 0] = __fdiv_rn(s0, nf);
        if (j0 + 1 < data.d) m[j0 + 1] = __fdiv_rn(s1, nf);
        if (j0 + 2 < data.d) m[j0 + 2] = __fdiv_rn(s2, nf);
        if (j0 + 3 < data.d) m[j0 + 3] = __fdiv_rn(s3, nf);
    } else if (means_prev != means) {  // an empty cluster keeps its previous mean
        const float *mp = means_prev + (size_t)c * data.d;
        for (int t = 0; t < 4; t++)
            if (j0 + t < data.d) m[j0 + t] = mp[j0 + t];
    }
}

cudaError_t launch_kmeans_accumulate_scan(const MatView &data, const int32_t *assign, int k, float *means, int64_t *counts,
                                          cudaStream_t st, const float *means_prev) {
    dim3 grid(k, (data.d_pad + kAccThreads * 4 - 1) / (kAccThreads * 4));
    kmeans_accumulate_scan_kernel<<<grid, kAccThreads, 0, st>>>(data, assign, means, counts, means_prev ? means_prev : means);
    return cudaGetLastError();
}

// ---- few centroids, many rows each (the reference's shapes: k <= 25 over <= 50 000 sampled rows) -------------------
// A (centroid, dimension) sum is one dependent chain over ~n/k rows, and with a handful of centroids there are only a
// handful of blocks: direct loads leave each chain waiting ~1 us of HBM latency for every few rows (measured 1.2 ms per
// update at 50 000 x 768, 30 GB/s).  So the member rows are first made contiguous per centroid (ordered member lists by
// walking the assignment, then a row gather at copy speed) and then streamed through a deep shared-memory ring by TMA
// bulk copies: one block per centroid, a producer thread keeps ~200 KB (8 stages x 32 rows) in flight, the consumer
// threads run their chains out of shared memory.
__global__ void kmeans_histogram_kernel(const int32_t *__restrict__ assign, size_t n, int k, uint32_t *__restrict__ counts) {
    extern __shared__ uint32_t s_cnt[];
    for (int i = threadIdx.x; i < k; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < n; r += (size_t)gridDim.x * blockDim.x)
        atomicAdd(&s_cnt[assign[r]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(&counts[i], s_cnt[i]);
}

// Block c: the rows assigned to centroid c, ascending, written at order[seg_off[c] ..); seg_off from the histogram.
// A thread owns eight consecutive rows of every window (their loads fly together), positions come from a block scan.
constexpr int kListThreads = 1024;
constexpr int kListRowsPerThread = 8;
__global__ void __launch_bounds__(kListThreads)
kmeans_member_lists_kernel(const int32_t *__restrict__ assign, size_t n, int k, const uint32_t *__restrict__ counts,
                           uint32_t *__restrict__ order, uint32_t *__restrict__ seg_off) {
    __shared__ uint32_t s_wsum[kListThreads / 32];
    __shared__ uint32_t s_base;
    const int c = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        uint32_t off = 0;
        for (int i = 0; i < c; i++) off += counts[i];
        s_base = off;
        seg_off[c] = off;
        if (c == k - 1) seg_off[k] = off + counts[c];
    }
    __syncthreads();
    uint32_t base = s_base;
    for (size_t w0 = 0; w0 < n; w0 += (size_t)kListThreads * kListRowsPerThread) {
        const size_t r0 = w0 + (size_t)threadIdx.x * kListRowsPerThread;
        uint32_t hits = 0, mine = 0;
#pragma unroll
        for (int t = 0; t < kListRowsPerThread; t++) {
            const bool hit = r0 + t < n && assign[r0 + t] == c;
            hits |= hit ? (1u << t) : 0u;
            mine += hit ? 1u : 0u;
        }
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += x;
        }
        __syncthreads();
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t pos = base + incl - mine, tot = 0;
        for (int w = 0; w < kListThreads / 32; w++) {
            if (w < warp) pos += s_wsum[w];
            tot += s_wsum[w];
        }
#pragma unroll
        for (int t = 0; t < kListRowsPerThread; t++)
            if (hits & (1u << t)) order[pos++] = (uint32_t)(r0 + t);
        base += tot;
    }
}

constexpr int kRingRows = 32;       // rows per stage
constexpr int kRingMaxStages = 8;   // as many as fit in ~200 KB of shared memory
__device__ __forceinline__ uint32_t km_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void km_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// codes / hdr: the member rows of all centroids, contiguous per centroid in row order (gathered), seg_off[k+1].
// blockDim = 32 * (consumer warps + 1): thread t < d_pad/4 owns dimensions 4t..4t+3; the last warp is the producer.
__global__ void __launch_bounds__(288)
kmeans_accumulate_ring_kernel(const uint8_t *__restrict__ codes, const float2 *__restrict__ hdr, int d, int d_pad,
                              const uint32_t *__restrict__ seg_off, float *means, int64_t *__restrict__ counts,
                              const float *means_prev, int consumer_warps, int kRingStages) {
    extern __shared__ __align__(128) unsigned char ring[];
    const uint32_t stage_bytes = (uint32_t)kRingRows * d_pad;
    const uint32_t hdr_stage = (kRingRows + 2) * 8;  // 16-byte aligned source: up to 8 bytes of lead-in
    unsigned char *s_codes = ring;
    unsigned char *s_hdr = ring + (size_t)kRingStages * stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_hdr + (size_t)kRingStages * hdr_stage);  // full[S], empty[S]
    const int c = blockIdx.x;
    const uint32_t beg = seg_off[c], end = seg_off[c + 1];
    const uint32_t rows = end - beg, groups = (rows + kRingRows - 1) / kRingRows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kRingStages; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(km_smem_u32(&bars[i])), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(km_smem_u32(&bars[kRingStages + i])), "r"(consumer_warps)
                         : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        counts[c] = (int64_t)rows;
    }
    __syncthreads();
    if (warp == consumer_warps) {
        if (lane == 0) {  // producer: keeps kRingStages groups of rows in flight
            for (uint32_t g = 0; g < groups; g++) {
                const uint32_t st = g % kRingStages, ph = (g / kRingStages) & 1u;
                if (g >= (uint32_t)kRingStages) km_mbar_wait(km_smem_u32(&bars[kRingStages + st]), ph ^ 1u);
                const uint32_t r0 = beg + g * kRingRows, nr = min((uint32_t)kRingRows, end - r0);
                const uint32_t cb = nr * (uint32_t)d_pad;
                const uintptr_t hsrc = reinterpret_cast<uintptr_t>(hdr + r0);
                const uintptr_t hal = hsrc & ~uintptr_t(15);
                const uint32_t hb = (uint32_t)(((hsrc - hal) + (uintptr_t)nr * 8 + 15) & ~uintptr_t(15));
                const uint32_t full = km_smem_u32(&bars[st]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(cb + hb) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 km_smem_u32(s_codes + (size_t)st * stage_bytes)),
                             "l"(codes + (size_t)r0 * d_pad), "r"(cb), "r"(full)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 km_smem_u32(s_hdr + (size_t)st * hdr_stage)),
                             "l"(reinterpret_cast<const void *>(hal)), "r"(hb), "r"(full)
                             : "memory");
            }
        }
        return;
    }
    const int j0 = threadIdx.x * 4;
    const bool live = j0 < d_pad;
    uint64_t s01 = km_pack(0.0f, 0.0f), s23 = km_pack(0.0f, 0.0f);  // k_means.go:60-65 zero-initialised sumVectors
    const uint64_t neg2p23 = km_pack(-8388608.0f, -8388608.0f), rcp2 = km_pack(0x1.010102p-8f, 0x1.010102p-8f),
                   neg255 = km_pack(-255.0f, -255.0f), one2 = km_pack(1.0f, 1.0f);
    for (uint32_t g = 0; g < groups; g++) {
        const uint32_t st = g % kRingStages, ph = (g / kRingStages) & 1u;
        km_mbar_wait(km_smem_u32(&bars[st]), ph);
        const uint32_t r0 = beg + g * kRingRows, nr = min((uint32_t)kRingRows, end - r0);
        const unsigned char *cs = s_codes + (size_t)st * stage_bytes;
        const float2 *hs = reinterpret_cast<const float2 *>(s_hdr + (size_t)st * hdr_stage + (reinterpret_cast<uintptr_t>(hdr + r0) & 15));
        if (live) {
#pragma unroll 8
            for (uint32_t i = 0; i < nr; i++) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(cs + (size_t)i * d_pad + j0);
                const float2 h = hs[i];
                const float range = __fsub_rn(h.y, h.x);
                const uint64_t mn2 = km_pack(h.x, h.x), range2 = km_pack(range, range);
                // compute.DequantizeVectorFloat32(data[i]) then sumVectors[c][j] += val  (k_means.go:81-84)
                s01 = km_accumulate_pair(s01, w, 0, mn2, range2, neg2p23, rcp2, neg255, one2);
                s23 = km_accumulate_pair(s23, w, 2, mn2, range2, neg2p23, rcp2, neg255, one2);
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(km_smem_u32(&bars[kRingStages + st])) : "memory");
    }
    if (!live) return;
    float s0, s1, s2, s3;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(s01));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(s2), "=f"(s3) : "l"(s23));
    float *m = means + (size_t)c * d;
    if (rows > 0) {  // :89-96
        const float nf = (float)(int64_t)rows;
        if (j0 + 0 < d) m[j0 + 0] = __fdiv_rn(s0, nf);
        if (j0 + 1 < d) m[j0 + 1] = __fdiv_rn(s1, nf);
        if (j0 + 2 < d) m[j0 + 2] = __fdiv_rn(s2, nf);
        if (j0 + 3 < d) m[j0 + 3] = __fdiv_rn(s3, nf);
    } else if (means_prev != means) {  // an empty cluster keeps its previous mean
        const float *mp = means_prev + (size_t)c * d;
        for (int t = 0; t < 4; t++)
            if (j0 + t < d) m[j0 + t] = mp[j0 + t];
    }
}

// Two warp roles instead of one.  In kmeans_accumulate_ring_kernel every consumer thread both dequantizes and adds, ~22
// instructions per row and thread, and a centroid's 768 dimensions sit on one SM: issue-bound at ~33 cycles per row.
// Here a block owns a 128-dimension slice of one centroid (6 blocks per centroid at 768-d): eight warps dequantize rows
// from the byte ring into a float32 ring (any order: 4 rows per warp), and one warp runs the 128 x 1 chains over the
// float32 rows with two packed additions per row.  Same operations on the same values in the same order per chain.
constexpr int kR2Slice = 128;        // bytes (= dimensions) of a row per block
constexpr int kR2Rows = 64;          // rows per stage (an mbarrier hand-off costs a few hundred cycles: amortize it)
constexpr int kR2ByteStages = 12, kR2FloatStages = 3;  // 96 KB of row bytes in flight per block; 96 KB of float32 rows
constexpr int kR2DequantWarps = 8;
constexpr int kR2Threads = 32 * (1 + kR2DequantWarps + 1);  // chain warp, dequant warps, producer warp
struct R2Smem {
    float f32[kR2FloatStages][kR2Rows][kR2Slice];               // 96 KB
    unsigned char u8[kR2ByteStages][kR2Rows][kR2Slice];         // 96 KB
    unsigned char hdr[kR2ByteStages][(kR2Rows + 2) * 8];        // row headers (with up to 8 bytes of lead-in)
    uint64_t u_full[kR2ByteStages], u_empty[kR2ByteStages], f_full[kR2FloatStages], f_empty[kR2FloatStages];
};
__device__ __forceinline__ void km_mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(km_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void km_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(km_smem_u32(bar)) : "memory");
}
// dequantized values of bytes k, k+1 of word as a packed pair: mn + (float32(q)/255) * range
__device__ __forceinline__ uint64_t km_dequant_pair(uint32_t word, int k, float mn, float range, uint64_t neg2p23, uint64_t rcp2,
                                                    uint64_t neg255) {
    const uint64_t bits = ((uint64_t)__byte_perm(word, 0x4B000000u, 0x7650u + (uint32_t)k + 1u) << 32) |
                          (uint64_t)__byte_perm(word, 0x4B000000u, 0x7650u + (uint32_t)k);
    const uint64_t x = km_add2(bits, neg2p23);
    const uint64_t r = km_mul2(x, rcp2);
    const uint64_t e = km_fma2(r, neg255, x);
    const uint64_t q = km_fma2(e, rcp2, r);
    float q_lo, q_hi;  // the product is rounded on its own (scalar: ptxas would contract the packed form)
    asm("mov.b64 {%0, %1}, %2;" : "=f"(q_lo), "=f"(q_hi) : "l"(q));
    return km_pack(__fadd_rn(mn, __fmul_rn(q_lo, range)), __fadd_rn(mn, __fmul_rn(q_hi, range)));
}

__global__ void __launch_bounds__(kR2Threads)
kmeans_accumulate_ring2_kernel(const __grid_constant__ CUtensorMap tm_codes, const float2 *__restrict__ hdr, int d, int d_pad,
                               const uint32_t *__restrict__ seg_off, float *means, int64_t *__restrict__ counts,
                               const float *means_prev, int dbg) {
    extern __shared__ __align__(128) unsigned char r2_raw[];
    R2Smem &sm = *reinterpret_cast<R2Smem *>(r2_raw);
    const int c = blockIdx.x, slice = blockIdx.y;
    const int col0 = slice * kR2Slice;                       // first byte / dimension of this slice
    const int width = min(kR2Slice, d_pad - col0);           // bytes of a row in this slice (multiple of 16)
    const uint32_t beg = seg_off[c], end = seg_off[c + 1];
    const uint32_t rows = end - beg, groups = (rows + kR2Rows - 1) / kR2Rows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kR2ByteStages; i++) {
            km_mbar_init(&sm.u_full[i], 1);
            km_mbar_init(&sm.u_empty[i], kR2DequantWarps);
        }
        for (int i = 0; i < kR2FloatStages; i++) {
            km_mbar_init(&sm.f_full[i], kR2DequantWarps);
            km_mbar_init(&sm.f_empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (slice == 0) counts[c] = (int64_t)rows;
    }
    __syncthreads();
    if (warp == 1 + kR2DequantWarps) {
        // ===== producer: one 2-D TMA tile (32 rows x this slice) and the group's headers per stage =====
        // (one bulk copy per 128-byte row slice was tried first: 32 tiny copies per stage cost ~2000 cycles)
        if (lane == 0) {
            for (uint32_t g = 0; g < groups; g++) {
                const uint32_t st = g % kR2ByteStages, ph = (g / kR2ByteStages) & 1u;
                if (g >= (uint32_t)kR2ByteStages) km_mbar_wait(km_smem_u32(&sm.u_empty[st]), ph ^ 1u);
                const uint32_t r0 = beg + g * kR2Rows, nr = min((uint32_t)kR2Rows, end - r0);
                const uintptr_t hsrc = reinterpret_cast<uintptr_t>(hdr + r0);
                const uintptr_t hal = hsrc & ~uintptr_t(15);
                const uint32_t hb = (uint32_t)(((hsrc - hal) + (uintptr_t)nr * 8 + 15) & ~uintptr_t(15));
                const uint32_t full = km_smem_u32(&sm.u_full[st]);
                // the tile always delivers the whole box (rows past the store read as zero and are not used)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full),
                             "r"((uint32_t)(kR2Rows * kR2Slice) + hb)
                             : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                        km_smem_u32(&sm.u8[st][0][0])),
                    "l"(&tm_codes), "r"(col0), "r"((int)r0), "r"(full)
                    : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 km_smem_u32(&sm.hdr[st][0])),
                             "l"(reinterpret_cast<const void *>(hal)), "r"(hb), "r"(full)
                             : "memory");
            }
        }
        return;
    }
    if (warp >= 1) {
        // ===== dequantize: warp w takes rows 4w .. 4w+3 of every group; lane l bytes 4l .. 4l+3 of the slice =====
        const int w = warp - 1;
        const uint64_t neg2p23 = km_pack(-8388608.0f, -8388608.0f), rcp2 = km_pack(0x1.010102p-8f, 0x1.010102p-8f),
                       neg255 = km_pack(-255.0f, -255.0f);
        const bool live = lane * 4 < width;
        for (uint32_t g = 0; g < groups; g++) {
            const uint32_t su = g % kR2ByteStages, pu = (g / kR2ByteStages) & 1u;
            const uint32_t sf = g % kR2FloatStages, pf = (g / kR2FloatStages) & 1u;
            const uint32_t r0 = beg + g * kR2Rows, nr = min((uint32_t)kR2Rows, end - r0);
            km_mbar_wait(km_smem_u32(&sm.u_full[su]), pu);
            if (g >= (uint32_t)kR2FloatStages) km_mbar_wait(km_smem_u32(&sm.f_empty[sf]), pf ^ 1u);
            const float2 *hs = reinterpret_cast<const float2 *>(&sm.hdr[su][reinterpret_cast<uintptr_t>(hdr + r0) & 15]);
#pragma unroll
            for (int t = 0; t < kR2Rows / kR2DequantWarps; t++) {
                const uint32_t i = (uint32_t)(w * (kR2Rows / kR2DequantWarps) + t);
                if (i < nr && live && !(dbg & 2)) {
                    const uint32_t word = *reinterpret_cast<const uint32_t *>(&sm.u8[su][i][lane * 4]);
                    const float2 h = hs[i];
                    const float range = __fsub_rn(h.y, h.x);
                    // compute.DequantizeVectorFloat32(data[i]) (k_means.go:81)
                    const uint64_t v01 = km_dequant_pair(word, 0, h.x, range, neg2p23, rcp2, neg255);
                    const uint64_t v23 = km_dequant_pair(word, 2, h.x, range, neg2p23, rcp2, neg255);
                    *reinterpret_cast<ulonglong2 *>(&sm.f32[sf][i][lane * 4]) = make_ulonglong2(v01, v23);
                }
            }
            __syncwarp();
            if (lane == 0) {
                km_mbar_arrive(&sm.u_empty[su]);
                km_mbar_arrive(&sm.f_full[sf]);  // (release: the stores above are ordered before the arrive)
            }
        }
        return;
    }
    // ===== chain warp: thread t owns dimensions col0 + 4t .. 4t+3; sumVectors[c][j] += val in row order (k_means.go:82-84) =====
    const bool live = lane * 4 < width;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;  // scalar adds: four independent 4-cycle chains (a dependent FADD2
                                                       // chain measured ~8x slower per row)
    for (uint32_t g = 0; g < groups; g++) {
        const uint32_t sf = g % kR2FloatStages, pf = (g / kR2FloatStages) & 1u;
        const uint32_t r0 = beg + g * kR2Rows, nr = min((uint32_t)kR2Rows, end - r0);
        km_mbar_wait(km_smem_u32(&sm.f_full[sf]), pf);
        if (live && !(dbg & 1)) {
#pragma unroll 8
            for (uint32_t i = 0; i < nr; i++) {
                const float4 v = *reinterpret_cast<const float4 *>(&sm.f32[sf][i][lane * 4]);
                s0 = __fadd_rn(s0, v.x);
                s1 = __fadd_rn(s1, v.y);
                s2 = __fadd_rn(s2, v.z);
                s3 = __fadd_rn(s3, v.w);
            }
        }
        __syncwarp();
        if (lane == 0) km_mbar_arrive(&sm.f_empty[sf]);
    }
    if (!live) return;
    const int j0 = col0 + lane * 4;
    float *m = means + (size_t)c * d;
    if (rows > 0) {  // :89-96
        const float nf = (float)(int64_t)rows;
        if (j0 + 0 < d) m[j0 + 0] = __fdiv_rn(s0, nf);
        if (j0 + 1 < d) m[j0 + 1] = __fdiv_rn(s1, nf);
        if (j0 + 2 < d) m[j0 + 2] = __fdiv_rn(s2, nf);
        if (j0 + 3 < d) m[j0 + 3] = __fdiv_rn(s3, nf);
    } else if (means_prev != means) {  // an empty cluster keeps its previous mean
        const float *mp = means_prev + (size_t)c * d;
        for (int t = 0; t < 4; t++)
            if (j0 + t < d) m[j0 + t] = mp[j0 + t];
    }
}

bool kmeans_ring_supported(int d_pad) { return d_pad >= 16 && d_pad <= 1024 && d_pad % 16 == 0; }
size_t kmeans_ring_scratch_bytes(size_t n, int k, int d_pad) {
    auto pad = [](size_t b) { return (b + 255) & ~size_t(255); };
    return pad((size_t)(k + 1) * 4) * 2 + pad(n * 4) + pad(n * (size_t)d_pad) + pad(n * 8) + pad(n * 8) + 1024;
}

// assign -> means (and counts) for few centroids: histogram, ordered member lists, gather, ring accumulate.
cudaError_t launch_kmeans_accumulate_ring(const MatView &data, const int32_t *assign, int k, float *means, int64_t *counts,
                                          const float *means_prev, void *scratch, cudaStream_t st) {
    auto pad = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t n = data.n;
    char *p = static_cast<char *>(scratch);
    uint32_t *hist = reinterpret_cast<uint32_t *>(p);
    p += pad((size_t)(k + 1) * 4);
    uint32_t *seg_off = reinterpret_cast<uint32_t *>(p);
    p += pad((size_t)(k + 1) * 4);
    uint32_t *order = reinterpret_cast<uint32_t *>(p);
    p += pad(n * 4);
    uint8_t *g_codes = reinterpret_cast<uint8_t *>(p);
    p += pad(n * (size_t)data.d_pad);
    float2 *g_hdr = reinterpret_cast<float2 *>(p);
    p += pad(n * 8);
    uint2 *g_sums = reinterpret_cast<uint2 *>(p);
    cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)(k + 1) * 4, st);
    if (e != cudaSuccess) return e;
    const unsigned hb = (unsigned)((n + 1023) / 1024 < 592 ? (n + 1023) / 1024 : 592);
    kmeans_histogram_kernel<<<hb ? hb : 1, 256, (size_t)k * 4, st>>>(assign, n, k, hist);
    kmeans_member_lists_kernel<<<k, kListThreads, 0, st>>>(assign, n, k, hist, order, seg_off);
    e = launch_gather_rows(data, order, n, g_codes, g_hdr, g_sums, nullptr, 0, nullptr, st);
    if (e != cudaSuccess) return e;
    if (!getenv("VS_KMEANS_RING1")) {  // (the one-role ring stays selectable for A/B timing)
        const size_t smem2 = sizeof(R2Smem) + 128;
        e = cudaFuncSetAttribute(kmeans_accumulate_ring2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return e;
        CUtensorMap tm;
        if (!make_u8_tile_map(&tm, g_codes, (uint64_t)data.d_pad, (uint64_t)n, (uint64_t)data.d_pad, kR2Slice, kR2Rows))
            return cudaErrorNotSupported;
        dim3 grid2(k, (data.d_pad + kR2Slice - 1) / kR2Slice);
        const char *dbg = getenv("VS_KMEANS_DBG");  // profiling aid: 1 = chain warp skips its additions, 2 = no dequantize
        kmeans_accumulate_ring2_kernel<<<grid2, kR2Threads, smem2, st>>>(tm, g_hdr, data.d, data.d_pad, seg_off, means, counts,
                                                                        means_prev ? means_prev : means, dbg ? atoi(dbg) : 0);
        return cudaGetLastError();
    }
    const int consumer_warps = (data.d_pad / 4 + 31) / 32;
    const size_t per_stage = (size_t)kRingRows * data.d_pad + (kRingRows + 2) * 8;
    int stages = (int)((200 * 1024) / per_stage);
    if (stages > kRingMaxStages) stages = kRingMaxStages;
    if (stages < 2) return cudaErrorInvalidValue;
    const size_t smem = (size_t)stages * per_stage + 2 * (size_t)stages * 8 + 128;
    e = cudaFuncSetAttribute(kmeans_accumulate_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kmeans_accumulate_ring_kernel<<<k, 32 * (consumer_warps + 1), smem, st>>>(g_codes, g_hdr, data.d, data.d_pad, seg_off, means, counts,
                                                                             means_prev ? means_prev : means, consumer_warps, stages);
    return cudaGetLastError();
}

// means_prev: the means before this iteration (an empty cluster keeps its previous mean); may be `means` itself.
cudaError_t launch_kmeans_accumulate(const MatView &data, const uint32_t *order, const uint32_t *seg_off, int k,
                                     float *means, int64_t *counts, cudaStream_t st, const float *means_prev) {
    dim3 grid(k, (data.d_pad + kAccThreads * 4 - 1) / (kAccThreads * 4));
    kmeans_accumulate_kernel<false><<<grid, kAccThreads, 0, st>>>(data, order, seg_off, means, counts, means_prev ? means_prev : means);
    return cudaGetLastError();
}

cudaError_t launch_kmeans_accumulate_relay(const MatView &data, const uint32_t *order, const uint32_t *seg_off, int k,
                                           float *sums, int64_t *counts, cudaStream_t st) {
    dim3 grid(k, (data.d_pad + kAccThreads * 4 - 1) / (kAccThreads * 4));
    kmeans_accumulate_kernel<true><<<grid, kAccThreads, 0, st>>>(data, order, seg_off, sums, counts, sums);
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
