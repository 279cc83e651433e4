// argmax.cu -- nearest-centroid assignment (K4 of SURVEY.md 2b), scan form for small/medium m.
//
// Replaces (*matrixContainer).MatrixCosineSimilarity, compute/cosine.go:70-125, as called from
// dnc/k_means.go:75,165, dnc/dnc.go:206,374,384,536 and server/upload.go:245: for every data row the
// index of the centroid with the largest float64 cosine, strict '>' from maxVal=-1.0/maxIdx=0, so the
// lowest index wins ties.  Only nearestIndexList is consumed upstream.
//
// Shape in the reference: m <= 25 centroids x n <= 10 000 rows per call -> HBM-bound.  Centroid codes
// sit in shared memory (tiles of kCentTile), each lane group streams one data row into registers with
// 128-bit loads and takes u8 dot products (dp4a) against every centroid of the tile.  The comparison
// runs on certified float64 intervals (common.cuh score_interval); a row whose winner cannot be
// separated from the runner-up by the error bound goes to the literal-arithmetic fix kernel.
#include "internal.h"

namespace vs {

#define FULL 0xFFFFFFFFu
constexpr int kArgWarps = 8;
constexpr int kCentTile = 32;  // centroids per shared-memory tile

struct Best {
    double c;      // best certified value so far
    double d;      // its half-width
    double other;  // max upper bound (c+delta) over all other candidates
    int idx;
};

__device__ __forceinline__ void best_merge(Best &a, const Best &b) {
    // a <- merge(a, b); candidates with idx < 0 are "none"
    bool b_wins = b.idx >= 0 && (a.idx < 0 || b.c > a.c || (b.c == a.c && b.idx < a.idx));
    double loser_hi;
    if (b_wins) {
        loser_hi = a.idx >= 0 ? a.c + a.d : -2.0;
        double o = fmax(a.other, b.other);
        a.c = b.c;
        a.d = b.d;
        a.idx = b.idx;
        a.other = fmax(o, loser_hi);
    } else {
        loser_hi = b.idx >= 0 ? b.c + b.d : -2.0;
        a.other = fmax(fmax(a.other, b.other), loser_hi);
    }
}

// G lanes per data row, CPL 16-byte chunks per lane; MASK: chunk index may exceed the row (generic d).
template <int G, int CPL, bool MASK>
__global__ void __launch_bounds__(kArgWarps * 32, 2)
argmax_kernel(MatView cent, MatView data, const uint32_t *__restrict__ canon, int32_t *__restrict__ idx_out,
              float *__restrict__ sims_out, uint32_t *__restrict__ worklist, unsigned int *__restrict__ work_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = data.d, d_pad = data.d_pad, CH = d_pad >> 4;
    uint4 *sh_codes = reinterpret_cast<uint4 *>(smem_raw);                                   // [kCentTile][CH]
    SideConst *sh_side = reinterpret_cast<SideConst *>(smem_raw + (size_t)kCentTile * d_pad);  // [kCentTile]
    int *sh_skip = reinterpret_cast<int *>(sh_side + kCentTile);                             // [kCentTile]

    constexpr int NG = 32 / G;       // rows per warp iteration
    constexpr int SL = kCentTile / G > 0 ? kCentTile / G : 1;  // dots kept per lane per tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / G, l = lane % G;
    const unsigned gmask = (G == 32) ? FULL : (((1u << G) - 1u) << (g * G));
    const double sqrtD = sqrt((double)D);
    const int M = (int)cent.n;

    const size_t rows_per_block_iter = (size_t)kArgWarps * NG;
    const bool single = M <= kCentTile;
    bool first_iter = true;
    for (size_t base = (size_t)blockIdx.x * rows_per_block_iter; base < data.n;
         base += (size_t)gridDim.x * rows_per_block_iter) {
        const size_t row = base + (size_t)warp * NG + g;
        const bool valid = row < data.n;
        uint4 v[CPL];
#pragma unroll
        for (int j = 0; j < CPL; j++) {
            int c = l + G * j;
            bool ok = valid && (!MASK || c < CH);
            v[j] = ok ? ld_stream_u4(data.codes + row * (size_t)d_pad + (size_t)c * 16) : make_uint4(0, 0, 0, 0);
        }
        SideConst y;
        {
            float2 h = valid ? data.hdr[row] : make_float2(0.f, 0.f);
            uint2 s = valid ? data.sums[row] : make_uint2(0, 0);
            y = make_side(h.x, h.y, s.x, s.y, D);
        }
        Best best;
        best.c = -2.0;
        best.d = 0.0;
        best.other = -2.0;
        best.idx = -1;
        bool bad = false;

        for (int m0 = 0; m0 < M; m0 += kCentTile) {
            const int mt = min(kCentTile, M - m0);
            if (!single || first_iter) {  // m <= kCentTile (the reference's shapes): the tile is loaded once
                __syncthreads();          // previous tile fully consumed
                for (int i = threadIdx.x; i < mt * CH; i += blockDim.x)
                    sh_codes[i] = reinterpret_cast<const uint4 *>(cent.codes + (size_t)m0 * d_pad)[i];
                for (int j = threadIdx.x; j < mt; j += blockDim.x) {
                    float2 h = cent.hdr[m0 + j];
                    uint2 s = cent.sums[m0 + j];
                    sh_side[j] = make_side(h.x, h.y, s.x, s.y, D);
                    sh_skip[j] = canon ? (canon[m0 + j] != (uint32_t)(m0 + j)) : 0;
                }
                __syncthreads();
                first_iter = false;
            }
            uint32_t dots[SL];
#pragma unroll
            for (int s = 0; s < SL; s++) dots[s] = 0;
            for (int j = 0; j < mt; j++) {
                const uint4 *cc = sh_codes + (size_t)j * CH;
                uint32_t acc = 0;
#pragma unroll
                for (int jj = 0; jj < CPL; jj++) {
                    int c = l + G * jj;
                    if (!MASK || c < CH) acc = dot16(v[jj], cc[c], acc);
                }
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
                if (G >= kCentTile) {
                    if (l == j) dots[0] = acc;
                } else {
#pragma unroll
                    for (int s = 0; s < SL; s++)
                        if (j == s * G + l) dots[s] = acc;
                }
            }
            // every lane finishes the (row, centroid) pairs it kept
#pragma unroll
            for (int s = 0; s < SL; s++) {
                int j = (G >= kCentTile) ? l : s * G + l;
                if (j < mt && !sh_skip[j]) {
                    double c, dl;
                    bool ok = score_interval(sh_side[j], y, dots[s], D, sqrtD, &c, &dl);
                    if (!ok) bad = true;
                    Best b;
                    b.c = ok ? c : 2.0;
                    b.d = ok ? dl : 0.0;
                    b.other = -2.0;
                    b.idx = m0 + j;
                    best_merge(best, b);
                }
            }
        }
        // reduce over the G lanes of the group
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            Best b;
            b.c = __shfl_xor_sync(FULL, best.c, o);
            b.d = __shfl_xor_sync(FULL, best.d, o);
            b.other = __shfl_xor_sync(FULL, best.other, o);
            b.idx = __shfl_xor_sync(FULL, best.idx, o);
            best_merge(best, b);
        }
        bad = (__ballot_sync(FULL, bad) & gmask) != 0;
        if (valid && l == 0) {
            // certain iff the winner clears every rival's upper bound and the -1.0 seed (cosine.go:101-102)
            bool certain = !bad && best.idx >= 0 && (best.c - best.d > best.other) && (best.c - best.d > -1.0);
            float simf = 0.f;
            if (sims_out && certain) {
                float lo = __double2float_rn(best.c - best.d), hi = __double2float_rn(best.c + best.d);
                if (lo != hi) certain = false;
                simf = hi;
            }
            idx_out[row] = certain ? best.idx : -1;
            if (sims_out) sims_out[row] = simf;
            if (!certain) worklist[atomicAdd(work_count, 1u)] = (uint32_t)row;
        }
    }
}

cudaError_t launch_argmax(const MatView &cent, const MatView &data, const uint32_t *canon, int32_t *idx_out,
                          float *sims_out, uint32_t *worklist, unsigned int *work_count, int sm_count,
                          cudaStream_t st) {
    const int ch = data.d_pad >> 4;
    const size_t smem = (size_t)kCentTile * data.d_pad + kCentTile * (sizeof(SideConst) + sizeof(int));
#define VS_ARGMAX_LAUNCH(G, CPL, MASK)                                                                         \
    do {                                                                                                       \
        auto kern = argmax_kernel<G, CPL, MASK>;                                                               \
        if (smem > 48 * 1024) {                                                                                \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                    \
        }                                                                                                      \
        size_t per = (size_t)kArgWarps * (32 / G);                                                             \
        size_t blocks = (data.n + per - 1) / per;                                                              \
        size_t maxb = (size_t)sm_count * 8;                                                                    \
        if (blocks > maxb) blocks = maxb;                                                                      \
        if (blocks < 1) blocks = 1;                                                                            \
        kern<<<(unsigned)blocks, kArgWarps * 32, smem, st>>>(cent, data, canon, idx_out, sims_out, worklist,   \
                                                             work_count);                                     \
    } while (0)
    if (ch == 48) VS_ARGMAX_LAUNCH(16, 3, false);
    else if (ch == 32) VS_ARGMAX_LAUNCH(32, 1, false);
    else if (ch == 64) VS_ARGMAX_LAUNCH(32, 2, false);
    else if (ch == 96) VS_ARGMAX_LAUNCH(32, 3, false);
    else if (ch == 24) VS_ARGMAX_LAUNCH(8, 3, false);
    else if (ch <= 256) VS_ARGMAX_LAUNCH(32, 8, true);
    else return cudaErrorInvalidValue;
#undef VS_ARGMAX_LAUNCH
    return cudaGetLastError();
}

// Literal reference arithmetic for the rows the interval test could not decide: one warp per row,
// the normalized row goes to shared memory, lanes split the centroids, each dot is sequential.
// cnorm = centroids after normalizeVector (float64 [m][d], from query_normalize_kernel).
__global__ void argmax_fix_kernel(MatView cent, MatView data, const double *__restrict__ cnorm,
                                  int32_t *__restrict__ idx_out, float *__restrict__ sims_out,
                                  const uint32_t *__restrict__ worklist, const unsigned int *__restrict__ work_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int D = data.d;
    double *B = reinterpret_cast<double *>(smem_raw) + (size_t)warp * D;
    const unsigned int nwork = *work_count;
    for (unsigned int wi = blockIdx.x * nwarp + warp; wi < nwork; wi += gridDim.x * nwarp) {
        const uint32_t row = worklist[wi];
        const uint8_t *codes = data.codes + (size_t)row * data.d_pad;
        const float2 h = data.hdr[row];
        const double mn = (double)h.x, range = __dsub_rn((double)h.y, (double)h.x);
        double norm = 0.0;
        if (lane == 0) {  // normalizeVector (cosine.go:138-149): sequential sum of squares
            for (int i = 0; i < D; i++) {
                double x = ref_dequant_f64(codes[i], mn, range);
                norm = __dadd_rn(norm, __dmul_rn(x, x));
            }
            norm = __dsqrt_rn(norm);
        }
        norm = __shfl_sync(FULL, norm, 0);
        for (int i = lane; i < D; i += 32) {
            double x = ref_dequant_f64(codes[i], mn, range);
            if (norm != 0.0) x = __ddiv_rn(x, norm);
            B[i] = x;
        }
        __syncwarp();
        double bestv = -1.0;  // cosine.go:101-102
        int besti = -1;       // -1 = still the seed (index 0)
        for (int j = lane; j < (int)cent.n; j += 32) {
            const double *A = cnorm + (size_t)j * D;
            double dot = 0.0;
            for (int k = 0; k < D; k++) dot = __dadd_rn(dot, __dmul_rn(A[k], B[k]));
            if (dot > bestv) {  // strict: within a lane j ascends, so the first maximum is kept
                bestv = dot;
                besti = j;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(FULL, bestv, o);
            int oi = __shfl_xor_sync(FULL, besti, o);
            bool take = (oi >= 0) && (besti < 0 || ov > bestv || (ov == bestv && oi < besti));
            if (take) {
                bestv = ov;
                besti = oi;
            }
        }
        if (lane == 0) {
            idx_out[row] = besti < 0 ? 0 : besti;
            if (sims_out) sims_out[row] = __double2float_rn(bestv);
        }
        __syncwarp();
    }
}

cudaError_t launch_argmax_fix(const MatView &cent, const MatView &data, const double *cnorm, int32_t *idx_out,
                              float *sims_out, const uint32_t *worklist, const unsigned int *work_count, int sm_count,
                              cudaStream_t st) {
    int warps = 4;
    while (warps > 1 && (size_t)warps * data.d * sizeof(double) > 96 * 1024) warps >>= 1;
    size_t smem = (size_t)warps * data.d * sizeof(double);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(argmax_fix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    argmax_fix_kernel<<<sm_count * 2, warps * 32, smem, st>>>(cent, data, cnorm, idx_out, sims_out, worklist, work_count);
    return cudaGetLastError();
}

// canon[j] = lowest index whose row (header + codes) is byte-identical to row j.  A duplicate can never
// win the strict '>' scan of cosine.go:114, so the assign kernel skips it instead of calling it a tie.
__global__ void canonical_rows_kernel(MatView m, uint32_t *canon) {
    const int j = blockIdx.x;
    __shared__ int s_first;
    if (threadIdx.x == 0) s_first = j;
    __syncthreads();
    const float2 hj = m.hdr[j];
    const uint2 sj = m.sums[j];
    const uint4 *cj = reinterpret_cast<const uint4 *>(m.codes + (size_t)j * m.d_pad);
    const int CH = m.d_pad >> 4;
    for (int i = threadIdx.x; i < j; i += blockDim.x) {
        const float2 hi = m.hdr[i];
        const uint2 si = m.sums[i];
        if (__float_as_uint(hi.x) != __float_as_uint(hj.x) || __float_as_uint(hi.y) != __float_as_uint(hj.y) ||
            si.x != sj.x || si.y != sj.y)
            continue;
        const uint4 *ci = reinterpret_cast<const uint4 *>(m.codes + (size_t)i * m.d_pad);
        bool same = true;
        for (int c = 0; c < CH && same; c++) {
            uint4 a = ci[c], b = cj[c];
            same = a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
        }
        if (same) atomicMin(&s_first, i);
    }
    __syncthreads();
    if (threadIdx.x == 0) canon[j] = (uint32_t)s_first;
}

cudaError_t launch_canonical_rows(const MatView &m, uint32_t *canon, cudaStream_t st) {
    canonical_rows_kernel<<<(unsigned)m.n, 128, 0, st>>>(m, canon);
    return cudaGetLastError();
}

cudaError_t argmax_set_certify_scale(float scale) { return cudaMemcpyToSymbol(c_certify_scale, &scale, sizeof(float)); }

}  // namespace vs
