// gemm.cu -- query batches as an int8 tensor-core GEMM (tcgen05.mma kind::i8, u8 x u8 -> s32 in TMEM) with a fused
// per-pair filter, for sm_100a.  Replaces, for nq queries at once, the loop of 1xN scans the reference runs per query
// (server/search.go:241-273 -> compute/cosine.go:13-57) when the list is the whole store (BASELINE config 3).
//
// The score matrix [nq x n] is never written.  The only term of the cosine identity (common.cuh) that couples a
// query and a row is the integer dot product sum(q*v); the tensor cores produce it exactly, and the epilogue turns
// "cos(q, v) >= tau_q" into "dot >= T(q, v)" with three FMAs per pair (T is a sum of three per-query x per-row
// products, see GemmRowConst / the column constants).  Pairs that pass are appended to a candidate list and finished
// by select_kernel with the certified float64 score of common.cuh (and literal reference arithmetic where the
// float32 rounding cannot be certified), so the emitted float32 similarities and ids are the reference's bits.
//
// tau_q comes from a pre-pass of the same GEMM over a strided sample of the store (MODE_GROUPMAX): the r-th largest of
// the per-16-row-group maxima is a lower bound of the r-th best score.  The result of a query is accepted only if k
// distinct documents at or above tau_q were found (then nothing outside the candidate list can belong to the top k);
// otherwise the query is flagged and the caller finishes it with the streaming scan.  Correctness therefore never
// depends on how good tau_q is -- only the amount of work does.
//
// Kernel structure (one CTA per SM, 384 threads):
//   warp 0      TMA producer: the 128-row store tile (K chunks of 128 bytes, SWIZZLE_128B) stays resident in shared
//               memory while every 128-query tile is streamed through a ring of 16 KB stages; per-tile column
//               constants arrive by cp.async.bulk.
//   warp 1      one thread issues tcgen05.mma.cta_group::1.kind::i8 (M=128 store rows, N=128 queries, K=32 per
//               instruction); tcgen05.commit releases smem stages and publishes accumulators.
//   warp 2      TMEM allocation (4 accumulator stages x 128 columns = 512 columns).
//   warps 4-11  epilogue: tcgen05.ld 32x32b, one store row per thread, 64 query columns per warp.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "internal.h"

namespace vs {
namespace {

constexpr int kTM = 128;                 // store rows per tile (TMEM lanes)
#ifndef VS_GEMM_TN
#define VS_GEMM_TN 256
#endif
constexpr int kTN = VS_GEMM_TN;          // queries per tile (accumulator columns)
constexpr int kChunkK = 128;             // K bytes per shared-memory chunk (one 128-byte swizzle atom)
constexpr int kChunkBytes = kTM * kChunkK;  // 16 KB: one K chunk of the resident store tile
constexpr int kStageBytes = kTN * kChunkK;  // one K chunk of a streamed query tile
constexpr int kAccStages = 512 / kTN;    // TMEM accumulator stages (512 columns in all)
constexpr int kEpiWarps = 16;
constexpr int kColsPerWarp = kTN / (kEpiWarps / 4);  // the epilogue warps of a lane quadrant split the columns
constexpr int kGemmThreads = (4 + kEpiWarps) * 32;
constexpr int kMaxKC = 8;                // d_pad <= 1024
constexpr int kMaxStages = 8;

constexpr int MODE_GROUPMAX = 0;
constexpr int MODE_FILTER = 1;

struct GemmParams {
    const float2 *row_hdr;
    const uint2 *row_sums;
    uint32_t n;                // store rows
    int D;
    int kc;                    // K chunks per row
    int nstages;               // streamed-operand stages
    uint32_t tile_first, tile_stride, tile_count;  // store tiles of this launch: tile_first + i * tile_stride
    uint32_t nq_tiles;         // query tiles
    uint32_t qsplit, qt_per;   // a work item = (store tile, range of qt_per query tiles); qsplit ranges per store tile
    const float4 *col_consts;  // [nq_tiles * 128]
    // MODE_FILTER: one bucket of cand_per_q (row, dot) pairs per query, filled through a per-query counter
    unsigned int *cand_cnt;    // [nq_pad]
    uint2 *cand_rowdot;        // [nq_pad][cand_per_q]
    uint32_t cand_per_q;
    // MODE_GROUPMAX
    float *gmax;               // [nq_pad][G]
    uint32_t G;
    int dbg;                   // profiling aid (VS_GEMM_DBG): 1 = epilogue skips the filter math, 2 = also skips tcgen05.ld
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    long long t0 = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done) {  // a broken pipeline must fault, not hang the device: ~5 s at 2 GHz
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000ll) __trap();
        }
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {  // one lane of a converged warp
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, u8 x u8 -> s32
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// 16 lanes x (8 x 256 bits): thread 4*r + c receives, for n = 0..7, columns 8n + 2c, 8n + 2c + 1 of lane r (registers
// 4n, 4n + 1) and of lane r + 8 (registers 4n + 2, 4n + 3).
__device__ __forceinline__ void tc_ld_16x256b_x8(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {  // same, n = 0..3
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8]) {  // same, n = 0..1
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x2.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// The wait names the destination registers of the load it completes, so no use of them can be scheduled above it.
__device__ __forceinline__ void tc_ld_wait8(uint32_t (&v)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
                 :
                 : "memory");
}

// Shared-memory matrix descriptor of a K-major [rows][128 B] SWIZZLE_128B tile (8-row groups 1024 B apart);
// kbytes = offset of this instruction's 32-byte K slice inside the 128-byte swizzle atom.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t tile_smem_addr, uint32_t kbytes) {
    const uint64_t start = ((tile_smem_addr + kbytes) & 0x3FFFFu) >> 4;
    return start | (1ull << 16) /* LBO (unused with swizzle) */ | (64ull << 32) /* SBO = 1024 B */ | (1ull << 46) /* version */ |
           (2ull << 61) /* SWIZZLE_128B */;
}
// Instruction descriptor, kind::i8: D = s32, A = B = unsigned 8-bit, both K-major, M = 128, N = 128.
constexpr uint32_t kIdesc = (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kTN >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);

// The integer dot product as a float without the (slow, XU-pipe) int->float conversion: 2^23 + (dot >> 3), exact
// because dot < 2^26.  The filter compares it with (T - margin) / 8 + 2^23.
__device__ __forceinline__ float dot_as_f8(uint32_t dot) { return __uint_as_float(0x4B000000u + (dot >> 3)); }
// dot rounded down to a multiple of 8, as a float (pre-pass: an approximate score is enough)
__device__ __forceinline__ float dot_approx(uint32_t dot) { return fmaf(dot_as_f8(dot), 8.0f, -67108864.0f); }

// Packed float32 pairs (Blackwell fma.rn.f32x2): one instruction evaluates the filter of two adjacent query columns.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// Column constants are stored per pair of adjacent queries as (x0, x1, y0, y1, z0, z1, w0, w1).
__host__ __device__ __forceinline__ size_t col_const_index(uint32_t q, int component) {
    return (size_t)(q >> 1) * 8 + (size_t)component * 2 + (q & 1);
}

__device__ __forceinline__ int f32_ordered(float f) {
    const int b = __float_as_int(f);
    return b >= 0 ? b : (b ^ 0x7FFFFFFF);
}
__device__ __forceinline__ float ordered_f32(int k) { return __int_as_float(k >= 0 ? k : (k ^ 0x7FFFFFFF)); }

// Per-store-row side of the filter.  With (b, b') the row header, S = b' - b, uy = 255*D*b + S*sum(v),
// Py = S^2*(D*sum(v^2) - sum(v)^2) + uy^2  (= 255^2 * D * |y|^2):
//   bp = sqrt(Py)/S,  cp = uy/S,  Bp = 255*D*b/S.
// A row whose header makes the identity unusable (S <= 0, non-finite, Py <= 0) gets Bp = +inf: it passes every
// filter (the certified score decides) and is left out of the group maxima.
struct GemmRowConst {
    float bp, cp, Bp, inv_bp;
    int usable;
};
__device__ __forceinline__ GemmRowConst gemm_row_const(float mn, float mx, uint32_t s1, uint32_t s2, int D, bool in_range) {
    GemmRowConst r;
    const double b = (double)mn, S = (double)mx - b;
    const double uy = (double)D * 255.0 * b + S * (double)s1;
    const long long I = (long long)D * (long long)s2 - (long long)s1 * (long long)s1;
    const double P = S * S * (double)I + uy * uy;
    const bool ok = in_range && (S > 0.0) && (P > 0.0) && (P < 1.0e300) && (S < 1.0e300);
    if (ok) {
        const double sp = sqrt(P);
        r.bp = (float)(sp / S);
        r.cp = (float)(uy / S);
        r.Bp = (float)((double)D * 255.0 * b / S);
        r.inv_bp = (float)(S / sp);
        r.usable = 1;
    } else {
        r.bp = 0.0f;
        r.cp = 0.0f;
        r.Bp = in_range ? __int_as_float(0x7F800000) : __int_as_float(0xFF800000);  // +inf: always passes; -inf: never
        r.inv_bp = 0.0f;
        r.usable = 0;
    }
    return r;
}

// a >= b, evaluated where it is written (volatile: neither hoisted nor merged with an equal comparison elsewhere)
__device__ __forceinline__ bool fge_opaque(float a, float b) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ge.f32 p, %1, %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(r)
        : "f"(a), "f"(b));
    return r != 0;
}

template <int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tm_rows, const __grid_constant__ CUtensorMap tm_queries, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t *sm = smem_raw + (base - raw);
    const int kc = p.kc, ns = p.nstages;
    const uint32_t sA = base;
    const uint32_t sB = sA + (uint32_t)kc * kChunkBytes;
    const uint32_t sQ = sB + (uint32_t)ns * kStageBytes;
    const uint32_t sBar = sQ + kAccStages * kTN * 16;
    float4 *q_consts = reinterpret_cast<float4 *>(sm + (sQ - base));
    // barrier slots
    const uint32_t a_full = sBar, a_empty = a_full + 8 * kMaxKC, b_full = a_empty + 8 * kMaxKC, b_empty = b_full + 8 * kMaxStages,
                   acc_full = b_empty + 8 * kMaxStages, acc_empty = acc_full + 8 * kAccStages, q_full = acc_empty + 8 * kAccStages;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(sm + (q_full + 8 * kAccStages - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_rows) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_queries) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kc; i++) {
            mbar_init(a_full + 8 * i, 1);
            mbar_init(a_empty + 8 * i, 1);
        }
        for (int i = 0; i < ns; i++) {
            mbar_init(b_full + 8 * i, 1);
            mbar_init(b_empty + 8 * i, 1);
        }
        for (int i = 0; i < kAccStages; i++) {
            mbar_init(acc_full + 8 * i, 1);
            mbar_init(acc_empty + 8 * i, kEpiWarps);
            mbar_init(q_full + 8 * i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void *)tmem_slot)),
                     "r"(kAccStages * kTN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t nqt = p.nq_tiles;
    const uint32_t n_items = p.tile_count * p.qsplit;

    // Every role loop is executed by its whole warp (waits are warp-convergent); the single-thread instructions
    // (TMA, tcgen05.mma, tcgen05.commit) are issued by one elected lane.  Ring positions are advanced
    // incrementally: no division on the issue path.
    if (warp == 0) {
        // ===== TMA producer =====
        const bool leader = elect_one();
        uint32_t acc_s = 0, acc_ph = 0, st = 0, bph = 0, li = 0;
        for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x, li++) {
            const uint32_t i = w / p.qsplit, qt0 = (w % p.qsplit) * p.qt_per, qt1 = min(nqt, qt0 + p.qt_per);
            const int row0 = (int)((p.tile_first + i * p.tile_stride) * kTM);
            for (int c = 0; c < kc; c++) {
                mbar_wait(a_empty + 8 * c, (li & 1) ^ 1);
                if (leader) {
                    mbar_expect_tx(a_full + 8 * c, kChunkBytes);
                    tma_load_2d(sA + c * kChunkBytes, &tm_rows, c * kChunkK, row0, a_full + 8 * c);
                }
            }
            for (uint32_t qt = qt0; qt < qt1; qt++) {
                mbar_wait(acc_empty + 8 * acc_s, acc_ph ^ 1);  // the epilogue has finished with this stage's constants
                if (leader) {
                    mbar_expect_tx(q_full + 8 * acc_s, kTN * 16);
                    bulk_load_1d(sQ + acc_s * kTN * 16, p.col_consts + (size_t)qt * kTN, kTN * 16, q_full + 8 * acc_s);
                }
                if (++acc_s == kAccStages) {
                    acc_s = 0;
                    acc_ph ^= 1;
                }
                for (int c = 0; c < kc; c++) {
                    mbar_wait(b_empty + 8 * st, bph ^ 1);
                    if (leader) {
                        mbar_expect_tx(b_full + 8 * st, kStageBytes);
                        tma_load_2d(sB + st * kStageBytes, &tm_queries, c * kChunkK, (int)(qt * kTN), b_full + 8 * st);
                    }
                    if (++st == (uint32_t)ns) {
                        st = 0;
                        bph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const bool leader = elect_one();
        const uint64_t desc_hi = (64ull << 32) | (1ull << 46) | (2ull << 61);  // SBO = 1024 B, version 1, SWIZZLE_128B
        const uint64_t a_desc0 = desc_hi | (1ull << 16) | (uint64_t)((sA & 0x3FFFFu) >> 4);
        const uint64_t b_desc0 = desc_hi | (1ull << 16) | (uint64_t)((sB & 0x3FFFFu) >> 4);
        uint32_t acc_s = 0, acc_ph = 0, st = 0, bph = 0, li = 0;
        for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x, li++) {
            const uint32_t qt0 = (w % p.qsplit) * p.qt_per, qt1 = min(nqt, qt0 + p.qt_per);
            for (uint32_t qt = qt0; qt < qt1; qt++) {
                mbar_wait(acc_empty + 8 * acc_s, acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc_s * kTN;
                for (int c = 0; c < kc; c++) {
                    if (qt == qt0) mbar_wait(a_full + 8 * c, li & 1);
                    mbar_wait(b_full + 8 * st, bph);
                    tc_fence_after();
                    if (leader) {
                        const uint64_t ad = a_desc0 + (uint64_t)(c * (kChunkBytes >> 4));
                        const uint64_t bd = b_desc0 + (uint64_t)(st * (kStageBytes >> 4));
#pragma unroll
                        for (int k = 0; k < kChunkK / 32; k++)  // +32 bytes of K inside the swizzle atom = +2 encoded
                            if (p.dbg != 5 || (c | k) == 0)  // profiling aid 5: one MMA per tile (epilogue-only timing)
                            tc_mma_i8(d_tmem, ad + 2 * k, bd + 2 * k, kIdesc, (c | k) != 0 ? 1u : 0u);
                        tc_commit(b_empty + 8 * st);
                        if (qt == qt1 - 1) tc_commit(a_empty + 8 * c);
                        if (c == kc - 1) tc_commit(acc_full + 8 * acc_s);
                    }
                    __syncwarp();
                    if (++st == (uint32_t)ns) {
                        st = 0;
                        bph ^= 1;
                    }
                }
                if (++acc_s == kAccStages) {
                    acc_s = 0;
                    acc_ph ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        // tcgen05.ld 16x256b hands every thread a few query columns of several store rows (the layout of an mma C
        // fragment: lane = 4*r + c holds columns 8n + 2c, 8n + 2c + 1 of rows r and r + 8), so a thread keeps the
        // constants of its 8 columns in registers for a whole 32-column pass: 8 shared-memory loads per pass instead
        // of one per column -- the shared-memory port belongs to the tensor cores and TMA.  Sixteen warps (four per
        // scheduler) at <= 96 registers hide the TMEM-load and FMA-chain latencies.
        const int e = warp - 4;
        const int qd = e & 3;       // TMEM lane quadrant this warp may read
        const int part = e >> 2;    // which slice of the query columns
        const int tc = lane & 3, tr = lane >> 2;
        uint32_t acc_s = 0, acc_ph = 0;
        for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x) {
            const uint32_t i = w / p.qsplit, qt0 = (w % p.qsplit) * p.qt_per, qt1 = min(nqt, qt0 + p.qt_per);
            const uint32_t tile = p.tile_first + i * p.tile_stride;
            const uint32_t row_own = tile * kTM + qd * 32 + lane;
            const bool in_range = row_own < p.n;
            float2 h = make_float2(0.f, 0.f);
            uint2 sm2 = make_uint2(0u, 0u);
            if (in_range) {
                h = p.row_hdr[row_own];
                sm2 = p.row_sums[row_own];
            }
            const GemmRowConst own = gemm_row_const(h.x, h.y, sm2.x, sm2.y, p.D, in_range);
            // this thread's four rows: tr + 8k of the quadrant
            float r_bp[4], r_cp[4], r_Bp[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                r_bp[k] = __shfl_sync(0xFFFFFFFFu, own.bp, tr + 8 * k);
                r_cp[k] = __shfl_sync(0xFFFFFFFFu, own.cp, tr + 8 * k);
                r_Bp[k] = __shfl_sync(0xFFFFFFFFu, own.Bp, tr + 8 * k);
            }
            if constexpr (MODE == MODE_GROUPMAX) {  // the pre-pass scores with 1/bp and must skip unusable rows
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float ib = __shfl_sync(0xFFFFFFFFu, own.inv_bp, tr + 8 * k);
                    const int ok = __shfl_sync(0xFFFFFFFFu, own.usable, tr + 8 * k);
                    r_bp[k] = ok ? ib : __int_as_float(0x7FC00000);  // NaN marks the row
                }
            }
            const uint32_t row_base = tile * kTM + qd * 32 + tr;
            for (uint32_t qt = qt0; qt < qt1; qt++) {
                mbar_wait(q_full + 8 * acc_s, acc_ph);
                mbar_wait(acc_full + 8 * acc_s, acc_ph);
                tc_fence_after();
                // Units of 16 TMEM lanes x 16 query columns (tcgen05.ld 16x256b.x2: 8 registers).  The load of the next
                // unit is in flight while the current one is filtered: the TMEM read latency, which bounded the 32-column
                // version, hides behind the arithmetic.  Registers 4n + {0,1}: row tr + 16 lh, columns 8n + 2 tc + {0,1};
                // 4n + {2,3}: the row 8 further down.
                constexpr int kUnits = (kColsPerWarp / 16) * 2;
                const uint32_t taddr0 = tmem_base + ((uint32_t)(qd * 32) << 16) + acc_s * kTN + part * kColsPerWarp;
                uint64_t cx[2], cy[2], cz[2], cw[2];
                auto process = [&](const uint32_t (&v)[8], const int lh, const int colbase) {
                    if (p.dbg == 1) {
                        if (v[0] == 0xFFFFFFFFu && v[7] == 0xFFFFFFFEu) p.cand_cnt[0] = 0;
                        return;
                    }
                    const float bpA = r_bp[2 * lh], cpA = r_cp[2 * lh], BpA = r_Bp[2 * lh];
                    const float bpB = r_bp[2 * lh + 1], cpB = r_cp[2 * lh + 1], BpB = r_Bp[2 * lh + 1];
                    if constexpr (MODE == MODE_FILTER) {
                        // a side-effect-free test first (four independent predicate chains); the rare emission code
                        // runs only in the threads that have a hit
                        bool any0 = false, any1 = false, any2 = false, any3 = false;
                        const uint64_t bpA2 = pack2(bpA, bpA), cpA2 = pack2(cpA, cpA), BpA2 = pack2(BpA, BpA);
                        const uint64_t bpB2 = pack2(bpB, bpB), cpB2 = pack2(cpB, cpB), BpB2 = pack2(BpB, BpB);
#pragma unroll
                        for (int n = 0; n < 2; n++) {  // columns (tau', -A', -e', 2^23*8 - m) / 8, two at a time
                            float TA0, TA1, TB0, TB1;
                            unpack2(fma2(cx[n], bpA2, fma2(cy[n], cpA2, fma2(cz[n], BpA2, cw[n]))), TA0, TA1);
                            unpack2(fma2(cx[n], bpB2, fma2(cy[n], cpB2, fma2(cz[n], BpB2, cw[n]))), TB0, TB1);
                            any0 = any0 || (dot_as_f8(v[4 * n + 0]) >= TA0);
                            any1 = any1 || (dot_as_f8(v[4 * n + 1]) >= TA1);
                            any2 = any2 || (dot_as_f8(v[4 * n + 2]) >= TB0);
                            any3 = any3 || (dot_as_f8(v[4 * n + 3]) >= TB1);
                        }
                        if (any0 || any1 || any2 || any3) {
#pragma unroll
                            for (int n = 0; n < 2; n++) {
                                float T[4];
                                unpack2(fma2(cx[n], bpA2, fma2(cy[n], cpA2, fma2(cz[n], BpA2, cw[n]))), T[0], T[1]);
                                unpack2(fma2(cx[n], bpB2, fma2(cy[n], cpB2, fma2(cz[n], BpB2, cw[n]))), T[2], T[3]);
#pragma unroll
                                for (int c = 0; c < 4; c++) {
                                    // (an opaque compare: written as plain C++ the compiler merges these tests with the
                                    // ones above and hoists all eight into the common path -- 6 extra FSETP per unit)
                                    if (fge_opaque(dot_as_f8(v[4 * n + c]), T[c])) {
                                        const uint32_t q = qt * kTN + colbase + 8 * n + 2 * tc + (c & 1);
                                        const unsigned int pos = atomicAdd(p.cand_cnt + q, 1u);
                                        if (pos < p.cand_per_q)
                                            p.cand_rowdot[(size_t)q * p.cand_per_q + pos] =
                                                make_uint2(row_base + 16 * lh + ((c & 2) ? 8 : 0), v[4 * n + c]);
                                    }
                                }
                            }
                        }
                    } else {
                        // maximum per query column over the 16 rows of this unit (8 groups per store tile)
                        const float ninf = __int_as_float(0xFF800000);
#pragma unroll
                        for (int n = 0; n < 2; n++) {
#pragma unroll
                            for (int c = 0; c < 2; c++) {
                                float xs[2], ys[2], zs[2];  // (aq, A', e'); r_bp holds 1/bp (NaN: skip the row)
                                unpack2(cx[n], xs[0], xs[1]);
                                unpack2(cy[n], ys[0], ys[1]);
                                unpack2(cz[n], zs[0], zs[1]);
                                float sa = fmaf(ys[c], cpA, fmaf(zs[c], BpA, dot_approx(v[4 * n + c]))) * (xs[c] * bpA);
                                float sb = fmaf(ys[c], cpB, fmaf(zs[c], BpB, dot_approx(v[4 * n + 2 + c]))) * (xs[c] * bpB);
                                if (!(sa == sa)) sa = ninf;
                                if (!(sb == sb)) sb = ninf;
                                float mx = fmaxf(sa, sb);
                                mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 4));
                                mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 8));
                                mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, 16));
                                if (tr == 0) {
                                    const uint32_t q = qt * kTN + colbase + 8 * n + 2 * tc + c;
                                    p.gmax[(size_t)q * p.G + (size_t)i * 8 + qd * 2 + lh] = mx;
                                }
                            }
                        }
                    }
                };
                if (p.dbg < 2) {
                    uint32_t va[8], vb[8];
                    tc_ld_16x256b_x2(taddr0, va);
#pragma unroll 1
                    for (int g = 0; g < kUnits / 2; g++) {  // a group = 16 columns: unit 2g (lanes 0-15), 2g+1 (16-31)
                        const int colbase = part * kColsPerWarp + g * 16;
                        // this thread's column pairs 4n + tc of the group: (x0,x1), (y0,y1), (z0,z1), (w0,w1) each
                        const ulonglong2 *qc = reinterpret_cast<const ulonglong2 *>(q_consts + acc_s * kTN + colbase + 2 * tc);
#pragma unroll
                        for (int n = 0; n < 2; n++) {
                            const ulonglong2 lo = qc[8 * n], hi = qc[8 * n + 1];
                            cx[n] = lo.x;
                            cy[n] = lo.y;
                            cz[n] = hi.x;
                            cw[n] = hi.y;
                        }
                        tc_ld_wait8(va);
                        tc_ld_16x256b_x2(taddr0 + (16u << 16) + (uint32_t)(g * 16), vb);
                        process(va, 0, colbase);
                        tc_ld_wait8(vb);
                        if (g + 1 < kUnits / 2) {
                            tc_ld_16x256b_x2(taddr0 + (uint32_t)((g + 1) * 16), va);
                        } else {
                            // the last TMEM read of this accumulator stage has landed in registers: hand the stage back
                            // to the MMA warp before filtering the last unit
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(acc_empty + 8 * acc_s);
                        }
                        process(vb, 1, colbase);
                    }
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + 8 * acc_s);
                }
                if (++acc_s == kAccStages) {
                    acc_s = 0;
                    acc_ph ^= 1;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kAccStages * kTN) : "memory");
    }
}

// ---- per-matrix bounds: max over usable rows of |bp|, |cp|, |Bp| (for the filter margin) ----------------------
__global__ void row_bounds_kernel(const float2 *hdr, const uint2 *sums, uint32_t n, int D, unsigned int *bounds) {
    float mb = 0.f, mc = 0.f, mB = 0.f;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const float2 h = hdr[r];
        const uint2 s = sums[r];
        const GemmRowConst rc = gemm_row_const(h.x, h.y, s.x, s.y, D, true);
        if (rc.usable) {
            mb = fmaxf(mb, fabsf(rc.bp));
            mc = fmaxf(mc, fabsf(rc.cp));
            mB = fmaxf(mB, fabsf(rc.Bp));
        }
    }
    for (int o = 16; o; o >>= 1) {
        mb = fmaxf(mb, __shfl_xor_sync(0xFFFFFFFFu, mb, o));
        mc = fmaxf(mc, __shfl_xor_sync(0xFFFFFFFFu, mc, o));
        mB = fmaxf(mB, __shfl_xor_sync(0xFFFFFFFFu, mB, o));
    }
    if ((threadIdx.x & 31) == 0) {  // non-negative floats order like their bit patterns
        atomicMax(bounds + 0, __float_as_uint(mb));
        atomicMax(bounds + 1, __float_as_uint(mc));
        atomicMax(bounds + 2, __float_as_uint(mB));
    }
}

// ---- per-query side ------------------------------------------------------------------------------------------
// With (a, a') the query header, R = a' - a, ux = 255*D*a + R*sum(q), Px = R^2*(D*sum(q^2) - sum(q)^2) + ux^2:
//   aq = R*D/sqrt(Px)   (cos = aq * (1/bp) * (dot - T0)),   Ap = 255*a/R,   ep = sum(q)/D
//   cos >= tau  <=>  dot >= T = tau/aq * bp - Ap * cp - ep * Bp
struct QuerySide {
    double aq, Ap, ep;
    int usable;
};
__device__ __forceinline__ QuerySide query_side(float mn, float mx, uint32_t s1, uint32_t s2, int D) {
    QuerySide q;
    const double a = (double)mn, R = (double)mx - a;
    const double ux = (double)D * 255.0 * a + R * (double)s1;
    const long long I = (long long)D * (long long)s2 - (long long)s1 * (long long)s1;
    const double P = R * R * (double)I + ux * ux;
    q.usable = (R > 0.0) && (P > 0.0) && (P < 1.0e300) && (R < 1.0e300) && (s1 > 0u);
    if (q.usable) {
        q.aq = R * (double)D / sqrt(P);
        q.Ap = 255.0 * a / R;
        q.ep = (double)s1 / (double)D;
    } else {
        q.aq = q.Ap = q.ep = 0.0;
    }
    return q;
}

__device__ __forceinline__ void store_col_const(float4 *base, uint32_t q, float4 v) {
    float *f = reinterpret_cast<float *>(base);
    f[col_const_index(q, 0)] = v.x;
    f[col_const_index(q, 1)] = v.y;
    f[col_const_index(q, 2)] = v.z;
    f[col_const_index(q, 3)] = v.w;
}

// Column constants of the pre-pass: (aq, A', e', 0); padding / unusable queries score -inf everywhere (aq = NaN).
__global__ void query_consts_groupmax_kernel(MatView queries, uint32_t nq_pad, float4 *out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq_pad) return;
    float4 o = make_float4(__int_as_float(0x7FC00000), 0.f, 0.f, 0.f);
    if (q < queries.n) {
        const float2 h = queries.hdr[q];
        const uint2 s = queries.sums[q];
        const QuerySide qs = query_side(h.x, h.y, s.x, s.y, queries.d);
        if (qs.usable) o = make_float4((float)qs.aq, (float)qs.Ap, (float)qs.ep, 0.f);
    }
    store_col_const(out, q, o);
}

// tau = (r-th largest group maximum) - slack; writes the filter's column constants (tau', -A', -e', 2^23*8 - m) / 8,
// tau itself, and flags unusable queries for the caller's fallback.  kth_key = ordered-int image of the r-th largest.
__device__ __forceinline__ void threshold_finish(const MatView &queries, uint32_t q, int kth_key, uint32_t G, uint32_t r,
                                                 const unsigned int *bounds, float4 *col_consts, float *tau_out, uint32_t *status) {
    const float ninf = __int_as_float(0xFF800000), pinf = __int_as_float(0x7F800000);
    float kth = ordered_f32(kth_key);
    if (G < r || !(kth == kth)) kth = ninf;
    const float2 h = queries.hdr[q];
    const uint2 s = queries.sums[q];
    const QuerySide qs = query_side(h.x, h.y, s.x, s.y, queries.d);
    if (!qs.usable) {
        store_col_const(col_consts, q, make_float4(0.f, 0.f, 0.f, pinf));
        tau_out[q] = pinf;
        status[q] = kStatusNeedMore;
        return;
    }
    status[q] = 0;
    // the group maxima are float32 evaluations: back off by more than their error
    const float tau = (kth == ninf) ? ninf : kth - 2.0e-5f * fmaxf(1.0f, fabsf(kth));
    tau_out[q] = tau;
    if (tau == ninf) {  // not enough groups: everything passes (small stores); -e' keeps out-of-range rows out
        store_col_const(col_consts, q, make_float4(0.f, 0.f, (float)(-qs.ep * 0.125), ninf));
        return;
    }
    const double tp = (double)tau / qs.aq;
    const double Bmax = (double)__uint_as_float(bounds[0]), Cmax = (double)__uint_as_float(bounds[1]),
                 BBmax = (double)__uint_as_float(bounds[2]);
    // float32 evaluation error of T (coefficient roundings, three FMAs, int->float of the dot) plus the float32
    // rounding of the reference's own score: < 2^-20 of the term magnitudes + a few units
    // plus the epilogue's own representation: the dot is compared as 2^23 + (dot >> 3) (up to 7 units low, so T
    // drops by 8) against T/8 + 2^23 accumulated at float32 ulp 1 there (three roundings: 12 units, take 24)
    const double m = ldexp(fabs(tp) * Bmax + fabs(qs.Ap) * Cmax + fabs(qs.ep) * BBmax, -20) + 16.0 + 8.0 + 24.0;
    store_col_const(col_consts, q, make_float4((float)(tp * 0.125), (float)(-qs.Ap * 0.125), (float)(-qs.ep * 0.125), (float)(8388608.0 - m * 0.125)));
}

// One block per query (many groups: large stores).  The r-th largest key by bisection over the ordered-int domain.
constexpr int kThrThreads = 128;
__global__ void __launch_bounds__(kThrThreads)
threshold_kernel(MatView queries, uint32_t nq_pad, const float *gmax, uint32_t G, uint32_t r, const unsigned int *bounds,
                 float4 *col_consts, float *tau_out, uint32_t *status) {
    extern __shared__ int s_keys[];
    __shared__ unsigned int s_cnt;
    const uint32_t q = blockIdx.x;
    if (q >= queries.n) {  // padding columns never pass
        if (threadIdx.x == 0) store_col_const(col_consts, q, make_float4(0.f, 0.f, 0.f, __int_as_float(0x7F800000)));
        return;
    }
    for (uint32_t g = threadIdx.x; g < G; g += kThrThreads) s_keys[g] = f32_ordered(gmax[(size_t)q * G + g]);
    __syncthreads();
    long long lo = (long long)(int)0x80000000, hi = 0x7FFFFFFFll;  // largest x with count(key >= x) >= r
    while (lo < hi) {
        const long long mid = lo + (hi - lo + 1) / 2;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        unsigned int c = 0;
        for (uint32_t g = threadIdx.x; g < G; g += kThrThreads) c += (long long)s_keys[g] >= mid;
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
        __syncthreads();
        if (s_cnt >= r) lo = mid;
        else hi = mid - 1;
        __syncthreads();
    }
    if (threadIdx.x == 0) threshold_finish(queries, q, (int)lo, G, r, bounds, col_consts, tau_out, status);
}

// One warp per query (few groups: many queries against a small store, the assignment shape): the same bisection with
// the keys in registers (KPL per lane) and warp reductions, no block barriers.
template <int KPL>
__global__ void __launch_bounds__(128)
threshold_warp_kernel(MatView queries, uint32_t nq_pad, const float *gmax, uint32_t G, uint32_t r, const unsigned int *bounds,
                      float4 *col_consts, float *tau_out, uint32_t *status) {
    const int lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < nq_pad; q += nwarps) {
        if (q >= queries.n) {
            if (lane == 0) store_col_const(col_consts, q, make_float4(0.f, 0.f, 0.f, __int_as_float(0x7F800000)));
            continue;
        }
        int keys[KPL];
#pragma unroll
        for (int j = 0; j < KPL; j++) {
            const uint32_t g = (uint32_t)j * 32 + lane;
            keys[j] = g < G ? f32_ordered(gmax[(size_t)q * G + g]) : (int)0x80000000;
        }
        // padding keys are the domain minimum: they only matter when fewer than r real keys exist (handled by G < r)
        long long lo = (long long)(int)0x80000000, hi = 0x7FFFFFFFll;
        while (lo < hi) {
            const int mid = (int)(lo + (hi - lo + 1) / 2);  // lo < mid <= hi: inside the int domain
            unsigned int c = 0;
#pragma unroll
            for (int j = 0; j < KPL; j++) c += keys[j] >= mid;
            c = __reduce_add_sync(0xFFFFFFFFu, c);
            if (c >= r) lo = mid;
            else hi = (long long)mid - 1;
        }
        if (lane == 0) threshold_finish(queries, q, (int)lo, G, r, bounds, col_consts, tau_out, status);
    }
}

// ---- candidate resolution -------------------------------------------------------------------------------------
constexpr int kSelThreads = 256;
constexpr int kSelCap = 4096;      // candidates per query the block kernel holds in shared memory (= the largest bucket)
constexpr int kSelWarpCap = 128;   // candidates per query the warp kernel holds in registers (4 per lane)
constexpr uint32_t kStatusDeferred = 0x100u;  // internal: the warp kernel hands a query with uncertified candidates on

struct SelEntry {
    uint32_t key;   // f32_to_key of the float32 similarity; 0 = removed
    uint32_t row;   // bit31 = float32 rounding not certified yet
    uint64_t id;
};

// One warp per query with at most kSelWarpCap candidates (the assignment shape: tens of candidates for each of many
// queries): certified scores in registers, then k rounds of "best remaining (similarity desc, id asc), drop every other
// entry of that document" (search.go:256-270) by warp reductions.  A query with a candidate whose float32 rounding is
// not certified is left to select_kernel (literal arithmetic needs the normalized query in shared memory).
// Lower bound of the key of an entry whose float32 rounding is not certified: the true similarity is the float32 just
// below the stored (upper) one -- two keys down at most, the key of -0 being unused -- or unknown (stored 2.0).
__device__ __forceinline__ uint32_t key_lower_bound(uint32_t key, bool flagged) {
    if (!flagged) return key;
    return (key == f32_to_key(2.0f) || key < 4u) ? 1u : key - 2u;
}

__global__ void __launch_bounds__(kSelThreads)
select_warp_kernel(MatView rows, const uint64_t *ids, uint64_t id_base, int unique_ids, MatView queries, const unsigned int *cand_cnt,
                   const uint2 *cand_rowdot, uint32_t cand_per_q, const float *tau, int k, uint64_t *out_ids, float *out_sims,
                   int32_t *out_counts, uint32_t *status) {
    const int lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (kSelThreads / 32);
    const int D = rows.d;
    for (uint32_t q = blockIdx.x * (kSelThreads / 32) + (threadIdx.x >> 5); q < (uint32_t)queries.n; q += nwarps) {
        const uint32_t st = status[q];
        if (st & kStatusNeedMore) continue;
        const uint32_t cnt = cand_cnt[q];
        if (cnt > (uint32_t)kSelWarpCap || cnt > cand_per_q) continue;  // the block kernel's
        const float2 qh = queries.hdr[q];
        const uint2 qs = queries.sums[q];
        const SideConst xq = make_side(qh.x, qh.y, qs.x, qs.y, D);
        uint32_t key[kSelWarpCap / 32];
        uint64_t id[kSelWarpCap / 32];
        uint32_t flags = 0;
        bool any_flag = false;
#pragma unroll
        for (int s = 0; s < kSelWarpCap / 32; s++) {
            const uint32_t e = (uint32_t)s * 32 + lane;
            key[s] = 0;
            id[s] = kEmptyId;
            if (e < cnt) {
                const uint2 rd = cand_rowdot[(size_t)q * cand_per_q + e];
                const float2 h = rows.hdr[rd.x];
                const uint2 sm = rows.sums[rd.x];
                bool flag;
                const float sim = score_fast(xq, h.x, h.y, sm.x, sm.y, rd.y, D, &flag);
                key[s] = f32_to_key(sim);
                id[s] = ids ? ids[rd.x] : id_base + rd.x;
                any_flag |= flag;
                flags |= flag ? (1u << s) : 0u;
            }
        }
        any_flag = __any_sync(0xFFFFFFFFu, any_flag);
        if (any_flag && unique_ids && cnt >= (uint32_t)k) {
            // With one entry per document the k-th largest lower bound L is a floor of the k-th best similarity: an
            // uncertified entry whose upper value is below L cannot be among the top k and needs no literal re-score.
            uint32_t lo = 0, hi = 0xFFFFFFFFu;  // largest x with count(lower bound >= x) >= k
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo) / 2 + 1;
                unsigned int c = 0;
#pragma unroll
                for (int s = 0; s < kSelWarpCap / 32; s++) c += key[s] != 0 && key_lower_bound(key[s], (flags >> s) & 1u) >= mid;
                c = __reduce_add_sync(0xFFFFFFFFu, c);
                if (c >= (unsigned int)k) lo = mid;
                else hi = mid - 1;
            }
            bool need = false;
#pragma unroll
            for (int s = 0; s < kSelWarpCap / 32; s++) need |= ((flags >> s) & 1u) && key[s] >= lo;
            any_flag = __any_sync(0xFFFFFFFFu, need);
        }
        if (any_flag) {
            if (lane == 0) status[q] = st | kStatusDeferred;
            continue;
        }
        const uint32_t tau_key = f32_to_key(tau[q]);
        int emitted = 0, emitted_above = 0;
        for (int round = 0; round < k; round++) {
            uint32_t bk = 0;
            uint64_t bid = kEmptyId;
#pragma unroll
            for (int s = 0; s < kSelWarpCap / 32; s++)
                if (key[s] != 0 && (bk == 0 || cand_better(key[s], id[s], bk, bid))) {
                    bk = key[s];
                    bid = id[s];
                }
            for (int o = 16; o; o >>= 1) {
                const uint32_t ok = __shfl_xor_sync(0xFFFFFFFFu, bk, o);
                const uint64_t oid = __shfl_xor_sync(0xFFFFFFFFu, bid, o);
                if (ok != 0 && (bk == 0 || cand_better(ok, oid, bk, bid))) {
                    bk = ok;
                    bid = oid;
                }
            }
            if (bk == 0) break;
            if (lane == 0) {
                out_ids[(size_t)q * k + round] = bid;
                out_sims[(size_t)q * k + round] = key_to_f32(bk);
            }
            emitted++;
            if (bk >= tau_key && bk != 1u) emitted_above++;
#pragma unroll
            for (int s = 0; s < kSelWarpCap / 32; s++)
                if (id[s] == bid) key[s] = 0;  // one hit per document
        }
        if (lane == 0) {
            out_counts[q] = emitted;
            if (emitted_above < k) status[q] = st | kStatusNeedMore;
        }
    }
}

// Blocks walk the queries the warp kernel left (more than kSelWarpCap candidates, or uncertified candidates; with
// all = 1 every query): certified scores of the candidates, literal re-score where needed, then the same k rounds.
__global__ void __launch_bounds__(kSelThreads)
select_kernel(MatView rows, const uint64_t *ids, uint64_t id_base, int unique_ids, MatView queries, const unsigned int *cand_cnt,
              const uint2 *cand_rowdot, uint32_t cand_per_q, int all, const float *tau, int k, uint64_t *out_ids, float *out_sims,
              int32_t *out_counts, uint32_t *status, unsigned long long *fix_counter) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    SelEntry *ent = reinterpret_cast<SelEntry *>(sel_smem);
    double *sh_qn = reinterpret_cast<double *>(sel_smem + sizeof(SelEntry) * kSelCap);
    __shared__ SideConst s_side;
    __shared__ double s_norm;
    __shared__ int s_any_flag;
    __shared__ uint32_t s_wkey[kSelThreads / 32];
    __shared__ uint64_t s_wid[kSelThreads / 32];
    __shared__ int s_widx[kSelThreads / 32];
    __shared__ int s_best;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = rows.d, d_pad = rows.d_pad;
    for (uint32_t q = blockIdx.x; q < (uint32_t)queries.n; q += gridDim.x) {
        const uint32_t st = status[q];
        if (st & kStatusNeedMore) continue;  // unusable query: the caller's fallback answers it
        const uint32_t cnt = cand_cnt[q];
        if (!all && cnt <= (uint32_t)kSelWarpCap && cnt <= cand_per_q && !(st & kStatusDeferred)) continue;  // the warp kernel answered it
        __syncthreads();  // everyone has read status[q]; the previous query's shared state is dead
        if (cnt > (uint32_t)kSelCap || cnt > cand_per_q) {
            if (threadIdx.x == 0) status[q] = kStatusNeedMore;
            continue;
        }
        if (threadIdx.x == 0) {
            const float2 h = queries.hdr[q];
            const uint2 s = queries.sums[q];
            s_side = make_side(h.x, h.y, s.x, s.y, D);
            s_any_flag = 0;
        }
        __syncthreads();
        const SideConst xq = s_side;
        for (uint32_t e = threadIdx.x; e < cnt; e += kSelThreads) {
            const uint2 rd = cand_rowdot[(size_t)q * cand_per_q + e];
            const float2 h = rows.hdr[rd.x];
            const uint2 s = rows.sums[rd.x];
            bool flag;
            const float sim = score_fast(xq, h.x, h.y, s.x, s.y, rd.y, D, &flag);
            ent[e].key = f32_to_key(sim);
            ent[e].row = rd.x | (flag ? kFlagBit : 0u);
            ent[e].id = ids ? ids[rd.x] : id_base + rd.x;
            if (flag) s_any_flag = 1;
        }
        __syncthreads();
        uint32_t floor_key = 0;  // uncertified entries below it cannot reach the top k (see select_warp_kernel)
        if (s_any_flag && unique_ids && cnt >= (uint32_t)k) {
            uint32_t lo = 0, hi = 0xFFFFFFFFu;
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo) / 2 + 1;
                int c = 0;
                for (uint32_t e = threadIdx.x; e < cnt; e += kSelThreads)
                    c += key_lower_bound(ent[e].key, (ent[e].row & kFlagBit) != 0) >= mid;
                for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
                if (lane == 0) s_widx[warp] = c;
                __syncthreads();
                int sum = 0;
                for (int w = 0; w < kSelThreads / 32; w++) sum += s_widx[w];
                __syncthreads();
                if (sum >= k) lo = mid;
                else hi = mid - 1;
            }
            floor_key = lo;
            int need = 0;
            for (uint32_t e = threadIdx.x; e < cnt; e += kSelThreads) need |= (ent[e].row & kFlagBit) && ent[e].key >= floor_key;
            if (__syncthreads_or(need) == 0 && threadIdx.x == 0) s_any_flag = 0;
            __syncthreads();
        }
        if (s_any_flag) {
            // literal path: normalizeVector of the query (compute/cosine.go:26,138-149), then each flagged row
            const uint8_t *qc = queries.codes + (size_t)q * queries.d_pad;
            const float2 qh = queries.hdr[q];
            const double mn = (double)qh.x, range = __dsub_rn((double)qh.y, (double)qh.x);
            for (int i = threadIdx.x; i < D; i += kSelThreads) sh_qn[i] = ref_dequant_f64(qc[i], mn, range);
            __syncthreads();
            if (warp == 0) {
                const double nsq = warp_ordered_sum(D, lane, [&](int i) { return __dmul_rn(sh_qn[i], sh_qn[i]); });
                if (lane == 0) s_norm = __dsqrt_rn(nsq);
            }
            __syncthreads();
            const double norm = s_norm;
            if (norm != 0.0)
                for (int i = threadIdx.x; i < D; i += kSelThreads) sh_qn[i] = __ddiv_rn(sh_qn[i], norm);
            __syncthreads();
            for (uint32_t e = warp; e < cnt; e += kSelThreads / 32) {
                const uint32_t r = ent[e].row;
                if ((r & kFlagBit) && ent[e].key >= floor_key) {
                    const uint32_t row = r & ~kFlagBit;
                    const float2 h = rows.hdr[row];
                    const double dot = warp_ref_cosine_row_f64(rows.codes + (size_t)row * d_pad, h.x, h.y, sh_qn, D, lane);
                    if (lane == 0) {
                        ent[e].key = f32_to_key(__double2float_rn(dot));
                        ent[e].row = row;
                        if (fix_counter) atomicAdd(fix_counter, 1ull);
                    }
                }
            }
            __syncthreads();
        }
        const uint32_t tau_key = f32_to_key(tau[q]);
        if (unique_ids && k > 16 && cnt <= 1024u) {
            // Many hits wanted from few candidates, one entry per document (the probe stage: k = nprobe up to 128): no
            // rounds.  Every candidate counts the candidates that rank before it -- (similarity desc, id asc) is a
            // strict total order here -- and the first k write themselves to their rank.
            int above = 0;
            for (uint32_t e = threadIdx.x; e < cnt; e += kSelThreads) {
                const uint32_t mk = ent[e].key;
                const uint64_t mi = ent[e].id;
                uint32_t rnk = 0;
                for (uint32_t j = 0; j < cnt; j++) rnk += cand_better(ent[j].key, ent[j].id, mk, mi) ? 1u : 0u;
                if (rnk < (uint32_t)k) {
                    out_ids[(size_t)q * k + rnk] = mi;
                    out_sims[(size_t)q * k + rnk] = key_to_f32(mk);
                    above += (mk >= tau_key && mk != 1u) ? 1 : 0;
                }
            }
            for (int o = 16; o; o >>= 1) above += __shfl_xor_sync(0xFFFFFFFFu, above, o);
            if (lane == 0) s_widx[warp] = above;
            __syncthreads();
            if (threadIdx.x == 0) {
                int sum = 0;
                for (int w = 0; w < kSelThreads / 32; w++) sum += s_widx[w];
                out_counts[q] = (int)min(cnt, (uint32_t)k);
                status[q] = (sum < k) ? kStatusNeedMore : 0u;
            }
            continue;
        }
        int emitted = 0, emitted_above = 0;
        for (int round = 0; round < k; round++) {
            uint32_t bk = 0;
            uint64_t bid = kEmptyId;
            int bidx = -1;
            for (uint32_t e = threadIdx.x; e < cnt; e += kSelThreads) {
                const uint32_t ek = ent[e].key;
                if (ek != 0 && (bidx < 0 || cand_better(ek, ent[e].id, bk, bid))) {
                    bk = ek;
                    bid = ent[e].id;
                    bidx = (int)e;
                }
            }
            for (int o = 16; o; o >>= 1) {
                const uint32_t ok = __shfl_xor_sync(0xFFFFFFFFu, bk, o);
                const uint64_t oid = __shfl_xor_sync(0xFFFFFFFFu, bid, o);
                const int oidx = __shfl_xor_sync(0xFFFFFFFFu, bidx, o);
                if (oidx >= 0 && (bidx < 0 || cand_better(ok, oid, bk, bid) || (ok == bk && oid == bid && oidx < bidx))) {
                    bk = ok;
                    bid = oid;
                    bidx = oidx;
                }
            }
            if (lane == 0) {
                s_wkey[warp] = bk;
                s_wid[warp] = bid;
                s_widx[warp] = bidx;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int best = -1;
                uint32_t k0 = 0;
                uint64_t i0 = kEmptyId;
                for (int w = 0; w < kSelThreads / 32; w++) {
                    if (s_widx[w] >= 0 && (best < 0 || cand_better(s_wkey[w], s_wid[w], k0, i0) ||
                                           (s_wkey[w] == k0 && s_wid[w] == i0 && s_widx[w] < best))) {
                        best = s_widx[w];
                        k0 = s_wkey[w];
                        i0 = s_wid[w];
                    }
                }
                s_best = best;
                if (best >= 0) {
                    out_ids[(size_t)q * k + round] = i0;
                    out_sims[(size_t)q * k + round] = key_to_f32(k0);
                }
            }
            __syncthreads();
            const int best = s_best;
            if (best < 0) break;
            const uint64_t win_id = ent[best].id;
            const uint32_t win_key = ent[best].key;
            emitted++;
            if (win_key >= tau_key && win_key != 1u) emitted_above++;
            __syncthreads();
            for (uint32_t e = threadIdx.x; e < cnt; e += kSelThreads)
                if (ent[e].id == win_id) ent[e].key = 0;  // one hit per document
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            out_counts[q] = emitted;
            // complete only if k distinct documents at or above tau were found: then no row outside the candidate
            // list (all of which score below tau) can belong to the top k
            status[q] = (emitted_above < k) ? kStatusNeedMore : 0u;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// [n][d_pad] uint8 codes, box = 128 rows x 128 bytes, SWIZZLE_128B; out-of-range rows / columns read as zero.
bool make_codes_map(CUtensorMap *tm, const MatView &m, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)m.d_pad, (cuuint64_t)m.n};
    cuuint64_t strides[1] = {(cuuint64_t)m.d_pad};
    cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(m.codes), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// A plain (unswizzled) 2-D tile map over a [rows][row_stride] byte matrix: box = box_cols bytes x box_rows rows.
// Out-of-range elements read as zero and still count towards the transaction bytes.  tm_out: CUtensorMap*.
bool make_u8_tile_map(void *tm_out, const uint8_t *base, uint64_t cols, uint64_t rows, uint64_t row_stride, uint32_t box_cols,
                      uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_stride};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(static_cast<CUtensorMap *>(tm_out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(base), dims, strides, box,
              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

namespace {

size_t gemm_smem_bytes(int kc, int ns) {
    return 1024 + (size_t)kc * kChunkBytes + (size_t)ns * kStageBytes + kAccStages * kTN * 16 +
           8 * (2 * kMaxKC + 2 * kMaxStages + 3 * kAccStages) + 64;
}

}  // namespace

bool gemm_store_supported(const MatView &rows) {
    return rows.d_pad <= kMaxKC * kChunkK && rows.n >= 1 && rows.n < 0x7FFFFF00ull && (reinterpret_cast<uintptr_t>(rows.codes) & 15) == 0;
}
bool gemm_supported(const MatView &rows, size_t nq) { return gemm_store_supported(rows) && nq >= 1 && nq <= (size_t)kMaxStageQueries; }

GemmPlan gemm_plan(const MatView &rows, size_t nq, size_t k, bool unique_ids, uint32_t sample_div, uint32_t min_sample_tiles,
                   size_t cand_per_query) {
    GemmPlan pl{};
    pl.nq_pad = (uint32_t)((nq + kTN - 1) / kTN * kTN);
    pl.tiles = (uint32_t)((rows.n + kTM - 1) / kTM);
    pl.rank = (uint32_t)(unique_ids ? k : 2 * k);
    // sample about 1/sample_div of the store, at least one tile (8 groups) per wanted rank and min_sample_tiles, at most
    // all of it
    uint32_t want = pl.tiles / sample_div;
    if (want < pl.rank) want = pl.rank;
    if (want < min_sample_tiles) want = min_sample_tiles;
    if (want > pl.tiles) want = pl.tiles;
    pl.sample_stride = pl.tiles / want;
    if (pl.sample_stride < 1) pl.sample_stride = 1;
    pl.sample_tiles = (pl.tiles + pl.sample_stride - 1) / pl.sample_stride;
    pl.G = pl.sample_tiles * 8;  // one group maximum per 16 store rows
    if (cand_per_query > (size_t)kSelCap) cand_per_query = kSelCap;
    if (cand_per_query < 32) cand_per_query = 32;
    pl.cand_per_q = (uint32_t)cand_per_query;
    return pl;
}

size_t gemm_scratch_bytes(const GemmPlan &pl, size_t nq) {
    auto pad = [](size_t b) { return (b + 255) & ~size_t(255); };
    return pad((size_t)pl.nq_pad * 16) + pad((size_t)pl.nq_pad * pl.G * 4) + pad(nq * 4) + pad(64) + pad((size_t)pl.nq_pad * 4) +
           pad((size_t)pl.nq_pad * pl.cand_per_q * 8) + 4096;
}

void gemm_take(char *base, const GemmPlan &pl, size_t nq, GemmBufs *b) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        off = (off + 255) & ~size_t(255);
        char *p = base + off;
        off += bytes;
        return p;
    };
    b->col_consts = reinterpret_cast<float4 *>(take((size_t)pl.nq_pad * 16));
    b->gmax = reinterpret_cast<float *>(take((size_t)pl.nq_pad * pl.G * 4));
    b->tau = reinterpret_cast<float *>(take(nq * 4));
    b->bounds = reinterpret_cast<unsigned int *>(take(64));
    b->cand_cnt = reinterpret_cast<unsigned int *>(take((size_t)pl.nq_pad * 4));
    b->cand_rowdot = reinterpret_cast<uint2 *>(take((size_t)pl.nq_pad * pl.cand_per_q * 8));
}

static cudaError_t launch_gemm(int mode, const CUtensorMap &tm_rows, const CUtensorMap &tm_q, const GemmParams &p, int sm_count,
                               cudaStream_t st) {
    const size_t smem = gemm_smem_bytes(p.kc, p.nstages);
    const uint32_t n_items = p.tile_count * p.qsplit;
    unsigned grid = n_items < (uint32_t)sm_count ? n_items : (unsigned)sm_count;
    if (grid == 0) return cudaSuccess;
    cudaError_t e;
    if (mode == MODE_FILTER) {
        e = cudaFuncSetAttribute(gemm_kernel<MODE_FILTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gemm_kernel<MODE_FILTER><<<grid, kGemmThreads, smem, st>>>(tm_rows, tm_q, p);
    } else {
        e = cudaFuncSetAttribute(gemm_kernel<MODE_GROUPMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        gemm_kernel<MODE_GROUPMAX><<<grid, kGemmThreads, smem, st>>>(tm_rows, tm_q, p);
    }
    return cudaGetLastError();
}

// A store with few tiles is cut into (store tile, query-tile range) work items so that every SM has several; a range keeps
// at least 4 query tiles so that reloading the resident store tile stays amortized.  Every range is non-empty.
static void gemm_split(GemmParams &p, int sm_count) {
    uint32_t qs = 1;
    if (p.tile_count < 4u * (uint32_t)sm_count && p.tile_count > 0) {
        qs = (4u * (uint32_t)sm_count + p.tile_count - 1) / p.tile_count;
        const uint32_t max_qs = p.nq_tiles / 4 > 1 ? p.nq_tiles / 4 : 1;
        if (qs > max_qs) qs = max_qs;
    }
    p.qt_per = (p.nq_tiles + qs - 1) / qs;
    p.qsplit = (p.nq_tiles + p.qt_per - 1) / p.qt_per;
}

static GemmParams gemm_params(const MatView &rows, const GemmPlan &pl, const GemmBufs &b) {
    GemmParams p{};
    p.row_hdr = rows.hdr;
    p.row_sums = rows.sums;
    p.n = (uint32_t)rows.n;
    p.D = rows.d;
    p.kc = (rows.d_pad + kChunkK - 1) / kChunkK;
    {   // kc * 16 KB + stages * kStageBytes + ~10 KB <= 227 KB
        const int room = (227 * 1024 - 10 * 1024 - p.kc * kChunkBytes) / kStageBytes;
        p.nstages = room < 6 ? room : 6;
    }
    p.nq_tiles = pl.nq_pad / kTN;
    p.col_consts = b.col_consts;
    p.cand_cnt = b.cand_cnt;
    p.cand_rowdot = b.cand_rowdot;
    p.cand_per_q = pl.cand_per_q;
    p.gmax = b.gmax;
    p.G = pl.G;
    const char *dbg = getenv("VS_GEMM_DBG");
    p.dbg = dbg ? atoi(dbg) : 0;
    return p;
}

// Phase 1a: thresholds from the sampled pre-pass (same GEMM, group maxima instead of the filter).
cudaError_t gemm_enqueue_prepass(const MatView &rows, const MatView &queries, const GemmPlan &pl, const GemmBufs &b, uint32_t *d_status,
                                 int sm_count, cudaStream_t st, uint64_t *launches) {
    CUtensorMap tm_rows, tm_q;
    if (!make_codes_map(&tm_rows, rows, kTM) || !make_codes_map(&tm_q, queries, kTN)) return cudaErrorNotSupported;
    cudaError_t e = cudaMemsetAsync(b.bounds, 0, 64, st);
    if (e != cudaSuccess) return e;
    row_bounds_kernel<<<sm_count * 4, 256, 0, st>>>(rows.hdr, rows.sums, (uint32_t)rows.n, rows.d, b.bounds);
    query_consts_groupmax_kernel<<<(pl.nq_pad + 127) / 128, 128, 0, st>>>(queries, pl.nq_pad, b.col_consts);
    GemmParams p = gemm_params(rows, pl, b);
    p.tile_first = 0;
    p.tile_stride = pl.sample_stride;
    p.tile_count = pl.sample_tiles;
    gemm_split(p, sm_count);
    e = launch_gemm(MODE_GROUPMAX, tm_rows, tm_q, p, sm_count, st);
    if (e != cudaSuccess) return e;
    if (pl.G <= 1024u) {
        const unsigned blocks = (pl.nq_pad + 3) / 4 < (unsigned)sm_count * 16 ? (pl.nq_pad + 3) / 4 : (unsigned)sm_count * 16;
        if (pl.G <= 256u)
            threshold_warp_kernel<8><<<blocks, 128, 0, st>>>(queries, pl.nq_pad, b.gmax, pl.G, pl.rank, b.bounds, b.col_consts, b.tau,
                                                             d_status);
        else
            threshold_warp_kernel<32><<<blocks, 128, 0, st>>>(queries, pl.nq_pad, b.gmax, pl.G, pl.rank, b.bounds, b.col_consts,
                                                              b.tau, d_status);
    } else {
        threshold_kernel<<<pl.nq_pad, kThrThreads, pl.G * sizeof(int), st>>>(queries, pl.nq_pad, b.gmax, pl.G, pl.rank, b.bounds,
                                                                             b.col_consts, b.tau, d_status);
    }
    e = cudaMemsetAsync(b.cand_cnt, 0, (size_t)pl.nq_pad * 4, st);
    if (e != cudaSuccess) return e;
    if (launches) *launches += 4;
    return cudaGetLastError();
}

// Phase 1b: the filtering GEMM over the whole store.  Asynchronous; the candidate count lands in b.bounds[4].
cudaError_t gemm_enqueue_filter(const MatView &rows, const MatView &queries, const GemmPlan &pl, const GemmBufs &b, int sm_count,
                                cudaStream_t st, uint64_t *launches) {
    CUtensorMap tm_rows, tm_q;
    if (!make_codes_map(&tm_rows, rows, kTM) || !make_codes_map(&tm_q, queries, kTN)) return cudaErrorNotSupported;
    GemmParams p = gemm_params(rows, pl, b);
    p.tile_first = 0;
    p.tile_stride = 1;
    p.tile_count = pl.tiles;
    gemm_split(p, sm_count);
    const cudaError_t e = launch_gemm(MODE_FILTER, tm_rows, tm_q, p, sm_count, st);
    if (e != cudaSuccess) return e;
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// Phase 2: finish the candidates of every query (no host involvement: the buckets are already grouped by query).
cudaError_t gemm_enqueue_select(const MatView &rows, const uint64_t *ids, uint64_t id_base, bool unique_ids, const MatView &queries, const GemmPlan &pl,
                                const GemmBufs &b, int k, uint64_t *d_ids, float *d_sims, int32_t *d_counts, uint32_t *d_status,
                                unsigned long long *fix_counter, int sm_count, cudaStream_t st, uint64_t *launches) {
    const uint32_t nq = (uint32_t)queries.n;
    // many queries with few candidates each (assignment): the warp kernel takes them, the block kernel the rest
    const bool two_level = k <= 32 && nq >= 4u * (uint32_t)sm_count;
    if (two_level) {
        const unsigned wb = (nq + 7) / 8 < (unsigned)sm_count * 8 ? (nq + 7) / 8 : (unsigned)sm_count * 8;
        select_warp_kernel<<<wb, kSelThreads, 0, st>>>(rows, ids, id_base, unique_ids ? 1 : 0, queries, b.cand_cnt, b.cand_rowdot, pl.cand_per_q, b.tau, k,
                                                       d_ids, d_sims, d_counts, d_status);
        if (launches) *launches += 1;
    }
    const size_t smem = sizeof(SelEntry) * kSelCap + (size_t)rows.d * 8;
    cudaError_t e = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = nq < (unsigned)sm_count * 3 ? nq : (unsigned)sm_count * 3;
    select_kernel<<<grid, kSelThreads, smem, st>>>(rows, ids, id_base, unique_ids ? 1 : 0, queries, b.cand_cnt, b.cand_rowdot, pl.cand_per_q,
                                                   two_level ? 0 : 1, b.tau, k, d_ids, d_sims, d_counts, d_status, fix_counter);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t gemm_set_certify_scale(float scale) { return cudaMemcpyToSymbol(c_certify_scale, &scale, sizeof(float)); }

}  // namespace vs
