// fused.cu -- one query, one launch: centroid scoring, nprobe cut, posting-list scan and top-k in a single
// cooperative kernel whose rows arrive through a TMA (cp.async.bulk) shared-memory ring.
//
// Replaces, for a single query: server/search.go:214 (centroid scoring, compute/cosine.go:13-57), :220-223 (sort and
// nprobe cut), :241-273 (posting-list scan, running sort, one hit per document, truncate).  scan.cu's two-launch form
// of the same path spends most of a single query's time on fixed costs: two launches, load -> score -> load rounds with
// nothing in flight while a warp scores, and a last block that merges 296 partial lists.  Here
//
//   * every block (one per SM) keeps a ring of 16-row stages in shared memory; one producer thread issues the bulk
//     copies (row codes, row headers, integer sums, document ids) for every stage of the ring at once and refills a
//     stage the moment its consumer warp releases it, so all of a block's rows are in flight from the first cycle and
//     no load waits for scoring;
//   * the probe stage and the list stage share the launch: blocks score C/gridDim centroids each, meet at a grid
//     barrier (cooperative launch: all blocks are resident), and every block selects the same nprobe best centroids
//     from the shared key array (group maxima give a threshold, the few survivors are sorted);
//   * a block publishes only `pub` (>= k + 6) distinct documents, so the last block sorts ~pub candidates instead of
//     merging gridDim lists: the best k distinct documents of the union are among the best k of every part.
//
// Scores are the certified integer-identity scores of common.cuh.  A query whose emitted window contains a candidate
// whose float32 rounding could not be certified is NOT repaired here (rare: ~1e-4 of rows at cos ~ 0.1): its status
// word gets the same bits scan.cu uses and the caller's resolve path (literal arithmetic) finishes it.
#include "internal.h"
#include "topk.cuh"

namespace vs {

namespace {

constexpr int kFusedWarps = 8;                        // consumer warps
constexpr int kFusedThreads = 32 * (kFusedWarps + 1);  // + the producer warp
constexpr int kFusedTileRows = 16;                    // rows per ring stage
constexpr int kFusedMaxStages = 16;
constexpr int kFusedSmemBudget = 227 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long fused_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One ring stage in shared memory.  The 8-byte-per-row side arrays are copied from a 16-byte aligned source address, so
// they carry `skew` (0 or 8) bytes of lead-in; the descriptor is written by the producer before it arms the barrier.
struct StageDesc {
    uint32_t row0;   // first store row of the stage
    uint32_t nrows;  // rows in the stage (<= kFusedTileRows)
    uint32_t skew;   // lead-in bytes of the side arrays
    uint32_t pad;
};

struct FusedCtl {
    uint64_t full[kFusedMaxStages];
    uint64_t empty[kFusedMaxStages];
    StageDesc desc[kFusedMaxStages];
    uint64_t seg_start[kMaxSeg];
    uint32_t seg_len[kMaxSeg];
    uint32_t seg_prefix[kMaxSeg + 1];  // rows before segment s in the query's concatenated row sequence
    SideConst qside;
    uint32_t nchunks;       // chunks of the current phase for this block
    int nseg;               // segments (probed lists) of the list stage
    uint32_t thr;           // selection threshold
    unsigned int cnt;       // survivors / collected candidates
    unsigned int overflow;
    unsigned int abort_status;  // non-zero: the probe stage could not be decided here; bits for the status word
    unsigned int is_last;
    int warp_cnt[kFusedWarps];
    alignas(16) uint32_t gmax[kFusedThreads];
};

#define FUSED_TRACE(slot)                                                                                   \
    do {                                                                                                    \
        if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 16 + (slot)] = fused_timer();          \
    } while (0)

// Integer dots of one stage: lane group g (G lanes) takes rows g*iters .. g*iters+iters-1; returns in lane g*G+it the
// dot of row g*iters+it.
template <int G, int CPL>
__device__ __forceinline__ uint32_t stage_dots(uint32_t s_codes, int nrows, int d_pad, const uint4 (&q)[CPL], int lane, int iters) {
    constexpr int U = (CPL <= 3) ? 4 : 2;
    const int g = lane / G, l = lane % G;
    uint32_t mydot = 0;
#pragma unroll 1
    for (int it0 = 0; it0 < iters; it0 += U) {
        uint4 v[U][CPL];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int r = g * iters + it0 + u;
            const bool ok = (it0 + u < iters) && (r < nrows);
            const uint32_t a = s_codes + (uint32_t)(ok ? r : 0) * (uint32_t)d_pad + (uint32_t)l * 16u;
#pragma unroll
            for (int j = 0; j < CPL; j++) v[u][j] = ok ? lds_u4(a + (uint32_t)(j * G * 16)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < CPL; j++) acc = dot16(v[u][j], q[j], acc);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
            if (l == it0 + u) mydot = acc;
        }
    }
    return mydot;
}

// The rows [r_lo, r_hi) of the query's concatenated segment sequence as chunks of <= chunk_rows (<= kFusedTileRows) rows
// that do not cross a segment.  walk(seg, first row within the segment, rows) is called per chunk, in order.
template <typename F>
__device__ __forceinline__ uint32_t for_each_chunk(const FusedCtl &ctl, int nseg, uint64_t r_lo, uint64_t r_hi, uint32_t chunk_rows,
                                                   F walk) {
    uint32_t n = 0;
    for (int s = 0; s < nseg; s++) {
        const uint64_t ps = ctl.seg_prefix[s], pe = ctl.seg_prefix[s + 1];
        if (pe <= r_lo) continue;
        if (ps >= r_hi) break;
        const uint32_t a = (uint32_t)((r_lo > ps ? r_lo : ps) - ps), b = (uint32_t)((r_hi < pe ? r_hi : pe) - ps);
        for (uint32_t o = a; o < b; o += chunk_rows) {
            walk(s, o, min(chunk_rows, b - o));
            n++;
        }
    }
    return n;
}

}  // namespace

template <int G, int CPL, int KPL>
__global__ void __launch_bounds__(kFusedThreads, 1)
fused_search_kernel(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char fsm[];
    constexpr int CAP = 32 * KPL;
    constexpr int NG = 32 / G;
    constexpr int TR = kFusedTileRows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = p.rows.d, d_pad = p.rows.d_pad;
    const uint32_t side_bytes = (TR + 2) * 8;
    const uint32_t stage_bytes = (uint32_t)p.stage_bytes;
    const int S = p.stages;
    unsigned char *ring = fsm;
    FusedCtl &ctl = *reinterpret_cast<FusedCtl *>(fsm + (size_t)S * stage_bytes);
    // the sort buffers overlay the ring: they are used only while no bulk copy is in flight
    SortSmem &ss = *reinterpret_cast<SortSmem *>(fsm);
    const CandBuf bufA{ss.key_a, ss.meta_a, ss.id_a};
    const CandBuf bufB{ss.key_b, ss.meta_b, ss.id_b};
    const uint32_t Gd = gridDim.x;
    const bool has_probe = p.npe > 0;
    const bool dedup = p.ids != nullptr;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; i++) {
            mbar_init(smem_u32(&ctl.full[i]), 1);
            mbar_init(smem_u32(&ctl.empty[i]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const float2 h = p.query.hdr[0];
        const uint2 s = p.query.sums[0];
        ctl.qside = make_side(h.x, h.y, s.x, s.y, D);
        ctl.abort_status = 0;
        ctl.overflow = 0;
        ctl.cnt = 0;
        ctl.thr = 0;
    }
    uint4 qreg[CPL];
#pragma unroll
    for (int j = 0; j < CPL; j++) qreg[j] = *reinterpret_cast<const uint4 *>(p.query.codes + ((lane % G) + G * j) * 16);
    __syncthreads();
    const SideConst xq = ctl.qside;
    FUSED_TRACE(0);

    uint32_t chunk_base = 0;  // ring position continues across the two phases (stage = chunk % S, parity from chunk / S)

    // One phase of ring traffic: the producer thread streams this block's chunks of `m` (a matrix: centroid table or
    // store), the consumer warps take chunk c = warp, warp + 8, ... and call `use(stage, desc)` on each.
    auto produce = [&](const MatView &m, const uint64_t *ids, int nseg, uint64_t r_lo, uint64_t r_hi, uint32_t chunk_rows) {
        uint32_t c = chunk_base;
        // the ring may have served as sort buffers (generic-proxy writes) since the last bulk copy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for_each_chunk(ctl, nseg, r_lo, r_hi, chunk_rows, [&](int s, uint32_t o, uint32_t nr) {
            const uint32_t st = c % (uint32_t)S;
            if (c >= (uint32_t)S) mbar_wait(smem_u32(&ctl.empty[st]), ((c / (uint32_t)S) & 1u) ^ 1u);
            const uint64_t row0 = ctl.seg_start[s] + o;
            const uint32_t skew = (uint32_t)(row0 & 1u) * 8u;
            const uint32_t cb = nr * (uint32_t)d_pad;
            const uint32_t sb = (skew + nr * 8u + 15u) & ~15u;
            ctl.desc[st] = StageDesc{(uint32_t)row0, nr, skew, 0u};
            const uint32_t full = smem_u32(&ctl.full[st]);
            const uint32_t dst = smem_u32(ring + (size_t)st * stage_bytes);
            mbar_expect_tx(full, cb + sb * (ids ? 3u : 2u));
            bulk_g2s(dst, m.codes + row0 * (uint64_t)d_pad, cb, full);
            const uint64_t ra = row0 & ~uint64_t(1);
            bulk_g2s(dst + (uint32_t)TR * d_pad, m.hdr + ra, sb, full);
            bulk_g2s(dst + (uint32_t)TR * d_pad + side_bytes, m.sums + ra, sb, full);
            if (ids) bulk_g2s(dst + (uint32_t)TR * d_pad + 2 * side_bytes, ids + ra, sb, full);
            c++;
        });
    };

    // ================= probe stage: score my share of the centroid table, meet, select =================
    if (has_probe) {
        const uint64_t C = p.cent.n;
        const uint64_t c_lo = (C * blockIdx.x) / Gd, c_hi = (C * (blockIdx.x + 1ull)) / Gd;
        // a block's share of the table is small (28 rows of 4096 on 148 SMs): chunks sized so that every warp gets one
        uint32_t pchunk = (uint32_t)((c_hi - c_lo + kFusedWarps - 1) / kFusedWarps);
        pchunk = pchunk < 2 ? 2 : (pchunk > (uint32_t)TR ? (uint32_t)TR : pchunk);
        if (threadIdx.x == 0) {
            ctl.seg_start[0] = 0;
            ctl.seg_len[0] = (uint32_t)C;
            ctl.seg_prefix[0] = 0;
            ctl.seg_prefix[1] = (uint32_t)C;
            ctl.nchunks = for_each_chunk(ctl, 1, c_lo, c_hi, pchunk, [](int, uint32_t, uint32_t) {});
        }
        __syncthreads();
        const uint32_t nch = ctl.nchunks;
        if (warp == kFusedWarps) {
            if (lane == 0) produce(p.cent, nullptr, 1, c_lo, c_hi, pchunk);
        } else {
            for (uint32_t c = warp; c < nch; c += kFusedWarps) {
                const uint32_t gc = chunk_base + c, st = gc % (uint32_t)S;
                mbar_wait(smem_u32(&ctl.full[st]), (gc / (uint32_t)S) & 1u);
                const StageDesc de = ctl.desc[st];
                const unsigned char *sp = ring + (size_t)st * stage_bytes;
                const int iters = ((int)de.nrows + NG - 1) / NG;
                const uint32_t mydot = stage_dots<G, CPL>(smem_u32(sp), (int)de.nrows, d_pad, qreg, lane, iters);
                const int myr = (lane / G) * iters + (lane % G);
                const bool valid = (lane % G) < iters && myr < (int)de.nrows;
                if (valid) {
                    const float2 h = *reinterpret_cast<const float2 *>(sp + TR * d_pad + de.skew + myr * 8);
                    const uint2 s = *reinterpret_cast<const uint2 *>(sp + TR * d_pad + side_bytes + de.skew + myr * 8);
                    bool flag;
                    const float sim = score_fast(xq, h.x, h.y, s.x, s.y, mydot, D, &flag);
                    const uint32_t ci = de.row0 + (uint32_t)myr;
                    p.keys[ci] = f32_to_key(sim);
                    if (flag) {
                        const unsigned int pos = atomicAdd(p.flag_cnt, 1u);
                        if (pos < (unsigned)kFusedFlagCap) p.flag_list[pos] = ci;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl.empty[st]));
            }
        }
        chunk_base += nch;
        FUSED_TRACE(1);
        // ---- grid barrier (cooperative launch: every block is resident) ----
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(&p.sync[0], 1u);
            while (ld_acquire_u32(&p.sync[0]) < Gd) {
            }
        }
        __syncthreads();
        FUSED_TRACE(2);
        // ---- every block selects the same npe best centroids (search.go:220-223; ties: lower index first) ----
        // group maxima: the npe-th largest of them is a lower bound of the npe-th largest key
        const uint32_t P = (uint32_t)(((C + kFusedThreads - 1) / kFusedThreads + 3) & ~uint64_t(3));
        const uint64_t k_lo = (uint64_t)threadIdx.x * P;
        uint32_t gm = 0;
        for (uint32_t i = 0; i < P && k_lo + i < C; i += 4) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p.keys + k_lo + i));
            gm = max(gm, v.x);
            if (k_lo + i + 1 < C) gm = max(gm, v.y);
            if (k_lo + i + 2 < C) gm = max(gm, v.z);
            if (k_lo + i + 3 < C) gm = max(gm, v.w);
        }
        ctl.gmax[threadIdx.x] = gm;
        __syncthreads();
        {
            int r = 0;
            for (int u = 0; u < kFusedThreads; u += 4) {
                const uint4 o = *reinterpret_cast<const uint4 *>(&ctl.gmax[u]);
                r += (o.x > gm) || (o.x == gm && u + 0 < (int)threadIdx.x);
                r += (o.y > gm) || (o.y == gm && u + 1 < (int)threadIdx.x);
                r += (o.z > gm) || (o.z == gm && u + 2 < (int)threadIdx.x);
                r += (o.w > gm) || (o.w == gm && u + 3 < (int)threadIdx.x);
            }
            if (r == p.npe - 1) ctl.thr = gm;  // stays 0 with fewer than npe groups: everything survives
        }
        __syncthreads();
        const uint32_t thr = ctl.thr;
        for (uint32_t i = 0; i < P && k_lo + i < C; i += 4) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p.keys + k_lo + i));
            const uint32_t kk[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (k_lo + i + t < C && kk[t] != 0 && kk[t] >= thr) {
                    const unsigned int pos = atomicAdd(&ctl.cnt, 1u);
                    if (pos < (unsigned)kOutCap) cand_put(bufA, (int)pos, kk[t], 0u, k_lo + i + t);
                    else ctl.overflow = 1;
                }
            }
        }
        __syncthreads();
        const int nsurv = (int)min(ctl.cnt, (unsigned)kOutCap);
        const bool sel_overflow = ctl.overflow != 0;
        __syncthreads();
        block_sort_small(bufA, nsurv, bufB, nsurv);
        const int npe = min(p.npe, nsurv);  // (= p.npe: every centroid has a non-zero key and the caller keeps npe < C)
        if (threadIdx.x == 0) ctl.nseg = npe;
        // an uncertified similarity inside the window (or a survivor list that did not fit): the caller's literal path
        {
            const unsigned int nf = __ldcg(p.flag_cnt);
            bool bad = sel_overflow || nf > (unsigned)kFusedFlagCap;
            if (!bad && (int)threadIdx.x < npe) {
                const uint32_t ci = (uint32_t)bufB.id[threadIdx.x];
                for (unsigned int j = 0; j < nf; j++) bad |= __ldcg(p.flag_list + j) == ci;
            }
            if (bad) ctl.abort_status = kStatusProbeAmbiguous;
        }
        __syncthreads();
        if ((int)threadIdx.x < npe) {
            const uint32_t L = (uint32_t)bufB.id[threadIdx.x];
            const uint64_t st = p.list_off[L];
            ctl.seg_start[threadIdx.x] = st;
            ctl.seg_len[threadIdx.x] = (uint32_t)(p.list_off[L + 1] - st);
            if (p.out_probe && blockIdx.x == 0) p.out_probe[threadIdx.x] = L;
        }
        __syncthreads();
        if (warp == 0) {  // rows before each segment
            uint32_t carry = 0;
            for (int base = 0; base < npe; base += 32) {
                const int s = base + lane;
                uint32_t v = s < npe ? ctl.seg_len[s] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, v, o);
                    if (lane >= o) v += t;
                }
                if (s < npe) ctl.seg_prefix[s + 1] = carry + v;
                carry += __shfl_sync(FULL, v, 31);
            }
            if (lane == 0) ctl.seg_prefix[0] = 0;
        }
        __syncthreads();
        FUSED_TRACE(3);
    } else if (threadIdx.x == 0) {
        ctl.seg_start[0] = p.flat_start;
        ctl.seg_len[0] = (uint32_t)p.flat_count;
        ctl.seg_prefix[0] = 0;
        ctl.seg_prefix[1] = (uint32_t)p.flat_count;
        ctl.nseg = 1;
    }
    __syncthreads();

    // ================= list stage: my share of the probed rows through the ring =================
    const bool aborted = ctl.abort_status != 0;
    WarpTopK<KPL> top;
    top.init();
    if (!aborted) {
        const int ns = ctl.nseg;
        const uint64_t Trows = ctl.seg_prefix[ns];
        const uint64_t r_lo = (Trows * blockIdx.x) / Gd, r_hi = (Trows * (blockIdx.x + 1ull)) / Gd;
        if (threadIdx.x == 0) ctl.nchunks = for_each_chunk(ctl, ns, r_lo, r_hi, (uint32_t)TR, [](int, uint32_t, uint32_t) {});
        __syncthreads();
        const uint32_t nch = ctl.nchunks;
        if (warp == kFusedWarps) {
            if (lane == 0) produce(p.rows, p.ids, ns, r_lo, r_hi, (uint32_t)TR);
        } else {
            for (uint32_t c = warp; c < nch; c += kFusedWarps) {
                const uint32_t gc = chunk_base + c, st = gc % (uint32_t)S;
                mbar_wait(smem_u32(&ctl.full[st]), (gc / (uint32_t)S) & 1u);
                const StageDesc de = ctl.desc[st];
                const unsigned char *sp = ring + (size_t)st * stage_bytes;
                const int iters = ((int)de.nrows + NG - 1) / NG;
                const uint32_t mydot = stage_dots<G, CPL>(smem_u32(sp), (int)de.nrows, d_pad, qreg, lane, iters);
                const int myr = (lane / G) * iters + (lane % G);
                const bool valid = (lane % G) < iters && myr < (int)de.nrows;
                uint32_t key = 0, meta = 0;
                uint64_t cid = kEmptyId;
                if (valid) {
                    const float2 h = *reinterpret_cast<const float2 *>(sp + TR * d_pad + de.skew + myr * 8);
                    const uint2 s = *reinterpret_cast<const uint2 *>(sp + TR * d_pad + side_bytes + de.skew + myr * 8);
                    bool flag;
                    const float sim = score_fast(xq, h.x, h.y, s.x, s.y, mydot, D, &flag);
                    key = f32_to_key(sim);
                    const uint32_t row = de.row0 + (uint32_t)myr;
                    meta = row | (flag ? kFlagBit : 0u);
                    cid = p.ids ? *reinterpret_cast<const uint64_t *>(sp + TR * d_pad + 2 * side_bytes + de.skew + myr * 8)
                                : p.id_base + row;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl.empty[st]));  // everything this warp needs is in registers
                top.offer(valid, key, meta, cid, lane, dedup);
            }
        }
        chunk_base += nch;
    }
    FUSED_TRACE(4);
    __syncthreads();  // every bulk copy of this block has landed and been consumed: the ring is free for the sort buffers

    // ================= block merge -> publish `pub` distinct documents -> last block finishes the query =================
    const int pub = p.pub;
    if (!aborted) {
        if (warp < kFusedWarps) {
            const int cnt = top.count();
            if (lane == 0) ctl.warp_cnt[warp] = cnt;
            top.store_soa(bufA, warp * CAP, lane);
        }
        __syncthreads();
        const CandBuf res = merge_lists(bufA, kFusedWarps, CAP, ctl.warp_cnt, bufB, CAP, dedup, ss.scan_tmp);
        uint4 *slot = p.partial + (size_t)blockIdx.x * pub;
        for (int e = threadIdx.x; e < pub; e += blockDim.x) {
            uint4 v;
            v.x = res.key[e];
            v.y = res.meta[e];
            v.z = (uint32_t)res.id[e];
            v.w = (uint32_t)(res.id[e] >> 32);
            slot[e] = v;
        }
    }
    FUSED_TRACE(5);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(&p.sync[1], 1u);
        ctl.is_last = (tk == Gd - 1) ? 1u : 0u;
        ctl.cnt = 0;
        ctl.overflow = 0;
        ctl.thr = 0;
    }
    __syncthreads();
    if (!ctl.is_last) return;
    __threadfence();
    if (threadIdx.x == 0) {  // re-arm for the next launch: every block has passed the barrier and taken its ticket
        p.sync[0] = 0;
        p.sync[1] = 0;
        *p.flag_cnt = 0;
    }
    FUSED_TRACE(6);
    if (aborted) {
        if (threadIdx.x == 0) {
            p.out_counts[0] = 0;
            p.out_status[0] = ctl.abort_status;
        }
        return;
    }
    // ---- all published entries (gridDim x pub): threshold from the slot heads, collect, sort, one hit per document ----
    const int total = (int)Gd * pub;
    uint32_t status = 0;
    bool slow = Gd > (uint32_t)kFusedThreads;
    uint32_t tkey = 0;
    CandBuf fin = bufB;
    if (!slow) {
        for (int u = threadIdx.x; u < kFusedThreads; u += blockDim.x)
            ctl.gmax[u] = (uint32_t)u < Gd ? __ldcg(p.partial + (size_t)u * pub).x : 0u;  // slot heads
        __syncthreads();
        if (threadIdx.x < Gd) {
            const uint32_t mine = ctl.gmax[threadIdx.x];
            int r = 0;
            for (int u = 0; u < kFusedThreads; u += 4) {
                const uint4 o = *reinterpret_cast<const uint4 *>(&ctl.gmax[u]);
                r += (o.x > mine) || (o.x == mine && u + 0 < (int)threadIdx.x);
                r += (o.y > mine) || (o.y == mine && u + 1 < (int)threadIdx.x);
                r += (o.z > mine) || (o.z == mine && u + 2 < (int)threadIdx.x);
                r += (o.w > mine) || (o.w == mine && u + 3 < (int)threadIdx.x);
            }
            if (r == pub - 1) ctl.thr = mine;  // the pub-th best head: at least pub entries reach it
        }
        __syncthreads();
        tkey = ctl.thr;
#pragma unroll 4
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const uint4 v = __ldcg(p.partial + e);
            if (v.x != 0 && v.x >= tkey) {
                const unsigned int pos = atomicAdd(&ctl.cnt, 1u);
                if (pos < (unsigned)kOutCap) cand_put(bufA, (int)pos, v.x, v.y, (uint64_t)v.z | ((uint64_t)v.w << 32));
                else ctl.overflow = 1;
            }
        }
        __syncthreads();
        slow = ctl.overflow != 0;
        if (!slow) {
            const int n = (int)ctl.cnt;
            __syncthreads();
            block_sort_small(bufA, n, bufB, max(n, CAP));
            if (dedup) {
                const int uniq = block_unique_compact(bufB, n, bufA, CAP, ss.scan_tmp);
                fin = bufA;
                slow = uniq < p.k && tkey != 0;  // duplicates across blocks ate the margin: look at every entry
            }
        }
    }
    if (slow) {
        __syncthreads();
        if (warp == 0) {
            top.init();
            for (uint32_t b = 0; b < Gd; b++) {
                for (int base = 0; base < pub; base += 32) {
                    const int r = base + lane;
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (r < pub) v = __ldcg(p.partial + (size_t)b * pub + r);
                    const uint64_t id = (uint64_t)v.z | ((uint64_t)v.w << 32);
                    const bool any = __any_sync(FULL, v.x != 0 && cand_better(v.x, id, top.thr_key, top.thr_id));
                    if (!any) break;
                    top.offer(v.x != 0, v.x, v.y, id, lane, dedup);
                }
            }
            top.store_soa(bufB, 0, lane);
        }
        fin = bufB;
        __syncthreads();
    }
    FUSED_TRACE(7);
    // ---- emit the first k (search.go:270); an uncertified score among them goes to the caller's literal path ----
    if (warp == 0) {
        int n = 0;
        bool anyflag = false;
        for (int base = 0; base < CAP; base += 32) {
            const int r = base + lane;
            const uint32_t kk = fin.key[r];
            const bool have = kk != 0;
            if (have && r < p.k) {
                p.out_ids[r] = fin.id[r];
                p.out_sims[r] = key_to_f32(kk);
                anyflag |= (fin.meta[r] & kFlagBit) != 0;
            }
            n += __popc(__ballot_sync(FULL, have));
        }
        anyflag = __any_sync(FULL, anyflag);
        if (anyflag) status |= kStatusListAmbiguous;
        if (lane == 0) {
            p.out_counts[0] = min(n, p.k);
            p.out_status[0] = status;
        }
    }
    FUSED_TRACE(8);
}

// ---------------------------------------------------------------------------------------------------
bool fused_supported(int d_pad) {
    switch (d_pad >> 4) {
        case 48: case 32: case 64: case 96: case 24: return (d_pad & 15) == 0;
        default: return false;
    }
}

static void fused_geometry(int d_pad, bool with_ids, int *stage_bytes, int *stages, size_t *smem) {
    (void)with_ids;
    int sb = kFusedTileRows * d_pad + 3 * (kFusedTileRows + 2) * 8;
    sb = (sb + 127) & ~127;
    const int ctl = (int)((sizeof(FusedCtl) + 127) & ~size_t(127));
    int S = (kFusedSmemBudget - ctl) / sb;
    if (S > kFusedMaxStages) S = kFusedMaxStages;
    // the sort buffers overlay the ring: it must be at least that large
    while ((size_t)S * sb < sizeof(SortSmem)) S++;
    *stage_bytes = sb;
    *stages = S;
    *smem = (size_t)S * sb + ctl;
}

int fused_pub(int k, int kpl) {
    const int cap = 32 * kpl;
    const int pub = ((k + 6 + 15) / 16) * 16;
    return pub < cap ? pub : cap;
}

template <int G, int CPL, int KPL>
static cudaError_t launch_fused_t(FusedParams p, int grid, cudaStream_t st) {
    size_t smem;
    fused_geometry(p.rows.d_pad, p.ids != nullptr, &p.stage_bytes, &p.stages, &smem);
    if (p.stages < 2) return cudaErrorInvalidValue;
    auto kern = fused_search_kernel<G, CPL, KPL>;
    static int max_grid = 0;  // per instantiation: blocks a cooperative launch can hold
    if (!max_grid) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0, dev = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFusedThreads, smem);
        if (e != cudaSuccess) return e;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (per_sm < 1) return cudaErrorCooperativeLaunchTooLarge;
        max_grid = sms;  // one block per SM (its shared-memory ring fills the SM)
    }
    if (grid > max_grid) grid = max_grid;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kFusedThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int KPL>
static cudaError_t launch_fused_k(const FusedParams &p, int grid, cudaStream_t st) {
    switch (p.rows.d_pad >> 4) {
        case 48: return launch_fused_t<16, 3, KPL>(p, grid, st);  // 768-d (nomic-embed-text)
        case 32: return launch_fused_t<32, 1, KPL>(p, grid, st);  // 512-d (noop/ai.go)
        case 64: return launch_fused_t<32, 2, KPL>(p, grid, st);  // 1024-d
        case 96: return launch_fused_t<32, 3, KPL>(p, grid, st);  // 1536-d
        case 24: return launch_fused_t<8, 3, KPL>(p, grid, st);   // 384-d
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t fused_set_certify_scale(float scale) { return cudaMemcpyToSymbol(c_certify_scale, &scale, sizeof(float)); }

cudaError_t launch_fused_search(const FusedParams &p, int kpl, int grid, cudaStream_t st) {
    switch (kpl) {
        case 1: return launch_fused_k<1>(p, grid, st);
        case 2: return launch_fused_k<2>(p, grid, st);
        case 4: return launch_fused_k<4>(p, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace vs
