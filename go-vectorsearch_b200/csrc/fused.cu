// fused.cu -- one query, one launch: centroid scoring, nprobe cut, posting-list scan and top-k in a single
// cooperative kernel whose rows arrive through a TMA (cp.async.bulk) shared-memory ring.
//
// Replaces, for a single query: server/search.go:214 (centroid scoring, compute/cosine.go:13-57), :220-223 (sort and
// nprobe cut), :241-273 (posting-list scan, running sort, one hit per document, truncate).  scan.cu's two-launch form
// of the same path spends most of a single query's time on fixed costs: two launches, load -> score -> load rounds with
// nothing in flight while a warp scores, and a last block that merges 296 partial lists.  Here
//
//   * every block (one per SM) keeps a ring of row-code stages in shared memory, filled by bulk copies
//     (cp.async.bulk, completion on an mbarrier).  Every warp owns its stage(s): it arms them at the start and re-arms a
//     stage with its next chunk the moment it has drained it into registers, so all of a block's rows are in flight
//     from the first cycle, no load waits for scoring and nothing waits for a single issuing thread.  The 24 bytes of
//     side data per row (header, integer sums, document id) are prefetched into registers one chunk ahead;
//   * the probe stage and the list stage share the launch: blocks score C/gridDim centroids each, meet at a grid
//     barrier (cooperative launch: all blocks are resident), and every block selects the same nprobe best centroids
//     from the shared key array (group maxima give a threshold, the few survivors are sorted);
//   * a block publishes only `pub` (>= k + 6) distinct documents, so the last block sorts ~pub candidates instead of
//     merging gridDim lists: the best k distinct documents of the union are among the best k of every part.
//
// Scores are the certified integer-identity scores of common.cuh.  A query whose emitted window contains a candidate
// whose float32 rounding could not be certified is NOT repaired here (rare: ~1e-4 of rows at cos ~ 0.1): its status
// word gets the same bits scan.cu uses and the caller's resolve path (literal arithmetic) finishes it.
#include <cstdlib>

#include "internal.h"
#include "ring.cuh"
#include "topk.cuh"

namespace vs {

namespace {

constexpr int kFusedWarps = 16;                       // 4 per scheduler (their latencies overlap); 512 threads leave 128
                                                      // registers per thread
constexpr int kFusedThreads = 32 * kFusedWarps;
constexpr int kFusedMaxStages = 2 * kFusedWarps;      // ring stages: stage w (and w + 16) belongs to warp w
constexpr int kFusedSmemBudget = 227 * 1024;

constexpr int kSegBufCap = 256;  // probe-stage survivors whose segment info is staged (>= kMaxSeg)

struct FusedCtl {
    uint64_t full[kFusedMaxStages];
    uint64_t seg_start[kMaxSeg];           // first store row of segment (probed list) s
    uint32_t seg_len[kMaxSeg];
    uint32_t seg_prefix[kMaxSeg + 1];      // rows before segment s in the query's concatenated row sequence
    uint32_t seg_a[kMaxSeg], seg_b[kMaxSeg];  // this block's rows of segment s: [a, b) within the segment
    uint32_t seg_cpre[kMaxSeg + 1];        // this block's chunks before segment s
    SideConst qside;
    unsigned long long t_scan_end, t_first;  // phase stamps (trace)
    uint32_t nchunks;       // chunks of this block
    int nseg;               // segments (probed lists) of the list stage
    uint32_t thr;           // selection threshold
    unsigned int cnt;       // survivors / collected candidates
    unsigned int overflow;
    unsigned int abort_status;  // non-zero: the probe stage could not be decided here; bits for the status word
    unsigned int is_last;
    int wcnt[kFusedWarps];      // candidates each warp kept
    alignas(16) uint32_t gmax[kFusedThreads];
    alignas(16) uint4 segbuf[kSegBufCap];  // (list start lo, hi, rows, uncertified) of the probe-stage survivors
};

#define FUSED_TRACE(slot)                                                                                   \
    do {                                                                                                    \
        if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 16 + (slot)] = fused_timer();          \
    } while (0)

// kreg[i] for a run-time i without spilling the array to local memory
__device__ __forceinline__ uint32_t kreg_at(const uint32_t (&kreg)[16], int i) {
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) v = (j == i) ? kreg[j] : v;
    return v;
}

// A key t such that at least m of the block's candidates have key >= t (0 when fewer than m groups hold a candidate):
// the m-th largest of the maxima of disjoint groups is a lower bound of the m-th largest candidate.  gm = the largest
// key among this thread's candidates.  Groups are 2..32 adjacent threads, at least 2m groups (so about 1.4 m candidates
// survive the cut).  All threads call; two barriers inside.
__device__ __forceinline__ uint32_t block_threshold(uint32_t gm, int m, FusedCtl &ctl) {
    int lg = 5;  // groups of gsz = 1 << lg lanes (shifts, not divisions: a division is ~50 dependent instructions)
    while (lg > 1 && (kFusedThreads >> lg) < 2 * m) lg--;
    const int gsz = 1 << lg;
    for (int o = 1; o < gsz; o <<= 1) gm = max(gm, __shfl_xor_sync(FULL, gm, o));
    const int ng = kFusedThreads >> lg;  // 16 .. 256 groups
    const int g = (int)threadIdx.x >> lg, part = (int)threadIdx.x & (gsz - 1);
    if (threadIdx.x == 0) ctl.thr = 0;
    if (part == 0) ctl.gmax[g] = gm;
    __syncthreads();
    // rank of every group maximum among all of them: the gsz lanes of a group share the comparisons
    const uint32_t mine = ctl.gmax[g];
    int r = 0;
    for (int u = part; u < ng; u += gsz) {
        const uint32_t o = ctl.gmax[u];
        r += (int)((o > mine) | ((o == mine) & (u < g)));
    }
    for (int o = 1; o < gsz; o <<= 1) r += __shfl_xor_sync(FULL, r, o);
    if (part == 0 && r == m - 1) ctl.thr = mine;
    __syncthreads();
    return ctl.thr;
}

// The candidates collected in a[0..ctl.cnt) -> best-first, one hit per document (dedup), in the returned buffer
// (at least `cap` entries, empty beyond what there is).  *slow: the cut cannot be trusted (the collection overflowed, or
// duplicates left fewer than m documents although candidates below the threshold were left out): the caller takes every
// candidate in turn instead.  All threads call.
__device__ __forceinline__ CandBuf finish_survivors(CandBuf a, CandBuf b, FusedCtl &ctl, int cap, int m, bool dedup,
                                                    uint32_t tkey, int *scan_tmp, bool *slow) {
    __syncthreads();
    const bool over = ctl.overflow != 0;
    const int n = (int)min(ctl.cnt, (unsigned)kOutCap);
    __syncthreads();
    *slow = over;
    if (over) return b;
    if (n <= kFusedThreads) {  // the usual case (about 1.4 m survivors): ordered by counting, one pass
        const int uniq = block_rank_small(a, n, b, cap, dedup);
        *slow = dedup && uniq < m && tkey != 0;
        return b;
    }
    block_sort_small(a, n, b, max(n, cap));
    if (!dedup) return b;
    const int uniq = block_unique_compact(b, n, a, cap, scan_tmp);
    *slow = uniq < m && tkey != 0;
    return a;
}

}  // namespace

// TR: rows per ring stage; SPW: stages per warp (1 or 2; the ring has 16 * SPW stages).
template <int G, int CPL, int KPL, int TR, int SPW>
__global__ void __launch_bounds__(kFusedThreads, 1)
fused_search_kernel(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char fsm[];
    constexpr int CAP = 32 * KPL;
    constexpr int WB = CAP + 32;  // candidates a warp can hold before it cuts its buffer to the best CAP
    constexpr int NG = 32 / G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = p.rows.d, d_pad = p.rows.d_pad;
    const uint32_t stage_bytes = (uint32_t)p.stage_bytes;
    constexpr uint32_t S = kFusedWarps * SPW;
    unsigned char *ring = fsm;
    // per-warp candidate buffers (outside the ring: they fill while the ring streams)
    uint64_t *cand_id = reinterpret_cast<uint64_t *>(fsm + (size_t)S * stage_bytes);
    uint32_t *cand_key = reinterpret_cast<uint32_t *>(cand_id + kFusedWarps * WB);
    uint32_t *cand_meta = cand_key + kFusedWarps * WB;
    FusedCtl &ctl = *reinterpret_cast<FusedCtl *>(cand_meta + kFusedWarps * WB);
    // the sort buffers overlay the ring: they are used only while no bulk copy is in flight
    SortSmem &ss = *reinterpret_cast<SortSmem *>(fsm);
    const CandBuf bufA{ss.key_a, ss.meta_a, ss.id_a};
    const CandBuf bufB{ss.key_b, ss.meta_b, ss.id_b};
    const uint32_t Gd = gridDim.x;
    const bool has_probe = p.npe > 0;
    const bool dedup = p.ids != nullptr;

    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < S; i++) mbar_init(smem_u32(&ctl.full[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const float2 h = p.query.hdr[0];
        const uint2 s = p.query.sums[0];
        ctl.qside = make_side(h.x, h.y, s.x, s.y, D);
        ctl.abort_status = 0;
        ctl.overflow = 0;
        ctl.cnt = 0;
        ctl.thr = 0;
        ctl.t_scan_end = 0;
        ctl.t_first = 0;
    }
    uint4 qreg[CPL];
#pragma unroll
    for (int j = 0; j < CPL; j++) qreg[j] = *reinterpret_cast<const uint4 *>(p.query.codes + ((lane % G) + G * j) * 16);

    // ================= probe stage: score my share of the centroid table, meet, select =================
    // A block's share of the table is small (28 rows of 4096 on 148 SMs, L2-resident): plain 128-bit loads issued by every
    // warp at once beat a ring here.  Tiles are sized so that every warp gets one, and a warp's first tile is loaded and
    // dotted BEFORE the block waits for thread 0's prologue (the query constants are only needed for the scoring).
    const uint64_t C = has_probe ? p.cent.n : 0;
    const uint32_t cq = (uint32_t)C / Gd, cr = (uint32_t)C % Gd;  // block b takes centroids [b*q + min(b,r), ...)
    const uint64_t c_lo = (uint64_t)blockIdx.x * cq + min(blockIdx.x, cr), c_hi = (uint64_t)(blockIdx.x + 1) * cq + min(blockIdx.x + 1, cr);
    uint32_t pchunk = (uint32_t)((c_hi - c_lo + kFusedWarps - 1) / kFusedWarps);
    pchunk = pchunk < (uint32_t)NG ? (uint32_t)NG : (pchunk > 32u ? 32u : pchunk);
    const uint32_t pnch = (uint32_t)((c_hi - c_lo + pchunk - 1) / pchunk);
    struct ProbeTile {
        uint64_t ci, l0, l1;
        float2 h;
        uint2 sums;
        uint32_t dot;
        bool valid;
    };
    auto probe_tile = [&](uint32_t c) {  // loads + integer dots of tile c; the finishing lane's side data beside them
        ProbeTile t;
        const uint64_t crow0 = c_lo + (uint64_t)c * pchunk;
        const int cn = (int)min((uint64_t)pchunk, c_hi - crow0);
        const int iters = (cn + NG - 1) / NG;
        const int myr = (lane / G) * iters + (lane % G);
        t.valid = (lane % G) < iters && myr < cn;
        t.ci = crow0 + (uint64_t)(t.valid ? myr : 0);
        t.h = make_float2(0.f, 0.f);
        t.sums = make_uint2(0, 0);
        t.l0 = t.l1 = 0;
        if (t.valid) {
            t.h = p.cent.hdr[t.ci];
            t.sums = p.cent.sums[t.ci];
            t.l0 = __ldg(p.list_off + t.ci);
            t.l1 = t.l0 + __ldg(p.list_len + t.ci);
        }
        t.dot = tile_dots<G, CPL>(p.cent.codes, (size_t)crow0, cn, d_pad, qreg, lane, iters);
        return t;
    };
    ProbeTile first_tile;
    first_tile.valid = false;
    if (has_probe && (uint32_t)warp < pnch) first_tile = probe_tile((uint32_t)warp);
    __syncthreads();
    const SideConst xq = ctl.qside;
    FUSED_TRACE(0);
    const long long clk0 = clock64();

    if (has_probe) {
        auto probe_score = [&](const ProbeTile &t) {
            if (!t.valid) return;
            bool flag;
            const float sim = score_fast(xq, t.h.x, t.h.y, t.sums.x, t.sums.y, t.dot, D, &flag);
            p.keys[t.ci] = f32_to_key(sim);
            p.seginfo[t.ci] = make_uint4((uint32_t)t.l0, (uint32_t)(t.l0 >> 32), (uint32_t)(t.l1 - t.l0), flag ? 1u : 0u);
        };
        probe_score(first_tile);
        for (uint32_t c = warp + kFusedWarps; c < pnch; c += kFusedWarps) probe_score(probe_tile(c));
        FUSED_TRACE(1);
        // ---- grid barrier (cooperative launch: every block is resident) ----
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.sync) : "memory");
            while (ld_acquire_u32(&p.sync[0]) < Gd) {
            }
        }
        __syncthreads();
        FUSED_TRACE(2);
        // ---- every block selects the same npe best centroids (search.go:220-223; ties: lower index first) ----
        const uint32_t P = (uint32_t)(((C + kFusedThreads - 1) / kFusedThreads + 3) & ~uint64_t(3));
        uint4 seg_first = make_uint4(0, 0, 0, 0);
        int seg_first_pos = -1;
        const uint64_t k_lo = (uint64_t)threadIdx.x * P;
        // survivors are staged with their list extent; a thread reserves room for all of its own with ONE atomic
        auto keep_all = [&](const uint32_t *kk, int cnt, uint32_t thr, uint64_t first) {
            unsigned mask = 0;
            for (int i = 0; i < cnt; i++) mask |= (unsigned)((kk[i] != 0) & (kk[i] >= thr)) << i;
            if (!mask) return;
            unsigned int pos = atomicAdd(&ctl.cnt, (unsigned)__popc(mask));
            while (mask) {
                const int i = __ffs(mask) - 1;
                mask &= mask - 1;
                if (pos < (unsigned)kSegBufCap) {
                    cand_put(bufA, (int)pos, kk[i], pos, first + i);
                    ctl.segbuf[pos] = __ldcg(p.seginfo + first + i);
                } else {
                    ctl.overflow = 1;
                }
                pos++;
            }
        };
        if (P <= 16) {  // (<= 8192 centroids) keys stay in registers between the two passes
            uint32_t kreg[16];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if ((uint32_t)i < P && k_lo + i < C) v = __ldcg(reinterpret_cast<const uint4 *>(p.keys + k_lo + i));
                kreg[i] = v.x;
                kreg[i + 1] = k_lo + i + 1 < C ? v.y : 0u;
                kreg[i + 2] = k_lo + i + 2 < C ? v.z : 0u;
                kreg[i + 3] = k_lo + i + 3 < C ? v.w : 0u;
            }
            uint32_t gm = 0;
#pragma unroll
            for (int i = 0; i < 16; i++) gm = max(gm, kreg[i]);
            const uint32_t thr = block_threshold(gm, p.npe, ctl);
            unsigned mask = 0;
#pragma unroll
            for (int i = 0; i < 16; i++) mask |= (unsigned)((kreg[i] != 0) & (kreg[i] >= thr)) << i;
            if (mask) {
                unsigned int pos = atomicAdd(&ctl.cnt, (unsigned)__popc(mask));
                // the first survivor's list extent stays in flight while the survivors are ranked (staged afterwards)
                const int i0 = __ffs(mask) - 1;
                if (pos < (unsigned)kSegBufCap) {
                    cand_put(bufA, (int)pos, kreg_at(kreg, i0), pos, k_lo + i0);
                    seg_first = __ldcg(p.seginfo + k_lo + i0);
                    seg_first_pos = (int)pos;
                } else {
                    ctl.overflow = 1;
                }
                pos++;
                mask &= mask - 1;
                while (mask) {  // (rare: two survivors among one thread's keys)
                    const int i = __ffs(mask) - 1;
                    mask &= mask - 1;
                    if (pos < (unsigned)kSegBufCap) {
                        cand_put(bufA, (int)pos, kreg_at(kreg, i), pos, k_lo + i);
                        ctl.segbuf[pos] = __ldcg(p.seginfo + k_lo + i);
                    } else {
                        ctl.overflow = 1;
                    }
                    pos++;
                }
            }
        } else {
            uint32_t gm = 0;
            for (uint32_t i = 0; i < P && k_lo + i < C; i += 4) {
                const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p.keys + k_lo + i));
                gm = max(gm, v.x);
                if (k_lo + i + 1 < C) gm = max(gm, v.y);
                if (k_lo + i + 2 < C) gm = max(gm, v.z);
                if (k_lo + i + 3 < C) gm = max(gm, v.w);
            }
            const uint32_t thr = block_threshold(gm, p.npe, ctl);
            for (uint32_t i = 0; i < P && k_lo + i < C; i += 4) {
                const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p.keys + k_lo + i));
                const uint32_t kk[4] = {v.x, v.y, v.z, v.w};
                const int cnt = (int)min((uint64_t)4, C - (k_lo + i));
                keep_all(kk, cnt, thr, k_lo + i);
            }
        }
        __syncthreads();
        const int nsurv = (int)min(ctl.cnt, (unsigned)kSegBufCap);
        const bool sel_overflow = ctl.overflow != 0;
        __syncthreads();
        block_rank_small(bufA, nsurv, bufB, nsurv, false);  // (nsurv <= kSegBufCap <= blockDim.x)
        if (seg_first_pos >= 0) ctl.segbuf[seg_first_pos] = seg_first;  // (its load was in flight during the ranking)
        __syncthreads();
        const int npe = min(p.npe, nsurv);  // (= p.npe: every centroid has a non-zero key and the caller keeps npe < C)
        if (threadIdx.x == 0) {
            ctl.nseg = npe;
            ctl.cnt = 0;  // the block selection after the list stage counts from zero again
            ctl.overflow = 0;
            if (sel_overflow) ctl.abort_status = kStatusProbeAmbiguous;  // more ties than the staging holds
        }
        if ((int)threadIdx.x < npe) {
            const uint4 sg = ctl.segbuf[bufB.meta[threadIdx.x]];
            ctl.seg_start[threadIdx.x] = (uint64_t)sg.x | ((uint64_t)sg.y << 32);
            ctl.seg_len[threadIdx.x] = sg.z;
            // an uncertified similarity inside the window: the caller's literal path decides this query
            if (sg.w) ctl.abort_status = kStatusProbeAmbiguous;
            if (p.out_probe && blockIdx.x == 0) p.out_probe[threadIdx.x] = (uint32_t)bufB.id[threadIdx.x];
        }
        __syncthreads();
    } else if (threadIdx.x == 0) {
        ctl.seg_start[0] = p.flat_start;
        ctl.seg_len[0] = (uint32_t)p.flat_count;
        ctl.nseg = 1;
    }
    __syncthreads();

    // ================= list stage: my share of the probed rows through the ring =================
    // This block's rows [r_lo, r_hi) of the query's concatenated segments are cut into chunks of <= 16 rows that do not
    // cross a segment.  Warp w owns the ring stages w and w + 16 and the chunks w, w + 16, w + 32, ... (chunk c in stage
    // c % 32): its lane 0 arms both stages at the start and re-arms a stage with the chunk two turns ahead as soon as the
    // warp has drained it into registers, before the scoring -- so one of a warp's stages is in flight while it works on
    // the other.  There is no producer warp and no hand-off between warps: nothing waits for a single issuing thread (one
    // thread needs ~0.45 us per chunk, which would pace the whole scan), and a stage's barrier is only ever waited on by the
    // warp that armed it, in order.
    // A warp appends every row that reaches its threshold to its own candidate buffer.  The threshold is 0 until the buffer
    // fills; then the buffer is cut to its best CAP distinct documents (sorted insertion, one hit per document) and the
    // CAP-th of them becomes the threshold.  A short scan (one query of BASELINE config 2 gives a warp ~35 rows) never
    // pays for an insertion; a long one cuts a few times (the rate of survivors falls like CAP / rows).
    const bool aborted = ctl.abort_status != 0;
    int cnt_w = 0;
    if (!aborted) {
        const int ns = ctl.nseg;
        if (warp == 0) {  // rows before each segment
            uint32_t carry = 0;
            for (int base = 0; base < ns; base += 32) {
                const int s = base + lane;
                uint32_t v = s < ns ? ctl.seg_len[s] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, v, o);
                    if (lane >= o) v += t;
                }
                if (s < ns) ctl.seg_prefix[s + 1] = carry + v;
                carry += __shfl_sync(FULL, v, 31);
            }
            if (lane == 0) ctl.seg_prefix[0] = 0;
            __syncwarp();
            const uint32_t Trows = ctl.seg_prefix[ns];
            const uint32_t rq = Trows / Gd, rr = Trows % Gd;  // block b takes rows [b*q + min(b,r), (b+1)*q + min(b+1,r))
            const uint64_t r_lo = (uint64_t)blockIdx.x * rq + min(blockIdx.x, rr), r_hi = (uint64_t)(blockIdx.x + 1) * rq + min(blockIdx.x + 1, rr);
            carry = 0;
            for (int base = 0; base < ns; base += 32) {  // my rows and chunks of each segment
                const int s = base + lane;
                uint32_t v = 0;
                if (s < ns) {
                    const uint64_t ps = ctl.seg_prefix[s], pe = ctl.seg_prefix[s + 1];
                    uint32_t a = 0, b = 0;
                    if (pe > r_lo && ps < r_hi) {
                        a = (uint32_t)((r_lo > ps ? r_lo : ps) - ps);
                        b = (uint32_t)((r_hi < pe ? r_hi : pe) - ps);
                    }
                    ctl.seg_a[s] = a;
                    ctl.seg_b[s] = b;
                    v = (b - a + TR - 1) / TR;
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, v, o);
                    if (lane >= o) v += t;
                }
                if (s < ns) ctl.seg_cpre[s + 1] = carry + v;
                carry += __shfl_sync(FULL, v, 31);
            }
            if (lane == 0) {
                ctl.seg_cpre[0] = 0;
                ctl.nchunks = carry;
            }
        }
        __syncthreads();
        FUSED_TRACE(3);
        const uint32_t nch = ctl.nchunks;
        // chunk c of this block: first store row and rows (every lane computes it: warp-uniform)
        auto locate = [&](uint32_t c, uint32_t &row0, uint32_t &nr) {
            int lo = 0, hi = ns;  // segment with seg_cpre[s] <= c < seg_cpre[s + 1]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (ctl.seg_cpre[mid] <= c) lo = mid;
                else hi = mid;
            }
            const uint32_t o = ctl.seg_a[lo] + (c - ctl.seg_cpre[lo]) * (uint32_t)TR;
            nr = min((uint32_t)TR, ctl.seg_b[lo] - o);
            row0 = (uint32_t)(ctl.seg_start[lo] + o);
        };
        // what a warp holds about a chunk it has armed: where it is and, per lane, the side data of the row that lane
        // will finish (header, integer sums, document id: 24 bytes per row, fetched while the codes are in flight)
        struct Pending {
            uint32_t row0, nr;
            float2 h;
            uint2 sums;
            uint64_t id;
        };
        auto arm = [&](uint32_t c, uint32_t slot, Pending &pd) {
            locate(c, pd.row0, pd.nr);
            if (lane == 0) {
                const uint32_t full = smem_u32(&ctl.full[warp + kFusedWarps * slot]);
                const uint32_t cb = pd.nr * (uint32_t)d_pad;
                mbar_expect_tx(full, cb);
                bulk_g2s(smem_u32(ring + (size_t)(warp + kFusedWarps * slot) * stage_bytes), p.rows.codes + (uint64_t)pd.row0 * d_pad, cb,
                         full);
            }
            const int iters = ((int)pd.nr + NG - 1) / NG;
            const int myr = (lane / G) * iters + (lane % G);
            pd.h = make_float2(0.f, 0.f);
            pd.sums = make_uint2(0, 0);
            pd.id = kEmptyId;
            if ((lane % G) < iters && myr < (int)pd.nr) {
                const uint32_t row = pd.row0 + (uint32_t)myr;
                pd.h = p.rows.hdr[row];
                pd.sums = p.rows.sums[row];
                pd.id = p.ids ? p.ids[row] : p.id_base + row;
            }
        };
        Pending pend[SPW];
        if (lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // (the ring served as sort buffers)
#pragma unroll
        for (int sl = 0; sl < SPW; sl++) {
            pend[sl].nr = 0;
            if ((uint32_t)warp + kFusedWarps * sl < nch) arm((uint32_t)warp + kFusedWarps * sl, sl, pend[sl]);
        }
        uint32_t thr_w = 0;
        uint64_t *wid = cand_id + warp * WB;
        uint32_t *wkey = cand_key + warp * WB, *wmeta = cand_meta + warp * WB;
        for (uint32_t j0 = 0; (uint32_t)warp + kFusedWarps * j0 < nch; j0 += SPW) {
#pragma unroll
            for (int sl = 0; sl < SPW; sl++) {
                const uint32_t c = (uint32_t)warp + kFusedWarps * (j0 + sl);
                if (c >= nch) break;
                const Pending cur = pend[sl];
                const uint32_t st = (uint32_t)warp + kFusedWarps * sl;
                mbar_wait(smem_u32(&ctl.full[st]), (j0 / SPW) & 1u);
                if (p.trace && c == 0 && lane == 0) ctl.t_first = fused_timer();
                const int iters = ((int)cur.nr + NG - 1) / NG;
                const uint32_t mydot = stage_dots<G, CPL>(smem_u32(ring + (size_t)st * stage_bytes), (int)cur.nr, d_pad, qreg, lane, iters);
                const int myr = (lane / G) * iters + (lane % G);
                const bool valid = (lane % G) < iters && myr < (int)cur.nr;
                __syncwarp();
                // the stage is drained into registers: re-arm it with this warp's chunk SPW turns ahead, before the scoring
                if (c + kFusedWarps * SPW < nch) arm(c + kFusedWarps * SPW, sl, pend[sl]);
                uint32_t key = 0, meta = 0;
                if (valid) {
                    bool flag;
                    const float sim = score_fast(xq, cur.h.x, cur.h.y, cur.sums.x, cur.sums.y, mydot, D, &flag);
                    key = f32_to_key(sim);
                    meta = (cur.row0 + (uint32_t)myr) | (flag ? kFlagBit : 0u);
                }
                unsigned m = __ballot_sync(FULL, valid && key >= thr_w);
                if (cnt_w + __popc(m) > WB) {
                    // cut the buffer to its best CAP distinct documents
                    WarpTopK<KPL> top;
                    top.init();
                    for (int base = 0; base < cnt_w; base += 32) {
                        const int e = base + lane;
                        const bool have = e < cnt_w;
                        top.offer(have, have ? wkey[e] : 0u, have ? wmeta[e] : 0u, have ? wid[e] : kEmptyId, lane, dedup);
                    }
                    __syncwarp();
#pragma unroll
                    for (int s2 = 0; s2 < KPL; s2++) {
                        wkey[s2 * 32 + lane] = top.skey[s2];
                        wmeta[s2 * 32 + lane] = top.meta[s2];
                        wid[s2 * 32 + lane] = top.id[s2];
                    }
                    cnt_w = top.count();
                    thr_w = top.thr_key;
                    __syncwarp();
                    m = __ballot_sync(FULL, valid && key >= thr_w);
                }
                if (m & (1u << lane)) {
                    const int slot = cnt_w + __popc(m & ((1u << lane) - 1u));
                    wkey[slot] = key;
                    wmeta[slot] = meta;
                    wid[slot] = cur.id;
                }
                cnt_w += __popc(m);
            }
        }
        if (lane == 0) {
            ctl.wcnt[warp] = cnt_w;
            if (p.trace) atomicMax(&ctl.t_scan_end, fused_timer());
        }
    }
    FUSED_TRACE(4);
    __syncthreads();  // every bulk copy of this block has landed and been consumed: the ring is free for the sort buffers
    if (p.trace && threadIdx.x == 0) {
        p.trace[(size_t)blockIdx.x * 16 + 9] = ctl.t_scan_end;
        p.trace[(size_t)blockIdx.x * 16 + 11] = ctl.t_first;
    }

    // ================= block selection -> publish `pub` distinct documents -> last block finishes the query =================
    const int pub = p.pub;
    // all candidates of the block, warp after warp (the unused tail of a buffer reads as empty)
    auto cand_at = [&](int e, uint32_t &k, uint32_t &m, uint64_t &id) {
        const int w = e / WB, i = e - w * WB;
        const bool have = i < ctl.wcnt[w];
        k = have ? cand_key[e] : 0u;
        m = have ? cand_meta[e] : 0u;
        id = have ? cand_id[e] : kEmptyId;
    };
    auto insert_all = [&](auto &&get, int n, const CandBuf &dst) {  // exact but slow: warp 0 takes every candidate in turn
        __syncthreads();
        if (warp == 0) {
            WarpTopK<KPL> top;
            top.init();
            for (int base = 0; base < n; base += 32) {
                const int e = base + lane;
                uint32_t k = 0, m = 0;
                uint64_t id = kEmptyId;
                if (e < n) get(e, k, m, id);
                top.offer(k != 0, k, m, id, lane, dedup);
            }
            top.store_soa(dst, 0, lane);
        }
        __syncthreads();
    };
    if (!aborted) {
        constexpr int NS = kFusedWarps * WB;
        uint32_t gm = 0;
        for (int e = threadIdx.x; e < NS; e += kFusedThreads) {
            uint32_t k, m;
            uint64_t id;
            cand_at(e, k, m, id);
            gm = max(gm, k);
        }
        const uint32_t thr = block_threshold(gm, pub, ctl);
        for (int e = threadIdx.x; e < NS; e += kFusedThreads) {
            uint32_t k, m;
            uint64_t id;
            cand_at(e, k, m, id);
            if (k != 0 && k >= thr) {
                const unsigned int pos = atomicAdd(&ctl.cnt, 1u);
                if (pos < (unsigned)kOutCap) cand_put(bufA, (int)pos, k, m, id);
                else ctl.overflow = 1;
            }
        }
        bool slow;
        CandBuf res = finish_survivors(bufA, bufB, ctl, CAP, pub, dedup, thr, ss.scan_tmp, &slow);
        if (slow) {
            insert_all(cand_at, NS, bufB);
            res = bufB;
        }
        if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 16 + 12] = fused_timer();
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
    if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 16 + 13] = (unsigned long long)(clock64() - clk0);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(&p.sync[1], 1u);
        ctl.is_last = (tk == Gd - 1) ? 1u : 0u;
        ctl.cnt = 0;
        ctl.overflow = 0;
    }
    __syncthreads();
    if (!ctl.is_last) return;
    __threadfence();
    if (threadIdx.x == 0) {  // re-arm for the next launch: every block has passed the barrier and taken its ticket
        p.sync[0] = 0;
        p.sync[1] = 0;
    }
    FUSED_TRACE(6);
    if (aborted) {
        if (threadIdx.x == 0) {
            p.out_counts[0] = 0;
            p.out_status[0] = ctl.abort_status;
        }
        return;
    }
    // ---- all published entries (gridDim x pub): threshold, collect, sort, one hit per document ----
    const int total = (int)Gd * pub;
    auto part_at = [&](int e, uint32_t &k, uint32_t &m, uint64_t &id) {
        const uint4 v = __ldcg(p.partial + e);
        k = v.x;
        m = v.y;
        id = (uint64_t)v.z | ((uint64_t)v.w << 32);
    };
    constexpr int kPer = 4;
    uint32_t tkey = 0;
    if (total <= kPer * kFusedThreads) {  // one round trip: the entries stay in registers between the two passes
        uint4 ent[kPer];
        uint32_t gm = 0;
#pragma unroll
        for (int u = 0; u < kPer; u++) {
            const int e = (int)threadIdx.x + u * kFusedThreads;
            ent[u] = e < total ? __ldcg(p.partial + e) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kPer; u++) gm = max(gm, ent[u].x);
        tkey = block_threshold(gm, p.k, ctl);
#pragma unroll
        for (int u = 0; u < kPer; u++) {
            if (ent[u].x != 0 && ent[u].x >= tkey) {
                const unsigned int pos = atomicAdd(&ctl.cnt, 1u);
                if (pos < (unsigned)kOutCap) cand_put(bufA, (int)pos, ent[u].x, ent[u].y, (uint64_t)ent[u].z | ((uint64_t)ent[u].w << 32));
                else ctl.overflow = 1;
            }
        }
    } else {
        uint32_t gm = 0;
        for (int e = threadIdx.x; e < total; e += kFusedThreads) gm = max(gm, __ldcg(&p.partial[e].x));
        tkey = block_threshold(gm, p.k, ctl);
        for (int e = threadIdx.x; e < total; e += kFusedThreads) {
            uint32_t k, m;
            uint64_t id;
            part_at(e, k, m, id);
            if (k != 0 && k >= tkey) {
                const unsigned int pos = atomicAdd(&ctl.cnt, 1u);
                if (pos < (unsigned)kOutCap) cand_put(bufA, (int)pos, k, m, id);
                else ctl.overflow = 1;
            }
        }
    }
    bool slow;
    CandBuf fin = finish_survivors(bufA, bufB, ctl, CAP, p.k, dedup, tkey, ss.scan_tmp, &slow);
    if (slow) {
        insert_all(part_at, total, bufB);
        fin = bufB;
    }
    FUSED_TRACE(7);
    // ---- emit the first k (search.go:270); an uncertified score among them goes to the caller's literal path ----
    if (warp == 0) {
        uint32_t status = 0;
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
// Ring geometry: 16 * spw stages of tr rows (codes only) beside the candidate buffers.  spw = 2 halves the stage.
static int fused_spw() {
    static const int v = getenv("VS_FUSED_SPW") ? atoi(getenv("VS_FUSED_SPW")) : 1;
    return v == 2 ? 2 : 1;
}
static int fused_tile_rows(int d_pad, int kpl, int spw) {
    const int tr = (d_pad <= 768 && kpl == 1) ? 16 : 8;
    return tr / spw;
}
static bool fused_geometry(int d_pad, int kpl, int spw, int *stage_bytes, size_t *smem) {
    const int tr = fused_tile_rows(d_pad, kpl, spw);
    const int sb = (tr * d_pad + 127) & ~127;
    const int wb = 32 * kpl + 32;
    const size_t fixed = (size_t)kFusedWarps * wb * 16 + ((sizeof(FusedCtl) + 127) & ~size_t(127));
    *stage_bytes = sb;
    *smem = (size_t)kFusedWarps * spw * sb + fixed;
    // the sort buffers overlay the ring
    return *smem <= (size_t)kFusedSmemBudget && (size_t)kFusedWarps * spw * sb >= sizeof(SortSmem);
}

bool fused_supported(int d_pad, int kpl) {
    if (d_pad & 15) return false;
    switch (d_pad >> 4) {
        case 48: case 32: case 64: case 96: case 24: break;
        default: return false;
    }
    int sb;
    size_t smem;
    return (kpl == 1 || kpl == 2 || kpl == 4) && fused_geometry(d_pad, kpl, fused_spw(), &sb, &smem);
}

// Distinct documents a block publishes: k is enough for the exact top k (the best k documents of a union are among the
// best k of every part); rounded up to whole 64-byte lines of 16-byte entries.
int fused_pub(int k, int kpl) {
    const int cap = 32 * kpl;
    const int pub = ((k + 3) / 4) * 4;
    return pub < cap ? pub : cap;
}

template <int G, int CPL, int KPL, int TR, int SPW>
static cudaError_t launch_fused_t(FusedParams p, int grid, cudaStream_t st) {
    size_t smem;
    if (!fused_geometry(p.rows.d_pad, KPL, SPW, &p.stage_bytes, &smem)) return cudaErrorInvalidValue;
    p.stages = kFusedWarps * SPW;
    auto kern = fused_search_kernel<G, CPL, KPL, TR, SPW>;
    static int max_grid_of[64] = {0};  // per instantiation and device (function attributes are per device): blocks a
    int dev = 0;                        // cooperative launch can hold
    cudaGetDevice(&dev);
    int &max_grid = max_grid_of[dev & 63];
    if (!max_grid) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFusedThreads, smem);
        if (e != cudaSuccess) return e;
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
    // Cooperative: the grid barrier needs every block resident, and only a cooperative launch guarantees that two such
    // kernels on different streams never hold half of the SMs each.  (VS_FUSED_COOP=0: measurement aid on an idle GPU.)
    static const int coop = getenv("VS_FUSED_COOP") ? atoi(getenv("VS_FUSED_COOP")) : 1;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = coop;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int G, int CPL, int KPL, int TR>
static cudaError_t launch_fused_s(const FusedParams &p, int grid, cudaStream_t st) {
    return fused_spw() == 2 ? launch_fused_t<G, CPL, KPL, TR / 2, 2>(p, grid, st) : launch_fused_t<G, CPL, KPL, TR, 1>(p, grid, st);
}

template <int KPL>
static cudaError_t launch_fused_k(const FusedParams &p, int grid, cudaStream_t st) {
    switch (p.rows.d_pad >> 4) {
        case 48: return launch_fused_s<16, 3, KPL, (KPL == 1 ? 16 : 8)>(p, grid, st);  // 768-d (nomic-embed-text)
        case 32: return launch_fused_s<32, 1, KPL, (KPL == 1 ? 16 : 8)>(p, grid, st);  // 512-d (noop/ai.go)
        case 64: return launch_fused_s<32, 2, KPL, 8>(p, grid, st);                    // 1024-d
        case 96: return launch_fused_s<32, 3, KPL, 8>(p, grid, st);                    // 1536-d
        case 24: return launch_fused_s<8, 3, KPL, (KPL == 1 ? 16 : 8)>(p, grid, st);   // 384-d
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
