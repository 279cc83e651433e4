// listmajor.cu -- the posting-list scan of a query BATCH, list-major: every probed list is read from HBM once and
// scored against all the queries of the batch that probe it.
//
// Replaces, for a batch: server/search.go:241-273 (posting-list scan, running sort, one hit per document, truncate)
// with compute/cosine.go:13-57 inside.  The reference runs one search per goroutine and reads the probed lists of every
// query separately (search.go:241-243); scan.cu's query-major kernel does the same on the device, and at 256 queries x
// 32 probes over 4096 lists ~57 % of its HBM traffic is lists read again for another query (8192 (query, list) pairs hit
// ~3540 distinct lists).  Here
//
//   1. the (query, list) pairs are inverted on the device into ITEMS: (list, up to 1024 of its rows, the queries that
//      probe it) -- lm_count / lm_items / lm_fill;
//   2. persistent blocks (one per SM) claim items in order.  An item's rows stream through a shared-memory ring of
//      32-row stages filled by bulk copies (cp.async.bulk, completion on an mbarrier).  With m queries on the item the
//      16 warps form r = 16 / m phase groups of m warps: the group of phase p takes the chunks p, p + r, p + 2r, ... and
//      owns its share of the ring stages; inside a group warp j scores the staged rows against query j (its codes live
//      in the warp's registers), so a row is read from HBM once and from shared memory m times.  The last warp of a
//      group to drain a stage re-arms it with the group's next chunk (no producer warp);
//   3. a warp keeps the rows that reach its threshold in its own candidate buffer (the same scheme as fused.cu); the
//      threshold starts at the query's RUNNING k-th best -- a per-query word in global memory raised (atomicMax) by every
//      warp that has found k distinct documents -- so after a query's first items almost nothing is kept;
//   4. what a warp kept is appended to the query's candidate list in global memory; lm_final picks the best k distinct
//      documents per query (threshold from group maxima, counting rank), emits them and raises the status bits.
//
// Exactness: thresholds only ever drop rows whose score is below a proven lower bound of the query's k-th best distinct
// document, so the emitted top k is the one the query-major path returns; scores are the certified integer-identity
// scores of common.cuh, and a query with an uncertified score inside its top k (or whose candidate list overflowed) gets
// the status bit that sends it to the caller's literal-arithmetic path, as everywhere else.
#include <cstdlib>

#include "internal.h"
#include "ring.cuh"
#include "topk.cuh"

namespace vs {

namespace {

constexpr int kLmWarps = 16;
constexpr int kLmThreads = 32 * kLmWarps;
constexpr int kLmStages = 8;        // ring stages of 32 rows (all 32 lanes finish a row's score; 16-row stages leave half idle)
// Rows per work item at most: a longer list is cut into equal parts.  An item costs a fixed ~3 us (the ring drains and
// refills, the warps load their queries), so fewer, longer items scan faster until the tail of the launch (one item per SM
// at the end) outweighs it.  VS_LM_SUB_ROWS overrides for experiments.
static uint32_t lm_sub_rows() {
    static const uint32_t v = [] {
        const char *e = getenv("VS_LM_SUB_ROWS");
        const long x = e ? atol(e) : 0;
        return (uint32_t)(x >= 64 && x <= (1 << 20) ? x : 1024);
    }();
    return v;
}
// items of a list of `len` rows: n parts of `per` rows (a multiple of 32) each, the last one shorter
__device__ __forceinline__ void lm_split(uint64_t len, uint32_t sub, uint32_t *n, uint32_t *per) {
    const uint32_t cnt = (uint32_t)((len + sub - 1) / sub);
    *n = cnt;
    *per = cnt ? (uint32_t)((((len + cnt - 1) / cnt) + 31) & ~(uint64_t)31) : 0u;
}
constexpr int kLmCap = 32;         // distinct documents a warp's buffer is cut to (k <= 32 on this path)
constexpr int kLmWB = kLmCap + 32; // candidates a warp can hold

// ---- 1. inversion of the probe lists -------------------------------------------------------------------------------
__global__ void lm_count_kernel(const uint32_t *__restrict__ probe, uint32_t npairs, uint32_t *__restrict__ count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npairs) atomicAdd(&count[probe[i]], 1u);
}

// One block: per list with at least one query, its items (<= sub rows each) and the offset of its query list.
__global__ void __launch_bounds__(1024) lm_items_kernel(const uint32_t *__restrict__ count, const uint64_t *__restrict__ list_off,
                                                       const uint64_t *__restrict__ list_len,
                                                       uint32_t C, uint32_t *__restrict__ pair_off, LmItem *__restrict__ items,
                                                       uint32_t *__restrict__ nitems, uint32_t items_cap, uint32_t sub) {
    __shared__ uint32_t s_items[64], s_pairs[64];
    const uint32_t per = (C + 1023u) / 1024u;
    const uint32_t lo = threadIdx.x * per, hi = min(C, lo + per);
    uint32_t my_items = 0, my_pairs = 0;
#pragma unroll 4
    for (uint32_t L = lo; L < hi; L++) {  // (both loads unconditional: independent, so they fly together)
        const uint32_t c = count[L];
        const uint64_t len = list_len[L];
        my_items += c ? (uint32_t)((len + sub - 1) / sub) : 0u;
        my_pairs += c;
    }
    // exclusive scans over the 1024 threads: shuffles inside a warp, the 32 warp totals by warp 0
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t vi = my_items, vp = my_pairs;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(FULL, vi, o), b = __shfl_up_sync(FULL, vp, o);
        if (lane >= o) {
            vi += a;
            vp += b;
        }
    }
    if (lane == 31) {
        s_items[warp] = vi;
        s_pairs[warp] = vp;
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t wi = s_items[lane], wp = s_pairs[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(FULL, wi, o), b = __shfl_up_sync(FULL, wp, o);
            if (lane >= o) {
                wi += a;
                wp += b;
            }
        }
        s_items[32 + lane] = wi;  // inclusive totals of warps 0..lane
        s_pairs[32 + lane] = wp;
    }
    __syncthreads();
    uint32_t it = vi - my_items + (warp ? s_items[32 + warp - 1] : 0u), pr = vp - my_pairs + (warp ? s_pairs[32 + warp - 1] : 0u);
    for (uint32_t L = lo; L < hi; L++) {
        const uint32_t c = count[L];
        pair_off[L] = pr;
        if (c) {
            const uint64_t st = list_off[L], len = list_len[L];
            uint32_t cnt, part;
            lm_split(len, sub, &cnt, &part);
            for (uint64_t o = 0; o < len; o += part) {
                if (it < items_cap) items[it] = LmItem{(uint32_t)(st + o), (uint32_t)min((uint64_t)part, len - o), pr, c};
                it++;
            }
            pr += c;
        }
    }
    if (threadIdx.x == 1023) *nitems = min(s_items[32 + 31], items_cap);
}

// The whole inversion in ONE single-block launch when the per-list counters fit in shared memory (C <= kLmFusedMaxC):
// histogram of the probe lists, items and pair offsets by the same scans as lm_items_kernel, the fill, the per-query
// side constants and the zeroing of the per-step counters -- instead of four memsets and four launches on the critical
// path of every step (each a few microseconds that no longer shrink when the store is sharded).
constexpr uint32_t kLmFusedMaxC = 12288;  // 2 x 4 B per list of dynamic shared memory
__global__ void __launch_bounds__(1024) lm_prepare_fused_kernel(const uint32_t *__restrict__ probe, uint32_t nq, uint32_t npe,
                                                               const uint64_t *__restrict__ list_off,
                                                               const uint64_t *__restrict__ list_len, uint32_t C, LmParams p,
                                                               uint32_t items_cap, uint32_t sub) {
    extern __shared__ uint32_t lm_sm[];
    uint32_t *hist = lm_sm, *poff = lm_sm + C;
    __shared__ uint32_t s_items[64], s_pairs[64];
    const uint32_t npairs = nq * npe;
    for (uint32_t L = threadIdx.x; L < C; L += 1024) hist[L] = 0;
    for (uint32_t q = threadIdx.x; q < nq; q += 1024) {
        p.gcnt[q] = 0;  // (gthr belongs to the seed kernel, which may run beside this one)
        const float2 h = p.queries.hdr[q];
        const uint2 sm = p.queries.sums[q];
        p.sides[q] = make_side(h.x, h.y, sm.x, sm.y, p.queries.d);
    }
    if (threadIdx.x == 0) *p.next_item = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < npairs; i += 1024) atomicAdd(&hist[probe[i]], 1u);
    __syncthreads();
    const uint32_t per = (C + 1023u) / 1024u;
    const uint32_t lo = threadIdx.x * per, hi = min(C, lo + per);
    uint32_t my_items = 0, my_pairs = 0;
#pragma unroll 4
    for (uint32_t L = lo; L < hi; L++) {
        const uint32_t c = hist[L];
        const uint64_t len = list_len[L];
        my_items += c ? (uint32_t)((len + sub - 1) / sub) : 0u;
        my_pairs += c;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t vi = my_items, vp = my_pairs;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(FULL, vi, o), b = __shfl_up_sync(FULL, vp, o);
        if (lane >= o) {
            vi += a;
            vp += b;
        }
    }
    if (lane == 31) {
        s_items[warp] = vi;
        s_pairs[warp] = vp;
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t wi = s_items[lane], wp = s_pairs[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(FULL, wi, o), b = __shfl_up_sync(FULL, wp, o);
            if (lane >= o) {
                wi += a;
                wp += b;
            }
        }
        s_items[32 + lane] = wi;
        s_pairs[32 + lane] = wp;
    }
    __syncthreads();
    uint32_t it = vi - my_items + (warp ? s_items[32 + warp - 1] : 0u), pr = vp - my_pairs + (warp ? s_pairs[32 + warp - 1] : 0u);
#pragma unroll 4
    for (uint32_t L = lo; L < hi; L++) {
        const uint32_t c = hist[L];
        const uint64_t st = list_off[L], len = list_len[L];
        poff[L] = pr;
        if (c) {
            uint32_t cnt, part;
            lm_split(len, sub, &cnt, &part);
            for (uint64_t o = 0; o < len; o += part) {
                if (it < items_cap) p.items[it] = LmItem{(uint32_t)(st + o), (uint32_t)min((uint64_t)part, len - o), pr, c};
                it++;
            }
            pr += c;
        }
    }
    if (threadIdx.x == 1023) *p.nitems = min(s_items[32 + 31], items_cap);
    __syncthreads();
    // the fill: hist[L] is consumed as a cursor (from the end)
    for (uint32_t i = threadIdx.x; i < npairs; i += 1024) {
        const uint32_t L = probe[i];
        const uint32_t pos = atomicSub(&hist[L], 1u) - 1u;
        p.pairs[poff[L] + pos] = i / npe;
    }
}

// count[L] is consumed as a cursor (filled from the end): pairs[pair_off[L] + ...] = query
__global__ void lm_fill_kernel(const uint32_t *__restrict__ probe, uint32_t npairs, uint32_t npe, uint32_t *__restrict__ count,
                               const uint32_t *__restrict__ pair_off, uint32_t *__restrict__ pairs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const uint32_t L = probe[i];
    const uint32_t pos = atomicSub(&count[L], 1u) - 1u;
    pairs[pair_off[L] + pos] = i / npe;
}

// The running bound of every query starts at the k-th best of its NEAREST list (an exact top k from the query-major
// kernel, ~1/npe of the rows): a lower bound of the k-th best over all its lists, one float32 step lower to be safe
// against an uncertified last digit.  Fewer than k documents there: no bound.
__global__ void lm_seed_kernel(const float *__restrict__ sims, const int32_t *__restrict__ counts, uint32_t nq, int k,
                               uint32_t *__restrict__ gthr) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint32_t t = 0;
    if (counts[q] >= k) {
        const uint32_t key = f32_to_key(sims[(size_t)q * k + k - 1]);
        t = key > 2u ? key - 1u : 0u;
    }
    gthr[q] = t;
}

// The same bound without the query-major kernel's machinery.  ANY k distinct documents with certified scores bound the
// k-th best score from below, so nothing here has to be an exact selection: one block of eight warps per query scores
// the first `sample` rows of the query's nearest list (probe[q][0]) in tiles of 32 rows, a warp sorts each of its tiles in
// registers (bitonic, shuffles) and contributes the tile's best 16 rows, and the block ranks the 256 contributions, one per
// document.  Rows whose float32 rounding is not certified are left out (the bound only gets lower).
constexpr int kSeedWarps = 8, kSeedKeep = 16, kSeedTiles = 16;  // 16 tiles x 32 rows = 512 rows at most
template <int G, int CPL>
__global__ void __launch_bounds__(32 * kSeedWarps, 2)
lm_seed_scan_kernel(MatView rows, const uint64_t *__restrict__ ids, uint64_t id_base, MatView queries,
                    const uint32_t *__restrict__ probe, uint32_t npe,
                    const uint64_t *__restrict__ list_off, const uint64_t *__restrict__ list_len, uint32_t sample, int k,
                    uint32_t *__restrict__ gthr) {
    constexpr int NG = 32 / G, TR = 32, N = kSeedTiles * kSeedKeep;  // 256 = blockDim.x
    __shared__ uint64_t s_id[N + 32];
    __shared__ uint32_t s_key[N + 32], s_meta[N + 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x;
    const int D = rows.d, d_pad = rows.d_pad;
    const bool dedup = ids != nullptr;
    // The sample: the first rows of the query's nearest list -- and, where that list is short, of its next lists in rank
    // order, until `sample` rows or kSeedTiles tiles are taken (a tile never spans two lists).  A query whose nearest list
    // holds fewer than k documents would otherwise start the scan without any bound.
    __shared__ uint64_t t_row0[kSeedTiles];
    __shared__ uint32_t t_nr[kSeedTiles];
    if (threadIdx.x == 0) {
        const uint32_t want = min(sample, (uint32_t)(kSeedTiles * TR));
        uint32_t taken = 0;
        int nt = 0;
        for (uint32_t j = 0; j < npe && nt < kSeedTiles && taken < want; j++) {
            const uint32_t L = probe[(size_t)q * npe + j];
            const uint64_t st = list_off[L];
            const uint32_t len = (uint32_t)min((uint64_t)(want - taken), list_len[L]);
            for (uint32_t r0 = 0; r0 < len && nt < kSeedTiles; r0 += TR) {
                t_row0[nt] = st + r0;
                t_nr[nt] = min((uint32_t)TR, len - r0);
                taken += t_nr[nt];
                nt++;
            }
        }
        for (; nt < kSeedTiles; nt++) t_nr[nt] = 0;
    }
    __syncthreads();
    uint4 qreg[CPL];
    const uint8_t *qc = queries.codes + (size_t)q * d_pad;
#pragma unroll
    for (int t = 0; t < CPL; t++) qreg[t] = *reinterpret_cast<const uint4 *>(qc + ((lane % G) + G * t) * 16);
    // (its own copy of the query's side constants: this kernel runs beside lm_prepare*, which writes p.sides)
    const float2 qh = queries.hdr[q];
    const uint2 qs = queries.sums[q];
    const SideConst xq = make_side(qh.x, qh.y, qs.x, qs.y, queries.d);
    for (int tile = warp; tile < kSeedTiles; tile += kSeedWarps) {
        uint32_t key = 0, meta = 0;
        uint64_t id = kEmptyId;
        if (t_nr[tile] != 0) {
            const int nr = (int)t_nr[tile];
            const uint64_t first = t_row0[tile];
            const int iters = (nr + NG - 1) / NG;
            const uint32_t mydot = tile_dots<G, CPL>(rows.codes, (size_t)first, nr, d_pad, qreg, lane, iters);
            const int myr = (lane / G) * iters + (lane % G);
            if ((lane % G) < iters && myr < nr) {
                const uint64_t row = first + (uint32_t)myr;
                const float2 h = rows.hdr[row];
                const uint2 sm = rows.sums[row];
                bool flag;
                const float sim = score_fast(xq, h.x, h.y, sm.x, sm.y, mydot, D, &flag);
                if (!flag) {
                    key = f32_to_key(sim);
                    meta = (uint32_t)row;
                    id = ids ? ids[row] : id_base + row;
                }
            }
            warp_sort32(key, meta, id, lane);
        }
        if (lane < kSeedKeep) {
            s_key[tile * kSeedKeep + lane] = key;
            s_meta[tile * kSeedKeep + lane] = meta;
            s_id[tile * kSeedKeep + lane] = id;
        }
    }
    __syncthreads();
    CandBuf src{s_key, s_meta, s_id};
    // ranked in place of a second buffer: only the k-th key is needed, so dst is a 32-entry window
    __shared__ uint64_t d_id[32];
    __shared__ uint32_t d_key[32], d_meta[32];
    CandBuf dst{d_key, d_meta, d_id};
    const int uniq = block_rank_small(src, N, dst, 32, dedup);
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        if (uniq >= k && k <= 32) {
            const uint32_t kk = d_key[k - 1];
            t = kk > 2u ? kk - 1u : 0u;
        }
        gthr[q] = t;
    }
}

// The query side of the score identity, once per query and step (compute/cosine.go:26,138-149 folded into integers).
__global__ void lm_side_kernel(MatView queries, SideConst *__restrict__ sides) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= queries.n) return;
    const float2 h = queries.hdr[q];
    const uint2 s = queries.sums[q];
    sides[q] = make_side(h.x, h.y, s.x, s.y, queries.d);
}

struct LmCtl {
    uint64_t full[kLmStages];
    unsigned int done[kLmStages];   // warps of the owning group that have drained the stage
    unsigned int base[kLmStages];   // completed phases of the stage's barrier before the current pass
    unsigned int item, next_item;
};

}  // namespace

// ---- 2./3. the scan ------------------------------------------------------------------------------------------------
// WARPS x STAGES: 16 x 8 = one block per SM; 8 x 4 = two independent blocks per SM with half the ring each, so that one
// block streams while the other sits in the fixed part of an item (ring drained, queries loading).
template <int G, int CPL, int TR, int WARPS, int STAGES>
__global__ void __launch_bounds__(32 * WARPS, 16 / WARPS)
lm_scan_kernel(const LmParams p) {
    extern __shared__ __align__(128) unsigned char lsm[];
    constexpr int kLmWarps = WARPS, kLmStages = STAGES;  // (shadow the defaults)
    constexpr int NG = 32 / G;
    constexpr int CAP = kLmCap, WB = kLmWB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = p.rows.d, d_pad = p.rows.d_pad;
    const uint32_t stage_bytes = (uint32_t)p.stage_bytes;
    unsigned char *ring = lsm;
    uint64_t *cand_id = reinterpret_cast<uint64_t *>(lsm + (size_t)kLmStages * stage_bytes);
    uint32_t *cand_key = reinterpret_cast<uint32_t *>(cand_id + kLmWarps * WB);
    uint32_t *cand_meta = cand_key + kLmWarps * WB;
    LmCtl &ctl = *reinterpret_cast<LmCtl *>(cand_meta + kLmWarps * WB);
    const bool dedup = p.ids != nullptr;
    const uint32_t nitems = *p.nitems;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kLmStages; i++) {
            mbar_init(smem_u32(&ctl.full[i]), 1);
            ctl.done[i] = 0;
            ctl.base[i] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ctl.item = atomicAdd(p.next_item, 1u);
    }
    __syncthreads();
    uint64_t *wid = cand_id + warp * WB;
    uint32_t *wkey = cand_key + warp * WB, *wmeta = cand_meta + warp * WB;

    for (;;) {
        const uint32_t item = ctl.item;
        if (item >= nitems) break;
        const LmItem it = p.items[item];
        if (threadIdx.x == 0) ctl.next_item = atomicAdd(p.next_item, 1u);  // (its latency hides behind this item)
        const uint32_t nchunks = (it.nrows + TR - 1) / TR;
        for (uint32_t pass0 = 0; pass0 < it.m; pass0 += kLmWarps) {
            const int m = (int)min((uint32_t)kLmWarps, it.m - pass0);  // queries scored in this pass, one warp each per phase
            const int r = min(kLmWarps / m, kLmStages);                 // phase groups (one query: 8 warps stream, 8 rest)
            const int depth = kLmStages / r;                            // ring stages a group owns
            const int j = warp % m, ph = warp / m;
            const bool active = ph < r;
            // ---- this warp's query ----
            uint4 qreg[CPL];
            SideConst xq;
            uint32_t q = 0, thr_w = 0;
            int cnt_w = 0;
            if (active) {
                q = p.pairs[it.pair_off + pass0 + j];
                const uint8_t *qc = p.queries.codes + (size_t)q * d_pad;
#pragma unroll
                for (int t = 0; t < CPL; t++) qreg[t] = *reinterpret_cast<const uint4 *>(qc + ((lane % G) + G * t) * 16);
                xq = p.sides[q];
                thr_w = __ldcg(p.gthr + q);
            }
            // chunk i of my group: rows [c * TR, ...) of the item with c = ph + i * r, in stage ph * depth + i % depth
            struct Pending {
                uint32_t nr;
                float2 h;
                uint2 sums;
                uint64_t id;
            };
            auto arm = [&](uint32_t c, uint32_t st) {  // (one lane)
                const uint32_t nr = min((uint32_t)TR, it.nrows - c * TR);
                const uint32_t full = smem_u32(&ctl.full[st]);
                const uint32_t cb = nr * (uint32_t)d_pad;
                mbar_expect_tx(full, cb);
                bulk_g2s(smem_u32(ring + (size_t)st * stage_bytes), p.rows.codes + (uint64_t)(it.row0 + c * TR) * d_pad, cb, full);
            };
            auto side = [&](uint32_t c, Pending &pd) {  // per lane: the side data of the row it will finish in chunk c
                pd.nr = min((uint32_t)TR, it.nrows - c * TR);
                const int iters = ((int)pd.nr + NG - 1) / NG;
                const int myr = (lane / G) * iters + (lane % G);
                pd.h = make_float2(0.f, 0.f);
                pd.sums = make_uint2(0, 0);
                pd.id = kEmptyId;
                if ((lane % G) < iters && myr < (int)pd.nr) {
                    const uint32_t row = it.row0 + c * TR + (uint32_t)myr;
                    pd.h = p.rows.hdr[row];
                    pd.sums = p.rows.sums[row];
                    pd.id = p.ids ? p.ids[row] : p.id_base + row;
                }
            };
            if (active) {
                if (j == 0 && lane == 0) {  // the group's first warp arms the group's stages
                    for (int i = 0; i < depth; i++) {
                        const uint32_t c = (uint32_t)ph + (uint32_t)i * r;
                        if (c < nchunks) arm(c, (uint32_t)(ph * depth + i));
                    }
                }
                Pending cur;
                if ((uint32_t)ph < nchunks) side((uint32_t)ph, cur);
                uint32_t slot = 0, turn = 0;  // chunk i of the group sits in stage ph * depth + i % depth, on ring turn i / depth
                for (uint32_t c = (uint32_t)ph; c < nchunks; c += r) {
                    const uint32_t st = (uint32_t)(ph * depth) + slot;
                    mbar_wait(smem_u32(&ctl.full[st]), (ctl.base[st] + turn) & 1u);
                    if (++slot == (uint32_t)depth) {
                        slot = 0;
                        turn++;
                    }
                    const int iters = ((int)cur.nr + NG - 1) / NG;
                    const uint32_t mydot = stage_dots<G, CPL>(smem_u32(ring + (size_t)st * stage_bytes), (int)cur.nr, d_pad, qreg, lane, iters);
                    const int myr = (lane / G) * iters + (lane % G);
                    const bool valid = (lane % G) < iters && myr < (int)cur.nr;
                    __syncwarp();
                    // the last warp of the group to drain the stage re-arms it with the group's chunk `depth` turns ahead
                    if (lane == 0) {
                        const unsigned int old = atomicAdd(&ctl.done[st], 1u);
                        if (old == (unsigned)m - 1u) {
                            ctl.done[st] = 0;
                            const uint32_t cn = c + (uint32_t)r * depth;
                            if (cn < nchunks) arm(cn, st);
                        }
                    }
                    const Pending now = cur;
                    if (c + r < nchunks) side(c + r, cur);  // next chunk's side data in flight during the scoring
                    uint32_t key = 0, meta = 0;
                    if (valid) {
                        bool flag;
                        const float sim = score_fast(xq, now.h.x, now.h.y, now.sums.x, now.sums.y, mydot, D, &flag);
                        key = f32_to_key(sim);
                        meta = (it.row0 + c * TR + (uint32_t)myr) | (flag ? kFlagBit : 0u);
                    }
                    unsigned mk = __ballot_sync(FULL, valid && key >= thr_w);
                    if (cnt_w + __popc(mk) > WB) {
                        // cut the buffer to its best CAP distinct documents; the CAP-th becomes the threshold
                        WarpTopK<1> top;
                        top.init();
                        for (int base = 0; base < cnt_w; base += 32) {
                            const int e = base + lane;
                            const bool have = e < cnt_w;
                            top.offer(have, have ? wkey[e] : 0u, have ? wmeta[e] : 0u, have ? wid[e] : kEmptyId, lane, dedup);
                        }
                        __syncwarp();
                        wkey[lane] = top.skey[0];
                        wmeta[lane] = top.meta[0];
                        wid[lane] = top.id[0];
                        cnt_w = top.count();
                        thr_w = max(thr_w, top.thr_key);
                        __syncwarp();
                        mk = __ballot_sync(FULL, valid && key >= thr_w);
                    }
                    if (mk & (1u << lane)) {
                        const int slot = cnt_w + __popc(mk & ((1u << lane) - 1u));
                        wkey[slot] = key;
                        wmeta[slot] = meta;
                        wid[slot] = now.id;
                    }
                    cnt_w += __popc(mk);
                }
                // ---- what this warp kept goes to the query's candidate list; k distinct documents raise its running bound ----
                __syncwarp();
                if (cnt_w >= p.k) {
                    WarpTopK<1> top;
                    top.init();
                    for (int base = 0; base < cnt_w; base += 32) {
                        const int e = base + lane;
                        const bool have = e < cnt_w;
                        top.offer(have, have ? wkey[e] : 0u, have ? wmeta[e] : 0u, have ? wid[e] : kEmptyId, lane, dedup);
                    }
                    __syncwarp();
                    wkey[lane] = top.skey[0];
                    wmeta[lane] = top.meta[0];
                    wid[lane] = top.id[0];
                    const int have = top.count();
                    cnt_w = min(have, p.pub);  // the best k distinct documents of a union are among the best k of every part
                    if (have >= p.k) {
                        const uint32_t kth = __shfl_sync(FULL, top.skey[0], p.k - 1);
                        if (lane == 0) atomicMax(p.gthr + q, kth);
                    }
                    __syncwarp();
                }
                if (cnt_w > 0) {
                    unsigned int at = 0;
                    if (lane == 0) at = atomicAdd(p.gcnt + q, (unsigned)cnt_w);
                    at = __shfl_sync(FULL, at, 0);
                    for (int e = lane; e < cnt_w; e += 32) {
                        if (at + e < (unsigned)p.gcap) {
                            uint4 v;
                            v.x = wkey[e];
                            v.y = wmeta[e];
                            v.z = (uint32_t)wid[e];
                            v.w = (uint32_t)(wid[e] >> 32);
                            p.gbuf[(size_t)q * p.gcap + at + e] = v;
                        }
                    }
                }
            }
            __syncthreads();  // every armed chunk has been drained by its whole group: the ring is idle
            if (threadIdx.x < kLmStages) {  // completed barrier phases of every stage, for the next pass / item
                const int st = (int)threadIdx.x, g = st / depth, d = st % depth;
                unsigned int uses = 0;
                if (g < r && (uint32_t)g < nchunks) {
                    const uint32_t n_g = (nchunks - (uint32_t)g + (uint32_t)r - 1u) / (uint32_t)r;  // chunks of group g
                    if ((uint32_t)d < n_g) uses = (n_g - (uint32_t)d + (uint32_t)depth - 1u) / (uint32_t)depth;
                }
                ctl.base[st] += uses;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) ctl.item = ctl.next_item;
        __syncthreads();
    }
}

// ---- 2b. the scan of a DENSE batch: several queries per probed list, scored on the tensor cores ------------------------
// From about four queries per probed list on (1024 queries x 32 probes over 4096 lists: eight) the dp4a scan above is bound
// by instruction issue, not by HBM: 28 warp instructions per (query, row) pair.  Here the staged rows of an item are the A
// operand and up to 16 of the item's queries the B operand of mma.sync.m16n8k32 (u8 x u8 -> s32, exact): a warp takes one
// 16-row tile of a stage against all the queries of the pass, ~1 warp instruction per pair, and the scan is HBM-bound again
// up to a few dozen queries per list.  (An m16n8k32 fragment wants, per lane, 4 consecutive K bytes of a row; a lane loads
// 16 B of its row instead and feeds the four words to two instructions -- the same permutation of K on both operands, which
// a dot product does not see -- so a row tile costs one 128-bit shared load per instruction.)
// Same ring as above (4 stages x 32 rows, one bulk copy each, two blocks per SM); warps 2s and 2s+1 own stage s, one row
// tile each, and the second of them to finish its instructions re-arms the stage before either scores.  A pair at or above
// the query's bound (the seed of lm_seed_scan_kernel; raised by whoever finds k documents) is appended to the query's
// candidate list directly -- there is no per-warp list when a warp serves 16 queries -- and lm_final_kernel picks the top k
// as before: the kept set is a superset of what the dp4a scan keeps, so the emitted hits are the same.
constexpr int kDnWarps = 8, kDnStages = 4, kDnRows = 32, kDnQ = 16;
constexpr int kDnReserveDefault = 0;  // blocks of the 2-per-SM grid left out (VS_LM_DENSE_RESERVE)
constexpr int kDnQStride(int d_pad) { return d_pad + 64; }  // query rows 64 B apart in bank phase: conflict-free 128-bit loads

struct DnCtl {
    uint64_t full[kDnStages];
    unsigned int done[kDnStages];
    unsigned int item, next_item;
};

__device__ __forceinline__ void mma_u8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// A float32 screen in front of the certified float64 score.  With r = 1/sqrt(P) per side the cosine is
//   c = (Rx rx D)(Ry ry) dot - (Rx rx s1x)(Ry ry s1y) + (ux rx)(uy ry)  =  t1 - t2 + t3,
// and a pair is skipped only if  t1 - t2 + t3 + E < bound  in float32, where E covers (a) the float32 evaluation, 2^-20 of
// |t1| + |t2| + |t3| and of the cancellation scales M r inside ux and uy (at most ~7 float32 roundings per term: 16 are
// budgeted), (b) everything the certified half-width delta of score_core can hold besides those terms (the per-side parts
// Ex, Ey, with |c| <= 2) and (c) one float32 step of the final rounding (4e-6).  So a skipped pair's reference score lies
// strictly below the query's bound whether or not its float32 rounding could have been certified; NaN or Inf anywhere (P <= 0,
// all-zero vectors, overflow) makes the comparison false and the pair takes the certified path.  ~16 instructions instead of
// ~150, for the 99 % of the pairs that are nowhere near the bound.
struct DnRowSide {
    float Ay, AyS, By, Myp, Ey;
};
__device__ __forceinline__ DnRowSide dn_row_side(float mn, float mx, uint32_t s1, uint32_t s2, int D) {
    const double a = (double)mn, R = (double)mx - a;
    const double DA = (double)D * (255.0 * a), rs = R * (double)s1, ux = DA + rs, Md = fabs(DA) + fabs(rs);
    const long long I = (long long)D * (long long)s2 - (long long)s1 * (long long)s1;
    const double r2i = R * R * (double)I, u2 = ux * ux, P = r2i + u2;
    const float rP = rsqrtf((float)P);
    const float eP = 2.0f * fabsf((float)r2i) + 4.0f * fabsf((float)ux) * (float)Md + (float)u2 + fabsf((float)P);
    DnRowSide o;
    o.Ay = (float)R * rP;
    o.AyS = o.Ay * (float)s1;
    o.By = (float)ux * rP;
    o.Myp = (float)Md * rP * 1.0001f;
    o.Ey = 1.5e-16f * (eP * rP * rP + 3.5f * (float)D * (255.0f * (fabsf(mn) + fabsf((float)R)) * rP) + 3.0f * (float)D + 16.0f) + 4.0e-6f;
    return o;
}
struct DnQuerySide {
    float Ax, Cx, Bx, Mxp, Ex, bound;
};
__device__ __forceinline__ DnQuerySide dn_query_side(const SideConst &x, int D, uint32_t thr_key) {
    const double rx = rsqrt(x.P);
    DnQuerySide o;
    o.Ax = (float)(x.R * rx * (double)D);
    o.Cx = (float)(x.R * rx * (double)x.s1);
    o.Bx = (float)(x.ux * rx);
    o.Mxp = (float)((double)x.M * rx) * 1.0001f;
    o.Ex = 1.5e-16f * ((float)x.epP + 3.5f * (float)D * (float)x.mgs + 3.0f * (float)D + 16.0f);
    o.bound = thr_key < 2u ? __int_as_float(0xFF800000) : key_to_f32(thr_key);
    return o;
}
// true: the pair cannot reach the query's bound
__device__ __forceinline__ bool dn_skip(const DnQuerySide &x, const DnRowSide &y, uint32_t dot) {
    const float t1 = (x.Ax * y.Ay) * __uint2float_rn(dot), t2 = x.Cx * y.AyS, t3 = x.Bx * y.By;
    const float mag = fabsf(t1) + fabsf(t2) + fabsf(t3) + fmaf(fabsf(y.By), x.Mxp, fabsf(x.Bx) * y.Myp);
    const float E = fmaf(9.5367431640625e-7f, mag, x.Ex + y.Ey);
    return (t1 - t2) + t3 + E < x.bound;
}

// (a broken ring must fault, not hang the device: ~4 s)
__device__ __forceinline__ void dn_wait(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 28)) __trap();
    } while (!done);
}

__global__ void __launch_bounds__(32 * kDnWarps, 2)
lm_dense_kernel(const LmParams p) {
    extern __shared__ __align__(128) unsigned char dsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = p.rows.d, d_pad = p.rows.d_pad;
    const int nj = d_pad >> 6;                      // 64 K bytes per step: two instructions
    const uint32_t stage_bytes = (uint32_t)(kDnRows * d_pad);
    const uint32_t qstride = (uint32_t)kDnQStride(d_pad);
    unsigned char *ring = dsm;
    unsigned char *qsm = dsm + (size_t)kDnStages * stage_bytes;                  // [kDnQ][qstride] query codes of the pass
    SideConst *sq = reinterpret_cast<SideConst *>(qsm + (size_t)kDnQ * qstride);  // [kDnQ]
    uint32_t *sqid = reinterpret_cast<uint32_t *>(sq + kDnQ);                     // [kDnQ] query numbers
    uint32_t *sthr = sqid + kDnQ;                                                 // [kDnQ] bounds (score keys)
    DnQuerySide *spre = reinterpret_cast<DnQuerySide *>(sthr + kDnQ);             // [kDnQ] float32 screen, query side
    DnCtl &ctl = *reinterpret_cast<DnCtl *>(spre + kDnQ);
    const uint32_t nitems = *p.nitems;
    const int st = warp >> 1, tile = warp & 1;      // my stage, my row tile
    const int g = lane >> 2, c4 = lane & 3;
    const uint32_t full = smem_u32(&ctl.full[st]);
    uint32_t phase = 0;                             // parity of my stage's barrier (only its two warps ever wait on it)

    if (threadIdx.x == 0) {
        for (int i = 0; i < kDnStages; i++) {
            mbar_init(smem_u32(&ctl.full[i]), 1);
            ctl.done[i] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ctl.item = atomicAdd(p.next_item, 1u);
    }
    __syncthreads();

    for (;;) {
        const uint32_t item = ctl.item;
        if (item >= nitems) break;
        const LmItem it = p.items[item];
        if (threadIdx.x == 0) ctl.next_item = atomicAdd(p.next_item, 1u);
        const uint32_t nchunks = (it.nrows + kDnRows - 1) / kDnRows;
        auto arm = [&](uint32_t c) {  // (one lane) chunk c of the item into my stage
            const uint32_t nr = min((uint32_t)kDnRows, it.nrows - c * kDnRows);
            const uint32_t cb = nr * (uint32_t)d_pad;
            mbar_expect_tx(full, cb);
            bulk_g2s(smem_u32(ring + (size_t)st * stage_bytes), p.rows.codes + (uint64_t)(it.row0 + c * kDnRows) * d_pad, cb, full);
        };
        for (uint32_t pass0 = 0; pass0 < it.m; pass0 += kDnQ) {
            const int mq = (int)min((uint32_t)kDnQ, it.m - pass0);
            const int ngr = (mq + 7) >> 3;
            // the rows start to stream while the queries of the pass are fetched
            if (tile == 0 && lane == 0 && (uint32_t)st < nchunks) arm((uint32_t)st);
            for (int i = threadIdx.x; i < mq * (d_pad >> 4); i += 32 * kDnWarps) {
                const int qi = i / (d_pad >> 4), ch = i % (d_pad >> 4);
                const uint32_t q = p.pairs[it.pair_off + pass0 + qi];
                *reinterpret_cast<uint4 *>(qsm + (size_t)qi * qstride + ch * 16) =
                    *reinterpret_cast<const uint4 *>(p.queries.codes + (size_t)q * d_pad + ch * 16);
            }
            if ((int)threadIdx.x < mq) {
                const uint32_t q = p.pairs[it.pair_off + pass0 + threadIdx.x];
                sqid[threadIdx.x] = q;
                sq[threadIdx.x] = p.sides[q];
                const uint32_t thr = __ldcg(p.gthr + q);
                sthr[threadIdx.x] = thr;
                spre[threadIdx.x] = dn_query_side(p.sides[q], D, thr);
            }
            __syncthreads();
            uint32_t fresh = 0;  // bound of query `lane` of the pass as last read from global memory (0: nothing newer)
            const uint32_t a_base = smem_u32(ring + (size_t)st * stage_bytes) + (uint32_t)((tile * 16 + g) * d_pad + c4 * 16);
            const uint32_t b_base = smem_u32(qsm) + (uint32_t)g * qstride + (uint32_t)c4 * 16u;
            for (uint32_t c = (uint32_t)st; c < nchunks; c += kDnStages) {
                // side data of this lane's two rows (tile rows g and g + 8), in flight during the wait
                const uint32_t nr = min((uint32_t)kDnRows, it.nrows - c * kDnRows);
                const uint32_t r0 = (uint32_t)(tile * 16 + g), r1 = r0 + 8;
                const uint32_t row0 = it.row0 + c * kDnRows + r0, row1 = row0 + 8;
                float2 h0 = make_float2(0.f, 0.f), h1 = h0;
                uint2 s0 = make_uint2(0, 0), s1 = s0;
                if (r0 < nr) {
                    h0 = p.rows.hdr[row0];
                    s0 = p.rows.sums[row0];
                }
                if (r1 < nr) {
                    h1 = p.rows.hdr[row1];
                    s1 = p.rows.sums[row1];
                }
                // the bounds of the pass's queries may have been raised since the pass began (by any block: the recomputation
                // below); a stale value is still a valid bound, so the unsynchronised update is harmless
                // (software-pipelined: the value loaded during the previous chunk is applied here and the next load is issued
                // for the chunk after this one, so the round trip to L2 never sits in front of anything -- on short per-rank
                // lists the scan is bound by latency, not by HBM, and a load used within the chunk cost a fifth of it)
                if (lane < mq && fresh > sthr[lane]) {
                    sthr[lane] = fresh;
                    spre[lane].bound = key_to_f32(fresh);
                }
                __syncwarp();
                if (p.tighten_at > 0 && lane < mq) fresh = __ldcg(p.gthr + sqid[lane]);
                dn_wait(full, phase);
                phase ^= 1u;
                int acc[2][4];
#pragma unroll
                for (int gr = 0; gr < 2; gr++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[gr][e] = 0;
                if ((uint32_t)(tile * 16) < nr) {
                    if (ngr == 1) {
#pragma unroll 4
                        for (int j = 0; j < nj; j++) {
                            const uint4 a0 = lds_u4(a_base + (uint32_t)j * 64u), a1 = lds_u4(a_base + (uint32_t)(8 * d_pad) + (uint32_t)j * 64u);
                            const uint4 b = lds_u4(b_base + (uint32_t)j * 64u);
                            mma_u8(acc[0], a0.x, a1.x, a0.y, a1.y, b.x, b.y);
                            mma_u8(acc[0], a0.z, a1.z, a0.w, a1.w, b.z, b.w);
                        }
                    } else {
#pragma unroll 4
                        for (int j = 0; j < nj; j++) {
                            const uint4 a0 = lds_u4(a_base + (uint32_t)j * 64u), a1 = lds_u4(a_base + (uint32_t)(8 * d_pad) + (uint32_t)j * 64u);
                            const uint4 b = lds_u4(b_base + (uint32_t)j * 64u), b2 = lds_u4(b_base + 8u * qstride + (uint32_t)j * 64u);
                            mma_u8(acc[0], a0.x, a1.x, a0.y, a1.y, b.x, b.y);
                            mma_u8(acc[0], a0.z, a1.z, a0.w, a1.w, b.z, b.w);
                            mma_u8(acc[1], a0.x, a1.x, a0.y, a1.y, b2.x, b2.y);
                            mma_u8(acc[1], a0.z, a1.z, a0.w, a1.w, b2.z, b2.w);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {  // the second warp of the stage to get here hands it to the next chunk
                    const unsigned int old = atomicAdd(&ctl.done[st], 1u);
                    if (old == 1u) {
                        ctl.done[st] = 0;
                        if (c + kDnStages < nchunks) arm(c + kDnStages);
                    }
                }
                const DnRowSide y0 = dn_row_side(h0.x, h0.y, s0.x, s0.y, D), y1 = dn_row_side(h1.x, h1.y, s1.x, s1.y, D);
                // ---- scores: accumulator e of group gr = (row g + 8 (e >> 1), query 8 gr + 2 c4 + (e & 1)) ----
                // the screen over this lane's pairs (static register indices), then the few that pass it one by one
                uint32_t need = 0, tq = 0, tn = 0;
#pragma unroll
                for (int gr = 0; gr < 2; gr++) {
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int qi = gr * 8 + c4 * 2 + (e & 1);
                        const bool hi = (e & 2) != 0;
                        if (gr < ngr && qi < mq && (hi ? r1 : r0) < nr && !dn_skip(spre[qi], hi ? y1 : y0, (uint32_t)acc[gr][e]))
                            need |= 1u << (gr * 4 + e);
                    }
                }
                while (need) {
                    const int bit = __ffs((int)need) - 1;
                    need &= need - 1u;
                    const int gr = bit >> 2, e = bit & 3;
                    const int lo4 = gr ? acc[1][0] : acc[0][0], lo5 = gr ? acc[1][1] : acc[0][1], lo6 = gr ? acc[1][2] : acc[0][2],
                              lo7 = gr ? acc[1][3] : acc[0][3];
                    const uint32_t dot = (uint32_t)((e & 2) ? ((e & 1) ? lo7 : lo6) : ((e & 1) ? lo5 : lo4));
                    const int qi = gr * 8 + c4 * 2 + (e & 1);
                    const bool hi = (e & 2) != 0;
                    const float2 h = hi ? h1 : h0;
                    const uint2 sm = hi ? s1 : s0;
                    bool flag;
                    const float sim = score_fast(sq[qi], h.x, h.y, sm.x, sm.y, dot, D, &flag);
                    const uint32_t key = f32_to_key(sim);
                    if (key < sthr[qi]) continue;
                    const uint32_t row = hi ? row1 : row0;
                    const uint32_t q = sqid[qi];
                    const uint64_t id = p.ids ? p.ids[row] : p.id_base + row;
                    const unsigned int at = atomicAdd(p.gcnt + q, 1u);
                    if (at < (unsigned)p.gcap) {
                        uint4 v;
                        v.x = key;
                        v.y = row | (flag ? kFlagBit : 0u);
                        v.z = (uint32_t)id;
                        v.w = (uint32_t)(id >> 32);
                        p.gbuf[(size_t)q * p.gcap + at] = v;
                    }
                    // the candidate that fills the list to 1x, 2x or 4x tighten_at makes its warp recompute the bound
                    if (p.tighten_at > 0 && (at + 1u == (unsigned)p.tighten_at || at + 1u == 2u * (unsigned)p.tighten_at || at + 1u == 4u * (unsigned)p.tighten_at)) {
                        tq = q;
                        tn = at + 1u;
                    }
                }
                // ---- a query's list has grown long: its bound becomes the k-th best certified document among what the list
                // holds so far (entries still in flight read as zero and are left out: the bound only gets lower) ----
                for (unsigned tm = __ballot_sync(FULL, tn != 0u); tm; tm &= tm - 1u) {
                    const int src = __ffs((int)tm) - 1;
                    const uint32_t q = __shfl_sync(FULL, tq, src), n = __shfl_sync(FULL, tn, src);
                    const uint4 *src_buf = p.gbuf + (size_t)q * p.gcap;
                    WarpTopK<1> top;
                    top.init();
                    for (uint32_t base = 0; base < n; base += 32) {
                        const uint32_t e2 = base + (uint32_t)lane;
                        uint4 v = make_uint4(0, 0, 0, 0);
                        if (e2 < n) v = __ldcg(src_buf + e2);
                        const bool have = v.x != 0u && !(v.y & kFlagBit);
                        top.offer(have, v.x, v.y, (uint64_t)v.z | ((uint64_t)v.w << 32), lane, p.ids != nullptr);
                    }
                    if (top.count() >= p.k) {
                        const uint32_t kth = __shfl_sync(FULL, top.skey[0], p.k - 1);
                        if (lane == 0 && kth > 2u) atomicMax(p.gthr + q, kth - 1u);
                    }
                }
            }
            __syncthreads();  // every chunk of the pass is scored: the ring is idle, the query buffers are free
        }
        if (threadIdx.x == 0) ctl.item = ctl.next_item;
        __syncthreads();
    }
}

// ---- 4. per query: best k distinct documents of its candidate list ----------------------------------------------------
constexpr int kLmFinalThreads = 256;
constexpr int kLmFinalSurv = 256;  // survivors of the threshold cut that are ranked (more: the caller's literal path)

__global__ void __launch_bounds__(kLmFinalThreads)
lm_final_kernel(const uint4 *__restrict__ gbuf, const unsigned int *__restrict__ gcnt, uint32_t gcap, int k, bool dedup, MatView rows,
                MatView queries, uint64_t *__restrict__ out_ids, float *__restrict__ out_sims, int32_t *__restrict__ out_counts,
                uint32_t *__restrict__ out_status, unsigned long long *fix_counter) {
    extern __shared__ __align__(16) unsigned char lf_raw[];
    double *sh_qn = reinterpret_cast<double *>(lf_raw);  // [d] normalized query (literal re-score only)
    __shared__ uint32_t s_gmax[kLmFinalThreads];
    __shared__ uint32_t s_thr;
    __shared__ unsigned int s_cnt, s_fix, s_bad;
    __shared__ double s_norm;
    __shared__ uint64_t a_id[kLmFinalSurv], b_id[kLmFinalSurv];
    __shared__ uint32_t a_key[kLmFinalSurv], a_meta[kLmFinalSurv], b_key[kLmFinalSurv], b_meta[kLmFinalSurv];
    const uint32_t q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int total = gcnt[q];
    const uint32_t n = min(total, gcap);
    const uint4 *src = gbuf + (size_t)q * gcap;
    uint32_t status = total > gcap ? kStatusListAmbiguous : 0u;  // candidates were lost: the literal path redoes the query
    // threshold: the k-th largest of the maxima of 64 groups of 4 threads
    uint32_t gm = 0;
    for (uint32_t e = threadIdx.x; e < n; e += kLmFinalThreads) gm = max(gm, __ldcg(&src[e].x));
    gm = max(gm, __shfl_xor_sync(FULL, gm, 1));
    gm = max(gm, __shfl_xor_sync(FULL, gm, 2));
    const int g = (int)threadIdx.x >> 2, part = (int)threadIdx.x & 3;
    if (threadIdx.x == 0) {
        s_thr = 0;
        s_cnt = 0;
        s_fix = 0;
        s_bad = 0;
    }
    if (part == 0) s_gmax[g] = gm;
    __syncthreads();
    {
        const uint32_t mine = s_gmax[g];
        int rk = 0;
        for (int u = part; u < kLmFinalThreads / 4; u += 4) {
            const uint32_t o = s_gmax[u];
            rk += (int)((o > mine) | ((o == mine) & (u < g)));
        }
        rk += __shfl_xor_sync(FULL, rk, 1);
        rk += __shfl_xor_sync(FULL, rk, 2);
        if (part == 0 && rk == k - 1) s_thr = mine;
    }
    __syncthreads();
    // The net is cast two keys below the k-th group maximum: a literal re-score moves an uncertified score down by at most
    // one float32 step (two keys across zero), so the k-th best after re-scoring still lies inside the net and nothing
    // outside it can matter.
    const uint32_t thr = s_thr > 2u ? s_thr - 2u : 0u;
    for (uint32_t e = threadIdx.x; e < n; e += kLmFinalThreads) {
        const uint4 v = __ldcg(src + e);
        if (v.x != 0 && v.x >= thr) {
            const unsigned int pos = atomicAdd(&s_cnt, 1u);
            if (pos < (unsigned)kLmFinalSurv) {
                a_key[pos] = v.x;
                a_meta[pos] = v.y;
                a_id[pos] = (uint64_t)v.z | ((uint64_t)v.w << 32);
            }
        }
    }
    __syncthreads();
    const unsigned int ns = s_cnt;
    if (ns > (unsigned)kLmFinalSurv) status |= kStatusListAmbiguous;
    const int nsurv = (int)min(ns, (unsigned)kLmFinalSurv);
    const CandBuf A{a_key, a_meta, a_id}, B{b_key, b_meta, b_id};
    int uniq = block_rank_small(A, nsurv, B, 32, dedup);
    // An uncertified score among the first k: every uncertified survivor is re-scored with the reference's own arithmetic
    // (compute/cosine.go:26-50, warp_ref_cosine_row_f64) and the survivors are ranked again.  Scores only move down, by at
    // most one float32 step; everything that was not a survivor lies below `thr`, so the new first k are exact as long as
    // the k-th of them still reaches `thr` (otherwise: the caller's literal path).
    {
        bool flag = false;
        if ((int)threadIdx.x < min(k, 32)) flag = b_key[threadIdx.x] != 0 && (b_meta[threadIdx.x] & kFlagBit);
        if (__syncthreads_or(flag)) {
            const int D = queries.d;
            const uint8_t *qc = queries.codes + (size_t)q * queries.d_pad;
            const float2 qh = queries.hdr[q];
            const double mn = (double)qh.x, range = __dsub_rn((double)qh.y, (double)qh.x);
            for (int i = threadIdx.x; i < D; i += kLmFinalThreads) sh_qn[i] = ref_dequant_f64(qc[i], mn, range);
            __syncthreads();
            if (warp == 0) {
                const double nsq = warp_ordered_sum(D, lane, [&](int i) { return __dmul_rn(sh_qn[i], sh_qn[i]); });
                if (lane == 0) s_norm = __dsqrt_rn(nsq);
            }
            __syncthreads();
            const double norm = s_norm;
            if (norm != 0.0)
                for (int i = threadIdx.x; i < D; i += kLmFinalThreads) sh_qn[i] = __ddiv_rn(sh_qn[i], norm);
            __syncthreads();
            for (int e = warp; e < nsurv; e += kLmFinalThreads / 32) {  // one warp per uncertified survivor (A still holds them)
                const uint32_t k0 = a_key[e], m0 = a_meta[e];
                if (k0 != 0 && (m0 & kFlagBit)) {
                    const size_t row = m0 & kMetaRowMask;
                    const float2 h = rows.hdr[row];
                    const double dot = warp_ref_cosine_row_f64(rows.codes + row * (size_t)rows.d_pad, h.x, h.y, sh_qn, D, lane);
                    if (lane == 0) {
                        const uint32_t k1 = f32_to_key(__double2float_rn(dot));
                        a_key[e] = k1;
                        a_meta[e] = m0 & kMetaRowMask;
                        if ((m0 & kSibBit) && k1 != k0) s_bad = 1;  // an equally scored row of the same document was dropped
                        atomicAdd(&s_fix, 1u);
                    }
                }
            }
            __syncthreads();
            uniq = block_rank_small(A, nsurv, B, 32, dedup);
            const uint32_t kth = (k <= 32 && uniq >= k) ? b_key[k - 1] : 0u;
            if (s_bad || (thr != 0 && n > (uint32_t)nsurv && kth < thr)) status |= kStatusListAmbiguous;
            if (threadIdx.x == 0 && fix_counter) atomicAdd(fix_counter, (unsigned long long)s_fix);
        }
    }
    // duplicates ate the margin of the cut although candidates below the threshold were left out: not decided here
    if (dedup && uniq < k && thr != 0 && n > (uint32_t)nsurv) status |= kStatusListAmbiguous;
    if (threadIdx.x < 32) {
        const int r = (int)threadIdx.x;
        const uint32_t kk = b_key[r];
        const bool have = kk != 0;
        bool flag = false;
        if (have && r < k) {
            out_ids[(size_t)q * k + r] = b_id[r];
            out_sims[(size_t)q * k + r] = key_to_f32(kk);
            flag = (b_meta[r] & kFlagBit) != 0;
        }
        const int cnt = __popc(__ballot_sync(FULL, have));
        if (__any_sync(FULL, flag)) status |= kStatusListAmbiguous;
        if (r == 0) {
            out_counts[q] = min(cnt, k);
            out_status[q] |= status;  // (the probe stage stored its own bits)
        }
    }
}

// ---------------------------------------------------------------------------------------------------
static int lm_tile_rows(int d_pad) { return d_pad <= 768 ? 32 : 16; }

static bool lm_geometry(int d_pad, int *stage_bytes, size_t *smem, int warps = kLmWarps, int stages = kLmStages) {
    const int sb = (lm_tile_rows(d_pad) * d_pad + 127) & ~127;
    *stage_bytes = sb;
    *smem = (size_t)stages * sb + (size_t)warps * kLmWB * 16 + ((sizeof(LmCtl) + 127) & ~size_t(127));
    return (*smem + 1024) * (size_t)(kLmWarps / warps) <= (size_t)227 * 1024;  // (1 KB per block is the system's)
}

bool lm_supported(int d_pad, int k) {
    if (d_pad & 15) return false;
    switch (d_pad >> 4) {
        case 48: case 32: case 64: case 96: case 24: break;
        default: return false;
    }
    int sb;
    size_t smem;
    return k >= 1 && k <= kLmCap && lm_geometry(d_pad, &sb, &smem);
}

size_t lm_items_cap(size_t n_rows, size_t nq, size_t npe) { return n_rows / lm_sub_rows() + nq * npe + 16; }

template <int G, int CPL, int TR, int WARPS, int STAGES>
static cudaError_t lm_launch_scan_t(LmParams p, int grid, cudaStream_t st) {
    size_t smem;
    if (!lm_geometry(p.rows.d_pad, &p.stage_bytes, &smem, WARPS, STAGES)) return cudaErrorInvalidValue;
    auto kern = lm_scan_kernel<G, CPL, TR, WARPS, STAGES>;
    static bool attr_of[64] = {false};  // (function attributes are per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_of[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_of[dev & 63] = true;
    }
    kern<<<grid * (kLmWarps / WARPS), 32 * WARPS, smem, st>>>(p);
    return cudaGetLastError();
}
template <int G, int CPL, int TR>
static cudaError_t lm_launch_scan(const LmParams &p, int grid, cudaStream_t st) {
    static const bool one_block = [] {
        const char *e = getenv("VS_LM_ONE_BLOCK");
        return e && atoi(e) != 0;
    }();
    int sb;
    size_t smem;
    if (!one_block && lm_geometry(p.rows.d_pad, &sb, &smem, 8, 4)) return lm_launch_scan_t<G, CPL, TR, 8, 4>(p, grid, st);
    return lm_launch_scan_t<G, CPL, TR, 16, 8>(p, grid, st);
}

// The list stage of a batch, list-major, in three steps (the caller brackets the scan with its profiling marks).
// probe: [nq][npe] list ids (already selected); the status words already hold the probe stage's bits.
cudaError_t lm_enqueue_prepare(const LmParams &p, const uint32_t *probe, uint32_t nq, uint32_t npe, uint32_t C,
                               const uint64_t *list_off, const uint64_t *list_len, uint32_t *count, uint32_t *pair_off,
                               uint32_t items_cap, cudaStream_t st,
                               uint64_t *launches) {
    const uint32_t npairs = nq * npe;
    cudaError_t e;
    if (C <= kLmFusedMaxC) {
        const size_t smem = (size_t)C * 8;
        if (smem > 40 * 1024) {
            e = cudaFuncSetAttribute(lm_prepare_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        lm_prepare_fused_kernel<<<1, 1024, smem, st>>>(probe, nq, npe, list_off, list_len, C, p, items_cap, lm_sub_rows());
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    if ((e = cudaMemsetAsync(count, 0, (size_t)C * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(p.gcnt, 0, (size_t)nq * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(p.next_item, 0, 4, st)) != cudaSuccess) return e;
    lm_count_kernel<<<(npairs + 255) / 256, 256, 0, st>>>(probe, npairs, count);
    lm_items_kernel<<<1, 1024, 0, st>>>(count, list_off, list_len, C, pair_off, p.items, p.nitems, items_cap, lm_sub_rows());
    lm_fill_kernel<<<(npairs + 255) / 256, 256, 0, st>>>(probe, npairs, npe, count, pair_off, p.pairs);
    lm_side_kernel<<<(nq + 127) / 128, 128, 0, st>>>(p.queries, p.sides);
    if (launches) *launches += 4;
    return cudaGetLastError();
}

cudaError_t lm_enqueue_seed(const LmParams &p, const float *first_list_sims, const int32_t *first_list_counts, uint32_t nq,
                            cudaStream_t st, uint64_t *launches) {
    lm_seed_kernel<<<(nq + 127) / 128, 128, 0, st>>>(first_list_sims, first_list_counts, nq, p.k, p.gthr);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t lm_enqueue_seed_scan(const LmParams &p, const uint32_t *probe, uint32_t nq, uint32_t npe, const uint64_t *list_off,
                                 const uint64_t *list_len, uint32_t sample, cudaStream_t st, uint64_t *launches) {
    if (launches) *launches += 1;
#define VS_LM_SEED(G, CPL)                                                                                                       \
    lm_seed_scan_kernel<G, CPL><<<nq, 32 * kSeedWarps, 0, st>>>(p.rows, p.ids, p.id_base, p.queries, probe, npe, list_off,          \
                                                                list_len, sample, p.k, p.gthr)
    switch (p.rows.d_pad >> 4) {
        case 48: VS_LM_SEED(16, 3); break;
        case 32: VS_LM_SEED(32, 1); break;
        case 64: VS_LM_SEED(32, 2); break;
        case 96: VS_LM_SEED(32, 3); break;
        case 24: VS_LM_SEED(8, 3); break;
        default: return cudaErrorInvalidValue;
    }
#undef VS_LM_SEED
    return cudaGetLastError();
}

// Dense form: d_pad a multiple of 64 (two instructions per step) and a 32-row stage of at most 24 KB.
bool lm_dense_supported(int d_pad) { return d_pad >= 64 && (d_pad & 63) == 0 && d_pad <= 768; }

static cudaError_t lm_launch_dense(const LmParams &p, int sm_count, cudaStream_t st) {
    const size_t smem = (size_t)kDnStages * kDnRows * p.rows.d_pad + (size_t)kDnQ * kDnQStride(p.rows.d_pad) + kDnQ * sizeof(SideConst) +
                        2 * kDnQ * 4 + kDnQ * sizeof(DnQuerySide) + ((sizeof(DnCtl) + 127) & ~size_t(127));
    static bool attr_of[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_of[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(lm_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_of[dev & 63] = true;
    }
    // Two blocks per SM fill every SM's shared memory and registers, so nothing of another search context (its probe
    // selection, inversion, seed, final selection: a quarter of a step's kernel time) can run beside the scan.  Leaving a few
    // blocks out -- those SMs then hold ONE scan block and have half their resources free -- lets the next step's small
    // kernels run during this step's scan; the scan itself is bound by HBM and does not need every block.
    static const int reserve = [] {
        const char *e = getenv("VS_LM_DENSE_RESERVE");
        const int v = e ? atoi(e) : -1;
        return v >= 0 && v <= 1024 ? v : kDnReserveDefault;
    }();
    int blocks = 2 * sm_count - reserve;
    if (blocks < sm_count) blocks = sm_count;
    lm_dense_kernel<<<blocks, 32 * kDnWarps, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t lm_enqueue_scan(const LmParams &p, int sm_count, cudaStream_t st, uint64_t *launches, bool dense) {
    if (launches) *launches += 1;
    if (dense) return lm_launch_dense(p, sm_count, st);
    switch (p.rows.d_pad >> 4) {
        case 48: return lm_launch_scan<16, 3, 32>(p, sm_count, st);  // 768-d (nomic-embed-text)
        case 32: return lm_launch_scan<32, 1, 32>(p, sm_count, st);  // 512-d (noop/ai.go)
        case 64: return lm_launch_scan<32, 2, 16>(p, sm_count, st);  // 1024-d
        case 96: return lm_launch_scan<32, 3, 16>(p, sm_count, st);  // 1536-d
        case 24: return lm_launch_scan<8, 3, 32>(p, sm_count, st);   // 384-d
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t lm_enqueue_final(const LmParams &p, uint32_t nq, uint64_t *out_ids, float *out_sims, int32_t *out_counts,
                             uint32_t *out_status, unsigned long long *fix_counter, cudaStream_t st, uint64_t *launches) {
    lm_final_kernel<<<nq, kLmFinalThreads, (size_t)p.rows.d * sizeof(double), st>>>(p.gbuf, p.gcnt, (uint32_t)p.gcap, p.k, p.ids != nullptr,
                                                                                    p.rows, p.queries, out_ids, out_sims, out_counts,
                                                                                    out_status, fix_counter);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t lm_set_certify_scale(float scale) { return cudaMemcpyToSymbol(c_certify_scale, &scale, sizeof(float)); }

}  // namespace vs
