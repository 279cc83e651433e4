// scan.cu -- the streaming uint8 scan with fused top-k (K3 + K5 of SURVEY.md 2b).
//
// Replaces: compute/cosine.go:13-57 (vector x matrix cosine) as driven by server/search.go:214
// (centroid scoring), :220-223 (nprobe cut), :241-273 (posting-list scan, running sort, dedup,
// truncate).  One kernel template serves the three shapes -- probe selection over the centroid
// table, the scan of the probed lists, and a flat scan -- as "segments" of a device matrix.
//
// Per row the only vector x vector work is the exact integer dot sum(q*v) (dp4a); the float64 cosine
// is rebuilt from it and per-row integer sums (common.cuh), with a rigorous error bound that tells
// when float32(score) is certain to equal the reference's.  Uncertain rows are flagged and resolved
// by the literal-arithmetic variant (EXACT=true) of the same kernel.
//
// Layout: a warp owns a tile of 32 consecutive rows. Lane groups of G lanes each stream one row per
// iteration with 128-bit loads (G*16 contiguous bytes per load instruction per group); the G-lane
// shuffle reduction leaves row (g*G+it)'s dot in lane g*G+it, so after G iterations every lane holds
// one row's dot and the float64 finishing math runs on all 32 lanes at once.
#include <cstdio>

#include "internal.h"
#include "topk.cuh"

namespace vs {


// ---------------------------------------------------------------------------------------------------
// Any dimension: whole warp per row (G = 32, NG = 1), query chunks read from shared memory.
__device__ __forceinline__ uint32_t tile_dots_generic(const uint8_t *__restrict__ codes, size_t row0, int nrows,
                                                      int d_pad, const uint4 *__restrict__ qs, int lane) {
    const int CH = d_pad >> 4;
    uint32_t mydot = 0;
    for (int r = 0; r < nrows; r++) {
        const uint8_t *p = codes + (row0 + r) * (size_t)d_pad;
        uint32_t acc = 0;
        for (int c = lane; c < CH; c += 32) acc = dot16(ld_stream_u4(p + c * 16), qs[c], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (lane == r) mydot = acc;
    }
    return mydot;
}

// ---------------------------------------------------------------------------------------------------
struct StageShared {
    uint32_t tile_prefix[kMaxSeg + 1];
    uint64_t seg_start[kMaxSeg];
    uint32_t seg_len[kMaxSeg];
    SideConst qside;
    double norm;
    unsigned int is_last;
    unsigned int need_fix;
    unsigned int extra_count;
    unsigned int overflow;
    int warp_cnt[kStageWarps];
    uint32_t tail_key;
    uint64_t tail_id;
};

// Work decomposition (both stages, any batch): the launch's tiles form one global sequence -- query 0's
// tiles, then query 1's, ... -- cut into gridDim.x equal contiguous ranges, one per block, so every SM
// gets the same amount of scanning whatever nq is.  Block b owns tiles [ceil(b*T/Gd), ceil((b+1)*T/Gd));
// tile t belongs to block floor(t*Gd/T).  A block walks the queries its range overlaps; per (query, block)
// pair it leaves one partial top list in slot q+b (unique because q and b both only grow along the walk),
// and the last block to finish a query merges that query's slots and emits.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define VS_TRACE(slot)                                                                  \
    do {                                                                                \
        if (p.trace && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 16 + (slot)] = globaltimer_ns(); \
    } while (0)

__device__ __forceinline__ uint32_t block_of_tile(uint64_t t, uint64_t T, uint32_t Gd) { return (uint32_t)((t * Gd) / T); }
__device__ __forceinline__ uint64_t first_tile_of_block(uint64_t b, uint64_t T, uint32_t Gd) { return (b * T + Gd - 1) / Gd; }

// G == 0 selects the generic (any d) tile routine; EXACT selects literal reference arithmetic.
template <int G, int CPL, int KPL, bool EXACT>
__global__ void __launch_bounds__(kStageWarps * 32, 2)
stage_kernel(const StageParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CAP = 32 * KPL;
    StageShared &sh = *reinterpret_cast<StageShared *>(smem_raw);
    size_t off = (sizeof(StageShared) + 15) & ~size_t(15);
    SortSmem &ss = *reinterpret_cast<SortSmem *>(smem_raw + off);
    off += (sizeof(SortSmem) + 15) & ~size_t(15);
    const CandBuf bufA{ss.key_a, ss.meta_a, ss.id_a};  // kSortCap entries
    const CandBuf bufB{ss.key_b, ss.meta_b, ss.id_b};  // kOutCap entries
    double *sh_qn = reinterpret_cast<double *>(smem_raw + off);  // [D] normalized query (fix path)
    off += (size_t)p.rows.d * sizeof(double);
    uint32_t *sh_qprefix = reinterpret_cast<uint32_t *>(smem_raw + off);  // [nq+1] when p.qtiles
    off += p.qtiles ? (((size_t)p.nq + 1) * 4 + 15) & ~size_t(15) : 0;
    uint4 *sh_q = reinterpret_cast<uint4 *>(smem_raw + off);  // generic path only

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = p.rows.d, d_pad = p.rows.d_pad;
    // Probe stage -> list stage of one search: the list stage is launched with programmatic stream serialization, so
    // its blocks are scheduled as the probe stage's blocks retire (while the last one still merges) and wait here
    // until that grid has completed and its writes (probe list, tile counts) are visible.
    if (p.pdl == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (p.pdl == 2) asm volatile("griddepcontrol.wait;" ::: "memory");
    constexpr int GG = (G == 0) ? 32 : G;
    const int iters = (G == 0 || EXACT) ? 32 : p.iters;  // rows per lane group in a tile
    const int tile_rows = (G == 0 || EXACT) ? 32 : (32 / GG) * iters;

    // ---- global tile space ----
    uint64_t T;
    if (p.qtiles) {
        if (warp == 0) {
            uint32_t carry = 0;
            for (int base = 0; base < p.nq; base += 32) {
                int s = base + lane;
                uint32_t v = 0;
                if (s < p.nq) v = p.qtiles[p.q_select ? p.q_select[s] : s];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_up_sync(FULL, v, o);
                    if (lane >= o) v += t;
                }
                if (s < p.nq) sh_qprefix[s + 1] = carry + v;
                carry += __shfl_sync(FULL, v, 31);
            }
            if (lane == 0) sh_qprefix[0] = 0;
        }
        __syncthreads();
        T = sh_qprefix[p.nq];
    } else {
        T = (uint64_t)p.uniform_tiles * p.nq;
    }
    if (T == 0) return;
    VS_TRACE(0);
    const uint32_t Gd = (uint64_t)gridDim.x < T ? gridDim.x : (uint32_t)T;  // never more blocks than tiles
    if (blockIdx.x >= Gd) return;
    uint64_t t0 = first_tile_of_block(blockIdx.x, T, Gd);
    const uint64_t t1 = first_tile_of_block((uint64_t)blockIdx.x + 1, T, Gd);
    if (t0 >= t1) return;
    int qslot;
    if (p.qtiles) {
        int lo = 0, hi = p.nq;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (sh_qprefix[mid] <= t0) lo = mid;
            else hi = mid;
        }
        qslot = lo;
    } else {
        qslot = (int)(t0 / p.uniform_tiles);
    }

    for (; t0 < t1; qslot++) {
        const uint64_t P0 = p.qtiles ? sh_qprefix[qslot] : (uint64_t)qslot * p.uniform_tiles;
        const uint64_t P1 = p.qtiles ? sh_qprefix[qslot + 1] : P0 + p.uniform_tiles;
        if (P1 <= t0) continue;  // a query with no tiles
        const uint64_t tend = t1 < P1 ? t1 : P1;
        const uint32_t ta = (uint32_t)(t0 - P0), tb = (uint32_t)(tend - P0);  // this block's tiles of this query
        t0 = tend;
        const int qi = p.q_select ? (int)p.q_select[qslot] : qslot;

        // ---- per-query prologue: segment table, query constants ----
        __syncthreads();
        const int nseg = p.seg_list ? p.nseg : 1;
        for (int s = threadIdx.x; s < nseg; s += blockDim.x) {
            uint64_t st, len;
            if (p.seg_list) {
                uint32_t L = p.seg_list[(size_t)qi * p.seg_stride + s];
                st = p.list_off[L];
                len = p.list_len[L];
                if (p.seg_cap && len > p.seg_cap) len = p.seg_cap;
            } else {
                st = p.single_start;
                len = p.single_count;
            }
            sh.seg_start[s] = st;
            sh.seg_len[s] = (uint32_t)len;
        }
        if (threadIdx.x == 0) {
            float2 h = p.queries.hdr[qi];
            uint2 s = p.queries.sums[qi];
            sh.qside = make_side(h.x, h.y, s.x, s.y, D);
        }
        if constexpr (G == 0 && !EXACT) {
            const uint4 *qsrc = reinterpret_cast<const uint4 *>(p.queries.codes + (size_t)qi * d_pad);
            for (int c = threadIdx.x; c < (d_pad >> 4); c += blockDim.x) sh_q[c] = qsrc[c];
        }
        __syncthreads();
        if (warp == 0) {  // inclusive scan of tiles per segment
            uint32_t carry = 0;
            for (int base = 0; base < nseg; base += 32) {
                int s = base + lane;
                uint32_t v = s < nseg ? (sh.seg_len[s] + tile_rows - 1) / tile_rows : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_up_sync(FULL, v, o);
                    if (lane >= o) v += t;
                }
                if (s < nseg) sh.tile_prefix[s + 1] = carry + v;
                carry += __shfl_sync(FULL, v, 31);
            }
            if (lane == 0) sh.tile_prefix[0] = 0;
        }
        __syncthreads();
        const uint32_t seg_tiles = sh.tile_prefix[nseg];
        const SideConst xq = sh.qside;

        uint4 qreg[CPL > 0 ? CPL : 1];
        if constexpr (G != 0 && !EXACT) {
            const uint8_t *qc = p.queries.codes + (size_t)qi * d_pad;
#pragma unroll
            for (int j = 0; j < CPL; j++) qreg[j] = *reinterpret_cast<const uint4 *>(qc + ((lane % G) + G * j) * 16);
        }
        const double *qn = EXACT ? p.qnorm + (size_t)qi * D : nullptr;

        WarpTopK<KPL> top;
        top.init();
        // one hit per document (search.go:260-268): removed BEFORE any list is cut to its capacity, at insert time and in
        // every merge.  Implicit ids (id_base + row) are distinct by construction.
        const bool dedup = (p.mode == 0) && (p.ids != nullptr);
        VS_TRACE(1);

        // ---- main loop: this block's tiles of this query, round-robin over its warps ----
        const uint32_t tb_eff = tb < seg_tiles ? tb : seg_tiles;  // a query padded to 1 tile may have none
        auto locate = [&](uint32_t t, size_t &row0, int &nrows) {
            int lo = 0, hi = nseg;  // segment with tile_prefix[seg] <= t < tile_prefix[seg+1]
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (sh.tile_prefix[mid] <= t) lo = mid;
                else hi = mid;
            }
            const uint32_t tin = t - sh.tile_prefix[lo];
            row0 = sh.seg_start[lo] + (size_t)tin * tile_rows;
            nrows = min((uint32_t)tile_rows, sh.seg_len[lo] - tin * tile_rows);
        };
        size_t row0 = 0;
        int nrows = 0;
        if (ta + warp < tb_eff) locate(ta + warp, row0, nrows);
        for (uint32_t t = ta + warp; t < tb_eff; t += kStageWarps) {
            // locate this warp's next tile one iteration ahead (keeps the shared-memory search off the load path)
            size_t next_row0 = 0;
            int next_nrows = 0;
            if (t + kStageWarps < tb_eff) {
                locate(t + kStageWarps, next_row0, next_nrows);
            }
            const int myr = (lane / GG) * iters + (lane % GG);  // tile row finished by this lane
            const bool valid = (lane % GG) < iters && myr < nrows;
            const size_t row = row0 + (valid ? myr : 0);

            float sim;
            bool flag = false;
            uint64_t cid = kEmptyId;
            if constexpr (EXACT) {
                sim = 0.0f;
                if (valid) {
                    float2 h = p.rows.hdr[row];
                    sim = ref_cosine_row(p.rows.codes + row * (size_t)d_pad, h.x, h.y, qn, D);
                }
            } else {
                // issue the per-row header / sums (and, while the list is still filling, id) loads ahead of the
                // code stream so their latency overlaps it instead of following it
                float2 h = valid ? p.rows.hdr[row] : make_float2(0.f, 0.f);
                uint2 s = valid ? p.rows.sums[row] : make_uint2(0, 0);
                if (top.thr_key == 0 && valid) cid = p.ids ? p.ids[row] : p.id_base + row;
                uint32_t mydot;
                if constexpr (G != 0) mydot = tile_dots<G, CPL>(p.rows.codes, row0, nrows, d_pad, qreg, lane, iters);
                else mydot = tile_dots_generic(p.rows.codes, row0, nrows, d_pad, sh_q, lane);
                sim = score_fast(xq, h.x, h.y, s.x, s.y, mydot, D, &flag);
            }
            uint32_t key = f32_to_key(sim);
            // lazy id: only rows that can still enter the list need their document id
            const bool cand = valid && key >= top.thr_key;
            if (cand && cid == kEmptyId) cid = p.ids ? p.ids[row] : p.id_base + row;
            top.offer(cand, key, (uint32_t)row | (flag ? kFlagBit : 0u), cid, lane, dedup);
            row0 = next_row0;
            nrows = next_nrows;
        }

        // ---- block merge: every warp's (sorted) list -> shared memory -> rank merge (duplicates out) -> partial slot ----
        CandBuf res = bufB;  // where the block's merged list (and later the query's final list) lives
        {
            const int cnt = top.count();
            if (lane == 0) sh.warp_cnt[warp] = cnt;
            top.store_soa(bufA, warp * CAP, lane);
            __syncthreads();
            VS_TRACE(2);
            res = merge_lists(bufA, kStageWarps, CAP, sh.warp_cnt, bufB, CAP, dedup, ss.scan_tmp);  // top CAP of the block
        }
        VS_TRACE(8);
        const uint32_t blo = block_of_tile(P0, T, Gd), bhi = block_of_tile(P1 - 1, T, Gd);
        {
            Cand *slot = p.partial + ((size_t)qslot + blockIdx.x) * CAP;
            for (int e = threadIdx.x; e < CAP; e += blockDim.x) {
                uint4 v;
                v.x = res.key[e];
                v.y = res.meta[e];
                v.z = (uint32_t)res.id[e];
                v.w = (uint32_t)(res.id[e] >> 32);
                *reinterpret_cast<uint4 *>(slot + e) = v;
            }
        }
        VS_TRACE(9);
        __threadfence();
        __syncthreads();
        VS_TRACE(10);
        if (threadIdx.x == 0) {
            unsigned int tk = atomicAdd(&p.tickets[qslot], 1u);
            sh.is_last = (tk == bhi - blo) ? 1u : 0u;
            sh.extra_count = 0;
            sh.overflow = 0;
        }
        __syncthreads();
        VS_TRACE(3);
        if (!sh.is_last) continue;

        // ---- last block of this query: merge its slots, fix, emit.  The final list ends up in bufB[0..CAP). ----
        __threadfence();
        {
            const Cand *base = p.partial + ((size_t)qslot + blo) * CAP;
            const int nslots = (int)(bhi - blo + 1);
            auto ld_cand = [&](int slot, int rank, uint32_t &k, uint32_t &m, uint64_t &id) {
                uint4 v = __ldcg(reinterpret_cast<const uint4 *>(base + (size_t)slot * CAP + rank));
                k = v.x;
                m = v.y;
                id = (uint64_t)v.z | ((uint64_t)v.w << 32);
            };
            if (nslots == 1) {
                // the block's own sorted list is still in res[0..CAP)
            } else if (nslots * CAP <= kSortCap) {
                // few slots: merge them all by rank
                __syncthreads();
                for (int e = threadIdx.x; e < nslots * CAP; e += blockDim.x) {
                    uint32_t k, m;
                    uint64_t id;
                    ld_cand(e / CAP, e % CAP, k, m, id);
                    cand_put(bufA, e, k, m, id);
                }
                __syncthreads();
                res = merge_lists(bufA, nslots, CAP, nullptr, bufB, CAP, dedup, ss.scan_tmp);
            } else {
                // many slots: the CAP-th best of the slot heads bounds the CAP-th best overall from below, so only
                // entries at least that good can matter; each slot is sorted, so they form a prefix of it.
                const int mh = (CAP + nslots - 1) / nslots;  // heads per slot so that at least CAP are gathered
                const int ng = nslots * mh;
                bool slow = ng > 2 * (int)blockDim.x;
                uint32_t tkey = 0;
                if (!slow) {
                    // each thread holds up to two gathered heads
                    uint32_t hk[2] = {0u, 0u}, hm[2] = {0u, 0u};
                    uint64_t hid[2] = {kEmptyId, kEmptyId};
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int e = threadIdx.x + u * blockDim.x;
                        if (e < ng) ld_cand(e / mh, e % mh, hk[u], hm[u], hid[u]);
                    }
                    VS_TRACE(11);
                    // exact CAP-th largest key among the heads, built bit by bit from block-wide counts
                    tkey = 0;
                    for (int bit = 31; bit >= 0; bit--) {
                        const uint32_t cand_t = tkey | (1u << bit);
                        int c = __syncthreads_count(hk[0] >= cand_t);
                        if (ng > (int)blockDim.x) c += __syncthreads_count(hk[1] >= cand_t);
                        if (c >= CAP) tkey = cand_t;
                    }
                    VS_TRACE(12);
                    // keep every entry whose key reaches the threshold: gathered heads first, then what the slots
                    // hold below their last gathered head (each slot is sorted, so these form a prefix)
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int e = threadIdx.x + u * blockDim.x;
                        if (e >= ng || hk[u] == 0 || hk[u] < tkey) continue;
                        unsigned int pos = atomicAdd(&sh.extra_count, 1u);
                        if (pos >= (unsigned)kOutCap) {
                            sh.overflow = 1;
                            continue;
                        }
                        cand_put(bufA, (int)pos, hk[u], hm[u], hid[u]);
                        if ((e % mh) == mh - 1) {  // last gathered head of its slot still qualifies: walk on
                            const int sidx = e / mh;
                            for (int r = mh; r < CAP; r++) {
                                uint32_t k, m;
                                uint64_t id;
                                ld_cand(sidx, r, k, m, id);
                                if (k == 0 || k < tkey) break;
                                pos = atomicAdd(&sh.extra_count, 1u);
                                if (pos >= (unsigned)kOutCap) {
                                    sh.overflow = 1;
                                    break;
                                }
                                cand_put(bufA, (int)pos, k, m, id);
                            }
                        }
                    }
                    __syncthreads();
                    VS_TRACE(13);
                    slow = sh.overflow != 0;
                    if (!slow && !dedup) {
                        block_sort_small(bufA, (int)sh.extra_count, bufB, CAP);  // final list -> bufB[0..CAP)
                        res = bufB;
                    } else if (!slow) {
                        // everything collected in order, then one hit per document, then the cut
                        const int n = (int)sh.extra_count;
                        block_sort_small(bufA, n, bufB, n);
                        const int uniq = block_unique_compact(bufB, n, bufA, CAP, ss.scan_tmp);
                        res = bufA;
                        // duplicates across blocks left fewer than CAP documents although entries below the threshold
                        // were never looked at: take every slot in turn instead (exact)
                        slow = uniq < CAP && tkey != 0;
                    }
                }
                if (slow) {
                    // degenerate (massive ties or an unusually wide grid): warp 0 inserts every slot in turn
                    __syncthreads();
                    if (warp == 0) {
                        top.init();
                        for (int sidx = 0; sidx < nslots; sidx++) {
                            for (int c = 0; c < KPL; c++) {
                                uint32_t k, m;
                                uint64_t id;
                                ld_cand(sidx, c * 32 + lane, k, m, id);
                                bool any = __any_sync(FULL, k != 0 && cand_better(k, id, top.thr_key, top.thr_id));
                                if (!any) break;
                                top.offer(k != 0, k, m, id, lane, dedup);
                            }
                        }
                        top.store_soa(bufB, 0, lane);
                    }
                    res = bufB;
                    __syncthreads();
                }
            }
        }
        VS_TRACE(4);
        if (warp == 0) {
            top.load_soa(res, 0, lane);
            if (lane == 0) p.tickets[qslot] = 0;  // re-arm for the next launch
        }
        VS_TRACE(5);

        // Emit; first (non-EXACT) re-score flagged candidates inside the emit window with literal arithmetic.
        uint32_t status = 0;
        for (int pass = 0; pass < 2; pass++) {
            bool need_fix = false;
            if (warp == 0) {
                int outpos[KPL];
                // (every list holds distinct documents: search.go:260-268 happened on the way, nothing to remove here)
                int base = 0;
                bool anyflag = false;
                uint32_t kth_key = 0;
                uint64_t kth_id = kEmptyId;
#pragma unroll
                for (int s = 0; s < KPL; s++) {
                    bool keep = top.skey[s] != 0;
                    unsigned m = __ballot_sync(FULL, keep);
                    outpos[s] = base + __popc(m & ((1u << lane) - 1u));
                    // every entry ranked before the k-th one must be certain
                    anyflag |= __any_sync(FULL, top.skey[s] != 0 && outpos[s] < p.k && (top.meta[s] & kFlagBit));
                    unsigned mk = __ballot_sync(FULL, keep && outpos[s] == p.k - 1);
                    if (mk) {
                        kth_key = __shfl_sync(FULL, top.skey[s], __ffs(mk) - 1);
                        kth_id = __shfl_sync(FULL, top.id[s], __ffs(mk) - 1);
                    }
                    base += __popc(m);
                }
                const bool full = top.thr_key != 0;
                if (pass == 0 && anyflag && !EXACT) {
                    need_fix = true;
                    top.store_soa(bufA, 0, lane);
                    if (lane == 0) {
                        sh.overflow = 0;
                        sh.tail_key = top.thr_key;
                        sh.tail_id = top.thr_id;
                    }
                } else {
                    if (anyflag) status |= p.status_bit;  // (EXACT never flags; defensive)
                    if (pass == 1 && full && kth_key != 0 && cand_better(sh.tail_key, sh.tail_id, kth_key, kth_id))
                        status |= p.status_bit;  // a re-scored entry fell below rows that were not kept
                    if (pass == 1 && sh.overflow) status |= p.status_bit;
                    if (p.mode == 1) {
                        uint32_t mytiles = 0;
#pragma unroll
                        for (int s = 0; s < KPL; s++) {
                            int r = s * 32 + lane;
                            if (r < p.k && top.skey[s] != 0) {
                                uint32_t L = (uint32_t)top.id[s];
                                p.out_probe[(size_t)qi * p.k + r] = L;
                                if (p.out_sims) p.out_sims[(size_t)qi * p.k + r] = key_to_f32(top.skey[s]);
                                if (p.out_qtiles) {
                                    uint32_t len = (uint32_t)p.next_list_len[L];
                                    mytiles += (len + p.next_tile_rows - 1) / p.next_tile_rows;
                                }
                            }
                        }
                        if (p.out_qtiles) {
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) mytiles += __shfl_xor_sync(FULL, mytiles, o);
                            if (lane == 0) p.out_qtiles[qi] = mytiles ? mytiles : 1u;
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < KPL; s++) {
                            bool keep = top.skey[s] != 0;
                            if (keep && outpos[s] < p.k) {
                                p.out_ids[(size_t)qi * p.k + outpos[s]] = top.id[s];
                                p.out_sims[(size_t)qi * p.k + outpos[s]] = key_to_f32(top.skey[s]);
                            }
                        }
                        if (base < p.k && full) status |= kStatusNeedMore;
                        if (lane == 0) p.out_counts[qi] = min(base, p.k);
                    }
                    if (lane == 0 && p.out_status) {
                        if (EXACT) p.out_status[qi] = (p.out_status[qi] & ~p.status_bit) | (status & kStatusNeedMore);
                        else p.out_status[qi] = p.status_init ? status : (p.out_status[qi] | status);
                    }
                }
                if (lane == 0) sh.need_fix = need_fix ? 1u : 0u;
            }
            __syncthreads();
            VS_TRACE(6 + (pass ? 1 : 0));
            if (!sh.need_fix) break;
            // normalizeVector of the query (compute/cosine.go:26,138-149), literal: the elementwise work is spread
            // over the threads, only the additions are chained in element order (warp_ordered_sum).
            {
                const uint8_t *qc = p.queries.codes + (size_t)qi * d_pad;
                const float2 qh = p.queries.hdr[qi];
                const double mn = (double)qh.x, range = __dsub_rn((double)qh.y, (double)qh.x);
                for (int i = threadIdx.x; i < D; i += blockDim.x) sh_qn[i] = ref_dequant_f64(qc[i], mn, range);
                __syncthreads();
                if (warp == 0) {
                    const double nsq = warp_ordered_sum(D, lane, [&](int i) { return __dmul_rn(sh_qn[i], sh_qn[i]); });
                    if (lane == 0) sh.norm = __dsqrt_rn(nsq);
                }
                __syncthreads();
                const double norm = sh.norm;
                if (norm != 0.0)
                    for (int i = threadIdx.x; i < D; i += blockDim.x) sh_qn[i] = __ddiv_rn(sh_qn[i], norm);
                __syncthreads();
                // one warp per flagged candidate
                for (int e = warp; e < CAP; e += kStageWarps) {
                    const uint32_t k0 = bufA.key[e], m0 = bufA.meta[e];
                    if (k0 != 0 && (m0 & kFlagBit)) {
                        const size_t row = m0 & kMetaRowMask;
                        const float2 h = p.rows.hdr[row];
                        const double dot = warp_ref_cosine_row_f64(p.rows.codes + row * (size_t)d_pad, h.x, h.y, sh_qn, D, lane);
                        if (lane == 0) {
                            const uint32_t k1 = f32_to_key(__double2float_rn(dot));
                            bufA.key[e] = k1;
                            bufA.meta[e] = m0 & kMetaRowMask;
                            // an equally scored, equally uncertain row of the same document was dropped on the way: if this
                            // one's true score is the lower neighbour, the document's best hit is not known here
                            if ((m0 & kSibBit) && k1 != k0) sh.overflow = 1;
                            if (p.fix_counter) atomicAdd(p.fix_counter, 1ull);
                        }
                    }
                }
                __syncthreads();
                block_sort_small(bufA, CAP, bufB, CAP);  // entries only moved down by at most one float32 step
                if (warp == 0) top.load_soa(bufB, 0, lane);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
int stage_cap(int kpl) { return 32 * kpl; }

template <int G, int CPL, int KPL, bool EXACT>
static cudaError_t launch_stage_t(const StageParams &p, int grid_blocks, cudaStream_t st) {
    size_t smem = ((sizeof(StageShared) + 15) & ~size_t(15)) + ((sizeof(SortSmem) + 15) & ~size_t(15)) +
                  (size_t)p.rows.d * sizeof(double);
    if (p.qtiles) smem += (((size_t)p.nq + 1) * 4 + 15) & ~size_t(15);
    if (G == 0) smem += (size_t)p.rows.d_pad;
    auto kern = stage_kernel<G, CPL, KPL, EXACT>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (p.pdl == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid_blocks);
        cfg.blockDim = dim3(kStageWarps * 32);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kern, p);
    }
    kern<<<grid_blocks, kStageWarps * 32, smem, st>>>(p);
    return cudaGetLastError();
}

template <int KPL>
static cudaError_t launch_stage_k(const StageParams &p, bool exact, int grid_blocks, cudaStream_t st) {
    if (exact) return launch_stage_t<0, 0, KPL, true>(p, grid_blocks, st);
    const int ch = p.rows.d_pad >> 4;
    switch (ch) {
        case 48: return launch_stage_t<16, 3, KPL, false>(p, grid_blocks, st);  // 768-d (nomic-embed-text)
        case 32: return launch_stage_t<32, 1, KPL, false>(p, grid_blocks, st);  // 512-d (noop/ai.go)
        case 64: return launch_stage_t<32, 2, KPL, false>(p, grid_blocks, st);  // 1024-d
        case 96: return launch_stage_t<32, 3, KPL, false>(p, grid_blocks, st);  // 1536-d
        case 24: return launch_stage_t<8, 3, KPL, false>(p, grid_blocks, st);   // 384-d
        default: return launch_stage_t<0, 0, KPL, false>(p, grid_blocks, st);
    }
}

// Lanes per row for this row width (0 = generic routine): decides the tile heights the host may choose.
int stage_lanes_per_row(int d_pad) {
    switch (d_pad >> 4) {
        case 48: return 16;
        case 32: case 64: case 96: return 32;
        case 24: return 8;
        default: return 0;
    }
}

cudaError_t launch_stage(const StageParams &p, int kpl, bool exact, int grid_blocks, cudaStream_t st) {
    if (p.nq > kMaxStageQueries) return cudaErrorInvalidValue;
    switch (kpl) {
        case 1: return launch_stage_k<1>(p, exact, grid_blocks, st);
        case 2: return launch_stage_k<2>(p, exact, grid_blocks, st);
        case 4: return launch_stage_k<4>(p, exact, grid_blocks, st);
        default: return cudaErrorInvalidValue;
    }
}

// ---------------------------------------------------------------------------------------------------
// normalizeVector of each query (compute/cosine.go:26,138-149) in literal float64: one thread sums
// sequentially, then the block divides.
__global__ void query_normalize_kernel(MatView q, double *qnorm, const unsigned int *only_if_nonzero) {
    __shared__ double s_norm;
    if (only_if_nonzero && *only_if_nonzero == 0) return;  // nothing will read the result (empty literal-path worklist)
    const int qi = blockIdx.x;
    const int D = q.d;
    const uint8_t *codes = q.codes + (size_t)qi * q.d_pad;
    const float2 h = q.hdr[qi];
    const double mn = (double)h.x, range = __dsub_rn((double)h.y, (double)h.x);
    if (threadIdx.x == 0) {
        double norm = 0.0;
        for (int i = 0; i < D; i++) {
            double x = ref_dequant_f64(codes[i], mn, range);
            norm = __dadd_rn(norm, __dmul_rn(x, x));
        }
        s_norm = __dsqrt_rn(norm);
    }
    __syncthreads();
    const double norm = s_norm;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        double x = ref_dequant_f64(codes[i], mn, range);
        if (norm != 0.0) x = __ddiv_rn(x, norm);
        qnorm[(size_t)qi * D + i] = x;
    }
}

cudaError_t launch_query_normalize(const MatView &queries, double *qnorm, cudaStream_t st, const unsigned int *only_if_nonzero) {
    query_normalize_kernel<<<(unsigned)queries.n, 128, 0, st>>>(queries, qnorm, only_if_nonzero);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Full 1xN scores (compute/cosine.go:13-57 as an API: sims[i] for every row).
template <int G, int CPL>
__global__ void __launch_bounds__(256, 3)
cosine_1xN_kernel(MatView rows, MatView query, float *sims, uint32_t *dots, uint32_t *worklist,
                  unsigned int *work_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *sh_q = reinterpret_cast<uint4 *>(smem_raw);
    __shared__ SideConst s_q;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = rows.d, d_pad = rows.d_pad;
    if (threadIdx.x == 0) {
        float2 h = query.hdr[0];
        uint2 s = query.sums[0];
        s_q = make_side(h.x, h.y, s.x, s.y, D);
    }
    if (G == 0) {
        const uint4 *qsrc = reinterpret_cast<const uint4 *>(query.codes);
        for (int c = threadIdx.x; c < (d_pad >> 4); c += blockDim.x) sh_q[c] = qsrc[c];
    }
    __syncthreads();
    const SideConst xq = s_q;
    uint4 qreg[CPL > 0 ? CPL : 1];
    if constexpr (G != 0) {
#pragma unroll
        for (int j = 0; j < CPL; j++) qreg[j] = *reinterpret_cast<const uint4 *>(query.codes + ((lane % G) + G * j) * 16);
    }
    const size_t ntiles = (rows.n + kTileRows - 1) / kTileRows;
    const size_t wstride = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t t = (size_t)blockIdx.x * (blockDim.x >> 5) + warp; t < ntiles; t += wstride) {
        const size_t row0 = t * kTileRows;
        const int nrows = (int)min((size_t)kTileRows, rows.n - row0);
        uint32_t mydot;
        if constexpr (G != 0) mydot = tile_dots<G, CPL>(rows.codes, row0, nrows, d_pad, qreg, lane, G);
        else mydot = tile_dots_generic(rows.codes, row0, nrows, d_pad, sh_q, lane);
        if (lane < nrows) {
            const size_t row = row0 + lane;
            if (dots) dots[row] = mydot;
            if (sims) {
                float2 h = rows.hdr[row];
                uint2 s = rows.sums[row];
                bool flag;
                float sim = score_fast(xq, h.x, h.y, s.x, s.y, mydot, D, &flag);
                sims[row] = sim;
                if (flag) worklist[atomicAdd(work_count, 1u)] = (uint32_t)row;
            }
        }
    }
}

cudaError_t launch_cosine_1xN(const MatView &rows, const MatView &query, float *sims, uint32_t *dots,
                              uint32_t *worklist, unsigned int *work_count, int sm_count, cudaStream_t st) {
    const size_t ntiles = (rows.n + kTileRows - 1) / kTileRows;
    size_t blocks = (ntiles + 7) / 8;
    const size_t maxb = (size_t)sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    const int ch = rows.d_pad >> 4;
    switch (ch) {
        case 48: cosine_1xN_kernel<16, 3><<<(unsigned)blocks, 256, 0, st>>>(rows, query, sims, dots, worklist, work_count); break;
        case 32: cosine_1xN_kernel<32, 1><<<(unsigned)blocks, 256, 0, st>>>(rows, query, sims, dots, worklist, work_count); break;
        case 64: cosine_1xN_kernel<32, 2><<<(unsigned)blocks, 256, 0, st>>>(rows, query, sims, dots, worklist, work_count); break;
        case 96: cosine_1xN_kernel<32, 3><<<(unsigned)blocks, 256, 0, st>>>(rows, query, sims, dots, worklist, work_count); break;
        case 24: cosine_1xN_kernel<8, 3><<<(unsigned)blocks, 256, 0, st>>>(rows, query, sims, dots, worklist, work_count); break;
        default: cosine_1xN_kernel<0, 0><<<(unsigned)blocks, 256, rows.d_pad, st>>>(rows, query, sims, dots, worklist, work_count); break;
    }
    return cudaGetLastError();
}

// Rows whose float32 rounding could not be certified: literal reference arithmetic, one warp per row.
__global__ void cosine_fix_kernel(MatView rows, const double *qnorm, float *sims, const uint32_t *worklist,
                                  const unsigned int *work_count) {
    const unsigned int n = *work_count;
    const int lane = threadIdx.x & 31;
    const unsigned int nwarps = gridDim.x * (blockDim.x >> 5);
    for (unsigned int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarps) {
        const uint32_t row = worklist[i];
        const float2 h = rows.hdr[row];
        const double dot = warp_ref_cosine_row_f64(rows.codes + (size_t)row * rows.d_pad, h.x, h.y, qnorm, rows.d, lane);
        if (lane == 0) sims[row] = __double2float_rn(dot);
    }
}

cudaError_t launch_cosine_fix(const MatView &rows, const double *qnorm, float *sims, const uint32_t *worklist,
                              const unsigned int *work_count, int sm_count, cudaStream_t st) {
    cosine_fix_kernel<<<sm_count * 4, 128, 0, st>>>(rows, qnorm, sims, worklist, work_count);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Multi-GPU merge: G shard-local hit lists per query -> global top-k (same order and dedup rule).  Shard g's arrays are
// either at a fixed stride from one base (one process per GPU: the all-gathered buffer) or anywhere (`bufs`: one packed
// buffer per shard, possibly in a PEER device's memory -- one process driving G devices reads them over NVLink).
//
// One process per GPU without a collective (`xs`): the shards' buffers are the PEER PROCESSES' memory, mapped through CUDA
// IPC, and the exchange is part of this kernel.  Every rank's hits are complete when its merge kernel starts (stream
// order), so the kernel first SIGNALS -- a system-scope release store of the step number into every peer's flag word for
// this rank, over NVLink -- and every block then WAITS (acquire loads of its own flag words) until all ranks have
// signalled this step, and only then reads the peers' hits in place.  No NCCL launch, no gathered copy: the transfer is
// the merge's own loads.  (The waiting blocks hold a warp each; nothing on this device depends on them.)
struct MergeExchange {
    uint32_t *const *signal;     // [G] address of this rank's flag word in every rank's memory (device array), or null
    const uint32_t *wait;        // [G] this rank's own flag words
    uint32_t step;               // the step being exchanged (flags only grow)
};
__global__ void topk_merge_kernel(const uint64_t *ids_in, const float *sims_in, const int32_t *counts_in,
                                  size_t rank_stride_bytes, const unsigned char *const *bufs, size_t ids_off, size_t sims_off,
                                  size_t counts_off, int G, int nq, int k, uint64_t *ids_out, float *sims_out,
                                  int32_t *counts_out, MergeExchange xs) {
    const int qi = blockIdx.x;
    const int lane = threadIdx.x;
    if (xs.signal || xs.wait) {
        if (xs.signal && qi == 0 && lane < G) {
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(xs.signal[lane]), "r"(xs.step) : "memory");
        }
        if (xs.wait && lane < G) {
            uint32_t seen;
            long long t0 = 0;
            do {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(xs.wait + lane) : "memory");
                if ((int32_t)(seen - xs.step) < 0) {  // a peer that never arrives must fault, not hang the device: ~10 s
                    const long long now = clock64();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > 20000000000ll) __trap();
                }
            } while ((int32_t)(seen - xs.step) < 0);
        }
        __syncwarp();
    }
    WarpTopK<4> top;
    top.init();
    for (int g = 0; g < G; g++) {
        const uint64_t *gi;
        const float *gs;
        const int32_t *gc;
        if (bufs) {
            const unsigned char *b = bufs[g];
            gi = reinterpret_cast<const uint64_t *>(b + ids_off);
            gs = reinterpret_cast<const float *>(b + sims_off);
            gc = reinterpret_cast<const int32_t *>(b + counts_off);
        } else if (rank_stride_bytes) {
            gi = reinterpret_cast<const uint64_t *>(reinterpret_cast<const char *>(ids_in) + g * rank_stride_bytes);
            gs = reinterpret_cast<const float *>(reinterpret_cast<const char *>(sims_in) + g * rank_stride_bytes);
            gc = reinterpret_cast<const int32_t *>(reinterpret_cast<const char *>(counts_in) + g * rank_stride_bytes);
        } else {
            gi = ids_in + (size_t)g * nq * k;
            gs = sims_in + (size_t)g * nq * k;
            gc = counts_in + (size_t)g * nq;
        }
        const int cnt = __ldcg(gc + qi);  // (L2 reads: a peer's buffer changes between launches, and within one after the wait)
        for (int base = 0; base < cnt; base += 32) {
            int j = base + lane;
            bool valid = j < cnt;
            size_t off = (size_t)qi * k + (valid ? j : 0);
            uint32_t key = valid ? f32_to_key(__ldcg(gs + off)) : 0u;
            uint64_t id = valid ? __ldcg(gi + off) : kEmptyId;
            top.offer(valid, key, 0u, id, lane, true);  // one hit per document across shards, before the cut
        }
    }
    int base = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        bool keep = top.skey[s] != 0;
        unsigned m = __ballot_sync(FULL, keep);
        int outpos = base + __popc(m & ((1u << lane) - 1u));
        if (keep && outpos < k) {
            ids_out[(size_t)qi * k + outpos] = top.id[s];
            sims_out[(size_t)qi * k + outpos] = key_to_f32(top.skey[s]);
        }
        base += __popc(m);
    }
    if (lane == 0) counts_out[qi] = min(base, k);
}

cudaError_t launch_topk_merge(const uint64_t *ids_in, const float *sims_in, const int32_t *counts_in, size_t rank_stride_bytes,
                              int G, int nq, int k, uint64_t *ids_out, float *sims_out, int32_t *counts_out, cudaStream_t st) {
    if (k > 128) return cudaErrorInvalidValue;
    topk_merge_kernel<<<nq, 32, 0, st>>>(ids_in, sims_in, counts_in, rank_stride_bytes, nullptr, 0, 0, 0, G, nq, k, ids_out, sims_out,
                                         counts_out, MergeExchange{nullptr, nullptr, 0u});
    return cudaGetLastError();
}

cudaError_t launch_topk_merge_ptrs(const unsigned char *const *bufs, size_t ids_off, size_t sims_off, size_t counts_off, int G, int nq,
                                   int k, uint64_t *ids_out, float *sims_out, int32_t *counts_out, cudaStream_t st) {
    if (k > 128) return cudaErrorInvalidValue;
    topk_merge_kernel<<<nq, 32, 0, st>>>(nullptr, nullptr, nullptr, 0, bufs, ids_off, sims_off, counts_off, G, nq, k, ids_out, sims_out,
                                         counts_out, MergeExchange{nullptr, nullptr, 0u});
    return cudaGetLastError();
}

cudaError_t launch_topk_merge_exchange(const unsigned char *const *bufs, size_t ids_off, size_t sims_off, size_t counts_off, int G, int nq,
                                       int k, uint64_t *ids_out, float *sims_out, int32_t *counts_out, uint32_t *const *signal,
                                       const uint32_t *wait, uint32_t step, cudaStream_t st) {
    if (k > 128 || G > 32) return cudaErrorInvalidValue;
    topk_merge_kernel<<<nq, 32, 0, st>>>(nullptr, nullptr, nullptr, 0, bufs, ids_off, sims_off, counts_off, G, nq, k, ids_out, sims_out,
                                         counts_out, MergeExchange{signal, wait, step});
    return cudaGetLastError();
}

// The same exchange with the waiting done by ONE warp in a launch of its own, the merge behind it in stream order: a rank
// that arrives early then holds one warp instead of one per query (the default, see vs_exchange_merge).
__global__ void exchange_wait_kernel(int G, MergeExchange xs) {
    const int lane = threadIdx.x;
    if (lane < G) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(xs.signal[lane]), "r"(xs.step) : "memory");
        uint32_t seen;
        long long t0 = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(xs.wait + lane) : "memory");
            if ((int32_t)(seen - xs.step) < 0) {
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 20000000000ll) __trap();
            }
        } while ((int32_t)(seen - xs.step) < 0);
    }
}
cudaError_t launch_topk_merge_exchange_split(const unsigned char *const *bufs, size_t ids_off, size_t sims_off, size_t counts_off, int G,
                                             int nq, int k, uint64_t *ids_out, float *sims_out, int32_t *counts_out,
                                             uint32_t *const *signal, const uint32_t *wait, uint32_t step, cudaStream_t st) {
    if (k > 128 || G > 32) return cudaErrorInvalidValue;
    exchange_wait_kernel<<<1, 32, 0, st>>>(G, MergeExchange{signal, wait, step});
    topk_merge_kernel<<<nq, 32, 0, st>>>(nullptr, nullptr, nullptr, 0, bufs, ids_off, sims_off, counts_off, G, nq, k, ids_out, sims_out,
                                         counts_out, MergeExchange{nullptr, nullptr, 0u});
    return cudaGetLastError();
}

cudaError_t scan_set_certify_scale(float scale) { return cudaMemcpyToSymbol(c_certify_scale, &scale, sizeof(float)); }

}  // namespace vs
