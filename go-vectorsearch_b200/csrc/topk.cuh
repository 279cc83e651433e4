// topk.cuh -- the fused top-k machinery shared by the scan kernels (scan.cu, fused.cu, listmajor.cu).
//
// Replaces server/search.go:250-270: append the batch's hits, sort by similarity (desc), keep ONE hit per
// document (the best), truncate to Count+Offset.  The reference removes duplicates BEFORE it truncates, so every
// list kept here holds distinct documents (`dedup` = the store carries explicit document ids): a warp list rejects
// or replaces on insert, and every merge removes duplicates before it cuts to its capacity.  A list cut to its
// capacity after de-duplication is the exact "best `cap` distinct documents" of what it has seen, and the best
// `cap` distinct documents of a union are among the best `cap` distinct documents of its parts.
#pragma once
#include "common.cuh"

namespace vs {

#ifndef FULL
#define FULL 0xFFFFFFFFu
#endif

// ---------------------------------------------------------------------------------------------------
// Tile dot products. A tile is NG*iters consecutive rows (NG = 32/G lane groups, `iters` rows per group):
// group g streams rows g*iters+it. Returns in lane g*G+it (it < iters) the integer dot of that row.
template <int G, int CPL>
__device__ __forceinline__ uint32_t tile_dots(const uint8_t *__restrict__ codes, size_t row0, int nrows, int d_pad,
                                              const uint4 (&q)[CPL], int lane, int iters) {
    constexpr int U = (CPL <= 3) ? 4 : 2;  // rows in flight per lane group
    const int g = lane / G, l = lane % G;
    uint32_t mydot = 0;
#pragma unroll 1
    for (int it0 = 0; it0 < iters; it0 += U) {
        uint4 v[U][CPL];
#pragma unroll
        for (int u = 0; u < U; u++) {
            int r = g * iters + it0 + u;
            bool ok = (it0 + u < iters) && (r < nrows);
            const uint8_t *p = codes + (row0 + (size_t)(ok ? r : 0)) * (size_t)d_pad + l * 16;
#pragma unroll
            for (int j = 0; j < CPL; j++) v[u][j] = ok ? ld_stream_u4(p + j * G * 16) : make_uint4(0, 0, 0, 0);
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

// ---------------------------------------------------------------------------------------------------
// Warp-distributed sorted top list: rank r lives in slot r/32 of lane r%32, best first.
template <int KPL>
struct WarpTopK {
    uint32_t skey[KPL];
    uint32_t meta[KPL];
    uint64_t id[KPL];
    uint32_t thr_key;  // worst kept entry (rank 32*KPL-1), warp-uniform
    uint64_t thr_id;

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < KPL; s++) {
            skey[s] = 0;
            meta[s] = 0;
            id[s] = kEmptyId;
        }
        thr_key = 0;
        thr_id = kEmptyId;
    }

    // Insert one (warp-uniform) candidate.  dedup: an entry of the same document already in the list either keeps its
    // place (it is at least as good) or is dropped in favour of the candidate (search.go:260-268).
    __device__ __forceinline__ void insert(uint32_t ck, uint32_t cm, uint64_t cid, int lane, bool dedup = false) {
        int drop = 32 * KPL - 1;  // the rank that leaves the list: the last one, or the same document's older entry
        if (dedup) {
#pragma unroll
            for (int s = 0; s < KPL; s++) {
                const unsigned m = __ballot_sync(FULL, skey[s] != 0 && id[s] == cid);
                if (m) {
                    const int l0 = __ffs(m) - 1;
                    const uint32_t ek = __shfl_sync(FULL, skey[s], l0);
                    if (ek > ck) return;  // same document, kept entry better
                    if (ek == ck) {       // equal float32 scores: the certain row represents the document
                        const uint32_t em = __shfl_sync(FULL, meta[s], l0);
                        if ((em & kFlagBit) && lane == l0) meta[s] = (cm & kFlagBit) ? (em | kSibBit) : cm;
                        return;
                    }
                    drop = s * 32 + l0;
                }
            }
        }
        int pos = 0;
#pragma unroll
        for (int s = 0; s < KPL; s++) {
            bool ahead = !cand_better(ck, cid, skey[s], id[s]);
            pos += __popc(__ballot_sync(FULL, ahead));
        }
        if (pos > drop) return;  // (drop is the last rank unless a worse entry of the same document was found, which
                                 // sorts after the candidate)
        uint32_t ck_prev = 0, cm_prev = 0;
        uint64_t cid_prev = 0;
        const int src = (lane + 31) & 31;
#pragma unroll
        for (int s = 0; s < KPL; s++) {
            uint32_t rk = __shfl_sync(FULL, skey[s], src);
            uint32_t rm = __shfl_sync(FULL, meta[s], src);
            uint64_t rid = __shfl_sync(FULL, id[s], src);
            uint32_t sk = lane == 0 ? ck_prev : rk;
            uint32_t sm = lane == 0 ? cm_prev : rm;
            uint64_t sid = lane == 0 ? cid_prev : rid;
            int r = s * 32 + lane;
            if (r == pos) {
                skey[s] = ck;
                meta[s] = cm;
                id[s] = cid;
            } else if (r > pos && r <= drop) {
                skey[s] = sk;
                meta[s] = sm;
                id[s] = sid;
            }
            ck_prev = rk;  // on lane 0: the old rank 32*s+31, which moves to rank 32*(s+1)
            cm_prev = rm;
            cid_prev = rid;
        }
        thr_key = __shfl_sync(FULL, skey[KPL - 1], 31);
        thr_id = __shfl_sync(FULL, id[KPL - 1], 31);
    }

    // Offer one candidate per lane.
    __device__ __forceinline__ void offer(bool valid, uint32_t ck, uint32_t cm, uint64_t cid, int lane, bool dedup = false) {
        bool pass = valid && cand_better(ck, cid, thr_key, thr_id);
        unsigned m = __ballot_sync(FULL, pass);
        while (m) {
            int src = __ffs(m) - 1;
            m &= m - 1;
            uint32_t bk = __shfl_sync(FULL, ck, src);
            uint32_t bm = __shfl_sync(FULL, cm, src);
            uint64_t bid = __shfl_sync(FULL, cid, src);
            insert(bk, bm, bid, lane, dedup);
        }
    }

    __device__ __forceinline__ int count() const {
        int cnt = 0;
#pragma unroll
        for (int s = 0; s < KPL; s++) cnt += __popc(__ballot_sync(FULL, skey[s] != 0));
        return cnt;
    }

    template <typename B>
    __device__ __forceinline__ void store_soa(B &b, int base, int lane) const {
#pragma unroll
        for (int s = 0; s < KPL; s++) {
            b.key[base + s * 32 + lane] = skey[s];
            b.meta[base + s * 32 + lane] = meta[s];
            b.id[base + s * 32 + lane] = id[s];
        }
    }
    template <typename B>
    __device__ __forceinline__ void load_soa(const B &b, int base, int lane) {
#pragma unroll
        for (int s = 0; s < KPL; s++) {
            skey[s] = b.key[base + s * 32 + lane];
            meta[s] = b.meta[base + s * 32 + lane];
            id[s] = b.id[base + s * 32 + lane];
        }
        thr_key = __shfl_sync(0xFFFFFFFFu, skey[KPL - 1], 31);
        thr_id = __shfl_sync(0xFFFFFFFFu, id[KPL - 1], 31);
    }
};

// ---------------------------------------------------------------------------------------------------
// Block-wide merge machinery.  Candidates live in shared memory as three arrays (key, meta, id).  A
// warp-shuffle insertion costs ~100 cycles per candidate and a barrier-per-step bitonic sort ~400 cycles
// per step, so ordering is done by RANK instead: every thread computes the final position of its own
// candidate (binary searches against the other sorted lists, or a count over an unsorted set) and
// scatters it -- no barriers inside, one at the end.
constexpr int kSortCap = 1024;  // input candidates per merge (>= warps per block * 128)
constexpr int kOutCap = 1024;   // output / collected-candidates buffer

struct CandBuf {
    uint32_t *key;
    uint32_t *meta;
    uint64_t *id;
};
struct SortSmem {
    uint64_t id_a[kSortCap];
    uint64_t id_b[kOutCap];
    uint32_t key_a[kSortCap];
    uint32_t meta_a[kSortCap];
    uint32_t key_b[kOutCap];
    uint32_t meta_b[kOutCap];
    int scan_tmp[33];
};

__device__ __forceinline__ void cand_put(const CandBuf &b, int i, uint32_t k, uint32_t m, uint64_t id) {
    b.key[i] = k;
    b.meta[i] = m;
    b.id[i] = id;
}

// e precedes f in the merged order; equal (key,id) pairs are ordered by (list, position) to keep ranks unique.
__device__ __forceinline__ bool cand_before(uint32_t ek, uint64_t eid, int eorder, uint32_t fk, uint64_t fid, int forder) {
    return ek > fk || (ek == fk && (eid < fid || (eid == fid && eorder < forder)));
}

// Merge nl sorted lists (list l occupies src[l*stride .. l*stride+len[l]) or, when len == nullptr, `stride`
// entries each) into dst[0..outcap) best-first.  One thread per input entry; ends with a barrier.
__device__ __forceinline__ void rank_merge(const CandBuf &src, int nl, int stride, const int *len, const CandBuf &dst,
                                           int outcap) {
    const int total = nl * stride;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int l = e / stride, pos = e - l * stride;
        const int mylen = len ? len[l] : stride;
        if (pos >= mylen) continue;
        const uint32_t k = src.key[e];
        const uint64_t id = src.id[e];
        int rank = pos;
        for (int o = 0; o < nl && rank < outcap; o++) {
            if (o == l) continue;
            const int base = o * stride;
            int lo = 0, hi = len ? len[o] : stride;
            while (lo < hi) {  // number of entries of list o that precede (k,id,l)
                const int mid = (lo + hi) >> 1;
                if (cand_before(src.key[base + mid], src.id[base + mid], o, k, id, l)) lo = mid + 1;
                else hi = mid;
            }
            rank += lo;
        }
        if (rank < outcap) cand_put(dst, rank, k, src.meta[e], id);
    }
    __syncthreads();
}

// Warp bitonic sort of 32 candidates held one per lane (registers + shuffles, no shared memory).
__device__ __forceinline__ void warp_sort32(uint32_t &k, uint32_t &m, uint64_t &id, int lane) {
#pragma unroll
    for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            const uint32_t ok = __shfl_xor_sync(FULL, k, j);
            const uint32_t om = __shfl_xor_sync(FULL, m, j);
            const uint64_t oid = __shfl_xor_sync(FULL, id, j);
            const bool up = (lane & kk) == 0;        // this run ends up best-first
            const bool lower = (lane & j) == 0;      // I hold the lower index of the pair
            const bool want_better = (up == lower);  // the better of the two belongs here
            const bool other_better = cand_better(ok, oid, k, id);
            if (other_better == want_better && !(ok == k && oid == id)) {
                k = ok;
                m = om;
                id = oid;
            }
        }
    }
}

// Order an unsorted set src[0..n) (n <= kOutCap, src has room for n rounded up to 32) into dst[0..outcap)
// best-first: every warp sorts 32-entry chunks in registers, then the chunks are rank-merged.  Barriers inside.
// `nthreads` threads of the block take part (the first ones; all of them must call).
__device__ __forceinline__ void block_sort_small(const CandBuf &src, int n, const CandBuf &dst, int outcap) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int nch = (n + 31) >> 5;
    for (int ch = warp; ch < nch; ch += nwarp) {
        const int e = ch * 32 + lane;
        uint32_t k = 0, m = 0;
        uint64_t id = kEmptyId;
        if (e < n) {
            k = src.key[e];
            m = src.meta[e];
            id = src.id[e];
        }
        warp_sort32(k, m, id, lane);
        cand_put(src, e, k, m, id);
    }
    for (int e = threadIdx.x; e < outcap; e += blockDim.x) cand_put(dst, e, 0u, 0u, kEmptyId);
    __syncthreads();
    rank_merge(src, nch, 32, nullptr, dst, outcap);
}

// One hit per document over a sorted list: srt[0..n) best-first (empty entries, key 0, only at the end; n <= 32 *
// blockDim.x) -> the first `outcap` entries whose document id does not occur earlier, compacted into dst[0..outcap)
// (the rest of dst emptied).  Returns the number of distinct documents among ALL n entries (block-uniform).
// scan_tmp: int[33].  Barriers inside; srt's meta words may be updated (see kSibBit).
__device__ __forceinline__ int block_unique_compact(const CandBuf &srt, int n, const CandBuf &dst, int outcap, int *scan_tmp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int e = threadIdx.x; e < outcap; e += blockDim.x) cand_put(dst, e, 0u, 0u, kEmptyId);
    if (threadIdx.x == 0) scan_tmp[32] = 0;
    const int rounds = (n + (int)blockDim.x - 1) / (int)blockDim.x;
    unsigned keepmask = 0;  // bit r: my entry of round r is its document's first (best) one
    for (int r = 0; r < rounds; r++) {
        const int e = r * (int)blockDim.x + (int)threadIdx.x;
        if (e >= n) break;
        const uint32_t k = srt.key[e];
        if (k == 0) continue;
        const uint64_t id = srt.id[e];
        bool keep = true;
        for (int j = 0; j < e; j++) {
            if (srt.id[j] != id) continue;
            keep = false;
            if (srt.key[j] == k) {  // equal float32 scores: the certain row represents the document
                const uint32_t mj = srt.meta[j], me = srt.meta[e];
                if (mj & kFlagBit) {
                    if (me & kFlagBit) atomicOr(&srt.meta[j], kSibBit);
                    else srt.meta[j] = me;
                }
            }
            break;
        }
        if (keep) keepmask |= 1u << r;
    }
    __syncthreads();
    for (int r = 0; r < rounds; r++) {
        const int e = r * (int)blockDim.x + (int)threadIdx.x;
        const bool keep = (keepmask >> r) & 1u;
        const unsigned bal = __ballot_sync(FULL, keep);
        if (lane == 0) scan_tmp[warp] = __popc(bal);
        __syncthreads();
        int base = scan_tmp[32];
        for (int w = 0; w < warp; w++) base += scan_tmp[w];
        const int pos = base + __popc(bal & ((1u << lane) - 1u));
        if (keep && pos < outcap) cand_put(dst, pos, srt.key[e], srt.meta[e], srt.id[e]);
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = scan_tmp[32];
            for (int w = 0; w < nwarp; w++) t += scan_tmp[w];
            scan_tmp[32] = t;
        }
        __syncthreads();
    }
    return scan_tmp[32];
}

// A small unsorted set (n <= blockDim.x) ordered by COUNTING, with the whole block at work: a lone warp runs a dependent
// instruction chain at ~5 cycles per instruction, so each candidate gets `tpe` adjacent lanes (as many as the block
// affords, up to 32); every lane compares the candidate with a strided share of the set, branch-free, and the lanes add
// their counts with shuffles.  With dedup a first pass finds, the same way, whether a better row of the same document
// exists (one hit per document, search.go:260-268; equal scores: see kSibBit).
// src[0..n) -> dst[0..outcap) best-first (empty beyond).  Returns the number of distinct documents.  src's keys of
// dropped candidates are zeroed.  Barriers inside; all threads call.
static __device__ __forceinline__ int block_rank_small(CandBuf src, int n, CandBuf dst, int outcap, bool dedup) {
    int lg = 5;  // tpe = 1 << lg lanes per candidate (shifts, not divisions: a division is ~50 dependent instructions)
    while (lg > 0 && (n << lg) > (int)blockDim.x) lg--;
    const int tpe = 1 << lg;
    const int e = (int)threadIdx.x >> lg, part = (int)threadIdx.x & (tpe - 1);
    for (int i = threadIdx.x; i < outcap; i += blockDim.x) cand_put(dst, i, 0u, 0u, kEmptyId);
    const bool mine = e < n;
    const uint32_t k = mine ? src.key[e] : 0u;
    const uint64_t id = mine ? src.id[e] : kEmptyId;
    bool live = k != 0;
    if (dedup) {
        // the best row of the same document that precedes this one (score desc, then position), packed for a max-reduce
        unsigned long long best = 0;
        for (int f = part; f < n; f += tpe) {
            const uint32_t fk = src.key[f];
            const uint64_t fid = src.id[f];
            const bool hit = live & (fid == id) & ((fk > k) | ((fk == k) & (f < e)));
            const unsigned long long cand = ((unsigned long long)fk << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)f);
            best = hit && cand > best ? cand : best;
        }
        for (int o = 1; o < tpe; o <<= 1) {
            const unsigned long long other = __shfl_xor_sync(FULL, best, o);
            best = other > best ? other : best;
        }
        if (best != 0) {
            live = false;
            const int f = (int)(0xFFFFFFFFu - (uint32_t)best);
            if (part == 0 && (uint32_t)(best >> 32) == k) {  // equal float32 scores: the certain row represents the document
                const uint32_t mf = src.meta[f], me = src.meta[e];
                if (mf & kFlagBit) {
                    if (me & kFlagBit) atomicOr(&src.meta[f], kSibBit);
                    else src.meta[f] = me;
                }
            }
        }
        __syncthreads();
        if (mine && !live && part == 0) src.key[e] = 0;  // dropped: out of the ranking below
    }
    const int uniq = __syncthreads_count(live && part == 0);
    int rank = 0;
    for (int f = part; f < n; f += tpe) {
        const uint32_t fk = src.key[f];
        const uint64_t fid = src.id[f];
        rank += (int)((fk > k) | ((fk == k) & ((fid < id) | ((fid == id) & (f < e)))));
    }
    for (int o = 1; o < tpe; o <<= 1) rank += __shfl_xor_sync(FULL, rank, o);
    if (live && part == 0 && rank < outcap) cand_put(dst, rank, k, src.meta[e], id);
    __syncthreads();
    return uniq;
}

// nl sorted lists (of distinct documents each, when dedup) in src -> the best `outcap` (distinct documents) of their
// union, best-first; entries beyond the available ones are empty.  tmp: a second buffer with room for nl*stride entries.
// Returns the buffer that holds the result: tmp (plain rank merge cut to outcap) or src (dedup: the whole union is
// ordered into tmp, then one hit per document is compacted back into src).  Barriers inside; all threads must call.
__device__ __forceinline__ CandBuf merge_lists(const CandBuf &src, int nl, int stride, const int *len, const CandBuf &tmp,
                                               int outcap, bool dedup, int *scan_tmp) {
    const int total = dedup ? nl * stride : outcap;
    for (int e = threadIdx.x; e < total; e += blockDim.x) cand_put(tmp, e, 0u, 0u, kEmptyId);
    __syncthreads();
    rank_merge(src, nl, stride, len, tmp, total);
    if (!dedup) return tmp;
    block_unique_compact(tmp, total, src, outcap, scan_tmp);
    return src;
}

}  // namespace vs
