// probe.cu -- probe selection for a batch of queries (server/search.go:202-227 for nq queries at once).
//
// The streaming stage kernel (scan.cu) re-reads the centroid table once per query; for a batch that is nq x C x 776 B
// of L2 traffic (203 MB at 64 queries x 4096 centroids) for a table of 3 MB.  Here the table is read once: a warp keeps
// eight centroid rows in registers and takes u8 dot products (dp4a) against every query of a shared-memory chunk, the
// certified float32 similarity (common.cuh) of every (query, centroid) pair is written as an order-preserving key, and
// a second kernel selects each query's nprobe best keys (similarity desc, centroid index asc) with a radix select.
// Pairs whose float32 rounding could not be certified are listed per query; the few of them that can reach the probe
// window are re-scored with literal reference arithmetic before the selection is accepted.
#include "internal.h"

namespace vs {
namespace {

#define FULL 0xFFFFFFFFu
constexpr int kPsWarps = 4;                          // warps per block of the score kernel
constexpr int kPsRows = kPsWarps * 8;                // centroid rows per work item (2 half-warps x 4 rows per warp)
// queries staged in shared memory per work item: 8 for small batches (the whole stage is a few microseconds of work,
// spread it over every SM), 32 for large ones (a row tile is loaded once per 32 queries instead of once per 8)
constexpr int kSelThreadsP = 256;
constexpr int kMaxProbe = 128;

// CPL = 16-byte chunks per lane; 16 lanes stream one row: d_pad = 256 * CPL bytes.
template <int CPL, int kPsQueries>
__global__ void __launch_bounds__(kPsWarps * 32)
probe_score_kernel(MatView cent, MatView queries, uint32_t *__restrict__ keys, size_t key_stride,
                   unsigned int *__restrict__ flag_cnt, uint32_t *__restrict__ flag_list, uint32_t flag_cap) {
    extern __shared__ __align__(16) unsigned char ps_smem[];
    constexpr int CH = 16 * CPL;  // 16-byte chunks per row
    uint4 *sh_q = reinterpret_cast<uint4 *>(ps_smem);                                       // [kPsQueries][CH]
    SideConst *sh_side = reinterpret_cast<SideConst *>(ps_smem + (size_t)kPsQueries * CH * 16);  // [kPsQueries]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hw = lane >> 4, l = lane & 15;
    const int D = cent.d, d_pad = cent.d_pad;
    const uint32_t C = (uint32_t)cent.n, nq = (uint32_t)queries.n;
    const uint32_t row_tiles = (C + kPsRows - 1) / kPsRows, q_chunks = (nq + kPsQueries - 1) / kPsQueries;
    for (uint32_t item = blockIdx.x; item < row_tiles * q_chunks; item += gridDim.x) {
        const uint32_t rt = item % row_tiles, qc = item / row_tiles;
        const uint32_t q0 = qc * kPsQueries;
        const int nqc = (int)min((uint32_t)kPsQueries, nq - q0);
        // the centroid rows do not depend on the staged queries: their loads fly while the chunk is staged
        const uint32_t row_base = rt * kPsRows + warp * 8 + hw * 4;
        uint4 r[4][CPL];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t row = row_base + i;
#pragma unroll
            for (int j = 0; j < CPL; j++)
                r[i][j] = row < C ? ld_stream_u4(cent.codes + (size_t)row * d_pad + (size_t)(l + 16 * j) * 16) : make_uint4(0, 0, 0, 0);
        }
        const uint32_t row_my = row_base + (l & 3);  // after the reduction lane l holds (query qg + l/4, row row_base + l%4)
        float2 h_my = make_float2(0.f, 0.f);
        uint2 s_my = make_uint2(0u, 0u);
        if (row_my < C) {
            h_my = cent.hdr[row_my];
            s_my = cent.sums[row_my];
        }
        __syncthreads();  // the previous item's shared reads are done
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(queries.codes + (size_t)q0 * d_pad);
            for (int i = threadIdx.x; i < nqc * CH; i += blockDim.x) sh_q[i] = src[i];
            for (int j = threadIdx.x; j < nqc; j += blockDim.x) {
                const float2 h = queries.hdr[q0 + j];
                const uint2 s = queries.sums[q0 + j];
                sh_side[j] = make_side(h.x, h.y, s.x, s.y, D);
            }
        }
        __syncthreads();
        for (int qg = 0; qg < nqc; qg += 4) {
            uint32_t v[16];
#pragma unroll
            for (int t = 0; t < 16; t++) v[t] = 0;
#pragma unroll
            for (int qq = 0; qq < 4; qq++) {
                const int q = min(qg + qq, nqc - 1);
#pragma unroll
                for (int j = 0; j < CPL; j++) {
                    const uint4 qv = sh_q[q * CH + l + 16 * j];
#pragma unroll
                    for (int i = 0; i < 4; i++) v[qq * 4 + i] = dot16(r[i][j], qv, v[qq * 4 + i]);
                }
            }
            // transpose-reduce 16 values over the 16 lanes of the half-warp: lane l ends with the total of value l
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const uint32_t send = (l & 8) ? v[t] : v[t + 8], keep = (l & 8) ? v[t + 8] : v[t];
                v[t] = keep + __shfl_xor_sync(FULL, send, 8);
            }
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint32_t send = (l & 4) ? v[t] : v[t + 4], keep = (l & 4) ? v[t + 4] : v[t];
                v[t] = keep + __shfl_xor_sync(FULL, send, 4);
            }
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const uint32_t send = (l & 2) ? v[t] : v[t + 2], keep = (l & 2) ? v[t + 2] : v[t];
                v[t] = keep + __shfl_xor_sync(FULL, send, 2);
            }
            {
                const uint32_t send = (l & 1) ? v[0] : v[1], keep = (l & 1) ? v[1] : v[0];
                v[0] = keep + __shfl_xor_sync(FULL, send, 1);
            }
            const int q = qg + (l >> 2);
            if (q < nqc && row_my < C) {
                bool flag;
                const float sim = score_fast(sh_side[q], h_my.x, h_my.y, s_my.x, s_my.y, v[0], D, &flag);
                keys[(size_t)(q0 + q) * key_stride + row_my] = f32_to_key(sim);
                if (flag) {
                    const unsigned int pos = atomicAdd(flag_cnt + q0 + q, 1u);
                    if (pos < flag_cap) flag_list[(size_t)(q0 + q) * flag_cap + pos] = row_my;
                }
            }
        }
    }
}

// One block per query: the k largest keys (ties by lowest centroid index) by bisection over the key row, a
// rank sort of the k survivors, literal re-scores where an uncertified pair could be among them, emission.
// KPT > 0: the thread's contiguous slice of the key row (<= KPT keys) lives in registers; KPT = 0: it is read from memory.
template <int KPT>
__global__ void __launch_bounds__(kSelThreadsP)
probe_select_kernel(MatView cent, MatView queries, uint32_t *__restrict__ keys, size_t key_stride,
                    const unsigned int *__restrict__ flag_cnt, const uint32_t *__restrict__ flag_list, uint32_t flag_cap, int k,
                    uint32_t *__restrict__ out_probe, float *__restrict__ out_sims, uint32_t *__restrict__ out_qtiles,
                    const uint64_t *__restrict__ next_list_len, uint32_t next_tile_rows, uint32_t *__restrict__ out_status,
                    uint32_t status_bit, int status_init, unsigned long long *fix_counter, uint32_t nseg, uint32_t seg_len,
                    uint32_t *__restrict__ cand_keys, uint32_t *__restrict__ cand_ids) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    double *sh_qn = reinterpret_cast<double *>(sel_smem);  // [D] normalized query (literal path only)
    __shared__ unsigned int s_wcount[2][kSelThreadsP / 32];
    __shared__ uint32_t s_key[kMaxProbe], s_id[kMaxProbe], o_key[kMaxProbe], o_id[kMaxProbe];
    __shared__ unsigned int s_nsel, s_wties[kSelThreadsP / 32], s_tiles;
    __shared__ int s_need_fix;
    __shared__ double s_norm;
    __shared__ unsigned char s_fstate[kProbeFlagCapMax];  // 0 = uncertified, 1 = re-scored, 2 = to re-score now

    // nseg > 1: the key row is cut into segments of seg_len keys, one block each; a block leaves its segment's k best
    // (exact after its own literal re-scores) in cand_keys / cand_ids and probe_final_kernel selects among them.
    const uint32_t q = blockIdx.x / nseg, seg = blockIdx.x % nseg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t C = (uint32_t)cent.n;
    const int D = cent.d, d_pad = cent.d_pad;
    const uint32_t seg_lo = min(C, seg * seg_len), seg_hi = min(C, seg_lo + seg_len);
    const uint32_t per = (seg_hi - seg_lo + kSelThreadsP - 1) / kSelThreadsP;
    const uint32_t lo = min(seg_hi, seg_lo + (uint32_t)tid * per), hi = min(seg_hi, lo + per);
    const int k_full = k;
    k = (int)min((uint32_t)k, seg_hi - seg_lo);  // a short segment keeps everything it has
    uint32_t *kq = keys + (size_t)q * key_stride;
    const unsigned int nflag_all = flag_cnt[q];
    const int nflag = (int)min(nflag_all, flag_cap);
    uint32_t status = nflag_all > flag_cap ? status_bit : 0u;  // unlisted uncertified pairs: the caller's literal path
    for (int f = tid; f < nflag; f += kSelThreadsP) s_fstate[f] = 0;
    if (tid == 0) s_tiles = 0;
    bool normalized = false;
    constexpr bool REG = KPT > 0;
    uint32_t rk[REG ? KPT : 1];
    uint32_t T = 0;
    for (;;) {
        if (REG) {  // key 0 = no entry: below every valid key
#pragma unroll
            for (int j = 0; j < (REG ? KPT : 1); j++) rk[j] = lo + j < hi ? kq[lo + j] : 0u;
        }
        // visits the thread's keys in index order (register copies are indexed statically); warp-uniform trip count
        auto for_keys = [&](auto &&body) {
            if (REG) {
#pragma unroll
                for (int j = 0; j < (REG ? KPT : 1); j++) body(rk[j], lo + j, lo + j < hi);
            } else {
                for (uint32_t j = 0; j < per; j++) body(lo + j < hi ? kq[lo + j] : 0u, lo + j, lo + j < hi);
            }
        };
        // ---- k-th largest key: bisection over the 32-bit key domain (largest x with count(key >= x) >= k).  A radix
        // select would be fewer passes, but similarities crowd into a handful of top-byte bins and its shared-memory
        // atomics serialize; counting needs no atomics at all.
        auto block_count = [&](uint32_t x, int buf) -> unsigned int {  // keys >= x in the whole row (x >= 1: padding is 0)
            unsigned int c = 0;
            if (REG) {
                unsigned int c4[4] = {0u, 0u, 0u, 0u};  // independent chains
#pragma unroll
                for (int j = 0; j < (REG ? KPT : 1); j++) c4[j & 3] += rk[j] >= x ? 1u : 0u;
                c = (c4[0] + c4[1]) + (c4[2] + c4[3]);
            } else {
                for_keys([&](uint32_t kk, uint32_t, bool valid) { c += (valid && kk >= x) ? 1u : 0u; });
            }
            c = __reduce_add_sync(FULL, c);
            if (lane == 0) s_wcount[buf][warp] = c;
            __syncthreads();
            unsigned int tot = 0;
#pragma unroll
            for (int w = 0; w < kSelThreadsP / 32; w++) tot += s_wcount[buf][w];
            return tot;
        };
        uint32_t klo = 1u, khi = 0xFFFFFFFFu;  // every valid key is >= 1 and k <= C: count(key >= 1) >= k
        int buf = 0;
        while (klo < khi) {
            const uint32_t mid = klo + (khi - klo) / 2 + 1;
            if (k > 0 && block_count(mid, buf) >= (unsigned int)k) klo = mid;
            else khi = mid - 1;
            buf ^= 1;
        }
        T = klo;  // the k-th largest key
        // how many of the keys equal to T belong to the k (lowest indices first)
        const uint32_t need_ties = (uint32_t)k - (T == 0xFFFFFFFFu ? 0u : block_count(T + 1u, buf));
        __syncthreads();
        // ---- collect: keys above T in any order, then the first need_ties ties in index order ----
        if (tid == 0) s_nsel = 0;
        __syncthreads();
        unsigned int my_ties = 0;
        for_keys([&](uint32_t kk, uint32_t idx, bool valid) {
            if (!valid) return;
            if (kk > T) {
                const unsigned int pos = atomicAdd(&s_nsel, 1u);
                s_key[pos] = kk;
                s_id[pos] = idx;
            } else if (kk == T) {
                my_ties++;
            }
        });
        unsigned int incl = my_ties;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int x = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) s_wties[warp] = incl;
        __syncthreads();
        unsigned int rank = incl - my_ties;
        for (int w = 0; w < warp; w++) rank += s_wties[w];
        const unsigned int base = s_nsel;  // = k - need_ties
        if (my_ties && rank < need_ties) {
            for_keys([&](uint32_t kk, uint32_t idx, bool valid) {
                if (valid && kk == T && rank < need_ties) {
                    s_key[base + rank] = T;
                    s_id[base + rank] = idx;
                    rank++;
                }
            });
        }
        if (tid == 0) s_need_fix = 0;
        __syncthreads();
        // ---- an uncertified pair matters iff its stored (upper) key reaches T: below T its true key is below T too ----
        for (int f = tid; f < nflag; f += kSelThreadsP) {
            const uint32_t c = flag_list[(size_t)q * flag_cap + f];
            if (s_fstate[f] == 0 && c >= seg_lo && c < seg_hi && kq[c] >= T) {
                s_fstate[f] = 2;
                s_need_fix = 1;
            }
        }
        __syncthreads();
        if (!s_need_fix) break;
        if (!normalized) {  // normalizeVector of the query (compute/cosine.go:26,138-149), literal
            const uint8_t *qc = queries.codes + (size_t)q * d_pad;
            const float2 qh = queries.hdr[q];
            const double mn = (double)qh.x, range = __dsub_rn((double)qh.y, (double)qh.x);
            for (int i = tid; i < D; i += kSelThreadsP) sh_qn[i] = ref_dequant_f64(qc[i], mn, range);
            __syncthreads();
            if (warp == 0) {
                const double nsq = warp_ordered_sum(D, lane, [&](int i) { return __dmul_rn(sh_qn[i], sh_qn[i]); });
                if (lane == 0) s_norm = __dsqrt_rn(nsq);
            }
            __syncthreads();
            const double norm = s_norm;
            if (norm != 0.0)
                for (int i = tid; i < D; i += kSelThreadsP) sh_qn[i] = __ddiv_rn(sh_qn[i], norm);
            __syncthreads();
            normalized = true;
        }
        for (int f = warp; f < nflag; f += kSelThreadsP / 32) {
            if (s_fstate[f] == 2) {
                const uint32_t c = flag_list[(size_t)q * flag_cap + f];
                const float2 h = cent.hdr[c];
                const double dot = warp_ref_cosine_row_f64(cent.codes + (size_t)c * d_pad, h.x, h.y, sh_qn, D, lane);
                if (lane == 0) {
                    kq[c] = f32_to_key(__double2float_rn(dot));
                    s_fstate[f] = 1;
                    if (fix_counter) atomicAdd(fix_counter, 1ull);
                }
            }
        }
        __syncthreads();  // the rewritten keys are visible to the whole block; select again
    }
    // ---- order the k survivors (similarity desc, centroid index asc) and emit ----
    if (tid < k) {
        const uint32_t mk = s_key[tid], mi = s_id[tid];
        int rnk = 0;
        for (int j = 0; j < k; j++) rnk += cand_better(s_key[j], (uint64_t)s_id[j], mk, (uint64_t)mi) ? 1 : 0;
        o_key[rnk] = mk;
        o_id[rnk] = mi;
    }
    __syncthreads();
    if (nseg > 1) {  // hand the segment's survivors to the final selection (key 0 = no entry)
        for (int r = tid; r < k_full; r += kSelThreadsP) {
            cand_keys[((size_t)q * nseg + seg) * k_full + r] = r < k ? o_key[r] : 0u;
            cand_ids[((size_t)q * nseg + seg) * k_full + r] = r < k ? o_id[r] : 0u;
        }
        return;
    }
    uint32_t mytiles = 0;
    if (tid < k) {
        const uint32_t L = o_id[tid];
        out_probe[(size_t)q * k + tid] = L;
        if (out_sims) out_sims[(size_t)q * k + tid] = key_to_f32(o_key[tid]);
        if (out_qtiles) {
            const uint32_t len = (uint32_t)next_list_len[L];
            mytiles = (len + next_tile_rows - 1) / next_tile_rows;
        }
    }
    if (out_qtiles) {
        for (int o = 16; o > 0; o >>= 1) mytiles += __shfl_xor_sync(FULL, mytiles, o);
        if (lane == 0 && mytiles) atomicAdd(&s_tiles, mytiles);
        __syncthreads();
        if (tid == 0) out_qtiles[q] = s_tiles ? s_tiles : 1u;
    }
    if (tid == 0 && out_status) out_status[q] = status_init ? status : (out_status[q] | status);
}

// Final selection among the nseg x k segment survivors of one query (all exact): the k best by (similarity desc,
// centroid index asc), found by bisection over the 64-bit composite (key, ~index), which is unique per entry.
__global__ void __launch_bounds__(kSelThreadsP)
probe_final_kernel(const uint32_t *__restrict__ cand_keys, const uint32_t *__restrict__ cand_ids, uint32_t ncand, uint32_t C,
                   const unsigned int *__restrict__ flag_cnt, uint32_t flag_cap, int k, uint32_t *__restrict__ out_probe,
                   float *__restrict__ out_sims, uint32_t *__restrict__ out_qtiles, const uint64_t *__restrict__ next_list_len,
                   uint32_t next_tile_rows, uint32_t *__restrict__ out_status, uint32_t status_bit, int status_init) {
    __shared__ unsigned int s_wcount[2][kSelThreadsP / 32];
    __shared__ uint32_t s_key[kMaxProbe], s_id[kMaxProbe], o_key[kMaxProbe], o_id[kMaxProbe];
    __shared__ unsigned int s_nsel, s_tiles;
    const uint32_t q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int PER = 16;  // ncand <= 32 segments x 128
    unsigned long long comp[PER];
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const uint32_t i = (uint32_t)j * kSelThreadsP + tid;
        comp[j] = 0ull;
        if (i < ncand) {
            const uint32_t kk = cand_keys[(size_t)q * ncand + i];
            if (kk != 0) comp[j] = ((unsigned long long)kk << 32) | (unsigned long long)(0xFFFFFFFFu - cand_ids[(size_t)q * ncand + i]);
        }
    }
    if (tid == 0) {
        s_nsel = 0;
        s_tiles = 0;
    }
    k = (int)min((uint32_t)k, C);
    unsigned long long lo = 1ull, hi = ~0ull;  // largest x with count(comp >= x) >= k
    int buf = 0;
    while (lo < hi) {
        const unsigned long long mid = lo + (hi - lo) / 2 + 1;
        unsigned int c = 0;
#pragma unroll
        for (int j = 0; j < PER; j++) c += comp[j] >= mid ? 1u : 0u;
        c = __reduce_add_sync(FULL, c);
        if (lane == 0) s_wcount[buf][warp] = c;
        __syncthreads();
        unsigned int tot = 0;
#pragma unroll
        for (int w = 0; w < kSelThreadsP / 32; w++) tot += s_wcount[buf][w];
        if (tot >= (unsigned int)k) lo = mid;
        else hi = mid - 1;
        buf ^= 1;
    }
#pragma unroll
    for (int j = 0; j < PER; j++) {
        if (comp[j] >= lo && comp[j] != 0ull) {
            const unsigned int pos = atomicAdd(&s_nsel, 1u);
            if (pos < (unsigned int)kMaxProbe) {
                s_key[pos] = (uint32_t)(comp[j] >> 32);
                s_id[pos] = 0xFFFFFFFFu - (uint32_t)comp[j];
            }
        }
    }
    __syncthreads();
    if (tid < k) {
        const uint32_t mk = s_key[tid], mi = s_id[tid];
        int rnk = 0;
        for (int j = 0; j < k; j++) rnk += cand_better(s_key[j], (uint64_t)s_id[j], mk, (uint64_t)mi) ? 1 : 0;
        o_key[rnk] = mk;
        o_id[rnk] = mi;
    }
    __syncthreads();
    uint32_t mytiles = 0;
    if (tid < k) {
        const uint32_t L = o_id[tid];
        out_probe[(size_t)q * k + tid] = L;
        if (out_sims) out_sims[(size_t)q * k + tid] = key_to_f32(o_key[tid]);
        if (out_qtiles) {
            const uint32_t len = (uint32_t)next_list_len[L];
            mytiles = (len + next_tile_rows - 1) / next_tile_rows;
        }
    }
    if (out_qtiles) {
        for (int o = 16; o > 0; o >>= 1) mytiles += __shfl_xor_sync(FULL, mytiles, o);
        if (lane == 0 && mytiles) atomicAdd(&s_tiles, mytiles);
        __syncthreads();
        if (tid == 0) out_qtiles[q] = s_tiles ? s_tiles : 1u;
    }
    if (tid == 0 && out_status) {
        const uint32_t status = flag_cnt[q] > flag_cap ? status_bit : 0u;  // unlisted uncertified pairs: the caller's literal path
        out_status[q] = status_init ? status : (out_status[q] | status);
    }
}

}  // namespace

cudaError_t probe_set_certify_scale(float scale) { return cudaMemcpyToSymbol(c_certify_scale, &scale, sizeof(float)); }

uint32_t probe_segments(size_t C) {  // key-row segments of at most 8192 keys (32 register-resident keys per thread)
    const size_t s = (C + 8191) / 8192;
    return (uint32_t)(s < 1 ? 1 : s);
}

bool probe_batch_supported(const MatView &cent, size_t nq, size_t k) {
    const int cpl = cent.d_pad / 256;
    return cent.d_pad % 256 == 0 && (cpl == 1 || cpl == 2 || cpl == 3 || cpl == 4 || cpl == 6) && k >= 1 && k <= (size_t)kMaxProbe &&
           k <= cent.n && probe_segments(cent.n) <= 32 && nq * cent.n <= ((size_t)32 << 20);
}

cudaError_t launch_probe_batch(const MatView &cent, const MatView &queries, int k, uint32_t *keys, unsigned int *flag_cnt,
                               uint32_t *flag_list, uint32_t flag_cap, uint32_t *cand_keys, uint32_t *cand_ids, uint32_t *out_probe,
                               float *out_sims, uint32_t *out_qtiles,
                               const uint64_t *next_list_len, uint32_t next_tile_rows, uint32_t *out_status, uint32_t status_bit,
                               int status_init, unsigned long long *fix_counter, int sm_count, cudaStream_t st) {
    const size_t nq = queries.n, C = cent.n;
    cudaError_t e = cudaMemsetAsync(flag_cnt, 0, nq * sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    const int qc = nq >= 128 ? 32 : 8;
    const size_t items = ((C + kPsRows - 1) / kPsRows) * ((nq + qc - 1) / qc);
    // (the work items are equal: a grid that does not divide them leaves the last wave part empty -- 1024 items on 592
    // blocks cost two items' time; VS_PROBE_GRID_MUL: blocks per SM at most, for experiments)
    static const int grid_mul = [] {
        const char *e = getenv("VS_PROBE_GRID_MUL");
        const int v = e ? atoi(e) : 0;
        return v >= 1 && v <= 32 ? v : 8;
    }();
    const unsigned grid = (unsigned)(items < (size_t)sm_count * grid_mul ? items : (size_t)sm_count * grid_mul);
    const size_t smem = (size_t)qc * cent.d_pad + qc * sizeof(SideConst);
#define VS_PROBE_SCORE2(CPL, QC)                                                                                              \
    do {                                                                                                                      \
        if (smem > 48 * 1024) {                                                                                               \
            e = cudaFuncSetAttribute(probe_score_kernel<CPL, QC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
            if (e != cudaSuccess) return e;                                                                                   \
        }                                                                                                                     \
        probe_score_kernel<CPL, QC><<<grid, kPsWarps * 32, smem, st>>>(cent, queries, keys, C, flag_cnt, flag_list, flag_cap); \
    } while (0)
#define VS_PROBE_SCORE(CPL)               \
    do {                                  \
        if (qc == 32) VS_PROBE_SCORE2(CPL, 32); \
        else VS_PROBE_SCORE2(CPL, 8);     \
    } while (0)
    switch (cent.d_pad / 256) {
        case 1: VS_PROBE_SCORE(1); break;
        case 2: VS_PROBE_SCORE(2); break;
        case 3: VS_PROBE_SCORE(3); break;
        case 4: VS_PROBE_SCORE(4); break;
        case 6: VS_PROBE_SCORE(6); break;
        default: return cudaErrorInvalidValue;
    }
#undef VS_PROBE_SCORE2
#undef VS_PROBE_SCORE
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t smem2 = (size_t)cent.d * sizeof(double);
    const uint32_t nseg = probe_segments(C);
    const uint32_t seg_len = (uint32_t)((C + nseg - 1) / nseg);
    const size_t per = ((size_t)seg_len + kSelThreadsP - 1) / kSelThreadsP;
#define VS_PROBE_SELECT(KPT)                                                                                                  \
    do {                                                                                                                      \
        if (smem2 > 40 * 1024) {                                                                                              \
            e = cudaFuncSetAttribute(probe_select_kernel<KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);      \
            if (e != cudaSuccess) return e;                                                                                   \
        }                                                                                                                     \
        probe_select_kernel<KPT><<<(unsigned)(nq * nseg), kSelThreadsP, smem2, st>>>(                                         \
            cent, queries, keys, C, flag_cnt, flag_list, flag_cap, k, out_probe, out_sims, out_qtiles, next_list_len,         \
            next_tile_rows, out_status, status_bit, status_init, fix_counter, nseg, seg_len, cand_keys, cand_ids);            \
    } while (0)
    if (per <= 8) VS_PROBE_SELECT(8);
    else if (per <= 16) VS_PROBE_SELECT(16);
    else VS_PROBE_SELECT(32);
#undef VS_PROBE_SELECT
    e = cudaGetLastError();
    if (e != cudaSuccess || nseg == 1) return e;
    probe_final_kernel<<<(unsigned)nq, kSelThreadsP, 0, st>>>(cand_keys, cand_ids, nseg * (uint32_t)k, (uint32_t)C, flag_cnt, flag_cap, k,
                                                             out_probe, out_sims, out_qtiles, next_list_len, next_tile_rows, out_status,
                                                             status_bit, status_init);
    return cudaGetLastError();
}

}  // namespace vs
