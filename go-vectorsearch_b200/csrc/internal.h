// internal.h -- host-side declarations shared by the .cu translation units of libvscuda.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <vector>

#include "common.cuh"

struct vs_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t launches = 0;
    uint64_t slowpath = 0;
    // scratch arena (device) grown on demand; pinned host staging
    void *scratch = nullptr;
    size_t scratch_cap = 0;
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    bool owns_stream = true;
    int *h_flags = nullptr;   // small pinned block of its own for flags read while other calls use `pinned`
    void *aux = nullptr;      // second device buffer (tensor-core assignment): lives beside the arena, grown on demand
    size_t aux_cap = 0;
    unsigned long long *d_fix_counter = nullptr;  // in-kernel literal re-scores (device)
    unsigned int *d_tickets = nullptr;            // [kMaxStageQueries] zeroed once; every launch re-arms them
    unsigned int *d_fused_sync = nullptr;         // [3] counters of the single-query kernel (fused.cu), same allocation
    unsigned long long *d_trace = nullptr;        // phase stamps of the last list-scan launch (when enabled)
    bool trace = false;
    // optional per-kernel timing of the list-scan stage
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;  // pairs (start, stop), recycled
    size_t prof_used = 0;
    cudaEvent_t phase_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // batch-search phase boundaries
    // a second stream for work that is independent of what runs on `stream` (fork / join by events), created on first use
    cudaStream_t side_stream = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
};

struct vs_matrix {
    std::atomic<int> refs{1};
    int device = 0;
    uint8_t *codes = nullptr;
    float2 *hdr = nullptr;
    uint2 *sums = nullptr;
    size_t n = 0;
    int d = 0;
    int d_pad = 0;
    bool owns = true;
    bool pooled = false;             // storage from the device memory pool (api.cu:pool_alloc) rather than cudaMalloc
    vs::MatView view() const { return vs::MatView{codes, hdr, sums, n, d, d_pad}; }
};

struct vs_index {
    vs_matrix *data = nullptr;       // rows grouped by list
    vs_matrix *centroids = nullptr;  // C rows
    uint64_t *doc_ids = nullptr;     // [n] device (may be null -> id = id_base + row)
    uint64_t id_base = 0;
    bool implicit_ids = false;       // built without document ids: rows are numbered id_base + primary-key order
    uint64_t *list_off = nullptr;    // [C+1] device: where every list starts (and the capacity boundary of the one before)
    uint64_t *list_len = nullptr;    // [C] device: rows every list holds; list_off[l] + list_len[l] <= list_off[l+1].  The
                                     // streaming loader's cursor array when there is one (fill_cursor), else owned
    size_t n = 0, C = 0;
    // Lists with room to grow (vs_index_create_empty / vs_index_fill* / vs_index_append): n is the CAPACITY in rows, filled the
    // rows placed so far, fill_cursor the rows placed per list (device, [C] + overflow flag); null for an index built in one
    // piece (filled unused, no holes).  While filled < n the store has holes between the lists and searches go by lists only.
    uint64_t *fill_cursor = nullptr;
    size_t filled = 0;
};

namespace vs {

constexpr int kMaxSeg = 128;      // max probed lists per query (the probe list is a top-k of capacity <= 128)
constexpr int kMaxStageQueries = 4096;  // queries per stage launch (their tile prefix lives in shared memory)
constexpr int kStageWarps = 8;    // warps per block of the stage kernel
constexpr int kTileRows = 32;     // rows scored per warp tile

// Parameters of one stage launch (probe selection, list scan or flat scan) for nq queries.
struct StageParams {
    MatView rows;                 // matrix being scanned
    const uint64_t *ids;          // per-row id (doc id) or null -> id_base + row
    uint64_t id_base;
    MatView queries;              // nq query rows
    const double *qnorm;          // EXACT only: [nq][d] normalized queries
    const uint32_t *q_select;     // optional: [nq_launch] indices into queries/outputs (resolve path) or null
    int nq;                       // queries in this launch (gridDim.y)
    // segments
    const uint32_t *seg_list;     // [q][seg_stride] list ids, or null = single segment
    int seg_stride;
    const uint64_t *list_off;     // list starts when seg_list != null
    const uint64_t *list_len;     // rows per list
    int nseg;
    uint64_t single_start, single_count;
    uint32_t seg_cap;             // when non-zero: at most this many rows of every segment are scanned (a sample)
    // work decomposition: tiles per query (stage 2: written by stage 1) or a uniform count
    const uint32_t *qtiles;       // [total queries] indexed like the outputs, or null
    uint32_t uniform_tiles;       // used when qtiles == null
    int iters;                    // rows per lane group in a tile (tile height = (32/G) * iters)
    // merge state
    Cand *partial;                // [nq + gridDim.x][CAP]: slot = query slot + block
    unsigned int *tickets;        // [nq], zeroed; reset by the last block
    // outputs
    int mode;                     // 0 = final hits, 1 = probe list
    int k;                        // hits (mode 0) or probes (mode 1) to emit
    uint64_t *out_ids;            // mode 0: [q][k]
    float *out_sims;              // mode 0: [q][k]; mode 1 (optional): [q][k]
    int32_t *out_counts;          // mode 0: [q]
    uint32_t *out_probe;          // mode 1: [q][k]
    uint32_t *out_qtiles;         // mode 1 (optional): [q] tiles the next stage will scan for this query
    const uint64_t *next_list_len;  // rows per list of the store scanned by the next stage (for out_qtiles)
    uint32_t next_tile_rows;
    unsigned long long *trace;    // optional [gridDim.x][8] globaltimer phase stamps (profiling aid)
    unsigned long long *fix_counter;  // device counter: candidates re-scored with literal arithmetic in-kernel
    uint32_t *out_status;         // [q], OR-ed with status_bit / need-more bit
    uint32_t status_bit;
    int status_init;              // 1: this stage stores the query's status word, 0: it ORs into it
    int pdl;                      // programmatic dependent launch: 1 = let the next stage's blocks be scheduled while
                                  // this one drains, 2 = launched that way: wait for the previous stage before reading
};

constexpr uint32_t kStatusProbeAmbiguous = 1u;
constexpr uint32_t kStatusListAmbiguous = 2u;
constexpr uint32_t kStatusNeedMore = 4u;

// api.cu: hooks for sharded.cu (one host thread driving several devices)
void internal_set_thread_device(int device);  // -1: back to the process default (vs_init)
int internal_thread_device();
int internal_sm_count();
int internal_fail(int code, const char *msg);

// scan.cu
cudaError_t launch_stage(const StageParams &p, int kpl, bool exact, int grid_blocks, cudaStream_t st);
int stage_lanes_per_row(int d_pad);
int stage_cap(int kpl);
cudaError_t launch_query_normalize(const MatView &queries, double *qnorm, cudaStream_t st,
                                   const unsigned int *only_if_nonzero = nullptr);
cudaError_t launch_cosine_1xN(const MatView &rows, const MatView &query, float *sims, uint32_t *dots,
                              uint32_t *worklist, unsigned int *work_count, int sm_count, cudaStream_t st);
cudaError_t launch_cosine_fix(const MatView &rows, const double *qnorm, float *sims, const uint32_t *worklist,
                              const unsigned int *work_count, int sm_count, cudaStream_t st);
cudaError_t launch_topk_merge(const uint64_t *ids_in, const float *sims_in, const int32_t *counts_in, size_t rank_stride_bytes,
                              int G, int nq, int k, uint64_t *ids_out, float *sims_out, int32_t *counts_out, cudaStream_t st);
cudaError_t launch_topk_merge_ptrs(const unsigned char *const *bufs, size_t ids_off, size_t sims_off, size_t counts_off, int G, int nq,
                                   int k, uint64_t *ids_out, float *sims_out, int32_t *counts_out, cudaStream_t st);

cudaError_t launch_topk_merge_exchange(const unsigned char *const *bufs, size_t ids_off, size_t sims_off, size_t counts_off, int G, int nq,
                                       int k, uint64_t *ids_out, float *sims_out, int32_t *counts_out, uint32_t *const *signal,
                                       const uint32_t *wait, uint32_t step, cudaStream_t st);

cudaError_t launch_topk_merge_exchange_split(const unsigned char *const *bufs, size_t ids_off, size_t sims_off, size_t counts_off, int G,
                                             int nq, int k, uint64_t *ids_out, float *sims_out, int32_t *counts_out,
                                             uint32_t *const *signal, const uint32_t *wait, uint32_t step, cudaStream_t st);
cudaError_t scan_set_certify_scale(float scale);
cudaError_t argmax_set_certify_scale(float scale);

cudaError_t gemm_set_certify_scale(float scale);

// fused.cu: one query in one cooperative launch (probe stage, grid barrier, selection, list stage through a TMA ring)
struct FusedParams {
    MatView rows;              // the store (rows grouped by list) or a flat matrix
    const uint64_t *ids;       // per-row document id, or null -> id_base + row (distinct: no de-duplication needed)
    uint64_t id_base;
    MatView cent;              // centroid table (npe > 0)
    const uint64_t *list_off;  // list starts (npe > 0)
    const uint64_t *list_len;  // rows per list
    MatView query;             // one row
    int npe;                   // lists to probe (< number of centroids, <= kMaxSeg); 0 = flat scan of [flat_start, +flat_count)
    uint64_t flat_start, flat_count;
    int k, pub;                // hits to emit; distinct documents a block publishes (fused_pub)
    uint32_t *keys;            // [C + 4] similarity keys of the centroids (16-byte aligned)
    uint4 *seginfo;            // [C] per centroid: (list start lo, hi, rows in the list, similarity uncertified)
    unsigned int *sync;        // [2] grid barrier counter and ticket counter: zero at launch; re-armed by the kernel
    uint4 *partial;            // [gridDim][pub] published candidates (key, meta, id lo, id hi)
    uint64_t *out_ids;         // [k]
    float *out_sims;           // [k]
    int32_t *out_counts;       // [1]
    uint32_t *out_status;      // [1] stored (kStatusProbeAmbiguous / kStatusListAmbiguous / 0)
    uint32_t *out_probe;       // optional [npe]
    unsigned long long *trace; // optional [gridDim][16] globaltimer stamps
    int stage_bytes, stages;   // ring geometry (filled by the launcher)
};
bool fused_supported(int d_pad, int kpl);
int fused_pub(int k, int kpl);
cudaError_t launch_fused_search(const FusedParams &p, int kpl, int grid, cudaStream_t st);
cudaError_t fused_set_certify_scale(float scale);

// listmajor.cu: the list stage of a query batch with every probed list read once for all the queries that probe it
struct LmItem {
    uint32_t row0;      // first store row
    uint32_t nrows;     // <= 1024 (a longer list is cut into several items)
    uint32_t pair_off;  // the item's queries: pairs[pair_off .. pair_off + m)
    uint32_t m;
};
struct LmParams {
    MatView rows;              // the store (rows grouped by list)
    const uint64_t *ids;       // per-row document id, or null -> id_base + row
    uint64_t id_base;
    MatView queries;           // the batch
    SideConst *sides;          // [nq] query side of the score identity
    uint32_t *pairs;           // [nq * npe] queries grouped by list
    LmItem *items;
    uint32_t *nitems;          // device word
    unsigned int *next_item;   // device word: work queue cursor
    uint32_t *gthr;            // [nq] running lower bound of the query's k-th best distinct document (score key)
    unsigned int *gcnt;        // [nq] candidates appended
    uint4 *gbuf;               // [nq][gcap] candidates (key, meta, id lo, id hi)
    int gcap;
    int k, pub;
    int stage_bytes;           // ring geometry (filled by the launcher)
    int tighten_at;            // dense form: candidates of a query after which its bound is recomputed (0: never); the first
                               // 4 * tighten_at entries of every query's list are zeroed before the scan
};
constexpr int kLmGbufCap = 4096;   // candidates per query the list-major scan may append (more: the literal path)
bool lm_supported(int d_pad, int k);
size_t lm_items_cap(size_t n_rows, size_t nq, size_t npe);
cudaError_t lm_enqueue_prepare(const LmParams &p, const uint32_t *probe, uint32_t nq, uint32_t npe, uint32_t C,
                               const uint64_t *list_off, const uint64_t *list_len, uint32_t *count, uint32_t *pair_off,
                               uint32_t items_cap, cudaStream_t st,
                               uint64_t *launches);
cudaError_t lm_enqueue_seed(const LmParams &p, const float *first_list_sims, const int32_t *first_list_counts, uint32_t nq,
                            cudaStream_t st, uint64_t *launches);
cudaError_t lm_enqueue_seed_scan(const LmParams &p, const uint32_t *probe, uint32_t nq, uint32_t npe, const uint64_t *list_off,
                                 const uint64_t *list_len, uint32_t sample, cudaStream_t st, uint64_t *launches);
bool lm_dense_supported(int d_pad);
cudaError_t lm_enqueue_scan(const LmParams &p, int sm_count, cudaStream_t st, uint64_t *launches, bool dense);
cudaError_t lm_enqueue_final(const LmParams &p, uint32_t nq, uint64_t *out_ids, float *out_sims, int32_t *out_counts,
                             uint32_t *out_status, unsigned long long *fix_counter, cudaStream_t st, uint64_t *launches);
cudaError_t lm_set_certify_scale(float scale);

// gemm.cu: query batches as a tcgen05 int8 GEMM with a fused filter (BASELINE config 3)
struct GemmPlan {
    uint32_t nq_pad;        // queries rounded up to the query tile
    uint32_t tiles;         // 128-row store tiles
    uint32_t rank;          // the threshold is the rank-th largest group maximum of the sample
    uint32_t sample_stride, sample_tiles, G;  // pre-pass sample: every sample_stride-th tile; G = 8 groups per tile
    uint32_t cand_per_q;    // candidate bucket capacity per query
};
struct GemmBufs {
    float4 *col_consts;
    float *gmax;
    float *tau;
    unsigned int *bounds;   // [0..2] row-side bounds
    unsigned int *cand_cnt; // [nq_pad] candidates emitted per query (may exceed the bucket: then the query needs more)
    uint2 *cand_rowdot;     // [nq_pad][cand_per_q] (store row, integer dot)
};
bool gemm_store_supported(const MatView &rows);
bool gemm_supported(const MatView &rows, size_t nq);
GemmPlan gemm_plan(const MatView &rows, size_t nq, size_t k, bool unique_ids, uint32_t sample_div, uint32_t min_sample_tiles,
                   size_t cand_per_query);
size_t gemm_scratch_bytes(const GemmPlan &pl, size_t nq);
void gemm_take(char *base, const GemmPlan &pl, size_t nq, GemmBufs *b);
cudaError_t gemm_enqueue_prepass(const MatView &rows, const MatView &queries, const GemmPlan &pl, const GemmBufs &b, uint32_t *d_status,
                                 int sm_count, cudaStream_t st, uint64_t *launches);
cudaError_t gemm_enqueue_filter(const MatView &rows, const MatView &queries, const GemmPlan &pl, const GemmBufs &b, int sm_count,
                                cudaStream_t st, uint64_t *launches);
cudaError_t gemm_enqueue_select(const MatView &rows, const uint64_t *ids, uint64_t id_base, bool unique_ids, const MatView &queries,
                                const GemmPlan &pl,
                                const GemmBufs &b, int k, uint64_t *d_ids, float *d_sims, int32_t *d_counts, uint32_t *d_status,
                                unsigned long long *fix_counter, int sm_count, cudaStream_t st, uint64_t *launches);

// probe.cu: probe selection for a batch of queries (the centroid table is read once, not once per query)
constexpr int kProbeFlagCapMax = 4096;  // uncertified (query, centroid) pairs listed per query: about 1 in 10^3 pairs is
inline uint32_t probe_flag_cap(size_t C) {  // uncertified (more where |cos| is small), only those near the top matter
    const size_t c = C / 16;
    return (uint32_t)(c < 64 ? 64 : (c > (size_t)kProbeFlagCapMax ? (size_t)kProbeFlagCapMax : c));
}
bool probe_batch_supported(const MatView &cent, size_t nq, size_t k);
uint32_t probe_segments(size_t C);  // the select stage cuts a key row into this many segments (1: no final pass)
cudaError_t launch_probe_batch(const MatView &cent, const MatView &queries, int k, uint32_t *keys, unsigned int *flag_cnt,
                               uint32_t *flag_list, uint32_t flag_cap, uint32_t *cand_keys, uint32_t *cand_ids, uint32_t *out_probe,
                               float *out_sims, uint32_t *out_qtiles,
                               const uint64_t *next_list_len, uint32_t next_tile_rows, uint32_t *out_status, uint32_t status_bit,
                               int status_init, unsigned long long *fix_counter, int sm_count, cudaStream_t st);
cudaError_t probe_set_certify_scale(float scale);

// gemm.cu: a plain 2-D TMA tile map over a byte matrix (tm_out: CUtensorMap*)
bool make_u8_tile_map(void *tm_out, const uint8_t *base, uint64_t cols, uint64_t rows, uint64_t row_stride, uint32_t box_cols,
                      uint32_t box_rows);

// quantize.cu
cudaError_t launch_quantize_f32(const float *in, size_t n, int d, uint8_t *out_rows, cudaStream_t st);
cudaError_t launch_quantize_f64(const double *in, size_t n, int d, uint8_t *out_rows, cudaStream_t st);
cudaError_t launch_quantize_f32_soa(const float *in, size_t n, int d, uint8_t *codes, int d_pad, float2 *hdr,
                                    uint2 *sums, cudaStream_t st);
cudaError_t launch_dequantize_f32(const uint8_t *rows, size_t n, int row_bytes, float *out, cudaStream_t st);
cudaError_t launch_dequantize_f64(const uint8_t *rows, size_t n, int row_bytes, double *out, cudaStream_t st);
cudaError_t launch_ingest(const uint8_t *rows, size_t n, int row_bytes, uint8_t *codes, int d_pad, float2 *hdr,
                          uint2 *sums, cudaStream_t st);
cudaError_t launch_export(const MatView &m, size_t first, size_t count, uint8_t *rows_out, cudaStream_t st);
cudaError_t launch_gather_rows(const MatView &src, const uint32_t *order, size_t n, uint8_t *codes, float2 *hdr,
                               uint2 *sums, const uint64_t *ids_in, uint64_t id_base, uint64_t *ids_out,
                               cudaStream_t st);
cudaError_t launch_merge_rows(const MatView &a, const uint64_t *a_ids, uint64_t a_id_base, const MatView &b, const uint64_t *b_ids,
                              uint64_t b_id_base, const uint32_t *order, size_t n, uint8_t *codes, float2 *hdr, uint2 *sums,
                              uint64_t *ids_out, cudaStream_t st);
cudaError_t launch_scatter_rows(const MatView &src, const uint32_t *order, const uint32_t *keys, const uint32_t *chunk_off,
                                const uint64_t *list_off, const uint64_t *cursor, uint8_t *codes, float2 *hdr, uint2 *sums,
                                const uint64_t *ids_in, uint64_t id_base, uint64_t *ids_out, unsigned int *overflow, cudaStream_t st);

// argmax.cu
cudaError_t launch_argmax(const MatView &cent, const MatView &data, const uint32_t *canon, int32_t *idx_out,
                          float *sims_out, uint32_t *worklist, unsigned int *work_count, int sm_count,
                          cudaStream_t st);
cudaError_t launch_argmax_fix(const MatView &cent, const MatView &data, const double *cnorm, int32_t *idx_out,
                              float *sims_out, const uint32_t *worklist, const unsigned int *work_count, int sm_count,
                              cudaStream_t st);
cudaError_t launch_canonical_rows(const MatView &m, uint32_t *canon, cudaStream_t st);

// kmeans.cu
cudaError_t launch_kmeans_accumulate(const MatView &data, const uint32_t *order, const uint32_t *seg_off, int k,
                                     float *means, int64_t *counts, cudaStream_t st, const float *means_prev = nullptr);
cudaError_t launch_kmeans_accumulate_scan(const MatView &data, const int32_t *assign, int k, float *means, int64_t *counts,
                                          cudaStream_t st, const float *means_prev = nullptr);
bool kmeans_ring_supported(int d_pad);
size_t kmeans_ring_scratch_bytes(size_t n, int k, int d_pad);
cudaError_t launch_kmeans_accumulate_ring(const MatView &data, const int32_t *assign, int k, float *means, int64_t *counts,
                                          const float *means_prev, void *scratch, cudaStream_t st);
cudaError_t launch_kmeans_accumulate_relay(const MatView &data, const uint32_t *order, const uint32_t *seg_off, int k,
                                           float *sums, int64_t *counts, cudaStream_t st);
cudaError_t launch_kmeans_finalize(const float *sums, const int64_t *counts, size_t k, int d, float *means, cudaStream_t st);
cudaError_t launch_recenter_clusters(const MatView &data, const uint32_t *order, const uint32_t *seg_off, int k, double *means,
                                     int64_t *counts, cudaStream_t st);
cudaError_t launch_recenter(const MatView &data, double *mean_out, cudaStream_t st);

}  // namespace vs
