/*
 * vscuda.h -- C ABI of libvscuda.so, the B200 (sm_100a) backend for go-vectorsearch's
 * similarity-search hot path.
 *
 * This is the drop-in boundary: the entry points below are what a `//go:build cuda`
 * sibling of compute/compute_gonum.go + compute/cosine_gonum.go binds through cgo
 * (see INTEGRATION.md and go-vectorsearch_b200/goshim/).  Plain pointers and sizes only.
 * Every entry point cites the reference interface (file:line under the reference repo)
 * it replaces.  All functions return VS_OK (0) or a negative VS_E* code; the message is
 * available from vs_last_error() (thread-local).  The Go shim turns a non-zero status into
 * panic()/logger.Fatalf exactly where the reference panics / Fatalf's.
 *
 * Row format ("row776"): the reference's own stored vector, compute/quantization.go:82-91 --
 * float32 min LE, float32 max LE, then D uint8 codes; 8+D bytes (776 for D=768).
 * "packed rows" = n such rows back to back (dnc/dataset.go:53-56 spool format).
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device and fails with
 * VS_ENODEV otherwise.
 */
#ifndef VSCUDA_H
#define VSCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_API __attribute__((visibility("default")))

#define VS_OK 0
#define VS_EINVAL (-1)   /* bad argument */
#define VS_EEMPTY (-2)   /* reference panics: "vector columns are empty" / "matrix rows are empty" (compute.go:13,26,30) */
#define VS_EDIM (-3)     /* reference Fatalf: "column size does not match" (cosine.go:19-21,77-79) */
#define VS_ECUDA (-4)    /* CUDA runtime error */
#define VS_ENODEV (-5)   /* no CUDA device / library not initialised */
#define VS_ENOMEM (-6)
#define VS_ERANGE (-7)   /* k / nprobe / dimension beyond what the kernels support */
#define VS_EFULL (-8)    /* vs_index_append: a list (or the store) has no room left; nothing was changed */

typedef struct vs_ctx vs_ctx;       /* one CUDA stream + scratch arena; one per goroutine/closure */
typedef struct vs_matrix vs_matrix; /* immutable device-resident quantized matrix, ref-counted */
typedef struct vs_index vs_index;   /* device-resident IVF-Flat index */

/* ---- lifecycle --------------------------------------------------------------------- */
VS_API int vs_init(int device);            /* select device (LOCAL_RANK); idempotent */
VS_API void vs_shutdown(void);
VS_API const char *vs_last_error(void);
VS_API int vs_device_info(char *name, size_t name_cap, int *sm_count, size_t *total_mem);
/* Matrices and index stores live in the device's stream-ordered memory pool, which keeps up to VS_POOL_KEEP_GB (environment,
 * default 24) of freed memory mapped for the next allocation.  This hands it back to the driver (e.g. before another
 * library in the process needs the room). */
VS_API int vs_release_cached_memory(void);

/* compute/cosine.go:60-66,129-135: VectorMatrixCosineSimilarity()/MatrixCosineSimilarity()
 * return (calculate, done).  calculate-closure creation == vs_ctx_create, done() == vs_ctx_destroy.
 * Contexts are independent and may be used concurrently from different threads (dnc/dnc.go:30-33). */
VS_API int vs_ctx_create(vs_ctx **out);
/* Same, but work is enqueued on a caller-owned cudaStream_t (e.g. the stream NCCL collectives run on). */
VS_API int vs_ctx_create_on_stream(void *cuda_stream, vs_ctx **out);
VS_API void vs_ctx_destroy(vs_ctx *ctx);
VS_API int vs_ctx_sync(vs_ctx *ctx);
VS_API void *vs_ctx_stream(vs_ctx *ctx);                 /* cudaStream_t, for callers that time with events */
VS_API uint64_t vs_ctx_launch_count(const vs_ctx *ctx);  /* kernels launched through this ctx so far */
VS_API uint64_t vs_ctx_slowpath_count(const vs_ctx *ctx);/* rows / queries that took the literal-arithmetic path */
/* Test hook: multiply every certification half-width by scale >= 1 so that rows are sent
 * through the literal reference-arithmetic paths; results must not change.  Process-wide; 1.0 = production. */
VS_API int vs_debug_set_certify_scale(float scale);
/* Test hook: from how many centroids on vs_argmax_MxN / vs_kmeans_step assign through the tensor-core GEMM
 * (default 256; below that the HBM-bound dp4a scan of the data rows is faster). */
VS_API int vs_debug_set_argmax_gemm_min(size_t min_centroids);
/* Test hook: 0 sends single-query searches through the two-launch streaming path (scan.cu) instead of the one-launch
 * fused kernel (fused.cu: probe stage, grid barrier, selection, TMA-ring list scan and top-k in one cooperative
 * launch; replaces server/search.go:214-273 for one query).  Default 1. */
VS_API int vs_debug_set_fused(int on);
/* Test hook: 0 sends the list stage of query batches through the query-major streaming scan (scan.cu) instead of the
 * list-major one (listmajor.cu: every probed list read once for all the queries of the batch that probe it; replaces
 * server/search.go:241-273 for a batch).  Default 1. */
VS_API int vs_debug_set_list_major(int on);
/* Test hook: the list-major scan takes its tensor-core form (lm_dense_kernel: the staged rows of a list against up to 16
 * of its queries per mma.sync.m16n8k32 pass) when the batch has at least this many (query, list) pairs per list; 0 = never
 * (always the dp4a form), 1 = whenever the row width allows.  Default 4.  Same hits either way. */
VS_API int vs_debug_set_lm_dense_min(int queries_per_list);
/* CUDA-event timing on the ctx stream (bench.py): start/stop bracket, elapsed in ms after sync. */
VS_API int vs_ctx_timer_start(vs_ctx *ctx);
VS_API int vs_ctx_timer_stop(vs_ctx *ctx, float *ms_out);
/* Per-kernel timing of the dominant kernel (the list-scan stage of vs_search*): when enabled, every
 * launch of it is bracketed by CUDA events on the ctx stream; read returns the summed duration and the
 * number of launches since the last read (synchronizes). */
VS_API int vs_ctx_profile_enable(vs_ctx *ctx, int on);
VS_API int vs_ctx_profile_read(vs_ctx *ctx, double *scan_ms_out, uint64_t *scan_launches_out);
/* Phase trace (profiling aid): when enabled, thread 0 of every block of each search stage stamps
 * %globaltimer (ns) into 16 slots: 0 start, 1 prologue done, 2 scan done, 8 block list sorted, 9 written,
 * 10 fenced, 3 ticket taken; last block of a query: 11 slot heads gathered, 12 sorted, 13 slots walked,
 * 4 merged, 5 final list, 6/7 emitted.  stage = 1 (probe) or 2 (list scan); out = blocks*16 values of the
 * most recent launch of that stage. */
VS_API int vs_ctx_trace_enable(vs_ctx *ctx, int on);
VS_API int vs_ctx_trace_read(vs_ctx *ctx, int stage, uint64_t *out, size_t max_blocks, size_t *blocks_out);

/* ---- compute/quantization.go ------------------------------------------------------- */
/* QuantizeMatrixFloat32 (quantization.go:142-148) / QuantizeVectorFloat32 (:82-91; n=1).
 * in: n*d floats (host); out: n*(8+d) bytes (host). */
VS_API int vs_quantize_f32(vs_ctx *ctx, const float *in, size_t n, size_t d, uint8_t *out);
/* QuantizeMatrixFloat64 (:150-156) / QuantizeVectorFloat64 (:93-102). */
VS_API int vs_quantize_f64(vs_ctx *ctx, const double *in, size_t n, size_t d, uint8_t *out);
/* DequantizeMatrixFloat32 (:166-172) / DequantizeVectorFloat32 (:114-122). */
VS_API int vs_dequantize_f32(vs_ctx *ctx, const uint8_t *rows, size_t n, size_t row_bytes, float *out);
/* DequantizeMatrixFloat64 (:174-180) / DequantizeVectorFloat64 (:124-132). */
VS_API int vs_dequantize_f64(vs_ctx *ctx, const uint8_t *rows, size_t n, size_t row_bytes, double *out);
/* Device-pointer forms of the above (no host copies, asynchronous on the ctx stream). */
VS_API int vs_quantize_f32_dev(vs_ctx *ctx, const float *d_in, size_t n, size_t d, uint8_t *d_out);
VS_API int vs_quantize_f64_dev(vs_ctx *ctx, const double *d_in, size_t n, size_t d, uint8_t *d_out);

/* ---- compute/compute.go: NewMatrix / NewVector / Clone ----------------------------- */
/* NewMatrix (compute.go:23-44): rows_packed = n rows of row_bytes (host).  Unlike the reference
 * nothing is dequantized: the device keeps codes[n][D], the 8-byte headers and two integer sums
 * per row.  VS_EEMPTY for n==0 or row_bytes<=8 (the reference's panics). */
VS_API int vs_matrix_create(vs_ctx *ctx, const uint8_t *rows_packed, size_t n, size_t row_bytes, vs_matrix **out);
VS_API int vs_matrix_create_dev(vs_ctx *ctx, const uint8_t *d_rows_packed, size_t n, size_t row_bytes, vs_matrix **out);
/* Quantize n*d device floats straight into a device matrix (QuantizeMatrixFloat32 + NewMatrix fused). */
VS_API int vs_matrix_from_f32_dev(vs_ctx *ctx, const float *d_in, size_t n, size_t d, vs_matrix **out);
/* Allocate an n x d matrix and fill row ranges from device float32 rows (bulk loaders). */
VS_API int vs_matrix_create_empty(vs_ctx *ctx, size_t n, size_t d, vs_matrix **out);
VS_API int vs_matrix_fill_f32_dev(vs_ctx *ctx, vs_matrix *m, size_t first_row, const float *d_in, size_t count);
/* Overwrite rows [first_row, first_row+count) from host row776 rows (asynchronous H2D + ingest on the ctx
 * stream; the host buffer must stay valid until the next vs_ctx_sync). Used to refill a query matrix. */
VS_API int vs_matrix_load_rows(vs_ctx *ctx, vs_matrix *m, size_t first_row, const uint8_t *rows_packed, size_t count);
/* Clone (compute.go:65-77): device matrices are immutable, so Clone is a reference-count bump. */
VS_API void vs_matrix_retain(vs_matrix *m);
VS_API void vs_matrix_release(vs_matrix *m);
VS_API size_t vs_matrix_rows(const vs_matrix *m);
VS_API size_t vs_matrix_cols(const vs_matrix *m);
/* Read rows back in row776 format (host out: count*(8+D) bytes). */
VS_API int vs_matrix_read_rows(vs_ctx *ctx, const vs_matrix *m, size_t first, size_t count, uint8_t *out);
/* The D&C row spool of the reference (dnc/dataset.go:19-56,122-146): a flat file of 8+D-byte rows, no header.
 * vs_matrix_load_spool builds a device matrix from rows [first_row, first_row+count) of the file (count 0 = to the end)
 * with double-buffered pinned reads overlapping the upload; a file that is not a whole number of rows is an error
 * (dataset.go:131-136).  vs_matrix_save_spool writes rows of a device matrix in the same format (append != 0: at the
 * end of an existing file, like createDataset.WriteRow). */
VS_API int vs_matrix_load_spool(vs_ctx *ctx, const char *path, size_t row_bytes, size_t first_row, size_t count, vs_matrix **out);
VS_API int vs_matrix_save_spool(vs_ctx *ctx, const vs_matrix *m, size_t first, size_t count, const char *path, int append);

/* ---- compute/cosine.go ------------------------------------------------------------- */
/* (*vectorContainer).MatrixCosineSimilarity (cosine.go:13-57): sims_out[i] = float32 cosine of the
 * query row776 against row i; n floats (host).  Bit-equal to the reference's default backend. */
VS_API int vs_cosine_1xN(vs_ctx *ctx, const uint8_t *q, size_t q_bytes, const vs_matrix *m, float *sims_out);
/* The integer dot product sum_j q[j]*v[j] (uint8 x uint8 -> uint32) the scores are built from. */
VS_API int vs_dot_1xN(vs_ctx *ctx, const uint8_t *q, size_t q_bytes, const vs_matrix *m, uint32_t *dots_out);
/* (*matrixContainer).MatrixCosineSimilarity (cosine.go:70-125): receiver = centroids, argument = data.
 * idx_out[i] = argmax over centroids for data row i (strict '>' from -1.0: lowest index wins ties);
 * sims_out (nullable; every reference call site discards it) = float32 of the winning cosine. */
VS_API int vs_argmax_MxN(vs_ctx *ctx, const vs_matrix *centroids, const vs_matrix *data, float *sims_out, int64_t *idx_out);
/* Same with the result left on the device (int32 indices), asynchronous. */
VS_API int vs_argmax_MxN_dev(vs_ctx *ctx, const vs_matrix *centroids, const vs_matrix *data, int32_t *d_idx_out);

/* ---- server/search.go:202-273: device-resident IVF-Flat probe-and-scan ------------- */
/* Build from rows already grouped by list: list_offsets[C+1] (CSR, rows of list c are
 * [list_offsets[c], list_offsets[c+1])), doc_ids[n] = Embedding.DocumentID (database/model.go:14),
 * centroids = C rows (database/model.go:35-37).  Host pointers. doc_ids may be NULL (id = row index). */
VS_API int vs_index_build(vs_ctx *ctx, const uint8_t *rows_packed, size_t n, size_t row_bytes, const uint64_t *doc_ids,
                   const uint64_t *list_offsets, const uint8_t *centroids_packed, size_t C, vs_index **out);
/* Build from ungrouped rows + the centroid index of each row (Embedding.CentroidID, model.go:16);
 * groups on device with a stable sort so rows keep primary-key order inside each list. */
VS_API int vs_index_build_assigned(vs_ctx *ctx, const uint8_t *rows_packed, size_t n, size_t row_bytes,
                            const uint64_t *doc_ids, const uint32_t *list_of_row, const uint8_t *centroids_packed,
                            size_t C, vs_index **out);
/* Device-resident build: data and centroids are vs_matrix handles, d_list_of_row/d_doc_ids device arrays
 * (d_doc_ids may be NULL -> id = id_base + row index). */
VS_API int vs_index_build_dev(vs_ctx *ctx, const vs_matrix *data, const int32_t *d_list_of_row, const uint64_t *d_doc_ids,
                       uint64_t id_base, const vs_matrix *centroids, vs_index **out);
/* Streaming loader: the database as a one-time loader (database/model.go:9-18: id, vector, document_id, centroid_id;
 * search.go:241-243 reads a list in primary-key order).  vs_index_create_empty reserves the grouped store at its final
 * size from the per-list row counts (host, [C] -- SELECT centroid_id, COUNT(*) ... GROUP BY centroid_id, the query of
 * dnc.go:465-470); vs_index_fill* then places chunks of rows, handed over in primary-key order, straight into their
 * lists.  HBM holds the store once plus one chunk, so a shard can be sized for nearly all of the 180 GB (the one-piece
 * builds above hold the ungrouped rows and the grouped copy at the same time).  Once every reserved row is placed the
 * result is identical to vs_index_build_assigned over the same rows; before that the index answers from the rows placed
 * so far (every list knows how many rows it holds; see vs_index_append for what that costs).  One loader thread per
 * index (the fill calls mutate it and are ordered); a fill that fails with VS_EINVAL because a list overflowed its
 * reservation leaves the index unusable -- release it.
 * vs_index_fill_dev: chunk = device matrix, d_list_of_row[m] int32 list index per row (e.g. from vs_argmax_MxN_dev),
 * d_doc_ids[m] or NULL (id = id_base + row index in the chunk); vs_index_fill: the same from host buffers. */
VS_API int vs_index_create_empty(vs_ctx *ctx, const vs_matrix *centroids, const uint64_t *list_counts, vs_index **out);
VS_API int vs_index_fill_dev(vs_ctx *ctx, vs_index *ix, const vs_matrix *chunk, const int32_t *d_list_of_row,
                      const uint64_t *d_doc_ids, uint64_t id_base);
VS_API int vs_index_fill(vs_ctx *ctx, vs_index *ix, const uint8_t *rows_packed, size_t n, size_t row_bytes,
                  const uint32_t *list_of_row, const uint64_t *doc_ids, uint64_t id_base);
VS_API void vs_index_release(vs_index *ix);
VS_API size_t vs_index_rows(const vs_index *ix);
VS_API size_t vs_index_lists(const vs_index *ix);
VS_API size_t vs_index_cols(const vs_index *ix);   /* vector size of the rows (row776 is 8 + this many bytes) */
/* Read back the CSR offsets (C+1 values) and rows [first, first+count) of the grouped store with their ids. */
VS_API int vs_index_list_offsets(vs_ctx *ctx, const vs_index *ix, uint64_t *offsets_out);
VS_API int vs_index_read_rows(vs_ctx *ctx, const vs_index *ix, size_t first, size_t count, uint8_t *rows_out, uint64_t *ids_out);
/* Upload's assignment and insert (server/upload.go:239-279): every new row (row776, host) goes to its nearest centroid
 * (centroids.MatrixCosineSimilarity(embeddings), upload.go:245; lowest index wins ties, cosine.go:114) and joins that
 * list behind the rows already there -- new embeddings get larger primary keys and search.go:241-243 streams a list in
 * primary-key order.  *out is a NEW index holding the old and the new rows (one pass at copy bandwidth: no re-sort of
 * the store, no host round trip); the old handle stays valid for the searches in flight and is released by the caller.
 * doc_ids[n] = Embedding.DocumentID of the new rows: required when the index was built with explicit ids, else NULL
 * continues the implicit numbering (id_base + row).  assign_out[n] (nullable) = the list index of every new row, what
 * upload.go:268-271 turns into Embedding.CentroidID.  Errors as NewMatrix / MatrixCosineSimilarity: VS_EEMPTY, VS_EDIM. */
VS_API int vs_index_upload(vs_ctx *ctx, const vs_index *ix, const uint8_t *rows_packed, size_t n, size_t row_bytes,
                    const uint64_t *doc_ids, int64_t *assign_out, vs_index **out);
/* Upload in place.  vs_index_upload copies the store once per call; a service that ingests continuously keeps room
 * behind every list instead.  vs_index_with_room: a copy of ix (one pass at copy bandwidth) in which list l can hold
 * len(l) + max(len(l) * percent / 100, min_rows) rows.  vs_index_append: the assignment of upload.go:245 and the insert,
 * in place -- every new row is written behind the last row of its nearest centroid's list (a few launches, cost
 * proportional to the new rows only); arguments and assign_out as vs_index_upload.  VS_EFULL when some list (or the store)
 * cannot take its new rows: NOTHING was changed; make a roomier copy with vs_index_with_room (or merge with
 * vs_index_upload) and append to that.  Searches see a list either before or after an append that runs concurrently on
 * another context; one appending thread per index.
 * While an index has room left its store has holes between the lists, so searches go list by list: nprobe >= number of
 * lists probes every list (up to 128 lists on the device, more through vs_search's host path); the whole-store entry
 * points (vs_index_search_batch_dev, the GEMM batch form of vs_search) are for indexes without holes.
 * vs_index_rows = rows present, vs_index_capacity = places, vs_index_list_lengths = rows per list (C values). */
VS_API int vs_index_with_room(vs_ctx *ctx, const vs_index *ix, size_t percent, size_t min_rows, vs_index **out);
VS_API int vs_index_append(vs_ctx *ctx, vs_index *ix, const uint8_t *rows_packed, size_t n, size_t row_bytes,
                    const uint64_t *doc_ids, int64_t *assign_out);
VS_API size_t vs_index_capacity(const vs_index *ix);
VS_API int vs_index_list_lengths(vs_ctx *ctx, const vs_index *ix, uint64_t *lengths_out);
/* Search (search.go:115-273 minus embedding/DB hops): nq query rows (row776, host), nprobe =
 * SearchRequest.Centroids (>= number of lists means "all"), k = Count+Offset.  Outputs (host):
 * ids_out[nq*k] document IDs, sims_out[nq*k] float32 similarities, counts_out[nq] valid entries.
 * Order: similarity desc (float32), then document ID asc; one entry per document (search.go:260-268), removed before the
 * list is cut to k like the reference does (search.go:259-271).  Any nq (batches above 4096 run in turns), any nprobe and
 * any k: up to 128 probed lists (or all of them) and up to 128 hits run in the fused kernels (one query: one launch,
 * csrc/fused.cu; a batch: list-major, csrc/listmajor.cu); wider requests (search.go:116-122 accepts any Centroids value,
 * and Count+Offset is unbounded by Offset) are served exactly through full similarity vectors and a host-side cut. */
VS_API int vs_search(vs_ctx *ctx, const vs_index *ix, const uint8_t *queries_packed, size_t nq, size_t nprobe, size_t k,
              uint64_t *ids_out, float *sims_out, int32_t *counts_out);
/* Brute force over a matrix (BASELINE config 1): same contract, ids = doc_ids[row] or row index. */
/* (vs_search with nprobe >= lists and vs_search_flat hand batches of >= 64 queries over >= 65536 rows to the GEMM
 * path below; smaller calls run one streaming scan per query.) */
VS_API int vs_search_flat(vs_ctx *ctx, const vs_matrix *m, const uint64_t *d_doc_ids, const uint8_t *queries_packed,
                   size_t nq, size_t k, uint64_t *ids_out, float *sims_out, int32_t *counts_out);
/* Query batches over the whole store as an int8 tensor-core GEMM (tcgen05.mma kind::i8) with a fused per-pair
 * filter (BASELINE config 3): the loop of per-query scans of search.go:241-273 for nq queries at once.  Same
 * contract and the same bits as vs_search_flat.  d <= 1024, nq <= 4096, k <= 128.  Queries the filter cannot
 * answer (degenerate header, fewer than k documents above its threshold) are finished by the streaming scan. */
VS_API int vs_search_flat_gemm(vs_ctx *ctx, const vs_matrix *m, const uint64_t *d_doc_ids, const uint8_t *queries_packed,
                        size_t nq, size_t k, uint64_t *ids_out, float *sims_out, int32_t *counts_out);
/* Device-resident form: queries are a device matrix, results stay in device buffers (d_ids[nq*k], d_sims[nq*k],
 * d_counts[nq]); ids = d_doc_ids[row] or id_base + row.  Synchronizes the ctx stream internally (candidate count).
 * stats_out (nullable, host uint64[8]): candidates emitted, queries finished by the scan, store tiles, sampled tiles,
 * then microseconds of the pre-pass, the filtering GEMM and the candidate resolution (CUDA events); [7] unused.
 * With vs_ctx_profile_enable the filtering GEMM launch is the kernel vs_ctx_profile_read reports. */
VS_API int vs_search_batch_dev(vs_ctx *ctx, const vs_matrix *m, const uint64_t *d_doc_ids, uint64_t id_base,
                        const vs_matrix *queries, size_t k, uint64_t *d_ids, float *d_sims, int32_t *d_counts,
                        uint64_t *stats_out);
/* The same over the whole store of an index (nprobe = "all lists"), ids = the index's document ids. */
VS_API int vs_index_search_batch_dev(vs_ctx *ctx, const vs_index *ix, const vs_matrix *queries, size_t k, uint64_t *d_ids,
                              float *d_sims, int32_t *d_counts, uint64_t *stats_out);
/* Device-resident search: queries already a device matrix; results stay in device buffers
 * (d_ids[nq*k], d_sims[nq*k], d_counts[nq], d_status[nq]); asynchronous on the ctx stream.
 * d_status bit0/bit1 = a float32 rounding could not be certified in the probe/list stage, bit2 =
 * fewer than k unique documents among the kept candidates: call vs_search_resolve to finish those
 * queries with literal reference arithmetic. */
VS_API int vs_search_dev(vs_ctx *ctx, const vs_index *ix, const vs_matrix *queries, size_t nprobe, size_t k,
                  uint64_t *d_ids, float *d_sims, int32_t *d_counts, uint32_t *d_status);
VS_API int vs_search_resolve(vs_ctx *ctx, const vs_index *ix, const vs_matrix *queries, size_t nprobe, size_t k,
                      uint64_t *d_ids, float *d_sims, int32_t *d_counts, uint32_t *d_status, int *n_resolved_out);
/* The two stages of vs_search_dev as separate calls, for an index striped over several devices (every device holds the
 * whole centroid table and 1/G of every posting list).  Device g runs vs_probe_dev on ITS share of the batch -- the
 * centroid stage of server/search.go:205-223: d_probe[nq*nprobe] list numbers in rank order, d_status[nq] the stage's
 * status bits -- the shares are exchanged (an all-gather of nq*(nprobe+1) words; the caller's transport), and every
 * device runs vs_search_dev_probed on the WHOLE batch with the gathered lists: the posting-list stage of
 * search.go:239-273 on its stripe.  d_status holds the gathered probe-stage words on entry and is left as vs_search_dev
 * leaves it.  The hits are the ones vs_search_dev returns (the probe stage is deterministic); needs 1 <= nprobe < lists. */
VS_API int vs_probe_dev(vs_ctx *ctx, const vs_index *ix, const vs_matrix *queries, size_t nprobe, uint32_t *d_probe,
                 uint32_t *d_status);
VS_API int vs_search_dev_probed(vs_ctx *ctx, const vs_index *ix, const vs_matrix *queries, size_t nprobe, size_t k,
                         const uint32_t *d_probe, uint64_t *d_ids, float *d_sims, int32_t *d_counts, uint32_t *d_status);
/* Stage 1 only (search.go:202-227): probe_out[nq*min(nprobe,C)] list indices in rank order (host). */
VS_API int vs_select_probes(vs_ctx *ctx, const vs_index *ix, const uint8_t *queries_packed, size_t nq, size_t nprobe,
                     uint32_t *probe_out, float *probe_sims_out);
/* ---- one host process, several GPUs (csrc/sharded.cu) ---------------------------------------------------------------
 * The reference is one Go process (main.go:31) that runs one search per goroutine (server/search.go:115): it cannot be
 * split into a process per GPU.  A vs_sharded handle lets that one process drive G devices of one NVLink / NVSwitch box:
 * rows are striped over the devices by primary key (key % G, so every posting list is split evenly for any probe set),
 * the centroid table is replicated, every device answers a batch on its stripe, and device 0 merges the shard-local
 * top-k lists, reading its peers' hit buffers in place over NVLink (peer access; no collective library, no staging copy).
 * Same results as one vs_index holding all the rows.  `devices` may name a device more than once (several stripes on one
 * GPU).  Calls on one handle must not overlap; different handles are independent. */
typedef struct vs_sharded vs_sharded;
VS_API int vs_sharded_create(const int *devices, size_t G, vs_sharded **out);
VS_API void vs_sharded_release(vs_sharded *sh);
VS_API size_t vs_sharded_rows(const vs_sharded *sh);
VS_API size_t vs_sharded_shards(const vs_sharded *sh);
VS_API size_t vs_sharded_shard_rows(const vs_sharded *sh, size_t shard);
/* Replaces the loader side of server/search.go:241-243 for G devices: rows in primary-key order (row i has key i) with
 * Embedding.CentroidID (database/model.go:16) as list_of_row; doc_ids NULL = the primary key is the document id. */
VS_API int vs_sharded_build_assigned(vs_sharded *sh, const uint8_t *rows776, size_t n, size_t row_bytes, const uint64_t *doc_ids,
                                     const uint32_t *list_of_row, const uint8_t *centroids776, size_t C);
/* Replaces server/upload.go:239-279 on the striped store: the new rows take the next primary keys, each joins the list
 * of its nearest centroid (upload.go:245) on the shard that owns its key.  assign_out (optional): int64[n]. */
VS_API int vs_sharded_upload(vs_sharded *sh, const uint8_t *rows776, size_t n, size_t row_bytes, const uint64_t *doc_ids,
                             int64_t *assign_out);
/* Replaces server/search.go:202-273 for nq queries (host buffers in and out, like vs_search); one search at a time. */
VS_API int vs_sharded_search(vs_sharded *sh, const uint8_t *queries776, size_t nq, size_t nprobe, size_t k, uint64_t *ids_out,
                             float *sims_out, int32_t *counts_out);
/* One search context over all the devices of a handle (a stream + scratch per device, merge buffers on device 0): what
 * one `calculate` closure / goroutine holds (server/search.go:230; compute/cosine.go:60-66 for the single-device form).
 * Searches through different contexts may overlap; uploads must not overlap searches. */
typedef struct vs_sharded_ctx vs_sharded_ctx;
VS_API int vs_sharded_ctx_create(vs_sharded *sh, vs_sharded_ctx **out);
VS_API void vs_sharded_ctx_destroy(vs_sharded_ctx *sc);
VS_API int vs_sharded_search_ctx(vs_sharded_ctx *sc, const uint8_t *queries776, size_t nq, size_t nprobe, size_t k, uint64_t *ids_out,
                                 float *sims_out, int32_t *counts_out);

/* Multi-GPU (row-striped shards): merge G gathered shard-local results per query into the global
 * top-k.  d_ids_in/d_sims_in [G][nq][k], d_counts_in [G][nq] device arrays. */
VS_API int vs_topk_merge_dev(vs_ctx *ctx, const uint64_t *d_ids_in, const float *d_sims_in, const int32_t *d_counts_in,
                      size_t G, size_t nq, size_t k, uint64_t *d_ids_out, float *d_sims_out, int32_t *d_counts_out);

/* Same for shard results packed one buffer per rank (a single all-gather): rank g's buffer starts at
 * d_packed + g*rank_stride_bytes and holds ids[nq*k] (uint64) at ids_off, sims[nq*k] (float32) at sims_off and
 * counts[nq] (int32) at counts_off (byte offsets, 8-byte aligned). */
VS_API int vs_topk_merge_packed_dev(vs_ctx *ctx, const void *d_packed, size_t rank_stride_bytes, size_t ids_off,
                                    size_t sims_off, size_t counts_off, size_t G, size_t nq, size_t k, uint64_t *d_ids_out,
                                    float *d_sims_out, int32_t *d_counts_out);

/* ---- dnc/k_means.go:67-117 and dnc/dnc.go:417-449 ---------------------------------- */
/* One Lloyd iteration.  data: device matrix; centroids_packed: k rows (host); means: [k][D] float32
 * (host, in/out: the state k_means.go:60-65 carries; an empty cluster keeps its previous mean).
 * Outputs (host): assign_out[n] (nullable), counts_out[k], new_centroids_out[k*(8+D)], *converged_out. */
VS_API int vs_kmeans_step(vs_ctx *ctx, const vs_matrix *data, const uint8_t *centroids_packed, size_t k, float *means,
                   int64_t *assign_out, int64_t *counts_out, uint8_t *new_centroids_out, int *converged_out);
/* kMeans (dnc/k_means.go:19-212) with every iteration on the device.  superset_rows[ks] = the distinct random data rows
 * the reference draws at :35-44 (ks = min(n, 5k)); the superset is iterated until the centroids' code bytes stop
 * changing or iter_limit iterations (:67-117), its first k centroids and their float32 means are kept (:125-154), and
 * the set is iterated the same way (:157-207).  centroids_out: k rows of 8+D bytes (host).
 * stats_out (nullable, [4]): iterations of the superset phase, of the set phase, microseconds spent assigning, updating.
 * The early returns of :20-26 (k <= 0, n <= k) are the binding's: they need no device work. */
VS_API int vs_kmeans(vs_ctx *ctx, const vs_matrix *data, size_t k, const uint64_t *superset_rows, size_t ks, size_t iter_limit,
              uint8_t *centroids_out, int64_t *stats_out);
/* One Lloyd iteration over a store cut into contiguous row blocks, one per GPU, bit-identical to one
 * device.  Every rank assigns its rows with vs_argmax_MxN_dev at the same time; vs_kmeans_accumulate_dev then continues
 * the per-centroid float32 sums [k][D] and int64 counts [k] (device buffers; zeros before the first block) over this
 * block's rows in row order, and the buffers travel to the next rank (NCCL send/recv).  After the last block
 * vs_kmeans_finish_dev turns them into the new float32 means (d_means in/out; an empty cluster keeps its previous
 * mean), the new centroid matrix and the convergence flag of k_means.go:102-108. */
VS_API int vs_kmeans_accumulate_dev(vs_ctx *ctx, const vs_matrix *data, size_t k, const int32_t *d_assign, float *d_sums,
                             int64_t *d_counts);
VS_API int vs_kmeans_finish_dev(vs_ctx *ctx, const vs_matrix *centroids, const float *d_sums, const int64_t *d_counts,
                         float *d_means, vs_matrix **new_centroids_out, int *converged_out);
/* Device-side pieces of the divide-and-conquer centroid build (dnc/dnc.go:300-400; the recursion itself is host logic,
 * see go-vectorsearch_b200/dnc.py DivideAndConquer).
 * vs_matrix_gather: a new matrix of the given rows of src in the given order (sample(), dnc/sampling.go:12-74).
 * vs_matrix_split_dev: stable partition of src by a device int32 assignment (dnc.go:363-389): children_out[j] gets the
 *   rows assigned to j in their original order (null when none), counts_out[j] their number.
 * vs_recenter_clusters_dev: recenterDbCentroid (dnc.go:402-456) for all k clusters of an assignment at once. */
VS_API int vs_matrix_gather(vs_ctx *ctx, const vs_matrix *src, const uint64_t *rows, size_t n, vs_matrix **out);
VS_API int vs_matrix_split_dev(vs_ctx *ctx, const vs_matrix *src, const int32_t *d_assign, size_t k, vs_matrix **children_out,
                        uint64_t *counts_out);
/* Host-pointer forms: vs_matrix_split = nearest of the k packed centroids for every row of src, then the stable
 * partition (the whole split loop of dnc.go:363-389); vs_reassign_recenter = the tail of KMeansDivideAndConquer
 * (dnc.go:177-291): every row to its nearest of the k new centroids (assignment to assign_out on the host and/or
 * d_assign_out on the device, both nullable), every centroid re-centred on its members. */
VS_API int vs_matrix_split(vs_ctx *ctx, const vs_matrix *src, const uint8_t *centroids_packed, size_t k, vs_matrix **children_out,
                    uint64_t *counts_out);
VS_API int vs_reassign_recenter(vs_ctx *ctx, const vs_matrix *data, const uint8_t *centroids_packed, size_t k, int32_t *assign_out,
                         int32_t *d_assign_out, uint8_t *centroids_out, int64_t *counts_out);
VS_API int vs_recenter_clusters_dev(vs_ctx *ctx, const vs_matrix *data, const int32_t *d_assign, size_t k, uint8_t *centroids_out,
                             int64_t *counts_out);
/* recenterDbCentroid (dnc.go:417-449): float64 mean of all rows of m in row order -> row776. */
VS_API int vs_recenter(vs_ctx *ctx, const vs_matrix *m, uint8_t *out_row);

/* ---- one process per GPU: the exchange of the shard-local hits inside the merge kernel -----------------
 * Replaces the all-gather (a collective launch per step for 32 KB per rank) of the striped index.  Every rank
 * allocates two hit slots + one flag word per rank (vs_exchange_create; slot_bytes = the packed [ids | sims | counts]
 * buffer of one step), publishes the allocation's CUDA IPC handle (vs_exchange_handle: 64 bytes, all-gathered by the
 * caller once over whatever channel the processes share) and maps the other ranks' allocations (vs_exchange_connect; a barrier between
 * create and connect is the caller's).  Per step every rank lets its search write the hits into
 * vs_exchange_slot(x, step & 1) and calls vs_exchange_merge on the same context: one warp signals the peers over NVLink
 * (release store of the step number into their flag words) and waits for their signals, then the merge kernel reads
 * their hits in place and writes the merged top k (order and one-hit-per-document rule of vs_search).  step = 1, 2, 3, ... identically on every
 * rank.  A peer that never arrives makes the kernel trap after ~10 s instead of hanging the device. */
typedef struct vs_exchange vs_exchange;
VS_API int vs_exchange_create(int rank, int world, size_t slot_bytes, vs_exchange **out);
VS_API int vs_exchange_handle(const vs_exchange *x, void *handle_out, size_t cap);
VS_API int vs_exchange_connect(vs_exchange *x, const void *handles);
VS_API void *vs_exchange_slot(const vs_exchange *x, int slot);
VS_API int vs_exchange_merge(vs_ctx *ctx, vs_exchange *x, uint32_t step, size_t ids_off, size_t sims_off, size_t counts_off,
                      size_t nq, size_t k, uint64_t *d_ids_out, float *d_sims_out, int32_t *d_counts_out);
VS_API void vs_exchange_release(vs_exchange *x);

#ifdef __cplusplus
}
#endif
#endif /* VSCUDA_H */
