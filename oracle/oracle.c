/*
 * oracle.c -- CPU restatement of go-vectorsearch's similarity-search hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (go-vectorsearch_b200/) never links, imports or calls anything in this directory.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors, and its Go
 * toolchain is absent from this image, so this restatement cannot be checked against
 * outputs of the reference itself.  It is pinned instead by (1) hand-derived known-answer
 * vectors in tests/golden/, (2) an independent numpy restatement (oracle/oracle_np.py)
 * that must agree bit-for-bit, both written from the cited reference lines.
 *
 * Build (oracle/Makefile): gcc -O2 -ffp-contract=off -fno-fast-math  (no FMA contraction,
 * SSE2 scalar IEEE arithmetic, left-to-right float64 accumulation) -- this models the
 * reference's DEFAULT backend (build tag !gonum && !gorgonia) as compiled by `go run .`
 * for GOAMD64=v1, which never fuses x*y+z.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORA_API __attribute__((visibility("default")))

/* The same source is compiled a second time as _build/liboracle_blas_proxy.so (oracle/Makefile: -O3 -ffast-math
 * -DORA_VECTORIZED_PROXY): there the compiler may reorder and vectorize the float64 sums and 1xN scoring is cloned
 * for AVX2+FMA with run-time dispatch.  That build is a TIMING PROXY for the reference's `gonum` build tag (BLAS
 * Dnrm2 / Dscal / Ddot, compute/cosine_gonum.go; gonum v0.16.0 is not in the image): a vectorised float64 CPU path
 * with an unspecified summation order, reported beside the exact build.  No parity claim rests on it. */
#ifdef ORA_VECTORIZED_PROXY
#define ORA_HOT __attribute__((target_clones("avx2,fma", "default")))
#else
#define ORA_HOT
#endif

/* Go's uint8(f) on amd64: CVTTSS2SL / CVTTSD2SQ then keep the low byte; NaN and
 * out-of-range inputs produce the "integer indefinite" pattern whose low byte is 0.
 * (compute/quantization.go:30,43: `valueQuantized = uint8(normalized * 255)`) */
static inline uint8_t go_u8_from_f32(float x) {
    int32_t i;
    if (x != x || x >= 2147483648.0f || x < -2147483648.0f) i = INT32_MIN;
    else i = (int32_t)x; /* truncation toward zero */
    return (uint8_t)(uint32_t)i;
}
static inline uint8_t go_u8_from_f64(double x) {
    int64_t i;
    if (x != x || x >= 9223372036854775808.0 || x < -9223372036854775808.0) i = INT64_MIN;
    else i = (int64_t)x;
    return (uint8_t)(uint64_t)i;
}

/* compute/quantization.go:21-32 QuantizeFloat32 */
ORA_API uint8_t ora_quantize_f32(float value, float min, float max) {
    if (value < min) value = min;
    else if (value > max) value = max;
    float normalized = (value - min) / (max - min);
    return go_u8_from_f32(normalized * 255);
}

/* compute/quantization.go:34-45 QuantizeFloat64 */
ORA_API uint8_t ora_quantize_f64(double value, double min, double max) {
    if (value < min) value = min;
    else if (value > max) value = max;
    double normalized = (value - min) / (max - min);
    return go_u8_from_f64(normalized * 255);
}

/* compute/quantization.go:55-61 DequantizeFloat32 */
ORA_API float ora_dequantize_f32(uint8_t q, float min, float max) {
    float normalized = (float)q / 255.0f;
    float range = max - min;
    float scaled = normalized * range; /* separate statement: no FMA */
    return min + scaled;
}

/* compute/quantization.go:63-69 DequantizeFloat64 */
ORA_API double ora_dequantize_f64(uint8_t q, double min, double max) {
    double normalized = (double)q / 255.0;
    double range = max - min;
    double scaled = normalized * range;
    return min + scaled;
}

/* compute/quantization.go:194-204 rangeFloat32: min and max both seeded at 0 (named results) */
ORA_API void ora_range_f32(const float *v, size_t d, float *pmin, float *pmax) {
    float mn = 0, mx = 0;
    for (size_t i = 0; i < d; i++) {
        if (v[i] < mn) mn = v[i];
        if (v[i] > mx) mx = v[i];
    }
    *pmin = mn; *pmax = mx;
}

/* compute/quantization.go:206-216 rangeFloat64 */
ORA_API void ora_range_f64(const double *v, size_t d, double *pmin, double *pmax) {
    double mn = 0, mx = 0;
    for (size_t i = 0; i < d; i++) {
        if (v[i] < mn) mn = v[i];
        if (v[i] > mx) mx = v[i];
    }
    *pmin = mn; *pmax = mx;
}

static inline void put_f32_le(uint8_t *p, float f) { uint32_t u; memcpy(&u, &f, 4); p[0] = u; p[1] = u >> 8; p[2] = u >> 16; p[3] = u >> 24; }
static inline float get_f32_le(const uint8_t *p) { uint32_t u = (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; float f; memcpy(&f, &u, 4); return f; }

/* compute/quantization.go:82-91 QuantizeVectorFloat32 -> out has 8+d bytes */
ORA_API void ora_quantize_vector_f32(const float *v, size_t d, uint8_t *out) {
    float mn, mx;
    ora_range_f32(v, d, &mn, &mx);
    put_f32_le(out, mn);
    put_f32_le(out + 4, mx);
    for (size_t i = 0; i < d; i++) out[8 + i] = ora_quantize_f32(v[i], mn, mx);
}

/* compute/quantization.go:93-102 QuantizeVectorFloat64: header is float32(min), float32(max),
 * codes use the unrounded float64 range */
ORA_API void ora_quantize_vector_f64(const double *v, size_t d, uint8_t *out) {
    double mn, mx;
    ora_range_f64(v, d, &mn, &mx);
    put_f32_le(out, (float)mn);
    put_f32_le(out + 4, (float)mx);
    for (size_t i = 0; i < d; i++) out[8 + i] = ora_quantize_f64(v[i], mn, mx);
}

/* compute/quantization.go:142-156 QuantizeMatrixFloat32/64 over packed rows */
ORA_API void ora_quantize_matrix_f32(const float *m, size_t n, size_t d, uint8_t *out) {
    for (size_t i = 0; i < n; i++) ora_quantize_vector_f32(m + i * d, d, out + i * (8 + d));
}
ORA_API void ora_quantize_matrix_f64(const double *m, size_t n, size_t d, uint8_t *out) {
    for (size_t i = 0; i < n; i++) ora_quantize_vector_f64(m + i * d, d, out + i * (8 + d));
}

/* compute/quantization.go:114-122 DequantizeVectorFloat32 */
ORA_API void ora_dequantize_vector_f32(const uint8_t *row, size_t row_bytes, float *out) {
    float mn = get_f32_le(row), mx = get_f32_le(row + 4);
    for (size_t i = 8; i < row_bytes; i++) out[i - 8] = ora_dequantize_f32(row[i], mn, mx);
}

/* compute/quantization.go:124-132 DequantizeVectorFloat64 (header widened from float32) */
ORA_API void ora_dequantize_vector_f64(const uint8_t *row, size_t row_bytes, double *out) {
    double mn = (double)get_f32_le(row), mx = (double)get_f32_le(row + 4);
    for (size_t i = 8; i < row_bytes; i++) out[i - 8] = ora_dequantize_f64(row[i], mn, mx);
}

/* compute/quantization.go:166-180 DequantizeMatrixFloat32/64 over packed rows */
ORA_API void ora_dequantize_matrix_f32(const uint8_t *rows, size_t n, size_t row_bytes, float *out) {
    for (size_t i = 0; i < n; i++) ora_dequantize_vector_f32(rows + i * row_bytes, row_bytes, out + i * (row_bytes - 8));
}
ORA_API void ora_dequantize_matrix_f64(const uint8_t *rows, size_t n, size_t row_bytes, double *out) {
    for (size_t i = 0; i < n; i++) ora_dequantize_vector_f64(rows + i * row_bytes, row_bytes, out + i * (row_bytes - 8));
}

/* compute/cosine.go:138-149 normalizeVector */
static void normalize_vector(double *vec, size_t d) {
    double norm = 0;
    for (size_t i = 0; i < d; i++) {
        double sq = vec[i] * vec[i];
        norm += sq;
    }
    norm = sqrt(norm);
    if (norm != 0) {
        for (size_t i = 0; i < d; i++) vec[i] /= norm;
    }
}

/* The integer dot product the device kernels are built on (SURVEY 8a note): not a
 * reference function, but the "integer dot products bit-exact" clause of north_star. */
ORA_API void ora_dot_u8_1xN(const uint8_t *q, const uint8_t *rows, size_t n, size_t row_bytes, uint32_t *dots) {
    for (size_t i = 0; i < n; i++) {
        uint32_t s = 0;
        const uint8_t *r = rows + i * row_bytes;
        for (size_t j = 8; j < row_bytes; j++) s += (uint32_t)q[j] * (uint32_t)r[j];
        dots[i] = s;
    }
}

/* compute/compute.go:10-21 NewVector + compute/compute.go:23-44 NewMatrix +
 * compute/cosine.go:13-57 (*vectorContainer).MatrixCosineSimilarity.
 * q: 8+d bytes; rows: n packed rows of row_bytes. Returns 0, or -1 for the reference's
 * panic cases (empty vector / empty matrix), -2 for the Fatalf dimension mismatch
 * (q_bytes != row_bytes). */
ORA_API ORA_HOT int ora_cosine_1xN(const uint8_t *q, size_t q_bytes, const uint8_t *rows, size_t n, size_t row_bytes, float *sims) {
    if (q_bytes <= 8) return -1;        /* compute.go:12-14 */
    if (n == 0) return -1;              /* compute.go:25-27 */
    if (row_bytes <= 8) return -1;      /* compute.go:29-31 */
    if (q_bytes != row_bytes) return -2; /* cosine.go:19-21 */
    size_t d = row_bytes - 8;
    double *A = (double *)malloc(d * sizeof(double));
    double *B = (double *)malloc(d * sizeof(double));
    ora_dequantize_vector_f64(q, q_bytes, A);
    normalize_vector(A, d); /* cosine.go:26 */
    for (size_t i = 0; i < n; i++) {
        ora_dequantize_vector_f64(rows + i * row_bytes, row_bytes, B);
        normalize_vector(B, d); /* cosine.go:29-33 */
        double dot = 0;          /* cosine.go:45-48 */
        for (size_t j = 0; j < d; j++) {
            double p = A[j] * B[j];
            dot += p;
        }
        sims[i] = (float)dot;    /* cosine.go:50 */
    }
    free(A); free(B);
    return 0;
}

/* compute/cosine.go:70-125 (*matrixContainer).MatrixCosineSimilarity: receiver = centroids
 * (m rows), argument = data (n rows). For each data row: argmax over centroids with strict
 * '>' (lowest index wins ties), maxVal seeded at -1.0, maxIdx at 0. */
ORA_API int ora_argmax_MxN(const uint8_t *cent, size_t m, size_t cent_bytes, const uint8_t *rows, size_t n,
                           size_t row_bytes, float *sims, int64_t *argmax) {
    if (m == 0 || n == 0) return -1;
    if (cent_bytes <= 8 || row_bytes <= 8) return -1;
    if (cent_bytes != row_bytes) return -2; /* cosine.go:77-79 */
    size_t d = row_bytes - 8;
    double *A = (double *)malloc(m * d * sizeof(double));
    double *B = (double *)malloc(d * sizeof(double));
    for (size_t j = 0; j < m; j++) {
        ora_dequantize_vector_f64(cent + j * cent_bytes, cent_bytes, A + j * d);
        normalize_vector(A + j * d, d); /* cosine.go:85-88 */
    }
    for (size_t i = 0; i < n; i++) {
        ora_dequantize_vector_f64(rows + i * row_bytes, row_bytes, B);
        normalize_vector(B, d); /* cosine.go:90-93 */
        double maxVal = -1.0;
        int64_t maxIdx = 0;
        for (size_t j = 0; j < m; j++) {
            const double *Arow = A + j * d;
            double dot = 0;
            for (size_t k = 0; k < d; k++) {
                double p = Arow[k] * B[k];
                dot += p;
            }
            if (dot > maxVal) { maxVal = dot; maxIdx = (int64_t)j; }
        }
        if (sims) sims[i] = (float)maxVal;
        argmax[i] = maxIdx;
    }
    free(A); free(B);
    return 0;
}

/* ---- server/search.go:202-273, restated on in-memory arrays -------------------------
 * Tie rule (this repo's contract; the reference's slices.SortFunc is unstable so ties are
 * undefined upstream): similarity desc compared as float32 with cmp.Compare semantics
 * (NaN sorts below every number), then ID asc. */
typedef struct { float sim; uint64_t id; } ora_hit;

static int hit_cmp(const void *pa, const void *pb) {
    const ora_hit *a = (const ora_hit *)pa, *b = (const ora_hit *)pb;
    int an = a->sim != a->sim, bn = b->sim != b->sim;
    if (an || bn) { /* cmp.Compare(b,a): NaN < everything, NaN == NaN */
        if (an && !bn) return 1;
        if (!an && bn) return -1;
    } else {
        if (a->sim > b->sim) return -1;
        if (a->sim < b->sim) return 1;
    }
    if (a->id < b->id) return -1;
    if (a->id > b->id) return 1;
    return 0;
}

/* search.go:202-227: score every centroid, sort desc, keep first min(nprobe, C).
 * probe_out receives centroid indices (table order stands in for centroid.ID). */
ORA_API int ora_select_probes(const uint8_t *q, const uint8_t *centroids, size_t C, size_t row_bytes,
                              size_t nprobe, uint32_t *probe_out, float *probe_sims_out) {
    if (C == 0) return 0; /* search.go:197-199 */
    float *sims = (float *)malloc(C * sizeof(float));
    int rc = ora_cosine_1xN(q, row_bytes, centroids, C, row_bytes, sims); /* search.go:214 */
    if (rc) { free(sims); return rc; }
    ora_hit *h = (ora_hit *)malloc(C * sizeof(ora_hit));
    for (size_t i = 0; i < C; i++) { h[i].sim = sims[i]; h[i].id = i; }
    qsort(h, C, sizeof(ora_hit), hit_cmp);  /* search.go:220-222 */
    size_t keep = nprobe < C ? nprobe : C;  /* search.go:223 */
    for (size_t i = 0; i < keep; i++) { probe_out[i] = (uint32_t)h[i].id; if (probe_sims_out) probe_sims_out[i] = h[i].sim; }
    free(h); free(sims);
    return (int)keep;
}

/* search.go:239-273: stream the rows whose list is probed, in row (primary-key) order,
 * BATCH_SIZE_DATABASE=1000 at a time (config/constants.go:6); per batch score, append,
 * sort, dedup by document keeping the best, truncate to k = Count+Offset.
 * list_of_row[i] = centroid index of row i; doc_ids[i] = documentID of row i.
 * Returns number of hits written (<= k). */
ORA_API int ora_search(const uint8_t *q, const uint8_t *centroids, size_t C, const uint8_t *rows, size_t n,
                       size_t row_bytes, const uint32_t *list_of_row, const uint64_t *doc_ids, size_t nprobe,
                       size_t k, uint64_t *ids_out, float *sims_out) {
    if (C == 0) return 0;
    uint32_t *probes = (uint32_t *)malloc((nprobe < C ? nprobe : C) * sizeof(uint32_t) + 4);
    int np = ora_select_probes(q, centroids, C, row_bytes, nprobe, probes, NULL);
    if (np < 0) { free(probes); return np; }
    uint8_t *probed = (uint8_t *)calloc(C, 1);
    for (int i = 0; i < np; i++) probed[probes[i]] = 1;
    const size_t BATCH = 1000;
    uint8_t *batch = (uint8_t *)malloc(BATCH * row_bytes);
    uint64_t *bdoc = (uint64_t *)malloc(BATCH * sizeof(uint64_t));
    float *bsim = (float *)malloc(BATCH * sizeof(float));
    ora_hit *closest = (ora_hit *)malloc((k + BATCH) * sizeof(ora_hit));
    size_t nclosest = 0, nb = 0;
    for (size_t i = 0; i <= n; i++) {
        if (i < n && probed[list_of_row[i]]) {
            memcpy(batch + nb * row_bytes, rows + i * row_bytes, row_bytes);
            bdoc[nb++] = doc_ids[i];
        }
        if (nb == BATCH || (i == n && nb > 0)) {
            ora_cosine_1xN(q, row_bytes, batch, nb, row_bytes, bsim);            /* search.go:249 */
            for (size_t j = 0; j < nb; j++) { closest[nclosest].sim = bsim[j]; closest[nclosest].id = bdoc[j]; nclosest++; }
            qsort(closest, nclosest, sizeof(ora_hit), hit_cmp);                    /* search.go:256-258 */
            size_t w = 0;                                                           /* search.go:260-268 */
            for (size_t j = 0; j < nclosest; j++) {
                int seen = 0;
                for (size_t t = 0; t < w; t++) if (closest[t].id == closest[j].id) { seen = 1; break; }
                if (!seen) { closest[w++] = closest[j]; if (w == k) break; }       /* :270 truncate */
            }
            nclosest = w;
            nb = 0;
        }
    }
    for (size_t j = 0; j < nclosest; j++) { ids_out[j] = closest[j].id; sims_out[j] = closest[j].sim; }
    free(probes); free(probed); free(batch); free(bdoc); free(bsim); free(closest);
    return (int)nclosest;
}

/* Brute force: config 1 of BASELINE.json (every row scored, no centroid stage). Same
 * list-stage logic as ora_search with all rows probed. */
ORA_API int ora_search_flat(const uint8_t *q, const uint8_t *rows, size_t n, size_t row_bytes,
                            const uint64_t *doc_ids, size_t k, uint64_t *ids_out, float *sims_out) {
    uint8_t cent[16];
    (void)cent;
    float *sims = (float *)malloc((n ? n : 1) * sizeof(float));
    ora_hit *h = (ora_hit *)malloc((n ? n : 1) * sizeof(ora_hit));
    if (n == 0) { free(sims); free(h); return 0; }
    int rc = ora_cosine_1xN(q, row_bytes, rows, n, row_bytes, sims);
    if (rc) { free(sims); free(h); return rc; }
    for (size_t i = 0; i < n; i++) { h[i].sim = sims[i]; h[i].id = doc_ids ? doc_ids[i] : i; }
    qsort(h, n, sizeof(ora_hit), hit_cmp);
    size_t w = 0;
    for (size_t j = 0; j < n && w < k; j++) {
        int seen = 0;
        for (size_t t = 0; t < w; t++) if (ids_out[t] == h[j].id) { seen = 1; break; }
        if (!seen) { ids_out[w] = h[j].id; sims_out[w] = h[j].sim; w++; }
    }
    free(sims); free(h);
    return (int)w;
}

/* ---- dnc/k_means.go:67-117: ONE Lloyd iteration -------------------------------------
 * data: n rows; centroids: k rows (current); means: [k][d] float32 state carried between
 * iterations (k_means.go:60-65; an empty cluster keeps its previous mean, :90-92).
 * Outputs: assign[n] (centroid index per row), counts[k], new_centroids[k][row_bytes],
 * returns 1 if converged (every centroid's code bytes [8:] unchanged, :102-108) else 0. */
ORA_API int ora_kmeans_step(const uint8_t *data, size_t n, const uint8_t *centroids, size_t k, size_t row_bytes,
                            float *means, int64_t *assign, int64_t *counts, uint8_t *new_centroids) {
    size_t d = row_bytes - 8;
    int rc = ora_argmax_MxN(centroids, k, row_bytes, data, n, row_bytes, NULL, assign); /* :73-77 */
    if (rc) return rc;
    float *sums = (float *)calloc(k * d, sizeof(float));
    float *vec = (float *)malloc(d * sizeof(float));
    for (size_t j = 0; j < k; j++) counts[j] = 0;
    for (size_t i = 0; i < n; i++) {                                /* :80-86 */
        ora_dequantize_vector_f32(data + i * row_bytes, row_bytes, vec);
        float *s = sums + (size_t)assign[i] * d;
        for (size_t j = 0; j < d; j++) s[j] += vec[j];
        counts[assign[i]]++;
    }
    for (size_t c = 0; c < k; c++) {                                /* :89-96 */
        if (counts[c] <= 0) continue;
        float fc = (float)counts[c];
        for (size_t j = 0; j < d; j++) means[c * d + j] = sums[c * d + j] / fc;
    }
    ora_quantize_matrix_f32(means, k, d, new_centroids);            /* :99 */
    int converged = 1;                                              /* :102-108 */
    for (size_t c = 0; c < k; c++)
        if (memcmp(new_centroids + c * row_bytes + 8, centroids + c * row_bytes + 8, d) != 0) { converged = 0; break; }
    free(sums); free(vec);
    return converged;
}

/* ---- dnc/dnc.go:417-449 recenterDbCentroid: float64 sum of dequantized member rows in
 * row order, divided by count, QuantizeVectorFloat64. count==0 divides by zero (NaN -> codes 0). */
ORA_API void ora_recenter(const uint8_t *rows, size_t n, size_t row_bytes, uint8_t *out) {
    size_t d = row_bytes - 8;
    double *sum = (double *)calloc(d ? d : 1, sizeof(double));
    double *vec = (double *)malloc((d ? d : 1) * sizeof(double));
    uint64_t count = 0;
    for (size_t i = 0; i < n; i++) {
        ora_dequantize_vector_f64(rows + i * row_bytes, row_bytes, vec);
        for (size_t j = 0; j < d; j++) sum[j] += vec[j];
        count++;
    }
    double fc = (double)count;
    for (size_t j = 0; j < d; j++) sum[j] = sum[j] / fc;
    ora_quantize_vector_f64(sum, d, out);
    free(sum); free(vec);
}

/* ---- timing helper for bench.py's cpu_baseline / --impl reference legs: the same
 * per-query arithmetic run over independent queries on `threads` pthreads (the Go server
 * runs one goroutine per in-flight request; queries are independent). ------------------ */
#include <pthread.h>
typedef struct {
    const uint8_t *qs; size_t nq; const uint8_t *centroids; size_t C; const uint8_t *rows; size_t n; size_t row_bytes;
    const uint32_t *list_of_row; const uint64_t *doc_ids; size_t nprobe; size_t k;
    uint64_t *ids_out; float *sims_out; int32_t *counts_out; size_t next; int err; pthread_mutex_t mu;
} many_job;

static void *many_worker(void *arg) {
    many_job *j = (many_job *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        size_t qi = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (qi >= j->nq) break;
        int c = ora_search(j->qs + qi * j->row_bytes, j->centroids, j->C, j->rows, j->n, j->row_bytes, j->list_of_row,
                           j->doc_ids, j->nprobe, j->k, j->ids_out + qi * j->k, j->sims_out + qi * j->k);
        if (c < 0) j->err = c; else j->counts_out[qi] = c;
    }
    return NULL;
}

ORA_API int ora_search_many(const uint8_t *qs, size_t nq, const uint8_t *centroids, size_t C, const uint8_t *rows,
                            size_t n, size_t row_bytes, const uint32_t *list_of_row, const uint64_t *doc_ids,
                            size_t nprobe, size_t k, uint64_t *ids_out, float *sims_out, int32_t *counts_out,
                            int threads) {
    many_job j = {qs, nq, centroids, C, rows, n, row_bytes, list_of_row, doc_ids, nprobe, k,
                  ids_out, sims_out, counts_out, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, many_worker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    return j.err;
}
