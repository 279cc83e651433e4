// api.cu -- the C ABI of libvscuda.so (include/vscuda.h): handles, arenas and launch orchestration.
// No torch types, no CPU fallback: every compute entry point needs a CUDA device.
#include <cub/device/device_radix_sort.cuh>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/vscuda.h"
#include "internal.h"

using namespace vs;

// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int g_device = -1;
static int g_sm_count = 0;
static std::mutex g_mu;
// A thread driving several devices (vs_sharded_*, sharded.cu) overrides the process default for the calls it makes.
static thread_local int tl_device = -1;
static int cur_device() { return tl_device >= 0 ? tl_device : g_device; }
namespace vs {
void internal_set_thread_device(int device) {
    tl_device = device;
    if (cur_device() >= 0) cudaSetDevice(cur_device());
}
int internal_thread_device() { return cur_device(); }
int internal_sm_count() { return g_sm_count; }
}  // namespace vs

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
namespace vs {
int internal_fail(int code, const char *msg) { return fail(code, "%s", msg); }
}  // namespace vs
#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) return fail(VS_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    } while (0)
#define VS(call)              \
    do {                      \
        int _r = (call);      \
        if (_r != VS_OK) return _r; \
    } while (0)

extern "C" const char *vs_last_error(void) { return g_err; }

extern "C" int vs_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(VS_ENODEV, "no CUDA device (%s); libvscuda has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
    if (device < 0 || device >= count) return fail(VS_EINVAL, "device %d out of range (%d devices)", device, count);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(VS_ENODEV, "device %s is sm_%d%d; libvscuda is built for sm_100a only", prop.name, prop.major, prop.minor);
    g_device = device;
    g_sm_count = prop.multiProcessorCount;
    return VS_OK;
}

extern "C" void vs_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_device >= 0) cudaDeviceSynchronize();
    g_device = -1;
}

extern "C" int vs_device_info(char *name, size_t name_cap, int *sm_count, size_t *total_mem) {
    if (g_device < 0) return fail(VS_ENODEV, "vs_init not called");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, g_device));
    if (name && name_cap) snprintf(name, name_cap, "%s", prop.name);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    return VS_OK;
}

static int need_dev() {
    if (cur_device() < 0) return fail(VS_ENODEV, "vs_init not called (or no CUDA device); libvscuda has no CPU fallback");
    cudaSetDevice(cur_device());
    return VS_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int vs_ctx_create(vs_ctx **out) {
    VS(need_dev());
    if (!out) return fail(VS_EINVAL, "out is null");
    vs_ctx *c = new vs_ctx();
    c->device = cur_device();
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    CU(cudaMalloc(&c->d_fix_counter, 8));
    CU(cudaMemset(c->d_fix_counter, 0, 8));
    CU(cudaMalloc(&c->d_tickets, (kMaxStageQueries + 8) * sizeof(unsigned int)));
    CU(cudaMemset(c->d_tickets, 0, (kMaxStageQueries + 8) * sizeof(unsigned int)));
    c->d_fused_sync = c->d_tickets + kMaxStageQueries;  // [0] grid barrier, [1] tickets, [2] uncertified-pair count
    *out = c;
    return VS_OK;
}

extern "C" int vs_ctx_create_on_stream(void *cuda_stream, vs_ctx **out) {
    VS(need_dev());
    if (!out) return fail(VS_EINVAL, "out is null");
    vs_ctx *c = new vs_ctx();
    c->device = cur_device();
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    c->owns_stream = false;
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    CU(cudaMalloc(&c->d_fix_counter, 8));
    CU(cudaMemset(c->d_fix_counter, 0, 8));
    CU(cudaMalloc(&c->d_tickets, (kMaxStageQueries + 8) * sizeof(unsigned int)));
    CU(cudaMemset(c->d_tickets, 0, (kMaxStageQueries + 8) * sizeof(unsigned int)));
    c->d_fused_sync = c->d_tickets + kMaxStageQueries;  // [0] grid barrier, [1] tickets, [2] uncertified-pair count
    *out = c;
    return VS_OK;
}

extern "C" void vs_ctx_destroy(vs_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->scratch) cudaFree(c->scratch);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->d_fix_counter) cudaFree(c->d_fix_counter);
    if (c->d_tickets) cudaFree(c->d_tickets);
    if (c->d_trace) cudaFree(c->d_trace);
    if (c->aux) cudaFree(c->aux);
    if (c->h_flags) cudaFreeHost(c->h_flags);
    for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : c->phase_ev)
        if (e) cudaEventDestroy(e);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    if (c->side_stream) {
        cudaStreamSynchronize(c->side_stream);
        cudaStreamDestroy(c->side_stream);
    }
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    if (c->join_ev) cudaEventDestroy(c->join_ev);
    if (c->owns_stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int vs_ctx_profile_enable(vs_ctx *c, int on) {
    if (!c) return fail(VS_EINVAL, "ctx is null");
    CU(cudaStreamSynchronize(c->stream));
    c->profile = on != 0;
    c->prof_used = 0;
    return VS_OK;
}

extern "C" int vs_ctx_profile_read(vs_ctx *c, double *ms_out, uint64_t *launches_out) {
    if (!c) return fail(VS_EINVAL, "ctx is null");
    CU(cudaStreamSynchronize(c->stream));
    double total = 0;
    for (size_t i = 0; i + 1 < c->prof_used; i += 2) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->prof_events[i], c->prof_events[i + 1]));
        total += ms;
    }
    if (ms_out) *ms_out = total;
    if (launches_out) *launches_out = c->prof_used / 2;
    c->prof_used = 0;
    return VS_OK;
}

constexpr size_t kTraceBlocks = 2048;
extern "C" int vs_ctx_trace_enable(vs_ctx *c, int on) {
    if (!c) return fail(VS_EINVAL, "ctx is null");
    CU(cudaStreamSynchronize(c->stream));
    if (on && !c->d_trace) {
        CU(cudaMalloc(&c->d_trace, 2 * kTraceBlocks * 16 * sizeof(unsigned long long)));
        CU(cudaMemset(c->d_trace, 0, 2 * kTraceBlocks * 16 * sizeof(unsigned long long)));
    }
    c->trace = on != 0;
    return VS_OK;
}
extern "C" int vs_ctx_trace_read(vs_ctx *c, int stage, uint64_t *out, size_t max_blocks, size_t *blocks_out) {
    if (!c || !out || !c->d_trace) return fail(VS_EINVAL, "trace not enabled");
    if (stage < 1 || stage > 2) return fail(VS_EINVAL, "stage must be 1 or 2");
    CU(cudaStreamSynchronize(c->stream));
    size_t nb = kTraceBlocks;  // (blocks that did not run leave zeros)
    if (nb > max_blocks) nb = max_blocks;
    CU(cudaMemcpy(out, c->d_trace + (size_t)(stage - 1) * kTraceBlocks * 16, nb * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (blocks_out) *blocks_out = nb;
    return VS_OK;
}

static int prof_mark(vs_ctx *c) {
    if (!c->profile) return VS_OK;
    if (c->prof_used == c->prof_events.size()) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        c->prof_events.push_back(e);
    }
    CU(cudaEventRecord(c->prof_events[c->prof_used++], c->stream));
    return VS_OK;
}

extern "C" int vs_ctx_sync(vs_ctx *c) {
    if (!c) return fail(VS_EINVAL, "ctx is null");
    CU(cudaStreamSynchronize(c->stream));
    return VS_OK;
}
extern "C" void *vs_ctx_stream(vs_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" uint64_t vs_ctx_launch_count(const vs_ctx *c) { return c ? c->launches : 0; }
extern "C" uint64_t vs_ctx_slowpath_count(const vs_ctx *c) {
    if (!c) return 0;
    unsigned long long dev = 0;  // candidates re-scored inside the search kernels (synchronizes the stream)
    cudaStreamSynchronize(c->stream);
    cudaMemcpy(&dev, c->d_fix_counter, 8, cudaMemcpyDeviceToHost);
    return c->slowpath + dev;
}
extern "C" int vs_debug_set_certify_scale(float scale) {
    if (need_dev()) return VS_ENODEV;
    if (!(scale >= 1.0f)) return fail(VS_EINVAL, "certify scale must be >= 1");
    cudaDeviceSynchronize();
    CU(vs::scan_set_certify_scale(scale));
    CU(vs::argmax_set_certify_scale(scale));
    CU(vs::gemm_set_certify_scale(scale));
    CU(vs::probe_set_certify_scale(scale));
    CU(vs::fused_set_certify_scale(scale));
    CU(vs::lm_set_certify_scale(scale));
    return VS_OK;
}

extern "C" int vs_ctx_timer_start(vs_ctx *c) {
    CU(cudaEventRecord(c->ev0, c->stream));
    return VS_OK;
}
extern "C" int vs_ctx_timer_stop(vs_ctx *c, float *ms) {
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaEventSynchronize(c->ev1));
    CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return VS_OK;
}

// Bump arena over the ctx scratch buffer. Growing synchronizes the stream first (kernels in flight may
// still be using the old buffer).
struct Arena {
    vs_ctx *c;
    size_t off = 0;
    Arena(vs_ctx *c_) : c(c_) {}
    int reserve(size_t bytes) {
        bytes += 4096;
        if (bytes <= c->scratch_cap) return VS_OK;
        CU(cudaStreamSynchronize(c->stream));
        if (c->scratch) CU(cudaFree(c->scratch));
        c->scratch = nullptr;
        c->scratch_cap = 0;
        size_t cap = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&c->scratch, cap);
        if (e != cudaSuccess) return fail(VS_ENOMEM, "cudaMalloc(%zu) for scratch: %s", cap, cudaGetErrorString(e));
        c->scratch_cap = cap;
        return VS_OK;
    }
    template <typename T>
    T *take(size_t count) {
        off = (off + 255) & ~size_t(255);
        T *p = reinterpret_cast<T *>(static_cast<char *>(c->scratch) + off);
        off += count * sizeof(T);
        return p;
    }
    static size_t pad(size_t bytes) { return (bytes + 255) & ~size_t(255); }
};

static int pinned_reserve(vs_ctx *c, size_t bytes) {
    if (bytes <= c->pinned_cap) return VS_OK;
    CU(cudaStreamSynchronize(c->stream));
    if (c->pinned) CU(cudaFreeHost(c->pinned));
    c->pinned = nullptr;
    c->pinned_cap = 0;
    cudaError_t e = cudaMallocHost(&c->pinned, bytes);
    if (e != cudaSuccess) return fail(VS_ENOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
    c->pinned_cap = bytes;
    return VS_OK;
}

#define LAUNCH(c, call)                                                                                  \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess) return fail(VS_ECUDA, "%s:%d launch %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
        (c)->launches++;                                                                                 \
    } while (0)

// ------------------------------------------------------------------------------------------------
// quantization
static int quantize_host(vs_ctx *c, const void *in, size_t n, size_t d, uint8_t *out, bool f64) {
    VS(need_dev());
    if (!c) return fail(VS_EINVAL, "ctx is null");
    if (n == 0) return VS_OK;
    if (!in || !out) return fail(VS_EINVAL, "null buffer");
    if (d > (1u << 20)) return fail(VS_ERANGE, "d=%zu too large", d);
    const size_t esz = f64 ? 8 : 4;
    const size_t rb = 8 + d;
    size_t chunk = (size_t(256) << 20) / (d * esz + rb + 1);
    if (chunk < 1) chunk = 1;
    if (chunk > n) chunk = n;
    Arena a(c);
    VS(a.reserve(Arena::pad(chunk * d * esz) + Arena::pad(chunk * rb) + 1024));
    char *d_in = a.take<char>(chunk * d * esz);
    uint8_t *d_out = a.take<uint8_t>(chunk * rb);
    for (size_t r0 = 0; r0 < n; r0 += chunk) {
        size_t m = n - r0 < chunk ? n - r0 : chunk;
        if (d) CU(cudaMemcpyAsync(d_in, (const char *)in + r0 * d * esz, m * d * esz, cudaMemcpyHostToDevice, c->stream));
        if (f64) LAUNCH(c, launch_quantize_f64((const double *)d_in, m, (int)d, d_out, c->stream));
        else LAUNCH(c, launch_quantize_f32((const float *)d_in, m, (int)d, d_out, c->stream));
        CU(cudaMemcpyAsync(out + r0 * rb, d_out, m * rb, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return VS_OK;
}

extern "C" int vs_quantize_f32(vs_ctx *c, const float *in, size_t n, size_t d, uint8_t *out) {
    return quantize_host(c, in, n, d, out, false);
}
extern "C" int vs_quantize_f64(vs_ctx *c, const double *in, size_t n, size_t d, uint8_t *out) {
    return quantize_host(c, in, n, d, out, true);
}
extern "C" int vs_quantize_f32_dev(vs_ctx *c, const float *d_in, size_t n, size_t d, uint8_t *d_out) {
    VS(need_dev());
    if (n == 0) return VS_OK;
    LAUNCH(c, launch_quantize_f32(d_in, n, (int)d, d_out, c->stream));
    return VS_OK;
}
extern "C" int vs_quantize_f64_dev(vs_ctx *c, const double *d_in, size_t n, size_t d, uint8_t *d_out) {
    VS(need_dev());
    if (n == 0) return VS_OK;
    LAUNCH(c, launch_quantize_f64(d_in, n, (int)d, d_out, c->stream));
    return VS_OK;
}

static int dequantize_host(vs_ctx *c, const uint8_t *rows, size_t n, size_t row_bytes, void *out, bool f64) {
    VS(need_dev());
    if (!c) return fail(VS_EINVAL, "ctx is null");
    if (n == 0) return VS_OK;
    if (row_bytes < 8) return fail(VS_EINVAL, "row_bytes=%zu < 8", row_bytes);
    const size_t d = row_bytes - 8;
    if (d == 0) return VS_OK;
    const size_t esz = f64 ? 8 : 4;
    size_t chunk = (size_t(256) << 20) / (d * esz + row_bytes);
    if (chunk < 1) chunk = 1;
    if (chunk > n) chunk = n;
    Arena a(c);
    VS(a.reserve(Arena::pad(chunk * d * esz) + Arena::pad(chunk * row_bytes) + 1024));
    uint8_t *d_rows = a.take<uint8_t>(chunk * row_bytes);
    char *d_out = a.take<char>(chunk * d * esz);
    for (size_t r0 = 0; r0 < n; r0 += chunk) {
        size_t m = n - r0 < chunk ? n - r0 : chunk;
        CU(cudaMemcpyAsync(d_rows, rows + r0 * row_bytes, m * row_bytes, cudaMemcpyHostToDevice, c->stream));
        if (f64) LAUNCH(c, launch_dequantize_f64(d_rows, m, (int)row_bytes, (double *)d_out, c->stream));
        else LAUNCH(c, launch_dequantize_f32(d_rows, m, (int)row_bytes, (float *)d_out, c->stream));
        CU(cudaMemcpyAsync((char *)out + r0 * d * esz, d_out, m * d * esz, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return VS_OK;
}
extern "C" int vs_dequantize_f32(vs_ctx *c, const uint8_t *rows, size_t n, size_t row_bytes, float *out) {
    return dequantize_host(c, rows, n, row_bytes, out, false);
}
extern "C" int vs_dequantize_f64(vs_ctx *c, const uint8_t *rows, size_t n, size_t row_bytes, double *out) {
    return dequantize_host(c, rows, n, row_bytes, out, true);
}

// ------------------------------------------------------------------------------------------------
// Matrix storage comes from the device's stream-ordered memory pool.  cudaMalloc / cudaFree cost milliseconds for the
// sizes of a store (mapping and unmapping the pages) and cudaFree stalls every stream: the divide-and-conquer build, which
// makes and drops a few thousand matrices, spent a third of its time there.  The pool keeps freed pages mapped (up to
// VS_POOL_KEEP_GB, default 24) and hands them out again; vs_release_cached_memory returns them to the driver.
static cudaStream_t g_pool_stream[64] = {};
static std::mutex g_pool_mu;
static cudaError_t pool_stream_for(int dev, cudaStream_t *out) {
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!g_pool_stream[dev]) {
        cudaMemPool_t pool;
        cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, dev);
        if (e != cudaSuccess) return e;
        const char *env = getenv("VS_POOL_KEEP_GB");
        uint64_t keep = (uint64_t)(env ? atoll(env) : 24) << 30;
        if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&g_pool_stream[dev], cudaStreamNonBlocking)) != cudaSuccess) return e;
    }
    *out = g_pool_stream[dev];
    return cudaSuccess;
}
// usable on every stream when it returns
static cudaError_t pool_alloc(void **p, size_t bytes) {
    const int dev = cur_device();
    cudaStream_t st;
    cudaError_t e = pool_stream_for(dev, &st);
    if (e != cudaSuccess) return e;
    e = cudaMallocAsync(p, bytes, st);
    if (e == cudaErrorMemoryAllocation) {  // give the cached pages back and try once more
        cudaGetLastError();
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            cudaDeviceSynchronize();
            cudaMemPoolTrimTo(pool, 0);
        }
        e = cudaMallocAsync(p, bytes, st);
    }
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}
// like cudaFree, waits for whatever may still be reading the memory (the device must be current)
static void pool_free_all(int dev, void *a, void *b, void *c3) {
    cudaStream_t st;
    if (pool_stream_for(dev, &st) != cudaSuccess) return;
    cudaDeviceSynchronize();
    if (a) cudaFreeAsync(a, st);
    if (b) cudaFreeAsync(b, st);
    if (c3) cudaFreeAsync(c3, st);
}
extern "C" int vs_release_cached_memory(void) {
    VS(need_dev());
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, cur_device()));
    CU(cudaDeviceSynchronize());
    CU(cudaMemPoolTrimTo(pool, 0));
    return VS_OK;
}

// matrices
constexpr size_t kPoolMaxBytes = size_t(1) << 30;
static void matrix_free_storage(vs_matrix *m) {
    if (m->pooled) {
        pool_free_all(m->device, m->codes, m->hdr, m->sums);
    } else {
        if (m->codes) cudaFree(m->codes);
        if (m->hdr) cudaFree(m->hdr);
        if (m->sums) cudaFree(m->sums);
    }
}
static int matrix_alloc(size_t n, size_t d, vs_matrix **out) {
    if (d > 4096) return fail(VS_ERANGE, "d=%zu: kernels support d <= 4096", d);
    if (n > 0x7FFFFFFFull) return fail(VS_ERANGE, "n=%zu: a device matrix holds < 2^31 rows", n);
    vs_matrix *m = new vs_matrix();
    m->device = cur_device();
    m->n = n;
    m->d = (int)d;
    m->d_pad = (int)((d + 15) & ~size_t(15));
    // store-sized matrices are rare and the pool maps fresh memory slowly (~0.15 s per GB measured): plain cudaMalloc there
    m->pooled = n * (size_t)m->d_pad < kPoolMaxBytes;
    auto alloc = [&](void **p, size_t bytes) { return m->pooled ? pool_alloc(p, bytes) : cudaMalloc(p, bytes); };
    cudaError_t e = alloc(reinterpret_cast<void **>(&m->codes), n * (size_t)m->d_pad + 256);
    if (e == cudaSuccess) e = alloc(reinterpret_cast<void **>(&m->hdr), n * sizeof(float2) + 256);
    if (e == cudaSuccess) e = alloc(reinterpret_cast<void **>(&m->sums), n * sizeof(uint2) + 256);
    if (e != cudaSuccess) {
        cudaGetLastError();
        matrix_free_storage(m);
        delete m;
        return fail(VS_ENOMEM, "cudaMalloc for %zu x %zu matrix: %s", n, d, cudaGetErrorString(e));
    }
    *out = m;
    return VS_OK;
}

extern "C" void vs_matrix_retain(vs_matrix *m) {
    if (m) m->refs.fetch_add(1);
}
extern "C" void vs_matrix_release(vs_matrix *m) {
    if (!m) return;
    if (m->refs.fetch_sub(1) == 1) {
        cudaSetDevice(m->device);
        if (m->owns) matrix_free_storage(m);
        delete m;
    }
}
extern "C" size_t vs_matrix_rows(const vs_matrix *m) { return m ? m->n : 0; }
extern "C" size_t vs_matrix_cols(const vs_matrix *m) { return m ? (size_t)m->d : 0; }

static int check_rows(size_t n, size_t row_bytes) {
    if (n == 0) return fail(VS_EEMPTY, "matrix rows are empty");          // compute.go:25-27
    if (row_bytes <= 8) return fail(VS_EEMPTY, "matrix columns are empty");  // compute.go:29-31
    return VS_OK;
}

extern "C" int vs_matrix_create_dev(vs_ctx *c, const uint8_t *d_rows, size_t n, size_t row_bytes, vs_matrix **out) {
    VS(need_dev());
    if (!c || !out) return fail(VS_EINVAL, "null argument");
    VS(check_rows(n, row_bytes));
    vs_matrix *m = nullptr;
    VS(matrix_alloc(n, row_bytes - 8, &m));
    cudaError_t e = launch_ingest(d_rows, n, (int)row_bytes, m->codes, m->d_pad, m->hdr, m->sums, c->stream);
    if (e != cudaSuccess) {
        vs_matrix_release(m);
        return fail(VS_ECUDA, "ingest: %s", cudaGetErrorString(e));
    }
    c->launches++;
    *out = m;
    return VS_OK;
}

extern "C" int vs_matrix_create(vs_ctx *c, const uint8_t *rows, size_t n, size_t row_bytes, vs_matrix **out) {
    VS(need_dev());
    if (!c || !out) return fail(VS_EINVAL, "null argument");
    VS(check_rows(n, row_bytes));
    if (!rows) return fail(VS_EINVAL, "rows is null");
    vs_matrix *m = nullptr;
    VS(matrix_alloc(n, row_bytes - 8, &m));
    size_t chunk = (size_t(256) << 20) / row_bytes;
    if (chunk > n) chunk = n;
    Arena a(c);
    int rc = a.reserve(Arena::pad(chunk * row_bytes) + 1024);
    if (rc != VS_OK) {
        vs_matrix_release(m);
        return rc;
    }
    uint8_t *stage = a.take<uint8_t>(chunk * row_bytes);
    for (size_t r0 = 0; r0 < n; r0 += chunk) {
        size_t cnt = n - r0 < chunk ? n - r0 : chunk;
        cudaError_t e = cudaMemcpyAsync(stage, rows + r0 * row_bytes, cnt * row_bytes, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess)
            e = launch_ingest(stage, cnt, (int)row_bytes, m->codes + r0 * (size_t)m->d_pad, m->d_pad, m->hdr + r0,
                              m->sums + r0, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            vs_matrix_release(m);
            return fail(VS_ECUDA, "matrix upload: %s", cudaGetErrorString(e));
        }
        c->launches++;
    }
    *out = m;
    return VS_OK;
}

extern "C" int vs_matrix_from_f32_dev(vs_ctx *c, const float *d_in, size_t n, size_t d, vs_matrix **out) {
    VS(need_dev());
    if (!c || !out) return fail(VS_EINVAL, "null argument");
    VS(check_rows(n, d + 8));
    vs_matrix *m = nullptr;
    VS(matrix_alloc(n, d, &m));
    cudaError_t e = launch_quantize_f32_soa(d_in, n, (int)d, m->codes, m->d_pad, m->hdr, m->sums, c->stream);
    if (e != cudaSuccess) {
        vs_matrix_release(m);
        return fail(VS_ECUDA, "quantize: %s", cudaGetErrorString(e));
    }
    c->launches++;
    *out = m;
    return VS_OK;
}

extern "C" int vs_matrix_create_empty(vs_ctx *c, size_t n, size_t d, vs_matrix **out) {
    VS(need_dev());
    if (!c || !out) return fail(VS_EINVAL, "null argument");
    VS(check_rows(n, d + 8));
    return matrix_alloc(n, d, out);
}

extern "C" int vs_matrix_fill_f32_dev(vs_ctx *c, vs_matrix *m, size_t first, const float *d_in, size_t count) {
    VS(need_dev());
    if (!c || !m || !d_in) return fail(VS_EINVAL, "null argument");
    if (first + count > m->n) return fail(VS_EINVAL, "row range out of bounds");
    if (count == 0) return VS_OK;
    LAUNCH(c, launch_quantize_f32_soa(d_in, count, m->d, m->codes + first * (size_t)m->d_pad, m->d_pad, m->hdr + first,
                                      m->sums + first, c->stream));
    return VS_OK;
}

extern "C" int vs_matrix_load_rows(vs_ctx *c, vs_matrix *m, size_t first, const uint8_t *rows, size_t count) {
    VS(need_dev());
    if (!c || !m || !rows) return fail(VS_EINVAL, "null argument");
    if (first + count > m->n) return fail(VS_EINVAL, "row range out of bounds");
    if (count == 0) return VS_OK;
    const size_t rb = 8 + (size_t)m->d;
    Arena a(c);
    VS(a.reserve(Arena::pad(count * rb) + 1024));
    uint8_t *stage = a.take<uint8_t>(count * rb);
    CU(cudaMemcpyAsync(stage, rows, count * rb, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, launch_ingest(stage, count, (int)rb, m->codes + first * (size_t)m->d_pad, m->d_pad, m->hdr + first,
                            m->sums + first, c->stream));
    return VS_OK;
}

extern "C" int vs_matrix_read_rows(vs_ctx *c, const vs_matrix *m, size_t first, size_t count, uint8_t *out) {
    VS(need_dev());
    if (!c || !m || !out) return fail(VS_EINVAL, "null argument");
    if (first + count > m->n) return fail(VS_EINVAL, "row range out of bounds");
    if (count == 0) return VS_OK;
    const size_t rb = 8 + (size_t)m->d;
    Arena a(c);
    VS(a.reserve(Arena::pad(count * rb) + 1024));
    uint8_t *buf = a.take<uint8_t>(count * rb);
    LAUNCH(c, launch_export(m->view(), first, count, buf, c->stream));
    CU(cudaMemcpyAsync(out, buf, count * rb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VS_OK;
}

// ---- the D&C row spool (dnc/dataset.go:19-56,122-146): a flat file of 8+D-byte rows, no header -----------------
// Loader: the file is read with pread into two pinned staging halves; while one half is on its way to the device
// (H2D + ingest kernel on the ctx stream) the next chunk is read into the other.  Replaces the per-row
// `make([]uint8) + io.ReadFull` of dataset.ReadRow and the [][]uint8 -> NewMatrix re-pack of chunkData (k_means.go:214-221).
constexpr size_t kSpoolChunkBytes = size_t(64) << 20;

extern "C" int vs_matrix_load_spool(vs_ctx *c, const char *path, size_t row_bytes, size_t first_row, size_t count,
                                    vs_matrix **out) {
    VS(need_dev());
    if (!c || !path || !out) return fail(VS_EINVAL, "null argument");
    if (row_bytes <= 8) return fail(VS_EEMPTY, "matrix columns are empty");  // compute.go:29-31
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(VS_EINVAL, "open %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) {
        close(fd);
        return fail(VS_EINVAL, "stat %s: %s", path, strerror(errno));
    }
    const size_t fsize = (size_t)st.st_size;
    if (fsize % row_bytes != 0) {  // dataset.ReadRow: io.ReadFull fails on a short row (dataset.go:131-136)
        close(fd);
        return fail(VS_EINVAL, "%s: %zu bytes is not a whole number of %zu-byte rows", path, fsize, row_bytes);
    }
    const size_t total = fsize / row_bytes;
    if (first_row > total) first_row = total;
    const size_t n = (count == 0 || first_row + count > total) ? total - first_row : count;
    if (n == 0) {
        close(fd);
        return fail(VS_EEMPTY, "matrix rows are empty");  // compute.go:25-27
    }
    vs_matrix *m = nullptr;
    int rc = matrix_alloc(n, row_bytes - 8, &m);
    size_t chunk = kSpoolChunkBytes / row_bytes;
    if (chunk < 1) chunk = 1;
    if (chunk > n) chunk = n;
    Arena a(c);
    if (rc == VS_OK) rc = a.reserve(2 * Arena::pad(chunk * row_bytes) + 1024);
    if (rc == VS_OK) rc = pinned_reserve(c, 2 * chunk * row_bytes);
    for (int i = 4; i < 6 && rc == VS_OK; i++)
        if (!c->phase_ev[i] && cudaEventCreateWithFlags(&c->phase_ev[i], cudaEventDisableTiming) != cudaSuccess)
            rc = fail(VS_ECUDA, "event create");
    if (rc != VS_OK) {
        close(fd);
        if (m) vs_matrix_release(m);
        return rc;
    }
    uint8_t *d_stage[2] = {a.take<uint8_t>(chunk * row_bytes), a.take<uint8_t>(chunk * row_bytes)};
    uint8_t *h_stage[2] = {static_cast<uint8_t *>(c->pinned), static_cast<uint8_t *>(c->pinned) + chunk * row_bytes};
    bool used[2] = {false, false};
    int b = 0;
    for (size_t r0 = 0; r0 < n && rc == VS_OK; r0 += chunk, b ^= 1) {
        const size_t cnt = n - r0 < chunk ? n - r0 : chunk;
        if (used[b] && cudaEventSynchronize(c->phase_ev[4 + b]) != cudaSuccess) rc = fail(VS_ECUDA, "spool: event sync");
        size_t got = 0;
        const size_t want = cnt * row_bytes;
        const off_t base = (off_t)((first_row + r0) * row_bytes);
        while (got < want && rc == VS_OK) {
            const ssize_t r = pread(fd, h_stage[b] + got, want - got, base + (off_t)got);
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) rc = fail(VS_EINVAL, "read %s: %s", path, r < 0 ? strerror(errno) : "unexpected end of file");
            else got += (size_t)r;
        }
        if (rc != VS_OK) break;
        cudaError_t e = cudaMemcpyAsync(d_stage[b], h_stage[b], want, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess)
            e = launch_ingest(d_stage[b], cnt, (int)row_bytes, m->codes + r0 * (size_t)m->d_pad, m->d_pad, m->hdr + r0, m->sums + r0,
                              c->stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->phase_ev[4 + b], c->stream);
        if (e != cudaSuccess) rc = fail(VS_ECUDA, "spool upload: %s", cudaGetErrorString(e));
        used[b] = true;
        c->launches++;
    }
    close(fd);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && rc == VS_OK) rc = fail(VS_ECUDA, "spool upload: stream sync");
    if (rc != VS_OK) {
        vs_matrix_release(m);
        return rc;
    }
    *out = m;
    return VS_OK;
}

// Writer: rows [first, first+count) of a device matrix as 8+D-byte rows, appended to (or replacing) a spool file --
// createDataset.WriteRow (dataset.go:53-56) for a whole child cluster at once.
extern "C" int vs_matrix_save_spool(vs_ctx *c, const vs_matrix *m, size_t first, size_t count, const char *path, int append) {
    VS(need_dev());
    if (!c || !m || !path) return fail(VS_EINVAL, "null argument");
    if (first + count > m->n) return fail(VS_EINVAL, "row range out of bounds");
    const size_t rb = 8 + (size_t)m->d;
    const int fd = open(path, O_WRONLY | O_CREAT | (append ? O_APPEND : O_TRUNC), 0644);
    if (fd < 0) return fail(VS_EINVAL, "open %s: %s", path, strerror(errno));
    size_t chunk = kSpoolChunkBytes / rb;
    if (chunk < 1) chunk = 1;
    if (chunk > count) chunk = count ? count : 1;
    Arena a(c);
    int rc = a.reserve(Arena::pad(chunk * rb) + 1024);
    if (rc == VS_OK) rc = pinned_reserve(c, chunk * rb);
    uint8_t *d_buf = rc == VS_OK ? a.take<uint8_t>(chunk * rb) : nullptr;
    for (size_t r0 = 0; r0 < count && rc == VS_OK; r0 += chunk) {
        const size_t cnt = count - r0 < chunk ? count - r0 : chunk;
        cudaError_t e = launch_export(m->view(), first + r0, cnt, d_buf, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->pinned, d_buf, cnt * rb, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            rc = fail(VS_ECUDA, "spool download: %s", cudaGetErrorString(e));
            break;
        }
        c->launches++;
        size_t put = 0;
        while (put < cnt * rb && rc == VS_OK) {
            const ssize_t w = write(fd, static_cast<uint8_t *>(c->pinned) + put, cnt * rb - put);
            if (w < 0 && errno == EINTR) continue;
            if (w <= 0) rc = fail(VS_EINVAL, "write %s: %s", path, strerror(errno));
            else put += (size_t)w;
        }
    }
    if (close(fd) != 0 && rc == VS_OK) rc = fail(VS_EINVAL, "close %s: %s", path, strerror(errno));
    return rc;
}

// A temporary (arena-resident) matrix built from host rows: queries, centroids.
static int temp_matrix(vs_ctx *c, Arena &a, const uint8_t *rows, size_t n, size_t row_bytes, MatView *out) {
    const size_t d = row_bytes - 8, d_pad = (d + 15) & ~size_t(15);
    uint8_t *stage = a.take<uint8_t>(n * row_bytes);
    uint8_t *codes = a.take<uint8_t>(n * d_pad);
    float2 *hdr = a.take<float2>(n);
    uint2 *sums = a.take<uint2>(n);
    CU(cudaMemcpyAsync(stage, rows, n * row_bytes, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, launch_ingest(stage, n, (int)row_bytes, codes, (int)d_pad, hdr, sums, c->stream));
    *out = MatView{codes, hdr, sums, n, (int)d, (int)d_pad};
    return VS_OK;
}
static size_t temp_matrix_bytes(size_t n, size_t row_bytes) {
    const size_t d = row_bytes - 8, d_pad = (d + 15) & ~size_t(15);
    return Arena::pad(n * row_bytes) + Arena::pad(n * d_pad) + 2 * Arena::pad(n * 8) + 1024;
}

// ------------------------------------------------------------------------------------------------
// cosine 1xN / dots
static int cosine_common(vs_ctx *c, const uint8_t *q, size_t q_bytes, const vs_matrix *m, float *sims_out,
                         uint32_t *dots_out) {
    VS(need_dev());
    if (!c || !m) return fail(VS_EINVAL, "null argument");
    if (q_bytes <= 8) return fail(VS_EEMPTY, "vector columns are empty");  // compute.go:12-14
    if (!q) return fail(VS_EINVAL, "q is null");
    if ((size_t)m->d != q_bytes - 8)  // cosine.go:19-21
        return fail(VS_EDIM, "vector/matrix column size does not match: %zu != %d", q_bytes - 8, m->d);
    const size_t n = m->n;
    Arena a(c);
    VS(a.reserve(temp_matrix_bytes(1, q_bytes) + Arena::pad(n * 4) * 2 + Arena::pad((size_t)m->d * 8) + 4096));
    MatView qv;
    VS(temp_matrix(c, a, q, 1, q_bytes, &qv));
    float *d_sims = sims_out ? a.take<float>(n) : nullptr;
    uint32_t *d_dots = dots_out ? a.take<uint32_t>(n) : nullptr;
    uint32_t *d_work = a.take<uint32_t>(n);
    unsigned int *d_count = a.take<unsigned int>(1);
    double *d_qn = a.take<double>(m->d);
    CU(cudaMemsetAsync(d_count, 0, sizeof(unsigned int), c->stream));
    LAUNCH(c, launch_cosine_1xN(m->view(), qv, d_sims, d_dots, d_work, d_count, g_sm_count, c->stream));
    if (sims_out) {
        unsigned int cnt = 0;
        CU(cudaMemcpyAsync(&cnt, d_count, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (cnt > 0) {
            LAUNCH(c, launch_query_normalize(qv, d_qn, c->stream));
            LAUNCH(c, launch_cosine_fix(m->view(), d_qn, d_sims, d_work, d_count, g_sm_count, c->stream));
            c->slowpath += cnt;
        }
        CU(cudaMemcpyAsync(sims_out, d_sims, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    if (dots_out) CU(cudaMemcpyAsync(dots_out, d_dots, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VS_OK;
}

extern "C" int vs_cosine_1xN(vs_ctx *c, const uint8_t *q, size_t q_bytes, const vs_matrix *m, float *sims_out) {
    if (!sims_out) return fail(VS_EINVAL, "sims_out is null");
    return cosine_common(c, q, q_bytes, m, sims_out, nullptr);
}
extern "C" int vs_dot_1xN(vs_ctx *c, const uint8_t *q, size_t q_bytes, const vs_matrix *m, uint32_t *dots_out) {
    if (!dots_out) return fail(VS_EINVAL, "dots_out is null");
    return cosine_common(c, q, q_bytes, m, nullptr, dots_out);
}

// ------------------------------------------------------------------------------------------------
// argmax MxN
static size_t g_argmax_gemm_min_centroids = 256;  // from here on the tensor cores beat the dp4a scan (vs_debug_set_argmax_gemm_min)
constexpr size_t kArgmaxGemmMinRows = 1024;
constexpr size_t kKMeansScanCentroids = 64;        // up to here the centroid update walks the assignment instead of sorting it
constexpr size_t kArgmaxInlineFixCentroids = 64;  // up to here the literal-arithmetic kernel is enqueued without asking the host
static int argmax_gemm_dev(vs_ctx *c, const MatView &cent, const MatView &data, int32_t *d_idx, float *d_sims, bool *done);

static int argmax_dev(vs_ctx *c, Arena &a, const MatView &cent, const MatView &data, int32_t *d_idx, float *d_sims) {
    const size_t n = data.n, M = cent.n;
    if (M >= g_argmax_gemm_min_centroids && n >= kArgmaxGemmMinRows && gemm_store_supported(cent) && gemm_store_supported(data)) {
        bool done = false;
        VS(argmax_gemm_dev(c, cent, data, d_idx, d_sims, &done));
        if (done) return VS_OK;
    }
    uint32_t *d_canon = a.take<uint32_t>(M);
    uint32_t *d_work = a.take<uint32_t>(n);
    unsigned int *d_count = a.take<unsigned int>(1);
    CU(cudaMemsetAsync(d_count, 0, sizeof(unsigned int), c->stream));
    LAUNCH(c, launch_canonical_rows(cent, d_canon, c->stream));
    LAUNCH(c, launch_argmax(cent, data, d_canon, d_idx, d_sims, d_work, d_count, g_sm_count, c->stream));
    if (M <= kArgmaxInlineFixCentroids) {
        // few centroids (the reference's own shapes): the literal-arithmetic kernel is enqueued unconditionally -- it
        // reads the worklist length on the device and does nothing when it is zero -- so the call never waits for the
        // host (k-means iterations stay back to back on the stream)
        double *d_cn = a.take<double>(M * (size_t)cent.d);
        LAUNCH(c, launch_query_normalize(cent, d_cn, c->stream, d_count));
        LAUNCH(c, launch_argmax_fix(cent, data, d_cn, d_idx, d_sims, d_work, d_count, g_sm_count, c->stream));
        return VS_OK;
    }
    unsigned int cnt = 0;
    CU(cudaMemcpyAsync(&cnt, d_count, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (cnt > 0) {
        double *d_cn = nullptr;  // normalized centroids: too big for the arena at large m, allocate
        CU(cudaMalloc(&d_cn, M * (size_t)cent.d * sizeof(double)));
        cudaError_t e = launch_query_normalize(cent, d_cn, c->stream);
        if (e == cudaSuccess) e = launch_argmax_fix(cent, data, d_cn, d_idx, d_sims, d_work, d_count, g_sm_count, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        cudaFree(d_cn);
        if (e != cudaSuccess) return fail(VS_ECUDA, "argmax fix: %s", cudaGetErrorString(e));
        c->launches += 2;
        c->slowpath += cnt;
    }
    return VS_OK;
}
static size_t argmax_bytes(size_t M, size_t n, size_t d) {
    return Arena::pad(M * 4) + Arena::pad(n * 4) + (M <= kArgmaxInlineFixCentroids ? Arena::pad(M * d * 8) : 0) + 1024;
}

static int argmax_check(const vs_matrix *cent, const vs_matrix *data) {
    if (!cent || !data) return fail(VS_EINVAL, "null matrix");
    if (cent->d != data->d)  // cosine.go:77-79
        return fail(VS_EDIM, "matrix/matrix column size does not match: %d != %d", cent->d, data->d);
    return VS_OK;
}

extern "C" int vs_argmax_MxN_dev(vs_ctx *c, const vs_matrix *cent, const vs_matrix *data, int32_t *d_idx_out) {
    VS(need_dev());
    if (!c || !d_idx_out) return fail(VS_EINVAL, "null argument");
    VS(argmax_check(cent, data));
    Arena a(c);
    VS(a.reserve(argmax_bytes(cent->n, data->n, (size_t)cent->d)));
    return argmax_dev(c, a, cent->view(), data->view(), d_idx_out, nullptr);
}

extern "C" int vs_argmax_MxN(vs_ctx *c, const vs_matrix *cent, const vs_matrix *data, float *sims_out, int64_t *idx_out) {
    VS(need_dev());
    if (!c || !idx_out) return fail(VS_EINVAL, "null argument");
    VS(argmax_check(cent, data));
    const size_t n = data->n;
    Arena a(c);
    VS(a.reserve(argmax_bytes(cent->n, n, (size_t)cent->d) + 2 * Arena::pad(n * 4) + 1024));
    int32_t *d_idx = a.take<int32_t>(n);
    float *d_sims = sims_out ? a.take<float>(n) : nullptr;
    VS(argmax_dev(c, a, cent->view(), data->view(), d_idx, d_sims));
    std::vector<int32_t> tmp(n);
    CU(cudaMemcpyAsync(tmp.data(), d_idx, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (sims_out) CU(cudaMemcpyAsync(sims_out, d_sims, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; i++) idx_out[i] = tmp[i];
    return VS_OK;
}

// ------------------------------------------------------------------------------------------------
// index
__global__ void iota_kernel(uint32_t *p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}
// off[c] = first position in sorted keys with key >= c, for c in [0, C]
__global__ void lower_bound_kernel(const uint32_t *keys, size_t n, uint64_t *off64, uint32_t *off32, size_t C) {
    size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (c > C) return;
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t mid = (lo + hi) >> 1;
        if (keys[mid] < (uint32_t)c) lo = mid + 1;
        else hi = mid;
    }
    if (off64) off64[c] = lo;
    if (off32) off32[c] = (uint32_t)lo;
}

// Stable sort of rows by key: d_order[i] = source row of sorted position i; d_sorted_keys optional.
// ws/ws_bytes: optional caller workspace of at least sort_rows_ws_bytes(n) (else allocated here).
static size_t sort_rows_ws_bytes(size_t n) {
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int64_t)n, 0, 32);
    return Arena::pad(n * 4 + 4) + Arena::pad(temp_bytes + 16);
}
static int sort_rows_by_key(vs_ctx *c, const uint32_t *d_keys, size_t n, int key_bits, uint32_t *d_order,
                            uint32_t *d_keys_sorted, void *ws = nullptr, size_t ws_bytes = 0) {
    const size_t need = sort_rows_ws_bytes(n);
    char *own = nullptr;
    if (!ws || ws_bytes < need) {
        CU(cudaMalloc(&own, need));
        ws = own;
    }
    uint32_t *d_iota = static_cast<uint32_t *>(ws);
    void *d_temp = static_cast<char *>(ws) + Arena::pad(n * 4 + 4);
    size_t temp_bytes = need - Arena::pad(n * 4 + 4);
    iota_kernel<<<g_sm_count * 4, 256, 0, c->stream>>>(d_iota, n);
    c->launches++;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_keys, d_keys_sorted, d_iota, d_order, (int64_t)n, 0, key_bits,
                                                    c->stream);
    if (own) {
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        cudaFree(own);
    }
    if (e != cudaSuccess) return fail(VS_ECUDA, "radix sort: %s", cudaGetErrorString(e));
    c->launches += 4;
    return VS_OK;
}

static int bits_for(size_t C) {
    int b = 1;
    while ((size_t(1) << b) < C + 1 && b < 32) b++;
    return b;
}

extern "C" void vs_index_release(vs_index *ix) {
    if (!ix) return;
    if (ix->data) cudaSetDevice(ix->data->device);
    if (ix->doc_ids) cudaFree(ix->doc_ids);
    if (ix->list_off) cudaFree(ix->list_off);
    if (ix->list_len && ix->list_len != ix->fill_cursor) cudaFree(ix->list_len);
    if (ix->fill_cursor) cudaFree(ix->fill_cursor);
    vs_matrix_release(ix->data);
    vs_matrix_release(ix->centroids);
    delete ix;
}
// rows present (an index with room to grow: rows placed so far; its capacity is vs_index_capacity)
extern "C" size_t vs_index_rows(const vs_index *ix) { return !ix ? 0 : ix->fill_cursor ? ix->filled : ix->n; }
extern "C" size_t vs_index_capacity(const vs_index *ix) { return ix ? ix->n : 0; }
// holes between the lists: whole-store scans (flat, GEMM batch) would score rows that are not there
static bool index_has_holes(const vs_index *ix) { return ix->fill_cursor && ix->filled != ix->n; }

__global__ void list_lens_kernel(const uint64_t *off, size_t C, uint64_t *len) {
    const size_t l = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (l < C) len[l] = off[l + 1] - off[l];
}
// list_len of an index built in one piece (lists back to back): the differences of list_off
static cudaError_t index_make_lens(vs_ctx *c, vs_index *ix) {
    if (ix->C == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&ix->list_len, ix->C * 8);
    if (e != cudaSuccess) return e;
    list_lens_kernel<<<(unsigned)((ix->C + 255) / 256), 256, 0, c->stream>>>(ix->list_off, ix->C, ix->list_len);
    return cudaGetLastError();
}
extern "C" size_t vs_index_lists(const vs_index *ix) { return ix ? ix->C : 0; }
extern "C" size_t vs_index_cols(const vs_index *ix) { return ix && ix->centroids ? (size_t)ix->centroids->d : 0; }

extern "C" int vs_index_list_offsets(vs_ctx *c, const vs_index *ix, uint64_t *out) {
    VS(need_dev());
    if (!c || !ix || !out) return fail(VS_EINVAL, "null argument");
    CU(cudaMemcpyAsync(out, ix->list_off, (ix->C + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VS_OK;
}

extern "C" int vs_index_list_lengths(vs_ctx *c, const vs_index *ix, uint64_t *out) {
    VS(need_dev());
    if (!c || !ix || !out) return fail(VS_EINVAL, "null argument");
    if (ix->C == 0) return VS_OK;
    CU(cudaMemcpyAsync(out, ix->list_len, ix->C * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VS_OK;
}

extern "C" int vs_index_read_rows(vs_ctx *c, const vs_index *ix, size_t first, size_t count, uint8_t *rows_out,
                                  uint64_t *ids_out) {
    VS(need_dev());
    if (!c || !ix) return fail(VS_EINVAL, "null argument");
    if (first + count > ix->n) return fail(VS_EINVAL, "row range out of bounds");
    if (rows_out) VS(vs_matrix_read_rows(c, ix->data, first, count, rows_out));
    if (ids_out) {
        if (ix->doc_ids) {
            CU(cudaMemcpyAsync(ids_out, ix->doc_ids + first, count * 8, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        } else {
            for (size_t i = 0; i < count; i++) ids_out[i] = ix->id_base + first + i;
        }
    }
    return VS_OK;
}

extern "C" int vs_index_build_dev(vs_ctx *c, const vs_matrix *data, const int32_t *d_list_of_row, const uint64_t *d_doc_ids,
                                  uint64_t id_base, const vs_matrix *centroids, vs_index **out) {
    VS(need_dev());
    if (!c || !data || !centroids || !d_list_of_row || !out) return fail(VS_EINVAL, "null argument");
    if (data->d != centroids->d) return fail(VS_EDIM, "data/centroid column size does not match: %d != %d", data->d, centroids->d);
    const size_t n = data->n, C = centroids->n;
    uint32_t *d_order = nullptr, *d_keys_sorted = nullptr;
    CU(cudaMalloc(&d_order, n * 4 + 4));
    if (cudaMalloc(&d_keys_sorted, n * 4 + 4) != cudaSuccess) {
        cudaFree(d_order);
        return fail(VS_ENOMEM, "cudaMalloc for the sorted list keys of %zu rows", n);
    }
    int rc = sort_rows_by_key(c, reinterpret_cast<const uint32_t *>(d_list_of_row), n, bits_for(C), d_order, d_keys_sorted);
    vs_index *ix = nullptr;
    vs_matrix *g = nullptr;
    if (rc == VS_OK) rc = matrix_alloc(n, data->d, &g);
    if (rc == VS_OK) {
        ix = new vs_index();
        ix->n = n;
        ix->C = C;
        ix->data = g;
        ix->centroids = const_cast<vs_matrix *>(centroids);
        vs_matrix_retain(ix->centroids);
        ix->id_base = id_base;
        ix->implicit_ids = d_doc_ids == nullptr;
        cudaError_t e = cudaMalloc(&ix->doc_ids, n * 8 + 8);
        if (e == cudaSuccess) e = cudaMalloc(&ix->list_off, (C + 1) * 8);
        if (e == cudaSuccess)
            e = launch_gather_rows(data->view(), d_order, n, g->codes, g->hdr, g->sums, d_doc_ids, id_base, ix->doc_ids, c->stream);
        if (e == cudaSuccess) {
            lower_bound_kernel<<<(unsigned)((C + 1 + 255) / 256), 256, 0, c->stream>>>(d_keys_sorted, n, ix->list_off, nullptr, C);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = index_make_lens(c, ix);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(VS_ECUDA, "index build: %s", cudaGetErrorString(e));
        c->launches += 2;
    }
    cudaFree(d_order);
    cudaFree(d_keys_sorted);
    if (rc != VS_OK) {
        if (ix) vs_index_release(ix);
        else if (g) vs_matrix_release(g);
        return rc;
    }
    *out = ix;
    return VS_OK;
}

extern "C" int vs_index_build_assigned(vs_ctx *c, const uint8_t *rows, size_t n, size_t row_bytes, const uint64_t *doc_ids,
                                       const uint32_t *list_of_row, const uint8_t *centroids, size_t C, vs_index **out) {
    VS(need_dev());
    if (!c || !rows || !list_of_row || !centroids || !out) return fail(VS_EINVAL, "null argument");
    VS(check_rows(n, row_bytes));
    VS(check_rows(C, row_bytes));
    for (size_t i = 0; i < n; i++)
        if (list_of_row[i] >= C) return fail(VS_EINVAL, "list_of_row[%zu]=%u >= C=%zu", i, list_of_row[i], C);
    vs_matrix *data = nullptr, *cent = nullptr;
    VS(vs_matrix_create(c, rows, n, row_bytes, &data));
    int rc = vs_matrix_create(c, centroids, C, row_bytes, &cent);
    int32_t *d_list = nullptr;
    uint64_t *d_ids = nullptr;
    if (rc == VS_OK && cudaMalloc(&d_list, n * 4 + 4) != cudaSuccess) rc = fail(VS_ENOMEM, "cudaMalloc list_of_row");
    if (rc == VS_OK && doc_ids && cudaMalloc(&d_ids, n * 8 + 8) != cudaSuccess) rc = fail(VS_ENOMEM, "cudaMalloc doc_ids");
    if (rc == VS_OK) {
        cudaError_t e = cudaMemcpyAsync(d_list, list_of_row, n * 4, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess && doc_ids) e = cudaMemcpyAsync(d_ids, doc_ids, n * 8, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(VS_ECUDA, "index upload: %s", cudaGetErrorString(e));
    }
    if (rc == VS_OK) rc = vs_index_build_dev(c, data, d_list, d_ids, 0, cent, out);
    if (d_list) cudaFree(d_list);
    if (d_ids) cudaFree(d_ids);
    vs_matrix_release(data);
    if (cent) vs_matrix_release(cent);
    return rc;
}

extern "C" int vs_index_build(vs_ctx *c, const uint8_t *rows, size_t n, size_t row_bytes, const uint64_t *doc_ids,
                              const uint64_t *list_offsets, const uint8_t *centroids, size_t C, vs_index **out) {
    VS(need_dev());
    if (!c || !rows || !list_offsets || !centroids || !out) return fail(VS_EINVAL, "null argument");
    VS(check_rows(n, row_bytes));
    VS(check_rows(C, row_bytes));
    if (list_offsets[0] != 0 || list_offsets[C] != n) return fail(VS_EINVAL, "list_offsets must start at 0 and end at n");
    for (size_t i = 0; i < C; i++)
        if (list_offsets[i] > list_offsets[i + 1]) return fail(VS_EINVAL, "list_offsets not monotone at %zu", i);
    vs_index *ix = new vs_index();
    ix->n = n;
    ix->C = C;
    int rc = vs_matrix_create(c, rows, n, row_bytes, &ix->data);
    if (rc == VS_OK) rc = vs_matrix_create(c, centroids, C, row_bytes, &ix->centroids);
    if (rc == VS_OK) {
        cudaError_t e = cudaMalloc(&ix->list_off, (C + 1) * 8);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ix->list_off, list_offsets, (C + 1) * 8, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess && doc_ids) {
            e = cudaMalloc(&ix->doc_ids, n * 8 + 8);
            if (e == cudaSuccess) e = cudaMemcpyAsync(ix->doc_ids, doc_ids, n * 8, cudaMemcpyHostToDevice, c->stream);
        }
        if (e == cudaSuccess) e = index_make_lens(c, ix);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(VS_ECUDA, "index upload: %s", cudaGetErrorString(e));
    }
    if (rc != VS_OK) {
        vs_index_release(ix);
        return rc;
    }
    *out = ix;
    return VS_OK;
}

// ---- Streaming loader: database -> HBM with centroid_id grouping (database/model.go:9-18; search.go:241-243) ----
__global__ void advance_cursor_kernel(const uint32_t *chunk_off, size_t C, uint64_t *cursor) {
    const size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (c < C) cursor[c] += chunk_off[c + 1] - chunk_off[c];
}

extern "C" int vs_index_create_empty(vs_ctx *c, const vs_matrix *centroids, const uint64_t *list_counts, vs_index **out) {
    VS(need_dev());
    if (!c || !centroids || !list_counts || !out) return fail(VS_EINVAL, "null argument");
    const size_t C = centroids->n;
    if (C == 0) return fail(VS_EEMPTY, "matrix rows are empty");  // compute.go:25-27
    std::vector<uint64_t> off(C + 1, 0);
    for (size_t l = 0; l < C; l++) off[l + 1] = off[l] + list_counts[l];
    const size_t n = off[C];
    vs_index *ix = new vs_index();
    ix->n = n;
    ix->C = C;
    ix->centroids = const_cast<vs_matrix *>(centroids);
    vs_matrix_retain(ix->centroids);
    int rc = matrix_alloc(n, (size_t)centroids->d, &ix->data);
    if (rc == VS_OK) {
        cudaError_t e = cudaMalloc(&ix->doc_ids, n * 8 + 8);
        if (e == cudaSuccess) e = cudaMalloc(&ix->list_off, (C + 1) * 8);
        if (e == cudaSuccess) e = cudaMalloc(&ix->fill_cursor, (C + 1) * 8);  // [C] cursors + the overflow flag
        if (e == cudaSuccess) e = cudaMemcpyAsync(ix->list_off, off.data(), (C + 1) * 8, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(ix->fill_cursor, 0, (C + 1) * 8, c->stream);
        ix->list_len = ix->fill_cursor;
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(VS_ECUDA, "empty index of %zu rows: %s", n, cudaGetErrorString(e));
    }
    if (rc != VS_OK) {
        vs_index_release(ix);
        return rc;
    }
    *out = ix;
    return VS_OK;
}

extern "C" int vs_index_fill_dev(vs_ctx *c, vs_index *ix, const vs_matrix *chunk, const int32_t *d_list_of_row, const uint64_t *d_doc_ids,
                                 uint64_t id_base) {
    VS(need_dev());
    if (!c || !ix || !chunk || !d_list_of_row) return fail(VS_EINVAL, "null argument");
    if (!ix->fill_cursor) return fail(VS_EINVAL, "the index was not created by vs_index_create_empty");
    if (chunk->d != ix->data->d) return fail(VS_EDIM, "data/centroid column size does not match: %d != %d", chunk->d, ix->data->d);
    const size_t m = chunk->n, C = ix->C;
    if (m == 0) return VS_OK;
    if (ix->filled + m > ix->n) return fail(VS_EINVAL, "%zu rows do not fit: %zu of %zu already placed", m, ix->filled, ix->n);
    const size_t sort_ws = sort_rows_ws_bytes(m);
    Arena a(c);
    VS(a.reserve(2 * Arena::pad(m * 4) + sort_ws + Arena::pad((C + 1) * 4) + 4096));
    uint32_t *d_order = a.take<uint32_t>(m);
    uint32_t *d_keys_sorted = a.take<uint32_t>(m);
    char *ws = a.take<char>(sort_ws);
    uint32_t *d_chunk_off = a.take<uint32_t>(C + 1);
    unsigned int *d_overflow = reinterpret_cast<unsigned int *>(ix->fill_cursor + C);
    // a list index >= C sorts behind every list and is caught by the row count below
    VS(sort_rows_by_key(c, reinterpret_cast<const uint32_t *>(d_list_of_row), m, 32, d_order, d_keys_sorted, ws, sort_ws));
    const unsigned cb = (unsigned)((C + 1 + 255) / 256);
    lower_bound_kernel<<<cb, 256, 0, c->stream>>>(d_keys_sorted, m, nullptr, d_chunk_off, C);
    CU(cudaGetLastError());
    c->launches++;
    VS(pinned_reserve(c, 64));
    unsigned int *h = static_cast<unsigned int *>(c->pinned);
    CU(cudaMemcpyAsync(h, d_chunk_off + C, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (h[0] != m) return fail(VS_EINVAL, "%zu row(s) of the chunk name a list >= %zu", m - (size_t)h[0], C);
    LAUNCH(c, launch_scatter_rows(chunk->view(), d_order, d_keys_sorted, d_chunk_off, ix->list_off, ix->fill_cursor, ix->data->codes,
                                  ix->data->hdr, ix->data->sums, d_doc_ids, id_base, ix->doc_ids, d_overflow, c->stream));
    advance_cursor_kernel<<<cb, 256, 0, c->stream>>>(d_chunk_off, C, ix->fill_cursor);
    CU(cudaGetLastError());
    c->launches++;
    CU(cudaMemcpyAsync(h, d_overflow, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (h[0] != 0) return fail(VS_EINVAL, "a list received more rows than vs_index_create_empty reserved for it");
    ix->filled += m;
    return VS_OK;
}

extern "C" int vs_index_fill(vs_ctx *c, vs_index *ix, const uint8_t *rows, size_t n, size_t row_bytes, const uint32_t *list_of_row,
                             const uint64_t *doc_ids, uint64_t id_base) {
    VS(need_dev());
    if (!c || !ix || !rows || !list_of_row) return fail(VS_EINVAL, "null argument");
    VS(check_rows(n, row_bytes));
    vs_matrix *chunk = nullptr;
    VS(vs_matrix_create(c, rows, n, row_bytes, &chunk));
    int32_t *d_list = nullptr;
    uint64_t *d_ids = nullptr;
    int rc = VS_OK;
    if (cudaMalloc(&d_list, n * 4 + 4) != cudaSuccess) rc = fail(VS_ENOMEM, "cudaMalloc list_of_row");
    if (rc == VS_OK && doc_ids && cudaMalloc(&d_ids, n * 8 + 8) != cudaSuccess) rc = fail(VS_ENOMEM, "cudaMalloc doc_ids");
    if (rc == VS_OK) {
        cudaError_t e = cudaMemcpyAsync(d_list, list_of_row, n * 4, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess && doc_ids) e = cudaMemcpyAsync(d_ids, doc_ids, n * 8, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(VS_ECUDA, "chunk upload: %s", cudaGetErrorString(e));
    }
    if (rc == VS_OK) rc = vs_index_fill_dev(c, ix, chunk, d_list, d_ids, id_base);
    cudaStreamSynchronize(c->stream);
    if (d_list) cudaFree(d_list);
    if (d_ids) cudaFree(d_ids);
    vs_matrix_release(chunk);
    return rc;
}

// ---- Upload: new rows into an existing index (server/upload.go:239-279) ----
// list_off of the merged store: every list keeps its rows and gets the uploaded rows assigned to it after them.
__global__ void merged_offsets_kernel(const uint64_t *old_off, const uint32_t *add_off, size_t C, uint64_t *new_off) {
    const size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (c <= C) new_off[c] = old_off[c] + add_off[c];
}
// order[i] = source of merged row i: a row of the old store, or (bit 31) an uploaded row, in upload order within a list.
__global__ void merge_order_kernel(const uint64_t *old_off, const uint32_t *add_off, const uint64_t *new_off,
                                   const uint32_t *add_order, size_t C, size_t n_total, uint32_t *order) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_total; i += (size_t)gridDim.x * blockDim.x) {
        size_t lo = 0, hi = C;  // first list boundary beyond i; new_off[0] = 0 <= i < new_off[C] = n_total
        while (lo < hi) {
            const size_t mid = (lo + hi) >> 1;
            if (new_off[mid] <= i) lo = mid + 1;
            else hi = mid;
        }
        const size_t l = lo - 1;
        const uint64_t p = i - new_off[l], old_len = old_off[l + 1] - old_off[l];
        order[i] = p < old_len ? (uint32_t)(old_off[l] + p) : (0x80000000u | add_order[add_off[l] + (p - old_len)]);
    }
}

static int index_upload_core(vs_ctx *c, const vs_index *ix, const vs_matrix *nm, const uint64_t *doc_ids, int64_t *assign_out,
                             vs_index *nx) {
    const size_t n = nm->n, C = ix->C, n_total = ix->n + n, d = (size_t)nm->d;
    const size_t sort_ws = sort_rows_ws_bytes(n);
    Arena a(c);
    VS(a.reserve(argmax_bytes(C, n, d) + 3 * Arena::pad(n * 4) + sort_ws + Arena::pad((C + 1) * 4) + Arena::pad(n_total * 4) +
                 Arena::pad(n * 8) + 4096));
    int32_t *d_assign = a.take<int32_t>(n);
    uint32_t *d_add_order = a.take<uint32_t>(n);
    uint32_t *d_keys_sorted = a.take<uint32_t>(n);
    char *ws = a.take<char>(sort_ws);
    uint32_t *d_add_off = a.take<uint32_t>(C + 1);
    uint32_t *d_order = a.take<uint32_t>(n_total);
    uint64_t *d_new_ids = doc_ids ? a.take<uint64_t>(n) : nullptr;
    // upload.go:245: centroids.MatrixCosineSimilarity(embeddings) -> nearest centroid of every new row, lowest index on ties
    VS(argmax_dev(c, a, ix->centroids->view(), nm->view(), d_assign, nullptr));
    VS(sort_rows_by_key(c, reinterpret_cast<const uint32_t *>(d_assign), n, bits_for(C), d_add_order, d_keys_sorted, ws, sort_ws));
    const unsigned cb = (unsigned)((C + 1 + 255) / 256);
    lower_bound_kernel<<<cb, 256, 0, c->stream>>>(d_keys_sorted, n, nullptr, d_add_off, C);
    CU(cudaGetLastError());
    merged_offsets_kernel<<<cb, 256, 0, c->stream>>>(ix->list_off, d_add_off, C, nx->list_off);
    CU(cudaGetLastError());
    merge_order_kernel<<<g_sm_count * 4, 256, 0, c->stream>>>(ix->list_off, d_add_off, nx->list_off, d_add_order, C, n_total, d_order);
    CU(cudaGetLastError());
    CU(index_make_lens(c, nx));
    c->launches += 4;
    if (doc_ids) CU(cudaMemcpyAsync(d_new_ids, doc_ids, n * 8, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, launch_merge_rows(ix->data->view(), ix->doc_ids, ix->id_base, nm->view(), d_new_ids, ix->id_base + ix->n, d_order, n_total,
                                nx->data->codes, nx->data->hdr, nx->data->sums, nx->doc_ids, c->stream));
    std::vector<int32_t> tmp;
    if (assign_out) {
        tmp.resize(n);
        CU(cudaMemcpyAsync(tmp.data(), d_assign, n * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < tmp.size(); i++) assign_out[i] = tmp[i];
    return VS_OK;
}

extern "C" int vs_index_upload(vs_ctx *c, const vs_index *ix, const uint8_t *rows, size_t n, size_t row_bytes, const uint64_t *doc_ids,
                               int64_t *assign_out, vs_index **out) {
    VS(need_dev());
    if (!c || !ix || !rows || !out) return fail(VS_EINVAL, "null argument");
    VS(check_rows(n, row_bytes));
    if (index_has_holes(ix)) return fail(VS_EINVAL, "the index has room to grow: vs_index_append adds rows in place");
    if (ix->C == 0) return fail(VS_EINVAL, "the index has no centroids");  // compute.go:26: NewMatrix panics on 0 rows
    if ((size_t)ix->centroids->d != row_bytes - 8)                          // cosine.go:77-79
        return fail(VS_EDIM, "matrix/matrix column size does not match: %d != %zu", ix->centroids->d, row_bytes - 8);
    const bool implicit = ix->implicit_ids || !ix->doc_ids;
    if (!implicit && !doc_ids) return fail(VS_EINVAL, "the index holds explicit document ids: doc_ids is required");
    if (ix->n + n > 0x7FFFFFFFull) return fail(VS_ERANGE, "n=%zu: a device matrix holds < 2^31 rows", ix->n + n);
    vs_matrix *nm = nullptr;
    VS(vs_matrix_create(c, rows, n, row_bytes, &nm));
    vs_index *nx = new vs_index();
    nx->n = ix->n + n;
    nx->C = ix->C;
    nx->centroids = ix->centroids;
    vs_matrix_retain(nx->centroids);
    nx->id_base = ix->id_base;
    nx->implicit_ids = implicit && !doc_ids;
    int rc = matrix_alloc(nx->n, (size_t)nm->d, &nx->data);
    if (rc == VS_OK && cudaMalloc(&nx->list_off, (nx->C + 1) * 8) != cudaSuccess) rc = fail(VS_ENOMEM, "cudaMalloc list offsets");
    if (rc == VS_OK && cudaMalloc(&nx->doc_ids, nx->n * 8 + 8) != cudaSuccess) rc = fail(VS_ENOMEM, "cudaMalloc document ids");
    if (rc == VS_OK) rc = index_upload_core(c, ix, nm, doc_ids, assign_out, nx);
    if (rc != VS_OK) cudaStreamSynchronize(c->stream);  // nothing may still read nm / nx when they are freed
    vs_matrix_release(nm);
    if (rc != VS_OK) {
        vs_index_release(nx);
        return rc;
    }
    *out = nx;
    return VS_OK;
}

// ---- Upload in place: lists with room to grow ---------------------------------------------------------------------
// vs_index_upload copies the whole store for every batch of new rows.  A service that ingests continuously keeps the
// store with slack behind every list instead (vs_index_with_room), appends new rows where their lists end
// (vs_index_append: the loader's scatter, preceded by upload.go:245's assignment) and copies only when a list is full.
__global__ void copy_lists_kernel(MatView src, const uint64_t *__restrict__ src_ids, uint64_t id_base,
                                  const uint64_t *__restrict__ src_off, const uint64_t *__restrict__ src_len,
                                  const uint64_t *__restrict__ dst_off, size_t C, uint8_t *__restrict__ codes,
                                  float2 *__restrict__ hdr, uint2 *__restrict__ sums, uint64_t *__restrict__ ids) {
    const size_t words_per_row = (size_t)src.d_pad >> 4;
    for (size_t l = blockIdx.y; l < C; l += gridDim.y) {
        const uint64_t s0 = src_off[l], d0 = dst_off[l], len = src_len[l];
        // the slice of this list that block x of the row copies
        const uint4 *sp = reinterpret_cast<const uint4 *>(src.codes + s0 * (size_t)src.d_pad);
        uint4 *dp = reinterpret_cast<uint4 *>(codes + d0 * (size_t)src.d_pad);
        const size_t words = len * words_per_row;
        for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) dp[w] = sp[w];
        for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < len; r += (size_t)gridDim.x * blockDim.x) {
            hdr[d0 + r] = src.hdr[s0 + r];
            sums[d0 + r] = src.sums[s0 + r];
            ids[d0 + r] = src_ids ? src_ids[s0 + r] : id_base + s0 + r;
        }
    }
}

extern "C" int vs_index_with_room(vs_ctx *c, const vs_index *ix, size_t percent, size_t min_rows, vs_index **out) {
    VS(need_dev());
    if (!c || !ix || !out) return fail(VS_EINVAL, "null argument");
    if (!ix->centroids || ix->C == 0) return fail(VS_EINVAL, "the index has no centroids");
    const size_t C = ix->C;
    std::vector<uint64_t> len(C), cap(C);
    CU(cudaMemcpyAsync(len.data(), ix->list_len, C * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    size_t live = 0;
    for (size_t l = 0; l < C; l++) {
        const size_t extra = len[l] * percent / 100;
        cap[l] = len[l] + (extra > min_rows ? extra : min_rows);
        live += len[l];
    }
    vs_index *nx = nullptr;
    VS(vs_index_create_empty(c, ix->centroids, cap.data(), &nx));
    nx->id_base = ix->id_base;
    nx->implicit_ids = ix->implicit_ids || !ix->doc_ids;
    if (nx->n > 0x3FFFFFFFull) {
        vs_index_release(nx);
        return fail(VS_ERANGE, "a device store holds < 2^30 rows per GPU");
    }
    const dim3 grid(8, (unsigned)(C < 4096 ? C : 4096));
    copy_lists_kernel<<<grid, 256, 0, c->stream>>>(ix->data->view(), ix->doc_ids, ix->id_base, ix->list_off, ix->list_len, nx->list_off, C,
                                                   nx->data->codes, nx->data->hdr, nx->data->sums, nx->doc_ids);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(nx->fill_cursor, ix->list_len, C * 8, cudaMemcpyDeviceToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    c->launches++;
    if (e != cudaSuccess) {
        vs_index_release(nx);
        return fail(VS_ECUDA, "index copy: %s", cudaGetErrorString(e));
    }
    nx->filled = live;
    *out = nx;
    return VS_OK;
}

// flag = 1 when some list cannot take the rows this chunk brings it
__global__ void lists_fit_kernel(const uint32_t *chunk_off, const uint64_t *list_off, const uint64_t *cursor, size_t C, unsigned int *flag) {
    const size_t l = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (l < C && cursor[l] + (chunk_off[l + 1] - chunk_off[l]) > list_off[l + 1] - list_off[l]) *flag = 1u;
}

extern "C" int vs_index_append(vs_ctx *c, vs_index *ix, const uint8_t *rows, size_t n, size_t row_bytes, const uint64_t *doc_ids,
                               int64_t *assign_out) {
    VS(need_dev());
    if (!c || !ix || !rows) return fail(VS_EINVAL, "null argument");
    VS(check_rows(n, row_bytes));
    if (!ix->fill_cursor) return fail(VS_EFULL, "the index was built in one piece: no room to grow (see vs_index_with_room)");
    if (ix->C == 0 || !ix->centroids) return fail(VS_EINVAL, "the index has no centroids");
    if ((size_t)ix->centroids->d != row_bytes - 8)  // cosine.go:77-79
        return fail(VS_EDIM, "matrix/matrix column size does not match: %d != %zu", ix->centroids->d, row_bytes - 8);
    if (!ix->implicit_ids && !doc_ids) return fail(VS_EINVAL, "the index holds explicit document ids: doc_ids is required");
    if (ix->filled + n > ix->n) return fail(VS_EFULL, "%zu rows do not fit: %zu of %zu places taken", n, ix->filled, ix->n);
    const size_t C = ix->C, d = (size_t)ix->centroids->d;
    vs_matrix *nm = nullptr;
    VS(vs_matrix_create(c, rows, n, row_bytes, &nm));
    auto body = [&]() -> int {
        const size_t sort_ws = sort_rows_ws_bytes(n);
        Arena a(c);
        VS(a.reserve(argmax_bytes(C, n, d) + 3 * Arena::pad(n * 4) + sort_ws + Arena::pad((C + 1) * 4) + Arena::pad(n * 8) + 4096));
        int32_t *d_assign = a.take<int32_t>(n);
        uint32_t *d_order = a.take<uint32_t>(n);
        uint32_t *d_keys_sorted = a.take<uint32_t>(n);
        char *ws = a.take<char>(sort_ws);
        uint32_t *d_chunk_off = a.take<uint32_t>(C + 1);
        uint64_t *d_new_ids = doc_ids ? a.take<uint64_t>(n) : nullptr;
        unsigned int *d_flag = a.take<unsigned int>(16);
        unsigned int *d_overflow = reinterpret_cast<unsigned int *>(ix->fill_cursor + C);
        // upload.go:245: nearest centroid of every new row, lowest index on ties
        VS(argmax_dev(c, a, ix->centroids->view(), nm->view(), d_assign, nullptr));
        VS(sort_rows_by_key(c, reinterpret_cast<const uint32_t *>(d_assign), n, bits_for(C), d_order, d_keys_sorted, ws, sort_ws));
        const unsigned cb = (unsigned)((C + 1 + 255) / 256);
        lower_bound_kernel<<<cb, 256, 0, c->stream>>>(d_keys_sorted, n, nullptr, d_chunk_off, C);
        CU(cudaGetLastError());
        CU(cudaMemsetAsync(d_flag, 0, 4, c->stream));
        lists_fit_kernel<<<cb, 256, 0, c->stream>>>(d_chunk_off, ix->list_off, ix->fill_cursor, C, d_flag);
        CU(cudaGetLastError());
        c->launches += 2;
        VS(pinned_reserve(c, 64));
        unsigned int *h = static_cast<unsigned int *>(c->pinned);
        CU(cudaMemcpyAsync(h, d_flag, 4, cudaMemcpyDeviceToHost, c->stream));
        std::vector<int32_t> tmp;
        if (assign_out) {
            tmp.resize(n);
            CU(cudaMemcpyAsync(tmp.data(), d_assign, n * 4, cudaMemcpyDeviceToHost, c->stream));
        }
        if (doc_ids) CU(cudaMemcpyAsync(d_new_ids, doc_ids, n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (h[0] != 0) return fail(VS_EFULL, "a list has no room for its new rows (the index is unchanged)");
        for (size_t i = 0; i < tmp.size(); i++) assign_out[i] = tmp[i];
        LAUNCH(c, launch_scatter_rows(nm->view(), d_order, d_keys_sorted, d_chunk_off, ix->list_off, ix->fill_cursor, ix->data->codes,
                                      ix->data->hdr, ix->data->sums, d_new_ids, ix->id_base + ix->filled, ix->doc_ids, d_overflow, c->stream));
        advance_cursor_kernel<<<cb, 256, 0, c->stream>>>(d_chunk_off, C, ix->fill_cursor);
        CU(cudaGetLastError());
        c->launches++;
        CU(cudaStreamSynchronize(c->stream));
        ix->filled += n;
        if (doc_ids) ix->implicit_ids = false;
        return VS_OK;
    };
    const int rc = body();
    if (rc != VS_OK) cudaStreamSynchronize(c->stream);
    vs_matrix_release(nm);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// search
static int kpl_for(size_t k) {
    if (k <= 32) return 1;
    if (k <= 64) return 2;
    if (k <= 128) return 4;
    return 0;
}

struct SearchBufs {
    uint32_t *probe;        // [nq][npe]
    uint32_t *qtiles;       // [nq] stage-2 tiles per query, written by stage 1
    Cand *partial1;         // stage-1 partial lists
    Cand *partial2;         // stage-2 partial lists
    unsigned int *tickets;  // [nq]
    double *qnorm;          // exact only
    uint32_t *q_select;     // exact only
    uint32_t *probe_keys;   // batched probe selection (probe.cu): [nq][C] similarity keys, or null
    unsigned int *flag_cnt; // [nq]
    uint32_t *flag_list;    // [nq][probe_flag_cap(C)]
    uint32_t *cand_keys, *cand_ids;  // [nq][segments][npe] segment survivors (many centroids only)
    // probe selection on the tensor cores (very many centroids, large batches): the GEMM pipeline with k = nprobe
    bool gemm_probe;
    GemmPlan gp;
    char *gp_scratch;
    uint64_t *gp_ids;       // [nq][npe]
    float *gp_sims;         // [nq][npe]
    int32_t *gp_counts;     // [nq]
    uint32_t *gp_status;    // [nq]
    // list stage of a batch, list-major (listmajor.cu)
    bool lm;
    LmParams lmp;
    uint32_t *lm_count, *lm_pair_off;
    uint32_t lm_items_cap;
    uint64_t *lm_seed_ids;   // exact top k of every query's nearest list (seeds the running bounds)
    float *lm_seed_sims;
    int32_t *lm_seed_counts;
    uint32_t *lm_seed_status;
    uint32_t *fused_keys;   // single query in one launch (fused.cu): [C + 4] centroid keys, [C] list extents + uncertified marks
    uint4 *fused_seginfo;
    int grid;               // blocks per stage launch
    int iters1, iters2;     // rows per lane group (tile height) of each stage
    int tile_rows1, tile_rows2;
};

// Tile height: 32 rows when there is plenty of work per warp, 16 or 8 when a launch would otherwise give
// each resident warp fewer than ~4 tiles (single-query latency).
static void pick_tile(int d_pad, size_t rows_total, int grid, int *iters, int *tile_rows) {
    const int G = stage_lanes_per_row(d_pad);
    if (G == 0) {
        *iters = 32;
        *tile_rows = 32;
        return;
    }
    const int NG = 32 / G;
    int tr = 32;
    const size_t warps = (size_t)grid * kStageWarps;
    while (tr > 8 && rows_total / tr < 4 * warps) tr >>= 1;
    *tile_rows = tr;
    *iters = tr / NG;
}

static int search_plan(const vs_index *ix, size_t nq, size_t npe, bool flat, SearchBufs *b) {
    b->grid = g_sm_count * 2;  // resident blocks per SM (__launch_bounds__ of stage_kernel)
    const int d_pad = ix->data->d_pad;
    pick_tile(d_pad, ix->C * nq, b->grid, &b->iters1, &b->tile_rows1);
    const size_t avg = ix->C ? (ix->n + ix->C - 1) / ix->C : 0;
    pick_tile(d_pad, (flat ? ix->n : npe * avg) * nq, b->grid, &b->iters2, &b->tile_rows2);
    return VS_OK;
}

constexpr size_t kProbeBatchMin = 8;  // from this many queries on the probe stage reads the centroid table once (probe.cu)
// Very many centroids x a large batch is a dense contraction (256 x 65 536 x 768 at config 5): the tensor-core pipeline
// of gemm.cu with k = nprobe (sampled thresholds, fused filter, certified selection) replaces the dp4a score kernel.
constexpr size_t kProbeGemmMinCentroids = 16384, kProbeGemmMinQueries = 64;
static bool use_probe_gemm(const vs_index *ix, size_t nq, size_t npe, bool flat) {
    return !flat && ix->centroids && nq >= kProbeGemmMinQueries && ix->C >= kProbeGemmMinCentroids && npe <= 128 &&
           gemm_supported(ix->centroids->view(), nq);
}
static GemmPlan probe_gemm_plan(const vs_index *ix, size_t nq, size_t npe) {
    return gemm_plan(ix->centroids->view(), nq, npe, true, 16, 1, 2048);
}
static size_t probe_gemm_bytes(const vs_index *ix, size_t nq, size_t npe) {
    const GemmPlan pl = probe_gemm_plan(ix, nq, npe);
    return gemm_scratch_bytes(pl, nq) + Arena::pad(nq * npe * 8) + Arena::pad(nq * npe * 4) + 2 * Arena::pad(nq * 4) + 1024;
}
static bool use_probe_batch(const vs_index *ix, size_t nq, size_t npe, bool flat) {
    return !flat && nq >= kProbeBatchMin && ix->centroids && probe_batch_supported(ix->centroids->view(), nq, npe);
}
static size_t probe_batch_bytes(size_t nq, size_t C, size_t npe) {
    return Arena::pad(nq * C * 4) + Arena::pad(nq * 4) + Arena::pad(nq * (size_t)probe_flag_cap(C) * 4) +
           2 * Arena::pad(nq * probe_segments(C) * npe * 4);
}

// The list stage of a batch goes list-major when lists are shared: with P (query, list) pairs over C lists the expected
// number of queries per probed list is (P/C) / (1 - exp(-P/C)): 1.27 at P = C/2, 2.3 at P = 2C (256 queries x 32 probes
// over 4096 lists).  Below that the query-major scan, which has no per-item cost, is as fast.
static bool g_lm_enabled = true;
extern "C" int vs_debug_set_list_major(int on) {
    g_lm_enabled = on != 0;
    return VS_OK;
}
// Queries per probed list (batch x nprobe / lists) from which the list-major scan takes its tensor-core form; 0 = never.
static long g_lm_dense_min = [] {
    const char *e = getenv("VS_LM_DENSE_MIN");
    return e ? atol(e) : 4l;
}();
extern "C" int vs_debug_set_lm_dense_min(int queries_per_list) {
    g_lm_dense_min = queries_per_list;
    return VS_OK;
}
static bool lm_eligible(const vs_index *ix, size_t nq, size_t npe, size_t k, bool flat) {
    // (VS_LM_MAX_NQ: measurement aid -- larger batches through the query-major scan)
    static const size_t max_nq = [] {
        const char *e = getenv("VS_LM_MAX_NQ");
        const long v = e ? atol(e) : 0;
        return (size_t)(v >= 16 && v <= (long)kMaxStageQueries ? v : kMaxStageQueries);
    }();
    if (!g_lm_enabled || flat || !ix->centroids || nq < 16 || nq > max_nq || nq < kProbeBatchMin) return false;
    if (nq * npe * 2 < ix->C || ix->n >= (1ull << 30) || ix->C >= (1ull << 30)) return false;
    return lm_supported(ix->data->d_pad, (int)k);
}
static size_t lm_bytes(const vs_index *ix, size_t nq, size_t npe) {
    return 2 * Arena::pad(ix->C * 4) + Arena::pad(lm_items_cap(ix->n, nq, npe) * sizeof(LmItem)) + Arena::pad(nq * npe * 4) +
           Arena::pad(nq * sizeof(SideConst)) + 4 * Arena::pad(nq * 4) + Arena::pad(nq * (size_t)kLmGbufCap * 16) + 2 * Arena::pad(8) +
           Arena::pad(nq * 32 * 8) + Arena::pad(nq * 32 * 4) + 4096;
}

static size_t search_bytes(size_t nq, size_t npe, int kpl1, int kpl2, int grid, size_t d, size_t C) {
    return Arena::pad((C + 4) * 4) + Arena::pad(C * 16) + Arena::pad(nq * npe * 4) + Arena::pad((nq + grid) * (size_t)32 * kpl1 * sizeof(Cand)) +
           Arena::pad((nq + grid) * (size_t)32 * kpl2 * sizeof(Cand)) + 3 * Arena::pad(nq * 4) + Arena::pad(nq * d * 8) + 4096;
}

// Probe list, next-stage tile count and status of one query from its certified top-npe centroids (tensor-core probe
// selection).  A query the filter could not answer (unusable header, fewer than npe centroids above its threshold) gets
// a valid placeholder list and the ambiguous bit: the caller's literal path redoes it.
__global__ void probe_from_topk_kernel(const uint64_t *ids, const float *sims, const int32_t *counts, const uint32_t *gstatus, int npe,
                                       uint32_t *out_probe, float *out_sims, uint32_t *out_qtiles, const uint64_t *list_len,
                                       uint32_t tile_rows, uint32_t *out_status) {
    __shared__ uint32_t s_tiles;
    const uint32_t q = blockIdx.x;
    const bool ok = !(gstatus[q] & kStatusNeedMore) && counts[q] == npe;
    if (threadIdx.x == 0) s_tiles = 0;
    __syncthreads();
    uint32_t mytiles = 0;
    for (int r = threadIdx.x; r < npe; r += blockDim.x) {
        const uint32_t L = ok ? (uint32_t)ids[(size_t)q * npe + r] : (uint32_t)r;
        out_probe[(size_t)q * npe + r] = L;
        if (out_sims) out_sims[(size_t)q * npe + r] = ok ? sims[(size_t)q * npe + r] : 0.0f;
        mytiles += (uint32_t)((list_len[L] + tile_rows - 1) / tile_rows);
    }
    for (int o = 16; o > 0; o >>= 1) mytiles += __shfl_xor_sync(0xFFFFFFFFu, mytiles, o);
    if ((threadIdx.x & 31) == 0 && mytiles) atomicAdd(&s_tiles, mytiles);
    __syncthreads();
    if (threadIdx.x == 0) {
        out_qtiles[q] = s_tiles ? s_tiles : 1u;
        out_status[q] = ok ? 0u : kStatusProbeAmbiguous;
    }
}

// Next-stage tile count of every query from probe lists that were selected elsewhere (one warp per query).  A list number
// outside the index (a corrupted exchange) becomes list 0 and flags the query for the literal path.
__global__ void probe_qtiles_kernel(uint32_t *probe, uint32_t nq, int npe, uint32_t C, const uint64_t *list_len, uint32_t tile_rows,
                                    uint32_t *out_qtiles, uint32_t *status) {
    const uint32_t q = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    uint32_t tiles = 0;
    bool bad = false;
    for (int r = lane; r < npe; r += 32) {
        uint32_t L = probe[(size_t)q * npe + r];
        if (L >= C) {
            bad = true;
            L = 0;
            probe[(size_t)q * npe + r] = 0;
        }
        tiles += (uint32_t)((list_len[L] + tile_rows - 1) / tile_rows);
    }
    for (int o = 16; o > 0; o >>= 1) tiles += __shfl_xor_sync(0xFFFFFFFFu, tiles, o);
    bad = __any_sync(0xFFFFFFFFu, bad);
    if (lane == 0) {
        out_qtiles[q] = tiles ? tiles : 1u;
        if (bad) status[q] |= kStatusProbeAmbiguous;
    }
}

// One query: probe stage, selection, list stage and top-k in ONE cooperative launch (fused.cu).  Queries it cannot decide
// (an uncertified score inside a window) get the usual status bits and are finished by the resolve path.
static bool g_fused_enabled = true;
extern "C" int vs_debug_set_fused(int on) {
    g_fused_enabled = on != 0;
    return VS_OK;
}
static bool fused_eligible(const vs_ctx *c, const vs_index *ix, size_t nq_launch, const uint32_t *d_select, size_t npe, bool flat,
                           bool exact, bool stage1_only, int kpl) {
    if (!g_fused_enabled || nq_launch != 1 || d_select || exact || stage1_only) return false;
    if (!fused_supported(ix->data->d_pad, kpl) || ix->n >= (1ull << 30)) return false;
    if (flat) return true;
    return ix->centroids && npe >= 1 && npe <= (size_t)kMaxSeg && npe < ix->C && ix->C < (1ull << 30);
}
static int fused_enqueue(vs_ctx *c, const vs_index *ix, const MatView &qv, size_t npe, size_t k, int kpl, bool flat,
                         const SearchBufs &b, uint64_t *d_ids, float *d_sims, int32_t *d_counts, uint32_t *d_status) {
    FusedParams p{};
    p.rows = ix->data->view();
    p.ids = ix->doc_ids;
    p.id_base = ix->id_base;
    if (!flat) {
        p.cent = ix->centroids->view();
        p.list_off = ix->list_off;
        p.list_len = ix->list_len;
        p.npe = (int)npe;
    } else {
        p.npe = 0;
        p.flat_start = 0;
        p.flat_count = ix->n;
    }
    p.query = qv;
    p.k = (int)k;
    p.pub = fused_pub((int)k, kpl);
    p.keys = b.fused_keys;
    p.seginfo = b.fused_seginfo;
    p.sync = c->d_fused_sync;
    p.partial = reinterpret_cast<uint4 *>(b.partial2);  // (1 + grid) * 32 * kpl entries of 16 bytes >= grid * pub
    p.out_ids = d_ids;
    p.out_sims = d_sims;
    p.out_counts = d_counts;
    p.out_status = d_status;
    p.out_probe = nullptr;
    p.trace = c->trace ? c->d_trace : nullptr;  // (both halves: the second holds per-chunk stamps)
    if (p.trace) CU(cudaMemsetAsync(p.trace, 0, 2 * kTraceBlocks * 16 * sizeof(unsigned long long), c->stream));
    VS(prof_mark(c));
    LAUNCH(c, launch_fused_search(p, kpl, g_sm_count, c->stream));
    VS(prof_mark(c));
    return VS_OK;
}

// Enqueue the two stages for nq_launch queries (all, or those listed in d_select).
static int search_enqueue(vs_ctx *c, const vs_index *ix, const MatView &qv, size_t nq_launch, const uint32_t *d_select,
                          size_t npe, size_t k, int kpl1, int kpl2, bool flat, bool exact, const SearchBufs &b,
                          uint64_t *d_ids, float *d_sims, int32_t *d_counts, uint32_t *d_status, float *d_probe_sims,
                          bool stage1_only, const uint32_t *d_probe_in = nullptr) {
    // d_probe_in: the probe lists were selected elsewhere (another device's share of a sharded probe stage, vs_probe_dev);
    // d_status already holds that stage's bits.  Only the plain batch form (no selection, certified arithmetic).
    const bool given = d_probe_in != nullptr && !flat && !exact && !d_select && nq_launch == qv.n && !stage1_only;
    if (!given && fused_eligible(c, ix, nq_launch, d_select, npe, flat, exact, stage1_only, kpl2))
        return fused_enqueue(c, ix, qv, npe, k, kpl2, flat, b, d_ids, d_sims, d_counts, d_status);
    StageParams p{};
    bool chained = false;
    p.queries = qv;
    p.qnorm = b.qnorm;
    p.q_select = d_select;
    p.nq = (int)nq_launch;
    p.tickets = c->d_tickets;
    p.out_status = d_status;
    p.fix_counter = c->d_fix_counter;
    const uint32_t tr1 = exact ? 32 : b.tile_rows1, tr2 = exact ? 32 : b.tile_rows2;
    if (given) {
        CU(cudaMemcpyAsync(b.probe, d_probe_in, qv.n * npe * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
        probe_qtiles_kernel<<<(unsigned)((qv.n + 3) / 4), 128, 0, c->stream>>>(b.probe, (uint32_t)qv.n, (int)npe, (uint32_t)ix->C, ix->list_len,
                                                                             tr2, b.qtiles, d_status);
        c->launches++;
    } else if (!flat && !exact && !d_select && b.gemm_probe && nq_launch == qv.n) {
        const MatView cent = ix->centroids->view();
        GemmBufs gb;
        gemm_take(b.gp_scratch, b.gp, qv.n, &gb);
        CU(gemm_enqueue_prepass(cent, qv, b.gp, gb, b.gp_status, g_sm_count, c->stream, &c->launches));
        CU(gemm_enqueue_filter(cent, qv, b.gp, gb, g_sm_count, c->stream, &c->launches));
        CU(gemm_enqueue_select(cent, nullptr, 0, true, qv, b.gp, gb, (int)npe, b.gp_ids, b.gp_sims, b.gp_counts, b.gp_status,
                               c->d_fix_counter, g_sm_count, c->stream, &c->launches));
        probe_from_topk_kernel<<<(unsigned)qv.n, 128, 0, c->stream>>>(b.gp_ids, b.gp_sims, b.gp_counts, b.gp_status, (int)npe, b.probe,
                                                                     d_probe_sims, b.qtiles, ix->list_len, tr2, d_status);
        c->launches++;
        if (stage1_only) return VS_OK;
    } else if (!flat && !exact && !d_select && b.probe_keys && nq_launch == qv.n) {
        // a batch: score every (query, centroid) pair with the table read once, then select per query (probe.cu)
        LAUNCH(c, launch_probe_batch(ix->centroids->view(), qv, (int)npe, b.probe_keys, b.flag_cnt, b.flag_list, probe_flag_cap(ix->C), b.cand_keys, b.cand_ids, b.probe,
                                     d_probe_sims, b.qtiles, ix->list_len, tr2, d_status, kStatusProbeAmbiguous, 1, c->d_fix_counter,
                                     g_sm_count, c->stream));
        c->launches++;
        if (stage1_only) return VS_OK;
    } else if (!flat) {
        p.rows = ix->centroids->view();
        p.ids = nullptr;
        p.id_base = 0;
        p.seg_list = nullptr;
        p.single_start = 0;
        p.single_count = ix->C;
        p.nseg = 1;
        p.qtiles = nullptr;
        p.uniform_tiles = (uint32_t)((ix->C + tr1 - 1) / tr1);
        p.iters = b.iters1;
        p.partial = b.partial1;
        p.mode = 1;
        p.k = (int)npe;
        p.out_probe = b.probe;
        p.out_sims = d_probe_sims;
        p.out_qtiles = b.qtiles;
        p.next_list_len = ix->list_len;
        p.next_tile_rows = tr2;
        p.status_bit = kStatusProbeAmbiguous;
        p.status_init = 1;
        p.trace = c->trace ? c->d_trace : nullptr;
        if (p.trace) CU(cudaMemsetAsync(p.trace, 0, kTraceBlocks * 16 * sizeof(unsigned long long), c->stream));
        // the list stage follows immediately on the stream: let its blocks be scheduled while this stage drains
        chained = !stage1_only && !c->trace && !c->profile && !(b.lm && !exact && !d_select && nq_launch == qv.n);
        p.pdl = chained ? 1 : 0;
        LAUNCH(c, launch_stage(p, kpl1, exact, b.grid, c->stream));
        if (stage1_only) return VS_OK;
    }
    if (b.lm && !flat && !exact && !d_select && nq_launch == qv.n) {
        // a batch that shares lists: every probed list is read once for all the queries that probe it (listmajor.cu)
        LmParams lp = b.lmp;
        lp.rows = ix->data->view();
        lp.ids = ix->doc_ids;
        lp.id_base = ix->id_base;
        lp.queries = qv;
        lp.k = (int)k;
        lp.pub = fused_pub((int)k, 1);
        static const bool seed_by_stage = getenv("VS_LM_SEED_STAGE") != nullptr;  // (the first form of the seed, kept for A/B timing)
        if (!seed_by_stage) {
            if (!c->side_stream) {
                CU(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
                CU(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
                CU(cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming));
            }
            CU(cudaEventRecord(c->fork_ev, c->stream));  // the probe lists are ready here: the seed below forks off this point
        }
        // several queries per probed list: the tensor-core form of the scan (listmajor.cu, lm_dense_kernel)
        const bool dense = g_lm_dense_min > 0 && lm_dense_supported(ix->data->d_pad) && qv.n * npe >= (size_t)g_lm_dense_min * ix->C;
        if (dense) {
            // a query's bound is recomputed from its own candidate list when that has grown to 512, 1024 and 2048 entries (a weak
            // seed then costs a longer list, not an overflow): those reads must not see a previous step's entries
            lp.tighten_at = 512;
            // (here, in front of the inversion, the zeroing hides behind the seed kernel of the side stream)
            CU(cudaMemset2DAsync(lp.gbuf, (size_t)lp.gcap * sizeof(uint4), 0, (size_t)4 * lp.tighten_at * sizeof(uint4), qv.n, c->stream));
        }
        CU(lm_enqueue_prepare(lp, b.probe, (uint32_t)qv.n, (uint32_t)npe, (uint32_t)ix->C, ix->list_off, ix->list_len, b.lm_count, b.lm_pair_off,
                              b.lm_items_cap, c->stream, &c->launches));
        // The k-th best document among the first 512 rows of every query's nearest list -- the ~2 % quantile of the query's
        // scores -- seeds the query's running bound, so that the list-major warps keep almost nothing from their first row
        // on (without it a warp, which sees ~128 rows of an item, keeps and sorts half of them).
        if (!seed_by_stage) {
            // beside the inversion (lm_enqueue_prepare above): both need only the probe lists, and each is a latency chain of
            // a few tens of microseconds on a fraction of the SMs
            CU(cudaStreamWaitEvent(c->side_stream, c->fork_ev, 0));
            static const uint32_t seed_rows = [] {  // (VS_LM_SEED_ROWS: measurement aid)
                const char *e = getenv("VS_LM_SEED_ROWS");
                const long v = e ? atol(e) : 0;
                return (uint32_t)(v >= 32 && v <= 512 ? v : 0);
            }();
            // 512 rows for the dp4a form (whose per-warp lists want a tight start); 256 for the tensor-core form, which
            // recomputes a query's bound from its candidate list as that grows -- there the seed kernel's own time (164 us
            // at 1024 queries x 512 rows, a constant of the step that does not shrink with the shard) matters more
            const uint32_t seed_n = seed_rows ? seed_rows : (dense ? 256u : 512u);
            CU(lm_enqueue_seed_scan(lp, b.probe, (uint32_t)qv.n, (uint32_t)npe, ix->list_off, ix->list_len, seed_n, c->side_stream,
                                    &c->launches));
            CU(cudaEventRecord(c->join_ev, c->side_stream));
            CU(cudaStreamWaitEvent(c->stream, c->join_ev, 0));
        } else {
            // an exact top k from the query-major kernel, first probe only
            StageParams sp{};
            sp.queries = qv;
            sp.nq = (int)qv.n;
            sp.tickets = c->d_tickets;
            sp.fix_counter = c->d_fix_counter;
            sp.rows = ix->data->view();
            sp.ids = ix->doc_ids;
            sp.id_base = ix->id_base;
            sp.seg_list = b.probe;
            sp.seg_stride = (int)npe;
            sp.list_off = ix->list_off;
            sp.list_len = ix->list_len;
            sp.nseg = 1;
            sp.seg_cap = 512;
            sp.qtiles = b.qtiles;  // (tiles of all npe lists: an over-estimate only spreads the same work over fewer blocks)
            sp.iters = b.iters2;
            sp.partial = b.partial2;
            sp.mode = 0;
            sp.k = (int)k;
            sp.out_ids = b.lm_seed_ids;
            sp.out_sims = b.lm_seed_sims;
            sp.out_counts = b.lm_seed_counts;
            sp.out_status = b.lm_seed_status;
            sp.status_bit = kStatusListAmbiguous;
            sp.status_init = 1;
            LAUNCH(c, launch_stage(sp, kpl2, false, b.grid, c->stream));
            CU(lm_enqueue_seed(lp, b.lm_seed_sims, b.lm_seed_counts, (uint32_t)qv.n, c->stream, &c->launches));
        }
        VS(prof_mark(c));
        CU(lm_enqueue_scan(lp, g_sm_count, c->stream, &c->launches, dense));
        VS(prof_mark(c));
        CU(lm_enqueue_final(lp, (uint32_t)qv.n, d_ids, d_sims, d_counts, d_status, c->d_fix_counter, c->stream, &c->launches));
        return VS_OK;
    }
    p.rows = ix->data->view();
    p.ids = ix->doc_ids;
    p.id_base = ix->id_base;
    if (flat) {
        p.seg_list = nullptr;
        p.single_start = 0;
        p.single_count = ix->n;
        p.nseg = 1;
        p.qtiles = nullptr;
        p.uniform_tiles = (uint32_t)((ix->n + tr2 - 1) / tr2);
        if (p.uniform_tiles == 0) p.uniform_tiles = 1;
    } else {
        p.seg_list = b.probe;
        p.seg_stride = (int)npe;
        p.list_off = ix->list_off;
        p.list_len = ix->list_len;
        p.nseg = (int)npe;
        p.qtiles = b.qtiles;
        p.uniform_tiles = 0;
    }
    p.iters = b.iters2;
    p.partial = b.partial2;
    p.mode = 0;
    p.k = (int)k;
    p.out_ids = d_ids;
    p.out_sims = d_sims;
    p.out_counts = d_counts;
    p.out_probe = nullptr;
    p.out_qtiles = nullptr;
    p.status_bit = kStatusListAmbiguous;
    p.status_init = flat ? 1 : 0;
    p.pdl = chained ? 2 : 0;
    p.trace = c->trace ? c->d_trace + kTraceBlocks * 16 : nullptr;
    if (p.trace) CU(cudaMemsetAsync(p.trace, 0, kTraceBlocks * 16 * sizeof(unsigned long long), c->stream));
    VS(prof_mark(c));
    LAUNCH(c, launch_stage(p, kpl2, exact, b.grid, c->stream));
    VS(prof_mark(c));
    return VS_OK;
}

struct SearchSetup {
    size_t npe, k;
    bool flat;
    int kpl1, kpl2;
    SearchBufs b;
};

static int search_setup(vs_ctx *c, Arena &a, const vs_index *ix, size_t nq, size_t nprobe, size_t k, size_t extra_bytes,
                        SearchSetup *s, bool rank_all = false) {
    if (nq == 0) return fail(VS_EINVAL, "nq == 0");
    if (nq > (size_t)kMaxStageQueries) return fail(VS_ERANGE, "nq=%zu: at most %d queries per call", nq, kMaxStageQueries);
    if (k == 0) return fail(VS_EINVAL, "k == 0");
    if (ix->n > 0x3FFFFFFFull) return fail(VS_ERANGE, "a device store holds < 2^30 rows per GPU");
    if (nprobe == 0) nprobe = 1;  // search.go:118-119
    // rank_all: stage 1 only, the caller wants the ranked list itself.  A store with holes between its lists is never
    // scanned as one piece: every list is probed instead (same rows, same result).
    s->flat = nprobe >= ix->C && !rank_all && !index_has_holes(ix);
    s->npe = nprobe >= ix->C ? ix->C : nprobe;
    s->kpl2 = kpl_for(k);
    s->kpl1 = s->flat ? 1 : kpl_for(s->npe);
    if (!s->kpl2) return fail(VS_ERANGE, "k=%zu: at most 128 hits (Count+Offset) per query", k);
    if (!s->kpl1)
        return fail(VS_ERANGE, "nprobe=%zu: at most 128 probed lists unless nprobe >= number of lists%s", nprobe,
                    index_has_holes(ix) ? " (and the lists are back to back: this index has room to grow; vs_search serves it)" : "");
    VS(search_plan(ix, nq, s->npe, s->flat, &s->b));
    VS(a.reserve(a.off + extra_bytes + search_bytes(nq, s->npe, s->kpl1, s->kpl2, s->b.grid, ix->data->d, ix->C) +
                 (use_probe_gemm(ix, nq, s->npe, s->flat) ? probe_gemm_bytes(ix, nq, s->npe)
                  : use_probe_batch(ix, nq, s->npe, s->flat) ? probe_batch_bytes(nq, ix->C, s->npe) : 0) +
                 (lm_eligible(ix, nq, s->npe, k, s->flat) ? lm_bytes(ix, nq, s->npe) : 0)));
    s->k = k;
    return VS_OK;
}

static void search_take(Arena &a, const vs_index *ix, size_t nq, SearchSetup *s) {
    s->b.probe = a.take<uint32_t>(nq * s->npe);
    s->b.qtiles = a.take<uint32_t>(nq);
    s->b.partial1 = a.take<Cand>((nq + s->b.grid) * (size_t)32 * s->kpl1);
    s->b.partial2 = a.take<Cand>((nq + s->b.grid) * (size_t)32 * s->kpl2);
    s->b.tickets = a.take<unsigned int>(nq);
    s->b.qnorm = a.take<double>(nq * (size_t)ix->data->d);
    s->b.q_select = a.take<uint32_t>(nq);
    s->b.lm = lm_eligible(ix, nq, s->npe, s->k, s->flat);
    if (s->b.lm) {
        LmParams &lp = s->b.lmp;
        lp = LmParams{};
        s->b.lm_count = a.take<uint32_t>(ix->C);
        s->b.lm_pair_off = a.take<uint32_t>(ix->C);
        s->b.lm_items_cap = (uint32_t)lm_items_cap(ix->n, nq, s->npe);
        lp.items = a.take<LmItem>(s->b.lm_items_cap);
        lp.pairs = a.take<uint32_t>(nq * s->npe);
        lp.sides = a.take<SideConst>(nq);
        lp.gthr = a.take<uint32_t>(nq);
        lp.gcnt = a.take<unsigned int>(nq);
        lp.gbuf = a.take<uint4>(nq * (size_t)kLmGbufCap);
        lp.nitems = a.take<uint32_t>(2);
        lp.next_item = a.take<unsigned int>(2);
        lp.gcap = kLmGbufCap;
        s->b.lm_seed_ids = a.take<uint64_t>(nq * 32);
        s->b.lm_seed_sims = a.take<float>(nq * 32);
        s->b.lm_seed_counts = a.take<int32_t>(nq);
        s->b.lm_seed_status = a.take<uint32_t>(nq);
    }
    s->b.fused_keys = a.take<uint32_t>(ix->C + 4);
    s->b.fused_seginfo = a.take<uint4>(ix->C);
    s->b.probe_keys = nullptr;
    s->b.gemm_probe = use_probe_gemm(ix, nq, s->npe, s->flat);
    if (s->b.gemm_probe) {
        s->b.gp = probe_gemm_plan(ix, nq, s->npe);
        s->b.gp_scratch = a.take<char>(gemm_scratch_bytes(s->b.gp, nq));
        s->b.gp_ids = a.take<uint64_t>(nq * s->npe);
        s->b.gp_sims = a.take<float>(nq * s->npe);
        s->b.gp_counts = a.take<int32_t>(nq);
        s->b.gp_status = a.take<uint32_t>(nq);
    } else if (use_probe_batch(ix, nq, s->npe, s->flat)) {
        s->b.probe_keys = a.take<uint32_t>(nq * ix->C);
        s->b.flag_cnt = a.take<unsigned int>(nq);
        s->b.flag_list = a.take<uint32_t>(nq * (size_t)probe_flag_cap(ix->C));
        s->b.cand_keys = a.take<uint32_t>(nq * probe_segments(ix->C) * s->npe);
        s->b.cand_ids = a.take<uint32_t>(nq * probe_segments(ix->C) * s->npe);
    }
}

// Finish the queries whose status is non-zero with literal arithmetic. h_status: host copy of d_status.
static int search_resolve_flagged(vs_ctx *c, const vs_index *ix, const MatView &qv, size_t nq, const uint32_t *h_status,
                                  const SearchSetup &s, size_t k, uint64_t *d_ids, float *d_sims, int32_t *d_counts,
                                  uint32_t *d_status, float *d_probe_sims, bool stage1_only, int *n_resolved) {
    std::vector<uint32_t> sel;
    for (size_t i = 0; i < nq; i++)
        if (h_status[i] & (kStatusProbeAmbiguous | kStatusListAmbiguous)) sel.push_back((uint32_t)i);
    if (n_resolved) *n_resolved = (int)sel.size();
    if (sel.empty()) return VS_OK;
    CU(cudaMemcpyAsync(s.b.q_select, sel.data(), sel.size() * 4, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, launch_query_normalize(qv, s.b.qnorm, c->stream));
    VS(search_enqueue(c, ix, qv, sel.size(), s.b.q_select, s.npe, k, s.kpl1, s.kpl2, s.flat, true, s.b, d_ids, d_sims,
                      d_counts, d_status, d_probe_sims, stage1_only));
    CU(cudaStreamSynchronize(c->stream));  // sel (host) must outlive the copy
    c->slowpath += sel.size();
    return VS_OK;
}

static int search_check(vs_ctx *c, const vs_index *ix) {
    VS(need_dev());
    if (!c || !ix) return fail(VS_EINVAL, "null argument");
    return VS_OK;
}

extern "C" int vs_search_dev(vs_ctx *c, const vs_index *ix, const vs_matrix *queries, size_t nprobe, size_t k,
                             uint64_t *d_ids, float *d_sims, int32_t *d_counts, uint32_t *d_status) {
    VS(search_check(c, ix));
    if (!queries || !d_ids || !d_sims || !d_counts || !d_status) return fail(VS_EINVAL, "null argument");
    if (queries->d != ix->data->d)
        return fail(VS_EDIM, "vector/matrix column size does not match: %d != %d", queries->d, ix->data->d);
    const size_t nq = queries->n;
    Arena a(c);
    SearchSetup s;
    VS(search_setup(c, a, ix, nq, nprobe, k, 0, &s));
    search_take(a, ix, nq, &s);
    return search_enqueue(c, ix, queries->view(), nq, nullptr, s.npe, k, s.kpl1, s.kpl2, s.flat, false, s.b, d_ids, d_sims,
                          d_counts, d_status, nullptr, false);
}

// The probe stage alone, device to device: the nprobe nearest lists of every query (search.go:205-223) and the stage's
// status bits.  With vs_search_dev_probed it lets G devices that hold stripes of one index share the probe stage: each
// scores the centroid table for 1/G of the batch, the lists are exchanged, and every device scans its stripe for all.
extern "C" int vs_probe_dev(vs_ctx *c, const vs_index *ix, const vs_matrix *queries, size_t nprobe, uint32_t *d_probe,
                            uint32_t *d_status) {
    VS(search_check(c, ix));
    if (!queries || !d_probe || !d_status) return fail(VS_EINVAL, "null argument");
    if (queries->d != ix->data->d)
        return fail(VS_EDIM, "vector/matrix column size does not match: %d != %d", queries->d, ix->data->d);
    if (!ix->centroids || nprobe == 0 || nprobe >= ix->C) return fail(VS_ERANGE, "vs_probe_dev: 1 <= nprobe < number of lists");
    const size_t nq = queries->n;
    Arena a(c);
    SearchSetup s;
    VS(search_setup(c, a, ix, nq, nprobe, 1, 0, &s, true));
    search_take(a, ix, nq, &s);
    VS(search_enqueue(c, ix, queries->view(), nq, nullptr, s.npe, 1, s.kpl1, s.kpl2, s.flat, false, s.b, nullptr, nullptr, nullptr,
                      d_status, nullptr, true));
    CU(cudaMemcpyAsync(d_probe, s.b.probe, nq * s.npe * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    return VS_OK;
}

// The list stage alone for probe lists selected elsewhere (vs_probe_dev on this or another device): d_probe = [nq][nprobe]
// list numbers, d_status = the probe stage's status words on entry, the search's on return (as vs_search_dev leaves them;
// vs_search_resolve finishes flagged queries the same way, it redoes both stages).
extern "C" int vs_search_dev_probed(vs_ctx *c, const vs_index *ix, const vs_matrix *queries, size_t nprobe, size_t k,
                                    const uint32_t *d_probe, uint64_t *d_ids, float *d_sims, int32_t *d_counts, uint32_t *d_status) {
    VS(search_check(c, ix));
    if (!queries || !d_probe || !d_ids || !d_sims || !d_counts || !d_status) return fail(VS_EINVAL, "null argument");
    if (queries->d != ix->data->d)
        return fail(VS_EDIM, "vector/matrix column size does not match: %d != %d", queries->d, ix->data->d);
    if (!ix->centroids || nprobe == 0 || nprobe >= ix->C) return fail(VS_ERANGE, "vs_search_dev_probed: 1 <= nprobe < number of lists");
    const size_t nq = queries->n;
    Arena a(c);
    SearchSetup s;
    VS(search_setup(c, a, ix, nq, nprobe, k, 0, &s));
    search_take(a, ix, nq, &s);
    return search_enqueue(c, ix, queries->view(), nq, nullptr, s.npe, k, s.kpl1, s.kpl2, s.flat, false, s.b, d_ids, d_sims,
                          d_counts, d_status, nullptr, false, d_probe);
}

extern "C" int vs_search_resolve(vs_ctx *c, const vs_index *ix, const vs_matrix *queries, size_t nprobe, size_t k,
                                 uint64_t *d_ids, float *d_sims, int32_t *d_counts, uint32_t *d_status, int *n_resolved_out) {
    VS(search_check(c, ix));
    if (!queries || !d_status) return fail(VS_EINVAL, "null argument");
    const size_t nq = queries->n;
    VS(pinned_reserve(c, nq * 4));
    uint32_t *h_status = static_cast<uint32_t *>(c->pinned);
    CU(cudaMemcpyAsync(h_status, d_status, nq * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    bool any = false;
    for (size_t i = 0; i < nq; i++) any |= (h_status[i] & (kStatusProbeAmbiguous | kStatusListAmbiguous)) != 0;
    if (n_resolved_out) *n_resolved_out = 0;
    if (!any) return VS_OK;
    Arena a(c);
    SearchSetup s;
    VS(search_setup(c, a, ix, nq, nprobe, k, 0, &s));
    search_take(a, ix, nq, &s);
    return search_resolve_flagged(c, ix, queries->view(), nq, h_status, s, k, d_ids, d_sims, d_counts, d_status, nullptr,
                                  false, n_resolved_out);
}

constexpr size_t kGemmMinQueries = 64;    // below this the per-query scans (HBM-bound) are faster
constexpr size_t kGemmMinRows = 1u << 16;
static int gemm_search_host(vs_ctx *c, const vs_index *ix, const uint8_t *queries, size_t nq, size_t k, uint64_t *ids_out,
                            float *sims_out, int32_t *counts_out);

static int search_host(vs_ctx *c, const vs_index *ix, const uint8_t *queries, size_t nq, size_t nprobe, size_t k,
                       uint64_t *ids_out, float *sims_out, int32_t *counts_out, uint32_t *probe_out, float *probe_sims_out,
                       bool stage1_only) {
    VS(search_check(c, ix));
    if (!queries) return fail(VS_EINVAL, "queries is null");
    // a query batch over the whole store is a dense contraction: tensor cores (gemm.cu) instead of nq scans
    if (!stage1_only && nprobe >= ix->C && nq >= kGemmMinQueries && ix->n >= kGemmMinRows && k <= 128 && !index_has_holes(ix) &&
        gemm_supported(ix->data->view(), nq))
        return gemm_search_host(c, ix, queries, nq, k, ids_out, sims_out, counts_out);
    const size_t row_bytes = 8 + (size_t)ix->data->d;
    Arena a(c);
    SearchSetup s;
    const size_t out_bytes = Arena::pad(nq * k * 8) + Arena::pad(nq * k * 4) + 2 * Arena::pad(nq * 4) + Arena::pad(nq * 128 * 4) + 4096;
    VS(search_setup(c, a, ix, nq, nprobe, k, temp_matrix_bytes(nq, row_bytes) + out_bytes, &s, stage1_only));
    MatView qv;
    VS(temp_matrix(c, a, queries, nq, row_bytes, &qv));
    search_take(a, ix, nq, &s);
    uint64_t *d_ids = a.take<uint64_t>(nq * k);
    float *d_sims = a.take<float>(nq * k);
    int32_t *d_counts = a.take<int32_t>(nq);
    uint32_t *d_status = a.take<uint32_t>(nq);
    float *d_psims = a.take<float>(nq * s.npe);
    VS(search_enqueue(c, ix, qv, nq, nullptr, s.npe, k, s.kpl1, s.kpl2, s.flat, false, s.b, d_ids, d_sims, d_counts, d_status,
                      stage1_only ? d_psims : nullptr, stage1_only));
    // hits, counts and status words sit next to each other in the arena: ONE copy into pinned memory brings them all back
    // (four separate copies into pageable memory cost more than a single-query search itself)
    const size_t span = stage1_only ? nq * 4 : (size_t)(reinterpret_cast<char *>(d_status + nq) - reinterpret_cast<char *>(d_ids));
    VS(pinned_reserve(c, span + 16));
    char *hp = static_cast<char *>(c->pinned);
    uint32_t *h_status = stage1_only ? reinterpret_cast<uint32_t *>(hp)
                                     : reinterpret_cast<uint32_t *>(hp + (reinterpret_cast<char *>(d_status) - reinterpret_cast<char *>(d_ids)));
    auto fetch = [&]() -> int {
        if (!stage1_only) {
            CU(cudaMemcpyAsync(hp, d_ids, span, cudaMemcpyDeviceToHost, c->stream));
        } else {
            CU(cudaMemcpyAsync(h_status, d_status, nq * 4, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaMemcpyAsync(probe_out, s.b.probe, nq * s.npe * 4, cudaMemcpyDeviceToHost, c->stream));
            if (probe_sims_out) CU(cudaMemcpyAsync(probe_sims_out, d_psims, nq * s.npe * 4, cudaMemcpyDeviceToHost, c->stream));
        }
        CU(cudaStreamSynchronize(c->stream));
        return VS_OK;
    };
    VS(fetch());
    bool any = false;
    for (size_t i = 0; i < nq; i++) any |= (h_status[i] & (kStatusProbeAmbiguous | kStatusListAmbiguous)) != 0;
    if (any) {
        int nres = 0;
        VS(search_resolve_flagged(c, ix, qv, nq, h_status, s, k, d_ids, d_sims, d_counts, d_status,
                                  stage1_only ? d_psims : nullptr, stage1_only, &nres));
        VS(fetch());
    }
    if (!stage1_only) {
        memcpy(ids_out, hp, nq * k * 8);
        memcpy(sims_out, hp + (reinterpret_cast<char *>(d_sims) - reinterpret_cast<char *>(d_ids)), nq * k * 4);
        memcpy(counts_out, hp + (reinterpret_cast<char *>(d_counts) - reinterpret_cast<char *>(d_ids)), nq * 4);
    }
    return VS_OK;
}

// Requests beyond what the fused top-k kernels hold (more than 128 probed lists without probing all of them, or more than
// 128 hits: server/search.go:116-122 accepts any Centroids value and Count+Offset is unbounded by Offset).  Rare and
// served for completeness, not speed: the exact float32 similarity of every centroid and of every row (vs_cosine_1xN's
// kernels, bit-equal to compute/cosine.go:13-57) comes back to the host, which cuts the probe list (search.go:220-223),
// walks the probed lists, sorts (similarity desc, id asc), keeps one hit per document and truncates (search.go:256-270).
static int search_wide_host(vs_ctx *c, const vs_index *ix, const uint8_t *queries, size_t nq, size_t nprobe, size_t k,
                            uint64_t *ids_out, float *sims_out, int32_t *counts_out) {
    VS(search_check(c, ix));
    if (!queries) return fail(VS_EINVAL, "queries is null");
    if (nq == 0 || k == 0) return fail(VS_EINVAL, "nq == 0 or k == 0");
    if (nprobe == 0) nprobe = 1;
    const size_t n = ix->n, C = ix->C, rb = 8 + (size_t)ix->data->d;
    const bool flat = (nprobe >= C && !index_has_holes(ix)) || !ix->centroids;
    if (nprobe > C) nprobe = C;
    std::vector<uint64_t> off(C + 1, 0), len(C, 0), ids;
    if (!flat) {
        CU(cudaMemcpy(off.data(), ix->list_off, (C + 1) * 8, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(len.data(), ix->list_len, C * 8, cudaMemcpyDeviceToHost));
    }
    if (ix->doc_ids) {
        ids.resize(n);
        CU(cudaMemcpy(ids.data(), ix->doc_ids, n * 8, cudaMemcpyDeviceToHost));
    }
    std::vector<float> csims(C), rsims(n);
    std::vector<uint32_t> order;
    struct Hit {
        uint32_t key;
        uint64_t id;
    };
    std::vector<Hit> hits;
    for (size_t q = 0; q < nq; q++) {
        const uint8_t *qrow = queries + q * rb;
        VS(vs_cosine_1xN(c, qrow, rb, ix->data, rsims.data()));
        hits.clear();
        auto take = [&](size_t lo, size_t hi) {
            for (size_t r = lo; r < hi; r++) hits.push_back(Hit{f32_to_key(rsims[r]), ix->doc_ids ? ids[r] : ix->id_base + r});
        };
        if (flat) {
            take(0, n);
        } else {
            VS(vs_cosine_1xN(c, qrow, rb, ix->centroids, csims.data()));
            order.resize(C);
            for (size_t i = 0; i < C; i++) order[i] = (uint32_t)i;
            std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
                const uint32_t ka = f32_to_key(csims[a]), kb = f32_to_key(csims[b]);
                return ka > kb || (ka == kb && a < b);
            });
            for (size_t s = 0; s < nprobe; s++) take(off[order[s]], off[order[s]] + len[order[s]]);
        }
        std::sort(hits.begin(), hits.end(), [](const Hit &a, const Hit &b) { return a.key > b.key || (a.key == b.key && a.id < b.id); });
        size_t cnt = 0;
        if (!ix->doc_ids) {  // ids are distinct
            for (; cnt < k && cnt < hits.size(); cnt++) {
                ids_out[q * k + cnt] = hits[cnt].id;
                sims_out[q * k + cnt] = key_to_f32(hits[cnt].key);
            }
        } else {
            std::vector<uint64_t> seen;  // (k is small next to the candidate list; sorted insert keeps the membership test cheap)
            for (size_t i = 0; i < hits.size() && cnt < k; i++) {
                auto it = std::lower_bound(seen.begin(), seen.end(), hits[i].id);
                if (it != seen.end() && *it == hits[i].id) continue;
                seen.insert(it, hits[i].id);
                ids_out[q * k + cnt] = hits[i].id;
                sims_out[q * k + cnt] = key_to_f32(hits[i].key);
                cnt++;
            }
        }
        counts_out[q] = (int32_t)cnt;
    }
    return VS_OK;
}

extern "C" int vs_search(vs_ctx *c, const vs_index *ix, const uint8_t *queries, size_t nq, size_t nprobe, size_t k,
                         uint64_t *ids_out, float *sims_out, int32_t *counts_out) {
    if (!ids_out || !sims_out || !counts_out) return fail(VS_EINVAL, "null output");
    if (ix && (k > 128 || (nprobe > 128 && (nprobe < ix->C || (index_has_holes(ix) && ix->C > 128)))))
        return search_wide_host(c, ix, queries, nq, nprobe, k, ids_out, sims_out, counts_out);
    // more queries than one launch takes: in turns of kMaxStageQueries
    const size_t rb = ix ? 8 + (size_t)ix->data->d : 0;
    for (size_t q0 = 0; q0 < nq || q0 == 0; q0 += kMaxStageQueries) {
        const size_t m = nq - q0 < (size_t)kMaxStageQueries ? nq - q0 : (size_t)kMaxStageQueries;
        VS(search_host(c, ix, queries ? queries + q0 * rb : nullptr, m, nprobe, k, ids_out + q0 * k, sims_out + q0 * k, counts_out + q0,
                       nullptr, nullptr, false));
        if (nq == 0) break;
    }
    return VS_OK;
}

extern "C" int vs_select_probes(vs_ctx *c, const vs_index *ix, const uint8_t *queries, size_t nq, size_t nprobe,
                                uint32_t *probe_out, float *probe_sims_out) {
    if (!probe_out) return fail(VS_EINVAL, "null output");
    return search_host(c, ix, queries, nq, nprobe, 1, nullptr, nullptr, nullptr, probe_out, probe_sims_out, true);
}

extern "C" int vs_search_flat(vs_ctx *c, const vs_matrix *m, const uint64_t *d_doc_ids, const uint8_t *queries, size_t nq,
                              size_t k, uint64_t *ids_out, float *sims_out, int32_t *counts_out) {
    VS(need_dev());
    if (!c || !m) return fail(VS_EINVAL, "null argument");
    // A flat scan is an index with one list and no centroid stage.
    vs_index ix;
    ix.data = const_cast<vs_matrix *>(m);
    ix.centroids = nullptr;
    ix.doc_ids = const_cast<uint64_t *>(d_doc_ids);
    ix.n = m->n;
    ix.C = 1;
    return vs_search(c, &ix, queries, nq, 1, k, ids_out, sims_out, counts_out);
}

// ------------------------------------------------------------------------------------------------
// query batches as a tensor-core GEMM (gemm.cu)
static int gemm_search_core(vs_ctx *c, Arena &a, const vs_index *ix, const MatView &qv, size_t k, const SearchSetup &s,
                            const GemmPlan &pl, uint64_t *d_ids, float *d_sims, int32_t *d_counts, uint32_t *d_status,
                            uint64_t *stats) {
    const size_t nq = qv.n;
    GemmBufs gb;
    gemm_take(a.take<char>(gemm_scratch_bytes(pl, nq)), pl, nq, &gb);
    const MatView rows = ix->data->view();
    auto phase = [&](int i) -> cudaError_t {
        if (!stats) return cudaSuccess;
        if (!c->phase_ev[i]) {
            const cudaError_t e = cudaEventCreate(&c->phase_ev[i]);
            if (e != cudaSuccess) return e;
        }
        return cudaEventRecord(c->phase_ev[i], c->stream);
    };
    CU(phase(0));
    CU(gemm_enqueue_prepass(rows, qv, pl, gb, d_status, g_sm_count, c->stream, &c->launches));
    CU(phase(1));
    VS(prof_mark(c));
    CU(gemm_enqueue_filter(rows, qv, pl, gb, g_sm_count, c->stream, &c->launches));
    VS(prof_mark(c));
    CU(phase(2));
    VS(pinned_reserve(c, nq * 8 + 64));
    uint32_t *h_status = static_cast<uint32_t *>(c->pinned);
    unsigned int *h_cnt = reinterpret_cast<unsigned int *>(h_status + nq);
    std::vector<uint32_t> sel;
    CU(gemm_enqueue_select(rows, ix->doc_ids, ix->id_base, ix->doc_ids == nullptr, qv, pl, gb, (int)k, d_ids, d_sims, d_counts, d_status, c->d_fix_counter,
                           g_sm_count, c->stream, &c->launches));
    CU(cudaMemcpyAsync(h_status, d_status, nq * 4, cudaMemcpyDeviceToHost, c->stream));
    if (stats) CU(cudaMemcpyAsync(h_cnt, gb.cand_cnt, nq * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < nq; i++)
        if (h_status[i] & kStatusNeedMore) sel.push_back((uint32_t)i);
    CU(phase(3));
    if (stats) {
        CU(cudaEventSynchronize(c->phase_ev[3]));
        stats[0] = 0;
        for (size_t i = 0; i < nq; i++) stats[0] += h_cnt[i];
        stats[1] = sel.size();
        stats[2] = pl.tiles;
        stats[3] = pl.sample_tiles;
        for (int i = 0; i < 3; i++) {  // microseconds: pre-pass, filtering GEMM, candidate resolution
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, c->phase_ev[i], c->phase_ev[i + 1]));
            stats[4 + i] = (uint64_t)(ms * 1000.0f);
        }
    }
    if (sel.empty()) return VS_OK;
    // queries the filter could not answer (unusable header, fewer than k documents above the threshold, overflow):
    // the streaming scan, then its own literal-arithmetic resolve where needed
    CU(cudaMemcpyAsync(s.b.q_select, sel.data(), sel.size() * 4, cudaMemcpyHostToDevice, c->stream));
    VS(search_enqueue(c, ix, qv, sel.size(), s.b.q_select, s.npe, k, s.kpl1, s.kpl2, true, false, s.b, d_ids, d_sims, d_counts,
                      d_status, nullptr, false));
    CU(cudaMemcpyAsync(h_status, d_status, nq * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < nq; i++) h_status[i] &= (kStatusProbeAmbiguous | kStatusListAmbiguous);
    bool any = false;
    for (uint32_t i : sel) any |= h_status[i] != 0;
    if (any) {
        std::vector<uint32_t> hs(nq, 0u);
        for (uint32_t i : sel) hs[i] = h_status[i];
        VS(search_resolve_flagged(c, ix, qv, nq, hs.data(), s, k, d_ids, d_sims, d_counts, d_status, nullptr, false, nullptr));
    }
    return VS_OK;
}

static int gemm_check(vs_ctx *c, const vs_matrix *m, size_t nq, size_t d, size_t k) {
    VS(need_dev());
    if (!c || !m) return fail(VS_EINVAL, "null argument");
    if (nq == 0 || k == 0) return fail(VS_EINVAL, "nq == 0 or k == 0");
    if ((size_t)m->d != d) return fail(VS_EDIM, "vector/matrix column size does not match: %zu != %d", d, m->d);
    if (!gemm_supported(m->view(), nq))
        return fail(VS_ERANGE, "batched GEMM search needs d <= 1024 and at most %d queries per call", kMaxStageQueries);
    return VS_OK;
}

static int gemm_search_dev(vs_ctx *c, const vs_index *ix, const vs_matrix *queries, size_t k, uint64_t *d_ids, float *d_sims,
                           int32_t *d_counts, uint64_t *stats_out) {
    if (!queries || !d_ids || !d_sims || !d_counts) return fail(VS_EINVAL, "null argument");
    VS(gemm_check(c, ix->data, queries->n, (size_t)queries->d, k));
    const size_t nq = queries->n;
    const GemmPlan pl = gemm_plan(ix->data->view(), nq, k, ix->doc_ids == nullptr, 64, (uint32_t)g_sm_count, 4096);
    Arena a(c);
    SearchSetup s;
    VS(search_setup(c, a, ix, nq, ix->C, k, gemm_scratch_bytes(pl, nq) + Arena::pad(nq * 4) + 4096, &s));
    search_take(a, ix, nq, &s);
    uint32_t *d_status = a.take<uint32_t>(nq);
    return gemm_search_core(c, a, ix, queries->view(), k, s, pl, d_ids, d_sims, d_counts, d_status, stats_out);
}

static int gemm_search_host(vs_ctx *c, const vs_index *ix, const uint8_t *queries, size_t nq, size_t k, uint64_t *ids_out,
                            float *sims_out, int32_t *counts_out) {
    if (!queries || !ids_out || !sims_out || !counts_out) return fail(VS_EINVAL, "null argument");
    VS(gemm_check(c, ix->data, nq, (size_t)ix->data->d, k));
    const size_t row_bytes = 8 + (size_t)ix->data->d;
    const GemmPlan pl = gemm_plan(ix->data->view(), nq, k, ix->doc_ids == nullptr, 64, (uint32_t)g_sm_count, 4096);
    Arena a(c);
    SearchSetup s;
    const size_t out_bytes = Arena::pad(nq * k * 8) + Arena::pad(nq * k * 4) + 2 * Arena::pad(nq * 4) + 4096;
    VS(search_setup(c, a, ix, nq, ix->C, k, gemm_scratch_bytes(pl, nq) + temp_matrix_bytes(nq, row_bytes) + out_bytes, &s));
    MatView qv;
    VS(temp_matrix(c, a, queries, nq, row_bytes, &qv));
    search_take(a, ix, nq, &s);
    uint64_t *d_ids = a.take<uint64_t>(nq * k);
    float *d_sims = a.take<float>(nq * k);
    int32_t *d_counts = a.take<int32_t>(nq);
    uint32_t *d_status = a.take<uint32_t>(nq);
    VS(gemm_search_core(c, a, ix, qv, k, s, pl, d_ids, d_sims, d_counts, d_status, nullptr));
    CU(cudaMemcpyAsync(ids_out, d_ids, nq * k * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(sims_out, d_sims, nq * k * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(counts_out, d_counts, nq * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VS_OK;
}

// a matrix as an index with one list and no centroid stage
static vs_index flat_view(const vs_matrix *m, const uint64_t *d_doc_ids, uint64_t id_base) {
    vs_index ix;
    ix.data = const_cast<vs_matrix *>(m);
    ix.centroids = nullptr;
    ix.doc_ids = const_cast<uint64_t *>(d_doc_ids);
    ix.id_base = id_base;
    ix.n = m->n;
    ix.C = 1;
    return ix;
}

extern "C" int vs_search_batch_dev(vs_ctx *c, const vs_matrix *m, const uint64_t *d_doc_ids, uint64_t id_base,
                                   const vs_matrix *queries, size_t k, uint64_t *d_ids, float *d_sims, int32_t *d_counts,
                                   uint64_t *stats_out) {
    VS(need_dev());
    if (!c || !m) return fail(VS_EINVAL, "null argument");
    const vs_index ix = flat_view(m, d_doc_ids, id_base);
    return gemm_search_dev(c, &ix, queries, k, d_ids, d_sims, d_counts, stats_out);
}

extern "C" int vs_index_search_batch_dev(vs_ctx *c, const vs_index *ix, const vs_matrix *queries, size_t k, uint64_t *d_ids,
                                         float *d_sims, int32_t *d_counts, uint64_t *stats_out) {
    VS(search_check(c, ix));
    if (index_has_holes(ix)) return fail(VS_EINVAL, "the index has room to grow: its store is not one piece (see vs_index_append)");
    return gemm_search_dev(c, ix, queries, k, d_ids, d_sims, d_counts, stats_out);
}

extern "C" int vs_search_flat_gemm(vs_ctx *c, const vs_matrix *m, const uint64_t *d_doc_ids, const uint8_t *queries, size_t nq,
                                   size_t k, uint64_t *ids_out, float *sims_out, int32_t *counts_out) {
    VS(need_dev());
    if (!c || !m) return fail(VS_EINVAL, "null argument");
    const vs_index ix = flat_view(m, d_doc_ids, 0);
    return gemm_search_host(c, &ix, queries, nq, k, ids_out, sims_out, counts_out);
}

// ------------------------------------------------------------------------------------------------
// Nearest centroid through the tensor cores (BASELINE config 4: assign for large k).  The centroids are the store, the
// data rows go through as query batches, and every row gets its certified top-2 (float32 similarities that carry the
// reference's bits, ties ordered by centroid index).  float32 rounding is monotonic, so sims[0] > sims[1] in float32
// implies the same strict order of the float64 values compute/cosine.go:114 compares, and sims[0] > -1.0f implies the
// winner beats the -1.0 seed (cosine.go:101-102): such a row is done.  Every other row (float32 tie, seed not beaten,
// unusable header, too many candidates) goes to the literal-arithmetic kernel of argmax.cu.  Byte-identical duplicate
// centroids can never win the strict '>' and are left out of the store.
__global__ void argmax_from_top2_kernel(const uint64_t *ids, const float *sims, const int32_t *counts, const uint32_t *status,
                                        uint32_t nb, uint32_t row0, int k, int32_t *idx_out, float *sims_out, uint32_t *worklist,
                                        unsigned int *work_count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    bool certain = !(status[i] & kStatusNeedMore) && counts[i] >= k;
    float s0 = 0.f;
    if (certain) {
        s0 = sims[(size_t)i * k];
        certain = s0 > -1.0f;  // false for NaN
        if (certain && k >= 2) {
            const float s1 = sims[(size_t)i * k + 1];
            certain = (s0 > s1) || (s1 != s1);  // NaN ranks below every number
        }
    }
    if (certain) {
        idx_out[row0 + i] = (int32_t)ids[(size_t)i * k];
        if (sims_out) sims_out[row0 + i] = s0;
    } else {
        worklist[atomicAdd(work_count, 1u)] = row0 + i;
    }
}

__global__ void scatter_assign_kernel(const uint32_t *work, unsigned int cnt, const int32_t *w_idx, const float *w_sims,
                                      int32_t *idx_out, float *sims_out) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    idx_out[work[i]] = w_idx[i];
    if (sims_out) sims_out[work[i]] = w_sims[i];
}

static int aux_reserve(vs_ctx *c, size_t bytes) {
    if (bytes <= c->aux_cap) return VS_OK;
    CU(cudaStreamSynchronize(c->stream));
    if (c->aux) CU(cudaFree(c->aux));
    c->aux = nullptr;
    c->aux_cap = 0;
    const cudaError_t e = cudaMalloc(&c->aux, bytes);
    if (e != cudaSuccess) return fail(VS_ENOMEM, "cudaMalloc(%zu) for the assignment scratch: %s", bytes, cudaGetErrorString(e));
    c->aux_cap = bytes;
    return VS_OK;
}

constexpr size_t kArgmaxGemmBatch = 32768;  // data rows per GEMM launch

static int argmax_gemm_dev(vs_ctx *c, const MatView &cent, const MatView &data, int32_t *d_idx, float *d_sims, bool *done) {
    *done = false;
    const size_t n = data.n, M = cent.n;
    const size_t nb_max = n < kArgmaxGemmBatch ? n : kArgmaxGemmBatch;
    const int k = 2;
    // plan for the largest batch; the store may shrink below when duplicates are dropped (only fewer tiles)
    GemmPlan pl = gemm_plan(cent, nb_max, k, true, 16, 1, 1024);
    auto pad = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t store_bytes = pad(M * (size_t)cent.d_pad) + pad(M * 8) * 2 + pad(M * 8);
    const size_t bytes = pad(M * 4) * 2 + store_bytes + gemm_scratch_bytes(pl, nb_max) + pad(nb_max * k * 8) + pad(nb_max * k * 4) +
                         pad(nb_max * 4) * 2 + pad(n * 4) + 4096;
    VS(aux_reserve(c, bytes));
    char *base = static_cast<char *>(c->aux);
    size_t off = 0;
    auto take = [&](size_t b) {
        char *p = base + off;
        off += pad(b);
        return p;
    };
    uint32_t *d_canon = reinterpret_cast<uint32_t *>(take(M * 4));
    uint32_t *d_keep = reinterpret_cast<uint32_t *>(take(M * 4));
    uint8_t *s_codes = reinterpret_cast<uint8_t *>(take(M * (size_t)cent.d_pad));
    float2 *s_hdr = reinterpret_cast<float2 *>(take(M * 8));
    uint2 *s_sums = reinterpret_cast<uint2 *>(take(M * 8));
    uint64_t *s_ids = reinterpret_cast<uint64_t *>(take(M * 8));
    char *g_scratch = take(gemm_scratch_bytes(pl, nb_max));
    uint64_t *t_ids = reinterpret_cast<uint64_t *>(take(nb_max * k * 8));
    float *t_sims = reinterpret_cast<float *>(take(nb_max * k * 4));
    int32_t *t_counts = reinterpret_cast<int32_t *>(take(nb_max * 4));
    uint32_t *t_status = reinterpret_cast<uint32_t *>(take(nb_max * 4));
    uint32_t *d_work = reinterpret_cast<uint32_t *>(take(n * 4));
    unsigned int *d_wcount = reinterpret_cast<unsigned int *>(take(64));

    // duplicates out (cosine.go:114: a later byte-identical centroid never wins)
    LAUNCH(c, launch_canonical_rows(cent, d_canon, c->stream));
    std::vector<uint32_t> canon(M), keep;
    CU(cudaMemcpyAsync(canon.data(), d_canon, M * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemsetAsync(d_wcount, 0, sizeof(unsigned int), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    keep.reserve(M);
    for (size_t j = 0; j < M; j++)
        if (canon[j] == (uint32_t)j) keep.push_back((uint32_t)j);
    if (keep.size() < 2) return VS_OK;  // a single distinct centroid: the scan form answers
    MatView store = cent;
    const uint64_t *store_ids = nullptr;
    if (keep.size() != M) {
        CU(cudaMemcpyAsync(d_keep, keep.data(), keep.size() * 4, cudaMemcpyHostToDevice, c->stream));
        LAUNCH(c, launch_gather_rows(cent, d_keep, keep.size(), s_codes, s_hdr, s_sums, nullptr, 0, s_ids, c->stream));
        store = MatView{s_codes, s_hdr, s_sums, keep.size(), cent.d, cent.d_pad};
        store_ids = s_ids;
    }
    VS(pinned_reserve(c, 64));
    unsigned int *h_count = static_cast<unsigned int *>(c->pinned);
    for (size_t r0 = 0; r0 < n; r0 += kArgmaxGemmBatch) {
        const size_t nb = n - r0 < kArgmaxGemmBatch ? n - r0 : kArgmaxGemmBatch;
        const MatView q{data.codes + r0 * (size_t)data.d_pad, data.hdr + r0, data.sums + r0, nb, data.d, data.d_pad};
        GemmPlan plb = gemm_plan(store, nb, k, true, 16, 1, 1024);
        if (plb.cand_per_q > pl.cand_per_q || plb.G > pl.G || plb.nq_pad > pl.nq_pad) return fail(VS_EINVAL, "assignment plan grew");
        GemmBufs gb;
        gemm_take(g_scratch, plb, nb, &gb);
        CU(gemm_enqueue_prepass(store, q, plb, gb, t_status, g_sm_count, c->stream, &c->launches));
        VS(prof_mark(c));
        CU(gemm_enqueue_filter(store, q, plb, gb, g_sm_count, c->stream, &c->launches));
        VS(prof_mark(c));
        CU(gemm_enqueue_select(store, store_ids, 0, true, q, plb, gb, k, t_ids, t_sims, t_counts, t_status, c->d_fix_counter, g_sm_count,
                               c->stream, &c->launches));
        argmax_from_top2_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, c->stream>>>(t_ids, t_sims, t_counts, t_status, (uint32_t)nb,
                                                                                    (uint32_t)r0, k, d_idx, d_sims, d_work, d_wcount);
        c->launches++;
    }
    CU(cudaMemcpyAsync(h_count, d_wcount, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const unsigned int cnt = *h_count;
    if (cnt > 0) {
        // Rows the float32 top-2 could not decide: the scan form on just those rows (float64 intervals separate almost
        // all of them), then literal arithmetic for what is left (true ties).  Rare, so plain allocations.
        char *tmp = nullptr;
        const size_t wb = pad((size_t)cnt * data.d_pad) + pad((size_t)cnt * 8) * 2 + pad((size_t)cnt * 4) * 3 + 256;
        CU(cudaMalloc(&tmp, wb));
        size_t o = 0;
        auto sub = [&](size_t bts) {
            char *p = tmp + o;
            o += pad(bts);
            return p;
        };
        uint8_t *w_codes = reinterpret_cast<uint8_t *>(sub((size_t)cnt * data.d_pad));
        float2 *w_hdr = reinterpret_cast<float2 *>(sub((size_t)cnt * 8));
        uint2 *w_sums = reinterpret_cast<uint2 *>(sub((size_t)cnt * 8));
        int32_t *w_idx = reinterpret_cast<int32_t *>(sub((size_t)cnt * 4));
        float *w_sims = reinterpret_cast<float *>(sub((size_t)cnt * 4));
        uint32_t *w_work = reinterpret_cast<uint32_t *>(sub((size_t)cnt * 4));
        unsigned int *w_count = reinterpret_cast<unsigned int *>(sub(64));
        const MatView wdata{w_codes, w_hdr, w_sums, cnt, data.d, data.d_pad};
        cudaError_t e = launch_gather_rows(data, d_work, cnt, w_codes, w_hdr, w_sums, nullptr, 0, nullptr, c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(w_count, 0, sizeof(unsigned int), c->stream);
        if (e == cudaSuccess) e = launch_argmax(cent, wdata, d_canon, w_idx, d_sims ? w_sims : nullptr, w_work, w_count, g_sm_count, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_count, w_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        c->launches += 2;
        const unsigned int cnt2 = e == cudaSuccess ? *h_count : 0;
        if (e == cudaSuccess && cnt2 > 0) {
            double *d_cn = nullptr;
            e = cudaMalloc(&d_cn, M * (size_t)cent.d * sizeof(double));
            if (e == cudaSuccess) e = launch_query_normalize(cent, d_cn, c->stream);
            if (e == cudaSuccess) e = launch_argmax_fix(cent, wdata, d_cn, w_idx, d_sims ? w_sims : nullptr, w_work, w_count, g_sm_count, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (d_cn) cudaFree(d_cn);
            c->launches += 2;
            c->slowpath += cnt2;
        }
        if (e == cudaSuccess) {
            scatter_assign_kernel<<<(cnt + 255) / 256, 256, 0, c->stream>>>(d_work, cnt, w_idx, w_sims, d_idx, d_sims);
            c->launches++;
            e = cudaStreamSynchronize(c->stream);
        }
        cudaFree(tmp);
        if (e != cudaSuccess) return fail(VS_ECUDA, "assignment fallback: %s", cudaGetErrorString(e));
    }
    *done = true;
    return VS_OK;
}

extern "C" int vs_debug_set_argmax_gemm_min(size_t min_centroids) {
    g_argmax_gemm_min_centroids = min_centroids < 2 ? 2 : min_centroids;
    return VS_OK;
}

extern "C" int vs_topk_merge_dev(vs_ctx *c, const uint64_t *d_ids_in, const float *d_sims_in, const int32_t *d_counts_in,
                                 size_t G, size_t nq, size_t k, uint64_t *d_ids_out, float *d_sims_out, int32_t *d_counts_out) {
    VS(need_dev());
    if (!c) return fail(VS_EINVAL, "ctx is null");
    if (k > 128) return fail(VS_ERANGE, "k=%zu > 128", k);
    if (nq == 0) return VS_OK;
    LAUNCH(c, launch_topk_merge(d_ids_in, d_sims_in, d_counts_in, 0, (int)G, (int)nq, (int)k, d_ids_out, d_sims_out, d_counts_out,
                                c->stream));
    return VS_OK;
}

extern "C" int vs_topk_merge_packed_dev(vs_ctx *c, const void *d_packed, size_t rank_stride_bytes, size_t ids_off, size_t sims_off,
                                        size_t counts_off, size_t G, size_t nq, size_t k, uint64_t *d_ids_out, float *d_sims_out,
                                        int32_t *d_counts_out) {
    VS(need_dev());
    if (!c || !d_packed) return fail(VS_EINVAL, "null argument");
    if (k > 128) return fail(VS_ERANGE, "k=%zu > 128", k);
    if ((sims_off | counts_off) & 3) return fail(VS_EINVAL, "sims_off / counts_off must be 4-byte aligned");
    if (((size_t)(uintptr_t)d_packed | ids_off | rank_stride_bytes) & 7)
        return fail(VS_EINVAL, "the packed buffer, ids_off and rank_stride_bytes must be 8-byte aligned (uint64 ids)");
    if (nq == 0) return VS_OK;
    const char *b = static_cast<const char *>(d_packed);
    LAUNCH(c, launch_topk_merge(reinterpret_cast<const uint64_t *>(b + ids_off), reinterpret_cast<const float *>(b + sims_off),
                                reinterpret_cast<const int32_t *>(b + counts_off), rank_stride_bytes, (int)G, (int)nq, (int)k,
                                d_ids_out, d_sims_out, d_counts_out, c->stream));
    return VS_OK;
}

// ------------------------------------------------------------------------------------------------
// k-means step / recenter
extern "C" int vs_kmeans_step(vs_ctx *c, const vs_matrix *data, const uint8_t *centroids, size_t k, float *means,
                              int64_t *assign_out, int64_t *counts_out, uint8_t *new_centroids_out, int *converged_out) {
    VS(need_dev());
    if (!c || !data || !centroids || !means || !counts_out || !new_centroids_out) return fail(VS_EINVAL, "null argument");
    if (k == 0) return fail(VS_EEMPTY, "matrix rows are empty");
    const size_t n = data->n, d = data->d, rb = 8 + d;
    Arena a(c);
    VS(a.reserve(temp_matrix_bytes(k, rb) + argmax_bytes(k, n, d) + 4 * Arena::pad(n * 4) + Arena::pad((k + 1) * 4) +
                 Arena::pad(k * d * 4) + Arena::pad(k * 8) + Arena::pad(k * rb) + 8192));
    MatView cv;
    VS(temp_matrix(c, a, centroids, k, rb, &cv));
    int32_t *d_assign = a.take<int32_t>(n);
    uint32_t *d_order = a.take<uint32_t>(n);
    uint32_t *d_sorted = a.take<uint32_t>(n);
    uint32_t *d_segoff = a.take<uint32_t>(k + 1);
    float *d_means = a.take<float>(k * d);
    int64_t *d_counts = a.take<int64_t>(k);
    uint8_t *d_newc = a.take<uint8_t>(k * rb);
    // k_means.go:73-77: nearest centroid of every row
    VS(argmax_dev(c, a, cv, data->view(), d_assign, nullptr));
    CU(cudaMemcpyAsync(d_means, means, k * d * 4, cudaMemcpyHostToDevice, c->stream));
    // k_means.go:80-96 accumulate + mean, :99 requantize
    if (k <= kKMeansScanCentroids) {
        LAUNCH(c, launch_kmeans_accumulate_scan(data->view(), d_assign, (int)k, d_means, d_counts, c->stream));
    } else {
        // member lists in ascending row order (stable sort by centroid index)
        VS(sort_rows_by_key(c, reinterpret_cast<const uint32_t *>(d_assign), n, bits_for(k), d_order, d_sorted));
        lower_bound_kernel<<<(unsigned)((k + 1 + 255) / 256), 256, 0, c->stream>>>(d_sorted, n, nullptr, d_segoff, k);
        c->launches++;
        LAUNCH(c, launch_kmeans_accumulate(data->view(), d_order, d_segoff, (int)k, d_means, d_counts, c->stream));
    }
    LAUNCH(c, launch_quantize_f32(d_means, k, (int)d, d_newc, c->stream));
    std::vector<int32_t> tmp;
    if (assign_out) {
        tmp.resize(n);
        CU(cudaMemcpyAsync(tmp.data(), d_assign, n * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaMemcpyAsync(means, d_means, k * d * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(counts_out, d_counts, k * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(new_centroids_out, d_newc, k * rb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (assign_out)
        for (size_t i = 0; i < n; i++) assign_out[i] = tmp[i];
    if (converged_out) {  // k_means.go:102-108: code bytes [8:] unchanged for every centroid
        int conv = 1;
        for (size_t j = 0; j < k && conv; j++)
            if (memcmp(new_centroids_out + j * rb + 8, centroids + j * rb + 8, d) != 0) conv = 0;
        *converged_out = conv;
    }
    return VS_OK;
}

// The whole kMeans of dnc/k_means.go:19-212 with every iteration on the device.  The random superset draw (:35-44) is the
// caller's (superset_rows); the superset phase iterates ks centroids until the code bytes stop changing (:67-117), the set
// phase keeps the first k (the sort at :132 compares counts that :111-113 already zeroed; Go's pdqsort leaves an
// all-equal slice untouched) together with their float32 means (:153-154) and iterates again (:157-207).
__global__ void codes_differ_kernel(const uint8_t *a, const uint8_t *b, size_t rows, int d, int d_pad, int *flag) {
    const size_t total = rows * (size_t)d;
    int diff = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total && !diff; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / d, j = i % d;
        diff = a[r * d_pad + j] != b[r * d_pad + j];
    }
    if (diff) *flag = 1;
}

extern "C" int vs_kmeans(vs_ctx *c, const vs_matrix *data, size_t k, const uint64_t *superset_rows, size_t ks, size_t iter_limit,
                         uint8_t *centroids_out, int64_t *stats_out) {
    VS(need_dev());
    if (!c || !data || !superset_rows || !centroids_out) return fail(VS_EINVAL, "null argument");
    const size_t n = data->n, d = data->d, rb = 8 + d, d_pad = data->d_pad;
    if (k == 0) return fail(VS_EEMPTY, "matrix rows are empty");
    if (ks < k || ks > n) return fail(VS_EINVAL, "superset of %zu rows for k=%zu, n=%zu", ks, k, n);
    if (n > 0xFFFFFFF0ull) return fail(VS_ERANGE, "n=%zu rows", n);
    std::vector<uint32_t> rows32(ks);
    for (size_t i = 0; i < ks; i++) {
        if (superset_rows[i] >= n) return fail(VS_EINVAL, "superset row %llu >= n", (unsigned long long)superset_rows[i]);
        rows32[i] = (uint32_t)superset_rows[i];
    }
    // few centroids over a store small enough to copy once per iteration (the reference's sampled shapes)
    const bool use_ring = ks <= kKMeansScanCentroids && kmeans_ring_supported((int)d_pad) && n * d_pad <= (size_t(512) << 20);
    Arena a(c);
    const size_t soa = Arena::pad(ks * d_pad) + 2 * Arena::pad(ks * 8);
    VS(a.reserve(2 * soa + Arena::pad(ks * 4) + argmax_bytes(ks, n, d) + 3 * Arena::pad(n * 4) + Arena::pad((ks + 1) * 4) +
                 2 * Arena::pad(ks * d * 4) + Arena::pad(ks * 8) + Arena::pad(k * rb) + sort_rows_ws_bytes(n) +
                 (use_ring ? kmeans_ring_scratch_bytes(n, (int)ks, (int)d_pad) : 0) + soa + 8192));
    struct Soa {
        uint8_t *codes;
        float2 *hdr;
        uint2 *sums;
    } buf[3];
    for (int i = 0; i < 3; i++) {
        buf[i].codes = a.take<uint8_t>(ks * d_pad);
        buf[i].hdr = a.take<float2>(ks);
        buf[i].sums = a.take<uint2>(ks);
    }
    uint32_t *d_rows = a.take<uint32_t>(ks);
    int32_t *d_assign = a.take<int32_t>(n);
    uint32_t *d_order = a.take<uint32_t>(n);
    uint32_t *d_sorted = a.take<uint32_t>(n);
    uint32_t *d_segoff = a.take<uint32_t>(ks + 1);
    float *d_means2[2] = {a.take<float>(ks * d), a.take<float>(ks * d)};
    int64_t *d_counts = a.take<int64_t>(ks);
    uint8_t *d_out = a.take<uint8_t>(k * rb);
    int *d_flag = a.take<int>(16);
    const size_t sort_ws_bytes = sort_rows_ws_bytes(n);
    char *d_sort_ws = a.take<char>(sort_ws_bytes);
    char *d_ring = use_ring ? a.take<char>(kmeans_ring_scratch_bytes(n, (int)ks, (int)d_pad)) : nullptr;
    CU(cudaMemcpyAsync(d_rows, rows32.data(), ks * 4, cudaMemcpyHostToDevice, c->stream));
    for (int i = 0; i < 3; i++) CU(cudaMemsetAsync(buf[i].codes, 0, ks * d_pad, c->stream));  // the padding columns stay zero
    LAUNCH(c, launch_gather_rows(data->view(), d_rows, ks, buf[0].codes, buf[0].hdr, buf[0].sums, nullptr, 0, nullptr, c->stream));
    CU(cudaMemsetAsync(d_means2[0], 0, ks * d * 4, c->stream));  // k_means.go:60-65
    CU(cudaMemsetAsync(d_means2[1], 0, ks * d * 4, c->stream));
    if (!c->h_flags) CU(cudaMallocHost(&c->h_flags, 64));  // not c->pinned: the assignment inside an iteration uses that
    volatile int *h_flag = c->h_flags;
    // per-slot events: start, assign done, iteration done (the flag of that iteration is on the host)
    cudaEvent_t ev[2][3];
    for (int sl = 0; sl < 2; sl++)
        for (int j = 0; j < 3; j++) CU(cudaEventCreate(&ev[sl][j]));
    int64_t iters[2] = {0, 0};
    double ms_assign = 0, ms_update = 0;
    // One Lloyd iteration, fully asynchronous: in -> out, means mi -> mo, convergence flag to h_flag[slot].
    auto enqueue = [&](const Soa &in, const Soa &out, const float *mi, float *mo, size_t m, int slot) -> int {
        const MatView cv{in.codes, in.hdr, in.sums, m, (int)d, (int)d_pad};
        CU(cudaEventRecord(ev[slot][0], c->stream));
        const size_t mark = a.off;
        VS(argmax_dev(c, a, cv, data->view(), d_assign, nullptr));  // :73-77
        a.off = mark;
        CU(cudaEventRecord(ev[slot][1], c->stream));
        if (use_ring) {  // few centroids: member rows made contiguous, then streamed through a shared-memory ring
            LAUNCH(c, launch_kmeans_accumulate_ring(data->view(), d_assign, (int)m, mo, d_counts, mi, d_ring, c->stream));  // :80-96
            c->launches += 3;
        } else if (m <= kKMeansScanCentroids) {  // (rows too many to copy) members found by walking the assignment
            LAUNCH(c, launch_kmeans_accumulate_scan(data->view(), d_assign, (int)m, mo, d_counts, c->stream, mi));
        } else {
            VS(sort_rows_by_key(c, reinterpret_cast<const uint32_t *>(d_assign), n, bits_for(m), d_order, d_sorted, d_sort_ws,
                                sort_ws_bytes));
            lower_bound_kernel<<<(unsigned)((m + 1 + 255) / 256), 256, 0, c->stream>>>(d_sorted, n, nullptr, d_segoff, m);
            c->launches++;
            LAUNCH(c, launch_kmeans_accumulate(data->view(), d_order, d_segoff, (int)m, mo, d_counts, c->stream, mi));  // :80-96
        }
        LAUNCH(c, launch_quantize_f32_soa(mo, m, (int)d, out.codes, (int)d_pad, out.hdr, out.sums, c->stream));     // :99
        CU(cudaMemsetAsync(d_flag + slot, 0, sizeof(int), c->stream));
        codes_differ_kernel<<<g_sm_count * 4, 256, 0, c->stream>>>(in.codes, out.codes, m, (int)d, (int)d_pad, d_flag + slot);  // :102-108
        c->launches++;
        CU(cudaMemcpyAsync(const_cast<int *>(h_flag) + slot, d_flag + slot, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaEventRecord(ev[slot][2], c->stream));
        return VS_OK;
    };
    // The loop of k_means.go:67-117 / :157-207.  Whether iteration i converged is known only on the host; iteration i+1
    // is enqueued before that answer arrives (its inputs are iteration i's outputs either way) so the device never waits
    // for the round trip.  Three centroid buffers and two mean buffers keep iteration i's results intact when the
    // speculative iteration turns out to be one too many: it is simply ignored.
    Soa in = buf[0], out = buf[1], spare = buf[2];
    float *mi = d_means2[0], *mo = d_means2[1];
    int rc = VS_OK;
    for (int phase = 0; phase < 2 && rc == VS_OK; phase++) {
        const size_t m = phase == 0 ? ks : k;
        if (iter_limit == 0) break;
        int slot = 0;
        rc = enqueue(in, out, mi, mo, m, slot);
        for (size_t it = 0; rc == VS_OK; it++) {
            const bool can_next = it + 1 < iter_limit;
            if (can_next) rc = enqueue(out, spare, mo, mi, m, slot ^ 1);
            if (rc != VS_OK) break;
            if (cudaEventSynchronize(ev[slot][2]) != cudaSuccess) {
                rc = fail(VS_ECUDA, "k-means: event sync");
                break;
            }
            const bool conv = h_flag[slot] == 0;
            float t0 = 0.f, t1 = 0.f;
            cudaEventElapsedTime(&t0, ev[slot][0], ev[slot][1]);
            cudaEventElapsedTime(&t1, ev[slot][1], ev[slot][2]);
            ms_assign += t0;
            ms_update += t1;
            iters[phase]++;
            // this iteration's output becomes the next input (or the result)
            const Soa done_in = in;
            in = out;
            out = spare;
            spare = done_in;
            float *t = mi;
            mi = mo;
            mo = t;
            slot ^= 1;
            if (conv || !can_next) break;
        }
        // `in` / `mi` now hold the phase's result; a speculative iteration may still be writing `out` / `mo`, in stream
        // order before anything enqueued below
    }
    for (int sl = 0; sl < 2; sl++)
        for (int j = 0; j < 3; j++) cudaEventDestroy(ev[sl][j]);
    VS(rc);
    const Soa cur = in;
    const MatView fin{cur.codes, cur.hdr, cur.sums, k, (int)d, (int)d_pad};
    LAUNCH(c, launch_export(fin, 0, k, d_out, c->stream));
    CU(cudaMemcpyAsync(centroids_out, d_out, k * rb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (stats_out) {
        stats_out[0] = iters[0];
        stats_out[1] = iters[1];
        stats_out[2] = (int64_t)(ms_assign * 1000.0);
        stats_out[3] = (int64_t)(ms_update * 1000.0);
    }
    return VS_OK;
}

// ---- one Lloyd iteration over a store cut into contiguous row blocks (one block per GPU) --------------------------
// Every rank assigns its own rows (vs_argmax_MxN_dev; k_means.go:73-77) at the same time.  vs_kmeans_accumulate_dev then
// continues the per-centroid float32 sums and counts from the values the previous block left in d_sums / d_counts
// (zeros for the first block).  Blocks are accumulated in row order, so the additions run in the reference's order
// (k_means.go:80-86) and the result is the same bits as on one device; the running sums travel from rank to rank
// (NCCL send/recv, shard.py).
extern "C" int vs_kmeans_accumulate_dev(vs_ctx *c, const vs_matrix *data, size_t k, const int32_t *d_assign, float *d_sums,
                                        int64_t *d_counts) {
    VS(need_dev());
    if (!c || !data || !d_assign || !d_sums || !d_counts) return fail(VS_EINVAL, "null argument");
    if (k == 0) return fail(VS_EEMPTY, "matrix rows are empty");
    const size_t n = data->n;
    if (n > 0xFFFFFFF0ull) return fail(VS_ERANGE, "n=%zu rows", n);
    Arena a(c);
    VS(a.reserve(2 * Arena::pad(n * 4) + Arena::pad((k + 1) * 4) + sort_rows_ws_bytes(n) + 4096));
    uint32_t *d_order = a.take<uint32_t>(n);
    uint32_t *d_sorted = a.take<uint32_t>(n);
    uint32_t *d_segoff = a.take<uint32_t>(k + 1);
    const size_t ws_bytes = sort_rows_ws_bytes(n);
    char *d_ws = a.take<char>(ws_bytes);
    VS(sort_rows_by_key(c, reinterpret_cast<const uint32_t *>(d_assign), n, bits_for(k), d_order, d_sorted, d_ws, ws_bytes));
    lower_bound_kernel<<<(unsigned)((k + 1 + 255) / 256), 256, 0, c->stream>>>(d_sorted, n, nullptr, d_segoff, k);
    c->launches++;
    LAUNCH(c, launch_kmeans_accumulate_relay(data->view(), d_order, d_segoff, (int)k, d_sums, d_counts, c->stream));
    return VS_OK;
}

// vs_kmeans_finish_dev (after the last block): means = sums / float32(count) where a cluster has members, else the
// previous mean (k_means.go:89-96); new centroids = QuantizeMatrixFloat32(means) (:99) as a new device matrix;
// *converged_out = every centroid's code bytes unchanged (:102-108).  d_means: [k][D] float32 state, in/out.
extern "C" int vs_kmeans_finish_dev(vs_ctx *c, const vs_matrix *centroids, const float *d_sums, const int64_t *d_counts,
                                    float *d_means, vs_matrix **new_centroids_out, int *converged_out) {
    VS(need_dev());
    if (!c || !centroids || !d_sums || !d_counts || !d_means || !new_centroids_out) return fail(VS_EINVAL, "null argument");
    const size_t k = centroids->n, d = centroids->d;
    vs_matrix *m = nullptr;
    VS(matrix_alloc(k, d, &m));
    Arena a(c);
    int rc = a.reserve(1024);
    int *d_flag = rc == VS_OK ? a.take<int>(16) : nullptr;
    cudaError_t e = rc == VS_OK ? launch_kmeans_finalize(d_sums, d_counts, k, (int)d, d_means, c->stream) : cudaSuccess;
    if (rc == VS_OK && e == cudaSuccess) e = cudaMemsetAsync(m->codes, 0, k * (size_t)m->d_pad, c->stream);
    if (rc == VS_OK && e == cudaSuccess) e = launch_quantize_f32_soa(d_means, k, (int)d, m->codes, m->d_pad, m->hdr, m->sums, c->stream);
    if (rc == VS_OK && e == cudaSuccess) e = cudaMemsetAsync(d_flag, 0, sizeof(int), c->stream);
    if (rc == VS_OK && e == cudaSuccess) {
        codes_differ_kernel<<<g_sm_count * 4, 256, 0, c->stream>>>(centroids->codes, m->codes, k, (int)d, m->d_pad, d_flag);
        e = cudaGetLastError();
    }
    int h_flag = 0;
    if (rc == VS_OK && e == cudaSuccess) e = cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (rc == VS_OK && e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (rc == VS_OK && e != cudaSuccess) rc = fail(VS_ECUDA, "k-means finish: %s", cudaGetErrorString(e));
    if (rc != VS_OK) {
        vs_matrix_release(m);
        return rc;
    }
    c->launches += 3;
    if (converged_out) *converged_out = h_flag == 0;
    *new_centroids_out = m;
    return VS_OK;
}

// ---- device-side pieces of the divide-and-conquer centroid build (dnc/dnc.go:300-400, 402-456) --------------------
// vs_matrix_gather: a new matrix made of the given rows of src, in the given order (sample(), dnc/sampling.go:12-74).
extern "C" int vs_matrix_gather(vs_ctx *c, const vs_matrix *src, const uint64_t *rows, size_t n, vs_matrix **out) {
    VS(need_dev());
    if (!c || !src || !rows || !out) return fail(VS_EINVAL, "null argument");
    if (n == 0) return fail(VS_EEMPTY, "matrix rows are empty");
    std::vector<uint32_t> r32(n);
    for (size_t i = 0; i < n; i++) {
        if (rows[i] >= src->n) return fail(VS_EINVAL, "row %llu >= %zu", (unsigned long long)rows[i], src->n);
        r32[i] = (uint32_t)rows[i];
    }
    vs_matrix *m = nullptr;
    VS(matrix_alloc(n, src->d, &m));
    Arena a(c);
    int rc = a.reserve(Arena::pad(n * 4) + 1024);
    if (rc == VS_OK) {
        uint32_t *d_rows = a.take<uint32_t>(n);
        cudaError_t e = cudaMemcpyAsync(d_rows, r32.data(), n * 4, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = launch_gather_rows(src->view(), d_rows, n, m->codes, m->hdr, m->sums, nullptr, 0, nullptr, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // r32 is read by the copy
        if (e != cudaSuccess) rc = fail(VS_ECUDA, "gather: %s", cudaGetErrorString(e));
        c->launches++;
    }
    if (rc != VS_OK) {
        vs_matrix_release(m);
        return rc;
    }
    *out = m;
    return VS_OK;
}

// vs_matrix_split_dev: the split loop of divideNconquer (dnc.go:363-389) once the rows have their nearest-centroid
// index: child j receives the rows with d_assign[row] == j in their original order (a stable partition; the reference
// appends rows to the child files in read order).  children_out[j] = null for a child without rows.
extern "C" int vs_matrix_split_dev(vs_ctx *c, const vs_matrix *src, const int32_t *d_assign, size_t k, vs_matrix **children_out,
                                   uint64_t *counts_out) {
    VS(need_dev());
    if (!c || !src || !d_assign || !children_out || !counts_out) return fail(VS_EINVAL, "null argument");
    if (k == 0) return fail(VS_EINVAL, "k == 0");
    const size_t n = src->n;
    for (size_t j = 0; j < k; j++) children_out[j] = nullptr;
    Arena a(c);
    VS(a.reserve(2 * Arena::pad(n * 4) + Arena::pad((k + 1) * 4) + sort_rows_ws_bytes(n) + 4096));
    uint32_t *d_order = a.take<uint32_t>(n);
    uint32_t *d_sorted = a.take<uint32_t>(n);
    uint32_t *d_segoff = a.take<uint32_t>(k + 1);
    const size_t ws_bytes = sort_rows_ws_bytes(n);
    char *d_ws = a.take<char>(ws_bytes);
    VS(sort_rows_by_key(c, reinterpret_cast<const uint32_t *>(d_assign), n, bits_for(k), d_order, d_sorted, d_ws, ws_bytes));
    lower_bound_kernel<<<(unsigned)((k + 1 + 255) / 256), 256, 0, c->stream>>>(d_sorted, n, nullptr, d_segoff, k);
    c->launches++;
    std::vector<uint32_t> off(k + 1);
    CU(cudaMemcpyAsync(off.data(), d_segoff, (k + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    int rc = VS_OK;
    for (size_t j = 0; j < k && rc == VS_OK; j++) {
        const size_t cnt = off[j + 1] - off[j];
        counts_out[j] = cnt;
        if (cnt == 0) continue;
        vs_matrix *m = nullptr;
        rc = matrix_alloc(cnt, src->d, &m);
        if (rc != VS_OK) break;
        const cudaError_t e = launch_gather_rows(src->view(), d_order + off[j], cnt, m->codes, m->hdr, m->sums, nullptr, 0, nullptr,
                                                 c->stream);
        if (e != cudaSuccess) {
            vs_matrix_release(m);
            rc = fail(VS_ECUDA, "split: %s", cudaGetErrorString(e));
            break;
        }
        c->launches++;
        children_out[j] = m;
    }
    if (rc == VS_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(VS_ECUDA, "split: stream sync");
    if (rc != VS_OK)
        for (size_t j = 0; j < k; j++)
            if (children_out[j]) {
                vs_matrix_release(children_out[j]);
                children_out[j] = nullptr;
            }
    return rc;
}

// vs_recenter_clusters_dev: recenterDbCentroid (dnc.go:402-456) for all k clusters of an assignment at once: the float64
// mean of each cluster's rows in row order, quantized with QuantizeVectorFloat64.  centroids_out: k rows (host).
extern "C" int vs_recenter_clusters_dev(vs_ctx *c, const vs_matrix *data, const int32_t *d_assign, size_t k, uint8_t *centroids_out,
                                        int64_t *counts_out) {
    VS(need_dev());
    if (!c || !data || !d_assign || !centroids_out) return fail(VS_EINVAL, "null argument");
    if (k == 0) return fail(VS_EINVAL, "k == 0");
    const size_t n = data->n, d = data->d, rb = 8 + d;
    Arena a(c);
    VS(a.reserve(2 * Arena::pad(n * 4) + Arena::pad((k + 1) * 4) + sort_rows_ws_bytes(n) + Arena::pad(k * d * 8) + Arena::pad(k * 8) +
                 Arena::pad(k * rb) + 4096));
    uint32_t *d_order = a.take<uint32_t>(n);
    uint32_t *d_sorted = a.take<uint32_t>(n);
    uint32_t *d_segoff = a.take<uint32_t>(k + 1);
    const size_t ws_bytes = sort_rows_ws_bytes(n);
    char *d_ws = a.take<char>(ws_bytes);
    double *d_means = a.take<double>(k * d);
    int64_t *d_counts = a.take<int64_t>(k);
    uint8_t *d_rows = a.take<uint8_t>(k * rb);
    VS(sort_rows_by_key(c, reinterpret_cast<const uint32_t *>(d_assign), n, bits_for(k), d_order, d_sorted, d_ws, ws_bytes));
    lower_bound_kernel<<<(unsigned)((k + 1 + 255) / 256), 256, 0, c->stream>>>(d_sorted, n, nullptr, d_segoff, k);
    c->launches++;
    LAUNCH(c, launch_recenter_clusters(data->view(), d_order, d_segoff, (int)k, d_means, d_counts, c->stream));
    LAUNCH(c, launch_quantize_f64(d_means, k, (int)d, d_rows, c->stream));
    CU(cudaMemcpyAsync(centroids_out, d_rows, k * rb, cudaMemcpyDeviceToHost, c->stream));
    if (counts_out) CU(cudaMemcpyAsync(counts_out, d_counts, k * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VS_OK;
}

// Host-pointer forms for the D&C driver: the centroids arrive as packed rows, the assignment stays on the device.
// vs_matrix_split = the split loop of divideNconquer (dnc.go:363-389): nearest of the k centroids for every row of src
// (cosine.go:70-125), then the stable partition above.
extern "C" int vs_matrix_split(vs_ctx *c, const vs_matrix *src, const uint8_t *centroids, size_t k, vs_matrix **children_out,
                               uint64_t *counts_out) {
    VS(need_dev());
    if (!c || !src || !centroids || !children_out || !counts_out) return fail(VS_EINVAL, "null argument");
    if (k == 0) return fail(VS_EEMPTY, "matrix rows are empty");
    const size_t n = src->n, rb = 8 + (size_t)src->d;
    int32_t *d_assign = nullptr;
    CU(cudaMalloc(&d_assign, n * 4 + 4));
    int rc;
    {
        Arena a(c);
        rc = a.reserve(temp_matrix_bytes(k, rb) + argmax_bytes(k, n, rb - 8) + 4096);
        MatView cv;
        if (rc == VS_OK) rc = temp_matrix(c, a, centroids, k, rb, &cv);
        if (rc == VS_OK) rc = argmax_dev(c, a, cv, src->view(), d_assign, nullptr);
    }
    if (rc == VS_OK) rc = vs_matrix_split_dev(c, src, d_assign, k, children_out, counts_out);
    cudaFree(d_assign);
    return rc;
}

// The tail of KMeansDivideAndConquer (dnc.go:177-291): every row to its nearest of the k new centroids, then every
// centroid re-centred on its members (recenterDbCentroid).  assign_out (host int32, nullable) / d_assign_out (device
// int32, nullable) receive the assignment; centroids_out the k re-centred rows; counts_out the cluster sizes.
extern "C" int vs_reassign_recenter(vs_ctx *c, const vs_matrix *data, const uint8_t *centroids, size_t k, int32_t *assign_out,
                                    int32_t *d_assign_out, uint8_t *centroids_out, int64_t *counts_out) {
    VS(need_dev());
    if (!c || !data || !centroids || !centroids_out) return fail(VS_EINVAL, "null argument");
    if (k == 0) return fail(VS_EEMPTY, "matrix rows are empty");
    const size_t n = data->n, rb = 8 + (size_t)data->d;
    int32_t *d_assign = d_assign_out;
    if (!d_assign) CU(cudaMalloc(&d_assign, n * 4 + 4));
    int rc;
    {
        Arena a(c);
        rc = a.reserve(temp_matrix_bytes(k, rb) + argmax_bytes(k, n, rb - 8) + 4096);
        MatView cv;
        if (rc == VS_OK) rc = temp_matrix(c, a, centroids, k, rb, &cv);
        if (rc == VS_OK) rc = argmax_dev(c, a, cv, data->view(), d_assign, nullptr);
    }
    if (rc == VS_OK) rc = vs_recenter_clusters_dev(c, data, d_assign, k, centroids_out, counts_out);
    if (rc == VS_OK && assign_out) {
        const cudaError_t e = cudaMemcpy(assign_out, d_assign, n * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(VS_ECUDA, "assignment download: %s", cudaGetErrorString(e));
    }
    if (!d_assign_out) cudaFree(d_assign);
    return rc;
}

extern "C" int vs_recenter(vs_ctx *c, const vs_matrix *m, uint8_t *out_row) {
    VS(need_dev());
    if (!c || !m || !out_row) return fail(VS_EINVAL, "null argument");
    const size_t d = m->d, rb = 8 + d;
    Arena a(c);
    VS(a.reserve(Arena::pad(d * 8) + Arena::pad(rb) + 1024));
    double *d_mean = a.take<double>(d);
    uint8_t *d_row = a.take<uint8_t>(rb);
    LAUNCH(c, launch_recenter(m->view(), d_mean, c->stream));
    LAUNCH(c, launch_quantize_f64(d_mean, 1, (int)d, d_row, c->stream));
    CU(cudaMemcpyAsync(out_row, d_row, rb, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return VS_OK;
}
