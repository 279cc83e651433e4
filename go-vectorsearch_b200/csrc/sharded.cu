// sharded.cu -- one host process, several GPUs: the row-striped index behind the C ABI (include/vscuda.h, vs_sharded_*).
//
// The reference is ONE Go process (main.go:31) whose searches run one per goroutine (server/search.go:115); it cannot
// start a process per GPU.  A vs_sharded handle lets that one process drive G devices: the rows of the store are striped
// over the devices by primary key (key % G: every posting list is split evenly for any probe set), the
// centroid table is replicated, every device answers the batch on its stripe (the same search path as a single index:
// fused.cu / listmajor.cu / scan.cu), and the shard-local top-k lists are merged ON device 0 by a kernel that reads the
// other devices' hit buffers directly over NVLink (peer access: no staging copy, no collective library, no second
// process).  Results match a single index holding all the rows, bit for bit (tests/test_gpu_sharded.py).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/vscuda.h"
#include "internal.h"

using namespace vs;

namespace {

struct DevScope {  // the calling thread's device for the duration of a scope
    int prev;
    explicit DevScope(int d) : prev(internal_thread_device()) { internal_set_thread_device(d); }
    ~DevScope() { internal_set_thread_device(prev); }
};

struct Shard {
    int device = 0;
    vs_ctx *ctx = nullptr;  // build / upload stream
    vs_index *ix = nullptr;
};
struct Lane {  // one search context's state on one device
    vs_ctx *ctx = nullptr;
    vs_matrix *queries = nullptr;  // staging matrix of the current batch (nq rows)
    size_t q_rows = 0;
    unsigned char *hits = nullptr;  // packed [ids nq*k u64 | sims nq*k f32 | counts nq i32 | status nq u32]
    size_t hits_cap = 0;
};

#define SH_CU(call)                                                                  \
    do {                                                                             \
        cudaError_t _e = (call);                                                     \
        if (_e != cudaSuccess) {                                                     \
            char _m[256];                                                            \
            snprintf(_m, sizeof(_m), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return internal_fail(VS_ECUDA, _m);                                      \
        }                                                                            \
    } while (0)
#define SH_VS(call)                 \
    do {                            \
        int _r = (call);            \
        if (_r != VS_OK) return _r; \
    } while (0)

}  // namespace

// One search context over all the devices: a stream + scratch per device and the merge buffers on device 0 -- what one
// `calculate` closure / goroutine holds (server/search.go:230).  Searches through different contexts may overlap.
struct vs_sharded_ctx {
    vs_sharded *sh = nullptr;
    std::vector<Lane> lanes;
    const unsigned char **d_bufs = nullptr;  // [G] device pointers to the lanes' packed hit buffers (device 0)
    uint64_t *d_ids = nullptr;
    float *d_sims = nullptr;
    int32_t *d_counts = nullptr;
    size_t out_cap = 0;
    void *pinned = nullptr;
    size_t pinned_cap = 0;
};

struct vs_sharded {
    std::vector<Shard> shards;
    size_t d = 0;           // columns
    size_t n_total = 0;     // rows over all shards = the next primary key
    size_t C = 0;
    bool explicit_ids = false;
    vs_sharded_ctx *default_ctx = nullptr;  // used by vs_sharded_search (one search at a time)
    std::mutex mu;
};

static size_t packed_sims_off(size_t nq, size_t k) { return nq * k * 8; }
static size_t packed_counts_off(size_t nq, size_t k) { return (packed_sims_off(nq, k) + nq * k * 4 + 7) & ~size_t(7); }
static size_t packed_status_off(size_t nq, size_t k) { return packed_counts_off(nq, k) + nq * 4; }
static size_t packed_bytes(size_t nq, size_t k) { return (packed_status_off(nq, k) + nq * 4 + 255) & ~size_t(255); }

extern "C" int vs_sharded_create(const int *devices, size_t G, vs_sharded **out) {
    if (!devices || !out || G == 0 || G > 64) return internal_fail(VS_EINVAL, "vs_sharded_create: devices / out null, or G not in 1..64");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return internal_fail(VS_ENODEV, "no CUDA device; libvscuda has no CPU fallback");
    for (size_t g = 0; g < G; g++)
        if (devices[g] < 0 || devices[g] >= count) return internal_fail(VS_EINVAL, "vs_sharded_create: device out of range");
    // the process default (and the SM count) comes from vs_init; take the first device if the caller never called it
    if (internal_thread_device() < 0) SH_VS(vs_init(devices[0]));
    vs_sharded *sh = new vs_sharded();
    sh->shards.resize(G);
    for (size_t g = 0; g < G; g++) {
        Shard &s = sh->shards[g];
        s.device = devices[g];
        DevScope sc(s.device);
        cudaDeviceProp prop;
        cudaError_t e = cudaGetDeviceProperties(&prop, s.device);
        if (e != cudaSuccess || prop.major < 10) {
            vs_sharded_release(sh);
            return internal_fail(VS_ENODEV, "vs_sharded_create: every device must be sm_100 (libvscuda is built for sm_100a only)");
        }
        for (size_t h = 0; h < G; h++) {  // every device reads (at least) its peers' hit buffers
            if (devices[h] == s.device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, s.device, devices[h]);
            if (!can) {
                vs_sharded_release(sh);
                return internal_fail(VS_ENODEV, "vs_sharded_create: the devices have no peer access to each other (NVLink / NVSwitch expected)");
            }
            e = cudaDeviceEnablePeerAccess(devices[h], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) {
                vs_sharded_release(sh);
                return internal_fail(VS_ECUDA, "vs_sharded_create: cudaDeviceEnablePeerAccess failed");
            }
        }
        int rc = vs_ctx_create(&s.ctx);
        if (rc != VS_OK) {
            vs_sharded_release(sh);
            return rc;
        }
    }
    *out = sh;
    return VS_OK;
}

extern "C" void vs_sharded_release(vs_sharded *sh) {
    if (!sh) return;
    if (sh->default_ctx) vs_sharded_ctx_destroy(sh->default_ctx);
    for (Shard &s : sh->shards) {
        DevScope sc(s.device);
        if (s.ctx) vs_ctx_sync(s.ctx);
        if (s.ix) vs_index_release(s.ix);
        if (s.ctx) vs_ctx_destroy(s.ctx);
    }
    delete sh;
}

extern "C" size_t vs_sharded_rows(const vs_sharded *sh) { return sh ? sh->n_total : 0; }
extern "C" size_t vs_sharded_shards(const vs_sharded *sh) { return sh ? sh->shards.size() : 0; }
extern "C" size_t vs_sharded_shard_rows(const vs_sharded *sh, size_t g) {
    return (sh && g < sh->shards.size() && sh->shards[g].ix) ? vs_index_rows(sh->shards[g].ix) : 0;
}

// Rows in primary-key order with the centroid index of each (Embedding.CentroidID, database/model.go:16): row i has the
// primary key i and goes to shard i % G.  doc_ids == NULL: the primary key is the document id.
extern "C" int vs_sharded_build_assigned(vs_sharded *sh, const uint8_t *rows, size_t n, size_t row_bytes, const uint64_t *doc_ids,
                                         const uint32_t *list_of_row, const uint8_t *centroids, size_t C) {
    if (!sh || !rows || !list_of_row || !centroids) return internal_fail(VS_EINVAL, "vs_sharded_build_assigned: null argument");
    if (row_bytes <= 8) return internal_fail(VS_EEMPTY, "vector columns are empty");
    const size_t G = sh->shards.size();
    std::vector<uint8_t> srows;
    std::vector<uint64_t> sids;
    std::vector<uint32_t> slist;
    for (size_t g = 0; g < G; g++) {
        Shard &s = sh->shards[g];
        const size_t m = n > g ? (n - g + G - 1) / G : 0;
        srows.resize(m * row_bytes);
        sids.resize(m);
        slist.resize(m);
        for (size_t j = 0, i = g; j < m; j++, i += G) {
            memcpy(srows.data() + j * row_bytes, rows + i * row_bytes, row_bytes);
            sids[j] = doc_ids ? doc_ids[i] : (uint64_t)i;
            slist[j] = list_of_row[i];
        }
        DevScope sc(s.device);
        if (s.ix) {
            vs_index_release(s.ix);
            s.ix = nullptr;
        }
        if (m == 0) return internal_fail(VS_EEMPTY, "vs_sharded_build_assigned: fewer rows than shards");
        SH_VS(vs_index_build_assigned(s.ctx, srows.data(), m, row_bytes, sids.data(), slist.data(), centroids, C, &s.ix));
    }
    sh->d = row_bytes - 8;
    sh->n_total = n;
    sh->C = C;
    sh->explicit_ids = doc_ids != nullptr;
    return VS_OK;
}

// Upload (server/upload.go:239-279) on the striped store: the new rows get the primary keys n_total .. n_total + n - 1,
// row with key p joins shard p % G, in the list of its nearest centroid (upload.go:245), behind the rows already there.
extern "C" int vs_sharded_upload(vs_sharded *sh, const uint8_t *rows, size_t n, size_t row_bytes, const uint64_t *doc_ids,
                                 int64_t *assign_out) {
    if (!sh || !rows) return internal_fail(VS_EINVAL, "vs_sharded_upload: null argument");
    if (sh->shards.empty() || !sh->shards[0].ix) return internal_fail(VS_EINVAL, "vs_sharded_upload: the index has not been built");
    if (row_bytes != sh->d + 8) return internal_fail(VS_EDIM, "vector/matrix column size does not match");
    const size_t G = sh->shards.size();
    std::vector<uint8_t> srows;
    std::vector<uint64_t> sids;
    std::vector<int64_t> sassign;
    for (size_t g = 0; g < G; g++) {
        Shard &s = sh->shards[g];
        const size_t first = (g + G - sh->n_total % G) % G;  // first uploaded row whose key lands on shard g
        const size_t m = n > first ? (n - first + G - 1) / G : 0;
        if (m == 0) continue;
        srows.resize(m * row_bytes);
        sids.resize(m);
        sassign.resize(m);
        for (size_t j = 0, i = first; j < m; j++, i += G) {
            memcpy(srows.data() + j * row_bytes, rows + i * row_bytes, row_bytes);
            sids[j] = doc_ids ? doc_ids[i] : (uint64_t)(sh->n_total + i);
        }
        DevScope sc(s.device);
        vs_index *nx = nullptr;
        SH_VS(vs_index_upload(s.ctx, s.ix, srows.data(), m, row_bytes, sids.data(), sassign.data(), &nx));
        vs_index_release(s.ix);
        s.ix = nx;
        if (assign_out)
            for (size_t j = 0, i = first; j < m; j++, i += G) assign_out[i] = sassign[j];
    }
    sh->n_total += n;
    return VS_OK;
}

extern "C" int vs_sharded_ctx_create(vs_sharded *sh, vs_sharded_ctx **out) {
    if (!sh || !out) return internal_fail(VS_EINVAL, "vs_sharded_ctx_create: null argument");
    vs_sharded_ctx *sc = new vs_sharded_ctx();
    sc->sh = sh;
    sc->lanes.resize(sh->shards.size());
    for (size_t g = 0; g < sh->shards.size(); g++) {
        DevScope ds(sh->shards[g].device);
        int rc = vs_ctx_create(&sc->lanes[g].ctx);
        if (rc != VS_OK) {
            vs_sharded_ctx_destroy(sc);
            return rc;
        }
    }
    *out = sc;
    return VS_OK;
}

extern "C" void vs_sharded_ctx_destroy(vs_sharded_ctx *sc) {
    if (!sc) return;
    for (size_t g = 0; g < sc->lanes.size(); g++) {
        Lane &l = sc->lanes[g];
        DevScope ds(sc->sh->shards[g].device);
        if (l.ctx) vs_ctx_sync(l.ctx);
        if (l.queries) vs_matrix_release(l.queries);
        if (l.hits) cudaFree(l.hits);
        if (l.ctx) vs_ctx_destroy(l.ctx);
    }
    if (!sc->lanes.empty()) {
        DevScope ds(sc->sh->shards[0].device);
        if (sc->d_bufs) cudaFree(sc->d_bufs);
        if (sc->d_ids) cudaFree(sc->d_ids);
        if (sc->d_sims) cudaFree(sc->d_sims);
        if (sc->d_counts) cudaFree(sc->d_counts);
        if (sc->pinned) cudaFreeHost(sc->pinned);
    }
    delete sc;
}

// server/search.go:115-273 for a batch of queries over the striped store: every device answers on its stripe, device 0
// merges the shard-local hits it reads from its peers, one copy brings the result back.
extern "C" int vs_sharded_search_ctx(vs_sharded_ctx *sc, const uint8_t *queries, size_t nq, size_t nprobe, size_t k, uint64_t *ids_out,
                                     float *sims_out, int32_t *counts_out) {
    if (!sc || !queries || !ids_out || !sims_out || !counts_out) return internal_fail(VS_EINVAL, "vs_sharded_search: null argument");
    vs_sharded *sh = sc->sh;
    if (sh->shards.empty() || !sh->shards[0].ix) return internal_fail(VS_EINVAL, "vs_sharded_search: the index has not been built");
    if (nq == 0 || k == 0) return internal_fail(VS_EINVAL, "nq == 0 or k == 0");
    if (k > 128) return internal_fail(VS_ERANGE, "k: at most 128 hits (Count+Offset) per query");
    const size_t G = sh->shards.size();
    const size_t hb = packed_bytes(nq, k), so = packed_sims_off(nq, k), co = packed_counts_off(nq, k), to = packed_status_off(nq, k);
    // 1. every device: the batch's rows, then both search stages on its stripe (all asynchronous)
    for (size_t g = 0; g < G; g++) {
        Shard &s = sh->shards[g];
        Lane &l = sc->lanes[g];
        DevScope ds(s.device);
        if (l.q_rows != nq) {
            SH_VS(vs_ctx_sync(l.ctx));
            if (l.queries) vs_matrix_release(l.queries);
            l.queries = nullptr;
            l.q_rows = 0;
            SH_VS(vs_matrix_create_empty(l.ctx, nq, sh->d, &l.queries));
            l.q_rows = nq;
        }
        if (l.hits_cap < hb) {
            SH_VS(vs_ctx_sync(l.ctx));
            if (l.hits) cudaFree(l.hits);
            l.hits = nullptr;
            l.hits_cap = 0;
            SH_CU(cudaMalloc(&l.hits, hb));
            l.hits_cap = hb;
        }
        SH_VS(vs_matrix_load_rows(l.ctx, l.queries, 0, queries, nq));
        SH_VS(vs_search_dev(l.ctx, s.ix, l.queries, nprobe, k, reinterpret_cast<uint64_t *>(l.hits), reinterpret_cast<float *>(l.hits + so),
                            reinterpret_cast<int32_t *>(l.hits + co), reinterpret_cast<uint32_t *>(l.hits + to)));
    }
    // 2. every device: wait, and finish the (rare) queries whose float32 rounding could not be certified
    for (size_t g = 0; g < G; g++) {
        Shard &s = sh->shards[g];
        Lane &l = sc->lanes[g];
        DevScope ds(s.device);
        SH_VS(vs_search_resolve(l.ctx, s.ix, l.queries, nprobe, k, reinterpret_cast<uint64_t *>(l.hits), reinterpret_cast<float *>(l.hits + so),
                                reinterpret_cast<int32_t *>(l.hits + co), reinterpret_cast<uint32_t *>(l.hits + to), nullptr));
        SH_VS(vs_ctx_sync(l.ctx));
    }
    // 3. device 0 merges: its kernel reads the peers' hit buffers in place (P2P loads over NVLink)
    DevScope ds(sh->shards[0].device);
    cudaStream_t st = static_cast<cudaStream_t>(vs_ctx_stream(sc->lanes[0].ctx));
    if (!sc->d_bufs) SH_CU(cudaMalloc(&sc->d_bufs, 64 * sizeof(void *)));
    if (sc->out_cap < nq * k) {
        if (sc->d_ids) cudaFree(sc->d_ids);
        if (sc->d_sims) cudaFree(sc->d_sims);
        if (sc->d_counts) cudaFree(sc->d_counts);
        sc->d_ids = nullptr;
        sc->d_sims = nullptr;
        sc->d_counts = nullptr;
        sc->out_cap = 0;
        SH_CU(cudaMalloc(&sc->d_ids, nq * k * 8));
        SH_CU(cudaMalloc(&sc->d_sims, nq * k * 4));
        SH_CU(cudaMalloc(&sc->d_counts, nq * k * 4));
        sc->out_cap = nq * k;
    }
    const size_t need = nq * k * 12 + nq * 4 + 64 * sizeof(void *);
    if (sc->pinned_cap < need) {
        if (sc->pinned) cudaFreeHost(sc->pinned);
        sc->pinned = nullptr;
        sc->pinned_cap = 0;
        SH_CU(cudaMallocHost(&sc->pinned, need));
        sc->pinned_cap = need;
    }
    char *hp = static_cast<char *>(sc->pinned);
    const unsigned char **h_bufs = reinterpret_cast<const unsigned char **>(hp + nq * k * 12 + nq * 4);
    for (size_t g = 0; g < G; g++) h_bufs[g] = sc->lanes[g].hits;
    SH_CU(cudaMemcpyAsync(sc->d_bufs, h_bufs, G * sizeof(void *), cudaMemcpyHostToDevice, st));
    SH_CU(launch_topk_merge_ptrs(sc->d_bufs, 0, so, co, (int)G, (int)nq, (int)k, sc->d_ids, sc->d_sims, sc->d_counts, st));
    SH_CU(cudaMemcpyAsync(hp, sc->d_ids, nq * k * 8, cudaMemcpyDeviceToHost, st));
    SH_CU(cudaMemcpyAsync(hp + nq * k * 8, sc->d_sims, nq * k * 4, cudaMemcpyDeviceToHost, st));
    SH_CU(cudaMemcpyAsync(hp + nq * k * 12, sc->d_counts, nq * 4, cudaMemcpyDeviceToHost, st));
    SH_CU(cudaStreamSynchronize(st));
    memcpy(ids_out, hp, nq * k * 8);
    memcpy(sims_out, hp + nq * k * 8, nq * k * 4);
    memcpy(counts_out, hp + nq * k * 12, nq * 4);
    return VS_OK;
}

// The same through the handle's own context, one search at a time (callers that overlap searches create contexts).
extern "C" int vs_sharded_search(vs_sharded *sh, const uint8_t *queries, size_t nq, size_t nprobe, size_t k, uint64_t *ids_out,
                                 float *sims_out, int32_t *counts_out) {
    if (!sh) return internal_fail(VS_EINVAL, "vs_sharded_search: null argument");
    std::lock_guard<std::mutex> lk(sh->mu);
    if (!sh->default_ctx) SH_VS(vs_sharded_ctx_create(sh, &sh->default_ctx));
    return vs_sharded_search_ctx(sh->default_ctx, queries, nq, nprobe, k, ids_out, sims_out, counts_out);
}

// ---- one process PER GPU: the exchange of the shard-local hits as part of the merge kernel (vs_exchange_*) ----------
// The torchrun form of the striped index (shard.py) used one NCCL all-gather of the packed hits per step and then the
// merge kernel.  The exchange is 32 KB per rank: pure latency, and a collective launch is most of it.  Here every rank
// maps every other rank's hit buffer through CUDA IPC once; per step a one-warp kernel signals the peers and waits for
// their signals (flag words in peer memory, system-scope release / acquire) and the merge kernel reads the peers' hits in
// place over NVLink (scan.cu: exchange_wait_kernel, topk_merge_kernel): the transfer is the merge's own loads.
// Layout of a rank's allocation: [slot 0 | slot 1 | flags: one 32-bit word per rank], slot_bytes each, the flags at
// 2 * slot_bytes.  Two slots: a rank may overwrite the slot of step s at step s + 2 because finishing the merge of step
// s + 1 means every peer had signalled s + 1, which each does only after its own merge of step s (stream order).
struct vs_exchange {
    int rank = 0, world = 1, device = 0;
    size_t slot_bytes = 0;
    unsigned char *local = nullptr;                 // this rank's allocation
    std::vector<unsigned char *> peer;              // [world] base pointers (peer[rank] == local)
    const unsigned char **d_bufs[2] = {nullptr, nullptr};  // device arrays [world]: slot pointers of every rank
    uint32_t **d_signal = nullptr;                  // device array [world]: &flags_of_rank_r[this rank]
    bool connected = false;
};

extern "C" int vs_exchange_create(int rank, int world, size_t slot_bytes, vs_exchange **out) {
    if (!out || world < 1 || world > 32 || rank < 0 || rank >= world || slot_bytes == 0)
        return internal_fail(VS_EINVAL, "vs_exchange_create: bad argument (1 <= world <= 32)");
    const int dev = internal_thread_device();
    if (dev < 0) return internal_fail(VS_ENODEV, "vs_init not called");
    cudaSetDevice(dev);
    vs_exchange *x = new vs_exchange();
    x->rank = rank;
    x->world = world;
    x->device = dev;
    x->slot_bytes = (slot_bytes + 255) & ~size_t(255);
    x->peer.assign((size_t)world, nullptr);
    const size_t total = 2 * x->slot_bytes + 256;
    if (cudaMalloc(&x->local, total) != cudaSuccess || cudaMemset(x->local, 0, total) != cudaSuccess) {
        delete x;
        return internal_fail(VS_ENOMEM, "vs_exchange_create: cudaMalloc");
    }
    x->peer[(size_t)rank] = x->local;
    *out = x;
    return VS_OK;
}

extern "C" int vs_exchange_handle(const vs_exchange *x, void *handle_out, size_t cap) {
    if (!x || !handle_out || cap < sizeof(cudaIpcMemHandle_t)) return internal_fail(VS_EINVAL, "vs_exchange_handle: 64 bytes needed");
    cudaSetDevice(x->device);
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, x->local) != cudaSuccess) return internal_fail(VS_ECUDA, "cudaIpcGetMemHandle");
    memcpy(handle_out, &h, sizeof(h));
    return VS_OK;
}

// handles: world x 64 bytes, rank-major (an all-gather of vs_exchange_handle's output).  Every rank must have finished
// vs_exchange_create before any rank connects (the caller's barrier), and a second barrier after connect makes sure no
// rank signals a peer that has not opened it yet -- signals go to the OWNER's memory, so only the first is needed for
// correctness; the flags start at zero.
extern "C" int vs_exchange_connect(vs_exchange *x, const void *handles) {
    if (!x || !handles) return internal_fail(VS_EINVAL, "vs_exchange_connect: null argument");
    cudaSetDevice(x->device);
    const unsigned char *hb = static_cast<const unsigned char *>(handles);
    for (int r = 0; r < x->world; r++) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hb + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            char msg[160];
            snprintf(msg, sizeof(msg), "vs_exchange_connect: cudaIpcOpenMemHandle of rank %d: %s", r, cudaGetErrorString(e));
            return internal_fail(VS_ECUDA, msg);
        }
        x->peer[(size_t)r] = static_cast<unsigned char *>(p);
    }
    std::vector<const unsigned char *> bufs((size_t)x->world);
    std::vector<uint32_t *> sig((size_t)x->world);
    for (int s = 0; s < 2; s++) {
        for (int r = 0; r < x->world; r++) bufs[(size_t)r] = x->peer[(size_t)r] + (size_t)s * x->slot_bytes;
        if (cudaMalloc(&x->d_bufs[s], (size_t)x->world * sizeof(void *)) != cudaSuccess ||
            cudaMemcpy(x->d_bufs[s], bufs.data(), (size_t)x->world * sizeof(void *), cudaMemcpyHostToDevice) != cudaSuccess)
            return internal_fail(VS_ECUDA, "vs_exchange_connect: pointer table");
    }
    for (int r = 0; r < x->world; r++)
        sig[(size_t)r] = reinterpret_cast<uint32_t *>(x->peer[(size_t)r] + 2 * x->slot_bytes) + x->rank;
    if (cudaMalloc(&x->d_signal, (size_t)x->world * sizeof(void *)) != cudaSuccess ||
        cudaMemcpy(x->d_signal, sig.data(), (size_t)x->world * sizeof(void *), cudaMemcpyHostToDevice) != cudaSuccess)
        return internal_fail(VS_ECUDA, "vs_exchange_connect: signal table");
    x->connected = true;
    return VS_OK;
}

extern "C" void *vs_exchange_slot(const vs_exchange *x, int slot) {
    return (x && (slot == 0 || slot == 1)) ? x->local + (size_t)slot * x->slot_bytes : nullptr;
}

// step: 1, 2, 3, ... (every rank the same sequence); the local hits of this step lie in slot (step & 1), written by
// earlier work on ctx's stream.  Asynchronous.
extern "C" int vs_exchange_merge(vs_ctx *c, vs_exchange *x, uint32_t step, size_t ids_off, size_t sims_off, size_t counts_off,
                                 size_t nq, size_t k, uint64_t *d_ids_out, float *d_sims_out, int32_t *d_counts_out) {
    if (!c || !x || !d_ids_out || !d_sims_out || !d_counts_out) return internal_fail(VS_EINVAL, "vs_exchange_merge: null argument");
    if (!x->connected) return internal_fail(VS_EINVAL, "vs_exchange_merge: not connected");
    if (k > 128) return internal_fail(VS_ERANGE, "k > 128");
    if ((sims_off | counts_off) & 3 || (ids_off & 7)) return internal_fail(VS_EINVAL, "vs_exchange_merge: misaligned offsets");
    if (nq == 0) return VS_OK;
    cudaSetDevice(x->device);
    const uint32_t *wait = reinterpret_cast<const uint32_t *>(x->local + 2 * x->slot_bytes);
    // Default: a one-warp launch signals and waits, the merge kernel behind it reads the peers' hits in place.  Measured on
    // 4 GPUs (config 2): 649.6 K queries/s, the NCCL all-gather form 648.6-650.2 K (end to end 641.6 K vs 630-632 K); with
    // the signal / wait inside the merge kernel itself (VS_EXCHANGE_ONE_KERNEL=1: one launch, but a rank that arrives
    // early then holds a warp per query and the list scans of the other contexts lose their second block per SM) 600 K.
    static const bool split = [] {
        const char *e = getenv("VS_EXCHANGE_ONE_KERNEL");
        return !(e && atoi(e) != 0);
    }();
    const cudaError_t e = split ? launch_topk_merge_exchange_split(x->d_bufs[step & 1u], ids_off, sims_off, counts_off, x->world, (int)nq,
                                                                   (int)k, d_ids_out, d_sims_out, d_counts_out, x->d_signal, wait, step,
                                                                   c->stream)
                                : launch_topk_merge_exchange(x->d_bufs[step & 1u], ids_off, sims_off, counts_off, x->world, (int)nq, (int)k,
                                                             d_ids_out, d_sims_out, d_counts_out, x->d_signal, wait, step, c->stream);
    if (e != cudaSuccess) return internal_fail(VS_ECUDA, cudaGetErrorString(e));
    c->launches++;
    return VS_OK;
}

extern "C" void vs_exchange_release(vs_exchange *x) {
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; r++)
        if (r != x->rank && x->peer[(size_t)r]) cudaIpcCloseMemHandle(x->peer[(size_t)r]);
    for (int s = 0; s < 2; s++)
        if (x->d_bufs[s]) cudaFree(x->d_bufs[s]);
    if (x->d_signal) cudaFree(x->d_signal);
    if (x->local) cudaFree(x->local);
    delete x;
}
