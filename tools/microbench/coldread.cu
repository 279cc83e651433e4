// coldread.cu -- how fast can ONE short launch pull B bytes per SM out of HBM?  (measurement aid, not product code)
// Every block reads one contiguous span at a random place of a large buffer, either with 128-bit loads (maximal
// memory-level parallelism) or through a cp.async.bulk shared-memory ring whose consumer only waits and releases.
// Prints the launch duration (CUDA events, L2 flushed before every launch) for several span sizes.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(1024, 1) ldg_kernel(const uint4 *base, const uint64_t *off, uint64_t n16, uint32_t *sink) {
    const uint4 *p = base + off[blockIdx.x];
    uint32_t acc = 0;
    for (uint64_t i = threadIdx.x; i < n16; i += 1024 * 8) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = (i + u * 1024 < n16) ? __ldcs(p + i + u * 1024) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 8; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) sink[blockIdx.x] = acc;
}

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint32_t bar, uint32_t par) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(par) : "memory");
    } while (!done);
}

// ring of S stages of `chunk` bytes; warp 1 lane 0 produces, warp 0 consumes (waits, optionally touches, releases)
__global__ void __launch_bounds__(64, 1) bulk_kernel(const unsigned char *base, const uint64_t *off, uint64_t bytes, uint32_t chunk, int S,
                                                     int side_copies, uint32_t *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + (size_t)S * chunk);
    uint64_t *empty = full + S;
    const unsigned char *src = base + off[blockIdx.x] * 16;
    const uint32_t nch = (uint32_t)((bytes + chunk - 1) / chunk);
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[i])) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 32) {
        for (uint32_t c = 0; c < nch; c++) {
            const uint32_t st = c % S;
            if (c >= (uint32_t)S) mwait(s32(&empty[st]), ((c / S) & 1u) ^ 1u);
            const uint64_t left = bytes - (uint64_t)c * chunk;
            const uint32_t nb = left < chunk ? (uint32_t)left : chunk;
            const uint32_t sideb = 144;
            const uint32_t main_b = nb - side_copies * sideb;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(nb) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + (size_t)st * chunk)),
                         "l"(src + (uint64_t)c * chunk), "r"(main_b), "r"(s32(&full[st])) : "memory");
            for (int k = 0; k < side_copies; k++)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 s32(sm + (size_t)st * chunk + main_b + k * sideb)),
                             "l"(src + (uint64_t)c * chunk + main_b + k * sideb), "r"(sideb), "r"(s32(&full[st])) : "memory");
        }
    } else if (threadIdx.x < 32) {
        uint32_t acc = 0;
        for (uint32_t c = 0; c < nch; c++) {
            const uint32_t st = c % S;
            mwait(s32(&full[st]), (c / S) & 1u);
            acc ^= *reinterpret_cast<const uint32_t *>(sm + (size_t)st * chunk + threadIdx.x * 4);
            __syncwarp();
            if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[st])) : "memory");
        }
        if (acc == 0x12345678u) sink[blockIdx.x] = acc;
    }
}

int main(int argc, char **argv) {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t buf_bytes = (size_t)8 << 30;
    unsigned char *buf;
    CK(cudaMalloc(&buf, buf_bytes));
    CK(cudaMemset(buf, 1, buf_bytes));
    unsigned char *flush;
    const size_t flush_bytes = (size_t)512 << 20;
    CK(cudaMalloc(&flush, flush_bytes));
    uint64_t *d_off;
    uint32_t *sink;
    CK(cudaMalloc(&d_off, sms * 8));
    CK(cudaMalloc(&sink, sms * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    srand(1);
    auto run = [&](const char *name, uint64_t bytes, int mode, uint32_t chunk, int S, int side) {
        std::vector<float> t;
        for (int rep = 0; rep < 12; rep++) {
            std::vector<uint64_t> off(sms);
            for (int b = 0; b < sms; b++) {
                uint64_t slots = (buf_bytes - bytes - 4096) / 4096;
                off[b] = ((((uint64_t)rand() << 16) ^ (uint64_t)rand()) % slots) * 256;  // in 16-byte units, 4 KB aligned
            }
            CK(cudaMemcpy(d_off, off.data(), sms * 8, cudaMemcpyHostToDevice));
            CK(cudaMemset(flush, rep, flush_bytes));  // evict the L2
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            if (mode == 0) ldg_kernel<<<sms, 1024>>>(reinterpret_cast<const uint4 *>(buf), d_off, bytes / 16, sink);
            else bulk_kernel<<<sms, 64, (size_t)S * chunk + 2 * S * 8>>>(buf, d_off, bytes, chunk, S, side, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            t.push_back(ms * 1e3f);
        }
        std::sort(t.begin(), t.end());
        const float med = t[t.size() / 2];
        printf("%-34s %8.1f KB/SM  total %7.1f MB  median %8.2f us  min %8.2f us  -> %6.0f GB/s at median\n", name, bytes / 1024.0,
               bytes * (double)sms / 1e6, med, t[0], bytes * (double)sms / (med * 1e-6) / 1e9);
    };
    // an empty launch for the fixed cost of this timing method
    run("ldg (1 KB/SM: launch cost)", 1024, 0, 0, 0, 0);
    for (uint64_t kb : {128, 410, 1024, 4096, 16384, 65536}) {
        run("ldg 128-bit, 1024 thr, unroll 8", kb * 1024, 0, 0, 0, 0);
        run("bulk ring 12.7KB x16", kb * 1024, 1, 12800, 16, 0);
        run("bulk ring 12.7KB x16 +3 side", kb * 1024, 1, 12800, 16, 3);
        run("bulk ring 25.6KB x8", kb * 1024, 1, 25600, 8, 0);
        run("bulk ring 51.2KB x4", kb * 1024, 1, 51200, 4, 0);
        run("bulk ring 6.4KB x32", kb * 1024, 1, 6400, 32, 0);
    }
    return 0;
}
