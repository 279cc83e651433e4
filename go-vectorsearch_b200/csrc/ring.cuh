// ring.cuh -- shared-memory ring plumbing shared by the streaming kernels (fused.cu, listmajor.cu): mbarrier and
// cp.async.bulk wrappers (sm_100a PTX), and the integer dot products of one staged chunk of rows.
#pragma once
#include "common.cuh"

namespace vs {

#ifndef FULL
#define FULL 0xFFFFFFFFu
#endif

static __device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
static __device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
static __device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
static __device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
static __device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ unsigned long long fused_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Integer dots of one stage: lane group g (G lanes) takes rows g*iters .. g*iters+iters-1; returns in lane g*G+it the
// dot of row g*iters+it.
template <int G, int CPL>
static __device__ __forceinline__ uint32_t stage_dots(uint32_t s_codes, int nrows, int d_pad, const uint4 (&q)[CPL], int lane, int iters) {
    constexpr int U = (CPL <= 3) ? 4 : 2;
    const int g = lane / G, l = lane % G;
    uint32_t mydot = 0;
#pragma unroll 1
    for (int it0 = 0; it0 < iters; it0 += U) {
        uint4 v[U][CPL];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int r = g * iters + it0 + u;
            const bool ok = (it0 + u < iters) && (r < nrows);
            const uint32_t a = s_codes + (uint32_t)(ok ? r : 0) * (uint32_t)d_pad + (uint32_t)l * 16u;
#pragma unroll
            for (int j = 0; j < CPL; j++) v[u][j] = ok ? lds_u4(a + (uint32_t)(j * G * 16)) : make_uint4(0, 0, 0, 0);
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


}  // namespace vs
