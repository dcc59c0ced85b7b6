// Shared device/host helpers for libvitrerank (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vitrerank.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvitrerank targets sm_100a (B200) only"
#endif

namespace vr {

// ---- host-side error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local long long g_launches;

#define VR_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            vr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                     \
            return VR_E_CUDA;                                                            \
        }                                                                                \
    } while (0)

#define VR_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            vr::set_error(__VA_ARGS__);  \
            return VR_E_INVALID;         \
        }                                \
    } while (0)

#define VR_LAUNCH_CHECK()                                  \
    do {                                                   \
        vr::g_launches++;                                  \
        VR_CHECK_CUDA(cudaGetLastError());                 \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers -----------------------------------------------------------------
#ifdef __CUDACC__

// Monotone map fp32 -> u32 (larger float -> larger integer; NaN with sign 0 sorts above +inf,
// which is where torch.argsort(descending=True) puts NaN).
__device__ __forceinline__ uint32_t ordered_bits(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
// (score, index) packed so that a descending u64 sort orders by score desc, then index asc.
__device__ __forceinline__ unsigned long long pack_key(float score, uint32_t idx) {
    return ((unsigned long long)ordered_bits(score) << 32) | (unsigned long long)(0xffffffffu - idx);
}
__device__ __forceinline__ uint32_t key_index(unsigned long long k) {
    return 0xffffffffu - (uint32_t)(k & 0xffffffffull);
}
__device__ __forceinline__ float key_score(unsigned long long k) {
    return from_ordered_bits((uint32_t)(k >> 32));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

// In-place descending bitonic sort of `n` (power of two, >= 32) u64 keys in shared memory by
// ONE warp.
__device__ __forceinline__ void warp_bitonic_sort_desc(unsigned long long* a, int n, int lane) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncwarp();
            for (int t = lane; t < (n >> 1); t += 32) {
                int lo = ((t / stride) * (stride << 1)) + (t % stride);
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                unsigned long long x = a[lo], y = a[hi];
                bool swap = desc ? (x < y) : (x > y);
                if (swap) {
                    a[lo] = y;
                    a[hi] = x;
                }
            }
        }
    }
    __syncwarp();
}

// ---- mbarrier + bulk async copy (TMA 1-D) --------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (16-byte aligned, size multiple of 16), completes on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

#endif  // __CUDACC__

}  // namespace vr
