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
                int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));   // stride is a power of two
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


// ---- torch's reduction order for `x.sum(dim=-1)` over a contiguous fp32 row -------------------
// The marginals are u = att / (att.sum() + 1e-5) (utilities/diml.py:110 etc.).  Sinkhorn's late-
// iteration error is dominated by a drift proportional to |sum(u)/sum(v) - 1|, which lives at the
// 1e-7 level, so a 1-ulp difference in att.sum() moves the stop test by ~10% ("n*
// fragility").  These functions therefore add in exactly the order of ATen's CPU sum kernel
// (aten/src/ATen/native/cpu/SumKernel.cpp: vectorized_inner_sum -> row_sum (ILP 4) ->
// multi_row_sum (4-level cascade), 8-lane fp32 vectors), verified bit-exact against torch 2.11
// on 200k random rows (tests/test_oracle_golden.py::test_torch_sum_order pins the formula).
__device__ __forceinline__ float torch_sum49(const float* x) {
    // 6 vectors of 8 + 1 tail element: lanes ((x0+x4)+x5) + x1 + x2 + x3, then tail + lanes in order
    float tot = x[48];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        float p0 = x[j] + x[32 + j];
        p0 += x[40 + j];
        p0 += x[8 + j];
        p0 += x[16 + j];
        p0 += x[24 + j];
        tot += p0;
    }
    return tot;
}

// General length (serial; one thread).  W = 8 lanes when n >= 8, else the scalar path (W = 1).
template <int W>
__device__ float torch_sum_inner_w(const float* x, int n) {
    constexpr int ILP = 4, LEVELS = 4;
    const int vec_size = n / W;
    const int size_ilp = vec_size / ILP;
    float acc[LEVELS][ILP][W];
    for (int a = 0; a < LEVELS; a++)
        for (int k = 0; k < ILP; k++)
            for (int l = 0; l < W; l++) acc[a][k][l] = 0.f;
    int lg = 0;
    while ((1 << lg) < size_ilp) lg++;  // CeilLog2
    const int level_power = max(4, lg / LEVELS);
    const int level_step = 1 << level_power;
    const int level_mask = level_step - 1;
    int i = 0;
    for (; i + level_step <= size_ilp;) {
        for (int j = 0; j < level_step; ++j, ++i)
            for (int k = 0; k < ILP; k++)
                for (int l = 0; l < W; l++) acc[0][k][l] += x[(i * ILP + k) * W + l];
        for (int j = 1; j < LEVELS; ++j) {
            for (int k = 0; k < ILP; k++)
                for (int l = 0; l < W; l++) {
                    acc[j][k][l] += acc[j - 1][k][l];
                    acc[j - 1][k][l] = 0.f;
                }
            const int mask = level_mask << (j * level_power);
            if ((i & mask) != 0) break;
        }
    }
    for (; i < size_ilp; ++i)
        for (int k = 0; k < ILP; k++)
            for (int l = 0; l < W; l++) acc[0][k][l] += x[(i * ILP + k) * W + l];
    for (int j = 1; j < LEVELS; ++j)
        for (int k = 0; k < ILP; k++)
            for (int l = 0; l < W; l++) acc[0][k][l] += acc[j][k][l];
    for (i = size_ilp * ILP; i < vec_size; ++i)
        for (int l = 0; l < W; l++) acc[0][0][l] += x[i * W + l];
    for (int k = 1; k < ILP; k++)
        for (int l = 0; l < W; l++) acc[0][0][l] += acc[0][k][l];
    float fin = 0.f;
    for (int k = vec_size * W; k < n; ++k) fin += x[k];
    for (int l = 0; l < W; l++) fin += acc[0][0][l];
    return fin;
}
__device__ __forceinline__ float torch_sum_inner(const float* x, int n) {
    return n >= 8 ? torch_sum_inner_w<8>(x, n) : torch_sum_inner_w<1>(x, n);
}

// The same order evaluated by a WARP: lane l < 8 carries vector lane l of ATen's 8-wide accumulators (the cascade's control
// flow is the same for all lanes), the lanes are folded in order at the end.  Every lane of the warp must call; all return
// the sum.  x may live in shared or global memory and must be visible to the whole warp.
__device__ __forceinline__ float torch_sum_inner_warp(const float* x, int n, int lane) {
    if (n < 8) {   // ATen's scalar path (ILP 4, no vector lanes)
        float s = 0.f;
        if (lane == 0) s = torch_sum_inner_w<1>(x, n);
        return __shfl_sync(0xffffffffu, s, 0);
    }
    constexpr int W = 8, ILP = 4, LEVELS = 4;
    const int l = lane & 7;
    const int vec_size = n / W;
    const int size_ilp = vec_size / ILP;
    float acc[LEVELS][ILP];
#pragma unroll
    for (int a = 0; a < LEVELS; a++)
#pragma unroll
        for (int k = 0; k < ILP; k++) acc[a][k] = 0.f;
    int lg = 0;
    while ((1 << lg) < size_ilp) lg++;
    const int level_power = max(4, lg / LEVELS);
    const int level_step = 1 << level_power;
    const int level_mask = level_step - 1;
    int i = 0;
    for (; i + level_step <= size_ilp;) {
        for (int j = 0; j < level_step; ++j, ++i)
#pragma unroll
            for (int k = 0; k < ILP; k++) acc[0][k] += x[(i * ILP + k) * W + l];
#pragma unroll
        for (int j = 1; j < LEVELS; ++j) {
#pragma unroll
            for (int k = 0; k < ILP; k++) {
                acc[j][k] += acc[j - 1][k];
                acc[j - 1][k] = 0.f;
            }
            const int mask = level_mask << (j * level_power);
            if ((i & mask) != 0) break;
        }
    }
    for (; i < size_ilp; ++i)
#pragma unroll
        for (int k = 0; k < ILP; k++) acc[0][k] += x[(i * ILP + k) * W + l];
#pragma unroll
    for (int j = 1; j < LEVELS; ++j)
#pragma unroll
        for (int k = 0; k < ILP; k++) acc[0][k] += acc[j][k];
    for (i = size_ilp * ILP; i < vec_size; ++i) acc[0][0] += x[i * W + l];
#pragma unroll
    for (int k = 1; k < ILP; k++) acc[0][0] += acc[0][k];
    float fin = 0.f;
    for (int k = vec_size * W; k < n; ++k) fin += x[k];
#pragma unroll
    for (int w = 0; w < W; w++) fin += __shfl_sync(0xffffffffu, acc[0][0], w);
    return fin;
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
