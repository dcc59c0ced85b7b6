// Microbenchmarks behind the Sinkhorn loop design of pair_fused.cu:
//   1. issue rate of FFMA, packed FFMA2 (fma.rn.f32x2) and FFMA2 + operand-duplicating MOV;
//   2. a stand-alone model of one Sinkhorn iteration in the "strip" layout: a thread owns 4 rows of a
//      pair's 49x49 Gibbs kernel in registers (row pass) and 4 columns in tensor memory (column pass),
//      2 pairs per warp, 16 pairs per CTA, the vectors r and c in shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/sk_bench tools/sk_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
typedef unsigned long long ull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ ull pack2(float lo, float hi) { ull r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(ull v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ ull ffma2(ull a, ull b, ull c) { ull d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ ull pk(uint32_t lo, uint32_t hi) { ull r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }

// ---------------- 1. issue-rate tests: 8 independent chains per thread ----------------
__global__ void rate_kernel(int mode, int iters, float* out, long long* cycles) {
    const int tid = threadIdx.x;
    float a[8], s = 1.0f + tid * 1e-6f;
    ull p[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = tid * 1e-3f + i; p[i] = pack2(a[i], a[i] + 0.5f); }
    ull s2 = pack2(s, s);
    float t = s;
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], s, 1e-7f * i + t);
    } else if (mode == 1) {
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = ffma2(p[i], s2, p[(i + 1) & 7]);
    } else if (mode == 2) {  // FFMA2 with a freshly duplicated operand every second instruction
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float lo, hi;
                unpack2(p[(i + 3) & 7], lo, hi);
                const ull d = pack2(lo, lo);
                p[i] = ffma2(p[i], d, p[i]);
                p[i + 1] = ffma2(p[i + 1], d, p[i + 1]);
            }
        }
    } else if (mode == 3) {  // 3-register FFMA, all operands varying
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], a[(i + 1) & 7], a[(i + 2) & 7]);
    } else if (mode == 4) {  // latency: one dependent FFMA chain (8 per iteration)
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[0] = fmaf(a[1 + (i & 3)], s, a[0]);
    } else if (mode == 5) {  // latency: one dependent FFMA2 chain
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 8; i++) p[0] = ffma2(p[1 + (i & 3)], s2, p[0]);
    } else if (mode == 6) {  // two dependent FFMA2 chains
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 4; i++) { p[0] = ffma2(p[2 + (i & 3)], s2, p[0]); p[1] = ffma2(p[3 + (i & 3)], s2, p[1]); }
    } else {  // four dependent FFMA chains
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < 2; i++) { a[0] = fmaf(a[4 + i], s, a[0]); a[1] = fmaf(a[5 + i], s, a[1]); a[2] = fmaf(a[6 + i], s, a[2]); a[3] = fmaf(a[4 + i], t, a[3]); }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo, hi; unpack2(p[i], lo, hi); r += a[i] + lo + hi; }
    out[blockIdx.x * blockDim.x + tid] = r;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---------------- 2. one Sinkhorn iteration in the strip layout ----------------
constexpr int W2_THREADS = 256;
constexpr int W2_PAIRS = 16;
constexpr int VP = 52;

template <int MODE>  // 0: FFMA2 + dup MOV; 1: scalar FFMA; 2: FFMA2, no column pass TMEM loads (K^T from registers: upper bound)
__global__ void __launch_bounds__(W2_THREADS, 1) strip_kernel(int iters, float* out, long long* cycles) {
    __shared__ __align__(16) float cs[W2_PAIRS * VP];
    __shared__ __align__(16) float rs[W2_PAIRS * VP];
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int half = lane >> 4, j = lane & 15, pair = warp * 2 + half;
    const bool active = j < 13;
    if (warp == 0) tmem_alloc(&tbase, 512);
    for (int i = tid; i < W2_PAIRS * VP; i += W2_THREADS) { cs[i] = 1.0f; rs[i] = 1.0f; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 196);
    ull K01[49], K23[49];
#pragma unroll
    for (int m = 0; m < 49; m++) {
        const float b = 0.01f + 0.001f * ((tid * 7 + m * 13) % 97);
        K01[m] = pack2(b, b * 1.01f);
        K23[m] = pack2(b * 1.02f, b * 1.03f);
        uint32_t kc[4];
        kc[0] = __float_as_uint(b); kc[1] = __float_as_uint(b * 1.01f); kc[2] = __float_as_uint(b * 1.02f); kc[3] = __float_as_uint(b * 1.03f);
        tmem_st4(taddr + 4 * m, kc);
    }
    tmem_wait_st();
    const float u0 = 0.02f, v0 = 0.02f;
    float r0 = 1.f, r1 = 1.f, r2 = 1.f, r3 = 1.f, esum = 0.f;
    const float4* c4 = reinterpret_cast<const float4*>(cs + pair * VP);
    const float4* r4 = reinterpret_cast<const float4*>(rs + pair * VP);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        // ---- row pass ----
        float y0, y1, y2, y3;
        if (MODE == 1) {
            y0 = y1 = y2 = y3 = 0.f;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const float4 cv = c4[i];
                const float cc[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float a, b, c, d;
                    unpack2(K01[4 * i + q], a, b);
                    unpack2(K23[4 * i + q], c, d);
                    y0 = fmaf(a, cc[q], y0); y1 = fmaf(b, cc[q], y1); y2 = fmaf(c, cc[q], y2); y3 = fmaf(d, cc[q], y3);
                }
            }
            {
                const float cl = cs[pair * VP + 48];
                float a, b, c, d;
                unpack2(K01[48], a, b);
                unpack2(K23[48], c, d);
                y0 = fmaf(a, cl, y0); y1 = fmaf(b, cl, y1); y2 = fmaf(c, cl, y2); y3 = fmaf(d, cl, y3);
            }
        } else {
            ull y01 = 0ull, y23 = 0ull;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const float4 cv = c4[i];
                const float cc[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const ull c2 = pack2(cc[q], cc[q]);
                    y01 = ffma2(K01[4 * i + q], c2, y01);
                    y23 = ffma2(K23[4 * i + q], c2, y23);
                }
            }
            {
                const float cl = cs[pair * VP + 48];
                const ull c2 = pack2(cl, cl);
                y01 = ffma2(K01[48], c2, y01);
                y23 = ffma2(K23[48], c2, y23);
            }
            unpack2(y01, y0, y1);
            unpack2(y23, y2, y3);
        }
        {
            const float n0 = u0 / y0, n1 = u0 / y1, n2 = u0 / y2, n3 = u0 / y3;
            esum += fabsf(n0 - r0) + fabsf(n1 - r1) + fabsf(n2 - r2) + fabsf(n3 - r3);
            r0 = n0; r1 = n1; r2 = n2; r3 = n3;
            if (active) *reinterpret_cast<float4*>(rs + pair * VP + 4 * j) = make_float4(n0, n1, n2, n3);
        }
        __syncwarp();
        // ---- column pass ----
        float x0, x1, x2, x3;
        {
            ull x01 = 0ull, x23 = 0ull;
            if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    const float4 rv = r4[i];
                    const float rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const ull q2 = pack2(rr[q], rr[q]);
                        x01 = ffma2(K23[4 * i + q], q2, x01);
                        x23 = ffma2(K01[4 * i + q], q2, x23);
                    }
                }
            } else {
                uint32_t ka[16], kb[16];
                tmem_ld16(taddr, ka);
#pragma unroll
                for (int i = 0; i < 12; i += 2) {
                    tmem_wait_ld();
                    tmem_ld16(taddr + 16 * (i + 1), kb);
                    {
                        const float4 rv = r4[i];
                        const float rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const ull q2 = pack2(rr[q], rr[q]);
                            x01 = ffma2(pk(ka[4 * q], ka[4 * q + 1]), q2, x01);
                            x23 = ffma2(pk(ka[4 * q + 2], ka[4 * q + 3]), q2, x23);
                        }
                    }
                    tmem_wait_ld();
                    if (i + 2 < 12) tmem_ld16(taddr + 16 * (i + 2), ka);
                    else tmem_ld4(taddr + 192, ka);
                    {
                        const float4 rv = r4[i + 1];
                        const float rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const ull q2 = pack2(rr[q], rr[q]);
                            x01 = ffma2(pk(kb[4 * q], kb[4 * q + 1]), q2, x01);
                            x23 = ffma2(pk(kb[4 * q + 2], kb[4 * q + 3]), q2, x23);
                        }
                    }
                }
                tmem_wait_ld();
                const float rl = rs[pair * VP + 48];
                const ull q2 = pack2(rl, rl);
                x01 = ffma2(pk(ka[0], ka[1]), q2, x01);
                x23 = ffma2(pk(ka[2], ka[3]), q2, x23);
            }
            unpack2(x01, x0, x1);
            unpack2(x23, x2, x3);
        }
        if (active) *reinterpret_cast<float4*>(cs + pair * VP + 4 * j) = make_float4(v0 / x0, v0 / x1, v0 / x2, v0 / x3);
        __syncwarp();
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * W2_THREADS + tid] = esum + r0 + cs[pair * VP + j];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// ---------------- 3. the S3 inner loop: 98 independent FFMA2 accumulators per thread and channel ----------------
template <int MODE>  // 0: as in the kernel (A row by LDS.128, F by LDS.32); 1: operands from registers only; 2: scalar FFMA
__global__ void __launch_bounds__(256, 1) s3_kernel(int nch, float* out, long long* cycles) {
    __shared__ __align__(16) float Aq[64 * 52];
    __shared__ __align__(16) float Fs[16 * 8 * 49 + 64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = lane & 15, ps = warp * 2 + (lane >> 4), jc = j < 13 ? j : 12;
    for (int i = tid; i < 64 * 52; i += 256) Aq[i] = 0.001f * (i % 97);
    for (int i = tid; i < 16 * 8 * 49 + 64; i += 256) Fs[i] = 0.002f * (i % 89);
    __syncthreads();
    ull K01[49], K23[49];
#pragma unroll
    for (int m = 0; m < 49; m++) K01[m] = K23[m] = 0ull;
    const long long t0 = clock64();
    for (int ch = 0; ch < nch; ch++) {
        const float* F = Fs + ps * 392 + 4 * jc + (ch & 7) * 49;
        const float f0 = F[0], f1 = F[1], f2 = F[2], f3 = F[3];
        if (MODE == 2) {
            float* k01 = reinterpret_cast<float*>(K01);
            float* k23 = reinterpret_cast<float*>(K23);
            const float4* A4 = reinterpret_cast<const float4*>(Aq + (ch & 63) * 52);
#pragma unroll
            for (int q = 0; q < 12; q++) {
                const float4 av = A4[q];
                const float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    k01[2 * (4 * q + t)] = fmaf(f0, aa[t], k01[2 * (4 * q + t)]);
                    k01[2 * (4 * q + t) + 1] = fmaf(f1, aa[t], k01[2 * (4 * q + t) + 1]);
                    k23[2 * (4 * q + t)] = fmaf(f2, aa[t], k23[2 * (4 * q + t)]);
                    k23[2 * (4 * q + t) + 1] = fmaf(f3, aa[t], k23[2 * (4 * q + t) + 1]);
                }
            }
        } else {
            const ull f01 = pack2(f0, f1), f23 = pack2(f2, f3);
            const float4* A4 = reinterpret_cast<const float4*>(Aq + (ch & 63) * 52);
#pragma unroll
            for (int q = 0; q < 12; q++) {
                float4 av;
                if (MODE == 0) av = A4[q];
                else av = make_float4(f0 + q, f1 + q, f2 + q, f3 + q);
                K01[4 * q + 0] = ffma2(f01, pack2(av.x, av.x), K01[4 * q + 0]);
                K23[4 * q + 0] = ffma2(f23, pack2(av.x, av.x), K23[4 * q + 0]);
                K01[4 * q + 1] = ffma2(f01, pack2(av.y, av.y), K01[4 * q + 1]);
                K23[4 * q + 1] = ffma2(f23, pack2(av.y, av.y), K23[4 * q + 1]);
                K01[4 * q + 2] = ffma2(f01, pack2(av.z, av.z), K01[4 * q + 2]);
                K23[4 * q + 2] = ffma2(f23, pack2(av.z, av.z), K23[4 * q + 2]);
                K01[4 * q + 3] = ffma2(f01, pack2(av.w, av.w), K01[4 * q + 3]);
                K23[4 * q + 3] = ffma2(f23, pack2(av.w, av.w), K23[4 * q + 3]);
            }
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int m = 0; m < 49; m++) { float a, b, c, d; unpack2(K01[m], a, b); unpack2(K23[m], c, d); r += a + b + c + d; }
    out[blockIdx.x * 256 + tid] = r;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}


// ---------------- 2b. the same iteration in a HALF-strip layout: 2 rows / 2 columns per lane, one pair per warp (25 of 32
// lanes), 16 warps per CTA = 4 warps per scheduler at <= 128 registers (the 4-row layout holds 2 warps per scheduler) ----------------
constexpr int W3_THREADS = 512;
__global__ void __launch_bounds__(W3_THREADS, 1) halfstrip_kernel(int iters, float* out, long long* cycles) {
    __shared__ __align__(16) float cs[W2_PAIRS * VP];
    __shared__ __align__(16) float rs[W2_PAIRS * VP];
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pair = warp;
    const bool active = lane < 25;
    if (warp == 0) tmem_alloc(&tbase, 512);
    for (int i = tid; i < W2_PAIRS * VP; i += W3_THREADS) { cs[i] = 1.0f; rs[i] = 1.0f; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // 4 warps share a TMEM lane quarter: 128 columns each (98 used: [s][2 owned columns])
    const uint32_t taddr = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 128);
    ull K01[49];
#pragma unroll
    for (int m = 0; m < 49; m++) {
        const float b = 0.01f + 0.001f * ((tid * 7 + m * 13) % 97);
        K01[m] = pack2(b, b * 1.01f);
    }
    for (int m = 0; m < 48; m += 2) {
        uint32_t kc[4];
        const float b = 0.01f + 0.001f * ((tid * 7 + m * 13) % 97);
        kc[0] = __float_as_uint(b); kc[1] = __float_as_uint(b * 1.01f); kc[2] = __float_as_uint(b * 1.02f); kc[3] = __float_as_uint(b * 1.03f);
        tmem_st4(taddr + 2 * m, kc);
    }
    {
        uint32_t kc[4] = {__float_as_uint(0.02f), __float_as_uint(0.021f), 0u, 0u};
        tmem_st4(taddr + 96, kc);
    }
    tmem_wait_st();
    const float u0 = 0.02f, v0 = 0.02f;
    float r0 = 1.f, r1 = 1.f, esum = 0.f;
    const float4* c4 = reinterpret_cast<const float4*>(cs + pair * VP);
    const float4* r4 = reinterpret_cast<const float4*>(rs + pair * VP);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        // ---- row pass: 49 FFMA2 from registers, 13 LDS.128 (whole warp reads one address: a broadcast) ----
        float y0, y1;
        {
            ull y01 = 0ull;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const float4 cv = c4[i];
                y01 = ffma2(K01[4 * i + 0], pack2(cv.x, cv.x), y01);
                y01 = ffma2(K01[4 * i + 1], pack2(cv.y, cv.y), y01);
                y01 = ffma2(K01[4 * i + 2], pack2(cv.z, cv.z), y01);
                y01 = ffma2(K01[4 * i + 3], pack2(cv.w, cv.w), y01);
            }
            const float cl = cs[pair * VP + 48];
            y01 = ffma2(K01[48], pack2(cl, cl), y01);
            unpack2(y01, y0, y1);
        }
        {
            const float n0 = u0 / y0, n1 = u0 / y1;
            esum += fabsf(n0 - r0) + fabsf(n1 - r1);
            r0 = n0; r1 = n1;
            if (active) *reinterpret_cast<float2*>(rs + pair * VP + 2 * lane) = make_float2(n0, n1);
        }
        __syncwarp();
        // ---- column pass: 2 owned columns x 49 rows from tensor memory (x16 loads = 8 rows each) ----
        float x0, x1;
        {
            ull x01 = 0ull;
            uint32_t ka[16], kb[16];
            tmem_ld16(taddr, ka);
#pragma unroll
            for (int i = 0; i < 6; i += 2) {
                tmem_wait_ld();
                tmem_ld16(taddr + 16 * (i + 1), kb);
                {
                    const float4 ra = r4[2 * i], rb = r4[2 * i + 1];
                    const float rr[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                    for (int q = 0; q < 8; q++) x01 = ffma2(pk(ka[2 * q], ka[2 * q + 1]), pack2(rr[q], rr[q]), x01);
                }
                tmem_wait_ld();
                if (i + 2 < 6) tmem_ld16(taddr + 16 * (i + 2), ka);
                else tmem_ld4(taddr + 96, ka);
                {
                    const float4 ra = r4[2 * i + 2], rb = r4[2 * i + 3];
                    const float rr[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                    for (int q = 0; q < 8; q++) x01 = ffma2(pk(kb[2 * q], kb[2 * q + 1]), pack2(rr[q], rr[q]), x01);
                }
            }
            tmem_wait_ld();
            const float rl = rs[pair * VP + 48];
            x01 = ffma2(pk(ka[0], ka[1]), pack2(rl, rl), x01);
            unpack2(x01, x0, x1);
        }
        if (active) *reinterpret_cast<float2*>(cs + pair * VP + 2 * lane) = make_float2(v0 / x0, v0 / x1);
        __syncwarp();
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * W3_THREADS + tid] = esum + r0 + cs[pair * VP + (lane & 15)];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    const int blocks = 148;
    float* out;
    long long* cyc;
    CHECK(cudaMalloc(&out, blocks * 1024 * 4));
    CHECK(cudaMalloc(&cyc, blocks * 8));
    long long h[blocks];
    const char* rn[8] = {"FFMA (reg, reg, const-ish)", "FFMA2", "FFMA2 + dup MOV per 2", "FFMA 3 distinct regs", "FFMA 1 chain", "FFMA2 1 chain", "FFMA2 2 chains", "FFMA 4 chains"};
    for (int threads = 128; threads <= 512; threads *= 2)
        for (int mode = 0; mode < 8; mode++) {
            const int iters = 4096;
            rate_kernel<<<blocks, threads>>>(mode, iters, out, cyc);
            CHECK(cudaDeviceSynchronize());
            CHECK(cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost));
            double mean = 0;
            for (int i = 0; i < blocks; i++) mean += (double)h[i];
            mean /= blocks;
            const double winst = (double)iters * 8 * (threads / 32) / 4;  // FMA-pipe warp instructions per SMSP
            printf("rate  %-28s %3d threads: %.3f cycles per FMA-pipe warp-instruction per SMSP\n", rn[mode], threads, mean / winst);
        }
    const int iters = 500;
    for (int mode = 0; mode < 3; mode++) {
        if (mode == 0) strip_kernel<0><<<blocks, W2_THREADS>>>(iters, out, cyc);
        if (mode == 1) strip_kernel<1><<<blocks, W2_THREADS>>>(iters, out, cyc);
        if (mode == 2) strip_kernel<2><<<blocks, W2_THREADS>>>(iters, out, cyc);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost));
        double mean = 0;
        for (int i = 0; i < blocks; i++) mean += (double)h[i] / iters;
        mean /= blocks;
        const char* sn[3] = {"FFMA2 + TMEM columns", "FFMA  + TMEM columns", "FFMA2, no TMEM (bound)"};
        printf("strip %-24s: %.1f cycles per iteration per CTA (16 pairs) = %.1f cycles per pair-iteration\n", sn[mode], mean, mean / 16);
    }
    {
        halfstrip_kernel<<<blocks, W3_THREADS>>>(iters, out, cyc);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost));
        double mean = 0;
        for (int i = 0; i < blocks; i++) mean += (double)h[i] / iters;
        mean /= blocks;
        printf("strip %-24s: %.1f cycles per iteration per CTA (16 pairs) = %.1f cycles per pair-iteration\n", "HALF strips, 16 warps", mean, mean / 16);
    }
    for (int mode = 0; mode < 3; mode++) {
        const int nch = 4096;
        if (mode == 0) s3_kernel<0><<<blocks, 256>>>(nch, out, cyc);
        if (mode == 1) s3_kernel<1><<<blocks, 256>>>(nch, out, cyc);
        if (mode == 2) s3_kernel<2><<<blocks, 256>>>(nch, out, cyc);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost));
        double mean = 0;
        for (int i = 0; i < blocks; i++) mean += (double)h[i] / nch;
        mean /= blocks;
        const char* sn[3] = {"FFMA2, LDS operands", "FFMA2, register operands", "FFMA, LDS operands"};
        printf("s3    %-26s: %.1f cycles per channel per CTA (8 warps x 96 FFMA2) -> %.2f cycles per FFMA2 per SMSP\n", sn[mode], mean, mean / (2 * 96));
    }
    printf("done\n");
    return 0;
}
