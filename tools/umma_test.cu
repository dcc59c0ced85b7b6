// Stand-alone check of the tcgen05.mma (kind::tf32) building block used by S3 of pair_fused.cu:
// D[128 x 64] = sum over k-steps of A[128 x 8] * B[64 x 8]^T, operands K-major without swizzle in
// shared memory, accumulator in tensor memory, with the 3-term split a = hi + lo (hi = a truncated to
// tf32, lo = a - hi):  D = Ahi*Bhi + Alo*Bhi + Ahi*Blo.
// Prints the error of the plain tf32 product and of the 3-term product against an fp64 reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_test tools/umma_test.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int M = 128, N = 64, KS = 8;       // one MMA: M x N x 8 (tf32)
constexpr int NK = 16;                       // k-steps (K = 128 channels)
constexpr int A_TILE = M * KS;               // floats per k-step
constexpr int B_TILE = N * KS;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: core matrix = 8 rows x 16 bytes, rows 16 B apart; LBO = distance between the two
// core matrices along K, SBO = distance between 8-row groups along M/N (cute/arch/mma_sm100_desc.hpp).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // layout type 0 = no swizzle, base offset 0
}
// element (row, k) of a tile with `rows` rows: float index inside the tile
__host__ __device__ inline int tile_index(int row, int k, int rows) {
    return (k / 4) * (rows * 4) + (row / 8) * 32 + (row % 8) * 4 + (k % 4);
}

__global__ void __launch_bounds__(128, 1) umma_kernel(const float* A, const float* B, float* D, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* Ahi = reinterpret_cast<float*>(smem);            // [NK][A_TILE]
    float* Alo = Ahi + NK * A_TILE;
    float* Bhi = Alo + NK * A_TILE;                         // [NK][B_TILE]
    float* Blo = Bhi + NK * B_TILE;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    // A[row][k] row-major [M][NK*8], B[n][k] row-major [N][NK*8]
    for (int i = tid; i < M * NK * KS; i += 128) {
        const int row = i / (NK * KS), k = i % (NK * KS);
        const float x = A[i];
        float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
        if (mode >= 3) hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);  // round to nearest tf32
        float lo = x - hi;
        if (mode >= 3) lo = __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xffffe000u);
        const int o = (k / KS) * A_TILE + tile_index(row, k % KS, M);
        Ahi[o] = (mode == 0) ? x : hi;
        Alo[o] = lo;
    }
    for (int i = tid; i < N * NK * KS; i += 128) {
        const int row = i / (NK * KS), k = i % (NK * KS);
        const float x = B[i];
        float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
        if (mode >= 3) hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
        float lo = x - hi;
        if (mode >= 3) lo = __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xffffe000u);
        const int o = (k / KS) * B_TILE + tile_index(row, k % KS, N);
        Bhi[o] = (mode == 0) ? x : hi;
        Blo[o] = lo;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tbase;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (mode == 6 && tid == 0) {   // issue-rate probe: precomputed descriptors, back-to-back MMAs
        const uint64_t ah = make_desc(smem_u32(Ahi), M * 16, 128), bh = make_desc(smem_u32(Bhi), N * 16, 128);
        for (int n = 1; n <= 64; n *= 2) {
            // shapes: n <= 16 uses N = 64; n == 32 uses N = 32 ... (see idx)
            const uint32_t nn = (n == 16) ? 128u : (n == 32 ? 32u : (n == 64 ? 16u : 64u));
            const uint32_t idx = (1u << 4) | (2u << 7) | (2u << 10) | ((nn >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
            const long long t0 = clock64();
            for (int i = 0; i < 32; i++)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tm), "l"(ah + (uint64_t)(i & 7) * 256), "l"(bh), "r"(idx), "r"(1u) : "memory");
            const long long t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            const int par = (n == 1 ? 0 : (n == 2 ? 1 : (n == 4 ? 0 : (n == 8 ? 1 : (n == 16 ? 0 : (n == 32 ? 1 : 0))))));
            asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DN2;\nbra W2;\nDN2:\n}\n" ::"r"(smem_u32(&bar)), "r"(par) : "memory");
            const long long t2 = clock64();
            printf("  32 back-to-back MMAs, N = %u: issue %lld cycles, until complete %lld cycles (%.1f per MMA)\n", nn, t1 - t0, t2 - t0, (double)(t2 - t0) / 32);
        }
    }
    long long t_issue0 = clock64(), t_issue1 = 0;
    if (tid == 0 && mode != 6) {
        for (int ks = 0; ks < NK; ks++) {
            const uint64_t ah = make_desc(smem_u32(Ahi + ks * A_TILE), M * 16, 128);
            const uint64_t al = make_desc(smem_u32(Alo + ks * A_TILE), M * 16, 128);
            const uint64_t bh = make_desc(smem_u32(Bhi + ks * B_TILE), N * 16, 128);
            const uint64_t bl = make_desc(smem_u32(Blo + ks * B_TILE), N * 16, 128);
            const uint32_t acc0 = ks > 0 ? 1u : 0u;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tm), "l"(ah), "l"(bh), "r"(idesc), "r"(acc0) : "memory");
            if (mode >= 2) {
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tm), "l"(al), "l"(bh), "r"(idesc), "r"(1u) : "memory");
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tm), "l"(ah), "l"(bl), "r"(idesc), "r"(1u) : "memory");
            }
            if (mode == 4)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tm), "l"(al), "l"(bl), "r"(idesc), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        t_issue1 = clock64();
    }
    // everyone waits for the MMAs
    if (mode != 6) asm volatile(
        "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DN;\nbra W;\nDN:\n}\n" ::"r"(smem_u32(&bar)), "r"(0)
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0 && mode != 6) {
        const long long t2 = clock64();
        const int nmma = NK * (mode >= 2 ? (mode == 4 ? 4 : 3) : 1);
        printf("  mode %d: %d MMAs (128x64x8 tf32): issue %lld cycles, issue+complete %lld cycles -> %.1f cycles per MMA\n", mode, nmma,
               t_issue1 - t_issue0, t2 - t_issue0, (double)(t2 - t_issue0) / nmma);
    }
    const uint32_t taddr = tm + ((uint32_t)(32 * warp) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; i++) D[tid * N + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128) : "memory");
}

// ---- fp16 variant: x' = 64 x = hi + lo in fp16 (lo may be subnormal: its absolute precision 2^-25 is far below
// the product scale), D = (Ahi*Bhi + Alo*Bhi + Ahi*Blo) / 4096, kind::f16, K = 16 per MMA ----
#include <cuda_fp16.h>
constexpr int KS16 = 16, NK16 = 8;
__host__ __device__ inline int tile_index16(int row, int k) { return (k / 8) * (128 * 8) + row * 8 + (k % 8); }   // halves, 128-row tile
__host__ __device__ inline int tile_index16b(int row, int k) { return (k / 8) * (64 * 8) + row * 8 + (k % 8); }   // 64-row tile
__global__ void __launch_bounds__(128, 1) umma16_kernel(const float* A, const float* B, float* D, int nterms) {
    extern __shared__ __align__(128) unsigned char smem[];
    __half* Ahi = reinterpret_cast<__half*>(smem);            // [NK16][128*16]
    __half* Alo = Ahi + NK16 * 128 * 16;
    __half* Bhi = Alo + NK16 * 128 * 16;                      // [NK16][64*16]
    __half* Blo = Bhi + NK16 * 64 * 16;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = NK16 * KS16;
    for (int i = tid; i < M * K; i += 128) {
        const int row = i / K, k = i % K;
        const float x = A[i] * 64.f;
        const __half h = __float2half_rn(x);
        const __half l = __float2half_rn(x - __half2float(h));
        const int o = (k / KS16) * (128 * 16) + tile_index16(row, k % KS16);
        Ahi[o] = h; Alo[o] = l;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int row = i / K, k = i % K;
        const float x = B[i] * 64.f;
        const __half h = __float2half_rn(x);
        const __half l = __float2half_rn(x - __half2float(h));
        const int o = (k / KS16) * (64 * 16) + tile_index16b(row, k % KS16);
        Bhi[o] = h; Blo[o] = l;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tbase;
    // D = F32 (bit 4), A = B = F16 (format 0), K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (tid == 0 && nterms == 3) {   // execution time of 24 MMAs versus the A-operand strides (values are garbage here)
        const uint32_t lbos[5] = {2048, 128, 128, 128, 128}, sbos[5] = {128, 256, 512, 1024, 2048};
        for (int v = 0; v < 5; v++) {
            const uint64_t ad = make_desc(smem_u32(Ahi), lbos[v], sbos[v]);
            const uint64_t bd = make_desc(smem_u32(Bhi), 64 * 16, 128);
            const long long t0 = clock64();
            for (int i = 0; i < 24; i++)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tm), "l"(ad + (uint64_t)(i & 3) * 16), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            const long long t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            asm volatile("{\n.reg .pred p;\nW5:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DN5;\nbra W5;\nDN5:\n}\n" ::"r"(smem_u32(&bar)), "r"(v & 1) : "memory");
            const long long t2 = clock64();
            printf("  24 f16 MMAs (128x64x16), A LBO %4u SBO %4u: issue %lld, until complete %lld cycles\n", lbos[v], sbos[v], t1 - t0, t2 - t0);
        }
    }
    const uint32_t wait_parity = (nterms == 3) ? 1u : 0u;
    if (tid == 0) {
        for (int ks = 0; ks < NK16; ks++) {
            const uint64_t ah = make_desc(smem_u32(Ahi + ks * 128 * 16), 128 * 16, 128);
            const uint64_t al = make_desc(smem_u32(Alo + ks * 128 * 16), 128 * 16, 128);
            const uint64_t bh = make_desc(smem_u32(Bhi + ks * 64 * 16), 64 * 16, 128);
            const uint64_t bl = make_desc(smem_u32(Blo + ks * 64 * 16), 64 * 16, 128);
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tm), "l"(al), "l"(bh), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tm), "l"(ah), "l"(bl), "r"(idesc), "r"(1u) : "memory");
            if (nterms == 4)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tm), "l"(al), "l"(bl), "r"(idesc), "r"(1u) : "memory");
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tm), "l"(ah), "l"(bh), "r"(idesc), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW3:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DN3;\nbra W3;\nDN3:\n}\n" ::"r"(smem_u32(&bar)), "r"(wait_parity) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tm + ((uint32_t)(32 * warp) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; i++) D[tid * N + c0 + i] = __uint_as_float(r[i]) * (1.0f / 4096.0f);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64) : "memory");
}

int main() {
    const int K = NK * KS;
    float *hA = (float*)malloc(M * K * 4), *hB = (float*)malloc(N * K * 4), *hD = (float*)malloc(M * N * 4);
    srand(1);
    // unit-norm rows like the L2-normalised patch features
    for (int r = 0; r < M; r++) {
        double s = 0;
        for (int k = 0; k < K; k++) { hA[r * K + k] = (float)rand() / RAND_MAX - 0.5f; s += (double)hA[r * K + k] * hA[r * K + k]; }
        for (int k = 0; k < K; k++) hA[r * K + k] /= (float)sqrt(s);
    }
    for (int r = 0; r < N; r++) {
        double s = 0;
        for (int k = 0; k < K; k++) { hB[r * K + k] = (float)rand() / RAND_MAX - 0.5f + (r < M ? 0.5f * hA[r * K + k] : 0.f); s += (double)hB[r * K + k] * hB[r * K + k]; }
        for (int k = 0; k < K; k++) hB[r * K + k] /= (float)sqrt(s);
    }
    float *dA, *dB, *dD;
    CHECK(cudaMalloc(&dA, M * K * 4)); CHECK(cudaMalloc(&dB, N * K * 4)); CHECK(cudaMalloc(&dD, M * N * 4));
    CHECK(cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(2 * NK * A_TILE + 2 * NK * B_TILE) * 4;
    CHECK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char* names[7] = {"1 x tf32 (raw fp32 operands)", "1 x tf32 (pre-truncated)", "3 x tf32 split (truncate)", "3 x tf32 split (round)", "4 x tf32 split (round)", "tf32-exact inputs (accumulation only)", "issue-rate probe"};
    for (int mode = 0; mode < 7; mode++) {
        if (mode == 5) {  // make the inputs exactly representable: what remains is the accumulation error
            for (int i = 0; i < M * K; i++) { uint32_t u; memcpy(&u, &hA[i], 4); u &= 0xffffe000u; memcpy(&hA[i], &u, 4); }
            for (int i = 0; i < N * K; i++) { uint32_t u; memcpy(&u, &hB[i], 4); u &= 0xffffe000u; memcpy(&hB[i], &u, 4); }
            CHECK(cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice));
            CHECK(cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice));
        }
        CHECK(cudaMemset(dD, 0, M * N * 4));
        umma_kernel<<<1, 128, smem>>>(dA, dB, dD, mode == 5 ? 1 : mode);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost));
        double maxe = 0, maxf = 0, bias = 0;
        for (int r = 0; r < M; r++)
            for (int n = 0; n < N; n++) {
                double ref = 0;
                float f = 0.f;
                for (int k = 0; k < K; k++) { ref += (double)hA[r * K + k] * hB[n * K + k]; f = fmaf(hA[r * K + k], hB[n * K + k], f); }
                maxe = fmax(maxe, fabs(hD[r * N + n] - ref));
                maxf = fmax(maxf, fabs((double)f - ref));
                bias += hD[r * N + n] - ref;
            }
        printf("%-30s: max |D - fp64| = %.3e (sequential fp32 FMA chain: %.3e), mean signed error %.3e, D[5][7] = %.8f\n", names[mode], maxe, maxf,
               bias / (M * N), hD[5 * N + 7]);
    }
    {   // fp16 split
        srand(1);
        for (int r = 0; r < M; r++) {
            double s2 = 0;
            for (int k = 0; k < K; k++) { hA[r * K + k] = (float)rand() / RAND_MAX - 0.5f; s2 += (double)hA[r * K + k] * hA[r * K + k]; }
            for (int k = 0; k < K; k++) hA[r * K + k] /= (float)sqrt(s2);
        }
        for (int r = 0; r < N; r++) {
            double s2 = 0;
            for (int k = 0; k < K; k++) { hB[r * K + k] = (float)rand() / RAND_MAX - 0.5f + (r < M ? 0.5f * hA[r * K + k] : 0.f); s2 += (double)hB[r * K + k] * hB[r * K + k]; }
            for (int k = 0; k < K; k++) hB[r * K + k] /= (float)sqrt(s2);
        }
        // a few tiny entries to exercise subnormal lo parts
        for (int k = 0; k < K; k += 7) hA[3 * K + k] *= 1e-3f;
        CHECK(cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice));
        CHECK(cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice));
        const size_t smem16 = (size_t)(2 * NK16 * 128 * 16 + 2 * NK16 * 64 * 16) * 2;
        CHECK(cudaFuncSetAttribute(umma16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
        for (int nterms = 3; nterms <= 4; nterms++) {
            umma16_kernel<<<1, 128, smem16>>>(dA, dB, dD, nterms);
            CHECK(cudaDeviceSynchronize());
            CHECK(cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost));
            double maxe = 0, maxf = 0, bias = 0;
            for (int r = 0; r < M; r++)
                for (int n = 0; n < N; n++) {
                    double ref = 0;
                    float f = 0.f;
                    for (int k = 0; k < K; k++) { ref += (double)hA[r * K + k] * hB[n * K + k]; f = fmaf(hA[r * K + k], hB[n * K + k], f); }
                    maxe = fmax(maxe, fabs(hD[r * N + n] - ref));
                    maxf = fmax(maxf, fabs((double)f - ref));
                    bias += hD[r * N + n] - ref;
                }
            printf("%d x fp16 split (x64 scaling)    : max |D - fp64| = %.3e (sequential fp32 FMA chain: %.3e), mean signed error %.3e\n", nterms, maxe,
                   maxf, bias / (M * N));
        }
    }
    printf("done\n");
    return 0;
}
