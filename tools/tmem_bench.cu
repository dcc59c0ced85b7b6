// Microbenchmark: tensor memory (TMEM) as per-thread scratch via tcgen05.st / tcgen05.ld versus the
// same access pattern from shared memory.  Pattern = the column pass of pair_fused_kernel: every
// thread of a 640-thread CTA reads 49 floats it owns and runs a dependent FMA chain over them.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tools/tmem_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#define R8(a, o) "%" #o "0"
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

constexpr int THREADS = 640;
constexpr int LD = THREADS + 1;

// mode 0: TMEM x32+x16+x1; mode 1: TMEM x32+x32 (64 cols); mode 2: shared memory [49][641]
__global__ void __launch_bounds__(THREADS, 1) bench(int mode, int iters, float* out, long long* cycles, int* bad) {
    extern __shared__ float Ksm[];
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tbase, 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 64);
    // fill: value(thread, j) = tid * 64 + j
    float v[64];
#pragma unroll
    for (int j = 0; j < 64; j++) v[j] = (float)(tid * 64 + j) * 1e-3f;
    tmem_st32(taddr, v);
    tmem_st32(taddr + 32, v + 32);
    tmem_wait_st();
    for (int j = 0; j < 49; j++) Ksm[j * LD + tid] = v[j];
    __syncthreads();
    // verify
    {
        float w[64];
        tmem_ld32(taddr, w);
        tmem_ld32(taddr + 32, w + 32);
        tmem_wait_ld();
        int b = 0;
#pragma unroll
        for (int j = 0; j < 64; j++) b += (w[j] != v[j]);
        if (b) atomicAdd(bad, b);
    }
    __syncthreads();
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        float x = acc * 1e-9f;
        if (mode == 0) {
            float k[32];
            tmem_ld32(taddr, k);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; j++) x = fmaf(k[j], 1.0001f, x);
            tmem_ld16(taddr + 32, k);
            tmem_ld1(taddr + 48, k + 16);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 17; j++) x = fmaf(k[j], 1.0001f, x);
        } else if (mode == 1) {
            float k[32];
            tmem_ld32(taddr, k);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; j++) x = fmaf(k[j], 1.0001f, x);
            tmem_ld32(taddr + 32, k);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; j++) x = fmaf(k[j], 1.0001f, x);
        } else if (mode == 2) {
#pragma unroll
            for (int j = 0; j < 49; j++) x = fmaf(Ksm[j * LD + tid], 1.0001f, x);
        } else if (mode == 3 || mode == 4) {   // LDS.128 broadcast of 52 floats: 1 address per warp / per-pair address (49 threads)
            const float4* p4 = reinterpret_cast<const float4*>(Ksm + (mode == 3 ? warp : tid / 49) * 52);
#pragma unroll
            for (int j = 0; j < 13; j++) { float4 q = p4[j]; x = fmaf(q.x, 1.0001f, x); x = fmaf(q.y, 1.0001f, x); x = fmaf(q.z, 1.0001f, x); x = fmaf(q.w, 1.0001f, x); }
        } else if (mode == 5 || mode == 6) {   // LDS.64 broadcast
            const float2* p2 = reinterpret_cast<const float2*>(Ksm + (mode == 5 ? warp : tid / 49) * 52);
#pragma unroll
            for (int j = 0; j < 26; j++) { float2 q = p2[j]; x = fmaf(q.x, 1.0001f, x); x = fmaf(q.y, 1.0001f, x); }
        } else {                               // LDS.32 broadcast, per-pair address
            const float* p1 = Ksm + (mode == 7 ? warp : tid / 49) * 52;
#pragma unroll
            for (int j = 0; j < 52; j++) x = fmaf(p1[j], 1.0001f, x);
        }
        acc += x;
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * THREADS + tid] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    const int blocks = 148, iters = 2000;
    float* out;
    long long* cyc;
    int* bad;
    CHECK(cudaMalloc(&out, blocks * THREADS * 4));
    CHECK(cudaMalloc(&cyc, blocks * 8));
    CHECK(cudaMalloc(&bad, 4));
    CHECK(cudaMemset(bad, 0, 4));
    const size_t smem = 49 * LD * 4;
    CHECK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char* names[9] = {"TMEM x32+x16+x1", "TMEM x32+x32", "shared [49][641]", "LDS.128 bcast warp", "LDS.128 bcast pair", "LDS.64 bcast warp", "LDS.64 bcast pair", "LDS.32 bcast warp", "LDS.32 bcast pair"};
    for (int rep = 0; rep < 1; rep++)
        for (int mode = 0; mode < 9; mode++) {
            bench<<<blocks, THREADS, smem>>>(mode, iters, out, cyc, bad);
            CHECK(cudaDeviceSynchronize());
            long long h[blocks];
            int hb;
            CHECK(cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost));
            CHECK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
            double mean = 0;
            for (int i = 0; i < blocks; i++) mean += (double)h[i] / iters;
            printf("%-18s: %.1f cycles per iteration per CTA (640 threads x 49 floats = %.0f B) -> %.1f B/clk/SM; readback mismatches %d\n",
                   names[mode], mean / blocks, 640.0 * 49 * 4, 640.0 * 49 * 4 / (mean / blocks), hb);
        }
    printf("done\n");
    return 0;
}
