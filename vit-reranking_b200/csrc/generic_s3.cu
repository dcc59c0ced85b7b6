// S3 for the shape-generic path on the tensor cores: sim[s][m] = sum_c F[c][s] * A[c][m] (utilities/diml.py:100) and the Gibbs
// kernel K = exp(-(1 - sim) / ot_temp) (:101-102) of every query / candidate pair, for any C % 16 == 0 and R <= 256 -- the
// ViT-B/16 shape of BASELINE.json configs[4] (C = 768, R = 196: 59 Mflop per pair) in particular.  It replaces the scalar fp32
// tile loop of generic_prepare_kernel, which spent ~85 % of a ViT-B/16 pass there.
//
// One CTA per pair, the recipe of pair_fused.cu's converter path at a larger scale: per 16-channel chunk the eight warps load
// the candidate's and the query's rows (coalesced along the patches), split them as 64 x = hi + lo in fp16 and store the
// K-major core-matrix operand tiles (A: up to 2 M-tiles of 128 candidate patches, B: the query patches padded to a multiple
// of 16); warp 7 / lane 0 issues 3 tcgen05.mma per M-tile (lo.hi + hi.lo + hi.hi, fp32 accumulation in tensor memory) and commits;
// two operand stages, `ready` / `mma_done` mbarriers hand them back and forth, the loads of the next chunk are in flight while
// this one is converted.  Read-out: every warp takes 32 accumulator rows, transposes 32-column pieces through a padded
// shared-memory tile and writes sim and K with full 128-byte rows.  |sim - fp32 chain| ~ 3e-7 (tools/umma_test.cu).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace vr {

constexpr int G3_THREADS = 512;
constexpr int G3_CONV_WARPS = 14;   // warps 0..13 convert (one (row, channel octet) item per thread), 14 = TMA producer, 15 = MMA issuer
constexpr float G3_SCALE = 64.0f;
constexpr int G3_NS = 4;   // raw-row staging ring (TMA): chunks in flight ahead of the converters

struct G3Args {
    const float* q_patches;
    const float* c_patches;
    const int32_t* cand_idx;
    int cand_stride;
    int64_t q_start, q_stride;
    int k, c, r, re, mt, rp16;
    float ot_temp;
    float* sim;   // [np, r, r]
    float* K;     // [np, re, re]
};

__device__ __forceinline__ void g3_split(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const float y0 = x0 * G3_SCALE, y1 = x1 * G3_SCALE;
    const __half2 h = __floats2half2_rn(y0, y1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// TMA = true (R % 4 == 0, 16-byte aligned banks): the 16 channel rows of a chunk are contiguous in both banks (12.5 KB at
// R = 196), so one thread fetches them with two cp.async.bulk per chunk into a 4-deep staging ring, three chunks ahead of the
// converters -- register loads one chunk ahead left the CTA (one per SM: its accumulators take 416 of the 512 TMEM columns)
// waiting a full HBM latency per chunk (83 us per pair; 2.1 TB/s over the GPU).
template <bool TMA>
__global__ void __launch_bounds__(G3_THREADS, 1) generic_sim_mma_kernel(G3Args a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t pair = blockIdx.x;
    const int64_t qi = pair / a.k;
    const int pi = (int)(pair % a.k);
    const int64_t qid = a.q_start + qi * a.q_stride;
    const int cand = a.cand_idx ? a.cand_idx[qi * a.cand_stride + pi] : pi;
    if (cand < 0) return;   // padded shortlist entry: generic_prepare_kernel writes the zero problem
    const int C = a.c, R = a.r, MT = a.mt, RP = a.rp16;
    // operand stage: [A hi | A lo | B hi | B lo]; A plane = [kcore 2][row group MT*16][128 B], B plane = [kcore 2][RP/8][128 B]
    const uint32_t planeA = (uint32_t)MT * 4096u, planeB = (uint32_t)RP * 32u;
    const uint32_t stage_bytes = 2u * planeA + 2u * planeB;
    unsigned char* stages = smem_raw;
    float* tile = reinterpret_cast<float*>(smem_raw + 2 * stage_bytes) + warp * (32 * 33);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * stage_bytes + 16 * 32 * 33 * 4);
    uint64_t* ready = bars;          // [2] operand stage stored by all warps
    uint64_t* mma_done = bars + 2;   // [2] MMAs of the stage completed
    uint64_t* s3_done = bars + 4;    // [1]
    uint64_t* full = bars + 5;       // [NS] raw rows of a chunk have landed
    uint64_t* empty = full + G3_NS;  // [NS] every warp has taken them into registers
    uint32_t* tmem_base = reinterpret_cast<uint32_t*>(empty + G3_NS);
    float* raw = reinterpret_cast<float*>(smem_raw + 2 * stage_bytes + 16 * 32 * 33 * 4 + 256);   // [NS][2][16 * R]
    const uint32_t raw_half = (uint32_t)(16 * R) * 4u;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(ready + i, G3_CONV_WARPS);
            mbar_init(mma_done + i, 1);
        }
        mbar_init(s3_done, 1);
        for (int i = 0; i < G3_NS; i++) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, G3_CONV_WARPS);
        }
        fence_mbar_init();
    }
    int ncols = 32;
    while (ncols < MT * RP) ncols <<= 1;
    if (warp == 0) tmem_alloc(tmem_base, (uint32_t)ncols);
    // rows beyond R of both operands stay zero for the whole kernel
    for (uint32_t e = tid; e < 2 * stage_bytes / 16; e += G3_THREADS) reinterpret_cast<uint4*>(stages)[e] = make_uint4(0u, 0u, 0u, 0u);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tmem0 = *tmem_base;

    const float* Fg = a.c_patches + (int64_t)cand * C * R;
    const float* Ag = a.q_patches + qid * (int64_t)C * R;
    // (kcore g, row s) items per operand: 2 R <= 448 = one per converter thread
    const int nitem = 2 * R;
    const bool conv = warp < G3_CONV_WARPS;
    const bool it_ok = conv && tid < nitem;
    const int it_g = it_ok ? tid / R : 0;
    const int it_s = it_ok ? tid - it_g * R : 0;
    const int NCH = C / 16;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(RP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t st_addr = smem_u32(stages);
    if (warp == G3_CONV_WARPS) {
        // ---- producer (TMA only): both banks' 16 rows of every chunk -> staging ring, NS - 1 chunks ahead of the converters ----
        if (TMA && lane == 0) {
            for (int ch = 0; ch < NCH; ch++) {
                const int stg = ch % G3_NS;
                if (ch >= G3_NS) mbar_wait(empty + stg, ((ch / G3_NS) - 1) & 1);
                mbar_expect_tx(full + stg, 2 * raw_half);
                bulk_g2s(raw + (size_t)stg * 2 * 16 * R, Fg + (int64_t)ch * 16 * R, raw_half, full + stg);
                bulk_g2s(raw + (size_t)stg * 2 * 16 * R + 16 * R, Ag + (int64_t)ch * 16 * R, raw_half, full + stg);
            }
        }
    } else if (warp == G3_CONV_WARPS + 1) {
        // ---- MMA issuer ----
        if (lane == 0) {
            for (int ch = 0; ch < NCH; ch++) {
                const int os = ch & 1;
                mbar_wait(ready + os, (ch >> 1) & 1);
                tmem_fence_after();
                const uint32_t base = st_addr + (uint32_t)os * stage_bytes;
                const uint64_t bhd = umma_desc(base + 2 * planeA, planeB / 2, 128);
                const uint64_t bld = umma_desc(base + 2 * planeA + planeB, planeB / 2, 128);
                for (int t = 0; t < MT; t++) {
                    const uint64_t ahd = umma_desc(base + (uint32_t)t * 2048u, planeA / 2, 128);
                    const uint64_t ald = umma_desc(base + planeA + (uint32_t)t * 2048u, planeA / 2, 128);
                    const uint32_t d = tmem0 + (uint32_t)(t * RP);
                    umma_f16_i(d, ald, bhd, idesc, ch > 0 ? 1u : 0u);   // small terms first
                    umma_f16_i(d, ahd, bld, idesc, 1u);
                    umma_f16_i(d, ahd, bhd, idesc, 1u);
                }
                umma_commit(smem_u32(mma_done + os));
                if (ch == NCH - 1) umma_commit(smem_u32(s3_done));
            }
        }
    } else {
        // ---- converters ----
        float xa[8], xb[8];
        auto load_chunk = [&](int ch) {
            if (TMA) {
                const int stg = ch % G3_NS;
                mbar_wait(full + stg, (ch / G3_NS) & 1);
                const float* ra = raw + (size_t)stg * 2 * 16 * R;
                const float* rb = ra + 16 * R;
                const int off = 8 * it_g * R + it_s;
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    xa[e] = it_ok ? ra[off + e * R] : 0.f;
                    xb[e] = it_ok ? rb[off + e * R] : 0.f;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(empty + stg));
                return;
            }
            const int64_t off = (int64_t)(ch * 16 + 8 * it_g) * R + it_s;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                xa[e] = it_ok ? __ldg(Fg + off + (int64_t)e * R) : 0.f;
                xb[e] = it_ok ? __ldg(Ag + off + (int64_t)e * R) : 0.f;
            }
        };
        const uint32_t offA = (uint32_t)it_g * (planeA / 2) + (uint32_t)(it_s >> 3) * 128u + (uint32_t)(it_s & 7) * 16u;
        const uint32_t offB = (uint32_t)it_g * (planeB / 2) + (uint32_t)(it_s >> 3) * 128u + (uint32_t)(it_s & 7) * 16u;
        load_chunk(0);
#pragma unroll 1
        for (int ch = 0; ch < NCH; ch++) {
            const int os = ch & 1;
            uint32_t ah[4], al[4], bh[4], bl[4];
#pragma unroll
            for (int w = 0; w < 4; w++) {
                g3_split(xa[2 * w], xa[2 * w + 1], ah[w], al[w]);
                g3_split(xb[2 * w], xb[2 * w + 1], bh[w], bl[w]);
            }
            if (ch + 1 < NCH) load_chunk(ch + 1);
            if (ch > 1) mbar_wait(mma_done + os, ((ch >> 1) - 1) & 1);   // the stage is free once the MMAs of chunk ch - 2 are done
            unsigned char* sg = stages + os * stage_bytes;
            if (it_ok) {
                *reinterpret_cast<uint4*>(sg + offA) = make_uint4(ah[0], ah[1], ah[2], ah[3]);
                *reinterpret_cast<uint4*>(sg + planeA + offA) = make_uint4(al[0], al[1], al[2], al[3]);
                *reinterpret_cast<uint4*>(sg + 2 * planeA + offB) = make_uint4(bh[0], bh[1], bh[2], bh[3]);
                *reinterpret_cast<uint4*>(sg + 2 * planeA + planeB + offB) = make_uint4(bl[0], bl[1], bl[2], bl[3]);
            }
            fence_proxy_async();   // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(ready + os));
        }
    }
    mbar_wait(s3_done, 0);
    tmem_fence_after();

    // ---- read-out: warp w owns accumulator rows 32 (w % 4) .. + 31 of M-tile (w / 4) % 2 and column half w / 8 ----
    const int t = (warp >> 2) & 1;
    const int chalf = warp >> 3;
    if (t < MT) {
        const int row0 = t * 128 + 32 * (warp & 3);
        const uint32_t tl = tmem0 + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(t * RP);
        float* simo = a.sim + pair * (int64_t)R * R;
        float* Ko = a.K + pair * (int64_t)a.re * a.re;
        constexpr float dscale = 1.0f / (G3_SCALE * G3_SCALE);
        const float ot = a.ot_temp;
        if (row0 < R) {
            const int cmid = ((RP / 32 + 1) / 2) * 32;   // columns [0, cmid) to half 0, [cmid, RP) to half 1
            for (int c0 = chalf ? cmid : 0; c0 < (chalf ? RP : cmid); c0 += 32) {
                uint32_t v[32];
                if (c0 + 32 <= RP) {
                    tmem_ld32(tl + (uint32_t)c0, v);
                } else {   // RP is a multiple of 16: a 16-column tail
                    tmem_ld16(tl + (uint32_t)c0, v);
#pragma unroll
                    for (int i = 16; i < 32; i++) v[i] = 0u;
                }
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; i++) tile[lane * 33 + i] = __uint_as_float(v[i]) * dscale;
                __syncwarp();
                const int m = c0 + lane;
                for (int rr = 0; rr < 32; rr++) {
                    const int s = row0 + rr;
                    if (s < R && m < R) {
                        const float x = tile[rr * 33 + lane];
                        simo[(int64_t)s * R + m] = x;
                        Ko[(int64_t)s * a.re + m] = expf(-(1.0f - x) / ot);
                    }
                }
                __syncwarp();
            }
        }
    }
    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, (uint32_t)ncols);
}

// ---- the same S3 from a RE-PACKED bank: no conversion per pair ----
// For a registered gallery the hi / lo halves of every image and 16-channel chunk are derived once (generic_repack_kernel) and
// kept as one block [plane hi | lo][k-core 2][row group RP / 8][128 B] per (image, chunk): RP * 64 bytes, 13,312 at R = 196 --
// 639 KB per image at C = 768, the size of the fp32 rows.  That block IS the B operand of the stage above; as the A operand its
// four (plane, k-core) runs go to the heads of the stage's four A runs (the row groups beyond RP / 8 stay zero).  A pair then
// needs five cp.async.bulk per chunk straight into one of four operand stages; one thread feeds, one issues, nobody converts.
struct G3PArgs {
    const unsigned char* packed;   // [n][C / 16][stage_bytes]
    const int32_t* cand_idx;
    int cand_stride;
    int64_t q_start, q_stride;
    int k, c, r, re, mt, rp16;
    float ot_temp;
    float* sim;
    float* K;
};

__global__ void __launch_bounds__(256) generic_repack_kernel(const float* __restrict__ patches, const float* __restrict__ centers,
                                                             int64_t first, int c, int r, int rp16, unsigned char* __restrict__ packed) {
    extern __shared__ float rp_cn[];   // [c]: the image's centre, L2-normalised (+ 32 floats for the reduction)
    const int64_t im = first + blockIdx.x;
    const float* F = patches + im * (int64_t)c * r;
    const uint32_t run = (uint32_t)rp16 * 16u;   // one (plane, k-core) run
    const int nch = c / 16;
    unsigned char* out = packed + im * (int64_t)nch * (4u * run);
    // (rows r .. rp16 - 1 were zeroed by the caller's memset)
    for (int item = threadIdx.x; item < nch * 2 * r; item += 256) {
        const int ch = item / (2 * r), rem = item - ch * 2 * r, g = rem / r, s = rem - g * r;
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = __ldg(F + (int64_t)(ch * 16 + 8 * g + e) * r + s);
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int w = 0; w < 4; w++) g3_split(x[2 * w], x[2 * w + 1], hi[w], lo[w]);
        unsigned char* blk = out + (size_t)ch * (4u * run) + (uint32_t)g * run + (uint32_t)(s >> 3) * 128u + (uint32_t)(s & 7) * 16u;
        *reinterpret_cast<uint4*>(blk) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(blk + 2u * run) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    // Patch R (a spare padding row when R % 16 != 0): the image's normalised centre.  In the MMA it turns row R of sim into
    // <candidate centre, query patches> and column R into <query centre, candidate patches> -- the cross-correlations of
    // utilities/diml.py:104-133 with use_cls_token -- at no cost (generic_fused.cu).
    if (centers && r < rp16) {
        float* red = rp_cn + c;
        const float* g = centers + im * (int64_t)c;
        float ng = 0.f;
        for (int cc = threadIdx.x; cc < c; cc += 256) {
            const float x = g[cc];
            rp_cn[cc] = x;
            ng += x * x;
        }
        {   // generic_prepare_kernel's block sum
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            ng = warp_sum(ng);
            __syncthreads();
            if (lane == 0) red[warp] = ng;
            __syncthreads();
            ng = 0.f;
            for (int i = 0; i < 8; i++) ng += red[i];
        }
        const float dg = fmaxf(sqrtf(ng), 1e-12f);
        for (int item = threadIdx.x; item < nch * 2; item += 256) {
            const int ch = item >> 1, g2 = item & 1;
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; e++) x[e] = rp_cn[ch * 16 + 8 * g2 + e] / dg;
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int w = 0; w < 4; w++) g3_split(x[2 * w], x[2 * w + 1], hi[w], lo[w]);
            unsigned char* blk = out + (size_t)ch * (4u * run) + (uint32_t)g2 * run + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
            *reinterpret_cast<uint4*>(blk) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(blk + 2u * run) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

constexpr int G3P_NS = 4;   // operand stages

__global__ void __launch_bounds__(G3_THREADS, 1) generic_sim_mma_packed_kernel(G3PArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t pair = blockIdx.x;
    const int64_t qi = pair / a.k;
    const int pi = (int)(pair % a.k);
    const int64_t qid = a.q_start + qi * a.q_stride;
    const int cand = a.cand_idx ? a.cand_idx[qi * a.cand_stride + pi] : pi;
    if (cand < 0) return;
    const int C = a.c, R = a.r, MT = a.mt, RP = a.rp16;
    const uint32_t planeA = (uint32_t)MT * 4096u, planeB = (uint32_t)RP * 32u;
    const uint32_t stage_bytes = 2u * planeA + 2u * planeB;
    unsigned char* stages = smem_raw;
    float* tile = reinterpret_cast<float*>(smem_raw + G3P_NS * stage_bytes) + warp * (32 * 33);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + G3P_NS * stage_bytes + 16 * 32 * 33 * 4);
    uint64_t* full = bars;                 // [NS] both halves of a chunk have landed
    uint64_t* mma_done = bars + G3P_NS;    // [NS] the MMAs that read the stage have completed
    uint64_t* s3_done = bars + 2 * G3P_NS;
    uint32_t* tmem_base = reinterpret_cast<uint32_t*>(bars + 2 * G3P_NS + 1);
    if (tid == 0) {
        for (int i = 0; i < G3P_NS; i++) {
            mbar_init(full + i, 1);
            mbar_init(mma_done + i, 1);
        }
        mbar_init(s3_done, 1);
        fence_mbar_init();
    }
    int ncols = 32;
    while (ncols < MT * RP) ncols <<= 1;
    if (warp == 0) tmem_alloc(tmem_base, (uint32_t)ncols);
    {   // the A row groups beyond RP / 8 are never copied: zero them in every stage, once
        const uint32_t runA = planeA / 2, padv = (runA - (uint32_t)RP * 16u) / 16u;
        for (uint32_t i = tid; i < G3P_NS * 4u * padv; i += G3_THREADS) {
            const uint32_t v = i % padv, pk = (i / padv) & 3u, stg = i / (4u * padv);
            *reinterpret_cast<uint4*>(stages + (size_t)stg * stage_bytes + (pk >> 1) * planeA + (pk & 1) * runA + (uint32_t)RP * 16u + v * 16u) =
                make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
    }
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tmem0 = *tmem_base;
    const int NCH = C / 16;
    const uint32_t run = (uint32_t)RP * 16u;
    const unsigned char* Ap = a.packed + (int64_t)cand * NCH * (4u * run);
    const unsigned char* Bp = a.packed + qid * (int64_t)NCH * (4u * run);
    if (warp == G3_CONV_WARPS) {
        if (lane == 0) {
            for (int ch = 0; ch < NCH; ch++) {
                const int stg = ch % G3P_NS;
                if (ch >= G3P_NS) mbar_wait(mma_done + stg, ((ch / G3P_NS) - 1) & 1);
                mbar_expect_tx(full + stg, 8u * run);
                unsigned char* dst = stages + (size_t)stg * stage_bytes;
#pragma unroll
                for (int pk = 0; pk < 4; pk++)   // (plane, k-core) runs of the candidate -> heads of the A runs
                    bulk_g2s(dst + (uint32_t)(pk >> 1) * planeA + (uint32_t)(pk & 1) * (planeA / 2), Ap + (size_t)ch * (4u * run) + (uint32_t)pk * run,
                             run, full + stg);
                bulk_g2s(dst + 2 * planeA, Bp + (size_t)ch * (4u * run), 4u * run, full + stg);
            }
        }
    } else if (warp == G3_CONV_WARPS + 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(RP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t st_addr = smem_u32(stages);
            for (int ch = 0; ch < NCH; ch++) {
                const int stg = ch % G3P_NS;
                mbar_wait(full + stg, (ch / G3P_NS) & 1);
                tmem_fence_after();
                const uint32_t base = st_addr + (uint32_t)stg * stage_bytes;
                const uint64_t bhd = umma_desc(base + 2 * planeA, planeB / 2, 128);
                const uint64_t bld = umma_desc(base + 2 * planeA + planeB, planeB / 2, 128);
                for (int t = 0; t < MT; t++) {
                    const uint64_t ahd = umma_desc(base + (uint32_t)t * 2048u, planeA / 2, 128);
                    const uint64_t ald = umma_desc(base + planeA + (uint32_t)t * 2048u, planeA / 2, 128);
                    const uint32_t d = tmem0 + (uint32_t)(t * RP);
                    umma_f16_i(d, ald, bhd, idesc, ch > 0 ? 1u : 0u);   // small terms first
                    umma_f16_i(d, ahd, bld, idesc, 1u);
                    umma_f16_i(d, ahd, bhd, idesc, 1u);
                }
                umma_commit(smem_u32(mma_done + stg));
                if (ch == NCH - 1) umma_commit(smem_u32(s3_done));
            }
        }
    }
    mbar_wait(s3_done, 0);
    tmem_fence_after();
    // ---- read-out (as in generic_sim_mma_kernel) ----
    const int t = (warp >> 2) & 1;
    const int chalf = warp >> 3;
    if (t < MT) {
        const int row0 = t * 128 + 32 * (warp & 3);
        const uint32_t tl = tmem0 + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(t * RP);
        float* simo = a.sim + pair * (int64_t)R * R;
        float* Ko = a.K + pair * (int64_t)a.re * a.re;
        constexpr float dscale = 1.0f / (G3_SCALE * G3_SCALE);
        const float ot = a.ot_temp;
        if (row0 < R) {
            const int cmid = ((RP / 32 + 1) / 2) * 32;
            for (int c0 = chalf ? cmid : 0; c0 < (chalf ? RP : cmid); c0 += 32) {
                uint32_t v[32];
                if (c0 + 32 <= RP) {
                    tmem_ld32(tl + (uint32_t)c0, v);
                } else {
                    tmem_ld16(tl + (uint32_t)c0, v);
#pragma unroll
                    for (int i = 16; i < 32; i++) v[i] = 0u;
                }
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; i++) tile[lane * 33 + i] = __uint_as_float(v[i]) * dscale;
                __syncwarp();
                const int m = c0 + lane;
                for (int rr = 0; rr < 32; rr++) {
                    const int s = row0 + rr;
                    if (s < R && m < R) {
                        const float x = tile[rr * 33 + lane];
                        simo[(int64_t)s * R + m] = x;
                        Ko[(int64_t)s * a.re + m] = expf(-(1.0f - x) / ot);
                    }
                }
                __syncwarp();
            }
        }
    }
    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, (uint32_t)ncols);
}

size_t generic_packed_image_bytes(int c, int r) {
    const size_t rp16 = (r + 15) / 16 * 16;
    return (size_t)(c / 16) * (rp16 * 64);
}

int generic_repack(const float* patches, const float* centers, int64_t n, int c, int r, void* packed, cudaStream_t st) {
    const int rp16 = (r + 15) / 16 * 16;
    VR_CHECK_CUDA(cudaMemsetAsync(packed, 0, (size_t)n * generic_packed_image_bytes(c, r), st));
    generic_repack_kernel<<<(unsigned)n, 256, (size_t)(c + 32) * 4, st>>>(patches, centers, 0, c, r, rp16, reinterpret_cast<unsigned char*>(packed));
    VR_LAUNCH_CHECK();
    return VR_OK;
}

bool generic_sim_mma_supported(int c, int r) {
    const char* e = getenv("VR_GENERIC_S3");   // VR_GENERIC_S3=fp32 keeps the scalar loop (A/B tests)
    if (e && e[0] == 'f') return false;
    return c >= 16 && c % 16 == 0 && r >= 8 && 2 * r <= G3_CONV_WARPS * 32;   // (one converter item per thread: R <= 224)
}

int generic_sim_mma(const GenArgs& g, int re, cudaStream_t st) {
    if (g.packed) {   // both roles of every image pre-split (registered bank): TMA + MMA only
        G3PArgs p{};
        p.packed = reinterpret_cast<const unsigned char*>(g.packed);
        p.cand_idx = g.cand_idx;
        p.cand_stride = g.cand_stride;
        p.q_start = g.q_start;
        p.q_stride = g.q_stride;
        p.k = g.k;
        p.c = g.c;
        p.r = g.r;
        p.re = re;
        p.mt = (g.r + 127) / 128;
        p.rp16 = (g.r + 15) / 16 * 16;
        p.ot_temp = g.p.ot_temp;
        p.sim = g.sim;
        p.K = g.K;
        const size_t stage = 2 * (size_t)p.mt * 4096 + 2 * (size_t)p.rp16 * 32;
        const size_t smem = G3P_NS * stage + 16 * 32 * 33 * 4 + 256;
        VR_CHECK_CUDA(cudaFuncSetAttribute(generic_sim_mma_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        generic_sim_mma_packed_kernel<<<(unsigned)(g.nq * g.k), G3_THREADS, smem, st>>>(p);
        VR_LAUNCH_CHECK();
        return VR_OK;
    }
    G3Args a{};
    a.q_patches = g.q_patches;
    a.c_patches = g.c_patches;
    a.cand_idx = g.cand_idx;
    a.cand_stride = g.cand_stride;
    a.q_start = g.q_start;
    a.q_stride = g.q_stride;
    a.k = g.k;
    a.c = g.c;
    a.r = g.r;
    a.re = re;
    a.mt = (g.r + 127) / 128;
    a.rp16 = (g.r + 15) / 16 * 16;
    a.ot_temp = g.p.ot_temp;
    a.sim = g.sim;
    a.K = g.K;
    const size_t stage = 2 * (size_t)a.mt * 4096 + 2 * (size_t)a.rp16 * 32;
    const size_t raw = (size_t)G3_NS * 2 * 16 * g.r * 4;
    const bool tma = g.r % 4 == 0 && ((uintptr_t)g.q_patches & 15) == 0 && ((uintptr_t)g.c_patches & 15) == 0 &&
                     2 * stage + 16 * 32 * 33 * 4 + 256 + raw <= 225 * 1024;
    const size_t smem = 2 * stage + 16 * 32 * 33 * 4 + 256 + (tma ? raw : 0);
    auto* kern = tma ? generic_sim_mma_kernel<true> : generic_sim_mma_kernel<false>;
    VR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)(g.nq * g.k), G3_THREADS, smem, st>>>(a);
    VR_LAUNCH_CHECK();
    return VR_OK;
}

}  // namespace vr
