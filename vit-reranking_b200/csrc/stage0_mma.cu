// S1 on the tensor cores: first-stage retrieval as a tcgen05 GEMM over the 128-d global embeddings with the top-K' select
// fused into the accumulator read-out, so that the N x N score matrix exists only in tensor memory.
// Replaces calc_similarity(stage=0) (utilities/diml.py:83-85), the self mask (evaluation/eval_cvt_diml.py:327) and the
// head of the full argsort (:329-332) for batches of queries; results are BIT-IDENTICAL to the fp32 path of
// stage0_topk.cu (same shortlists, same order, same fp32 scores), which stays the path for small batches, other embedding
// widths and long shortlists.
//
// How a tensor-core product can return the fp32 path's lists.  Operands are split as 64 x = hi + lo in fp16 and
// score ~ (hi.hi + lo.hi + hi.lo) / 4096 with fp32 accumulation (the S3 recipe of pair_fused.cu): |approx - fp32 chain| <= EPS
// = 2e-5 |q| max|g| (worst-case bound; typical 3e-7).  The approximate scores only SHORTLIST: per query the best kp + 33 by
// approximate score are re-scored with the canonical fp32 FMA chain (sequential over the channels, the arithmetic of
// stage0_topk.cu), re-sorted, and the list is accepted when its kp-th canonical score exceeds the best approximate score left
// outside by more than EPS -- then no outsider can belong to (or tie into) the true list.  Rows that fail the test, overflow
// their candidate buffer or hold non-finite / out-of-range values are redone by an exact fp32 fallback kernel; their count is
// reported (vr_stage0_stats).
//
// Four launches + one memset per call, no host synchronisation:
//   pack        centres -> fp16 hi / lo planes in the UMMA K-major core-matrix layout (gallery: 256-row tiles, queries: 128-row
//               blocks), row norms, range flags.  31 MB each way at SOP scale: ~20 us, so it is redone per call, not cached;
//   pass 1      GEMM; read-out keeps only the maximum of every L consecutive scores of a row ("segment maxima").  The
//               (kp + 34)-th largest segment maximum T of a row is a lower bound of its (kp + 34)-th best score (those maxima are
//               distinct scores) -- found by a 32-step bisection on the ordered bit patterns (thresh kernel);
//   pass 2      the same GEMM again (2 x 0.94 Tflop at SOP: cheaper than keeping 14.6 GB of scores); read-out appends the few
//               scores >= T (~1.3 x need for shuffled data, more when classes are contiguous) to the row's candidate buffer;
//   final       one warp per row: sort candidates by approximate score, canonical re-score, test, write-out / fail list;
//   fallback    persistent CTAs walk the fail list: canonical scores of the whole row into a scratch row, warp row-select.
//
// GEMM kernel: one CTA = 128 queries x a range of 256-column gallery tiles.  Warp 0 = TMA producer (one 32 KB cp.async.bulk per
// stage: the packed bank stores a tile's stages contiguously), warp 1 = MMA issuer (24 tcgen05.mma M128 N256 K16 per tile into
// one of two 256-column accumulators), warps 2..5 = read-out (thread = TMEM lane = query row: every thread filters ITS OWN
// query's scores, no cross-thread traffic at all).  A operand (64 KB) resident per CTA, B ring 4 x 32 KB.
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "stage0_select.cuh"
#include "umma.cuh"

namespace vr {

typedef unsigned long long ull;

constexpr int M0_C = 128;                     // embedding width this path is built for
constexpr int M0_BM = 128;                    // queries per CTA (= TMEM lanes)
constexpr int M0_BN = 256;                    // gallery columns per tile (= MMA N)
constexpr int M0_NCH = M0_C / 16;             // 8 K-chunks of 16 channels
constexpr int M0_STAGE_CH = 2;                // chunks per B stage
constexpr int M0_SPT = M0_NCH / M0_STAGE_CH;  // 4 stages per tile
constexpr int M0_NST = 4;                     // B ring depth
constexpr int M0_PLANE_B = 2 * 32 * 128;      // bytes of one (chunk, plane) of a gallery tile: [kcore 2][row group 32][128 B]
constexpr int M0_PLANE_A = 2 * 16 * 128;      // ... of a query block
constexpr int M0_STAGE_BYTES = M0_STAGE_CH * 2 * M0_PLANE_B;   // 32,768
constexpr int M0_TILE_BYTES = M0_NCH * 2 * M0_PLANE_B;         // 131,072 per 256 gallery rows
constexpr int M0_ABLK_BYTES = M0_NCH * 2 * M0_PLANE_A;         // 65,536 per 128 queries
constexpr int M0_EW2 = 4;                     // read-out warps per TMEM lane quarter in pass 2 (pass 1: one)
constexpr float M0_SCALE = 64.0f;             // operands are scaled by 64 before the fp16 split, products by 1/4096
constexpr float M0_INV = 1.0f / 4096.0f;
constexpr int M0_SLACK = 12;                  // candidates re-scored beyond kp
constexpr float M0_EPS = 2e-5f;               // bound on |tensor-core score - fp32 chain| for unit-norm rows
constexpr size_t M0_SMEM = 1024 + (size_t)M0_ABLK_BYTES + (size_t)M0_NST * M0_STAGE_BYTES + 256;
constexpr uint32_t M0_IDESC = (1u << 4) | ((uint32_t)(M0_BN >> 3) << 17) | ((uint32_t)(M0_BM >> 4) << 24);

// flags[0] = rows redone by the fallback, [1] = max |g|^2 (float bits), [2] = 1 if a value cannot be split (non-finite or
// |x| > 900), [3] = rows whose candidate buffer overflowed
struct S0MArgs {
    const unsigned char* gpk;
    const unsigned char* qpk;
    int64_t nq, n;
    int ntiles, tiles_per, nsplit, ngroups, S_per, S_total, cap;
    uint32_t* segmax;     // [nq][S_total] ordered score bits
    const uint32_t* thr;  // [nq] ordered bits of T
    uint32_t* cnt;        // [nq][nsplit][EW2] candidates each (split, read-out warp) found (may exceed cap: overflow)
    ull* buf;             // [nq][nsplit][EW2][cap]
};

__device__ __forceinline__ void split_f16x2_s0(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const float y0 = x0 * M0_SCALE, y1 = x1 * M0_SCALE;
    const __half2 h = __floats2half2_rn(y0, y1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// rows [0, rows_pad) of the packed operand: row i = src row (start + i * stride), zero beyond n_rows.
// RG = 8-row groups per tile (32: gallery tiles of 256 rows, 16: query blocks of 128 rows).  One thread per row.
template <int RG>
__global__ void stage0_pack_kernel(const float* __restrict__ src, int64_t n_rows, int64_t rows_pad, int64_t start, int64_t stride,
                                   unsigned char* __restrict__ out, uint32_t* __restrict__ flags, int is_gallery) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_pad) return;
    const int64_t tile = i / (8 * RG);
    const int rg = (int)((i % (8 * RG)) / 8), r8 = (int)(i % 8);
    unsigned char* base = out + tile * (int64_t)(M0_NCH * 2 * 2 * RG * 128) + rg * 128 + r8 * 16;
    const bool real = i < n_rows;
    const float4* s = reinterpret_cast<const float4*>(src + (real ? (start + i * stride) : 0) * M0_C);
    float n2 = 0.f;
    bool bad = false;
#pragma unroll 4
    for (int o = 0; o < 16; o++) {   // octet o = channels 8o .. 8o+7 = (chunk o >> 1, kcore o & 1)
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (real) {
            a = __ldg(s + 2 * o);
            b = __ldg(s + 2 * o + 1);
        }
        const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            split_f16x2_s0(x[2 * w], x[2 * w + 1], hi[w], lo[w]);
            n2 = fmaf(x[2 * w], x[2 * w], n2);
            n2 = fmaf(x[2 * w + 1], x[2 * w + 1], n2);
            bad |= !(fabsf(x[2 * w]) <= 900.f) | !(fabsf(x[2 * w + 1]) <= 900.f);
        }
        unsigned char* p = base + (size_t)((((o >> 1) * 2 + 0) * 2 + (o & 1)) * RG) * 128;
        *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(p + (size_t)2 * RG * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (bad) atomicOr(flags + 2, 1u);
    if (is_gallery && real && n2 == n2) atomicMax(flags + 1, __float_as_uint(n2));   // non-negative floats order like their bits
}

__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// PASS 1: segment maxima.  PASS 2: candidates >= T.  EW = read-out warps per TMEM lane quarter (a warp can only reach the 32
// lanes of quarter warp % 4): the EW warps of a quarter split the 256 columns of a tile.  Pass 1 is one FMNMX per score and
// runs at the tensor pipe's pace with EW = 1; pass 2 is a compare + branch per score, latency-bound with one warp per
// scheduler (7.0 ms at SOP), so it runs four (1.9 ms).
template <int PASS, int EW>
__global__ void __launch_bounds__(64 + 128 * EW, 1) stage0_mma_kernel(S0MArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* As = sm;
    unsigned char* Bs = sm + M0_ABLK_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Bs + M0_NST * M0_STAGE_BYTES);
    uint64_t* a_full = bars;                 // [1]
    uint64_t* b_full = bars + 1;             // [NST]
    uint64_t* b_empty = b_full + M0_NST;     // [NST]
    uint64_t* acc_full = b_empty + M0_NST;   // [2]
    uint64_t* acc_empty = acc_full + 2;      // [2]
    uint32_t* tmem_base = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int split = blockIdx.x, qb = blockIdx.y;
    const int tile_lo = split * a.tiles_per;
    const int ntl = max(0, min(a.ntiles, tile_lo + a.tiles_per) - tile_lo);

    if (tid == 0) {
        mbar_init(a_full, 1);
        for (int i = 0; i < M0_NST; i++) {
            mbar_init(b_full + i, 1);
            mbar_init(b_empty + i, 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(acc_full + i, 1);
            mbar_init(acc_empty + i, 4 * EW);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_base, 512);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tmem0 = *tmem_base;

    if (warp == 0) {
        if (lane == 0 && ntl > 0) {
            mbar_expect_tx(a_full, M0_ABLK_BYTES);
            bulk_g2s(As, a.qpk + (int64_t)qb * M0_ABLK_BYTES, M0_ABLK_BYTES, a_full);
            int it = 0;
            for (int g = 0; g < a.ngroups; g++)
                for (int tt = g; tt < ntl; tt += a.ngroups)
                    for (int s = 0; s < M0_SPT; s++, it++) {
                        const int st = it % M0_NST;
                        if (it >= M0_NST) mbar_wait(b_empty + st, ((it / M0_NST) - 1) & 1);
                        mbar_expect_tx(b_full + st, M0_STAGE_BYTES);
                        bulk_g2s(Bs + st * M0_STAGE_BYTES, a.gpk + (int64_t)(tile_lo + tt) * M0_TILE_BYTES + s * M0_STAGE_BYTES,
                                 M0_STAGE_BYTES, b_full + st);
                    }
        }
    } else if (warp == 1) {
        if (lane == 0 && ntl > 0) {
            const uint32_t a_addr = smem_u32(As), b_addr = smem_u32(Bs);
            mbar_wait(a_full, 0);
            for (int t = 0; t < ntl; t++) {
                const int buf = t & 1;
                if (t >= 2) mbar_wait(acc_empty + buf, ((t >> 1) - 1) & 1);
                tmem_fence_after();
                const uint32_t d = tmem0 + (uint32_t)(buf * M0_BN);
#pragma unroll
                for (int s = 0; s < M0_SPT; s++) {
                    const int it = t * M0_SPT + s, st = it % M0_NST;
                    mbar_wait(b_full + st, (it / M0_NST) & 1);
                    tmem_fence_after();
#pragma unroll
                    for (int lc = 0; lc < M0_STAGE_CH; lc++) {
                        const int ch = s * M0_STAGE_CH + lc;
                        const uint64_t ah = umma_desc(a_addr + (uint32_t)((ch * 2 + 0) * M0_PLANE_A), 16 * 128, 128);
                        const uint64_t al = umma_desc(a_addr + (uint32_t)((ch * 2 + 1) * M0_PLANE_A), 16 * 128, 128);
                        const uint64_t bh = umma_desc(b_addr + (uint32_t)(st * M0_STAGE_BYTES + (lc * 2 + 0) * M0_PLANE_B), 32 * 128, 128);
                        const uint64_t bl = umma_desc(b_addr + (uint32_t)(st * M0_STAGE_BYTES + (lc * 2 + 1) * M0_PLANE_B), 32 * 128, 128);
                        umma_f16_i(d, al, bh, M0_IDESC, ch > 0 ? 1u : 0u);   // small terms first
                        umma_f16_i(d, ah, bl, M0_IDESC, 1u);
                        umma_f16_i(d, ah, bh, M0_IDESC, 1u);
                    }
                    umma_commit(smem_u32(b_empty + st));
                }
                umma_commit(smem_u32(acc_full + buf));
            }
        }
    } else {
        // ---- read-out: thread = TMEM lane = query row ----
        // Tiles are visited group by group (group g = tiles g, g + G, g + 2G, .. of this split's range).  PASS 1 keeps 64
        // running maxima per thread, one per column residue mod 64, and stores them at the end of a group: segment (g, l) =
        // every column of group g with (column mod 64) == l.  Interleaving at both levels spreads a contiguous block of
        // same-class images (test sets are ordered by class) over many segments instead of letting it raise a few maxima.
        constexpr int PPW = (M0_BN / 32) / EW;   // 32-column pieces of a tile per warp
        const int lq = warp & 3, h = (warp - 2) >> 2, pc0 = h * PPW;
        const int64_t row = (int64_t)qb * M0_BM + 32 * lq + lane;
        const bool rowok = row < a.nq;
        const uint32_t tl = tmem0 + ((uint32_t)(32 * lq) << 16);
        float Ts = INFINITY;
        uint32_t* seg = nullptr;
        ull* mybuf = nullptr;
        if (PASS == 1) {
            seg = a.segmax + (rowok ? row : 0) * (int64_t)a.S_total + (int64_t)split * a.S_per;
        } else {
            if (rowok) Ts = from_ordered_bits(a.thr[row]) * 4096.0f;
            mybuf = a.buf + (((rowok ? row : 0) * (int64_t)a.nsplit + split) * EW + h) * (int64_t)a.cap;
        }
        uint32_t cnt = 0u;
        int t = 0;   // sequence number of the tile (accumulator buffer = t & 1)
        for (int g = 0; g < a.ngroups; g++) {
            static_assert(PASS == 2 || EW == 1, "pass 1 keeps the 64 column residues of a row in one thread");
            float m[64];
#pragma unroll
            for (int i = 0; i < 64; i++) m[i] = -INFINITY;
            for (int tt = g; tt < ntl; tt += a.ngroups, t++) {
                const int buf = t & 1;
                mbar_wait(acc_full + buf, (t >> 1) & 1);
                tmem_fence_after();
                const int64_t col0 = (int64_t)(tile_lo + tt) * M0_BN;
                const bool edge = col0 + M0_BN > a.n;     // tile holds padding columns (zero rows of the packed bank)
                uint32_t va[32], vb[32];
                tmem_ld32(tl + (uint32_t)(buf * M0_BN + pc0 * 32), va);
#pragma unroll
                for (int lp = 0; lp < PPW; lp++) {
                    const int pc = pc0 + lp;
                    uint32_t* cur = (lp & 1) ? vb : va;
                    uint32_t* nxt = (lp & 1) ? va : vb;
                    tmem_wait_ld();
                    if (lp + 1 < PPW) tmem_ld32(tl + (uint32_t)(buf * M0_BN + (pc + 1) * 32), nxt);
                    const int64_t c0 = col0 + pc * 32;
                    if (edge) {
#pragma unroll
                        for (int i = 0; i < 32; i++)
                            if (c0 + i >= a.n) cur[i] = 0xff800000u;   // -inf
                    }
                    if (PASS == 1) {
#pragma unroll
                        for (int i = 0; i < 32; i++) m[(lp & 1) * 32 + i] = fmaxf(m[(lp & 1) * 32 + i], __uint_as_float(cur[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const float v = __uint_as_float(cur[i]);
                            if (v >= Ts) {   // rare: ~2 need of n scores
                                if (cnt < (uint32_t)a.cap) mybuf[cnt] = pack_key(v * M0_INV, (uint32_t)(c0 + i));
                                cnt++;
                            }
                        }
                    }
                }
                tmem_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(acc_empty + buf));
            }
            if (PASS == 1 && rowok) {
#pragma unroll
                for (int i = 0; i < 64; i++) seg[g * 64 + i] = ordered_bits(m[i] * M0_INV);
            }
        }
        if (PASS == 2 && rowok) a.cnt[(row * a.nsplit + split) * EW + h] = cnt;
    }
    tmem_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem0, 512);
}

// T of every row: the need-th largest of its S_total segment maxima (32-step bisection on the ordered bit patterns; a row's
// maxima are 2 KB and stay in L1 between the steps).  One warp per row.
__global__ void stage0_thresh_kernel(const uint32_t* __restrict__ segmax, int64_t nq, int S_total, int need, uint32_t* __restrict__ thr) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nq) return;
    const uint32_t* s = segmax + row * (int64_t)S_total;
    uint32_t t = 0u;
    if (S_total <= 1024) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = (lane + 32 * i < S_total) ? s[lane + 32 * i] : 0u;
        for (int bit = 31; bit >= 0; bit--) {
            const uint32_t cand = t | (1u << bit);
            int c = 0;
#pragma unroll
            for (int i = 0; i < 32; i++) c += (v[i] >= cand) ? 1 : 0;
            if (__reduce_add_sync(0xffffffffu, c) >= need) t = cand;
        }
    } else {
        for (int bit = 31; bit >= 0; bit--) {
            const uint32_t cand = t | (1u << bit);
            int c = 0;
            for (int i = lane; i < S_total; i += 32) c += (s[i] >= cand) ? 1 : 0;
            if (__reduce_add_sync(0xffffffffu, c) >= need) t = cand;
        }
    }
    if (lane == 0) thr[row] = t;
}

struct S0FArgs {
    const float* q_centers;
    const int64_t* self_idx;
    const float* centers;
    int64_t q_start, q_stride, nq, n;
    int kp, need, cap, capF, P2, nsplit;
    const uint32_t* cnt;
    const ull* buf;
    uint32_t* flags;
    int32_t* fail_list;
    int32_t* out_idx;
    float* out_score;
};

constexpr int M0_FWARPS = 8;
constexpr int M0_TLD = 68;                    // row stride (floats) of the transposed half-rows: LDS.128 of a quarter-warp hits distinct banks

// One warp per row.  The candidates of all splits are gathered into shared memory; the need-th best approximate score a* is
// found by bisection on the ordered bit patterns (held in registers, VPL per lane); the candidates with approximate score >= a*
// are re-scored with the canonical fp32 chain (acc = fmaf(q[c], g[c], acc), c ascending: the arithmetic of stage0_topk.cu) and
// sorted; the list is accepted when its kp-th fp32 score exceeds a* + eps (every other gallery item has approximate score
// < a*, hence fp32 score < a* + eps: it can neither belong to the list nor tie into it).
template <int VPL>
__global__ void __launch_bounds__(M0_FWARPS * 32) stage0_final_kernel(S0FArgs f) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * M0_FWARPS + warp;
    if (row >= f.nq) return;
    // per warp: [union: all candidates (capF keys) | transposed half-rows of 32 candidates ([32][68] floats)] [b2: P2 keys] [q: 128]
    const size_t ubytes = (size_t)f.capF * 8 > (size_t)32 * M0_TLD * 4 ? (size_t)f.capF * 8 : (size_t)32 * M0_TLD * 4;
    unsigned char* wbase = smem_raw + (size_t)warp * (ubytes + (size_t)f.P2 * 8 + M0_C * 4);
    ull* b = reinterpret_cast<ull*>(wbase);                     // all candidates (unsorted); dead after the compaction
    float* tile = reinterpret_cast<float*>(wbase);              // ... then the transposed candidate rows
    ull* b2 = reinterpret_cast<ull*>(wbase + ubytes);           // the re-scored candidates, finally by fp32 score
    float* qs = reinterpret_cast<float*>(wbase + ubytes + (size_t)f.P2 * 8);
    const float* q = f.q_centers ? f.q_centers + row * M0_C : f.centers + (f.q_start + row * f.q_stride) * M0_C;
    long long self = -1;
    if (f.self_idx) self = f.self_idx[row];
    else if (!f.q_centers) self = f.q_start + row * f.q_stride;
    const float4 q4 = __ldg(reinterpret_cast<const float4*>(q) + lane);
    reinterpret_cast<float4*>(qs)[lane] = q4;
    const float qn2 = warp_sum(q4.x * q4.x + q4.y * q4.y + q4.z * q4.z + q4.w * q4.w);
    bool fail = f.flags[2] != 0u, overflow = false;
    int c = 0;
    {
        // counts of the sub-lists (lane s holds sub-list s; groups of 32), exclusive prefix by shuffles, then the first 64
        // entries of every sub-list are requested before any is stored: one memory latency per row instead of one per sub-list
        const ull* rbuf = f.buf + row * (int64_t)f.nsplit * (int64_t)f.cap;
        for (int s0 = 0; s0 < f.nsplit && !overflow; s0 += 32) {
            const int ns = min(32, f.nsplit - s0);
            const uint32_t mine = lane < ns ? f.cnt[row * f.nsplit + s0 + lane] : 0u;
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            if (__any_sync(0xffffffffu, mine > (uint32_t)f.cap) || c + (int)tot > f.capF) {
                overflow = true;
                break;
            }
            for (int sb = 0; sb < ns; sb += 8) {
                ull x[8][2];
                uint32_t cs[8], off[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int sl = min(sb + u, ns - 1);
                    cs[u] = sb + u < ns ? __shfl_sync(0xffffffffu, mine, sl) : 0u;
                    off[u] = __shfl_sync(0xffffffffu, incl - mine, sl);
                    const ull* src = rbuf + (int64_t)(s0 + sl) * f.cap;
#pragma unroll
                    for (int w = 0; w < 2; w++) x[u][w] = (uint32_t)(lane + 32 * w) < cs[u] ? src[lane + 32 * w] : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
#pragma unroll
                    for (int w = 0; w < 2; w++)
                        if ((uint32_t)(lane + 32 * w) < cs[u]) b[c + off[u] + lane + 32 * w] = x[u][w];
                    if (cs[u] > 64u) {   // long sub-list (a class block): the rest, plainly
                        const ull* src = rbuf + (int64_t)(s0 + sb + u) * f.cap;
                        for (int e = 64 + lane; e < (int)cs[u]; e += 32) b[c + off[u] + e] = src[e];
                    }
                }
            }
            c += (int)tot;
        }
    }
    if (overflow && lane == 0) atomicAdd(f.flags + 3, 1u);
    fail |= overflow | (c < f.need);
    __syncwarp();
    uint32_t t = 0u;
    int m = 0;
    if (!fail) {
        uint32_t v[VPL];
        uint32_t all_and = 0xffffffffu, all_or = 0u;
#pragma unroll
        for (int i = 0; i < VPL; i++) {
            v[i] = 0u;
            if (32 * i < c) {   // (warp-uniform)
                v[i] = (lane + 32 * i < c) ? (uint32_t)(b[lane + 32 * i] >> 32) : 0u;
                if (lane + 32 * i < c) {
                    all_and &= v[i];
                    all_or |= v[i];
                }
            }
        }
        // bits on which every candidate agrees need no bisection step: start below the common prefix
        all_and = __reduce_and_sync(0xffffffffu, all_and);
        all_or = __reduce_or_sync(0xffffffffu, all_or);
        const uint32_t differ = all_and ^ all_or;
        const int top = differ ? 31 - __clz(differ) : -1;
        t = top >= 31 ? 0u : (all_and & ~((2u << top) - 1u));
        for (int bit = top; bit >= 0; bit--) {
            const uint32_t cand = t | (1u << bit);
            int k = 0;
#pragma unroll
            for (int i = 0; i < VPL; i++)
                if (32 * i < c) k += (v[i] >= cand) ? 1 : 0;
            if (__reduce_add_sync(0xffffffffu, k) >= f.need) t = cand;
        }
        // compact the candidates with approximate score >= a* (exact ties at a* come along), then re-score them
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int i = 0; i < VPL; i++) {
            const bool take = v[i] >= t && t != 0u;
            const unsigned bal = __ballot_sync(0xffffffffu, take);
            const int pos = m + __popc(bal & lt);
            if (take && pos < f.P2) b2[pos] = b[lane + 32 * i];
            m += __popc(bal);
        }
        fail = m > f.P2 || m < f.need;   // more exact ties at a* than the list holds
    }
    if (!fail) {
        __syncwarp();
        // 32 candidates at a time, in two half-rows of 64 channels: the 256-byte half-rows are fetched by half-warps (one
        // instruction = two candidates, coalesced) and laid out [candidate][68] in shared memory (over the dead candidate
        // buffer), then lane l continues the chain of candidate l from there (stride 68 words: the LDS.128 of a quarter-warp
        // falls on distinct banks).  One row per lane straight from L2 costs 32 cache lines per load instruction.
        const int hl = lane >> 4, l16 = lane & 15;
        for (int base = 0; base < m; base += 32) {
            const int e = base + lane;
            const uint32_t idx = e < m ? key_index(b2[e]) : 0u;
            float acc = 0.f;
#pragma unroll
            for (int hc = 0; hc < 2; hc++) {
                float4 tv[16];
#pragma unroll
                for (int r = 0; r < 16; r++) {
                    const int er = base + 2 * r + hl;
                    const uint32_t ir = __shfl_sync(0xffffffffu, idx, 2 * r + hl);
                    tv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (er < m) tv[r] = __ldg(reinterpret_cast<const float4*>(f.centers + (int64_t)ir * M0_C + hc * 64) + l16);
                }
                __syncwarp();   // the chains of the previous half are done with the tile
#pragma unroll
                for (int r = 0; r < 16; r++) *reinterpret_cast<float4*>(tile + (2 * r + hl) * M0_TLD + 4 * l16) = tv[r];
                __syncwarp();
                const float4* g = reinterpret_cast<const float4*>(tile + lane * M0_TLD);
#pragma unroll 8
                for (int k4 = 0; k4 < 16; k4++) {
                    const float4 gv = g[k4];
                    const float4 qv = reinterpret_cast<const float4*>(qs)[hc * 16 + k4];
                    acc = fmaf(qv.x, gv.x, acc);
                    acc = fmaf(qv.y, gv.y, acc);
                    acc = fmaf(qv.z, gv.z, acc);
                    acc = fmaf(qv.w, gv.w, acc);
                }
            }
            if (e < m) b2[e] = pack_key((long long)idx == self ? -100.0f : acc, idx);
        }
        __syncwarp();
        for (int e = m + lane; e < f.P2; e += 32) b2[e] = 0ull;
        warp_bitonic_sort_desc(b2, f.P2, lane);
        const float eps = M0_EPS * sqrtf(qn2 * __uint_as_float(f.flags[1])) + 1e-30f;
        fail = !(key_score(b2[f.kp - 1]) > from_ordered_bits(t) + eps);
    }
    if (fail) {
        if (lane == 0) f.fail_list[atomicAdd(f.flags + 0, 1u)] = (int32_t)row;
        return;
    }
    for (int e = lane; e < f.kp; e += 32) {
        const ull key = b2[e];
        f.out_idx[row * f.kp + e] = (int32_t)key_index(key);
        f.out_score[row * f.kp + e] = key_score(key);
    }
}

// Exact redo of the rows on the fail list: persistent CTAs; all 256 threads write the row's canonical scores to the CTA's
// scratch row, then warp 0 selects (the same routine as stage0_rowselect_kernel).
__global__ void __launch_bounds__(256) stage0_fallback_kernel(S0FArgs f, float* __restrict__ scratch, int64_t ld, int P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    ull* b = reinterpret_cast<ull*>(smem_raw);
    float* qs = reinterpret_cast<float*>(b + P);
    const int tid = threadIdx.x;
    const uint32_t nfail = f.flags[0];
    float* srow = scratch + (int64_t)blockIdx.x * ld;
    for (uint32_t i = blockIdx.x; i < nfail; i += gridDim.x) {
        const int64_t row = f.fail_list[i];
        const float* q = f.q_centers ? f.q_centers + row * M0_C : f.centers + (f.q_start + row * f.q_stride) * M0_C;
        long long self = -1;
        if (f.self_idx) self = f.self_idx[row];
        else if (!f.q_centers) self = f.q_start + row * f.q_stride;
        __syncthreads();
        if (tid < M0_C) qs[tid] = q[tid];
        __syncthreads();
        for (int64_t col = tid; col < f.n; col += 256) {
            const float4* g = reinterpret_cast<const float4*>(f.centers + col * M0_C);
            float acc = 0.f;
#pragma unroll 8
            for (int k4 = 0; k4 < M0_C / 4; k4++) {
                const float4 gv = __ldg(g + k4);
                const float4 qv = reinterpret_cast<const float4*>(qs)[k4];
                acc = fmaf(qv.x, gv.x, acc);
                acc = fmaf(qv.y, gv.y, acc);
                acc = fmaf(qv.z, gv.z, acc);
                acc = fmaf(qv.w, gv.w, acc);
            }
            srow[col] = acc;
        }
        __syncthreads();
        if (tid < 32) warp_rowselect_stream(srow, f.n, ld, self, f.kp, P, b, tid, 0ull, f.out_idx + row * f.kp, f.out_score + row * f.kp);
    }
}

// ---- host side ----
struct S0MPlan {
    int qblocks, ntiles, tiles_per, nsplit, ngroups, S_per, S_total, need, cap, capF, P2, Pfb, fb_ctas;
    int64_t n_pad, q_pad, ld;
    size_t off_gpk, off_qpk, off_flags, off_seg, off_thr, off_cnt, off_buf, off_fail, off_scratch, total;
};

static int pow2_at_least(int x, int lo) {
    int p = lo;
    while (p < x) p <<= 1;
    return p;
}

static S0MPlan plan_s0m(int64_t nq, int64_t n, int kp, int sms) {
    S0MPlan p{};
    p.qblocks = (int)((nq + M0_BM - 1) / M0_BM);
    p.ntiles = (int)((n + M0_BN - 1) / M0_BN);
    p.need = kp + M0_SLACK + 1;
    p.capF = 2 * pow2_at_least(3 * p.need, 512);  // candidates the final kernel takes per row (1,024 up to kp = 157, else 2,048)
    p.cap = p.capF / 4;                           // candidates one (split, read-out warp) sub-list keeps per row
    p.P2 = pow2_at_least(p.need + 32, 32);        // re-scored list (room for exact ties at the cut)
    p.Pfb = pow2_at_least(kp + 64, 32);
    // gallery splits: enough CTAs for ~4 waves when the query blocks alone are few, at least 8 tiles per CTA, and among the
    // admissible counts the one that fills its last wave best
    const int hi = std::max(1, p.ntiles / 8);
    const int want = (int)std::min<int64_t>(hi, std::max<int64_t>(1, (4ll * sms + p.qblocks - 1) / p.qblocks));
    int best = want;
    double best_eff = 0.0;
    for (int s = std::max(1, want - 2); s <= std::min(hi, want + 3); s++) {
        const int tp = (p.ntiles + s - 1) / s;
        const int ns = (p.ntiles + tp - 1) / tp;
        const int64_t ctas = (int64_t)ns * p.qblocks;
        const double eff = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms);
        if (eff > best_eff + 0.02) {
            best_eff = eff;
            best = s;
        }
    }
    p.tiles_per = (p.ntiles + best - 1) / best;
    p.nsplit = (p.ntiles + p.tiles_per - 1) / p.tiles_per;
    // segments per row: about 3.5 x need over all splits, 64 per tile group
    p.ngroups = std::max(1, std::min(p.tiles_per, (int)((3.5 * p.need) / (64.0 * p.nsplit) + 0.5)));
    p.S_per = p.ngroups * 64;
    p.S_total = p.S_per * p.nsplit;
    p.n_pad = (int64_t)p.ntiles * M0_BN;
    p.q_pad = (int64_t)p.qblocks * M0_BM;
    p.ld = (n + 3) & ~(int64_t)3;
    p.fb_ctas = 2 * sms;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += align_up(bytes, 1024);
        return at;
    };
    p.off_flags = take(64);
    p.off_gpk = take((size_t)p.ntiles * M0_TILE_BYTES);
    p.off_qpk = take((size_t)p.qblocks * M0_ABLK_BYTES);
    p.off_seg = take((size_t)nq * p.S_total * 4);
    p.off_thr = take((size_t)nq * 4);
    p.off_cnt = take((size_t)nq * p.nsplit * M0_EW2 * 4);
    p.off_buf = take((size_t)nq * p.nsplit * M0_EW2 * p.cap * 8);
    p.off_fail = take((size_t)nq * 4);
    p.off_scratch = take((size_t)p.fb_ctas * p.ld * 4);
    p.total = o + 1024;
    return p;
}

bool stage0_mma_supported(int64_t nq, int64_t n, int c, int kp) {
    const char* e = getenv("VR_STAGE0");   // VR_STAGE0=sgemm forces the fp32 path (A/B tests); read per call
    if (e && e[0] == 's') return false;
    // enough queries to fill 128-row blocks, enough gallery for >= 2 x need segments of >= 8 scores, shortlists that keep the
    // candidate buffers small; everything else takes the fp32 path
    return c == M0_C && nq >= 256 && kp >= 1 && kp <= 256 && n >= 8ll * (kp + M0_SLACK + 1) && n >= 2048 && n < 0x7fffffffll;
}

size_t stage0_mma_workspace_bytes(int64_t nq, int64_t n, int kp, int sms) { return plan_s0m(nq, n, kp, sms).total; }

int stage0_mma_topk(const float* q_centers, const int64_t* self_idx, const float* centers, int64_t q_start, int64_t q_stride,
                    int64_t nq, int64_t n, int kp, int32_t* out_idx, float* out_score, void* ws, size_t ws_bytes, int sms,
                    uint32_t* stats_dev, cudaStream_t st) {
    const S0MPlan p = plan_s0m(nq, n, kp, sms);
    if (p.total > ws_bytes) {
        set_error("stage0 (tensor-core path): workspace %zu < %zu", ws_bytes, p.total);
        return VR_E_WORKSPACE;
    }
    unsigned char* w = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
    uint32_t* flags = reinterpret_cast<uint32_t*>(w + p.off_flags);
    unsigned char* gpk = w + p.off_gpk;
    unsigned char* qpk = w + p.off_qpk;
    uint32_t* segmax = reinterpret_cast<uint32_t*>(w + p.off_seg);
    uint32_t* thr = reinterpret_cast<uint32_t*>(w + p.off_thr);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(w + p.off_cnt);
    ull* buf = reinterpret_cast<ull*>(w + p.off_buf);
    int32_t* fail_list = reinterpret_cast<int32_t*>(w + p.off_fail);
    float* scratch = reinterpret_cast<float*>(w + p.off_scratch);

    VR_CHECK_CUDA(cudaMemsetAsync(flags, 0, 64, st));
    // the last split may own fewer tiles: its unused slots must read 0 (the lowest key)
    VR_CHECK_CUDA(cudaMemsetAsync(segmax, 0, (size_t)nq * p.S_total * 4, st));
    stage0_pack_kernel<32><<<(unsigned)((p.n_pad + 255) / 256), 256, 0, st>>>(centers, n, p.n_pad, 0, 1, gpk, flags, 1);
    VR_LAUNCH_CHECK();
    stage0_pack_kernel<16><<<(unsigned)((p.q_pad + 255) / 256), 256, 0, st>>>(q_centers ? q_centers : centers, nq, p.q_pad,
                                                                               q_centers ? 0 : q_start, q_centers ? 1 : q_stride, qpk,
                                                                               flags, 0);
    VR_LAUNCH_CHECK();
    S0MArgs a{};
    a.gpk = gpk;
    a.qpk = qpk;
    a.nq = nq;
    a.n = n;
    a.ntiles = p.ntiles;
    a.tiles_per = p.tiles_per;
    a.nsplit = p.nsplit;
    a.ngroups = p.ngroups;
    a.S_per = p.S_per;
    a.S_total = p.S_total;
    a.cap = p.cap;
    a.segmax = segmax;
    a.thr = thr;
    a.cnt = cnt;
    a.buf = buf;
    VR_CHECK_CUDA(cudaFuncSetAttribute(stage0_mma_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M0_SMEM));
    VR_CHECK_CUDA(cudaFuncSetAttribute(stage0_mma_kernel<2, M0_EW2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M0_SMEM));
    const dim3 grid((unsigned)p.nsplit, (unsigned)p.qblocks);
    stage0_mma_kernel<1, 1><<<grid, 64 + 128, M0_SMEM, st>>>(a);
    VR_LAUNCH_CHECK();
    stage0_thresh_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(segmax, nq, p.S_total, p.need, thr);
    VR_LAUNCH_CHECK();
    stage0_mma_kernel<2, M0_EW2><<<grid, 64 + 128 * M0_EW2, M0_SMEM, st>>>(a);
    VR_LAUNCH_CHECK();
    S0FArgs f{};
    f.q_centers = q_centers;
    f.self_idx = self_idx;
    f.centers = centers;
    f.q_start = q_start;
    f.q_stride = q_stride;
    f.nq = nq;
    f.n = n;
    f.kp = kp;
    f.need = p.need;
    f.cap = p.cap;
    f.capF = p.capF;
    f.P2 = p.P2;
    f.nsplit = p.nsplit * M0_EW2;   // sub-lists per row
    f.cnt = cnt;
    f.buf = buf;
    f.flags = flags;
    f.fail_list = fail_list;
    f.out_idx = out_idx;
    f.out_score = out_score;
    const size_t smem_f = (size_t)M0_FWARPS * (std::max((size_t)p.capF * 8, (size_t)32 * M0_TLD * 4) + (size_t)p.P2 * 8 + M0_C * 4);
    auto* fk = p.capF <= 1024 ? stage0_final_kernel<32> : stage0_final_kernel<64>;   // capF = 1,024 (kp <= 157) or 2,048
    VR_CHECK_CUDA(cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
    fk<<<(unsigned)((nq + M0_FWARPS - 1) / M0_FWARPS), M0_FWARPS * 32, smem_f, st>>>(f);
    VR_LAUNCH_CHECK();
    const size_t smem_fb = (size_t)p.Pfb * 8 + M0_C * 4;
    stage0_fallback_kernel<<<(unsigned)p.fb_ctas, 256, smem_fb, st>>>(f, scratch, p.ld, p.Pfb);
    VR_LAUNCH_CHECK();
    if (stats_dev) VR_CHECK_CUDA(cudaMemcpyAsync(stats_dev, flags, 16, cudaMemcpyDeviceToDevice, st));
    return VR_OK;
}

}  // namespace vr
