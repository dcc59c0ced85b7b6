// S1, fp32 paths: first-stage retrieval = fp32 cosine GEMM over the global embeddings with a
// streaming per-query top-K' select.  Replaces calc_similarity(stage=0)
// (utilities/diml.py:83-85), the self mask (evaluation/eval_cvt_diml.py:327) and the head
// of the full argsort (:329-332).  Batches over 128-d embeddings take the tensor-core path of
// stage0_mma.cu (score matrix in tensor memory only); this file is the canonical fp32 arithmetic
// that path reproduces, and serves everything it does not cover:
//   - a few queries (< 256): stage0_select_kernel, the select fused into the score tile (every BM x 64 tile
//     is filtered in registers against the per-row running threshold; survivors go to a per-row buffer in
//     shared memory that one warp re-sorts before it overflows); no score matrix in memory;
//   - batches with other widths or shortlists beyond 256: stage0_scores_kernel writes a [rows, N] score
//     chunk to the workspace (HBM) and stage0_rowselect_kernel streams each row once.
//
// Arithmetic: plain fp32 FMA, sequential over the C channels (no TF32: the top-K sets
// must match the reference's fp32 path, SURVEY.md section 7 hard part 3).
#include <algorithm>
#include "common.cuh"
#include "kernels.h"
#include "stage0_select.cuh"

namespace vr {

constexpr int S0_BN = 64;      // gallery columns per tile
constexpr int S0_BK = 32;      // channels per pipeline stage
constexpr int S0_LD = 36;      // padded smem row stride (floats): conflict-free LDS.128
constexpr int S0_THREADS = 256;
constexpr int S0_STAGES = 3;   // cp.async ring: one CTA barrier per chunk instead of two

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct Stage0Args {
    const float* q_centers;   // [nq, C] or nullptr (then bank centres at q_start + i*q_stride)
    const int64_t* self_idx;  // [nq] or nullptr
    const float* centers;     // [N, C]
    int64_t q_start, q_stride, nq, n;
    int c, kp, P, nsplit;
    unsigned long long* partial;  // [nq, nsplit, kp] keys when nsplit > 1
    int32_t* out_idx;             // [nq, kp]
    float* out_score;             // [nq, kp]
};

template <int BM, int TY, int TX>
__global__ void __launch_bounds__(S0_THREADS, 2) stage0_select_kernel(Stage0Args a) {
    constexpr int RM = BM / TY;
    constexpr int CN = S0_BN / TX;
    static_assert(TY * TX == S0_THREADS, "thread grid");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);                 // [3][BM][LD]
    float* Bs = As + S0_STAGES * BM * S0_LD;                        // [3][BN][LD]
    unsigned long long* buf = reinterpret_cast<unsigned long long*>(Bs + S0_STAGES * S0_BN * S0_LD);  // [BM][P]
    unsigned long long* thr = buf + (size_t)BM * a.P;               // [BM]
    int* cnt = reinterpret_cast<int*>(thr + BM);                    // [BM]
    long long* selfs = reinterpret_cast<long long*>(cnt + BM + (BM & 1));  // [BM]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid % TX, ty = tid / TX;
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int P = a.P, kp = a.kp, C = a.c;
    // gallery range of this split, in whole tiles
    const int64_t tiles_total = (a.n + S0_BN - 1) / S0_BN;
    const int64_t tiles_per = (tiles_total + a.nsplit - 1) / a.nsplit;
    const int64_t tile_lo = (int64_t)blockIdx.y * tiles_per;
    const int64_t tile_hi = min(tiles_total, tile_lo + tiles_per);
    const int ntiles = (int)max((int64_t)0, tile_hi - tile_lo);
    const int KC = (C + S0_BK - 1) / S0_BK;

    for (int r = tid; r < BM; r += S0_THREADS) {
        cnt[r] = 0;
        thr[r] = 0ull;
        int64_t g = row0 + r;
        long long s = -1;
        if (g < a.nq) {
            if (a.self_idx) s = a.self_idx[g];
            else if (!a.q_centers) s = a.q_start + g * a.q_stride;
        }
        selfs[r] = s;
    }

    auto issue = [&](int it) {
        const int tile = it / KC, kc = it % KC, st = it % S0_STAGES;
        const int64_t n0 = (tile_lo + tile) * S0_BN;
        const int k0 = kc * S0_BK;
        float* as = As + st * BM * S0_LD;
        float* bs = Bs + st * S0_BN * S0_LD;
        for (int e = tid; e < BM * (S0_BK / 4); e += S0_THREADS) {
            int r = e / (S0_BK / 4), q4 = e % (S0_BK / 4);
            int64_t g = row0 + r;
            int k = k0 + q4 * 4;
            const float* src = a.centers;
            int bytes = 0;
            if (g < a.nq && k < C) {
                src = a.q_centers ? a.q_centers + g * C + k : a.centers + (a.q_start + g * a.q_stride) * C + k;
                bytes = 16;
            }
            cp_async16(as + r * S0_LD + q4 * 4, src, bytes);
        }
        for (int e = tid; e < S0_BN * (S0_BK / 4); e += S0_THREADS) {
            int r = e / (S0_BK / 4), q4 = e % (S0_BK / 4);
            int64_t g = n0 + r;
            int k = k0 + q4 * 4;
            const float* src = a.centers;
            int bytes = 0;
            if (g < a.n && k < C) {
                src = a.centers + g * C + k;
                bytes = 16;
            }
            cp_async16(bs + r * S0_LD + q4 * 4, src, bytes);
        }
        cp_async_commit();
    };

    float acc[RM][CN];
#pragma unroll
    for (int i = 0; i < RM; i++)
#pragma unroll
        for (int j = 0; j < CN; j++) acc[i][j] = 0.f;

    const int total = ntiles * KC;
    if (total > 0) issue(0);
    if (total > 1) issue(1);
    for (int it = 0; it < total; it++) {
        if (it + 1 < total) cp_async_wait<1>();
        else cp_async_wait<0>();
        // one barrier per chunk: chunk `it` has landed for everybody, and everybody is done with chunk it-1, whose
        // stage is the one chunk it+2 goes to (it also publishes cnt / thr / selfs and the sorts of the last tile)
        __syncthreads();
        if (it + 2 < total) issue(it + 2);
        const float* as = As + (it % S0_STAGES) * BM * S0_LD;
        const float* bs = Bs + (it % S0_STAGES) * S0_BN * S0_LD;
#pragma unroll
        for (int k4 = 0; k4 < S0_BK / 4; k4++) {
            float4 av[RM], bv[CN];
#pragma unroll
            for (int i = 0; i < RM; i++) av[i] = *reinterpret_cast<const float4*>(as + (ty + TY * i) * S0_LD + k4 * 4);
#pragma unroll
            for (int j = 0; j < CN; j++) bv[j] = *reinterpret_cast<const float4*>(bs + (tx + TX * j) * S0_LD + k4 * 4);
#pragma unroll
            for (int i = 0; i < RM; i++)
#pragma unroll
                for (int j = 0; j < CN; j++) {
                    acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]);
                    acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
                    acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]);
                    acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
                }
        }
        if (it % KC == KC - 1) {
            // ---- fused select on the finished BM x 64 score tile ----
            const int64_t n0 = (tile_lo + it / KC) * S0_BN;
#pragma unroll
            for (int i = 0; i < RM; i++) {
                const int r = ty + TY * i;
                const bool rowok = (row0 + r) < a.nq;
                const unsigned long long t = thr[r];
                const long long self = selfs[r];
#pragma unroll
                for (int j = 0; j < CN; j++) {
                    const int64_t col = n0 + tx + TX * j;
                    float s = acc[i][j];
                    acc[i][j] = 0.f;
                    if (rowok && col < a.n) {
                        if (col == self) s = -100.0f;  // eval_cvt_diml.py:327
                        unsigned long long key = pack_key(s, (uint32_t)col);
                        if (key > t) {
                            int pos = atomicAdd(&cnt[r], 1);
                            buf[(size_t)r * P + pos] = key;
                        }
                    }
                }
            }
            __syncthreads();
            for (int r = warp; r < BM; r += S0_THREADS / 32) {
                int n = cnt[r];
                if (n > P - S0_BN) {
                    unsigned long long* b = buf + (size_t)r * P;
                    for (int e = n + lane; e < P; e += 32) b[e] = 0ull;
                    warp_bitonic_sort_desc(b, P, lane);
                    if (lane == 0) {
                        cnt[r] = min(n, kp);
                        if (n >= kp) thr[r] = b[kp - 1];
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- final per-row sort and write-out ----
    for (int r = warp; r < BM; r += S0_THREADS / 32) {
        const int64_t g = row0 + r;
        if (g >= a.nq) continue;
        int n = cnt[r];
        unsigned long long* b = buf + (size_t)r * P;
        for (int e = n + lane; e < P; e += 32) b[e] = 0ull;
        warp_bitonic_sort_desc(b, P, lane);
        if (a.nsplit > 1) {
            unsigned long long* dst = a.partial + ((size_t)g * a.nsplit + blockIdx.y) * kp;
            for (int e = lane; e < kp; e += 32) dst[e] = b[e];
        } else {
            for (int e = lane; e < kp; e += 32) {
                unsigned long long key = b[e];
                a.out_idx[g * kp + e] = key ? (int32_t)key_index(key) : -1;
                a.out_score[g * kp + e] = key ? key_score(key) : 0.f;
            }
        }
    }
}

// ---- large query batches: score chunk by a register-tiled SGEMM, then one warp per row selects ----
// The fused kernel above keeps BM x P keys of candidate buffers in shared memory, which caps its row block at 32 and its
// register tile at 2 x 4: 6 LDS.128 per 32 FMAs, shared-memory bound, and the buffer sorts of single warps stall whole
// CTAs at the chunk barrier.  For batches that fill the GPU the work is split instead: this kernel is a plain 128 x 64
// SGEMM tile (8 x 4 per thread: 12 LDS.128 per 128 FMAs) that writes a [rows, ld] score chunk to the workspace, and
// stage0_rowselect_kernel streams each row once.  Every score is the same sequential fp32 FMA chain over the channels, so
// both paths return identical shortlists.
constexpr int S0G_BM = 128;
template <int TY, int TX>
__global__ void __launch_bounds__(S0_THREADS, 2) stage0_scores_kernel(Stage0Args a, int64_t r_begin, int64_t rows, int tiles_per,
                                                                      float* __restrict__ scores, int64_t ld) {
    constexpr int BM = S0G_BM;
    constexpr int RM = BM / TY;
    constexpr int CN = S0_BN / TX;
    static_assert(TY * TX == S0_THREADS, "thread grid");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);                 // [3][BM][LD]
    float* Bs = As + S0_STAGES * BM * S0_LD;                        // [3][BN][LD]
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int64_t row0 = (int64_t)blockIdx.y * BM;                  // row of the chunk
    const int C = a.c;
    const int64_t tiles_total = (a.n + S0_BN - 1) / S0_BN;
    const int64_t tile_lo = (int64_t)blockIdx.x * tiles_per;
    const int ntiles = (int)max((int64_t)0, min(tiles_total, tile_lo + tiles_per) - tile_lo);
    const int KC = (C + S0_BK - 1) / S0_BK;

    auto issue = [&](int it) {
        const int tile = it / KC, kc = it % KC, st = it % S0_STAGES;
        const int64_t n0 = (tile_lo + tile) * S0_BN;
        const int k0 = kc * S0_BK;
        float* as = As + st * BM * S0_LD;
        float* bs = Bs + st * S0_BN * S0_LD;
        for (int e = tid; e < BM * (S0_BK / 4); e += S0_THREADS) {
            int r = e / (S0_BK / 4), q4 = e % (S0_BK / 4);
            int64_t g = r_begin + row0 + r;
            int k = k0 + q4 * 4;
            const float* src = a.centers;
            int bytes = 0;
            if (row0 + r < rows && k < C) {
                src = a.q_centers ? a.q_centers + g * C + k : a.centers + (a.q_start + g * a.q_stride) * C + k;
                bytes = 16;
            }
            cp_async16(as + r * S0_LD + q4 * 4, src, bytes);
        }
        for (int e = tid; e < S0_BN * (S0_BK / 4); e += S0_THREADS) {
            int r = e / (S0_BK / 4), q4 = e % (S0_BK / 4);
            int64_t g = n0 + r;
            int k = k0 + q4 * 4;
            const float* src = a.centers;
            int bytes = 0;
            if (g < a.n && k < C) {
                src = a.centers + g * C + k;
                bytes = 16;
            }
            cp_async16(bs + r * S0_LD + q4 * 4, src, bytes);
        }
        cp_async_commit();
    };

    float acc[RM][CN];
#pragma unroll
    for (int i = 0; i < RM; i++)
#pragma unroll
        for (int j = 0; j < CN; j++) acc[i][j] = 0.f;

    const int total = ntiles * KC;
    if (total > 0) issue(0);
    if (total > 1) issue(1);
    for (int it = 0; it < total; it++) {
        if (it + 1 < total) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncthreads();   // chunk `it` has landed for everybody, and everybody is done with the stage chunk it+2 goes to
        if (it + 2 < total) issue(it + 2);
        const float* as = As + (it % S0_STAGES) * BM * S0_LD;
        const float* bs = Bs + (it % S0_STAGES) * S0_BN * S0_LD;
#pragma unroll
        for (int k4 = 0; k4 < S0_BK / 4; k4++) {
            float4 av[RM], bv[CN];
#pragma unroll
            for (int i = 0; i < RM; i++) av[i] = *reinterpret_cast<const float4*>(as + (ty + TY * i) * S0_LD + k4 * 4);
#pragma unroll
            for (int j = 0; j < CN; j++) bv[j] = *reinterpret_cast<const float4*>(bs + (tx + TX * j) * S0_LD + k4 * 4);
#pragma unroll
            for (int i = 0; i < RM; i++)
#pragma unroll
                for (int j = 0; j < CN; j++) {
                    acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]);
                    acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
                    acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]);
                    acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
                }
        }
        if (it % KC == KC - 1) {
            const int64_t n0 = (tile_lo + it / KC) * S0_BN;
#pragma unroll
            for (int i = 0; i < RM; i++) {
                const int64_t r = row0 + ty + TY * i;
#pragma unroll
                for (int j = 0; j < CN; j++) {
                    const int64_t col = n0 + tx + TX * j;
                    if (r < rows && col < a.n) scores[r * ld + col] = acc[i][j];
                    acc[i][j] = 0.f;
                }
            }
        }
    }
}

// One warp per row of the score chunk: self mask (eval_cvt_diml.py:327), streaming top-kp select with a running
// threshold (survivors go to a P-key buffer that is re-sorted when 64 more might not fit), final sort, write-out.
// M > 0 (kp <= 32 M): a first pass over the row keeps the M best scores of every lane in registers; the smallest of those
// 32 M values is a lower bound of the kp-th best score, so the second pass starts with that threshold and the buffer
// normally fills once (~1.5 kp survivors) instead of being re-sorted ~6 times while the threshold climbs from zero.
template <int M>
__global__ void stage0_rowselect_kernel(Stage0Args a, int64_t r_begin, int64_t rows, const float* __restrict__ scores, int64_t ld) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int P = a.P, kp = a.kp;
    unsigned long long* b = reinterpret_cast<unsigned long long*>(smem_raw) + (size_t)warp * P;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= rows) return;
    const int64_t g = r_begin + r;
    long long self = -1;
    if (a.self_idx) self = a.self_idx[g];
    else if (!a.q_centers) self = a.q_start + g * a.q_stride;
    const float* srow = scores + r * ld;
    unsigned long long thr = 0ull;
    if (M > 0) {
        uint32_t top[M > 0 ? M : 1];   // this lane's best ordered scores, descending
#pragma unroll
        for (int i = 0; i < M; i++) top[i] = 0u;
        auto load4p = [&](float2 (&d)[4], int64_t base) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t col0 = base + 64 * u + 2 * lane;
                d[u] = make_float2(0.f, 0.f);
                if (col0 < ld) d[u] = *reinterpret_cast<const float2*>(srow + col0);   // (kept in L2 for the second pass)
            }
        };
        float2 cur[4], nxt[4];
        load4p(cur, 0);
        for (int64_t base = 0; base < a.n; base += 256) {
            if (base + 256 < a.n) load4p(nxt, base + 256);
#pragma unroll
            for (int u = 0; u < 4; u++) {
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const int64_t col = base + 64 * u + 2 * lane + t;
                    float s = t ? cur[u].y : cur[u].x;
                    if (col == self) s = -100.0f;
                    uint32_t o = col < a.n ? ordered_bits(s) : 0u;
                    if (o > top[M - 1]) {
#pragma unroll
                        for (int i = 0; i < M; i++)
                            if (o > top[i]) { const uint32_t x = top[i]; top[i] = o; o = x; }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) cur[u] = nxt[u];
        }
        const uint32_t t0 = __reduce_min_sync(0xffffffffu, top[M - 1]);   // at least 32 M >= kp scores are >= t0
        const unsigned long long t0key = (unsigned long long)t0 << 32;     // the smallest key with that score
        thr = t0key ? t0key - 1ull : 0ull;                                // pass 2 keeps key > thr, i.e. key >= t0key
    }
    warp_rowselect_stream(srow, a.n, ld, self, kp, P, b, lane, thr, a.out_idx + g * kp, a.out_score + g * kp);
}

// Merge the nsplit partial shortlists of every query (one warp per query).
__global__ void stage0_merge_kernel(const unsigned long long* partial, int64_t nq, int nsplit, int kp, int P2,
                                    int32_t* out_idx, float* out_score) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long* b = reinterpret_cast<unsigned long long*>(smem_raw) + (size_t)warp * P2;
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (g >= nq) return;
    const int m = nsplit * kp;
    const unsigned long long* src = partial + (size_t)g * m;
    for (int e = lane; e < P2; e += 32) b[e] = (e < m) ? src[e] : 0ull;
    warp_bitonic_sort_desc(b, P2, lane);
    for (int e = lane; e < kp; e += 32) {
        unsigned long long key = b[e];
        out_idx[g * kp + e] = key ? (int32_t)key_index(key) : -1;
        out_score[g * kp + e] = key ? key_score(key) : 0.f;
    }
}

// One query against the whole gallery: calc_similarity(stage=0), utilities/diml.py:83-85.
__global__ void global_similarity_kernel(const float* __restrict__ q, const float* __restrict__ centers, int64_t n,
                                         int c, float* __restrict__ sim) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* g = centers + row * c;
    float acc = 0.f;
    for (int k = lane; k < c; k += 32) acc = fmaf(q[k], g[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) sim[row] = acc;
}

static int next_pow2(int x) {
    int p = 32;
    while (p < x) p <<= 1;
    return p;
}

struct Stage0Plan {
    int bm, P, nsplit;
    size_t smem, ws_bytes;
};

static Stage0Plan plan_stage0(int64_t nq, int64_t n, int kp, int sms) {
    Stage0Plan pl{};
    pl.P = next_pow2(kp + S0_BN);
    // largest row block whose candidate buffers fit beside the operand tiles with two CTAs per SM.  The gallery is
    // split (below) only when the row blocks alone cannot fill the SMs: every split starts its threshold from zero and
    // pays the full series of buffer sorts again, and the sorts, not the FMAs, dominate this kernel.
    const size_t budget = 110 * 1024;
    pl.bm = 64;
    for (;;) {
        size_t ops = (size_t)S0_STAGES * (pl.bm + S0_BN) * S0_LD * 4;
        size_t keys = (size_t)pl.bm * pl.P * 8 + (size_t)pl.bm * (8 + 4 + 8) + 16;
        pl.smem = ops + keys;
        if (pl.smem <= budget || pl.bm == 8) break;
        pl.bm >>= 1;
    }
    int64_t row_blocks = (nq + pl.bm - 1) / pl.bm;
    int64_t tiles = (n + S0_BN - 1) / S0_BN;
    int ns = 1;
    if (row_blocks < (int64_t)sms) ns = (int)(((int64_t)sms + row_blocks - 1) / row_blocks);
    if (ns > 8) ns = 8;
    if (ns > tiles / 4) ns = (int)(tiles / 4);
    if (ns < 1) ns = 1;
    pl.nsplit = ns;
    pl.ws_bytes = ns > 1 ? (size_t)nq * ns * kp * 8 : 0;
    return pl;
}

// Two-kernel path: used for batches of at least 256 queries.
struct Stage0GemmPlan {
    bool use;
    int64_t rows, ld;   // rows per chunk (multiple of 128), row stride of the score chunk (floats, multiple of 4)
    int P, warps;
    size_t smem_sel, ws_bytes;
};
static Stage0GemmPlan plan_stage0_gemm(int64_t nq, int64_t n, int kp) {
    Stage0GemmPlan g{};
    g.use = nq >= 256;
    if (!g.use) return g;
    g.ld = (n + 3) & ~(int64_t)3;
    // bytes of scores per chunk: large enough that the row select has a warp for every scheduler slot of the GPU
    // (8,192 rows); the HBM round trip of the chunk costs ~3 % of the SGEMM's time
    const int64_t target = 1ll << 30;
    int64_t rows = target / (g.ld * 4) / S0G_BM * S0G_BM;
    rows = std::max<int64_t>(rows, S0G_BM);
    rows = std::min<int64_t>(rows, 8192);
    rows = std::min<int64_t>(rows, (nq + S0G_BM - 1) / S0G_BM * S0G_BM);
    g.rows = rows;
    g.P = next_pow2(kp + 64);
    g.warps = 8;
    while (g.warps > 1 && (size_t)g.warps * g.P * 8 > 96 * 1024) g.warps >>= 1;
    g.smem_sel = (size_t)g.warps * g.P * 8;
    g.ws_bytes = (size_t)rows * g.ld * 4;
    return g;
}

size_t stage0_workspace_bytes(int64_t nq, int64_t n, int c, int kp, int sms) {
    const size_t fused = plan_stage0(nq, n, kp, sms).ws_bytes, gemm = plan_stage0_gemm(nq, n, kp).ws_bytes;
    const size_t mma = stage0_mma_supported(nq, n, c, kp) ? stage0_mma_workspace_bytes(nq, n, kp, sms) : 0;
    // the tensor-core path, when it applies, is the only one that runs: its workspace alone decides
    return align_up(mma ? mma : std::max(fused, gemm), 256) + 256;
}

template <int BM, int TY, int TX>
static int launch_select(const Stage0Args& a, const Stage0Plan& pl, cudaStream_t st) {
    VR_CHECK_CUDA(cudaFuncSetAttribute(stage0_select_kernel<BM, TY, TX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)pl.smem));
    dim3 grid((unsigned)((a.nq + BM - 1) / BM), (unsigned)pl.nsplit);
    stage0_select_kernel<BM, TY, TX><<<grid, S0_THREADS, pl.smem, st>>>(a);
    VR_LAUNCH_CHECK();
    return VR_OK;
}

int stage0_topk(const float* q_centers, const int64_t* self_idx, const float* centers, int64_t q_start,
                int64_t q_stride, int64_t nq, int64_t n, int c, int kp, int32_t* out_idx, float* out_score,
                void* ws, size_t ws_bytes, int sms, uint32_t* stats_dev, cudaStream_t st) {
    VR_REQUIRE(nq > 0 && n > 0 && kp > 0, "stage0: empty problem (nq=%lld n=%lld kp=%d)", (long long)nq,
               (long long)n, kp);
    VR_REQUIRE(c % 4 == 0, "stage0: embed dim %d must be a multiple of 4", c);
    VR_REQUIRE(kp <= 1984, "stage0: shortlist length %d exceeds 1984", kp);
    VR_REQUIRE(n < 0xffffffffll, "stage0: gallery too large");
    // batches over 128-d embeddings: tcgen05 GEMM with the select fused into the accumulator read-out (stage0_mma.cu),
    // bit-identical lists; the score matrix is never written to memory
    if (stage0_mma_supported(nq, n, c, kp))
        return stage0_mma_topk(q_centers, self_idx, centers, q_start, q_stride, nq, n, kp, out_idx, out_score, ws, ws_bytes, sms,
                               stats_dev, st);
    if (stats_dev) VR_CHECK_CUDA(cudaMemsetAsync(stats_dev, 0, 16, st));
    Stage0Plan pl = plan_stage0(nq, n, kp, sms);
    if (pl.smem > 227 * 1024) {
        set_error("stage0: shortlist %d needs %zu B of shared memory", kp, pl.smem);
        return VR_E_INVALID;
    }
    if (pl.ws_bytes > ws_bytes) {
        set_error("stage0: workspace %zu < %zu", ws_bytes, pl.ws_bytes);
        return VR_E_WORKSPACE;
    }
    Stage0Args a{};
    a.q_centers = q_centers;
    a.self_idx = self_idx;
    a.centers = centers;
    a.q_start = q_start;
    a.q_stride = q_stride;
    a.nq = nq;
    a.n = n;
    a.c = c;
    a.kp = kp;
    const Stage0GemmPlan gp = plan_stage0_gemm(nq, n, kp);
    if (gp.use && gp.ws_bytes <= ws_bytes) {
        a.P = gp.P;
        a.nsplit = 1;
        a.partial = nullptr;
        a.out_idx = out_idx;
        a.out_score = out_score;
        float* scores = reinterpret_cast<float*>(ws);
        const size_t smem_g = (size_t)S0_STAGES * (S0G_BM + S0_BN) * S0_LD * 4;
        VR_CHECK_CUDA(cudaFuncSetAttribute(stage0_scores_kernel<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
        // the pre-pass reads every row twice: worth it while the rows in flight (one per resident warp) come back from L2,
        // not for SOP-sized rows of 242 KB (measured: Cars196 1.06 -> 0.80 ms, SOP 34.8 -> 37.0 ms)
        const int selM = n > 16384 ? 0 : kp <= 128 ? 4 : kp <= 256 ? 8 : 0;
        auto* selk = selM == 4 ? stage0_rowselect_kernel<4> : selM == 8 ? stage0_rowselect_kernel<8> : stage0_rowselect_kernel<0>;
        VR_CHECK_CUDA(cudaFuncSetAttribute(selk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gp.smem_sel));
        const int64_t tiles_total = (n + S0_BN - 1) / S0_BN;
        for (int64_t r0 = 0; r0 < nq; r0 += gp.rows) {
            const int64_t rows = std::min(gp.rows, nq - r0);
            const int64_t row_blocks = (rows + S0G_BM - 1) / S0G_BM;
            // column groups: about sixteen CTAs per SM in the grid (a short tail), at least two column tiles per CTA
            int64_t groups = std::min<int64_t>(tiles_total, std::max<int64_t>(1, (16ll * sms + row_blocks - 1) / row_blocks));
            int tiles_per = (int)std::max<int64_t>(2, (tiles_total + groups - 1) / groups);
            groups = (tiles_total + tiles_per - 1) / tiles_per;
            dim3 grid((unsigned)groups, (unsigned)row_blocks);
            stage0_scores_kernel<16, 16><<<grid, S0_THREADS, smem_g, st>>>(a, r0, rows, tiles_per, scores, gp.ld);
            VR_LAUNCH_CHECK();
            selk<<<(unsigned)((rows + gp.warps - 1) / gp.warps), gp.warps * 32, gp.smem_sel, st>>>(a, r0, rows, scores, gp.ld);
            VR_LAUNCH_CHECK();
        }
        return VR_OK;
    }
    a.P = pl.P;
    a.nsplit = pl.nsplit;
    a.partial = reinterpret_cast<unsigned long long*>(ws);
    a.out_idx = out_idx;
    a.out_score = out_score;
    int rc;
    switch (pl.bm) {
        case 64: rc = launch_select<64, 16, 16>(a, pl, st); break;
        case 32: rc = launch_select<32, 16, 16>(a, pl, st); break;
        case 16: rc = launch_select<16, 16, 16>(a, pl, st); break;
        default: rc = launch_select<8, 8, 32>(a, pl, st); break;
    }
    if (rc) return rc;
    if (pl.nsplit > 1) {
        int P2 = next_pow2(pl.nsplit * kp);
        int warps = 4;
        while (warps > 1 && (size_t)warps * P2 * 8 > 96 * 1024) warps >>= 1;
        size_t smem = (size_t)warps * P2 * 8;
        VR_REQUIRE(smem <= 200 * 1024, "stage0 merge: shortlist too long");
        VR_CHECK_CUDA(cudaFuncSetAttribute(stage0_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        unsigned blocks = (unsigned)((nq + warps - 1) / warps);
        stage0_merge_kernel<<<blocks, warps * 32, smem, st>>>(a.partial, nq, pl.nsplit, kp, P2, out_idx, out_score);
        VR_LAUNCH_CHECK();
    }
    return VR_OK;
}

int global_similarity(const float* q, const float* centers, int64_t n, int c, float* sim, cudaStream_t st) {
    VR_REQUIRE(n > 0 && c > 0, "global_similarity: empty problem");
    unsigned blocks = (unsigned)((n + 7) / 8);
    global_similarity_kernel<<<blocks, 256, 0, st>>>(q, centers, n, c, sim);
    VR_LAUNCH_CHECK();
    return VR_OK;
}

}  // namespace vr
