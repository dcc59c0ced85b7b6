// Attention-rollout producer on the device (evaluation/eval_cvt_diml.py:54-146): the marginals of
// --use_rollout are joint[-1].mean(1) of the chained, filtered, pooled attention maps of every transformer block.
//
//   filter_attention_map [:74-108]   heads fused (min / max), the discard_ratio smallest entries of every image (cls row and
//                                    column included) found
//                                    (torch.topk(largest=False)) and -- as in the reference's fancy-index assignment --
//                                    the UNION of their coordinates over the batch zeroed in every image
//   resize_attn_map [:54-71]         cls row / column dropped (stage 2), both token axes pooled to grid x grid
//                                    (AdaptiveAvgPool2d: row-major window sum / count, the key axis first)
//   get_attention_rollout [:111-146] + identity, rows normalised by their ATen-order sums, chained with bmm
//
// A block's attention is [B, heads, T, T'] fp32 -- 630 MB for 64 images at CvT's first stage -- so everything up to the
// pooled [B, g^2, g^2] map is byte work bound by HBM: one pass fuses the heads, stores the fused map and counts the top
// 11 key bits per image; two more passes over the fused map finish an exact radix select of the k-th smallest value
// (ties: lowest index first); one pass marks the union mask; the pooling pass reads the fused map once more.  The chain
// (13 matrices of 49 x 49 per image) is one small CTA per image with the FMA chains of the CPU bmm.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vr {

namespace {

constexpr int RO_THREADS = 256;
constexpr int RO_ITEMS = 16;    // elements per thread and CTA pass
constexpr int RO_BINS = 2048;

struct RoSelect {          // per image
    uint32_t prefix;       // key bits decided so far
    uint32_t krem;         // rank (1-based) of the wanted element among the keys that share the prefix
    uint32_t eq;           // after the last pass: how many keys equal the threshold
    uint32_t pad;
};

__device__ __forceinline__ float fuse_heads(const float* __restrict__ p, int heads, int64_t head_stride, int mode) {
    float m = __ldg(p);
    for (int h = 1; h < heads; h++) {
        const float x = __ldg(p + h * head_stride);
        // torch.min / torch.max propagate NaN
        if (mode == 2) m = (x < m || x != x) ? x : m;
        else m = (x > m || x != x) ? x : m;
    }
    return m;
}

// PASS 0: fuse the heads, store the fused map, histogram of key bits 31..21.  PASS 1 / 2: histogram of bits 20..10 / 9..0 of the
// keys that share the prefix found so far.  VEC: 16-byte loads and stores (maps whose size is a multiple of 4 floats).
template <int PASS>
__device__ __forceinline__ void ro_count(uint32_t* sh, uint32_t key, uint32_t prefix) {
    if (PASS == 0) {
        atomicAdd(&sh[key >> 21], 1u);   // (warp-aggregating equal bins with match.any first was twice as slow: 770 against 381 us)
    } else if (PASS == 1) {
        if ((key >> 21) == prefix) atomicAdd(&sh[(key >> 10) & 2047u], 1u);
    } else {
        if ((key >> 10) == prefix) atomicAdd(&sh[key & 1023u], 1u);
    }
}

template <int PASS, bool VEC>
__global__ void __launch_bounds__(RO_THREADS) rollout_hist_kernel(const float* __restrict__ probs, float* __restrict__ fused,
                                                                   uint32_t* __restrict__ hist, const RoSelect* __restrict__ sel, int heads,
                                                                   int ht, int wt, int mode) {
    __shared__ uint32_t sh[RO_BINS];
    const int b = blockIdx.y;
    const int64_t hw = (int64_t)ht * wt;   // the whole map, cls row / column included: the reference filters before it drops them
    for (int i = threadIdx.x; i < RO_BINS; i += RO_THREADS) sh[i] = 0u;
    __syncthreads();
    const uint32_t prefix = PASS ? sel[b].prefix : 0u;
    const int64_t base = (int64_t)blockIdx.x * (RO_THREADS * RO_ITEMS);
    if (VEC) {
        for (int it = 0; it < RO_ITEMS / 4; it++) {
            const int64_t i = base + ((int64_t)it * RO_THREADS + threadIdx.x) * 4;
            if (i >= hw) break;
            float4 v;
            if (PASS == 0) {
                const float* p = probs + (int64_t)b * heads * hw + i;
                v = __ldcs(reinterpret_cast<const float4*>(p));
                for (int h = 1; h < heads; h++) {
                    const float4 x = __ldcs(reinterpret_cast<const float4*>(p + h * hw));
                    if (mode == 2) {   // torch.min / torch.max propagate NaN
                        v.x = (x.x < v.x || x.x != x.x) ? x.x : v.x; v.y = (x.y < v.y || x.y != x.y) ? x.y : v.y;
                        v.z = (x.z < v.z || x.z != x.z) ? x.z : v.z; v.w = (x.w < v.w || x.w != x.w) ? x.w : v.w;
                    } else {
                        v.x = (x.x > v.x || x.x != x.x) ? x.x : v.x; v.y = (x.y > v.y || x.y != x.y) ? x.y : v.y;
                        v.z = (x.z > v.z || x.z != x.z) ? x.z : v.z; v.w = (x.w > v.w || x.w != x.w) ? x.w : v.w;
                    }
                }
                *reinterpret_cast<float4*>(fused + (int64_t)b * hw + i) = v;
            } else {
                v = *reinterpret_cast<const float4*>(fused + (int64_t)b * hw + i);
            }
            ro_count<PASS>(sh, ordered_bits(v.x), prefix);
            ro_count<PASS>(sh, ordered_bits(v.y), prefix);
            ro_count<PASS>(sh, ordered_bits(v.z), prefix);
            ro_count<PASS>(sh, ordered_bits(v.w), prefix);
        }
    } else {
        for (int it = 0; it < RO_ITEMS; it++) {
            const int64_t i = base + (int64_t)it * RO_THREADS + threadIdx.x;
            if (i >= hw) break;
            float v;
            if (PASS == 0) {
                v = fuse_heads(probs + (int64_t)b * heads * hw + i, heads, hw, mode);
                fused[(int64_t)b * hw + i] = v;
            } else {
                v = fused[(int64_t)b * hw + i];
            }
            ro_count<PASS>(sh, ordered_bits(v), prefix);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RO_BINS; i += RO_THREADS)
        if (sh[i]) atomicAdd(&hist[(int64_t)b * RO_BINS + i], sh[i]);
}

// The bin that holds the krem-th smallest key; the histogram is cleared for the next pass.
template <int PASS>
__global__ void __launch_bounds__(RO_THREADS) rollout_pick_kernel(uint32_t* __restrict__ hist, RoSelect* __restrict__ sel, uint32_t k) {
    __shared__ uint32_t part[RO_THREADS];
    const int b = blockIdx.x, t = threadIdx.x;
    uint32_t* h = hist + (int64_t)b * RO_BINS;
    constexpr int PER = RO_BINS / RO_THREADS;
    uint32_t c[PER], s = 0;
#pragma unroll
    for (int i = 0; i < PER; i++) {
        c[i] = h[t * PER + i];
        s += c[i];
        h[t * PER + i] = 0u;
    }
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        const uint32_t want = PASS ? sel[b].krem : k;
        uint32_t before = 0;
        int tt = 0;
        while (tt < RO_THREADS - 1 && before + part[tt] < want) before += part[tt++];
        part[0] = (uint32_t)tt;
        part[1] = before;
        part[2] = want;
    }
    __syncthreads();
    if (t == (int)part[0]) {
        uint32_t before = part[1];
        const uint32_t want = part[2];
        int i = 0;
        while (i < PER - 1 && before + c[i] < want) before += c[i++];
        const uint32_t bin = (uint32_t)(t * PER + i);
        RoSelect r = sel[b];
        r.prefix = PASS == 0 ? bin : (PASS == 1 ? ((r.prefix << 11) | bin) : ((r.prefix << 10) | bin));
        r.krem = want - before;
        r.eq = c[i];
        sel[b] = r;
    }
}

// mask[i] = 1 where image b discards coordinate i: below the threshold, or equal to it when all equals are taken.
template <bool VEC>
__global__ void __launch_bounds__(RO_THREADS) rollout_mask_kernel(const float* __restrict__ fused, const RoSelect* __restrict__ sel,
                                                                   unsigned char* __restrict__ mask, int64_t hw) {
    const int b = blockIdx.y;
    const RoSelect s = sel[b];
    const bool all_eq = s.krem == s.eq;
    const int64_t base = (int64_t)blockIdx.x * (RO_THREADS * RO_ITEMS);
    auto hit = [&](float v) {
        const uint32_t key = ordered_bits(v);
        return key < s.prefix || (key == s.prefix && all_eq);
    };
    if (VEC) {
        for (int it = 0; it < RO_ITEMS / 4; it++) {
            const int64_t i = base + ((int64_t)it * RO_THREADS + threadIdx.x) * 4;
            if (i >= hw) break;
            const float4 v = *reinterpret_cast<const float4*>(fused + (int64_t)b * hw + i);
            if (hit(v.x)) mask[i] = 1;
            if (hit(v.y)) mask[i + 1] = 1;
            if (hit(v.z)) mask[i + 2] = 1;
            if (hit(v.w)) mask[i + 3] = 1;
        }
    } else {
        for (int it = 0; it < RO_ITEMS; it++) {
            const int64_t i = base + (int64_t)it * RO_THREADS + threadIdx.x;
            if (i >= hw) break;
            if (hit(fused[(int64_t)b * hw + i])) mask[i] = 1;
        }
    }
}

// Only some of the keys equal to the threshold belong to the k smallest: the first krem of them in index order.
__global__ void __launch_bounds__(RO_THREADS) rollout_ties_kernel(const float* __restrict__ fused, const RoSelect* __restrict__ sel,
                                                                   unsigned char* __restrict__ mask, int64_t hw) {
    __shared__ uint32_t wcount[RO_THREADS / 32];
    __shared__ uint32_t taken;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const RoSelect s = sel[b];
    if (s.krem == s.eq) return;
    if (t == 0) taken = 0;
    __syncthreads();
    for (int64_t base = 0; base < hw; base += RO_THREADS) {
        const int64_t i = base + t;
        const bool eq = i < hw && ordered_bits(fused[(int64_t)b * hw + i]) == s.prefix;
        const uint32_t bal = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) wcount[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = taken;
        for (int w = 0; w < warp; w++) before += wcount[w];
        before += __popc(bal & ((1u << lane) - 1u));
        if (eq && before < s.krem) mask[i] = 1;
        __syncthreads();
        if (t == 0) {
            uint32_t tot = taken;
            for (int w = 0; w < RO_THREADS / 32; w++) tot += wcount[w];
            taken = tot;
        }
        __syncthreads();
        if (taken >= s.krem) return;
    }
}

// (o < out <= 16 and in < 2^16: 32-bit arithmetic holds the products)
__device__ __forceinline__ int win_lo(int o, int in, int out) { return (o * in) / out; }
__device__ __forceinline__ int win_hi(int o, int in, int out) { return ((o + 1) * in + out - 1) / out; }

// out[b][hp][wp]: the masked fused map pooled over the key axis (rows of ws x ws -> g x g, when ws > g), then over the query axis
// (hs x hs -> g x g, when hs > g).  One CTA per (hp, image): the query rows of the window are staged a few at a time (coalesced, the
// mask applied), every (row, wp) key-window sum / kh / kw is taken by one thread and kept in shared memory, and thread wp < g^2
// finally adds its column over the rows in row-major order and divides -- AdaptiveAvgPool2d's arithmetic on the CPU.
__global__ void __launch_bounds__(RO_THREADS) rollout_pool_kernel(const float* __restrict__ fused, const unsigned char* __restrict__ mask,
                                                                   float* __restrict__ out, int H, int W, int hs, int ws, int g, int drop,
                                                                   int hsingle, int rc) {
    extern __shared__ float ro_sm[];
    const int b = blockIdx.y, hp = blockIdx.x, t = threadIdx.x;
    const int g2h = hs > g ? g * g : H, g2w = ws > g ? g * g : W;   // (an axis that is already g x g is kept)
    const int wt = W + drop;
    const int64_t hw = (int64_t)(H + drop) * wt;                     // fused and mask cover the whole map; row / column 0 are skipped
    int hy0, hy1, hx0, hx1;
    if (hs > g) {
        const int oy = hp / g, ox = hp % g;
        hy0 = win_lo(oy, hs, g); hy1 = win_hi(oy, hs, g);
        hx0 = win_lo(ox, hs, g); hx1 = win_hi(ox, hs, g);
    } else {
        hy0 = hp / hs; hy1 = hy0 + 1;
        hx0 = hp % hs; hx1 = hx0 + 1;
    }
    const int kwh = hx1 - hx0, nrows = (hy1 - hy0) * kwh;
    float* rows = ro_sm;                       // [rc][W]
    float* out1 = ro_sm + (size_t)rc * W;      // [nrows][g2w]
    __shared__ int wwin[RO_THREADS][4];        // the key window of every output cell
    if (t < g2w && ws > g) {
        const int oy = t / g, ox = t % g;
        wwin[t][0] = win_lo(oy, ws, g); wwin[t][1] = win_hi(oy, ws, g);
        wwin[t][2] = win_lo(ox, ws, g); wwin[t][3] = win_hi(ox, ws, g);
    }
    for (int r0 = 0; r0 < nrows; r0 += rc) {
        const int nr = min(rc, nrows - r0);
        __syncthreads();
        const bool vec = drop == 0 && (W & 3) == 0;   // rows start on 16-byte boundaries
        const int lane = t & 31, warp = t >> 5;
        for (int rr = warp; rr < nr; rr += RO_THREADS / 32) {   // a warp per query row: the row's address is computed once
            const int ri = r0 + rr;
            const int64_t h = (int64_t)(hy0 + ri / kwh) * hs + (hx0 + ri % kwh);
            const float* fr = fused + (int64_t)b * hw + (h + drop) * wt + drop;
            const unsigned char* mr = mask + (h + drop) * wt + drop;
            float* dst = rows + (size_t)rr * W;
            if (vec) {
#pragma unroll 4
                for (int w4 = lane; w4 < (W >> 2); w4 += 32) {
                    float4 v = __ldcs(reinterpret_cast<const float4*>(fr) + w4);
                    const uchar4 m = *(reinterpret_cast<const uchar4*>(mr) + w4);
                    if (m.x) v.x = 0.f;
                    if (m.y) v.y = 0.f;
                    if (m.z) v.z = 0.f;
                    if (m.w) v.w = 0.f;
                    *(reinterpret_cast<float4*>(dst) + w4) = v;
                }
            } else {
#pragma unroll 4
                for (int w = lane; w < W; w += 32) {
                    const float v = __ldcs(fr + w);
                    dst[w] = mr[w] ? 0.f : v;
                }
            }
        }
        __syncthreads();
        for (int e = t; e < nr * g2w; e += RO_THREADS) {
            const int rr = e / g2w, wp = e - rr * g2w;
            const float* rb = rows + (size_t)rr * W;
            float v;
            if (ws > g) {
                const int wy0 = wwin[wp][0], wy1 = wwin[wp][1], wx0 = wwin[wp][2], wx1 = wwin[wp][3];
                float s = 0.f;
                for (int y = wy0; y < wy1; y++)
                    for (int x = wx0; x < wx1; x++) s += rb[y * ws + x];
                v = (s / (float)(wy1 - wy0)) / (float)(wx1 - wx0);   // (ATen: sum / kh / kw)
            } else {
                v = rb[wp];
            }
            out1[(size_t)(r0 + rr) * g2w + wp] = v;
        }
    }
    __syncthreads();
    if (t < g2w) {
        float acc = 0.f;
        for (int r = 0; r < nrows; r++) acc += out1[(size_t)r * g2w + t];
        // The reference pools the query axis on a permuted VIEW [B, g^2, hs, hs] whose strides are channels-last: for B >= 2 ATen
        // then runs its channels-last kernel, which divides the window sum ONCE by kh * kw; for B = 1 the view counts as
        // contiguous and the plain kernel divides by kh, then by kw -- as does the channels-last kernel's scalar tail, the
        // channels beyond the last whole vector of 8.  (All equal for the power-of-two windows of every reference model.)
        if (hs > g)
            acc = (hsingle && t < (g2w & ~7)) ? acc / (float)((hy1 - hy0) * kwh) : (acc / (float)(hy1 - hy0)) / (float)kwh;
        out[((int64_t)b * g2h + hp) * g2w + t] = acc;
    }
}

// joint[0] = M0, joint[j] = Mj joint[j - 1] with Mj = (mats[j] + I) / rowsum when use_res: one CTA per image, the matrices in
// shared memory, every product entry one FMA chain over the inner index.
__global__ void __launch_bounds__(RO_THREADS) rollout_chain_kernel(const float* __restrict__ mats, int J, int64_t B, int n, int use_res,
                                                                    float* __restrict__ joints) {
    extern __shared__ float sm[];
    float* M = sm;                 // [n][n + 1]
    float* P = M + n * (n + 1);    // previous joint
    float* Q = P + n * (n + 1);    // next joint
    float* rs = Q + n * (n + 1);   // [n] row sums
    const int64_t b = blockIdx.x;
    const int t = threadIdx.x, ld = n + 1;
    for (int j = 0; j < J; j++) {
        const float* src = mats + ((int64_t)j * B + b) * n * n;
        for (int i = t; i < n * n; i += RO_THREADS) {
            const int r = i / n, c = i - r * n;
            float x = src[i];
            if (use_res) x = x + (r == c ? 1.0f : 0.0f);
            M[r * ld + c] = x;
        }
        __syncthreads();
        if (use_res) {
            for (int r = t; r < n; r += RO_THREADS) rs[r] = torch_sum_inner(M + r * ld, n);
            __syncthreads();
            for (int i = t; i < n * n; i += RO_THREADS) {
                const int r = i / n, c = i - r * n;
                M[r * ld + c] = M[r * ld + c] / rs[r];
            }
            __syncthreads();
        }
        float* dst = joints + ((int64_t)j * B + b) * n * n;
        if (j == 0) {
            for (int i = t; i < n * n; i += RO_THREADS) {
                const int r = i / n, c = i - r * n;
                const float x = M[r * ld + c];
                P[r * ld + c] = x;
                dst[i] = x;
            }
        } else {
            for (int i = t; i < n * n; i += RO_THREADS) {
                const int r = i / n, c = i - r * n;
                float a = 0.f;
                for (int k = 0; k < n; k++) a = fmaf(M[r * ld + k], P[k * ld + c], a);
                Q[r * ld + c] = a;
                dst[i] = a;
            }
            __syncthreads();
            float* tmp = P;
            P = Q;
            Q = tmp;
        }
        __syncthreads();
    }
}

struct RoWs {
    float* fused;
    uint32_t* hist;
    RoSelect* sel;
    unsigned char* mask;
    size_t bytes;
};
RoWs ro_carve(void* base, int64_t b, int64_t hw) {
    RoWs w{};
    size_t off = 0;
    auto take = [&](size_t n) {
        size_t o = off;
        off = align_up(off + n, 256);
        return base ? reinterpret_cast<unsigned char*>(base) + o : nullptr;
    };
    w.fused = reinterpret_cast<float*>(take((size_t)b * hw * 4));
    w.hist = reinterpret_cast<uint32_t*>(take((size_t)b * RO_BINS * 4));
    w.sel = reinterpret_cast<RoSelect*>(take((size_t)b * sizeof(RoSelect)));
    w.mask = take((size_t)hw);
    w.bytes = off + 256;
    return w;
}

int isqrt_exact(int x) {
    int s = 0;
    while ((int64_t)(s + 1) * (s + 1) <= x) s++;
    return s;
}

}  // namespace

size_t rollout_block_workspace_bytes(int64_t b, int ht, int wt, int drop_cls) {
    (void)drop_cls;
    return ro_carve(nullptr, b, (int64_t)ht * wt).bytes;
}

int rollout_block(const float* probs, int64_t b, int heads, int ht, int wt, int drop_cls, int grid, int64_t n_discard, int fusion,
                  float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    VR_REQUIRE(probs && out && b > 0 && b < 65536 && heads > 0 && grid > 0, "rollout_block: bad arguments");
    VR_REQUIRE(fusion == 1 || fusion == 2, "rollout_block: head fusion must be max (1) or min (2)");
    drop_cls = drop_cls ? 1 : 0;
    const int H = ht - drop_cls, W = wt - drop_cls;
    VR_REQUIRE(H > 0 && W > 0, "rollout_block: empty map");
    const int hs = isqrt_exact(H), ws_ = isqrt_exact(W);
    VR_REQUIRE(hs * hs == H && ws_ * ws_ == W, "rollout_block: %d x %d tokens are not square grids", H, W);
    VR_REQUIRE(hs >= grid && ws_ >= grid && grid * grid <= RO_THREADS, "rollout_block: token grids %d / %d against a %d x %d target", hs, ws_,
               grid, grid);
    const int64_t hw = (int64_t)ht * wt;
    VR_REQUIRE(n_discard >= 0 && n_discard <= hw && hw < ((int64_t)1 << 32), "rollout_block: discard count out of range");
    RoWs w = ro_carve(ws, b, hw);
    if (w.bytes > ws_bytes) {
        set_error("rollout_block: workspace %zu < %zu", ws_bytes, w.bytes);
        return VR_E_WORKSPACE;
    }
    const dim3 grid_e((unsigned)((hw + RO_THREADS * RO_ITEMS - 1) / (RO_THREADS * RO_ITEMS)), (unsigned)b);
    VR_CHECK_CUDA(cudaMemsetAsync(w.hist, 0, (size_t)b * RO_BINS * 4, st));
    VR_CHECK_CUDA(cudaMemsetAsync(w.mask, 0, (size_t)hw, st));
    VR_CHECK_CUDA(cudaMemsetAsync(w.sel, 0, (size_t)b * sizeof(RoSelect), st));
    const bool vec = (hw & 3) == 0 && (reinterpret_cast<uintptr_t>(probs) & 15) == 0;
#define RO_HIST(PASS)                                                                                                          \
    do {                                                                                                                       \
        if (vec) rollout_hist_kernel<PASS, true><<<grid_e, RO_THREADS, 0, st>>>(probs, w.fused, w.hist, w.sel, heads, ht, wt, fusion); \
        else rollout_hist_kernel<PASS, false><<<grid_e, RO_THREADS, 0, st>>>(probs, w.fused, w.hist, w.sel, heads, ht, wt, fusion);    \
        VR_LAUNCH_CHECK();                                                                                                     \
    } while (0)
    RO_HIST(0);
    if (n_discard > 0) {
        rollout_pick_kernel<0><<<(unsigned)b, RO_THREADS, 0, st>>>(w.hist, w.sel, (uint32_t)n_discard);
        VR_LAUNCH_CHECK();
        RO_HIST(1);
        rollout_pick_kernel<1><<<(unsigned)b, RO_THREADS, 0, st>>>(w.hist, w.sel, 0u);
        VR_LAUNCH_CHECK();
        RO_HIST(2);
        rollout_pick_kernel<2><<<(unsigned)b, RO_THREADS, 0, st>>>(w.hist, w.sel, 0u);
        VR_LAUNCH_CHECK();
        if (vec) rollout_mask_kernel<true><<<grid_e, RO_THREADS, 0, st>>>(w.fused, w.sel, w.mask, hw);
        else rollout_mask_kernel<false><<<grid_e, RO_THREADS, 0, st>>>(w.fused, w.sel, w.mask, hw);
        VR_LAUNCH_CHECK();
        rollout_ties_kernel<<<(unsigned)b, RO_THREADS, 0, st>>>(w.fused, w.sel, w.mask, hw);
        VR_LAUNCH_CHECK();
    }
#undef RO_HIST
    const int g2h = hs > grid ? grid * grid : H, g2w = ws_ > grid ? grid * grid : W;
    const int kmax = hs > grid ? (hs + grid - 1) / grid + 1 : 1;     // longest window side of the query axis
    const int nrows_max = kmax * kmax;
    const int rc = std::max(1, std::min(nrows_max, 4096 / W));      // query rows staged at a time (<= 16 KB: 6 CTAs per SM)
    const size_t smem = ((size_t)rc * W + (size_t)nrows_max * g2w) * 4;
    VR_REQUIRE(smem <= 200 * 1024, "rollout_block: %d x %d tokens need %zu bytes of shared memory", H, W, smem);
    if (smem > 40 * 1024)   // (the kernel also holds 4 KB of static shared memory)
        VR_CHECK_CUDA(cudaFuncSetAttribute(rollout_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rollout_pool_kernel<<<dim3((unsigned)g2h, (unsigned)b), RO_THREADS, smem, st>>>(w.fused, w.mask, out, H, W, hs, ws_, grid, drop_cls,
                                                                                     b >= 2 ? 1 : 0, rc);
    VR_LAUNCH_CHECK();
    return VR_OK;
}

int rollout_chain(const float* mats, int n_mats, int64_t b, int n, int use_res, float* joints, cudaStream_t st) {
    VR_REQUIRE(mats && joints && n_mats > 0 && b > 0 && n > 0 && n <= 128, "rollout_chain: bad arguments (n <= 128)");
    const size_t smem = ((size_t)3 * n * (n + 1) + n) * 4;
    if (smem > 48 * 1024) VR_CHECK_CUDA(cudaFuncSetAttribute(rollout_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rollout_chain_kernel<<<(unsigned)b, RO_THREADS, smem, st>>>(mats, n_mats, b, n, use_res ? 1 : 0, joints);
    VR_LAUNCH_CHECK();
    return VR_OK;
}

}  // namespace vr
