// Shape-generic OT path: any (C, R), any number of candidates per query, any [b, m, n] for
// the direct Sinkhorn call.  Same arithmetic as pair_fused.cu, but the Gibbs kernel, sim and
// the scaling vectors live in a caller-provided global workspace (L2-resident for typical
// sizes) and the batch-global stop of utilities/diml.py:50-52 is taken between launches:
//
//   prepare   (1 launch)   sim, K(_ext), u, v per pair            diml.py:100-133 / :339-354
//   iterate   (<= max_iter) r = u/(K c), c = v/(K^T r), sum|dr|   diml.py:47-49
//   decide    (<= max_iter) per-query mean over all pairs < thresh -> done flag   :50-52
//   finish    (1 launch)   T = r c^T * K, sim_r, score            diml.py:53,142-143
//
// iterate/decide pairs for iterations after a query stopped return immediately on its
// `done` flag, so no host synchronisation is needed.  Used for everything the fused
// R=49/C=128/K<=104 kernel does not cover (K=1000 shortlists, 14x14 grids, C=768, the
// direct Sinkhorn() call).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vr {


__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; i++) t += red[i];
    return t;
}
__device__ __forceinline__ float block_reduce_max(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = -INFINITY;
    for (int i = 0; i < nw; i++) t = fmaxf(t, red[i]);
    return t;
}

constexpr int GP_THREADS = 256;
constexpr int GP_T = 64;   // output tile
constexpr int GP_KC = 16;  // channels per step

__global__ void __launch_bounds__(GP_THREADS) generic_prepare_kernel(GenArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Fs = reinterpret_cast<float*>(smem_raw);  // [KC][T]
    float* As = Fs + GP_KC * GP_T;                   // [KC][T]
    float* red = As + GP_KC * GP_T;                  // [32]
    float* qcs = red + 32;                           // [C]
    float* gcs = qcs + a.c;                          // [C]
    float* ccu = gcs + a.c;                          // [R]
    float* ccv = ccu + a.r;                          // [R]

    const int tid = threadIdx.x;
    const int64_t pair = blockIdx.x;
    const int64_t qi = pair / a.k;
    const int pi = (int)(pair % a.k);
    const int64_t qid = a.q_start + qi * a.q_stride;
    const int C = a.c, R = a.r;
    const bool full = a.p.ot_part > 0.999f;
    const int Re = full ? R : R + 1;
    const float bins = a.part_bin;   // 1 - ot_part, rounded as the reference rounds it (partial_ot_bin)
    const int mode = a.p.mode;
    if (pi == 0 && tid == 0) {
        a.done[qi] = 0;
        a.niter[qi] = 0;
    }
    const int cand = a.cand_idx ? a.cand_idx[qi * a.cand_stride + pi] : pi;
    float* uo = a.u + pair * Re;
    float* vo = a.v + pair * Re;
    for (int i = tid; i < Re; i += GP_THREADS) {
        a.rv[pair * Re + i] = 1.f;
        a.cv[pair * Re + i] = 1.f;
    }
    if (tid == 0) a.e[pair] = 0.f;
    if (cand < 0) {  // padded shortlist entry: a zero problem that never contributes
        for (int i = tid; i < Re; i += GP_THREADS) {
            uo[i] = 0.f;
            vo[i] = 0.f;
            a.rv[pair * Re + i] = 0.f;
            a.cv[pair * Re + i] = 0.f;
        }
        if (a.K)   // (nullptr: marginals only, for generic_fused.cu)
            for (int i = tid; i < Re * Re; i += GP_THREADS) a.K[pair * Re * Re + i] = 0.f;
        if (a.sim)
            for (int i = tid; i < R * R; i += GP_THREADS) a.sim[pair * R * R + i] = 0.f;
        if (tid == 0) a.e[pair] = -1.f;  // marks the pair as skipped for iterate / decide
        return;
    }
    const float* Ag = a.q_patches + qid * (int64_t)C * R;
    const float* Fg = a.c_patches + (int64_t)cand * C * R;
    float* simo = a.sim ? a.sim + pair * (int64_t)R * R : nullptr;   // (nullptr: marginals only, for generic_fused.cu)
    float* Ko = a.K ? a.K + pair * (int64_t)Re * Re : nullptr;

    // ---- sim[s][m] = sum_c F[c][s] * A[c][m], K = exp(-(1 - sim) / ot_temp) ----
    const int tx = tid & 15, ty = tid >> 4;
    for (int s0 = 0; s0 < (a.sim_done ? 0 : R); s0 += GP_T) {   // (tensor-core S3 already wrote sim and K: generic_s3.cu)
        for (int m0 = 0; m0 < R; m0 += GP_T) {
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
            for (int c0 = 0; c0 < C; c0 += GP_KC) {
                __syncthreads();
                for (int e = tid; e < GP_KC * GP_T; e += GP_THREADS) {
                    const int kk = e / GP_T, x = e % GP_T;
                    const int c = c0 + kk;
                    Fs[e] = (c < C && s0 + x < R) ? Fg[(int64_t)c * R + s0 + x] : 0.f;
                    As[e] = (c < C && m0 + x < R) ? Ag[(int64_t)c * R + m0 + x] : 0.f;
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < GP_KC; kk++) {
                    const float4 f = *reinterpret_cast<const float4*>(Fs + kk * GP_T + ty * 4);
                    const float4 av = *reinterpret_cast<const float4*>(As + kk * GP_T + tx * 4);
                    const float fv[4] = {f.x, f.y, f.z, f.w};
                    const float aw[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(fv[i], aw[j], acc[i][j]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int s = s0 + ty * 4 + i, m = m0 + tx * 4 + j;
                    if (s < R && m < R) {
                        simo[(int64_t)s * R + m] = acc[i][j];
                        Ko[(int64_t)s * Re + m] = expf(-(1.0f - acc[i][j]) / a.p.ot_temp);
                    }
                }
        }
    }
    if (!full && a.K) {  // utilities/diml.py:62-73
        for (int i = tid; i < R; i += GP_THREADS) {
            Ko[(int64_t)i * Re + R] = bins;
            Ko[(int64_t)R * Re + i] = bins;
        }
        if (tid == 0) Ko[(int64_t)R * Re + R] = 0.f;
    }

    // ---- marginals ----
    const bool need_cc = mode >= VR_MODE_INVERSE;
    const bool cls = a.p.use_cls_token != 0;
    if (need_cc) {
        float nq = 0.f, ng = 0.f;
        for (int c = tid; c < C; c += GP_THREADS) {
            float x, g;
            if (cls) {
                x = a.q_centers[qid * C + c];
                g = a.c_centers[(int64_t)cand * C + c];
            } else {
                float sx = 0.f, sg = 0.f;
                for (int m = 0; m < R; m++) {
                    sx += Ag[(int64_t)c * R + m];
                    sg += Fg[(int64_t)c * R + m];
                }
                x = sx / (float)R;
                g = sg / (float)R;
            }
            qcs[c] = x;
            gcs[c] = g;
            nq += x * x;
            ng += g * g;
        }
        nq = block_reduce_sum(nq, red);
        ng = block_reduce_sum(ng, red);
        const float dq = fmaxf(sqrtf(nq), 1e-12f), dg = fmaxf(sqrtf(ng), 1e-12f);
        for (int c = tid; c < C; c += GP_THREADS) {
            qcs[c] = qcs[c] / dq;
            gcs[c] = gcs[c] / dg;
        }
        __syncthreads();
        if (a.cc_stages > 0) {
            // The two mat-vecs read both fp32 images once (1.2 MB at C = 768, R = 196): the 16 channel rows of a chunk are contiguous
            // in both banks, so one thread streams them through a ring of bulk copies while every thread extends its two chains --
            // the same FMA order as the loop below, at the speed the rows arrive instead of a dependent global load per FMA.
            const int NS = a.cc_stages, NCH = C / GP_KC;
            const uint32_t chunk_bytes = (uint32_t)(GP_KC * R * 4);
            float* ring = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ccv + R) + 15) & ~(uintptr_t)15);   // [NS][2][16 * R]
            uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)NS * 2 * GP_KC * R);   // [NS]
            if (tid == 0) {
                for (int i = 0; i < NS; i++) mbar_init(full + i, 1);
                fence_mbar_init();
            }
            __syncthreads();
            if (tid == 0)
                for (int i = 0; i < NS && i < NCH; i++) {
                    mbar_expect_tx(full + i, 2 * chunk_bytes);
                    bulk_g2s(ring + (size_t)(2 * i) * GP_KC * R, Fg + (int64_t)i * GP_KC * R, chunk_bytes, full + i);
                    bulk_g2s(ring + (size_t)(2 * i + 1) * GP_KC * R, Ag + (int64_t)i * GP_KC * R, chunk_bytes, full + i);
                }
            float xs[4] = {0.f, 0.f, 0.f, 0.f}, ys[4] = {0.f, 0.f, 0.f, 0.f};   // R <= 1,024: up to 4 patches per thread
            for (int ch = 0; ch < NCH; ch++) {
                const int stg = ch % NS;
                mbar_wait(full + stg, (uint32_t)(ch / NS) & 1u);
                const float* Fr = ring + (size_t)(2 * stg) * GP_KC * R;
                const float* Ar = Fr + (size_t)GP_KC * R;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int s = tid + j * GP_THREADS;
                    if (s < R) {
#pragma unroll
                        for (int kk = 0; kk < GP_KC; kk++) {
                            xs[j] = fmaf(qcs[ch * GP_KC + kk], Fr[kk * R + s], xs[j]);
                            ys[j] = fmaf(Ar[kk * R + s], gcs[ch * GP_KC + kk], ys[j]);
                        }
                    }
                }
                __syncthreads();   // the stage is free
                if (tid == 0 && ch + NS < NCH) {
                    mbar_expect_tx(full + stg, 2 * chunk_bytes);
                    bulk_g2s(ring + (size_t)(2 * stg) * GP_KC * R, Fg + (int64_t)(ch + NS) * GP_KC * R, chunk_bytes, full + stg);
                    bulk_g2s(ring + (size_t)(2 * stg + 1) * GP_KC * R, Ag + (int64_t)(ch + NS) * GP_KC * R, chunk_bytes, full + stg);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int s = tid + j * GP_THREADS;
                if (s < R) {
                    ccu[s] = xs[j];
                    ccv[s] = ys[j];
                }
            }
        } else {
            for (int s = tid; s < R; s += GP_THREADS) {
                float x = 0.f, y = 0.f;
                for (int c = 0; c < C; c++) {
                    x = fmaf(qcs[c], Fg[(int64_t)c * R + s], x);
                    y = fmaf(Ag[(int64_t)c * R + s], gcs[c], y);
                }
                ccu[s] = x;
                ccv[s] = y;
            }
        }
        __syncthreads();
    }
    // numerators, strided over threads; sums via block reductions
    float mu = 0.f, mv = 0.f;
    if (mode == VR_MODE_SOFT) {
        float lu = -INFINITY, lv = -INFINITY;
        for (int s = tid; s < R; s += GP_THREADS) {
            lu = fmaxf(lu, ccu[s]);
            lv = fmaxf(lv, ccv[s]);
        }
        mu = block_reduce_max(lu, red);
        mv = block_reduce_max(lv, red);
    }
    for (int s = tid; s < R; s += GP_THREADS) {
        float x, y;
        switch (mode) {
            case VR_MODE_UNIFORM: x = y = (float)(1.0 / (double)R); break;
            case VR_MODE_ROLLOUT:
                x = fmaxf(a.c_rollout[(int64_t)cand * R + s], 0.f);
                y = fmaxf(a.q_rollout[qid * R + s], 0.f);
                break;
            case VR_MODE_INVERSE:
                x = expf(-fmaxf(ccu[s], 0.f) / a.p.temperature);
                y = expf(-fmaxf(ccv[s], 0.f) / a.p.temperature);
                break;
            case VR_MODE_MINUS:
                x = 1.f - fmaxf(ccu[s], 0.f);
                y = 1.f - fmaxf(ccv[s], 0.f);
                break;
            case VR_MODE_SOFT:
                x = expf(ccu[s] - mu);
                y = expf(ccv[s] - mv);
                break;
            default:
                x = fmaxf(ccu[s], 0.f);
                y = fmaxf(ccv[s], 0.f);
                break;
        }
        uo[s] = x;
        vo[s] = y;
    }
    // row sums in torch's reduction order (common.cuh: torch_sum_inner); one thread each
    auto row_sums = [&](float& su, float& sv) {
        __syncthreads();
        if (tid == 0) red[0] = torch_sum_inner(uo, R);
        if (tid == 32) red[1] = torch_sum_inner(vo, R);
        __syncthreads();
        su = red[0];
        sv = red[1];
        __syncthreads();
    };
    float su = 0.f, sv = 0.f;
    if (mode == VR_MODE_SOFT) {  // softmax, then the common /(sum + 1e-5)
        row_sums(su, sv);
        for (int s = tid; s < R; s += GP_THREADS) {
            uo[s] = uo[s] / su;
            vo[s] = vo[s] / sv;
        }
    }
    if (mode != VR_MODE_UNIFORM) {
        row_sums(su, sv);
        su += 1e-5f;
        sv += 1e-5f;
        for (int s = tid; s < R; s += GP_THREADS) {
            uo[s] = uo[s] / su;
            vo[s] = vo[s] / sv;
        }
    }
    if (!full && tid == 0) {
        uo[R] = bins;
        vo[R] = bins;
    }
    __syncthreads();
    if (a.out_u)
        for (int s = tid; s < R; s += GP_THREADS) {
            a.out_u[pair * R + s] = uo[s];
            a.out_v[pair * R + s] = vo[s];
        }
    if (a.out_cc && (mode == VR_MODE_MINUS || mode == VR_MODE_SOFT || mode == VR_MODE_RELU))
        for (int s = tid; s < R; s += GP_THREADS) a.out_cc[pair * R + s] = (mode == VR_MODE_MINUS) ? ccu[s] : ccv[s];
}

// One Sinkhorn iteration for every pair whose query has not stopped.  rows x cols problem.
struct IterArgs {
    const float* K;   // [np, rows, cols]
    const float* u;   // [np, rows]
    const float* v;   // [np, cols]
    float* rv;        // [np, rows]
    float* cv;        // [np, cols]
    float* e;         // [np]
    int32_t* done;    // [nq]
    int32_t* niter;   // [nq]
    int rows, cols, k;
    float thresh;
};

constexpr int GI_THREADS = 256;

// Same arithmetic order as the reference's CPU bmm (and as pair_fused.cu): every output is one
// sequential FMA chain over the inner index, followed by an IEEE division.  K is staged in
// shared memory with a padded row (conflict-free for both passes) when it fits, else read
// from global memory.
template <bool STAGED>
__global__ void __launch_bounds__(GI_THREADS) generic_iter_kernel(IterArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* cs = reinterpret_cast<float*>(smem_raw);  // [cols]
    float* rs = cs + a.cols;                          // [rows]
    float* red = rs + a.rows;                         // [32]
    float* Ks = red + 32;                             // [rows][cols + 1] when STAGED
    const int64_t pair = blockIdx.x;
    const int64_t qi = pair / a.k;
    if (a.done[qi] || a.e[pair] < 0.f) return;
    const int tid = threadIdx.x;
    const int rows = a.rows, cols = a.cols, ld = STAGED ? cols + 1 : cols;
    const float* Kg = a.K + pair * (int64_t)rows * cols;
    const float* u = a.u + pair * rows;
    const float* v = a.v + pair * cols;
    float* rv = a.rv + pair * rows;
    float* cv = a.cv + pair * cols;
    for (int m = tid; m < cols; m += GI_THREADS) cs[m] = cv[m];
    if (STAGED)
        for (int i = tid; i < rows * cols; i += GI_THREADS) Ks[(i / cols) * ld + (i % cols)] = Kg[i];
    __syncthreads();
    const float* K = STAGED ? Ks : Kg;
    // torch.bmm switches to a plain (unfused multiply, then add) loop below 400 multiply-adds per
    // matrix (aten/src/ATen/native/LinearAlgebra.cpp, bmm_out_or_baddbmm_); follow it so that the
    // result stays bit-identical for tiny problems too.
    const bool fused = (int64_t)rows * cols >= 400;
    float e = 0.f;
    for (int s = tid; s < rows; s += GI_THREADS) {
        float y = 0.f;
        if (fused)
            for (int m = 0; m < cols; m++) y = fmaf(K[(int64_t)s * ld + m], cs[m], y);
        else
            for (int m = 0; m < cols; m++) y = __fadd_rn(y, __fmul_rn(K[(int64_t)s * ld + m], cs[m]));
        const float rn = u[s] / y;
        e += fabsf(rn - rv[s]);
        rv[s] = rn;
        rs[s] = rn;
    }
    __syncthreads();
    for (int m = tid; m < cols; m += GI_THREADS) {
        float x = 0.f;
        if (fused)
            for (int s = 0; s < rows; s++) x = fmaf(K[(int64_t)s * ld + m], rs[s], x);
        else
            for (int s = 0; s < rows; s++) x = __fadd_rn(x, __fmul_rn(K[(int64_t)s * ld + m], rs[s]));
        cv[m] = v[m] / x;
    }
    e = block_reduce_sum(e, red);
    if (tid == 0) a.e[pair] = e;
}

__global__ void __launch_bounds__(256) generic_decide_kernel(IterArgs a, int it, float* dbg_err, int max_iter) {
    __shared__ float red[32];
    const int64_t qi = blockIdx.x;
    if (a.done[qi]) return;
    float s = 0.f;
    // padded shortlist entries carry e = -1 and add nothing; NaN / inf propagate into the mean, so that `mean < thresh`
    // is false and the query runs all its iterations like the reference's `nan < thresh` (diml.py:50-52)
    for (int i = threadIdx.x; i < a.k; i += 256) {
        const float e = a.e[qi * a.k + i];
        s += (e < 0.f) ? 0.f : e;
    }
    s = block_reduce_sum(s, red);
    if (threadIdx.x == 0) {
        a.niter[qi] = it + 1;
        const float mean = s / ((float)a.k * (float)a.rows);
        if (dbg_err) dbg_err[qi * max_iter + it] = mean;
        if (mean < a.thresh) a.done[qi] = 1;
    }
}

struct FinishArgs {
    const float* K;
    const float* sim;  // [np, r, r] or nullptr
    const float* rv;
    const float* cv;
    int rows, cols, r;
    float* out_T;      // [np, rows, cols] or nullptr
    float* out_simr;   // [np, r, r] or nullptr
    float* out_score;  // [np] or nullptr
};

__global__ void __launch_bounds__(256) generic_finish_kernel(FinishArgs a) {
    __shared__ float red[32];
    const int64_t pair = blockIdx.x;
    const int rows = a.rows, cols = a.cols, R = a.r;
    const float* K = a.K + pair * (int64_t)rows * cols;
    const float* rv = a.rv + pair * rows;
    const float* cv = a.cv + pair * cols;
    float sc = 0.f;
    // full-OT score-only calls whose rows are whole float4s (R = 196, 48, ...): a warp per row, 16-byte loads, no index arithmetic
    if (!a.out_T && !a.out_simr && a.sim && rows == R && cols == R && (R & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(a.K) | reinterpret_cast<uintptr_t>(a.sim) | reinterpret_cast<uintptr_t>(a.cv)) & 15) == 0) {
        const float* sim = a.sim + pair * (int64_t)R * R;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int s = warp; s < R; s += 8) {
            const float rs = rv[s];
            const float4* K4 = reinterpret_cast<const float4*>(K + (int64_t)s * R);
            const float4* S4 = reinterpret_cast<const float4*>(sim + (int64_t)s * R);
            const float4* C4 = reinterpret_cast<const float4*>(cv);
            for (int m4 = lane; m4 < R / 4; m4 += 32) {
                const float4 k = __ldcs(K4 + m4), x = __ldcs(S4 + m4), c = C4[m4];
                sc += ((rs * c.x) * k.x) * x.x;
                sc += ((rs * c.y) * k.y) * x.y;
                sc += ((rs * c.z) * k.z) * x.z;
                sc += ((rs * c.w) * k.w) * x.w;
            }
        }
        sc = block_reduce_sum(sc, red);
        if (threadIdx.x == 0 && a.out_score) a.out_score[pair] = sc;
        return;
    }
    for (int i = threadIdx.x; i < rows * cols; i += 256) {
        const int s = i / cols, m = i % cols;
        const float T = (rv[s] * cv[m]) * K[i];
        if (a.out_T) a.out_T[pair * (int64_t)rows * cols + i] = T;
        if (a.sim && s < R && m < R) {
            const float sr = T * a.sim[pair * (int64_t)R * R + (int64_t)s * R + m];
            sc += sr;
            if (a.out_simr) a.out_simr[pair * (int64_t)R * R + (int64_t)s * R + m] = sr;
        }
    }
    if (a.out_score) {
        sc = block_reduce_sum(sc, red);
        if (threadIdx.x == 0) a.out_score[pair] = sc;
    }
}

__global__ void generic_init_kernel(float* rv, float* cv, float* e, int64_t nr, int64_t nc, int64_t np, int32_t* done,
                                    int32_t* niter, int64_t nq) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) e[i] = 0.f;
    if (i < nr) rv[i] = 1.f;
    if (i < nc) cv[i] = 1.f;
    if (i < nq) {
        done[i] = 0;
        niter[i] = 0;
    }
}

__global__ void copy_niter_kernel(const int32_t* src, int32_t* dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// ---- workspace carving ----------------------------------------------------------------
constexpr int SKC_T = 8;   // iterations per launch of the chunked solver

struct GenWs {
    float *sim, *K, *u, *v, *rv, *cv, *e;
    float *ehist, *rhist, *chist;   // chunked solver: sum |dr| and the (r, c) state of every iteration of the current chunk
    int32_t *done, *niter, *tstar;
    size_t bytes;
};

// mode 0: the direct Sinkhorn call (K, u, v are the caller's); 1: the rerank path; 2: generic_fused_rerank only (u and v for the
// cross-correlation marginals, but no sim, K or state history: rhist holds the per-iteration scores)
static GenWs carve(void* base, int64_t nq, int64_t np, int r, int rows, int cols, int mode) {
    const bool with_sim = mode == 1;
    GenWs w{};
    size_t off = 0;
    auto take = [&](size_t n) {
        size_t o = off;
        off = align_up(off + n, 256);
        return base ? reinterpret_cast<unsigned char*>(base) + o : nullptr;
    };
    w.sim = reinterpret_cast<float*>(take(with_sim ? (size_t)np * r * r * 4 : 0));
    w.K = reinterpret_cast<float*>(take(with_sim ? (size_t)np * rows * cols * 4 : 0));
    w.u = reinterpret_cast<float*>(take(mode ? (size_t)np * rows * 4 : 0));
    w.v = reinterpret_cast<float*>(take(mode ? (size_t)np * cols * 4 : 0));
    w.rv = reinterpret_cast<float*>(take((size_t)np * rows * 4));
    w.cv = reinterpret_cast<float*>(take((size_t)np * cols * 4));
    w.e = reinterpret_cast<float*>(take((size_t)np * 4));
    w.done = reinterpret_cast<int32_t*>(take((size_t)nq * 4));
    w.niter = reinterpret_cast<int32_t*>(take((size_t)nq * 4));
    w.tstar = reinterpret_cast<int32_t*>(take((size_t)nq * 4));
    w.ehist = reinterpret_cast<float*>(take((size_t)np * SKC_T * 4));
    w.rhist = reinterpret_cast<float*>(take((size_t)np * SKC_T * (mode == 2 ? 1 : rows) * 4));
    w.chist = reinterpret_cast<float*>(take(mode == 2 ? 0 : (size_t)np * SKC_T * cols * 4));
    w.bytes = off + 256;
    return w;
}

size_t generic_rerank_workspace_bytes(int64_t nq, int k, int r, const vr_ot_params* p) {
    const int re = (p->ot_part > 0.999f) ? r : r + 1;
    return carve(nullptr, nq, nq * k, r, re, re, 1).bytes;
}

// What generic_rerank needs when it takes generic_fused.cu (a.packed set, scores only, generic_fused_supported): 1.6 KB per pair
// instead of 2 R^2 floats.  (3.2 KB with the marginal buffers.)
size_t generic_fused_workspace_bytes(int64_t nq, int k, int r) { return carve(nullptr, nq, nq * k, r, r, r, 2).bytes; }

size_t generic_sinkhorn_workspace_bytes(int64_t b, int m, int n) { return carve(nullptr, 1, b, 0, m, n, 0).bytes; }

// ---- the whole loop in ONE launch: a CTA per pair keeps K in shared memory for all iterations ----
// The multi-launch scheme above re-stages K (154 KB for a 14 x 14 grid) from global memory in every iteration and always issues
// 2 x max_iter launches.  Here the k CTAs of a query stay resident and exchange their sum |dr| through tagged 8-byte words in
// L2 (the protocol of pair_fused.cu's global transport: tag = (query + 1, iteration + 1), value = the float's bits; single-copy
// atomic, no fence, the reader re-reads until every tag matches).  The sum of iteration t is published after its row pass and
// looked at after the row pass of iteration t + 1, so its L2 round trip hides behind a column and a row pass; r and c are
// double-buffered, so the state of iteration t is intact when its test fires.  Every CTA adds the k values in index order:
// all reach the same decision.  Needs the k CTAs of a query co-resident (k <= SMs) and K in shared memory (rows <= ~236).
__host__ __device__ inline int skp_ld(int cols) {
    int q = (cols + 3) / 4;
    if ((q & 1) == 0) q++;
    return 4 * q;
}
constexpr int SKP_MAXK = 128;          // words per exchange step
constexpr int SKP_SLOTS = 8;           // steps in flight (iteration & 7)
constexpr int SKP_RING = 512;          // query slots of the exchange buffer (4 MB / (8 x 128 x 8 B))

struct SkpArgs {
    IterArgs it;
    unsigned long long* part;
    int max_iter;
    float* dbg_err;
};

__global__ void __launch_bounds__(GI_THREADS, 1) generic_sk_persistent_kernel(SkpArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const IterArgs& a = p.it;
    // row stride: a multiple of 4 floats with ld / 4 odd -- the row pass reads K with LDS.128 (lane = row: the 8 lanes of a
    // quarter-warp fall on 8 distinct 4-bank groups), the column pass with LDS.32 (lane = column: consecutive words)
    const int rows = a.rows, cols = a.cols, ld = skp_ld(cols);
    const int rp = (rows + 3) & ~3, cp = (cols + 3) & ~3;
    float* Ks = reinterpret_cast<float*>(smem_raw);   // [rows][ld]
    float* cs = Ks + (size_t)rows * ld;               // [2][cp]
    float* rs = cs + 2 * cp;                          // [2][rp]
    float* us = rs + 2 * rp;                          // [rp]
    float* vs = us + rp;                              // [cp]
    float* red = vs + cp;                             // [32]
    float* vals = red + 32;                           // [SKP_MAXK]
    __shared__ int s_stop, s_first;
    __shared__ __align__(8) uint64_t kbar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t pair = blockIdx.x;
    const int64_t qi = pair / a.k;
    const int pi = (int)(pair % a.k);
    if (a.e[pair] < 0.f) return;                      // padded shortlist entry: not part of the problem
    const float* Kg = a.K + pair * (int64_t)rows * cols;
    // K -> shared memory once: one bulk copy when the rows need no padding (cols = ld, e.g. 196) and the matrix is 16-byte aligned,
    // else row by row through registers with a few loads in flight per thread
    const bool bulk = ld == cols && ((reinterpret_cast<uintptr_t>(Kg) & 15) == 0) && ((size_t)rows * cols * 4 < (1u << 20));
    if (bulk) {
        if (tid == 0) {
            mbar_init(&kbar, 1);
            fence_mbar_init();
            mbar_expect_tx(&kbar, (uint32_t)(rows * cols * 4));
            bulk_g2s(Ks, Kg, (uint32_t)(rows * cols * 4), &kbar);
        }
    } else {
        for (int s = warp; s < rows; s += GI_THREADS / 32) {
            float* dst = Ks + (size_t)s * ld;
            const float* src = Kg + (int64_t)s * cols;
            for (int m = lane; m < ld; m += 32) dst[m] = m < cols ? __ldg(src + m) : 0.f;
        }
    }
    for (int s = tid; s < rows; s += GI_THREADS) {
        us[s] = a.u[pair * rows + s];
        rs[rp + s] = a.rv[pair * rows + s];           // "r of iteration -1" (ones, diml.py:43)
    }
    for (int m = tid; m < cp; m += GI_THREADS) {
        vs[m] = m < cols ? a.v[pair * cols + m] : 0.f;
        cs[cp + m] = m < cols ? a.cv[pair * cols + m] : 0.f;   // "c of iteration -1" (ones, diml.py:44); padding stays 0
        cs[m] = 0.f;
    }
    // the members of this query that take part (padded entries carry e = -1), and the first of them (it reports)
    if (warp == 0) {
        int first = a.k;
        for (int i = lane; i < a.k; i += 32) {
            const bool act = a.e[qi * a.k + i] >= 0.f;
            vals[i] = act ? 1.f : 0.f;
            if (act) first = min(first, i);
        }
        first = __reduce_min_sync(0xffffffffu, first);
        if (lane == 0) s_first = first;
    }
    __syncthreads();
    if (bulk) mbar_wait(&kbar, 0);
    const bool reporter = pi == s_first;
    uint32_t actmask[SKP_MAXK / 32];
#pragma unroll
    for (int w = 0; w < SKP_MAXK / 32; w++) {
        const int i = lane + 32 * w;
        actmask[w] = (i < a.k && vals[i] != 0.f) ? 1u : 0u;   // this lane's members (lane + 32 w)
    }
    __syncthreads();
    unsigned long long* part = p.part + (size_t)(qi & (SKP_RING - 1)) * (SKP_SLOTS * SKP_MAXK);
    const uint32_t qtag = (uint32_t)(qi + 1) << 8;
    const bool fused = (int64_t)rows * cols >= 400;   // torch.bmm's small-matrix path multiplies and adds unfused (see above)
    const float denom = (float)a.k * (float)rows;
    // warp 0: the batch mean of |dr| of iteration `step` (waits until every member has published it)
    auto gather_mean = [&](int step) -> float {
        const uint32_t want = qtag | (uint32_t)(step + 1);
        const unsigned long long* src = part + (step & (SKP_SLOTS - 1)) * SKP_MAXK;
        long long t0 = 0;
#pragma unroll
        for (int w = 0; w < SKP_MAXK / 32; w++) {
            float v = 0.f;
            if (actmask[w]) {
                for (;;) {
                    unsigned long long x;
                    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(x) : "l"(src + lane + 32 * w) : "memory");
                    if ((uint32_t)(x >> 32) == want) {
                        v = __uint_as_float((uint32_t)x);
                        break;
                    }
                    if (t0 == 0) t0 = clock64();
                    if (clock64() - t0 > 8000000000ll) __trap();   // ~4 s: the group is not co-resident; fail loudly
                    __nanosleep(64);
                }
            }
            vals[lane + 32 * w] = v;
        }
        __syncwarp();
        float s = 0.f;
        for (int i = 0; i < a.k; i++) s += vals[i];   // index order: identical in every CTA of the query
        __syncwarp();
        return s / denom;
    };
    int niter = p.max_iter, fin = (p.max_iter - 1) & 1;
    if (p.max_iter == 0) fin = 1;
    for (int it = 0; it < p.max_iter; it++) {
        const int cur = it & 1, prv = cur ^ 1;
        // row pass: r = u / (K c)
        float e = 0.f;
        for (int s = tid; s < rows; s += GI_THREADS) {
            const float4* Kr = reinterpret_cast<const float4*>(Ks + (size_t)s * ld);
            const float4* c = reinterpret_cast<const float4*>(cs + prv * cp);
            float y = 0.f;
            if (fused) {   // (padding columns: K = 0 and c = 0 add exact zeros at the end of the chain)
#pragma unroll 4
                for (int m4 = 0; m4 < cp / 4; m4++) {
                    const float4 kv = Kr[m4], cv = c[m4];
                    y = fmaf(kv.x, cv.x, y);
                    y = fmaf(kv.y, cv.y, y);
                    y = fmaf(kv.z, cv.z, y);
                    y = fmaf(kv.w, cv.w, y);
                }
            } else {
                const float* K1 = Ks + (size_t)s * ld;
                const float* c1 = cs + prv * cp;
                for (int m = 0; m < cols; m++) y = __fadd_rn(y, __fmul_rn(K1[m], c1[m]));
            }
            const float rn = us[s] / y;
            e += fabsf(rn - rs[prv * rp + s]);
            rs[cur * rp + s] = rn;
        }
        e = block_reduce_sum(e, red);                 // (includes the barrier that publishes r)
        if (tid == 0) {
            const unsigned long long w = ((unsigned long long)(qtag | (uint32_t)(it + 1)) << 32) | (unsigned long long)__float_as_uint(e);
            asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(part + (it & (SKP_SLOTS - 1)) * SKP_MAXK + pi), "l"(w) : "memory");
        }
        if (it >= 1) {   // the test of iteration it - 1
            if (warp == 0) {
                const float mean = gather_mean(it - 1);
                if (lane == 0) {
                    if (reporter && p.dbg_err) p.dbg_err[qi * p.max_iter + it - 1] = mean;
                    s_stop = mean < a.thresh ? 1 : 0;             // NaN: no stop, like `nan < thresh`
                }
            }
            __syncthreads();
            if (s_stop) {   // iteration it - 1 was the last one: its r and c are still in place
                niter = it;
                fin = prv;
                break;
            }
        }
        // column pass: c = v / (K^T r)
        for (int m = tid; m < cols; m += GI_THREADS) {
            const float* r = rs + cur * rp;
            float x = 0.f;
            if (fused) {   // r by float4: a broadcast LDS.32 per row would cost the load pipe as much as K itself
                const float4* r4 = reinterpret_cast<const float4*>(r);
                const float* Kc = Ks + m;
                int s = 0;
#pragma unroll 2
                for (; s + 4 <= rows; s += 4) {
                    const float4 rr = r4[s >> 2];
                    x = fmaf(Kc[(size_t)(s + 0) * ld], rr.x, x);
                    x = fmaf(Kc[(size_t)(s + 1) * ld], rr.y, x);
                    x = fmaf(Kc[(size_t)(s + 2) * ld], rr.z, x);
                    x = fmaf(Kc[(size_t)(s + 3) * ld], rr.w, x);
                }
                for (; s < rows; s++) x = fmaf(Kc[(size_t)s * ld], r[s], x);
            } else {
                for (int s = 0; s < rows; s++) x = __fadd_rn(x, __fmul_rn(Ks[(size_t)s * ld + m], r[s]));
            }
            cs[cur * cp + m] = vs[m] / x;
        }
        __syncthreads();
    }
    if (reporter && p.dbg_err && niter == p.max_iter && p.max_iter >= 1 && warp == 0) {
        // the last iteration's value was not needed for a decision: report it for the trace all the same
        const float mean = gather_mean(p.max_iter - 1);
        if (lane == 0) p.dbg_err[qi * p.max_iter + p.max_iter - 1] = mean;
    }
    for (int s = tid; s < rows; s += GI_THREADS) a.rv[pair * rows + s] = rs[fin * rp + s];
    for (int m = tid; m < cols; m += GI_THREADS) a.cv[pair * cols + m] = cs[fin * cp + m];
    if (reporter && tid == 0) a.niter[qi] = niter;
}

// ---- the loop in chunks of SKC_T iterations, no exchange at all ----
// The pairs of a query only meet in the stop test (the batch mean of |dr|); between tests they evolve independently.  A CTA
// therefore runs SKC_T iterations of its pair on its own -- K staged once per chunk -- and records sum |dr| AND the (r, c) state
// of every iteration; `decide` then finds, per query, the first iteration of the chunk whose batch mean is below the threshold,
// and `select` copies that iteration's state back (or the last one, to continue from).  No CTA waits for another: any number of
// pairs per query, every SM busy (the persistent kernel above keeps 100 of 148 SMs busy at K = 100 and spends half its warp time
// behind the exchange poll), at the price of up to SKC_T - 1 iterations run in vain per query.
struct ChunkArgs {
    IterArgs it;
    float *ehist, *rhist, *chist;
    int32_t* tstar;
    int it0, max_iter;
    float* dbg_err;
};

__global__ void __launch_bounds__(GI_THREADS, 1) generic_sk_chunk_kernel(ChunkArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const IterArgs& a = p.it;
    const int rows = a.rows, cols = a.cols, ld = skp_ld(cols);
    const int rp = (rows + 3) & ~3, cp = (cols + 3) & ~3;
    float* Ks = reinterpret_cast<float*>(smem_raw);   // [rows][ld]
    float* cs = Ks + (size_t)rows * ld;               // [2][cp]
    float* rs = cs + 2 * cp;                          // [2][rp]
    float* us = rs + 2 * rp;                          // [rp]
    float* vs = us + rp;                              // [cp]
    float* red = vs + cp;                             // [32]
    __shared__ __align__(8) uint64_t kbar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t pair = blockIdx.x;
    const int64_t qi = pair / a.k;
    if (a.done[qi] || a.e[pair] < 0.f) return;
    const float* Kg = a.K + pair * (int64_t)rows * cols;
    const bool bulk = ld == cols && ((reinterpret_cast<uintptr_t>(Kg) & 15) == 0) && ((size_t)rows * cols * 4 < (1u << 20));
    if (bulk) {
        if (tid == 0) {
            mbar_init(&kbar, 1);
            fence_mbar_init();
            mbar_expect_tx(&kbar, (uint32_t)(rows * cols * 4));
            bulk_g2s(Ks, Kg, (uint32_t)(rows * cols * 4), &kbar);
        }
    } else {
        for (int s = warp; s < rows; s += GI_THREADS / 32) {
            float* dst = Ks + (size_t)s * ld;
            const float* src = Kg + (int64_t)s * cols;
            for (int m = lane; m < ld; m += 32) dst[m] = m < cols ? __ldg(src + m) : 0.f;
        }
    }
    for (int s = tid; s < rows; s += GI_THREADS) {
        us[s] = a.u[pair * rows + s];
        rs[rp + s] = a.rv[pair * rows + s];           // r of the iteration before this chunk
    }
    for (int m = tid; m < cp; m += GI_THREADS) {
        vs[m] = m < cols ? a.v[pair * cols + m] : 0.f;
        cs[cp + m] = m < cols ? a.cv[pair * cols + m] : 0.f;
        cs[m] = 0.f;
    }
    __syncthreads();
    if (bulk) mbar_wait(&kbar, 0);
    const bool fused = (int64_t)rows * cols >= 400;   // torch.bmm's small-matrix path multiplies and adds unfused (see above)
    const int nt = min(SKC_T, p.max_iter - p.it0);
    for (int t = 0; t < nt; t++) {
        const int cur = t & 1, prv = cur ^ 1;
        float e = 0.f;
        for (int s = tid; s < rows; s += GI_THREADS) {
            const float4* Kr = reinterpret_cast<const float4*>(Ks + (size_t)s * ld);
            const float4* c = reinterpret_cast<const float4*>(cs + prv * cp);
            float y = 0.f;
            if (fused) {
#pragma unroll 4
                for (int m4 = 0; m4 < cp / 4; m4++) {
                    const float4 kv = Kr[m4], cv = c[m4];
                    y = fmaf(kv.x, cv.x, y);
                    y = fmaf(kv.y, cv.y, y);
                    y = fmaf(kv.z, cv.z, y);
                    y = fmaf(kv.w, cv.w, y);
                }
            } else {
                const float* K1 = Ks + (size_t)s * ld;
                const float* c1 = cs + prv * cp;
                for (int m = 0; m < cols; m++) y = __fadd_rn(y, __fmul_rn(K1[m], c1[m]));
            }
            const float rn = us[s] / y;
            e += fabsf(rn - rs[prv * rp + s]);
            rs[cur * rp + s] = rn;
            p.rhist[(pair * SKC_T + t) * rows + s] = rn;
        }
        e = block_reduce_sum(e, red);
        if (tid == 0) p.ehist[pair * SKC_T + t] = e;
        for (int m = tid; m < cols; m += GI_THREADS) {
            const float* r = rs + cur * rp;
            float x = 0.f;
            if (fused) {   // r by float4: a broadcast LDS.32 per row would cost the load pipe as much as K itself
                const float4* r4 = reinterpret_cast<const float4*>(r);
                const float* Kc = Ks + m;
                int s = 0;
#pragma unroll 2
                for (; s + 4 <= rows; s += 4) {
                    const float4 rr = r4[s >> 2];
                    x = fmaf(Kc[(size_t)(s + 0) * ld], rr.x, x);
                    x = fmaf(Kc[(size_t)(s + 1) * ld], rr.y, x);
                    x = fmaf(Kc[(size_t)(s + 2) * ld], rr.z, x);
                    x = fmaf(Kc[(size_t)(s + 3) * ld], rr.w, x);
                }
                for (; s < rows; s++) x = fmaf(Kc[(size_t)s * ld], r[s], x);
            } else {
                for (int s = 0; s < rows; s++) x = __fadd_rn(x, __fmul_rn(Ks[(size_t)s * ld + m], r[s]));
            }
            const float cn = vs[m] / x;
            cs[cur * cp + m] = cn;
            p.chist[(pair * SKC_T + t) * cols + m] = cn;
        }
        __syncthreads();
    }
}

// Per query: the first iteration of the chunk whose batch mean of |dr| is below the threshold (diml.py:50-52).
__global__ void __launch_bounds__(256) generic_sk_decide_chunk_kernel(ChunkArgs p) {
    __shared__ float red[32];
    const IterArgs& a = p.it;
    const int64_t qi = blockIdx.x;
    if (a.done[qi]) {
        if (threadIdx.x == 0) p.tstar[qi] = -1;   // finished in an earlier chunk: its state is final
        return;
    }
    const int nt = min(SKC_T, p.max_iter - p.it0);
    int tstar = nt - 1, stop = 0;
    for (int t = 0; t < nt && !stop; t++) {
        float s = 0.f;
        for (int i = threadIdx.x; i < a.k; i += 256)
            if (a.e[qi * a.k + i] >= 0.f) s += p.ehist[(qi * a.k + i) * SKC_T + t];   // NaN / inf propagate: no stop
        s = block_reduce_sum(s, red);
        const float mean = s / ((float)a.k * (float)a.rows);
        if (threadIdx.x == 0 && p.dbg_err) p.dbg_err[qi * p.max_iter + p.it0 + t] = mean;
        if (mean < a.thresh) {
            tstar = t;
            stop = 1;
        }
    }
    if (threadIdx.x == 0) {
        p.tstar[qi] = tstar;
        a.niter[qi] = p.it0 + tstar + 1;
        if (stop) a.done[qi] = 1;
    }
}

// Per pair: the state of iteration tstar of the chunk becomes the pair's (r, c) -- final, or the state the next chunk continues from.
__global__ void __launch_bounds__(128) generic_sk_select_kernel(ChunkArgs p) {
    const IterArgs& a = p.it;
    const int64_t pair = blockIdx.x;
    const int64_t qi = pair / a.k;
    const int t = p.tstar[qi];
    if (t < 0 || a.e[pair] < 0.f) return;
    for (int s = threadIdx.x; s < a.rows; s += 128) a.rv[pair * a.rows + s] = p.rhist[(pair * SKC_T + t) * a.rows + s];
    for (int m = threadIdx.x; m < a.cols; m += 128) a.cv[pair * a.cols + m] = p.chist[(pair * SKC_T + t) * a.cols + m];
}

static size_t skp_smem(int rows, int cols) {
    const int ld = skp_ld(cols), rp = (rows + 3) & ~3, cp = (cols + 3) & ~3;
    return ((size_t)rows * ld + 3 * (size_t)cp + 3 * (size_t)rp + 32 + SKP_MAXK) * 4;
}

// The persistent kernel applies when K fits in shared memory, a query has at most 128 members that can all be resident, and
// the iteration / query counts fit the tag.  VR_GENERIC_SK=launches keeps the multi-launch scheme (A/B tests).
static bool skp_usable(const IterArgs& it, int64_t nq, int max_iter, int* resident_out) {
    const char* e = getenv("VR_GENERIC_SK");
    if (!e || e[0] != 'p') return false;   // opt-in: VR_GENERIC_SK=persistent
    const size_t smem = skp_smem(it.rows, it.cols);
    if (smem > 225 * 1024 || it.k > SKP_MAXK || max_iter > 254 || max_iter < 1 || nq >= (1ll << 23)) return false;
    if (cudaFuncSetAttribute(generic_sk_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, generic_sk_persistent_kernel, GI_THREADS, smem) != cudaSuccess)
        return false;
    *resident_out = per_sm * sms;
    return per_sm * sms >= it.k;   // the members of a query wait for one another: all must fit on the device
}

static int run_iterations(const IterArgs& it, const GenWs& w, int64_t nq, int64_t np, int max_iter, float* dbg_err, cudaStream_t st) {
    const size_t base = (size_t)(it.rows + it.cols + 32) * 4;
    const size_t staged = base + (size_t)it.rows * (it.cols + 1) * 4;
    const bool use_staged = staged <= 200 * 1024;
    const size_t smem = use_staged ? staged : base;
    VR_REQUIRE(base <= 48 * 1024, "sinkhorn: %d x %d too large", it.rows, it.cols);
    {   // default: chunks of SKC_T iterations without any exchange (VR_GENERIC_SK=persistent | launches select the others)
        const char* e = getenv("VR_GENERIC_SK");
        const size_t smem_c = skp_smem(it.rows, it.cols);
        if ((!e || e[0] == 'c') && smem_c <= 225 * 1024 && max_iter >= 1) {
            VR_CHECK_CUDA(cudaFuncSetAttribute(generic_sk_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
            ChunkArgs c{it, w.ehist, w.rhist, w.chist, w.tstar, 0, max_iter, dbg_err};
            for (int it0 = 0; it0 < max_iter; it0 += SKC_T) {
                c.it0 = it0;
                generic_sk_chunk_kernel<<<(unsigned)np, GI_THREADS, smem_c, st>>>(c);
                VR_LAUNCH_CHECK();
                generic_sk_decide_chunk_kernel<<<(unsigned)nq, 256, 0, st>>>(c);
                VR_LAUNCH_CHECK();
                generic_sk_select_kernel<<<(unsigned)np, 128, 0, st>>>(c);
                VR_LAUNCH_CHECK();
            }
            return VR_OK;
        }
    }
    int resident = 0;
    if (skp_usable(it, nq, max_iter, &resident)) {
        SkpArgs p{it, nullptr, max_iter, dbg_err};
        size_t xbytes = 0;
        int rc = pair_exchange_begin(st, &p.part, &xbytes);
        if (rc) return rc;
        if (xbytes >= (size_t)SKP_RING * SKP_SLOTS * SKP_MAXK * 8) {
            generic_sk_persistent_kernel<<<(unsigned)np, GI_THREADS, skp_smem(it.rows, it.cols), st>>>(p);
            const cudaError_t le = cudaGetLastError();
            vr::g_launches++;
            rc = pair_exchange_end(st);
            if (le != cudaSuccess) {
                set_error("generic_sk_persistent_kernel: %s", cudaGetErrorString(le));
                return VR_E_CUDA;
            }
            return rc;
        }
        if ((rc = pair_exchange_end(st))) return rc;
    }
    if (use_staged && smem > 48 * 1024)
        VR_CHECK_CUDA(cudaFuncSetAttribute(generic_iter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    for (int i = 0; i < max_iter; i++) {
        if (use_staged)
            generic_iter_kernel<true><<<(unsigned)np, GI_THREADS, smem, st>>>(it);
        else
            generic_iter_kernel<false><<<(unsigned)np, GI_THREADS, smem, st>>>(it);
        VR_LAUNCH_CHECK();
        generic_decide_kernel<<<(unsigned)nq, 256, 0, st>>>(it, i, dbg_err, max_iter);
        VR_LAUNCH_CHECK();
    }
    return VR_OK;
}

// Shared memory of generic_prepare_kernel; decides whether the cross-correlation rows go through the bulk-copy ring (a.cc_stages).
static size_t prepare_smem(GenArgs& a) {
    const size_t base = (size_t)(2 * GP_KC * GP_T + 32 + 2 * a.c + 2 * a.r) * 4;
    a.cc_stages = 0;
    const char* e = getenv("VR_GENERIC_CC");
    if (a.p.mode < VR_MODE_INVERSE || (e && e[0] == 'l') || a.c % GP_KC != 0 || a.r > 4 * GP_THREADS ||
        ((reinterpret_cast<uintptr_t>(a.q_patches) | reinterpret_cast<uintptr_t>(a.c_patches)) & 15) != 0)
        return base;
    const size_t stage = (size_t)2 * GP_KC * a.r * 4;
    int ns = 4;
    while (ns > 1 && base + 16 + ns * stage + ns * 8 > 100 * 1024) ns--;   // two CTAs per SM
    if (base + 16 + ns * stage + ns * 8 > 100 * 1024) return base;
    a.cc_stages = ns;
    return base + 16 + ns * stage + ns * 8;
}

int generic_rerank(GenArgs a, void* ws, size_t ws_bytes, cudaStream_t st) {
    VR_REQUIRE(a.nq > 0 && a.k > 0, "generic_rerank: empty problem");
    VR_REQUIRE(a.c >= 1 && a.r >= 1 && a.r <= 1024 && a.c <= 4096, "generic_rerank: unsupported shape C=%d R=%d", a.c,
               a.r);
    const int64_t np = a.nq * a.k;
    VR_REQUIRE(np < 0x7fffffffll, "generic_rerank: too many pairs in one call (%lld)", (long long)np);
    const int re = (a.p.ot_part > 0.999f) ? a.r : a.r + 1;
    a.part_bin = partial_ot_bin(a.p.ot_part);
    const bool fused = a.packed && !a.out_T && !a.out_simr && !a.out_u && !a.out_cc && generic_fused_supported(a.c, a.r, &a.p) &&
                       (a.p.mode != VR_MODE_ROLLOUT || a.c_rollout);
    GenWs w = carve(ws, a.nq, np, a.r, re, re, fused ? 2 : 1);
    if (w.bytes > ws_bytes) {
        // not enough for all queries at once: as many per round as fit (the queries of a call are independent problems)
        int64_t fit = a.nq;
        while (fit > 1 && carve(nullptr, fit, fit * a.k, a.r, re, re, fused ? 2 : 1).bytes > ws_bytes) fit >>= 1;
        if (a.nq == 1 || carve(nullptr, fit, fit * a.k, a.r, re, re, fused ? 2 : 1).bytes > ws_bytes) {
            set_error("generic_rerank: workspace %zu < %zu (one query needs %zu)", ws_bytes, w.bytes,
                      carve(nullptr, 1, a.k, a.r, re, re, fused ? 2 : 1).bytes);
            return VR_E_WORKSPACE;
        }
        for (int64_t lo = 0; lo < a.nq; lo += fit) {
            GenArgs b = a;
            b.nq = std::min(fit, a.nq - lo);
            b.q_start = a.q_start + lo * a.q_stride;
            if (a.cand_idx) b.cand_idx = a.cand_idx + lo * a.cand_stride;
            const int64_t p0 = lo * a.k;
            if (a.out_score) b.out_score = a.out_score + p0;
            if (a.out_niter) b.out_niter = a.out_niter + lo;
            if (a.out_u) b.out_u = a.out_u + p0 * a.r;
            if (a.out_v) b.out_v = a.out_v + p0 * a.r;
            if (a.out_T) b.out_T = a.out_T + p0 * re * re;
            if (a.out_simr) b.out_simr = a.out_simr + p0 * a.r * a.r;
            if (a.out_cc) b.out_cc = a.out_cc + p0 * a.r;
            if (a.dbg_err) b.dbg_err = a.dbg_err + lo * a.p.max_iter;
            int rc = generic_rerank(b, ws, ws_bytes, st);
            if (rc) return rc;
        }
        return VR_OK;
    }
    a.sim = w.sim; a.K = w.K; a.u = w.u; a.v = w.v; a.rv = w.rv; a.cv = w.cv; a.e = w.e;
    a.done = w.done; a.niter = w.niter;
    a.sim_done = 0;
    if (fused) {
        // registered bank with its operand copy, scores only: S3 and S4 in one kernel, nothing but the scores leaves the SMs
        if (a.p.mode >= VR_MODE_INVERSE && !(a.packed_centers && a.p.use_cls_token)) {
            // cross-correlation marginals from the fp32 rows (centres = patch means): generic_prepare_kernel, marginals only
            a.sim = nullptr;
            a.K = nullptr;
            a.sim_done = 1;
            const size_t smem_p = prepare_smem(a);
            if (smem_p > 48 * 1024)
                VR_CHECK_CUDA(cudaFuncSetAttribute(generic_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
            generic_prepare_kernel<<<(unsigned)np, GP_THREADS, smem_p, st>>>(a);
            VR_LAUNCH_CHECK();
        }
        int rc = generic_fused_rerank(a, w.done, w.tstar, reinterpret_cast<int32_t*>(w.e), w.ehist, w.rhist, w.niter, st);
        if (rc) return rc;
        if (a.out_niter) {
            copy_niter_kernel<<<(unsigned)((a.nq + 255) / 256), 256, 0, st>>>(w.niter, a.out_niter, a.nq);
            VR_LAUNCH_CHECK();
        }
        return VR_OK;
    }
    if (generic_sim_mma_supported(a.c, a.r)) {   // S3 on the tensor cores (C % 16 == 0, R <= 256)
        int rc = generic_sim_mma(a, re, st);
        if (rc) return rc;
        a.sim_done = 1;
    }
    const size_t smem = prepare_smem(a);
    if (smem > 48 * 1024)
        VR_CHECK_CUDA(cudaFuncSetAttribute(generic_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    generic_prepare_kernel<<<(unsigned)np, GP_THREADS, smem, st>>>(a);
    VR_LAUNCH_CHECK();
    IterArgs it{w.K, w.u, w.v, w.rv, w.cv, w.e, w.done, w.niter, re, re, a.k, a.p.thresh};
    int rc = run_iterations(it, w, a.nq, np, a.p.max_iter, a.dbg_err, st);
    if (rc) return rc;
    FinishArgs f{w.K, w.sim, w.rv, w.cv, re, re, a.r, a.out_T, a.out_simr, a.out_score};
    generic_finish_kernel<<<(unsigned)np, 256, 0, st>>>(f);
    VR_LAUNCH_CHECK();
    if (a.out_niter) {
        copy_niter_kernel<<<(unsigned)((a.nq + 255) / 256), 256, 0, st>>>(w.niter, a.out_niter, a.nq);
        VR_LAUNCH_CHECK();
    }
    return VR_OK;
}

int generic_sinkhorn(const float* K, const float* u, const float* v, int64_t b, int m, int n, int max_iter,
                     float thresh, float* T, int32_t* niter, void* ws, size_t ws_bytes, cudaStream_t st) {
    VR_REQUIRE(b > 0 && m > 0 && n > 0 && b < 0x7fffffffll, "sinkhorn: bad shape [%lld, %d, %d]", (long long)b, m, n);
    GenWs w = carve(ws, 1, b, 0, m, n, 0);
    if (w.bytes > ws_bytes) {
        set_error("sinkhorn: workspace %zu < %zu", ws_bytes, w.bytes);
        return VR_E_WORKSPACE;
    }
    const int64_t mx = b * (int64_t)(m > n ? m : n);
    generic_init_kernel<<<(unsigned)((mx + 255) / 256), 256, 0, st>>>(w.rv, w.cv, w.e, b * m, b * n, b, w.done, w.niter, 1);
    VR_LAUNCH_CHECK();
    IterArgs it{K, u, v, w.rv, w.cv, w.e, w.done, w.niter, m, n, (int)b, thresh};
    int rc = run_iterations(it, w, 1, b, max_iter, nullptr, st);
    if (rc) return rc;
    FinishArgs f{K, nullptr, w.rv, w.cv, m, n, 0, T, nullptr, nullptr};
    generic_finish_kernel<<<(unsigned)b, 256, 0, st>>>(f);
    VR_LAUNCH_CHECK();
    if (niter) {
        copy_niter_kernel<<<1, 32, 0, st>>>(w.niter, niter, 1);
        VR_LAUNCH_CHECK();
    }
    return VR_OK;
}

}  // namespace vr
