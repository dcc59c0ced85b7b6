// S5b: blend + per-query re-sort + rank metrics, one warp per query.
//
// Replaces evaluation/eval_cvt_diml.py:357-372 (rank_in_tops = argsort(sim + approx_sim[top]),
// final_tops = cat(top[rank][:t], approx_tops[t:])) and evaluation/metrics.py:26-47
// (get_metrics_rank: r1, R-precision and MAP@R over final_tops[:num_pos], where num_pos
// counts the query itself).  Recall@1/2/4/8 is an extension named by BASELINE.json.
// Only final_tops[:num_pos] is ever read by the reference, so the first-stage shortlist of
// length kp >= max(k, num_pos) carries everything needed; nothing N-long is materialised.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace vr {

constexpr int FN_WARPS = 4;
constexpr int FN_METRICS = 8;  // r1, rp, mapr, R@1, R@2, R@4, R@8, count

struct FinalizeArgs {
    int64_t q_start, q_stride, nq;
    int k, kp, n_trunc;
    const int32_t* approx_idx;   // [nq, kp]
    const float* approx_score;   // [nq, kp]
    const float* ot_score;       // [nq, k] (nullptr when k == 0)
    const int64_t* labels;       // [N]
    const int32_t* num_pos;      // [N]
    int32_t truncs[16];
    int32_t* out_rank;           // [nq, k] or nullptr
    double* per_query;           // [nq, n_trunc, FN_METRICS]
};

__global__ void __launch_bounds__(FN_WARPS * 32) finalize_kernel(FinalizeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = a.k, kp = a.kp;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw) + (size_t)warp * k;
    int32_t* rer = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned long long*>(smem_raw) + (size_t)FN_WARPS * k) +
                   (size_t)warp * k;
    // precision@j terms of MAP@R, fp32 like the reference's tensor (metrics.py:43-45)
    float* terms = reinterpret_cast<float*>(reinterpret_cast<int32_t*>(reinterpret_cast<unsigned long long*>(smem_raw) + (size_t)FN_WARPS * k) +
                                            (size_t)FN_WARPS * k) + (size_t)warp * kp;
    const int64_t qi = (int64_t)blockIdx.x * FN_WARPS + warp;
    if (qi >= a.nq) return;
    const int64_t qid = a.q_start + qi * a.q_stride;
    const int32_t* aidx = a.approx_idx + qi * kp;
    const float* asc = a.approx_score + qi * kp;
    const int64_t ql = a.labels ? a.labels[qid] : 0;
    const int np = a.labels ? a.num_pos[qid] : 0;

    // number of valid shortlist entries (rows are padded with -1 when the gallery is small)
    int nvalid = 0;
    for (int e = lane; e < kp; e += 32) nvalid += (aidx[e] >= 0) ? 1 : 0;
    nvalid = __reduce_add_sync(0xffffffffu, nvalid);
    const int keff = min(k, nvalid);

    // ---- blend + argsort(descending) by ranking (eval_cvt_diml.py:357) ----
    for (int e = lane; e < keff; e += 32) keys[e] = pack_key(a.ot_score[qi * k + e] + asc[e], (uint32_t)e);
    __syncwarp();
    for (int e = lane; e < keff; e += 32) {
        const unsigned long long me = keys[e];
        int rank = 0;
        for (int o = 0; o < keff; o++) rank += (keys[o] > me) ? 1 : 0;
        rer[rank] = aidx[e];
    }
    __syncwarp();
    if (a.out_rank)
        for (int e = lane; e < k; e += 32) a.out_rank[qi * k + e] = e < keff ? rer[e] : -1;
    if (!a.labels) return;   // blend_rank: the reranked order only

    // ---- metrics per trunc (metrics.py:26-47) ----
    const int upto = min(np, nvalid);
    for (int ti = 0; ti < a.n_trunc; ti++) {
        const int t = a.truncs[ti];
        const int tt = min(t, keff);  // entries taken from the reranked list
        int cum = 0;                  // hits before the current 32-wide window
        unsigned first8 = 0;
        const int lim = max(upto, min(8, nvalid));  // Recall@8 looks at 8 entries even if num_pos < 8
        for (int base = 0; base < lim; base += 32) {
            const int j = base + lane;
            bool hit = false;
            if (j < lim) {
                int src = -1;
                if (j < tt) src = rer[j];
                else {
                    const int jj = j + (t - tt);  // cat(top[rank][:t], approx_tops[t:])
                    src = jj < kp ? aidx[jj] : -1;
                }
                hit = src >= 0 && a.labels[src] == ql;
            }
            if (base == 0) first8 = __ballot_sync(0xffffffffu, hit) & 0xffu;
            hit = hit && j < upto;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (j < upto) {   // (cum * eq) / k_idx: fp32 / int64 -> fp32 division (metrics.py:43-45)
                const int c = cum + __popc(m & (0xffffffffu >> (31 - lane)));
                terms[j] = hit ? (float)c / (float)(j + 1) : 0.f;
            }
            cum += __popc(m);
        }
        __syncwarp();
        // torch.mean over the num_pos terms = ATen's fp32 sum (cascade order, common.cuh) followed by one fp32 division
        const float ap = torch_sum_inner_warp(terms, upto, lane);
        __syncwarp();
        if (lane == 0) {
            double* out = a.per_query + (qi * a.n_trunc + ti) * FN_METRICS;
            out[0] = (first8 & 1u) ? 1.0 : 0.0;
            out[1] = (double)((float)cum / (float)np);
            out[2] = (double)(ap / (float)upto);
            out[3] = (first8 & 0x1u) ? 1.0 : 0.0;
            out[4] = (first8 & 0x3u) ? 1.0 : 0.0;
            out[5] = (first8 & 0xfu) ? 1.0 : 0.0;
            out[6] = (first8 & 0xffu) ? 1.0 : 0.0;
            out[7] = 1.0;
        }
    }
}

// Deterministic column sums of per_query [nq, cols] into tallies[cols] (added).
__global__ void __launch_bounds__(256) tally_kernel(const double* per_query, int64_t nq, int cols, double* tallies) {
    __shared__ double red[256];
    const int col = blockIdx.x;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < nq; i += 256) s += per_query[i * cols + col];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) tallies[col] += red[0];
}


// Direct call: evaluation/metrics.py:26-47 for ONE ranked list (tops may be N long; only
// tops[:num_pos] is read).  out = {r1, rp, mapr}.  `terms` is scratch for min(n_tops, n_labels) floats.
__global__ void __launch_bounds__(256) metrics_rank_kernel(const int64_t* tops, int64_t n_tops, int64_t qlabel,
                                                           const int64_t* labels, int64_t n_labels, float* terms,
                                                           double* out) {
    __shared__ int red[8];
    __shared__ int s_np;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cnt = 0;
    for (int64_t i = tid; i < n_labels; i += 256) cnt += (labels[i] == qlabel) ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) red[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int i = 0; i < 8; i++) t += red[i];
        s_np = t;
    }
    __syncthreads();
    if (warp != 0) return;
    const int np = s_np;
    const int64_t upto = min((int64_t)np, n_tops);
    int cum = 0;
    bool first = false;
    for (int64_t base = 0; base < upto; base += 32) {
        const int64_t j = base + lane;
        bool hit = false;
        if (j < upto) {
            const int64_t src = tops[j];
            hit = src >= 0 && src < n_labels && labels[src] == qlabel;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (base == 0) first = (m & 1u) != 0;
        if (j < upto) {
            const int c = cum + __popc(m & (0xffffffffu >> (31 - lane)));
            terms[j] = hit ? (float)c / (float)(j + 1) : 0.f;   // fp32, like (cum * eq) / k_idx (metrics.py:43-45)
        }
        cum += __popc(m);
    }
    __syncwarp();
    // torch.mean = ATen's fp32 cascade sum, then one fp32 division
    const float ap = upto > 0 ? torch_sum_inner_warp(terms, (int)upto, lane) : 0.f;
    if (lane == 0) {
        out[0] = first ? 1.0 : 0.0;
        out[1] = np > 0 ? (double)((float)cum / (float)np) : 0.0;
        out[2] = upto > 0 ? (double)(ap / (float)upto) : 0.0;
    }
}

int metrics_rank(const int64_t* tops, int64_t n_tops, int64_t qlabel, const int64_t* labels, int64_t n_labels,
                 double* out, cudaStream_t st) {
    VR_REQUIRE(tops && labels && out && n_tops > 0 && n_labels > 0, "metrics_rank: bad arguments");
    VR_REQUIRE(n_labels < 0x7fffffffll, "metrics_rank: gallery too large");
    float* terms = nullptr;   // stream-ordered scratch: num_pos is only known on the device
    VR_CHECK_CUDA(cudaMallocAsync(&terms, (size_t)std::min(n_tops, n_labels) * sizeof(float), st));
    metrics_rank_kernel<<<1, 256, 0, st>>>(tops, n_tops, qlabel, labels, n_labels, terms, out);
    VR_LAUNCH_CHECK();
    VR_CHECK_CUDA(cudaFreeAsync(terms, st));
    return VR_OK;
}

// num_pos[i] = #{j : labels[j] == labels[i]} (metrics.py:34, counts the query itself) and the largest such count,
// by brute force over label tiles in shared memory: n^2 64-bit compares (3.7e9 at SOP size: ~0.2 ms), no sort, no hash
// table, no host pass over the labels.
constexpr int NP_THREADS = 256;
constexpr int NP_TILE = 2048;
__global__ void __launch_bounds__(NP_THREADS) num_pos_kernel(const int64_t* __restrict__ labels, int64_t n, int32_t* __restrict__ num_pos,
                                                             int32_t* __restrict__ max_out) {
    __shared__ int64_t tile[NP_TILE];
    const int64_t i = (int64_t)blockIdx.x * NP_THREADS + threadIdx.x;
    const int64_t mine = i < n ? labels[i] : 0;
    int cnt = 0;
    for (int64_t base = 0; base < n; base += NP_TILE) {
        const int m = (int)min((int64_t)NP_TILE, n - base);
        __syncthreads();
        for (int t = threadIdx.x; t < m; t += NP_THREADS) tile[t] = labels[base + t];
        __syncthreads();
#pragma unroll 8
        for (int t = 0; t < m; t++) cnt += (tile[t] == mine) ? 1 : 0;
    }
    if (i < n) num_pos[i] = cnt;
    cnt = i < n ? cnt : 0;
    cnt = __reduce_max_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) atomicMax(max_out, cnt);
}

int num_pos_counts(const int64_t* labels, int64_t n, int32_t* num_pos, int32_t* max_dev, cudaStream_t st) {
    VR_REQUIRE(labels && num_pos && max_dev && n > 0 && n < 0x7fffffffll, "num_pos: bad arguments");
    VR_CHECK_CUDA(cudaMemsetAsync(max_dev, 0, sizeof(int32_t), st));
    num_pos_kernel<<<(unsigned)((n + NP_THREADS - 1) / NP_THREADS), NP_THREADS, 0, st>>>(labels, n, num_pos, max_dev);
    VR_LAUNCH_CHECK();
    return VR_OK;
}

size_t finalize_workspace_bytes(int64_t nq, int n_trunc) {
    return align_up((size_t)nq * n_trunc * FN_METRICS * sizeof(double), 256) + 256;
}

int finalize(int64_t q_start, int64_t q_stride, int64_t nq, int k, int kp, const int32_t* approx_idx,
             const float* approx_score, const float* ot_score, const int64_t* labels, const int32_t* num_pos,
             const int32_t* truncs, int n_trunc, int32_t* out_rank, double* tallies, void* ws, size_t ws_bytes,
             cudaStream_t st) {
    VR_REQUIRE(nq > 0 && kp > 0 && k >= 0 && k <= kp, "finalize: bad sizes nq=%lld k=%d kp=%d", (long long)nq, k, kp);
    VR_REQUIRE(n_trunc >= 1 && n_trunc <= 16, "finalize: 1..16 trunc values supported, got %d", n_trunc);
    VR_REQUIRE(labels && num_pos, "finalize: labels / num_pos not registered");
    VR_REQUIRE(k == 0 || ot_score, "finalize: ot_score missing");
    if (finalize_workspace_bytes(nq, n_trunc) > ws_bytes) {
        set_error("finalize: workspace %zu < %zu", ws_bytes, finalize_workspace_bytes(nq, n_trunc));
        return VR_E_WORKSPACE;
    }
    FinalizeArgs a{};
    a.q_start = q_start;
    a.q_stride = q_stride;
    a.nq = nq;
    a.k = k;
    a.kp = kp;
    a.n_trunc = n_trunc;
    a.approx_idx = approx_idx;
    a.approx_score = approx_score;
    a.ot_score = ot_score;
    a.labels = labels;
    a.num_pos = num_pos;
    for (int i = 0; i < n_trunc; i++) {
        VR_REQUIRE(truncs[i] >= 0 && truncs[i] <= k, "finalize: trunc %d outside 0..k=%d", truncs[i], k);
        a.truncs[i] = truncs[i];
    }
    a.out_rank = out_rank;
    a.per_query = reinterpret_cast<double*>(ws);
    size_t smem = (size_t)FN_WARPS * k * (8 + 4) + (size_t)FN_WARPS * kp * 4 + 16;
    VR_REQUIRE(smem <= 200 * 1024, "finalize: k=%d too large", k);
    if (smem > 48 * 1024)
        VR_CHECK_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    finalize_kernel<<<(unsigned)((nq + FN_WARPS - 1) / FN_WARPS), FN_WARPS * 32, smem, st>>>(a);
    VR_LAUNCH_CHECK();
    tally_kernel<<<n_trunc * FN_METRICS, 256, 0, st>>>(a.per_query, nq, n_trunc * FN_METRICS, tallies);
    VR_LAUNCH_CHECK();
    return VR_OK;
}

// eval_cvt_diml.py:357 / training_tools/val.py:197 alone: rank = argsort(ot_score + approx_score[:, :k], descending)
// (ties: lower position first), out_rank[q, :] = approx_idx[q, rank].  No labels, no metrics.
int blend_rank(int64_t nq, int k, int kp, const int32_t* approx_idx, const float* approx_score, const float* ot_score,
               int32_t* out_rank, cudaStream_t st) {
    VR_REQUIRE(nq > 0 && k >= 1 && k <= kp && approx_idx && approx_score && ot_score && out_rank, "blend_rank: bad arguments");
    FinalizeArgs a{};
    a.q_stride = 1;
    a.nq = nq;
    a.k = k;
    a.kp = kp;
    a.approx_idx = approx_idx;
    a.approx_score = approx_score;
    a.ot_score = ot_score;
    a.out_rank = out_rank;
    size_t smem = (size_t)FN_WARPS * k * (8 + 4) + (size_t)FN_WARPS * kp * 4 + 16;
    VR_REQUIRE(smem <= 200 * 1024, "blend_rank: k=%d too large", k);
    if (smem > 48 * 1024)
        VR_CHECK_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    finalize_kernel<<<(unsigned)((nq + FN_WARPS - 1) / FN_WARPS), FN_WARPS * 32, smem, st>>>(a);
    VR_LAUNCH_CHECK();
    return VR_OK;
}

}  // namespace vr
