// Stand-alone driver of pair_fused_kernel for timing experiments (no Python, no torch):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DPR_TIMING -o tools/pair_bench tools/pair_bench.cu
//   ./tools/pair_bench [n_images] [n_queries] [k] [sigma]
// Synthetic class-structured patch banks (same recipe as vitrerank/synth.py, different RNG), candidates
// of a query = the k images after it.  Prints the kernel time and, with -DPR_TIMING, the mean cycle
// count of each phase of cluster rank 0.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <random>
#include <vector>

#include "../vit-reranking_b200/csrc/pair_fused.cu"

namespace vr {
thread_local long long g_launches = 0;
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fprintf(stderr, "\n");
}
}  // namespace vr

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 2048, nq = argc > 2 ? atoi(argv[2]) : 1024, k = argc > 3 ? atoi(argv[3]) : 100;
    const float sigma = argc > 4 ? atof(argv[4]) : 0.6f;
    const int C = 128, R = 49, classes = std::max(2, n / 83);
    std::mt19937 rng(0);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> proto((size_t)classes * C * R), bank((size_t)n * C * R), roll((size_t)n * R);
    for (auto& x : proto) x = nd(rng);
    for (int i = 0; i < n; i++) {
        const int cls = i % classes;
        float* b = &bank[(size_t)i * C * R];
        for (int e = 0; e < C * R; e++) b[e] = proto[(size_t)cls * C * R + e] + sigma * nd(rng);
        for (int r = 0; r < R; r++) {  // L2-normalise every patch over the channels
            double s = 0;
            for (int c = 0; c < C; c++) s += (double)b[c * R + r] * b[c * R + r];
            const float inv = 1.f / (float)std::sqrt(s);
            for (int c = 0; c < C; c++) b[c * R + r] *= inv;
        }
        double z = 0;
        for (int r = 0; r < R; r++) { roll[(size_t)i * R + r] = std::exp(nd(rng)); z += roll[(size_t)i * R + r]; }
        for (int r = 0; r < R; r++) roll[(size_t)i * R + r] /= (float)z;
    }
    std::vector<int32_t> cand((size_t)nq * k);
    for (int q = 0; q < nq; q++)
        for (int j = 0; j < k; j++) cand[(size_t)q * k + j] = (q + 1 + j * 7) % n;
    float *d_bank, *d_roll, *d_score;
    int32_t *d_cand, *d_niter;
    long long* d_clk;
    CK(cudaMalloc(&d_bank, bank.size() * 4));
    CK(cudaMalloc(&d_roll, roll.size() * 4));
    CK(cudaMalloc(&d_score, (size_t)nq * k * 4));
    CK(cudaMalloc(&d_cand, cand.size() * 4));
    CK(cudaMalloc(&d_niter, nq * 4));
    CK(cudaMalloc(&d_clk, (size_t)nq * (16 + 28) * 8));
    CK(cudaMemset(d_clk, 0, (size_t)nq * (16 + 28) * 8));
    CK(cudaMemcpy(d_bank, bank.data(), bank.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_roll, roll.data(), roll.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_cand, cand.data(), cand.size() * 4, cudaMemcpyHostToDevice));
    vr::PairArgs a{};
    a.q_patches = d_bank; a.c_patches = d_bank; a.q_rollout = d_roll; a.c_rollout = d_roll;
    a.cand_idx = d_cand; a.cand_stride = k; a.q_start = 0; a.q_stride = 1; a.k = k;
    a.p.mode = VR_MODE_ROLLOUT; a.p.use_cls_token = 1; a.p.ot_temp = 0.05f; a.p.temperature = 0.1f; a.p.ot_part = 1.0f;
    a.p.max_iter = argc > 5 ? atoi(argv[5]) : 100; a.p.thresh = argc > 7 ? atof(argv[7]) : 0.1f;
    a.out_score = d_score; a.out_niter = d_niter;
#ifdef PR_TIMING
    a.dbg_clk = d_clk;
#endif
    if (argc > 6 && argv[6][0] == 'p') {   // feed S3 from the re-packed bank (TMA path), as vr_rerank_scores does
        void* packed;
        CK(cudaMalloc(&packed, vr::pair_fused_packed_bytes(n)));
        if (vr::pair_fused_repack(d_bank, n, 0, n, packed, 0)) return 1;
        a.c_packed_a = packed;
        a.q_packed_b = (const char*)packed + vr::pair_fused_packed_bytes(n) / 2;
        printf("re-packed bank (TMA-fed S3)\n");
    }
    int mc = 0;
    if (vr::pair_fused_max_clusters(&mc)) return 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        if (vr::pair_fused_launch(a, nq, 0)) return 1;
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    std::vector<int32_t> niter(nq);
    std::vector<float> score((size_t)nq * k);
    CK(cudaMemcpy(niter.data(), d_niter, nq * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(score.data(), d_score, score.size() * 4, cudaMemcpyDeviceToHost));
    double mi = 0, cs = 0;
    for (int q = 0; q < nq; q++) mi += niter[q];
    for (auto s : score) cs += s;
    mi /= nq;
    printf("n=%d nq=%d k=%d: %.3f ms, %.2f M pairs/s, %.2f us/query/cluster-slot (%d clusters), mean n*=%.1f, score checksum %.6f\n", n, nq, k,
           best, (double)nq * k / best / 1e3, best * 1e3 * mc / nq, mc, mi, cs);
#ifdef PR_TIMING
#ifdef PR_SKEW
    {
        std::vector<long long> sk((size_t)nq * 28);
        CK(cudaMemcpy(sk.data(), d_clk + (size_t)nq * 16, sk.size() * 8, cudaMemcpyDeviceToHost));
        double sp[3] = {0, 0, 0}, pro = 0, lp = 0;
        double late[7] = {0};
        for (int q = 0; q < nq; q++)
            for (int e = 0; e < 3; e++) {
                long long lo = 1ll << 62, hi = 0; int arg = 0;
                for (int r = 0; r < 7; r++) { long long t = sk[((size_t)q * 7 + r) * 4 + e]; lo = std::min(lo, t); if (t > hi) { hi = t; arg = r; } }
                sp[e] += (double)(hi - lo);
                if (e == 1) late[arg] += 1;
            }
        for (int q = 0; q < nq; q++)
            for (int r = 0; r < 7; r++) { pro += (double)(sk[((size_t)q * 7 + r) * 4 + 1] - sk[((size_t)q * 7 + r) * 4 + 0]); lp += (double)(sk[((size_t)q * 7 + r) * 4 + 2] - sk[((size_t)q * 7 + r) * 4 + 1]); }
        printf("  skew over the 7 CTAs of a query (ns, max - min): start %.0f, loop entry %.0f, loop exit %.0f; mean prologue %.0f ns, loop %.0f ns\n",
               sp[0] / nq, sp[1] / nq, sp[2] / nq, pro / nq / 7, lp / nq / 7);
        printf("  last CTA at loop entry by rank:");
        for (int r = 0; r < 7; r++) printf(" %.1f%%", 100 * late[r] / nq);
        printf("\n");
    }
#endif
    std::vector<long long> clk((size_t)nq * 16);
    CK(cudaMemcpy(clk.data(), d_clk, clk.size() * 8, cudaMemcpyDeviceToHost));
    const char* names[7] = {"cluster.sync", "S2+S3", "gibbs+K^T+TMEM", "marginals", "sinkhorn loop", "drain+dealloc", "score"};
    double tot = 0, ph[7] = {0};
    for (int q = 0; q < nq; q++)
        for (int i = 0; i < 7; i++) ph[i] += (double)(clk[q * 16 + i + 1] - clk[q * 16 + i]);
    for (int i = 0; i < 7; i++) tot += ph[i];
    for (int i = 0; i < 7; i++) printf("  %-16s %9.0f cycles (%4.1f%%)%s\n", names[i], ph[i] / nq, 100 * ph[i] / tot,
                                        i == 4 ? "" : "");
    double it3[3] = {0};
    for (int q = 0; q < nq; q++)
        for (int i = 0; i < 3; i++) it3[i] += (double)(clk[q * 16 + 9 + i] - clk[q * 16 + 8 + i]);
    double s3[5] = {0};
    for (int q = 0; q < nq; q++)
        for (int i = 0; i < 5; i++) s3[i] += (double)clk[q * 16 + 10 + i];
    printf("  S3 per chunk: (packed: wait TMA, issue MMAs, wait MMAs + reissue) / (direct: -, load+split, wait MMA, stores, -) %.0f %.0f %.0f %.0f %.0f cycles\n", s3[0] / nq / 8,
           s3[1] / nq / 8, s3[2] / nq / 8, s3[3] / nq / 8, s3[4] / nq / 8);
    double wc = 0, wn = 0;
    for (int q = 0; q < nq; q++) { wc += (double)clk[q * 16 + 9]; wn += (double)clk[q * 16 + 15]; }
    printf("  exchange re-polls of thread 0 per query: %.1f in steps 0..3, %.1f later\n", wc / nq, wn / nq);
    double fs = 0;
    for (int q = 0; q < nq; q++) fs += (double)clk[q * 16 + 8];
    printf("  fetch check (first look at the fetched partials, thread 0): %.0f cycles per iteration\n", fs / nq / (mi + 2));
    printf("  loop per iteration: %.0f cycles; whole CTA after setup: %.0f cycles\n", ph[4] / nq / mi, tot / nq);
#endif
    return 0;
}
