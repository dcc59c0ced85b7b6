// Internal (non-ABI) interfaces between the translation units of libvitrerank.
#pragma once
#include "common.cuh"

namespace vr {

struct PairArgs {
    const float* q_patches;   // query i block at q_patches + qid * C * R
    const float* q_centers;   // + qid * C   (cc modes with use_cls_token)
    const float* q_rollout;   // + qid * R   (rollout mode)
    const float* c_patches;   // candidate banks, indexed by candidate id
    const float* c_centers;
    const float* c_rollout;
    const int32_t* cand_idx;  // [nq, cand_stride] or nullptr (identity)
    int cand_stride;
    int64_t q_start, q_stride;
    int k;
    vr_ot_params p;
    float* out_score;         // [nq, k]
    int32_t* out_niter;       // [nq] or nullptr
    float *out_u, *out_v, *out_T, *out_simr, *out_cc;  // optional, [nq*k, ...]
    float* dbg_err;           // optional [nq, max_iter]: the stop-test value of every iteration
    long long* dbg_clk;       // optional [nq, 16]: phase clocks (only read by builds with -DPR_TIMING)
    unsigned long long* ex_part;   // exchange buffer of the global transport (set by pair_fused_launch)
    int group_ctas;           // CTAs per query of the wide path, ceil(k / 16) (set by pair_fused_launch)
    int halves;               // half-CTAs (8 pair slots) per query of the packed score-only launch, 0 = whole CTAs (set by pair_fused_launch)
    float part_bin;           // partial OT: 1 - ot_part as the reference rounds it (set by pair_fused_launch)
    const void* c_packed_a;   // re-packed candidate bank (pair_fused_repack) or nullptr: convert on the fly
    const void* q_packed_b;   // re-packed query bank, indexed by the query id
    int packed_centers;       // both operand copies carry the images' normalised centres as patch 50 (pack_image)
};

struct GenArgs {
    const float* q_patches;
    const float* q_centers;
    const float* q_rollout;
    const float* c_patches;
    const float* c_centers;
    const float* c_rollout;
    const int32_t* cand_idx;
    int cand_stride;
    int64_t q_start, q_stride, nq;
    int k, c, r;
    vr_ot_params p;
    float part_bin;   // 1 - ot_part as the reference rounds it (set by generic_rerank)
    const void* packed; // re-packed registered bank for generic_sim_mma (generic_repack), both roles; nullptr: convert per pair
    int packed_centers; // the operand copy carries every image's normalised centre as patch R (R % 16 != 0)
    int cc_stages;    // > 0: generic_prepare_kernel streams the fp32 rows for the cross-correlation marginals through a ring of this many
                      // bulk-copy stages (set by generic_rerank)
    int sim_done;     // sim and K were written by generic_sim_mma (tensor cores): generic_prepare_kernel skips its fp32 loop
    // workspace
    float* sim;    // [np, r, r]
    float* K;      // [np, re, re]
    float* u;      // [np, re]
    float* v;      // [np, re]
    float* rv;     // [np, rows]
    float* cv;     // [np, cols]
    float* e;      // [np]
    int32_t* done; // [nq]
    int32_t* niter;// [nq]
    // outputs
    float* out_score;
    int32_t* out_niter;
    float *out_u, *out_v, *out_T, *out_simr, *out_cc;
    float* dbg_err;
};

// stage0_topk.cu
// stats_dev (nullable): 4 uint32 on the device = {rows redone by the exact fallback, max |g|^2 bits, unsplittable-value flag,
// rows whose candidate buffer overflowed} of the tensor-core path; zeros for the fp32 paths
size_t stage0_workspace_bytes(int64_t nq, int64_t n, int c, int kp, int sms);
int stage0_topk(const float* q_centers, const int64_t* self_idx, const float* centers, int64_t q_start,
                int64_t q_stride, int64_t nq, int64_t n, int c, int kp, int32_t* out_idx, float* out_score,
                void* ws, size_t ws_bytes, int sms, uint32_t* stats_dev, cudaStream_t st);

// stage0_mma.cu
bool stage0_mma_supported(int64_t nq, int64_t n, int c, int kp);
size_t stage0_mma_workspace_bytes(int64_t nq, int64_t n, int kp, int sms);
int stage0_mma_topk(const float* q_centers, const int64_t* self_idx, const float* centers, int64_t q_start, int64_t q_stride,
                    int64_t nq, int64_t n, int kp, int32_t* out_idx, float* out_score, void* ws, size_t ws_bytes, int sms,
                    uint32_t* stats_dev, cudaStream_t st);
int global_similarity(const float* q, const float* centers, int64_t n, int c, float* sim, cudaStream_t st);

// pair_fused.cu
int pair_fused_max_clusters(int* out);
bool pair_fused_supports(int c, int r, int k, const vr_ot_params* p, bool scores_only);
float partial_ot_bin(float ot_part);   // 1 - ot_part rounded as diml.py:61 rounds it
bool pair_fused_supports_wide(int c, int r, int k, const vr_ot_params* p);   // 112 < k <= 1024, scores only
int pair_fused_launch(const PairArgs& a, int64_t nq, cudaStream_t st);
int pair_exchange_begin(cudaStream_t st, unsigned long long** part, size_t* bytes);   // exchange buffer + launch serialisation
int pair_exchange_end(cudaStream_t st);
int pair_fused_ctx_open(int device);    // vr_create / vr_destroy: the last context on a device frees the exchange buffer
void pair_fused_ctx_close(int device);
size_t pair_fused_packed_bytes(int64_t n);   // both roles
int pair_fused_repack(const float* patches, const float* centers, int64_t n, int64_t first, int64_t count, void* packed, cudaStream_t st);
int bank_ingest(const float* tokens, const float* centers_raw, int channel_major, int64_t n, int64_t first, int64_t count, int h,
                int w, int grid, int c, float* patches, float* centers, void* packed, cudaStream_t st);

// generic_ot.cu
size_t generic_rerank_workspace_bytes(int64_t nq, int k, int r, const vr_ot_params* p);
size_t generic_fused_workspace_bytes(int64_t nq, int k, int r);   // when generic_rerank takes generic_fused.cu
size_t generic_sinkhorn_workspace_bytes(int64_t b, int m, int n);
int generic_rerank(GenArgs a, void* ws, size_t ws_bytes, cudaStream_t st);
int generic_sinkhorn(const float* K, const float* u, const float* v, int64_t b, int m, int n, int max_iter,
                     float thresh, float* T, int32_t* niter, void* ws, size_t ws_bytes, cudaStream_t st);

// generic_s3.cu
bool generic_sim_mma_supported(int c, int r);
int generic_sim_mma(const GenArgs& g, int re, cudaStream_t st);
size_t generic_packed_image_bytes(int c, int r);
int generic_repack(const float* patches, const float* centers, int64_t n, int c, int r, void* packed, cudaStream_t st);

// generic_fused.cu: S3 + S4 in one kernel from the re-packed bank (full OT, rollout / uniform marginals, scores only)
bool generic_fused_supported(int c, int r, const vr_ot_params* p);
int generic_fused_rerank(const GenArgs& g, int32_t* list0, int32_t* list1, int32_t* counts, float* ehist, float* shist,
                         int32_t* niter, cudaStream_t st);

// rollout.cu: the attention-rollout producer (eval_cvt_diml.py:54-146)
size_t rollout_block_workspace_bytes(int64_t b, int ht, int wt, int drop_cls);
int rollout_block(const float* probs, int64_t b, int heads, int ht, int wt, int drop_cls, int grid, int64_t n_discard, int fusion,
                  float* out, void* ws, size_t ws_bytes, cudaStream_t st);
int rollout_chain(const float* mats, int n_mats, int64_t b, int n, int use_res, float* joints, cudaStream_t st);

// finalize.cu
size_t finalize_workspace_bytes(int64_t nq, int n_trunc);
int finalize(int64_t q_start, int64_t q_stride, int64_t nq, int k, int kp, const int32_t* approx_idx,
             const float* approx_score, const float* ot_score, const int64_t* labels, const int32_t* num_pos,
             const int32_t* truncs, int n_trunc, int32_t* out_rank, double* tallies, void* ws, size_t ws_bytes,
             cudaStream_t st);

int blend_rank(int64_t nq, int k, int kp, const int32_t* approx_idx, const float* approx_score, const float* ot_score,
               int32_t* out_rank, cudaStream_t st);

int num_pos_counts(const int64_t* labels, int64_t n, int32_t* num_pos, int32_t* max_dev, cudaStream_t st);

int metrics_rank(const int64_t* tops, int64_t n_tops, int64_t qlabel, const int64_t* labels, int64_t n_labels,
                 double* out, cudaStream_t st);

}  // namespace vr
