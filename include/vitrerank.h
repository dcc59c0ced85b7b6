/*
 * vitrerank.h -- C ABI of libvitrerank.so: the DIML structural-similarity rerank path of
 * cazhang/vit-reranking as hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI: the path is plain Python functions (utilities/diml.py,
 * evaluation/eval_cvt_diml.py, evaluation/metrics.py).  Each entry point below names the
 * reference function (file:line under the reference checkout) whose arithmetic it
 * replaces; the Python drop-in modules in vit-reranking_b200/{utilities,evaluation}/ keep
 * the reference's names and signatures and call these through ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative VR_E_* code otherwise;
 *     vr_last_error() returns a thread-local message for the last failure;
 *   - all tensor arguments are raw DEVICE pointers to contiguous fp32 / int32 / int64
 *     data unless the name ends in _host; the caller owns every buffer (inputs, outputs,
 *     workspace); the library frees only what vr_create / vr_*_host allocated;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); no entry
 *     point synchronises the host except the *_host ones, which return finished results;
 *   - layouts follow the reference: patches [N, C, R] (R contiguous), centers [N, C],
 *     rollout [N, R], labels [N] int64; sim / T matrices are [pairs, rows = candidate
 *     patch, cols = query patch] (utilities/diml.py:100).
 *   - no CPU fallback exists: without a CUDA device every compute entry fails.
 */
#ifndef VITRERANK_H_
#define VITRERANK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VR_ABI_VERSION 4

#if defined(__GNUC__)
#define VR_API __attribute__((visibility("default")))
#else
#define VR_API
#endif

enum {
    VR_OK = 0,
    VR_E_INVALID = -1,   /* bad argument (shape, null pointer, unsupported size) */
    VR_E_CUDA = -2,      /* a CUDA runtime call failed; message has the CUDA error */
    VR_E_NOBANK = -3,    /* vr_bank_register has not been called on this context */
    VR_E_WORKSPACE = -4  /* workspace too small; see the matching *_workspace_bytes */
};

/* Marginal modes: the branch order of utilities/diml.py:104-133 plus the rollout variant
 * utilities/diml.py:344-354. */
enum {
    VR_MODE_ROLLOUT = 0, /* u,v = relu(rollout)/(sum+1e-5)            diml.py:351-354 */
    VR_MODE_UNIFORM = 1, /* 1/R                                        diml.py:104-106 */
    VR_MODE_INVERSE = 2, /* exp(-relu(cc)/temperature)/(sum+1e-5)      diml.py:107-113 */
    VR_MODE_MINUS = 3,   /* (1-relu(cc))/(sum+1e-5)                    diml.py:114-121 */
    VR_MODE_SOFT = 4,    /* softmax(cc)/(sum+1e-5)                     diml.py:122-127 */
    VR_MODE_RELU = 5     /* relu(cc)/(sum+1e-5)                        diml.py:128-133 */
};

typedef struct vr_ot_params {
    int32_t mode;          /* VR_MODE_* */
    int32_t use_cls_token; /* cc modes: 1 = use the given centres, 0 = mean of patches (diml.py:87-96) */
    float ot_temp;         /* Gibbs kernel temperature, 0.05 in evaluate (eval_cvt_diml.py:341) */
    float temperature;     /* inverse-mode temperature (--temperature, diml.py:109) */
    float ot_part;         /* > 0.999: full OT (diml.py:79,135); else one dummy point (diml.py:59-75) */
    int32_t max_iter;      /* 100 (diml.py:42) */
    float thresh;          /* 0.1, absolute, on the batch mean of |r - r_prev| (diml.py:45-52) */
} vr_ot_params;

typedef struct vr_ctx vr_ctx;

/* Library / device -------------------------------------------------------------------- */
VR_API int vr_abi_version(void);
VR_API const char* vr_last_error(void);
/* The dummy point's mass of partial OT, 1 - ot_part, rounded as utilities/diml.py:61 rounds it (double subtraction of the
 * Python float, one rounding to fp32); ot_part travels as fp32, the decimal the caller typed is recovered.  Host arithmetic. */
VR_API float vr_partial_ot_pad(float ot_part);
/* Creates a context bound to CUDA device `device` (queries SM count, cluster support). */
VR_API int vr_create(int device, vr_ctx** out);
VR_API int vr_destroy(vr_ctx* ctx);
/* Number of SMs and the number of 8-CTA clusters of the fused pair kernel that can be
 * co-resident (cudaOccupancyMaxActiveClusters); for reporting. */
VR_API int vr_device_info(vr_ctx* ctx, int32_t* sm_count, int32_t* max_active_clusters);

/* Gallery -----------------------------------------------------------------------------
 * Replaces the bank construction at eval_cvt_diml.py:299-308 (banks are already
 * L2-normalised by the caller as at :304-305).  rollout / labels / num_pos may be NULL
 * when the mode / entry points used do not need them.  num_pos[i] = #{j : labels[j] ==
 * labels[i]} (metrics.py:34), int32.
 * Only pointers are recorded here.  The first fused rerank afterwards derives a library-owned
 * fp16 copy of `patches` in the operand layout of the patch-similarity kernel (64 KB per image,
 * on that call's stream); register again after modifying a registered bank in place. */
VR_API int vr_bank_register(vr_ctx* ctx, const float* patches, const float* centers, const float* rollout,
                     const int64_t* labels, const int32_t* num_pos, int64_t n, int32_t c, int32_t r);

/* Derives the library-owned operand copy of images [first, first + count) of the registered bank now, on `stream`
 * (ranges in ascending order without gaps; once [0, n) is covered the fused rerank uses the copy as it is).  Lets a caller
 * that fills the bank piecewise (uploads, all-gathers) overlap the re-pack with the transfers and with stage 0; later
 * rerank calls must be ordered after it by the caller (same stream or an event).  The copy also carries every image's normalised
 * centre (the cross-correlation marginals then come out of the patch-similarity MMA): patches AND centres of the range must be
 * valid on `stream`.  No-op for shapes without a fused kernel. */
VR_API int vr_bank_prepare(vr_ctx* ctx, int64_t first, int64_t count, void* stream);

/* Bank ingest: the step between the backbone and the path (evaluation/eval_cvt_diml.py:269-278 head output -> permute ->
 * AdaptiveAvgPool2d(grid_size); :304 F.normalize(feature_bank, dim=1); :305 F.normalize(feature_bank_center, dim=1)).
 * tokens: [count, h * w, C] (channel_major = 0: the head projection's output) or [count, C, h * w] (channel_major = 1: the
 * feature maps of the trained-model branch :286-289); centers_raw (nullable): [count, C] un-normalised global embeddings.
 * Writes rows [first, first + count) of the REGISTERED patch / centre banks (they are the destination: register the empty
 * buffers first; h, w must be whole multiples of the registered grid) and, for 128 x 49 banks, the library's operand copy
 * of those images, so no later re-pack pass is needed (ascending ranges without gaps, like vr_bank_prepare).  The results
 * equal torch's CPU AdaptiveAvgPool2d + F.normalize bit for bit. */
VR_API int vr_bank_ingest(vr_ctx* ctx, const float* tokens, const float* centers_raw, int32_t channel_major, int64_t first,
                          int64_t count, int32_t h, int32_t w, void* stream);

/* Attaches labels / class counts (device, [n]) to the registered bank without touching the banks or the operand copy
 * (for banks filled by vr_bank_ingest, whose labels arrive with the batches). */
VR_API int vr_bank_labels(vr_ctx* ctx, const int64_t* labels, const int32_t* num_pos);

/* num_pos[i] = #{j : labels[j] == labels[i]}: the `num_pos = torch.sum(gallery_label == query_label)` of
 * evaluation/metrics.py:34 for every gallery item at once (device int32 [n]; counts the item itself).  When
 * max_num_pos_host is not NULL the call synchronises `stream` and stores the largest count there (the first-stage
 * shortlist must be at least that long, see vr_finalize). */
VR_API int vr_num_pos(vr_ctx* ctx, const int64_t* labels, int64_t n, int32_t* num_pos, int32_t* max_num_pos_host,
                      void* stream);

/* S1: first-stage retrieval -----------------------------------------------------------
 * Replaces, for a batch of queries, calc_similarity(stage=0) (diml.py:83-85), the self
 * mask approx_sim[idx] = -100 (eval_cvt_diml.py:327) and the head of the full argsort
 * (:329-332): out_idx[i, :] are the kp best gallery indices of query i in descending
 * score order (ties: lower index first), out_score the matching fp32 scores; rows are
 * padded with idx -1 when n < kp.  For batches (>= 256 queries) over 128-d embeddings with
 * kp <= 256 the scores are a tcgen05 GEMM whose accumulators are filtered as they leave
 * tensor memory: the N x N score matrix is never written to memory, and the lists are
 * bit-identical to the fp32 FMA-chain path (shortlist by the tensor-core score, canonical
 * fp32 re-score, rounding-bound acceptance test, exact fallback for rows that fail it).
 * A few queries take a fused fp32 kernel (no matrix in memory either); batches with other
 * widths or longer shortlists write [rows, N] score chunks to the workspace.
 * Query i is gallery item q_start + i * q_stride (self-masked) when q_centers is NULL;
 * otherwise q_centers is [nq, c] and self_idx (nullable, int64 [nq]) names the gallery
 * item to mask for each query (the query != gallery case of training_tools/val.py:159-190). */
VR_API size_t vr_stage0_workspace_bytes(vr_ctx* ctx, int64_t nq, int32_t kp);
VR_API int vr_stage0_topk(vr_ctx* ctx, const float* q_centers, const int64_t* self_idx, int64_t q_start,
                   int64_t q_stride, int64_t nq, int32_t kp, int32_t* out_idx, float* out_score,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Counters of the last vr_stage0_topk call on this context (synchronises `stream`): out4[0] = rows the tensor-core path
 * handed to its exact fp32 fallback, out4[3] = rows whose candidate buffer overflowed (a subset), out4[2] = 1 when the
 * centres held values the fp16 split cannot represent (then every row is redone), out4[1] = bits of max |g|^2. */
VR_API int vr_stage0_stats(vr_ctx* ctx, uint32_t* out4_host, void* stream);

/* S2-S5a: structural scores of query/candidate pairs ------------------------------------
 * Replaces the stage-1 call at eval_cvt_diml.py:334-351 for a batch of queries taken from
 * the registered bank: candidate gather (:335-336,346-348), patch similarity
 * (diml.py:100/:339), Gibbs kernel (:101-102/:341-342), marginals (:104-133/:344-354),
 * Sinkhorn (:42-54, or Sinkhorn_partial :59-75 when ot_part <= 0.999) with its
 * batch-global stop per query, and the score sum(T * sim) (:142-143/:361-362).
 * cand_idx is [nq, cand_stride] int32 (first k entries of each row are used; -1 entries
 * are skipped and score 0).  out_score is [nq, k]; out_niter (nullable) [nq] receives the
 * number of Sinkhorn iterations each query ran.
 * Shapes other than C = 128, R = 49 take the shape-generic kernels.  For a registered bank with C % 16 == 0 and
 * 20 <= R <= 224 the library keeps an operand copy (fp16 hi / lo halves, C x ceil16(R) x 4 bytes per image, at most
 * 48 GB, allocated by the first vr_rerank_workspace_bytes / vr_rerank_scores call after a registration); with it,
 * full OT and rollout / uniform marginals, S3 + S4 run in one kernel and the workspace is 1.6 KB per pair, otherwise
 * 2 (R + 1)^2 floats per pair.  A smaller workspace than vr_rerank_workspace_bytes returns is legal as long as one
 * query fits: the queries are then processed in rounds. */
VR_API size_t vr_rerank_workspace_bytes(vr_ctx* ctx, int64_t nq, int32_t k, const vr_ot_params* p);
VR_API int vr_rerank_scores(vr_ctx* ctx, int64_t q_start, int64_t q_stride, int64_t nq, int32_t k,
                     const int32_t* cand_idx, int32_t cand_stride, const vr_ot_params* p,
                     float* out_score, int32_t* out_niter, void* workspace, size_t workspace_bytes,
                     void* stream);

/* The same for queries that are NOT gallery items (training_tools/val.py:159-190: MSLS query images against the database
 * images of a city): q_patches [nq, c, r], q_centers [nq, c] (cc modes with use_cls_token), q_rollout [nq, r] (rollout mode);
 * candidates come from the registered bank.  Pair with vr_stage0_topk(q_centers, self_idx = NULL). */
VR_API int vr_rerank_scores_queries(vr_ctx* ctx, const float* q_patches, const float* q_centers, const float* q_rollout,
                             int64_t nq, int32_t k, const int32_t* cand_idx, int32_t cand_stride, const vr_ot_params* p,
                             float* out_score, int32_t* out_niter, void* workspace, size_t workspace_bytes, void* stream);

/* S5b: blend, re-sort, metrics --------------------------------------------------------
 * Replaces eval_cvt_diml.py:357-372 and evaluation/metrics.py:26-47 for a batch of
 * queries: total = ot_score + approx_score, descending argsort (NaN first, ties: lower
 * position first), and for every trunc t: final = reranked[:t] ++ approx_tops[t:]
 * (t = 0: approx_tops), r1 / R-precision / MAP@R over final[:num_pos], plus Recall@1,2,4,8
 * (extension).  approx_idx/approx_score are the [nq, kp] outputs of vr_stage0_topk (kp >=
 * max(k, max num_pos) unless the gallery is smaller).  out_rank (nullable) [nq, k]
 * receives the reranked candidate indices.  Adds into tallies [n_trunc, 8] doubles:
 * sum r1, sum rp, sum mapr, sum R@1, R@2, R@4, R@8, query count (caller zeroes them). */
VR_API size_t vr_finalize_workspace_bytes(vr_ctx* ctx, int64_t nq, int32_t n_trunc);
VR_API int vr_finalize(vr_ctx* ctx, int64_t q_start, int64_t q_stride, int64_t nq, int32_t k, int32_t kp,
                const int32_t* approx_idx, const float* approx_score, const float* ot_score,
                const int32_t* trunc_nums_host, int32_t n_trunc, int32_t* out_rank, double* tallies,
                void* workspace, size_t workspace_bytes, void* stream);

/* The blend + re-sort alone (eval_cvt_diml.py:357, training_tools/val.py:197): out_rank[q, :] = approx_idx[q, rank] with
 * rank = argsort(ot_score[q] + approx_score[q, :k], descending) (NaN first, ties: lower position first).  No labels needed. */
VR_API int vr_blend_rank(vr_ctx* ctx, int64_t nq, int32_t k, int32_t kp, const int32_t* approx_idx, const float* approx_score,
                  const float* ot_score, int32_t* out_rank, void* stream);

/* Attention-rollout producer (the input of --use_rollout; evaluation/eval_cvt_diml.py:54-146) ---------------
 * vr_rollout_block replaces filter_attention_map(probs, discard_ratio, head_fusion) (:74-108) followed by
 * resize_attn_map(., resize, stage, grid) (:54-71) for one transformer block: probs is the block's attention
 * [b, heads, ht, wt] fp32 (device); the heads are fused (fusion 1 = max, 2 = min); the n_discard =
 * int(ht * wt * discard_ratio) smallest entries of every image are found (ties: lower index first) and the UNION of
 * their coordinates over the batch is zeroed in every image, as the reference's fancy-index assignment does; with
 * drop_cls (stage 2, :57-58) row 0 and column 0 are then dropped (H = ht - 1, W = wt - 1); both token axes, square
 * grids, are pooled to grid x grid with AdaptiveAvgPool2d's arithmetic, the key axis first.  out: [b, grid^2, grid^2]
 * (query cell, key cell).  Workspace: vr_rollout_block_workspace_bytes (the fused map, b * ht * wt floats).
 * vr_rollout_chain replaces :130-141: mats [n_mats, b, n, n] (the stacked block maps); with use_res the identity is
 * added and every row divided by its sum (ATen's summation order); joints [n_mats, b, n, n] receives joint[0] =
 * mats[0], joint[j] = mats[j] joint[j - 1] (every entry one FMA chain over the inner index; the reference's CPU
 * bmm may round differently).  The marginal of an image is joints[n_mats - 1].mean(1). */
VR_API size_t vr_rollout_block_workspace_bytes(int64_t b, int32_t ht, int32_t wt, int32_t drop_cls);
VR_API int vr_rollout_block(vr_ctx* ctx, const float* probs, int64_t b, int32_t heads, int32_t ht, int32_t wt, int32_t drop_cls,
                     int32_t grid, int64_t n_discard, int32_t fusion, float* out, void* workspace, size_t workspace_bytes,
                     void* stream);
VR_API int vr_rollout_chain(vr_ctx* ctx, const float* mats, int32_t n_mats, int64_t b, int32_t n, int32_t use_res, float* joints,
                     void* stream);

/* Direct calls (the per-call surface of utilities/diml.py) ------------------------------
 * vr_sinkhorn replaces Sinkhorn(K, u, v, iter) (diml.py:42-54) for any [b, m, n] batch:
 * T [b, m, n]; niter (nullable) one int32.  Workspace: vr_sinkhorn_workspace_bytes. */
VR_API size_t vr_sinkhorn_workspace_bytes(int64_t b, int32_t m, int32_t n);
VR_API int vr_sinkhorn(const float* K, const float* u, const float* v, int64_t b, int32_t m, int32_t n,
                int32_t max_iter, float thresh, float* T, int32_t* niter, void* workspace,
                size_t workspace_bytes, void* stream);

/* vr_calc_similarity replaces stage 1 of calc_similarity (diml.py:86-147) and of
 * calc_similarity_cvt_rollout (diml.py:331-366) for explicit tensors: anchor [c, r],
 * anchor_center [c] (may be NULL in rollout/uniform mode), q_rollout [r] (rollout mode),
 * fb [n, c, r], fb_center [n, c], c_rollout [n, r].  Outputs: score [n]; optional (NULL to
 * skip) u [n, r], v [n, r], T [n, re, re] with re = r (full) or r + 1 (partial),
 * sim_r [n, r, r], cc [n, r] (minus: cc_u; soft / relu: cc_v; other modes untouched),
 * niter one int32.  The Sinkhorn stop is global over the n candidates, as in the reference. */
VR_API size_t vr_calc_similarity_workspace_bytes(int64_t n, int32_t c, int32_t r, const vr_ot_params* p);
VR_API int vr_calc_similarity(vr_ctx* ctx, const float* anchor, const float* anchor_center, const float* q_rollout,
                       const float* fb, const float* fb_center, const float* c_rollout, int64_t n,
                       int32_t c, int32_t r, const vr_ot_params* p, float* score, float* u, float* v,
                       float* T, float* sim_r, float* cc, int32_t* niter, void* workspace,
                       size_t workspace_bytes, void* stream);

/* vr_global_similarity replaces calc_similarity(stage=0) (diml.py:83-85) for one query:
 * sim[n] = sum_c q[c] * centers[n, c]. */
VR_API int vr_global_similarity(const float* q_center, const float* centers, int64_t n, int32_t c, float* sim,
                         void* stream);

/* vr_metrics_rank replaces get_metrics_rank(tops, query_label, gallery_label)
 * (evaluation/metrics.py:26-47) for one ranked list: tops [n_tops] int64 (device), labels [n_labels]
 * int64 (device); out3 (device, 3 doubles) = r1, R-precision, MAP@R. */
VR_API int vr_metrics_rank(const int64_t* tops, int64_t n_tops, int64_t query_label, const int64_t* labels,
                           int64_t n_labels, double* out3, void* stream);

/* Whole pass with HOST buffers (the end-to-end entry) -----------------------------------
 * Replaces the query loop eval_cvt_diml.py:308-372 over host-resident banks, as the
 * reference keeps them (feature_bank and rollout_list live on the CPU, :278,:256): copies
 * the banks to the device, runs S1..S5 for queries q_start + i * q_stride (i < nq) and
 * returns tallies_host [n_trunc, 8] (same layout as vr_finalize; not yet divided by
 * N/100).  per_query_niter_host (nullable) [nq].  Device memory is taken from an arena
 * owned by ctx (grown on demand, released by vr_destroy).  Synchronous. */
VR_API int vr_evaluate_host(vr_ctx* ctx, const float* patches_host, const float* centers_host,
                     const float* rollout_host, const int64_t* labels_host, int64_t n, int32_t c, int32_t r,
                     int64_t q_start, int64_t q_stride, int64_t nq, const int32_t* trunc_nums_host,
                     int32_t n_trunc, const vr_ot_params* p, double* tallies_host,
                     int32_t* per_query_niter_host);

/* Same pass over the bank registered with vr_bank_register (device-resident inputs);
 * tallies_host as above.  kp_hint = 0 lets the library size the first-stage shortlist
 * from max_num_pos (the largest num_pos value, supplied by the caller). */
VR_API int vr_evaluate_registered(vr_ctx* ctx, int64_t q_start, int64_t q_stride, int64_t nq,
                           const int32_t* trunc_nums_host, int32_t n_trunc, int32_t max_num_pos,
                           const vr_ot_params* p, double* tallies_host, int32_t* per_query_niter_host,
                           void* stream);

/* Diagnostics: when buf is non-NULL, vr_rerank_scores / vr_calc_similarity also write the value of
 * the stop test (batch mean of |r - r_prev|, diml.py:50) of every iteration they run into
 * buf[query * max_iter + iteration] (device memory, caller-owned, nq * max_iter floats).  NULL turns
 * it off.  Used by the parity harness to show how close to the 0.1 threshold a stop was. */
VR_API int vr_debug_err_trace(vr_ctx* ctx, float* buf);

/* Number of kernels this library launched since the last call (resets the counter). */
VR_API int64_t vr_take_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VITRERANK_H_ */
