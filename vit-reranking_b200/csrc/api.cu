// C ABI of libvitrerank.so (see include/vitrerank.h for the contract and the reference
// file:line each entry replaces).
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace vr {

static thread_local char g_err[512] = "";
thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace vr

struct vr_ctx {
    int device = 0;
    int sms = 0;
    int max_clusters = 0;
    // registered gallery (device pointers owned by the caller, or by the arena for *_host)
    const float* patches = nullptr;
    const float* centers = nullptr;
    const float* rollout = nullptr;
    const int64_t* labels = nullptr;
    const int32_t* num_pos = nullptr;
    int64_t n = 0;
    int c = 0, r = 0;
    // arena: named device buffers that only grow
    std::unordered_map<std::string, std::pair<void*, size_t>> arena;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // uploads the patch bank piece by piece while stage 0 runs (vr_evaluate_host)
    cudaStream_t prep_stream = nullptr;   // re-packs every piece as it lands, under the upload of the next one
    cudaEvent_t piece_ready[8] = {};      // recorded on copy_stream after each piece
    cudaEvent_t small_ready = nullptr;    // recorded on own_stream after the centres / rollout / labels of vr_evaluate_host went up
    cudaEvent_t patches_ready = nullptr;  // recorded on prep_stream; non-null pending_wait makes the next rerank wait for it
    bool pending_wait = false;
    float* dbg_err = nullptr;  // see vr_debug_err_trace
    bool packed_valid = false; // the fp16 re-pack of `patches` (arena "packed") matches the registered bank
    bool gpacked_valid = false; // the same for the generic path's operand copy (arena "gpacked")
    int64_t packed_hi = 0;     // images [0, packed_hi) have been re-packed by vr_bank_prepare calls since the registration
    int32_t* pinned = nullptr; // 16 pinned int32 for small device -> host reads (max num_pos)
};

using namespace vr;

static int arena_get(vr_ctx* ctx, const char* name, size_t bytes, void** out) {
    auto& slot = ctx->arena[name];
    if (slot.second < bytes || slot.first == nullptr) {
        if (slot.first) VR_CHECK_CUDA(cudaFree(slot.first));
        slot.first = nullptr;
        slot.second = 0;
        size_t want = align_up(bytes ? bytes : 256, 256);
        VR_CHECK_CUDA(cudaMalloc(&slot.first, want));
        slot.second = want;
    }
    *out = slot.first;
    return VR_OK;
}

static int check_params(const vr_ot_params* p) {
    VR_REQUIRE(p != nullptr, "ot params missing");
    VR_REQUIRE(p->mode >= VR_MODE_ROLLOUT && p->mode <= VR_MODE_RELU, "unknown marginal mode %d", p->mode);
    VR_REQUIRE(p->ot_temp > 0.f, "ot_temp must be positive");
    VR_REQUIRE(p->max_iter >= 0 && p->max_iter <= 100000, "max_iter out of range");
    VR_REQUIRE(p->ot_part > 0.999f || (p->ot_part >= 0.f && p->ot_part < 1.f), "ot_part must be in [0, 1]");
    VR_REQUIRE(p->mode != VR_MODE_INVERSE || p->temperature != 0.f, "temperature must be non-zero");
    return VR_OK;
}

extern "C" {

int vr_abi_version(void) { return VR_ABI_VERSION; }

float vr_partial_ot_pad(float ot_part) { return partial_ot_bin(ot_part); }

const char* vr_last_error(void) { return g_err; }

int64_t vr_take_launch_count(void) {
    long long v = g_launches;
    g_launches = 0;
    return v;
}

int vr_create(int device, vr_ctx** out) {
    VR_REQUIRE(out != nullptr, "vr_create: out is null");
    int count = 0;
    VR_CHECK_CUDA(cudaGetDeviceCount(&count));
    VR_REQUIRE(device >= 0 && device < count, "vr_create: device %d not present (%d devices)", device, count);
    VR_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VR_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    VR_REQUIRE(prop.major == 10, "vr_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
               prop.major, prop.minor);
    vr_ctx* ctx = new vr_ctx();
    ctx->device = device;
    ctx->sms = prop.multiProcessorCount;
    int mc = 0;
    int rc = pair_fused_max_clusters(&mc);
    if (rc) {
        delete ctx;
        return rc;
    }
    ctx->max_clusters = mc;
    if ((rc = pair_fused_ctx_open(device))) {
        delete ctx;
        return rc;
    }
    rc = (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
          cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
          cudaStreamCreateWithFlags(&ctx->prep_stream, cudaStreamNonBlocking) == cudaSuccess &&
          cudaEventCreateWithFlags(&ctx->patches_ready, cudaEventDisableTiming) == cudaSuccess &&
          cudaEventCreateWithFlags(&ctx->small_ready, cudaEventDisableTiming) == cudaSuccess)
             ? VR_OK
             : VR_E_CUDA;
    for (int i = 0; i < 8 && rc == VR_OK; i++)
        if (cudaEventCreateWithFlags(&ctx->piece_ready[i], cudaEventDisableTiming) != cudaSuccess) rc = VR_E_CUDA;
    if (rc == VR_OK && cudaHostAlloc((void**)&ctx->pinned, 64, cudaHostAllocDefault) != cudaSuccess) rc = VR_E_CUDA;
    if (rc) {
        set_error("vr_create: cudaStreamCreate failed");
        pair_fused_ctx_close(device);
        delete ctx;
        return rc;
    }
    *out = ctx;
    return VR_OK;
}

int vr_destroy(vr_ctx* ctx) {
    if (!ctx) return VR_OK;
    cudaSetDevice(ctx->device);
    for (auto& kv : ctx->arena)
        if (kv.second.first) cudaFree(kv.second.first);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->prep_stream) cudaStreamDestroy(ctx->prep_stream);
    for (int i = 0; i < 8; i++)
        if (ctx->piece_ready[i]) cudaEventDestroy(ctx->piece_ready[i]);
    if (ctx->patches_ready) cudaEventDestroy(ctx->patches_ready);
    if (ctx->small_ready) cudaEventDestroy(ctx->small_ready);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    pair_fused_ctx_close(ctx->device);
    delete ctx;
    return VR_OK;
}

int vr_num_pos(vr_ctx* ctx, const int64_t* labels, int64_t n, int32_t* num_pos, int32_t* max_num_pos_host, void* stream) {
    VR_REQUIRE(ctx, "ctx is null");
    VR_REQUIRE(labels && num_pos && n > 0, "num_pos: bad arguments");
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    void* d_max = nullptr;
    int rc = arena_get(ctx, "np_max", 256, &d_max);
    if (rc) return rc;
    if ((rc = num_pos_counts(labels, n, num_pos, (int32_t*)d_max, st))) return rc;
    if (max_num_pos_host) {
        VR_CHECK_CUDA(cudaMemcpyAsync(ctx->pinned, d_max, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        VR_CHECK_CUDA(cudaStreamSynchronize(st));
        *max_num_pos_host = ctx->pinned[0];
    }
    return VR_OK;
}

int vr_debug_err_trace(vr_ctx* ctx, float* buf) {
    VR_REQUIRE(ctx, "ctx is null");
    ctx->dbg_err = buf;
    return VR_OK;
}

int vr_device_info(vr_ctx* ctx, int32_t* sm_count, int32_t* max_active_clusters) {
    VR_REQUIRE(ctx, "ctx is null");
    if (sm_count) *sm_count = ctx->sms;
    if (max_active_clusters) *max_active_clusters = ctx->max_clusters;
    return VR_OK;
}

int vr_bank_register(vr_ctx* ctx, const float* patches, const float* centers, const float* rollout,
                     const int64_t* labels, const int32_t* num_pos, int64_t n, int32_t c, int32_t r) {
    VR_REQUIRE(ctx, "ctx is null");
    VR_REQUIRE(patches && centers, "bank_register: patches and centers are required");
    VR_REQUIRE(n > 0 && c > 0 && r > 0, "bank_register: bad shape [%lld, %d, %d]", (long long)n, c, r);
    VR_REQUIRE(((uintptr_t)patches & 15) == 0 && ((uintptr_t)centers & 15) == 0, "bank_register: banks must be 16-byte aligned");
    ctx->patches = patches;
    ctx->centers = centers;
    ctx->rollout = rollout;
    ctx->labels = labels;
    ctx->num_pos = num_pos;
    ctx->n = n;
    ctx->c = c;
    ctx->r = r;
    ctx->packed_valid = false;   // re-packed by vr_bank_prepare, or lazily by the first fused rerank on that call's stream
    ctx->gpacked_valid = false;
    ctx->packed_hi = 0;
    return VR_OK;
}

int vr_bank_prepare(vr_ctx* ctx, int64_t first, int64_t count, void* stream) {
    VR_REQUIRE(ctx, "ctx is null");
    if (!ctx->patches) {
        set_error("bank_prepare: no bank registered");
        return VR_E_NOBANK;
    }
    if (ctx->c != 128 || ctx->r != 49) return VR_OK;   // only the fused 7x7 / 128-channel kernel keeps an operand copy
    VR_REQUIRE(first >= 0 && count > 0 && first + count <= ctx->n, "bank_prepare: range [%lld, %lld) outside the bank",
               (long long)first, (long long)(first + count));
    VR_REQUIRE(first <= ctx->packed_hi, "bank_prepare: ranges must be prepared in ascending order without gaps");
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    void* packed = nullptr;
    int rc = arena_get(ctx, "packed", pair_fused_packed_bytes(ctx->n), &packed);
    if (rc) return rc;
    if ((rc = pair_fused_repack(ctx->patches, ctx->centers, ctx->n, first, count, packed, (cudaStream_t)stream))) return rc;
    ctx->packed_hi = std::max(ctx->packed_hi, first + count);
    if (ctx->packed_hi >= ctx->n) ctx->packed_valid = true;
    return VR_OK;
}

int vr_bank_labels(vr_ctx* ctx, const int64_t* labels, const int32_t* num_pos) {
    VR_REQUIRE(ctx, "ctx is null");
    if (!ctx->patches) {
        set_error("bank_labels: no bank registered");
        return VR_E_NOBANK;
    }
    ctx->labels = labels;
    ctx->num_pos = num_pos;
    return VR_OK;
}

int vr_bank_ingest(vr_ctx* ctx, const float* tokens, const float* centers_raw, int32_t channel_major, int64_t first,
                   int64_t count, int32_t h, int32_t w, void* stream) {
    VR_REQUIRE(ctx, "ctx is null");
    if (!ctx->patches) {
        set_error("bank_ingest: register the (empty) destination banks first");
        return VR_E_NOBANK;
    }
    int grid = 1;
    while (grid * grid < ctx->r) grid++;
    VR_REQUIRE(grid * grid == ctx->r, "bank_ingest: the registered bank has %d patches per image, not a square grid", ctx->r);
    VR_REQUIRE(tokens && count > 0 && first >= 0 && first + count <= ctx->n, "bank_ingest: range [%lld, %lld) outside the bank",
               (long long)first, (long long)(first + count));
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    void* packed = nullptr;
    const bool fused_shape = ctx->c == 128 && ctx->r == 49;
    if (fused_shape) {
        VR_REQUIRE(first <= ctx->packed_hi, "bank_ingest: ranges must be ingested in ascending order without gaps");
        int rc = arena_get(ctx, "packed", pair_fused_packed_bytes(ctx->n), &packed);
        if (rc) return rc;
    }
    int rc = bank_ingest(tokens, centers_raw, channel_major, ctx->n, first, count, h, w, grid, ctx->c, const_cast<float*>(ctx->patches),
                         centers_raw ? const_cast<float*>(ctx->centers) : nullptr, packed, (cudaStream_t)stream);
    if (rc) return rc;
    ctx->gpacked_valid = false;   // the bank changed under the generic path's operand copy
    if (fused_shape) {
        ctx->packed_hi = std::max(ctx->packed_hi, first + count);
        if (ctx->packed_hi >= ctx->n) ctx->packed_valid = true;
    }
    return VR_OK;
}

size_t vr_stage0_workspace_bytes(vr_ctx* ctx, int64_t nq, int32_t kp) {
    if (!ctx || ctx->n <= 0) return 0;
    return stage0_workspace_bytes(nq, ctx->n, ctx->c, kp, ctx->sms);
}

int vr_stage0_topk(vr_ctx* ctx, const float* q_centers, const int64_t* self_idx, int64_t q_start, int64_t q_stride,
                   int64_t nq, int32_t kp, int32_t* out_idx, float* out_score, void* workspace,
                   size_t workspace_bytes, void* stream) {
    VR_REQUIRE(ctx, "ctx is null");
    if (!ctx->centers) {
        set_error("stage0: no bank registered");
        return VR_E_NOBANK;
    }
    VR_REQUIRE(out_idx && out_score, "stage0: outputs are null");
    if (!q_centers)
        VR_REQUIRE(q_start >= 0 && q_start + (nq - 1) * q_stride < ctx->n && q_start + (nq - 1) * q_stride >= 0,
                   "stage0: query range outside the gallery");
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    void* stats = nullptr;
    int rc = arena_get(ctx, "s0_stats", 256, &stats);
    if (rc) return rc;
    return stage0_topk(q_centers, self_idx, ctx->centers, q_start, q_stride, nq, ctx->n, ctx->c, kp, out_idx, out_score,
                       workspace, workspace_bytes, ctx->sms, (uint32_t*)stats, (cudaStream_t)stream);
}

int vr_stage0_stats(vr_ctx* ctx, uint32_t* out4_host, void* stream) {
    VR_REQUIRE(ctx && out4_host, "stage0_stats: bad arguments");
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    void* stats = nullptr;
    int rc = arena_get(ctx, "s0_stats", 256, &stats);
    if (rc) return rc;
    VR_CHECK_CUDA(cudaMemcpyAsync(ctx->pinned + 4, stats, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    VR_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    memcpy(out4_host, ctx->pinned + 4, 16);
    return VR_OK;
}

// The generic path's operand copy of the registered bank (generic_s3.cu: both MMA roles, c x rp16 x 4 bytes per image -- 639 KB at
// C = 768, R = 196), allocated on first use when it fits in 48 GB; nullptr: every pair converts its rows on the fly
// (shape not supported by the tensor-core S3, VR_GENERIC_PACK=0, or no memory).  Contents: ctx->gpacked_valid.
static void* generic_operand_copy(vr_ctx* ctx) {
    if (!ctx->patches || !generic_sim_mma_supported(ctx->c, ctx->r)) return nullptr;
    const char* e = getenv("VR_GENERIC_PACK");
    if (e && e[0] == '0') return nullptr;
    const size_t need = (size_t)ctx->n * generic_packed_image_bytes(ctx->c, ctx->r);
    if (need > ((size_t)48 << 30)) return nullptr;
    auto& slot = ctx->arena["gpacked"];
    if (slot.first && slot.second >= need) return slot.first;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    if (slot.first) cudaFree(slot.first);
    slot.first = nullptr;
    slot.second = 0;
    ctx->gpacked_valid = false;
    void* gp = nullptr;
    if (cudaMalloc(&gp, need) != cudaSuccess) {
        (void)cudaGetLastError();   // not enough memory: the converter path needs none
        return nullptr;
    }
    slot.first = gp;
    slot.second = need;
    return gp;
}

// A score-only rerank of queries FROM the registered bank takes generic_fused.cu (S3 + S4 in one kernel, 1.6 KB of workspace per pair)
static bool generic_fused_ready(vr_ctx* ctx, const vr_ot_params* p) {
    return generic_fused_supported(ctx->c, ctx->r, p) && (p->mode != VR_MODE_ROLLOUT || ctx->rollout) && generic_operand_copy(ctx) != nullptr;
}

size_t vr_rerank_workspace_bytes(vr_ctx* ctx, int64_t nq, int32_t k, const vr_ot_params* p) {
    if (!ctx || !p || ctx->n <= 0) return 0;
    if (pair_fused_supports(ctx->c, ctx->r, k, p, !ctx->dbg_err) || (!ctx->dbg_err && pair_fused_supports_wide(ctx->c, ctx->r, k, p))) return 256;
    // (queries from other banks -- vr_rerank_scores_queries -- take the separate kernels: with less than their 2 R^2 floats per
    // pair for all queries at once they run as many queries per round as fit, never less than one)
    if (generic_fused_ready(ctx, p)) return std::max(generic_fused_workspace_bytes(nq, k, ctx->r), generic_rerank_workspace_bytes(1, k, ctx->r, p));
    return generic_rerank_workspace_bytes(nq, k, ctx->r, p);
}

// Queries either from the registered bank (q_* null: items q_start + i * q_stride) or from explicit banks [nq, ...]
// (the query != gallery case of training_tools/val.py:159-190).
static int rerank_scores_impl(vr_ctx* ctx, const float* q_patches, const float* q_centers, const float* q_rollout, int64_t q_start,
                              int64_t q_stride, int64_t nq, int32_t k, const int32_t* cand_idx, int32_t cand_stride,
                              const vr_ot_params* p, float* out_score, int32_t* out_niter, void* workspace, size_t workspace_bytes,
                              void* stream) {
    VR_REQUIRE(ctx, "ctx is null");
    if (!ctx->patches) {
        set_error("rerank: no bank registered");
        return VR_E_NOBANK;
    }
    int rc = check_params(p);
    if (rc) return rc;
    const bool ext = q_patches != nullptr;
    VR_REQUIRE(nq > 0 && k > 0 && cand_idx && out_score && cand_stride >= k, "rerank: bad arguments");
    VR_REQUIRE(p->mode != VR_MODE_ROLLOUT || (ctx->rollout && (!ext || q_rollout)), "rerank: rollout mode needs the rollout banks");
    VR_REQUIRE(!(ext && p->mode >= VR_MODE_INVERSE && p->use_cls_token && !q_centers), "rerank: query centres missing");
    if (!ext)
        VR_REQUIRE(q_start >= 0 && q_start + (nq - 1) * q_stride < ctx->n && q_start + (nq - 1) * q_stride >= 0,
                   "rerank: query range outside the gallery");
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->pending_wait) {   // vr_evaluate_host: the patch bank is still being uploaded on the copy stream
        VR_CHECK_CUDA(cudaStreamWaitEvent(st, ctx->patches_ready, 0));
        ctx->pending_wait = false;
    }
    // (the err trace of vr_debug_err_trace is a diagnostics output: shortlists beyond 112 then take the generic solver)
    if (pair_fused_supports(ctx->c, ctx->r, k, p, !ctx->dbg_err) || (!ctx->dbg_err && pair_fused_supports_wide(ctx->c, ctx->r, k, p))) {
        PairArgs a{};
        a.q_patches = ext ? q_patches : ctx->patches;
        a.q_centers = ext ? q_centers : ctx->centers;
        a.q_rollout = ext ? q_rollout : ctx->rollout;
        a.c_patches = ctx->patches;
        a.c_centers = ctx->centers;
        a.c_rollout = ctx->rollout;
        a.cand_idx = cand_idx;
        a.cand_stride = cand_stride;
        a.q_start = ext ? 0 : q_start;
        a.q_stride = ext ? 1 : q_stride;
        a.k = k;
        a.p = *p;
        a.out_score = out_score;
        a.out_niter = out_niter;
        a.dbg_err = ctx->dbg_err;
        // one-time re-pack of the registered bank into the fp16 operand planes of S3 (a registered bank must not be
        // modified in place without registering it again)
        void* packed = nullptr;
        if ((rc = arena_get(ctx, "packed", pair_fused_packed_bytes(ctx->n), &packed))) return rc;
        if (!ctx->packed_valid) {
            if ((rc = pair_fused_repack(ctx->patches, ctx->centers, ctx->n, 0, ctx->n, packed, st))) return rc;
            ctx->packed_valid = true;
        }
        a.c_packed_a = packed;
        a.q_packed_b = (const char*)packed + pair_fused_packed_bytes(ctx->n) / 2;
        if (ext) {   // explicit query bank: its operand planes are derived per call
            void* packed_q = nullptr;
            if ((rc = arena_get(ctx, "packed_q", pair_fused_packed_bytes(nq), &packed_q))) return rc;
            if ((rc = pair_fused_repack(q_patches, q_centers, nq, 0, nq, packed_q, st))) return rc;
            a.q_packed_b = (const char*)packed_q + pair_fused_packed_bytes(nq) / 2;
        }
        // the operand copies carry the images' normalised centres (pack_image): cross-correlations with cls centres come out of the MMA
        a.packed_centers = (ctx->centers && (!ext || q_centers) && !(getenv("VR_PAIR_CC") && getenv("VR_PAIR_CC")[0] == 'f')) ? 1 : 0;
        return pair_fused_launch(a, nq, st);
    }
    GenArgs g{};
    g.q_patches = ext ? q_patches : ctx->patches;
    g.q_centers = ext ? q_centers : ctx->centers;
    g.q_rollout = ext ? q_rollout : ctx->rollout;
    g.c_patches = ctx->patches;
    g.c_centers = ctx->centers;
    g.c_rollout = ctx->rollout;
    g.cand_idx = cand_idx;
    g.cand_stride = cand_stride;
    g.q_start = ext ? 0 : q_start;
    g.q_stride = ext ? 1 : q_stride;
    g.nq = nq;
    g.k = k;
    g.c = ctx->c;
    g.r = ctx->r;
    g.p = *p;
    g.out_score = out_score;
    g.out_niter = out_niter;
    g.dbg_err = ctx->dbg_err;
    // operand copy of the registered bank for the tensor-core S3 of the generic path, derived once per registration
    g.packed = nullptr;
    if (!ext) {
        void* gp = generic_operand_copy(ctx);
        if (gp) {
            if (!ctx->gpacked_valid) {
                if ((rc = generic_repack(ctx->patches, ctx->centers, ctx->n, ctx->c, ctx->r, gp, st))) return rc;
                ctx->gpacked_valid = true;
            }
            g.packed = gp;
            g.packed_centers = ctx->centers && (ctx->r % 16) != 0;
        }
    }
    return generic_rerank(g, workspace, workspace_bytes, st);
}

int vr_rerank_scores(vr_ctx* ctx, int64_t q_start, int64_t q_stride, int64_t nq, int32_t k, const int32_t* cand_idx,
                     int32_t cand_stride, const vr_ot_params* p, float* out_score, int32_t* out_niter,
                     void* workspace, size_t workspace_bytes, void* stream) {
    return rerank_scores_impl(ctx, nullptr, nullptr, nullptr, q_start, q_stride, nq, k, cand_idx, cand_stride, p, out_score,
                              out_niter, workspace, workspace_bytes, stream);
}

int vr_rerank_scores_queries(vr_ctx* ctx, const float* q_patches, const float* q_centers, const float* q_rollout, int64_t nq,
                             int32_t k, const int32_t* cand_idx, int32_t cand_stride, const vr_ot_params* p, float* out_score,
                             int32_t* out_niter, void* workspace, size_t workspace_bytes, void* stream) {
    VR_REQUIRE(q_patches, "rerank: query patches missing");
    VR_REQUIRE(((uintptr_t)q_patches & 15) == 0, "rerank: query banks must be 16-byte aligned");
    return rerank_scores_impl(ctx, q_patches, q_centers, q_rollout, 0, 1, nq, k, cand_idx, cand_stride, p, out_score, out_niter,
                              workspace, workspace_bytes, stream);
}

size_t vr_finalize_workspace_bytes(vr_ctx* ctx, int64_t nq, int32_t n_trunc) {
    (void)ctx;
    return finalize_workspace_bytes(nq, n_trunc);
}

int vr_finalize(vr_ctx* ctx, int64_t q_start, int64_t q_stride, int64_t nq, int32_t k, int32_t kp,
                const int32_t* approx_idx, const float* approx_score, const float* ot_score,
                const int32_t* trunc_nums_host, int32_t n_trunc, int32_t* out_rank, double* tallies, void* workspace,
                size_t workspace_bytes, void* stream) {
    VR_REQUIRE(ctx, "ctx is null");
    if (!ctx->labels || !ctx->num_pos) {
        set_error("finalize: labels / num_pos not registered");
        return VR_E_NOBANK;
    }
    VR_REQUIRE(approx_idx && approx_score && trunc_nums_host && tallies, "finalize: null argument");
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    return finalize(q_start, q_stride, nq, k, kp, approx_idx, approx_score, ot_score, ctx->labels, ctx->num_pos,
                    trunc_nums_host, n_trunc, out_rank, tallies, workspace, workspace_bytes, (cudaStream_t)stream);
}

int vr_blend_rank(vr_ctx* ctx, int64_t nq, int32_t k, int32_t kp, const int32_t* approx_idx, const float* approx_score,
                  const float* ot_score, int32_t* out_rank, void* stream) {
    VR_REQUIRE(ctx, "ctx is null");
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    return blend_rank(nq, k, kp, approx_idx, approx_score, ot_score, out_rank, (cudaStream_t)stream);
}

size_t vr_rollout_block_workspace_bytes(int64_t b, int32_t ht, int32_t wt, int32_t drop_cls) {
    if (b <= 0 || ht <= 0 || wt <= 0) return 0;
    return rollout_block_workspace_bytes(b, ht, wt, drop_cls ? 1 : 0);
}

int vr_rollout_block(vr_ctx* ctx, const float* probs, int64_t b, int32_t heads, int32_t ht, int32_t wt, int32_t drop_cls, int32_t grid,
                     int64_t n_discard, int32_t fusion, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    VR_REQUIRE(ctx, "ctx is null");
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    return rollout_block(probs, b, heads, ht, wt, drop_cls, grid, n_discard, fusion, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int vr_rollout_chain(vr_ctx* ctx, const float* mats, int32_t n_mats, int64_t b, int32_t n, int32_t use_res, float* joints, void* stream) {
    VR_REQUIRE(ctx, "ctx is null");
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    return rollout_chain(mats, n_mats, b, n, use_res, joints, (cudaStream_t)stream);
}

size_t vr_sinkhorn_workspace_bytes(int64_t b, int32_t m, int32_t n) { return generic_sinkhorn_workspace_bytes(b, m, n); }

int vr_sinkhorn(const float* K, const float* u, const float* v, int64_t b, int32_t m, int32_t n, int32_t max_iter,
                float thresh, float* T, int32_t* niter, void* workspace, size_t workspace_bytes, void* stream) {
    VR_REQUIRE(K && u && v && T, "sinkhorn: null tensor");
    return generic_sinkhorn(K, u, v, b, m, n, max_iter, thresh, T, niter, workspace, workspace_bytes,
                            (cudaStream_t)stream);
}

size_t vr_calc_similarity_workspace_bytes(int64_t n, int32_t c, int32_t r, const vr_ot_params* p) {
    if (!p || n <= 0 || n >= 0x7fffffff) return 0;
    (void)c;
    // always the generic solver's size: the fused kernel needs none, but a misaligned view of a shape it supports is
    // routed to the generic solver by vr_calc_similarity
    return generic_rerank_workspace_bytes(1, (int)n, r, p);
}

int vr_calc_similarity(vr_ctx* ctx, const float* anchor, const float* anchor_center, const float* q_rollout,
                       const float* fb, const float* fb_center, const float* c_rollout, int64_t n, int32_t c,
                       int32_t r, const vr_ot_params* p, float* score, float* u, float* v, float* T, float* sim_r,
                       float* cc, int32_t* niter, void* workspace, size_t workspace_bytes, void* stream) {
    VR_REQUIRE(ctx, "ctx is null");
    int rc = check_params(p);
    if (rc) return rc;
    VR_REQUIRE(anchor && fb && score, "calc_similarity: anchor, fb and score are required");
    VR_REQUIRE(n > 0 && n < 0x7fffffff && c > 0 && r > 0, "calc_similarity: bad shape");
    VR_REQUIRE((u == nullptr) == (v == nullptr), "calc_similarity: u and v go together");
    const bool need_cc = p->mode >= VR_MODE_INVERSE;
    VR_REQUIRE(!(need_cc && p->use_cls_token) || (anchor_center && fb_center),
               "calc_similarity: centres required with use_cls_token");
    VR_REQUIRE(p->mode != VR_MODE_ROLLOUT || (q_rollout && c_rollout), "calc_similarity: rollout marginals missing");
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (pair_fused_supports(c, r, (int)n, p, false) && ((uintptr_t)anchor & 15) == 0 && ((uintptr_t)fb & 15) == 0) {
        PairArgs a{};
        a.q_patches = anchor;
        a.q_centers = anchor_center;
        a.q_rollout = q_rollout;
        a.c_patches = fb;
        a.c_centers = fb_center;
        a.c_rollout = c_rollout;
        a.cand_idx = nullptr;
        a.cand_stride = (int)n;
        a.q_start = 0;
        a.q_stride = 0;
        a.k = (int)n;
        a.p = *p;
        a.out_score = score;
        a.out_niter = niter;
        a.out_u = u;
        a.out_v = v;
        a.out_T = T;
        a.out_simr = sim_r;
        a.out_cc = cc;
        a.dbg_err = ctx->dbg_err;
        return pair_fused_launch(a, 1, st);
    }
    GenArgs g{};
    g.q_patches = anchor;
    g.q_centers = anchor_center;
    g.q_rollout = q_rollout;
    g.c_patches = fb;
    g.c_centers = fb_center;
    g.c_rollout = c_rollout;
    g.cand_idx = nullptr;
    g.cand_stride = (int)n;
    g.q_start = 0;
    g.q_stride = 0;
    g.nq = 1;
    g.k = (int)n;
    g.c = c;
    g.r = r;
    g.p = *p;
    g.out_score = score;
    g.out_niter = niter;
    g.out_u = u;
    g.out_v = v;
    g.out_T = T;
    g.out_simr = sim_r;
    g.out_cc = cc;
    g.dbg_err = ctx->dbg_err;
    return generic_rerank(g, workspace, workspace_bytes, st);
}

int vr_global_similarity(const float* q_center, const float* centers, int64_t n, int32_t c, float* sim, void* stream) {
    VR_REQUIRE(q_center && centers && sim, "global_similarity: null tensor");
    return global_similarity(q_center, centers, n, c, sim, (cudaStream_t)stream);
}

int vr_metrics_rank(const int64_t* tops, int64_t n_tops, int64_t query_label, const int64_t* labels, int64_t n_labels,
                    double* out3, void* stream) {
    return metrics_rank(tops, n_tops, query_label, labels, n_labels, out3, (cudaStream_t)stream);
}

int vr_evaluate_registered(vr_ctx* ctx, int64_t q_start, int64_t q_stride, int64_t nq, const int32_t* trunc_nums_host,
                           int32_t n_trunc, int32_t max_num_pos, const vr_ot_params* p, double* tallies_host,
                           int32_t* per_query_niter_host, void* stream) {
    VR_REQUIRE(ctx, "ctx is null");
    if (!ctx->patches || !ctx->labels || !ctx->num_pos) {
        set_error("evaluate: bank with labels and num_pos must be registered");
        return VR_E_NOBANK;
    }
    int rc = check_params(p);
    if (rc) return rc;
    VR_REQUIRE(nq > 0 && trunc_nums_host && n_trunc >= 1 && n_trunc <= 16 && tallies_host, "evaluate: bad arguments");
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int kmax = 0;
    for (int i = 0; i < n_trunc; i++) kmax = std::max(kmax, (int)trunc_nums_host[i]);
    // eval_cvt_diml.py:332: top_inds = approx_tops[:max(trunc_nums)]; the gallery may be smaller
    const int k = (int)std::min<int64_t>(kmax, ctx->n);
    int kp = std::max(k, (int)std::min<int64_t>(max_num_pos, ctx->n));
    kp = std::max(kp, (int)std::min<int64_t>(8, ctx->n));
    std::vector<int32_t> truncs(n_trunc);
    for (int i = 0; i < n_trunc; i++) truncs[i] = std::min((int)trunc_nums_host[i], k);

    // chunk the queries so that per-chunk buffers stay bounded (shortlists of up to 1 GB per chunk: the SOP pass is ONE
    // chunk, so that the whole first stage runs while a host bank is still being uploaded)
    int64_t chunk = std::min<int64_t>(nq, std::max<int64_t>(16384, ((int64_t)1 << 30) / ((int64_t)kp * 12)));
    const bool fused = k > 0 && (pair_fused_supports(ctx->c, ctx->r, k, p, !ctx->dbg_err) ||
                                 (!ctx->dbg_err && pair_fused_supports_wide(ctx->c, ctx->r, k, p)));
    const bool gfused = k > 0 && !fused && generic_fused_ready(ctx, p);
    if (k > 0 && !fused) {
        size_t per_q = gfused ? generic_fused_workspace_bytes(1, k, ctx->r) : generic_rerank_workspace_bytes(1, k, ctx->r, p);
        chunk = std::max<int64_t>(1, std::min<int64_t>(chunk, (int64_t)((size_t)1536 * 1024 * 1024 / per_q)));
    }
    void *d_idx, *d_sc, *d_ot, *d_nit, *d_tal, *d_ws0, *d_ws1, *d_ws2;
    size_t ws0 = stage0_workspace_bytes(chunk, ctx->n, ctx->c, kp, ctx->sms);
    size_t ws1 = k > 0 ? (fused ? 256 : vr_rerank_workspace_bytes(ctx, chunk, k, p)) : 256;
    size_t ws2 = finalize_workspace_bytes(chunk, n_trunc);
    if ((rc = arena_get(ctx, "ev_idx", (size_t)chunk * kp * 4, &d_idx))) return rc;
    if ((rc = arena_get(ctx, "ev_sc", (size_t)chunk * kp * 4, &d_sc))) return rc;
    if ((rc = arena_get(ctx, "ev_ot", (size_t)chunk * std::max(k, 1) * 4, &d_ot))) return rc;
    if ((rc = arena_get(ctx, "ev_nit", (size_t)nq * 4, &d_nit))) return rc;
    if ((rc = arena_get(ctx, "ev_tal", (size_t)n_trunc * 8 * sizeof(double), &d_tal))) return rc;
    if ((rc = arena_get(ctx, "ev_ws0", ws0, &d_ws0))) return rc;
    if ((rc = arena_get(ctx, "ev_ws1", ws1, &d_ws1))) return rc;
    if ((rc = arena_get(ctx, "ev_ws2", ws2, &d_ws2))) return rc;
    VR_CHECK_CUDA(cudaMemsetAsync(d_tal, 0, (size_t)n_trunc * 8 * sizeof(double), st));
    VR_CHECK_CUDA(cudaMemsetAsync(d_nit, 0, (size_t)nq * 4, st));

    for (int64_t lo = 0; lo < nq; lo += chunk) {
        const int64_t cnt = std::min(chunk, nq - lo);
        const int64_t qs = q_start + lo * q_stride;
        rc = vr_stage0_topk(ctx, nullptr, nullptr, qs, q_stride, cnt, kp, (int32_t*)d_idx, (float*)d_sc, d_ws0, ws0, st);
        if (rc) return rc;
        if (k > 0) {
            rc = vr_rerank_scores(ctx, qs, q_stride, cnt, k, (const int32_t*)d_idx, kp, p, (float*)d_ot,
                                  (int32_t*)d_nit + lo, d_ws1, ws1, st);
            if (rc) return rc;
        }
        rc = vr_finalize(ctx, qs, q_stride, cnt, k, kp, (const int32_t*)d_idx, (const float*)d_sc,
                         k > 0 ? (const float*)d_ot : nullptr, truncs.data(), n_trunc, nullptr, (double*)d_tal, d_ws2,
                         ws2, st);
        if (rc) return rc;
    }
    VR_CHECK_CUDA(cudaMemcpyAsync(tallies_host, d_tal, (size_t)n_trunc * 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (per_query_niter_host)
        VR_CHECK_CUDA(cudaMemcpyAsync(per_query_niter_host, d_nit, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    VR_CHECK_CUDA(cudaStreamSynchronize(st));
    return VR_OK;
}

int vr_evaluate_host(vr_ctx* ctx, const float* patches_host, const float* centers_host, const float* rollout_host,
                     const int64_t* labels_host, int64_t n, int32_t c, int32_t r, int64_t q_start, int64_t q_stride,
                     int64_t nq, const int32_t* trunc_nums_host, int32_t n_trunc, const vr_ot_params* p,
                     double* tallies_host, int32_t* per_query_niter_host) {
    VR_REQUIRE(ctx, "ctx is null");
    VR_REQUIRE(patches_host && centers_host && labels_host, "evaluate_host: patches, centers and labels are required");
    VR_REQUIRE(n > 0 && c > 0 && r > 0, "evaluate_host: bad shape");
    VR_CHECK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->own_stream;
    int rc;
    void *d_p, *d_c, *d_r = nullptr, *d_l, *d_np;
    const size_t bp = (size_t)n * c * r * 4, bc = (size_t)n * c * 4, br = (size_t)n * r * 4;
    if ((rc = arena_get(ctx, "h_patches", bp, &d_p))) return rc;
    if ((rc = arena_get(ctx, "h_centers", bc, &d_c))) return rc;
    if (rollout_host && (rc = arena_get(ctx, "h_rollout", br, &d_r))) return rc;
    if ((rc = arena_get(ctx, "h_labels", (size_t)n * 8, &d_l))) return rc;
    if ((rc = arena_get(ctx, "h_numpos", (size_t)n * 4, &d_np))) return rc;
    // labels first: their class counts decide the shortlist length, and only the maximum comes back (4 bytes)
    VR_CHECK_CUDA(cudaMemcpyAsync(d_l, labels_host, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    VR_CHECK_CUDA(cudaMemcpyAsync(d_c, centers_host, bc, cudaMemcpyHostToDevice, st));
    if (rollout_host) VR_CHECK_CUDA(cudaMemcpyAsync(d_r, rollout_host, br, cudaMemcpyHostToDevice, st));
    // (the re-pack of a piece stores every image's centre in its operand copy: prep_stream needs the centres)
    VR_CHECK_CUDA(cudaEventRecord(ctx->small_ready, st));
    VR_CHECK_CUDA(cudaStreamWaitEvent(ctx->prep_stream, ctx->small_ready, 0));
    rc = vr_bank_register(ctx, (const float*)d_p, (const float*)d_c, (const float*)d_r, (const int64_t*)d_l,
                          (const int32_t*)d_np, n, c, r);
    if (rc) return rc;
    // The patch bank goes up on its own stream in pieces: stage 0 needs the centres only and runs meanwhile, and every
    // piece is re-packed into the operand layout of the fused kernel (vr_bank_prepare) while the next one is in flight.
    const int pieces = (int)std::min<int64_t>(8, std::max<int64_t>(1, n / 512));
    const int64_t per = (n + pieces - 1) / pieces;
    for (int i = 0; i < pieces; i++) {
        const int64_t lo = i * per, cnt = std::min(per, n - lo);
        if (cnt <= 0) break;
        VR_CHECK_CUDA(cudaMemcpyAsync((char*)d_p + (size_t)lo * c * r * 4, (const char*)patches_host + (size_t)lo * c * r * 4,
                                      (size_t)cnt * c * r * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
        VR_CHECK_CUDA(cudaEventRecord(ctx->piece_ready[i], ctx->copy_stream));
        VR_CHECK_CUDA(cudaStreamWaitEvent(ctx->prep_stream, ctx->piece_ready[i], 0));
        if ((rc = vr_bank_prepare(ctx, lo, cnt, ctx->prep_stream))) return rc;
    }
    VR_CHECK_CUDA(cudaEventRecord(ctx->patches_ready, ctx->prep_stream));
    // num_pos[i] = #{j : label[j] == label[i]} (metrics.py:34) on the device
    int32_t max_np = 1;
    if ((rc = vr_num_pos(ctx, (const int64_t*)d_l, n, (int32_t*)d_np, &max_np, st))) return rc;
    ctx->pending_wait = true;
    rc = vr_evaluate_registered(ctx, q_start, q_stride, nq, trunc_nums_host, n_trunc, max_np, p, tallies_host,
                                per_query_niter_host, st);
    if (ctx->pending_wait) {   // no rerank ran (K = 0): still do not return before the upload has finished
        ctx->pending_wait = false;
        VR_CHECK_CUDA(cudaStreamSynchronize(ctx->prep_stream));
    }
    return rc;
}

}  // extern "C"
