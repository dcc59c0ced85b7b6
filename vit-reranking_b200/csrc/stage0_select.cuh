// Streaming top-kp select of one score row by ONE warp (shared by stage0_topk.cu and the fallback of stage0_mma.cu).
#pragma once
#include "common.cuh"

namespace vr {

// srow[0 .. n) fp32 scores (readable up to the even row stride ld), self = index to mask with -100 (eval_cvt_diml.py:327) or -1,
// b = P-key buffer of this warp in shared memory (P a power of two >= kp + 64), thr = keys <= thr are known to be outside the
// best kp (0: nothing known).  Writes the kp best (index, score) in descending (score, then lower index) order; idx -1 pads.
__device__ __forceinline__ void warp_rowselect_stream(const float* __restrict__ srow, int64_t n, int64_t ld, long long self, int kp,
                                                      int P, unsigned long long* b, int lane, unsigned long long thr,
                                                      int32_t* __restrict__ out_idx, float* __restrict__ out_score) {
    int cnt = 0;
    const unsigned lt = (1u << lane) - 1u;
    // 256 scores per step, as four groups of 64 (the buffer check is per group); the next step's loads are in flight
    // while this one is filtered, so the L2 / HBM latency is paid once per row, not once per group
    auto load4 = [&](float2 (&d)[4], int64_t base) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t col0 = base + 64 * u + 2 * lane;
            d[u] = make_float2(0.f, 0.f);
            if (col0 < ld) d[u] = __ldcs(reinterpret_cast<const float2*>(srow + col0));   // ld is even: col0 + 1 < ld too
        }
    };
    float2 cur[4], nxt[4];
    load4(cur, 0);
    for (int64_t base = 0; base < n; base += 256) {
        if (base + 256 < n) load4(nxt, base + 256);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (base + 64 * u >= n) break;
            if (cnt > P - 64) {
                for (int e = cnt + lane; e < P; e += 32) b[e] = 0ull;
                warp_bitonic_sort_desc(b, P, lane);
                cnt = min(cnt, kp);
                if (cnt >= kp) thr = b[kp - 1];
                __syncwarp();
            }
            const int64_t col0 = base + 64 * u + 2 * lane;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int64_t col = col0 + t;
                float s = t ? cur[u].y : cur[u].x;
                if (col == self) s = -100.0f;
                const unsigned long long key = pack_key(s, (uint32_t)col);
                const bool take = col < n && key > thr;
                const unsigned m = __ballot_sync(0xffffffffu, take);
                if (take) b[cnt + __popc(m & lt)] = key;
                cnt += __popc(m);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) cur[u] = nxt[u];
    }
    for (int e = cnt + lane; e < P; e += 32) b[e] = 0ull;
    warp_bitonic_sort_desc(b, P, lane);
    for (int e = lane; e < kp; e += 32) {
        const unsigned long long key = b[e];
        out_idx[e] = key ? (int32_t)key_index(key) : -1;
        out_score[e] = key ? key_score(key) : 0.f;
    }
}

}  // namespace vr
