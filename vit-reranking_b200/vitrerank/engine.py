"""Host side of the B200 rerank path: torch owns device memory and streams, every
computation goes through the C ABI of libvitrerank.so (no torch math on the path).

Mirrors the query loop of the reference (evaluation/eval_cvt_diml.py:308-416) as a batched
pipeline over the registered gallery:

    S1  stage0_topk      global cosine + self mask + top-K' select      (:325-332)
    S2-S5a rerank_scores gather + patch sim + Sinkhorn + score          (:334-351)
    S5b finalize         blend, re-sort, r1 / RP / MAP@R tallies        (:357-372)
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np
import torch

from . import _lib
from ._lib import OTParamsStruct, VR_MODE, check, lib

METRIC_COLS = ("r1", "rp", "mapr", "recall@1", "recall@2", "recall@4", "recall@8", "count")


@dataclasses.dataclass
class OTParams:
    mode: str = "rollout"          # rollout | uniform | inverse | minus | soft | relu
    use_cls_token: bool = False
    ot_temp: float = 0.05          # eval_cvt_diml.py:341 / diml.py:325
    temperature: float = 1.0       # diml.py:77 default; --temperature 0.1 in the north-star run
    ot_part: float = 1.0
    max_iter: int = 100            # diml.py:42
    thresh: float = 1e-1           # diml.py:45

    def struct(self) -> OTParamsStruct:
        return OTParamsStruct(VR_MODE[self.mode], int(self.use_cls_token), float(self.ot_temp),
                              float(self.temperature), float(self.ot_part), int(self.max_iter),
                              float(self.thresh))

    @staticmethod
    def from_flags(use_rollout=False, use_uniform=False, use_inverse=False, use_minus=False, use_soft=False,
                   temperature=1.0, use_cls_token=False, ot_part=1.0, ot_temp=0.05) -> "OTParams":
        """Branch order of evaluate (eval_cvt_diml.py:201,334-351) and of calc_similarity
        (diml.py:80-81,104-133): with use_rollout the rollout branch runs and ignores the
        cross-correlation flags."""
        if use_minus:
            use_inverse = False
        if use_uniform:
            mode = "uniform"
        elif use_rollout:
            mode = "rollout"
        elif use_inverse:
            mode = "inverse"
        elif use_minus:
            mode = "minus"
        elif use_soft:
            mode = "soft"
        else:
            mode = "relu"
        return OTParams(mode=mode, use_cls_token=use_cls_token, ot_temp=ot_temp, temperature=temperature,
                        ot_part=ot_part)


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32(t, device):
    if t is None:
        return None
    return t.to(device=device, dtype=torch.float32).contiguous()


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.VitRerankError("vitrerank needs a CUDA device (B200, sm_100a); there is no CPU path")
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise _lib.VitRerankError(f"vitrerank runs on CUDA devices only, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class RerankEngine:
    """One context per CUDA device."""

    _instances: dict = {}

    def __init__(self, device=None):
        self.device = require_cuda(device)
        h = C.c_void_p()
        check(lib.vr_create(self.device.index, C.byref(h)), "vr_create")
        self._h = h
        self._bank = None
        self._ws = {}
        sm, mc = C.c_int32(), C.c_int32()
        check(lib.vr_device_info(self._h, C.byref(sm), C.byref(mc)))
        self.sm_count, self.max_active_clusters = sm.value, mc.value

    @classmethod
    def get(cls, device=None) -> "RerankEngine":
        dev = require_cuda(device)
        if dev.index not in cls._instances:
            cls._instances[dev.index] = cls(dev)
        return cls._instances[dev.index]

    def close(self):
        if getattr(self, "_h", None):
            lib.vr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def err_trace(self, nq, max_iter=100):
        """Diagnostics: returns a [nq, max_iter] device tensor that the next rerank_scores /
        calc_similarity calls fill with the stop-test value of every iteration (NaN = not run).
        Call err_trace(0) to switch it off."""
        if nq <= 0:
            self._trace = None
            check(lib.vr_debug_err_trace(self._h, C.c_void_p(0)))
            return None
        self._trace = torch.full((nq, max_iter), float('nan'), dtype=torch.float32, device=self.device)
        check(lib.vr_debug_err_trace(self._h, _ptr(self._trace)))
        return self._trace

    # ---- gallery -----------------------------------------------------------------------
    def num_pos(self, labels):
        """(num_pos [N] int32 on the device, its maximum): num_pos[i] = #{j: labels[j] == labels[i]}, the
        `torch.sum(gallery_label == query_label)` of metrics.py:34 for every item (counts the item itself)."""
        labels = labels.to(device=self.device, dtype=torch.int64).contiguous()
        out = torch.empty(labels.numel(), dtype=torch.int32, device=self.device)
        mx = C.c_int32(0)
        check(lib.vr_num_pos(self._h, _ptr(labels), labels.numel(), _ptr(out), C.byref(mx), _stream(self.device)),
              "vr_num_pos")
        return out, int(mx.value)

    def register(self, patches, centers, rollout=None, labels=None, num_pos=None, max_num_pos=None):
        """patches [N,C,R], centers [N,C], rollout [N,R], labels [N]; moved to the device if
        needed (the reference keeps the patch bank on the host, eval_cvt_diml.py:278).  num_pos / max_num_pos
        (from a previous num_pos() call on the same labels) skip the label count."""
        dev = self.device
        patches, centers, rollout = _f32(patches, dev), _f32(centers, dev), _f32(rollout, dev)
        n, c, r = patches.shape
        max_np = 1
        if labels is not None:
            labels = labels.to(device=dev, dtype=torch.int64).contiguous()
            if num_pos is None or max_num_pos is None:
                num_pos, max_np = self.num_pos(labels)
            else:
                num_pos, max_np = num_pos.to(device=dev, dtype=torch.int32).contiguous(), int(max_num_pos)
        else:
            num_pos = None
        check(lib.vr_bank_register(self._h, _ptr(patches), _ptr(centers), _ptr(rollout), _ptr(labels),
                                   _ptr(num_pos), n, c, r), "vr_bank_register")
        self._bank = dict(patches=patches, centers=centers, rollout=rollout, labels=labels, num_pos=num_pos,
                          n=n, c=c, r=r, max_num_pos=max_np)
        return self

    def prepare_bank(self, first=0, count=None, stream=None):
        """Derive the library's operand copy of images [first, first + count) now (vr_bank_prepare), on `stream`
        (a torch.cuda.Stream; default: the current one).  Optional: the first fused rerank does it otherwise."""
        b = self.bank
        count = b["n"] - first if count is None else count
        sp = C.c_void_p(stream.cuda_stream) if stream is not None else _stream(self.device)
        check(lib.vr_bank_prepare(self._h, first, count, sp), "vr_bank_prepare")

    def new_bank(self, n, c, grid, with_rollout=False):
        """Allocate and register empty banks for n images (patches [n, c, grid*grid], centres [n, c], rollout [n, grid*grid]):
        the destination of ingest().  Labels are attached later with register_labels()."""
        dev = self.device
        r = grid * grid
        patches = torch.empty(n, c, r, dtype=torch.float32, device=dev)
        centers = torch.empty(n, c, dtype=torch.float32, device=dev)
        rollout = torch.empty(n, r, dtype=torch.float32, device=dev) if with_rollout else None
        check(lib.vr_bank_register(self._h, _ptr(patches), _ptr(centers), _ptr(rollout), _ptr(None), _ptr(None), n, c, r),
              "vr_bank_register")
        self._bank = dict(patches=patches, centers=centers, rollout=rollout, labels=None, num_pos=None, n=n, c=c, r=r,
                          max_num_pos=1)
        return self

    def ingest(self, tokens, centers_raw, first, h=None, w=None, channel_major=False):
        """One batch of backbone outputs into rows [first, first + B) of the registered banks (vr_bank_ingest): tokens
        [B, L, C] (the head projection's output; channel_major: [B, C, L]), centers_raw [B, C] (or None).  Pools the h x w token
        map to the bank's grid, normalises per patch / per centre and writes the operand copy of the fused kernel."""
        b = self.bank
        tokens = _f32(tokens, self.device)
        centers_raw = _f32(centers_raw, self.device)
        cnt = tokens.shape[0]
        L = tokens.shape[2] if channel_major else tokens.shape[1]
        if h is None:
            h = w = int(round(L ** 0.5))
        assert h * w == L and (tokens.shape[1] if channel_major else tokens.shape[2]) == b["c"]
        check(lib.vr_bank_ingest(self._h, _ptr(tokens), _ptr(centers_raw), int(channel_major), first, cnt, h, w,
                                 _stream(self.device)), "vr_bank_ingest")

    def register_labels(self, labels):
        """Attach labels (and their class counts) to banks filled by ingest(); the operand copy is kept."""
        b = self.bank
        labels = labels.to(device=self.device, dtype=torch.int64).contiguous()
        num_pos, max_np = self.num_pos(labels)
        check(lib.vr_bank_labels(self._h, _ptr(labels), _ptr(num_pos)), "vr_bank_labels")
        b.update(labels=labels, num_pos=num_pos, max_num_pos=max_np)
        return self

    @property
    def bank(self):
        if self._bank is None:
            raise _lib.VitRerankError("no gallery registered")
        return self._bank

    def _workspace(self, name, nbytes):
        t = self._ws.get(name)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._ws[name] = t
        return t

    # ---- S1 ------------------------------------------------------------------------------
    def stage0_topk(self, kp, q_start=0, q_stride=1, nq=None, q_centers=None, self_idx=None):
        b = self.bank
        if q_centers is not None:
            q_centers = _f32(q_centers, self.device)
            nq = q_centers.shape[0]
            if self_idx is not None:
                self_idx = self_idx.to(device=self.device, dtype=torch.int64).contiguous()
        elif nq is None:
            nq = (b["n"] - q_start + q_stride - 1) // q_stride
        idx = torch.empty(nq, kp, dtype=torch.int32, device=self.device)
        score = torch.empty(nq, kp, dtype=torch.float32, device=self.device)
        nb = lib.vr_stage0_workspace_bytes(self._h, nq, kp)
        ws = self._workspace("s0", nb)
        check(lib.vr_stage0_topk(self._h, _ptr(q_centers), _ptr(self_idx), q_start, q_stride, nq, kp, _ptr(idx),
                                 _ptr(score), _ptr(ws), ws.numel(), _stream(self.device)), "vr_stage0_topk")
        return idx, score

    def stage0_stats(self):
        """Counters of the last stage0_topk call (synchronises): rows the tensor-core path handed to its exact fp32
        fallback, rows whose candidate buffer overflowed, whether the centres held values the fp16 split cannot hold."""
        out = (C.c_uint32 * 4)()
        check(lib.vr_stage0_stats(self._h, out, _stream(self.device)), "vr_stage0_stats")
        return {"fallback_rows": int(out[0]), "overflow_rows": int(out[3]), "unsplittable": int(out[2])}

    # ---- S2-S5a ----------------------------------------------------------------------------
    def rerank_scores(self, cand_idx, k, params: OTParams, q_start=0, q_stride=1, workspace_bytes=None):
        """workspace_bytes: a smaller workspace than vr_rerank_workspace_bytes asks for (the shape-generic path then runs as
        many queries per round as fit)."""
        cand_idx = cand_idx.to(device=self.device, dtype=torch.int32).contiguous()
        nq, stride = cand_idx.shape
        ps = params.struct()
        score = torch.empty(nq, k, dtype=torch.float32, device=self.device)
        niter = torch.zeros(nq, dtype=torch.int32, device=self.device)
        nb = lib.vr_rerank_workspace_bytes(self._h, nq, k, C.byref(ps)) if workspace_bytes is None else int(workspace_bytes)
        ws = self._workspace("s1", nb) if workspace_bytes is None else torch.empty(nb, dtype=torch.uint8, device=self.device)
        check(lib.vr_rerank_scores(self._h, q_start, q_stride, nq, k, _ptr(cand_idx), stride, C.byref(ps),
                                   _ptr(score), _ptr(niter), _ptr(ws), ws.numel(), _stream(self.device)),
              "vr_rerank_scores")
        return score, niter

    def rerank_scores_queries(self, q_patches, q_centers, cand_idx, k, params: OTParams, q_rollout=None):
        """rerank_scores for queries that are not gallery items (training_tools/val.py:159-190): q_patches [nq, C, R],
        q_centers [nq, C]; candidates from the registered bank."""
        q_patches, q_centers, q_rollout = _f32(q_patches, self.device), _f32(q_centers, self.device), _f32(q_rollout, self.device)
        cand_idx = cand_idx.to(device=self.device, dtype=torch.int32).contiguous()
        nq, stride = cand_idx.shape
        assert q_patches.shape[0] == nq
        ps = params.struct()
        score = torch.empty(nq, k, dtype=torch.float32, device=self.device)
        niter = torch.zeros(nq, dtype=torch.int32, device=self.device)
        nb = lib.vr_rerank_workspace_bytes(self._h, nq, k, C.byref(ps))
        # (for a bank whose own queries take the fused generic kernel that is the small figure; queries from outside take the
        # separate kernels, which run as many queries per round as the workspace holds: give them up to 2 GB)
        re = self._bank["r"] + (0 if params.ot_part > 0.999 else 1)
        nb = max(nb, min(nq * k * (2 * re * re + 20 * re + 64) * 4 + 65536, 2 << 30))
        ws = self._workspace("s1", nb)
        check(lib.vr_rerank_scores_queries(self._h, _ptr(q_patches), _ptr(q_centers), _ptr(q_rollout), nq, k, _ptr(cand_idx),
                                           stride, C.byref(ps), _ptr(score), _ptr(niter), _ptr(ws), ws.numel(),
                                           _stream(self.device)), "vr_rerank_scores_queries")
        return score, niter

    # ---- S5b ---------------------------------------------------------------------------------
    def finalize(self, approx_idx, approx_score, ot_score, k, trunc_nums, q_start=0, q_stride=1, tallies=None,
                 want_rank=False, want_per_query=False):
        """Returns (tallies, rank) or, with want_per_query, (tallies, rank, per_query [nq, len(trunc_nums), 8] float64:
        the r1 / rp / mapr / recall@1,2,4,8 / 1 of every query, columns METRIC_COLS)."""
        nq, kp = approx_idx.shape
        nt = len(trunc_nums)
        if tallies is None:
            tallies = torch.zeros(nt, 8, dtype=torch.float64, device=self.device)
        rank = torch.empty(nq, k, dtype=torch.int32, device=self.device) if (want_rank and k > 0) else None
        tr = (C.c_int32 * nt)(*[int(t) for t in trunc_nums])
        nb = lib.vr_finalize_workspace_bytes(self._h, nq, nt)
        ws = self._workspace("s2", nb)
        check(lib.vr_finalize(self._h, q_start, q_stride, nq, k, kp, _ptr(approx_idx), _ptr(approx_score),
                              _ptr(ot_score), tr, nt, _ptr(rank), _ptr(tallies), _ptr(ws), ws.numel(),
                              _stream(self.device)), "vr_finalize")
        if want_per_query:   # the head of the workspace is the per-query table the tallies were summed from
            pq = ws[:nq * nt * 8 * 8].view(torch.float64).view(nq, nt, 8).clone()
            return tallies, rank, pq
        return tallies, rank

    def blend_rank(self, approx_idx, approx_score, ot_score, k):
        """approx_idx[q, argsort(ot_score[q] + approx_score[q, :k], descending)] (vr_blend_rank): [nq, k] int32."""
        nq, kp = approx_idx.shape
        rank = torch.empty(nq, k, dtype=torch.int32, device=self.device)
        check(lib.vr_blend_rank(self._h, nq, k, kp, _ptr(approx_idx), _ptr(approx_score), _ptr(ot_score), _ptr(rank),
                                _stream(self.device)), "vr_blend_rank")
        return rank

    # ---- attention-rollout producer (eval_cvt_diml.py:54-146) ----------------------------------------
    def rollout_block(self, probs, drop_cls, grid=7, discard_ratio=0.1, head_fusion="min"):
        """filter_attention_map + resize_attn_map of one block (vr_rollout_block): probs [B, heads, T, T'] fp32 on the
        engine's device -> [B, grid^2, grid^2]."""
        assert probs.is_cuda and probs.dtype == torch.float32 and probs.dim() == 4
        probs = probs.contiguous()
        b, heads, ht, wt = probs.shape
        d = 1 if drop_cls else 0
        n_discard = int(ht * wt * discard_ratio)                        # eval_cvt_diml.py:91 (the map still holds the cls row / column)
        g2 = grid * grid
        out = torch.empty(b, g2, g2, dtype=torch.float32, device=self.device)
        nb = lib.vr_rollout_block_workspace_bytes(b, ht, wt, d)
        ws = self._workspace("ro", nb)
        check(lib.vr_rollout_block(self._h, _ptr(probs), b, heads, ht, wt, d, grid, n_discard, {"max": 1, "min": 2}[head_fusion],
                                   _ptr(out), _ptr(ws), ws.numel(), _stream(self.device)), "vr_rollout_block")
        return out

    def rollout_chain(self, mats, use_res=True):
        """mats [J, B, n, n] -> joints [J, B, n, n] (vr_rollout_chain): identity + row normalisation when use_res, then the
        running product joint[j] = mats[j] @ joint[j - 1]."""
        mats = _f32(mats, self.device)
        j, b, n, n2 = mats.shape
        assert n == n2
        joints = torch.empty_like(mats)
        check(lib.vr_rollout_chain(self._h, _ptr(mats), j, b, n, int(bool(use_res)), _ptr(joints), _stream(self.device)),
              "vr_rollout_chain")
        return joints

    # ---- whole pass over the registered (device-resident) gallery ---------------------------------
    def evaluate(self, trunc_nums, params: OTParams, q_start=0, q_stride=1, nq=None, want_niter=False):
        """Tallies [len(trunc_nums), 8] (numpy float64, columns METRIC_COLS; sums, not yet
        divided by N/100) for queries q_start + i*q_stride."""
        b = self.bank
        if nq is None:
            nq = (b["n"] - q_start + q_stride - 1) // q_stride
        nt = len(trunc_nums)
        tr = (C.c_int32 * nt)(*[int(t) for t in trunc_nums])
        out = np.zeros((nt, 8), dtype=np.float64)
        nit = np.zeros(nq, dtype=np.int32) if want_niter else None
        ps = params.struct()
        check(lib.vr_evaluate_registered(self._h, q_start, q_stride, nq, tr, nt, b["max_num_pos"], C.byref(ps),
                                         out.ctypes.data_as(C.POINTER(C.c_double)),
                                         C.c_void_p(nit.ctypes.data if nit is not None else 0),
                                         _stream(self.device)), "vr_evaluate_registered")
        return (out, nit) if want_niter else out

    # ---- whole pass from HOST buffers (end-to-end entry) ---------------------------------------
    def evaluate_host(self, patches, centers, rollout, labels, trunc_nums, params: OTParams, q_start=0,
                      q_stride=1, nq=None, want_niter=False):
        """Same as evaluate() but the banks are CPU tensors (pinned for full PCIe speed), as the
        reference keeps them; host->device copies and the device->host read of the tallies are
        inside the call."""
        for t in (patches, centers, labels):
            if t.device.type != "cpu":
                raise _lib.VitRerankError("evaluate_host takes CPU tensors")
        patches = patches.contiguous().float()
        centers = centers.contiguous().float()
        rollout = None if rollout is None else rollout.contiguous().float()
        labels = labels.contiguous().long()
        n, c, r = patches.shape
        if nq is None:
            nq = (n - q_start + q_stride - 1) // q_stride
        nt = len(trunc_nums)
        tr = (C.c_int32 * nt)(*[int(t) for t in trunc_nums])
        out = np.zeros((nt, 8), dtype=np.float64)
        nit = np.zeros(nq, dtype=np.int32) if want_niter else None
        ps = params.struct()
        check(lib.vr_evaluate_host(self._h, _ptr(patches), _ptr(centers), _ptr(rollout), _ptr(labels), n, c, r,
                                   q_start, q_stride, nq, tr, nt, C.byref(ps),
                                   out.ctypes.data_as(C.POINTER(C.c_double)),
                                   C.c_void_p(nit.ctypes.data if nit is not None else 0)), "vr_evaluate_host")
        self._bank = None  # the context now points at its own copy of the banks
        return (out, nit) if want_niter else out

    # ---- direct calls (utilities/diml.py surface) ---------------------------------------------------
    def sinkhorn(self, K, u, v, max_iter=100, thresh=1e-1):
        K, u, v = _f32(K, self.device), _f32(u, self.device), _f32(v, self.device)
        b, m, n = K.shape
        T = torch.empty_like(K)
        niter = torch.zeros(1, dtype=torch.int32, device=self.device)
        nb = lib.vr_sinkhorn_workspace_bytes(b, m, n)
        ws = self._workspace("sk", nb)
        check(lib.vr_sinkhorn(_ptr(K), _ptr(u), _ptr(v), b, m, n, int(max_iter), float(thresh), _ptr(T),
                              _ptr(niter), _ptr(ws), ws.numel(), _stream(self.device)), "vr_sinkhorn")
        return T, niter

    def global_similarity(self, q_center, centers):
        q_center, centers = _f32(q_center, self.device), _f32(centers, self.device)
        n, c = centers.shape
        sim = torch.empty(n, dtype=torch.float32, device=self.device)
        check(lib.vr_global_similarity(_ptr(q_center), _ptr(centers), n, c, _ptr(sim), _stream(self.device)),
              "vr_global_similarity")
        return sim

    def calc_similarity(self, anchor, anchor_center, fb, fb_center, params: OTParams, q_rollout=None,
                        c_rollout=None, want_uv=True):
        dev = self.device
        anchor, fb = _f32(anchor, dev), _f32(fb, dev)
        anchor_center, fb_center = _f32(anchor_center, dev), _f32(fb_center, dev)
        q_rollout, c_rollout = _f32(q_rollout, dev), _f32(c_rollout, dev)
        n, c, r = fb.shape
        assert anchor.shape == (c, r), f"anchor {tuple(anchor.shape)} vs fb {tuple(fb.shape)}"
        re = r if params.ot_part > 0.999 else r + 1
        score = torch.empty(n, dtype=torch.float32, device=dev)
        niter = torch.zeros(1, dtype=torch.int32, device=dev)
        u = v = T = sim_r = cc = None
        if want_uv:
            u = torch.empty(n, r, dtype=torch.float32, device=dev)
            v = torch.empty(n, r, dtype=torch.float32, device=dev)
            T = torch.empty(n, re, re, dtype=torch.float32, device=dev)
            sim_r = torch.empty(n, r, r, dtype=torch.float32, device=dev)
            if params.mode in ("minus", "soft", "relu"):
                cc = torch.empty(n, r, dtype=torch.float32, device=dev)
        ps = params.struct()
        nb = lib.vr_calc_similarity_workspace_bytes(n, c, r, C.byref(ps))
        ws = self._workspace("cs", nb)
        check(lib.vr_calc_similarity(self._h, _ptr(anchor), _ptr(anchor_center), _ptr(q_rollout), _ptr(fb),
                                     _ptr(fb_center), _ptr(c_rollout), n, c, r, C.byref(ps), _ptr(score), _ptr(u),
                                     _ptr(v), _ptr(T), _ptr(sim_r), _ptr(cc), _ptr(niter), _ptr(ws), ws.numel(),
                                     _stream(dev)), "vr_calc_similarity")
        return score, (u, v, T, sim_r, cc), niter
