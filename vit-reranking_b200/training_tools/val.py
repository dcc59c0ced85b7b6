"""The rerank core of the reference's training_tools/val.py::val (MSLS place recognition): query images against the
database images of one city, query != gallery.  val() itself (MSLS dataset classes, faiss index bookkeeping, recall@N
against GPS ground truth, [:68-157,:200-260]) is outside the rerank path; its per-city block [:159-199] is
`rerank_city` below, as one batched pass on the GPU:

    approx_sim = einsum('nc,mc->nm', q, db); preds = argsort(desc)[:, :max(n_values)]        [:173-175]
    per query: calc_similarity(anchor, anchor_center, feats, feats_center, stage=1, use_minus=True, use_cls_token=True,
               temperature=0.1, ot_temp=0.05, ot_part=1.0)                                  [:187-195]
               rank = argsort(sim + approx_sim[idx, top_ind], desc); final = cat(top_ind[rank], preds[idx][T:])   [:197-199]
"""
from __future__ import annotations

import torch

from vitrerank.engine import OTParams, RerankEngine


def rerank_city(q_global, db_global, q_dense, db_dense, trunc_nums=None, n_values=(1, 5, 10, 20, 50, 100), device=None,
                return_scores=False):
    """q_global [Q, C], db_global [M, C] (as val() holds them: NOT re-normalised), q_dense [Q, C, R], db_dense [M, C, R]
    (already L2-normalised per location, [:129-130]).  Returns (preds [Q, max(n_values)], final_tops [Q, max(n_values)] or
    None when max(trunc_nums) == 0) as int64 tensors on the device, like the reference's `preds` / `final_tops`."""
    trunc_nums = trunc_nums or [0, 100]
    eng = RerankEngine.get(device)
    dev = eng.device
    m = db_global.shape[0]
    npred = min(max(n_values), m)
    t = min(max(trunc_nums), npred)
    eng.register(db_dense, db_global, None, None)
    idx, approx = eng.stage0_topk(npred, q_centers=q_global)                     # no self mask: queries are not db items
    preds = idx.long()
    if t <= 0:
        return (preds, None, None) if return_scores else (preds, None)
    p = OTParams.from_flags(use_minus=True, use_cls_token=True, temperature=0.1, ot_temp=0.05, ot_part=1.0)
    score, _ = eng.rerank_scores_queries(q_dense, q_global, idx[:, :t].contiguous(), t, p)
    # rank = argsort(sim + approx_sim[idx, top_ind], descending); exact ties: lower position first (torch's order is unspecified)
    reranked = eng.blend_rank(idx, approx, score, t)
    final = torch.cat([reranked.long(), preds[:, t:]], dim=1)
    return (preds, final, score) if return_scores else (preds, final)
