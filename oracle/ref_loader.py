"""The REAL reference functions of the path, loaded by file path (TEST INFRASTRUCTURE ONLY).

`load()` returns (diml_module, metrics_module) from the reference checkout (/root/reference in the build container)
or from oracle/_ref/reference (the copy made by oracle/make_ref.py, which travels to the GPU box), or None when
neither exists.  The modules are loaded by path under private names: `import evaluation.metrics` would pull faiss
through the package __init__, and the names `utilities` / `evaluation` belong to the drop-in.

`reference_loop` drives those functions exactly as evaluation/eval_cvt_diml.py:316-372,402-416 does, over pre-built
CPU banks: it is what `bench.py --impl reference` times (cpu_baseline.kind = "reference") and what the parity tests
compare the restatement (rerank_oracle.evaluate_banks) with on the GPU box.
"""
from __future__ import annotations

import importlib.util
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.environ.get("VITRERANK_REFERENCE", "/root/reference"), os.path.join(HERE, "_ref", "reference")]
_cache = {}


def root():
    for c in CANDIDATES:
        if c and os.path.exists(os.path.join(c, "utilities", "diml.py")) and \
                os.path.exists(os.path.join(c, "evaluation", "metrics.py")):
            return c
    return None


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    r = root()
    if r is None:
        return None
    if r not in _cache:
        _cache[r] = (_load("_vr_ref_diml", os.path.join(r, "utilities", "diml.py")),
                     _load("_vr_ref_metrics", os.path.join(r, "evaluation", "metrics.py")))
    return _cache[r]


def reference_loop(patches, centers, rollout, labels, trunc_nums, use_rollout=False, query_ids=None, **flags):
    """eval_cvt_diml.py:316-372 + :402-416 with the reference's own calc_similarity /
    calc_similarity_cvt_rollout / get_metrics_rank (CPU tensors).  flags: use_uniform, use_inverse, temperature,
    use_cls_token, use_minus, ot_part.  Returns the reference's dict plus per-query records."""
    D, M = load()
    n = patches.shape[0]
    qids = range(n) if query_ids is None else [int(q) for q in query_ids]
    sums = {t: [0.0, 0.0, 0.0] for t in trunc_nums}
    kmax = max(trunc_nums)
    per_q = []
    for idx in qids:
        anchor_center, anchor = centers[idx], patches[idx]
        approx_sim, _ = D.calc_similarity(None, anchor_center, None, centers, 0)          # :325
        approx_sim = approx_sim.clone()
        approx_sim[idx] = -100                                                            # :327
        approx_tops = torch.argsort(approx_sim, descending=True)                          # :329
        rec = {"q": idx, "metrics": {}}
        if kmax > 0:
            top_inds = approx_tops[:kmax]                                                 # :332
            if use_rollout:                                                               # :344-351
                sim, _ = D.calc_similarity_cvt_rollout(anchor_center, anchor, rollout[idx], centers[top_inds],
                                                       patches[top_inds], rollout[top_inds], stage=1,
                                                       use_uniform=flags.get("use_uniform", False), use_ot=True,
                                                       ot_part=flags.get("ot_part", 0.1))
            else:                                                                         # :335-343
                sim, _ = D.calc_similarity(anchor, anchor_center, patches[top_inds], centers[top_inds], stage=1,
                                           use_uniform=flags.get("use_uniform", False),
                                           use_inverse=flags.get("use_inverse", False),
                                           temperature=flags.get("temperature", 1.0),
                                           use_cls_token=flags.get("use_cls_token", False), ot_temp=0.05,
                                           use_minus=flags.get("use_minus", False), ot_part=flags.get("ot_part", 0.1))
            rank_in_tops = torch.argsort(sim + approx_sim[top_inds], descending=True)     # :357
            rec.update(top=top_inds.clone(), score=sim.clone())
        for t in trunc_nums:                                                              # :359-372
            if t == 0:
                final_tops = approx_tops
            else:
                final_tops = torch.cat([top_inds[rank_in_tops][:t], approx_tops[t:]], dim=0)
            r1, rp, mapr = M.get_metrics_rank(final_tops, labels[idx], labels)
            rec["metrics"][t] = (float(r1), float(rp), float(mapr))
            s = sums[t]
            s[0] += r1
            s[1] += rp
            s[2] += mapr
        per_q.append(rec)
    scale = float(n / 100)                                                                # :403-405
    return {"r1": [float(sums[t][0]) / scale for t in trunc_nums], "rp": [float(sums[t][1]) / scale for t in trunc_nums],
            "mapr": [float(sums[t][2]) / scale for t in trunc_nums], "per_query": per_q, "n_queries": len(qids)}
