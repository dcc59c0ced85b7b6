"""Parity counters: CUDA outputs against the UN-FORCED oracle (TEST INFRASTRUCTURE ONLY).

The oracle (rerank_oracle.evaluate_banks, dump=True) runs the reference's loop as the reference
does -- its own stop test decides the Sinkhorn iteration count of every query.  Nothing here
re-runs it at the CUDA path's iteration counts: a query whose count differs is COUNTED (with the
distance of the oracle's err from the 0.1 threshold at the decisive iteration), and its pairs
stay in the score statistics.  Used by tests/test_gpu_fullpass.py and by bench.py's cpu_baseline
leg (the `parity` object of the JSON line).

Gates named by BASELINE.json north_star: first-stage top-K sets bit-exact, per-pair OT scores
within 1e-4 relative, Recall / MAP@R identical.
"""
from __future__ import annotations

import numpy as np

SCORE_RTOL = 1e-4
TIE = 1e-6


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-12)


def compare(dumps, idx, score, niter, k, trunc_nums=None, per_query=None, thresh=0.1):
    """dumps: the oracle's per-query records (same order as the rows of the CUDA arrays);
    idx [nq, >=k] first-stage shortlist, score [nq, k] OT scores in shortlist order, niter [nq];
    per_query (optional) [nq, len(trunc_nums), >=3] the CUDA path's per-query r1 / rp / mapr.
    Returns a dict of plain numbers."""
    idx = np.asarray(idx)
    score = np.asarray(score)
    niter = np.asarray(niter)
    nq = len(dumps)
    c = {"queries": nq, "k": int(k), "pairs": 0, "stage0_set_mismatch": 0, "stage0_set_mismatch_beyond_tie": 0,
         "stage0_boundary_near_ties": 0, "niter_equal": 0, "niter_off_by_one": 0, "niter_off_by_more": 0,
         "flips_outside_2pct_band": 0, "max_flip_err_distance": 0.0, "pairs_over_1e-4": 0,
         "pairs_over_1e-4_in_equal_niter_queries": 0, "max_rel_err": 0.0, "max_rel_err_equal_niter": 0.0,
         "queries_with_pairs_over_1e-4": 0, "metric_mismatch_queries": 0, "metric_mismatch_queries_equal_niter": 0,
         "stage0_order_differs_within_tie": 0, "stage0_order_differs_beyond_tie": 0,
         "metric_mismatch_unexplained": 0,
         "mean_niter_cuda": float(niter[:nq].mean()) if nq else 0.0,
         "mean_niter_oracle": float(np.mean([d["n_iter"] for d in dumps])) if nq else 0.0}
    dt = {t: [0.0, 0.0, 0.0] for t in (trunc_nums or [])}
    for q, d in enumerate(dumps):
        top = d["top"].numpy()
        kk = len(top)
        mine_set, ref_set = set(idx[q, :kk].tolist()), set(top.tolist())
        if d.get("gap", 1.0) < TIE:
            c["stage0_boundary_near_ties"] += 1
        comparable = mine_set == ref_set
        if not comparable:
            c["stage0_set_mismatch"] += 1
            if not d.get("gap", 1.0) < TIE:
                c["stage0_set_mismatch_beyond_tie"] += 1
        # order inside the shortlist: torch's matvec and the CUDA FMA chain round the global scores differently, so two
        # candidates whose scores agree to 1e-6 may swap places (which can move MAP@R of the un-reranked list, trunc 0)
        order_tie = False
        if comparable and not np.array_equal(idx[q, :kk], top):
            ref_sc = {int(cand): float(sc) for cand, sc in zip(top, d["approx"].numpy())}
            worst = max(abs(ref_sc[int(a)] - ref_sc[int(b)]) for a, b in zip(idx[q, :kk], top) if a != b)
            if worst < TIE:
                order_tie = True
                c["stage0_order_differs_within_tie"] += 1
            else:
                c["stage0_order_differs_beyond_tie"] += 1
        n_ref, n_mine = int(d["n_iter"]), int(niter[q])
        equal = n_ref == n_mine
        if equal:
            c["niter_equal"] += 1
        elif abs(n_ref - n_mine) == 1:
            c["niter_off_by_one"] += 1
            errs = d["errs"]
            e = errs[n_ref - 2] if n_mine < n_ref else errs[n_ref - 1]
            dist = abs(e - thresh) / thresh
            c["max_flip_err_distance"] = max(c["max_flip_err_distance"], float(dist))
            if dist > 0.02:
                c["flips_outside_2pct_band"] += 1
        else:
            c["niter_off_by_more"] += 1
        if comparable:
            pos = {int(cand): i for i, cand in enumerate(idx[q, :kk])}
            mine = np.array([score[q, pos[int(cand)]] for cand in top])
            rel = _rel(mine, d["score"].numpy())
            over = int((rel > SCORE_RTOL).sum())
            c["pairs"] += kk
            c["pairs_over_1e-4"] += over
            c["queries_with_pairs_over_1e-4"] += 1 if over else 0
            c["max_rel_err"] = max(c["max_rel_err"], float(rel.max()))
            if equal:
                c["pairs_over_1e-4_in_equal_niter_queries"] += over
                c["max_rel_err_equal_niter"] = max(c["max_rel_err_equal_niter"], float(rel.max()))
        if per_query is not None and trunc_nums:
            bad = False
            for ti, t in enumerate(trunc_nums):
                ref = d["metrics"][t]
                got = per_query[q, ti, :3]
                for j in range(3):
                    dt[t][j] += float(got[j]) - float(ref[j])
                    if float(got[j]) != float(ref[j]):
                        bad = True
            if bad:
                c["metric_mismatch_queries"] += 1
                if equal and comparable:
                    c["metric_mismatch_queries_equal_niter"] += 1
                    if not order_tie:     # same shortlist, same iteration count, same first-stage order: must be identical
                        c["metric_mismatch_unexplained"] += 1
    if dt:
        # sums over the sampled queries of (cuda - oracle), per trunc: r1, rp, mapr (not yet divided by N/100)
        c["tally_delta"] = {str(t): v for t, v in dt.items()}
    return c
