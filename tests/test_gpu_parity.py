"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded
inputs, and against the golden fixtures produced by the real reference.

Gates (BASELINE.json north_star): first-stage top-K index sets bit-exact (ties at the K/K+1
boundary closer than 1e-6 are reported, not hidden), per-pair OT scores within 1e-4
relative, Sinkhorn iteration counts equal, R@1 / RP / MAP@R identical.
"""
import os

import numpy as np
import pytest
import torch

from oracle import rerank_oracle as O
from vitrerank import synth

from golden.cases import CALC_CASES, LOOP_CASES

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-4    # north_star gate: per-pair OT scores within 1e-4 relative (queries with the reference's iteration count)
FLIP_RTOL = 1e-2     # a query one Sinkhorn iteration apart from the oracle is compared UN-FORCED; one more iteration of the
                     # reference itself moves its scores by up to 2e-3 (measured, profiles/r2_parity.md); the full-size tests
                     # (tests/test_gpu_fullpass.py) count such queries and pairs


def score_gate(n_mine, n_ref):
    return SCORE_RTOL if int(n_mine) == int(n_ref) else FLIP_RTOL


def plan_tol(n_mine, n_ref):
    """rtol for the transport plan T / sim_r against the un-forced oracle."""
    return 2e-4 if int(n_mine) == int(n_ref) else 5e-2


@pytest.fixture(scope="module")
def eng():
    from vitrerank.engine import RerankEngine
    return RerankEngine.get("cuda:0")


def params(**kw):
    from vitrerank.engine import OTParams
    return OTParams(**kw)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-12)


def stop_ok(n_mine, n_ref, errs, tol=0.02):
    """The reference's stop test sits at the fp32 noise floor (DESIGN.md "n* fragility": its own
    fp32 and fp64 runs disagree on ~15% of queries), so the iteration count may differ by ONE,
    and only when the oracle's err at the decisive iteration is within tol (relative) of 0.1."""
    if n_mine == n_ref:
        return True
    if abs(n_mine - n_ref) != 1:
        return False
    e = errs[n_ref - 2] if n_mine < n_ref else errs[n_ref - 1]
    return abs(e - 0.1) <= tol * 0.1


# ---------------------------------------------------------------------------------------------
def test_library_loads_on_gpu(eng):
    assert eng.sm_count > 0
    assert eng.max_active_clusters > 0


@pytest.mark.parametrize("n,c,kp", [(300, 128, 100), (1000, 128, 128), (700, 64, 40), (50, 128, 100),
                                    (2500, 128, 1000)])
def test_stage0_topk(eng, n, c, kp):
    g = synth.make_gallery(n, c, 4, classes=max(2, n // 30), seed=n + kp, sigma=0.6)
    eng.register(g.patches, g.centers, None, g.labels)
    idx, score = eng.stage0_topk(kp)
    idx, score = idx.cpu().numpy(), score.cpu().numpy()
    keff = min(kp, n)
    near_tie_rows = 0
    for q in range(n):
        sim = O.global_similarity(g.centers[q], g.centers).clone()
        sim[q] = O.SELF_MASK
        order = torch.argsort(sim, descending=True).numpy()
        ref = sim.numpy()
        got = idx[q, :keff]
        assert (idx[q, keff:] == -1).all()
        np.testing.assert_allclose(score[q, :keff], ref[got], rtol=0, atol=2e-6)
        assert (np.diff(score[q, :keff]) <= 0).all(), "shortlist not sorted"
        if set(got.tolist()) != set(order[:keff].tolist()):
            # only legal when the boundary is a numerical tie
            diff = set(got.tolist()) ^ set(order[:keff].tolist())
            kth = ref[order[keff - 1]]
            assert all(abs(ref[i] - kth) < 1e-6 for i in diff), f"query {q}: top-{keff} set differs beyond a tie"
            near_tie_rows += 1
    assert near_tie_rows <= max(1, n // 200)


def test_stage0_batch_and_small_paths_agree(eng):
    """>= 256 queries go through the SGEMM + row-select kernels, fewer through the fused select kernel:
    same fp32 FMA chains, same keys, so the shortlists must be bit-identical; the same holds for explicit
    query centres with a self index (the query != gallery form of S1) and for an interleaved shard."""
    n, kp = 777, 124
    g = synth.make_gallery(n, 128, 4, classes=25, seed=3, sigma=0.6)
    eng.register(g.patches, g.centers, None, g.labels)
    idx, score = eng.stage0_topk(kp)                       # batch path
    for lo in range(0, n, 200):                            # fused path, 200 (or 177) queries at a time
        cnt = min(200, n - lo)
        i2, s2 = eng.stage0_topk(kp, q_start=lo, nq=cnt)
        assert torch.equal(i2, idx[lo:lo + cnt]) and torch.equal(s2, score[lo:lo + cnt])
    i3, s3 = eng.stage0_topk(kp, q_centers=g.centers, self_idx=torch.arange(n))
    assert torch.equal(i3, idx) and torch.equal(s3, score)
    i4, s4 = eng.stage0_topk(kp, q_start=1, q_stride=2)    # 388 queries: batch path on a strided shard
    assert torch.equal(i4, idx[1::2]) and torch.equal(s4, score[1::2])
    i5, _ = eng.stage0_topk(kp, q_centers=g.centers[:300])  # no self index: nothing is masked
    assert (i5[:, 0].cpu() == torch.arange(300)).all()


def _oracle_pair(g, mode, **kw):
    return O.structural_similarity(g.patches[0], g.centers[0], g.patches[1:], g.centers[1:], mode,
                                   q_rollout=g.rollout[0], c_rollout=g.rollout[1:], trace=True, **kw)


PAIR_CASES = [
    ("rollout", dict(), 100, 0.6),
    ("rollout", dict(), 37, 1.0),
    ("rollout", dict(ot_part=0.3), 20, 0.6),
    ("uniform", dict(), 16, 0.6),
    ("inverse", dict(temperature=0.1, use_cls_token=True), 100, 0.6),
    ("inverse", dict(temperature=0.1, use_cls_token=False), 24, 0.8),
    ("minus", dict(use_cls_token=True), 24, 0.6),
    ("minus", dict(use_cls_token=True, ot_part=0.5), 24, 0.6),
    ("soft", dict(use_cls_token=True), 12, 0.6),
    ("relu", dict(use_cls_token=False), 12, 0.6),
    ("rollout", dict(), 104, 0.3),
    # shortlist lengths around the CTA / warp boundaries of the 7 x 16 pair slots
    ("rollout", dict(), 1, 0.6),
    ("rollout", dict(), 17, 0.6),
    ("rollout", dict(), 97, 0.6),
    ("rollout", dict(), 112, 0.6),
    ("uniform", dict(), 33, 1.0),
]


@pytest.mark.parametrize("mode,kw,k,sigma", PAIR_CASES)
def test_calc_similarity_fused(eng, mode, kw, k, sigma):
    g = synth.make_gallery(k + 1, 128, 49, classes=3, seed=1000 + k, sigma=sigma)
    ref_score, ref_uv, (n_ref, errs) = _oracle_pair(g, mode, **kw)
    p = params(mode=mode, **kw)
    score, uv, niter = eng.calc_similarity(g.patches[0], g.centers[0], g.patches[1:], g.centers[1:], p,
                                           q_rollout=g.rollout[0], c_rollout=g.rollout[1:])
    torch.cuda.synchronize()
    np.testing.assert_allclose(uv[0].cpu(), ref_uv[0], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(uv[1].cpu(), ref_uv[1], rtol=1e-5, atol=1e-7)
    if ref_uv[4] is not None:
        np.testing.assert_allclose(uv[4].cpu(), ref_uv[4], rtol=1e-5, atol=2e-6)
    assert stop_ok(int(niter), n_ref, errs), f"n* {int(niter)} vs oracle {n_ref}, errs tail {errs[-3:]}"
    # un-forced: the oracle keeps ITS iteration count
    assert rel_err(score.cpu(), ref_score).max() < score_gate(niter, n_ref)
    np.testing.assert_allclose(uv[2].cpu(), ref_uv[2], rtol=plan_tol(niter, n_ref), atol=1e-9)
    np.testing.assert_allclose(uv[3].cpu(), ref_uv[3], rtol=plan_tol(niter, n_ref), atol=1e-8)


@pytest.mark.parametrize("k", [100, 37, 130])
def test_stop_test_thresholds_and_iteration_caps(eng, k):
    """The stop test runs two iterations late on fixed-point sums whose format follows thresh * K * 49: thresholds over
    nine decades and iteration caps around the lag (1, 2, 3) must give the iteration count the oracle's err trace
    dictates, and the scores of exactly that many iterations."""
    g = synth.make_gallery(k + 1, 128, 49, classes=3, seed=4242 + k, sigma=0.6)
    _, _, (_, errs) = _oracle_pair(g, "rollout", force_iters=60)       # err of iterations 1..60, no stop

    def expected(thresh, max_iter):
        for t in range(min(max_iter, 60)):
            if errs[t] < thresh:
                return t + 1
        return max_iter

    cases = [(1e9, 100), (1e3, 100), (10.0, 100), (1.0, 100), (0.3, 100), (0.1, 100), (1e-2, 60), (0.0, 12),
             (0.0, 1), (0.0, 2), (0.0, 3), (1e9, 1), (1e9, 2), (10.0, 2), (10.0, 3)]
    for thresh, max_iter in cases:
        p = params(mode="rollout", thresh=thresh, max_iter=max_iter)
        score, uv, niter = eng.calc_similarity(g.patches[0], g.centers[0], g.patches[1:], g.centers[1:], p,
                                               q_rollout=g.rollout[0], c_rollout=g.rollout[1:])
        n_exp = expected(thresh, max_iter)
        n_got = int(niter)
        if n_got != n_exp:   # only legal one iteration off, at a near tie of err and thresh
            assert abs(n_got - n_exp) == 1, (thresh, max_iter, n_got, n_exp)
            e = errs[min(n_got, n_exp) - 1]
            assert abs(e - thresh) <= 0.02 * thresh, (thresh, max_iter, n_got, n_exp, e)
        # the reference has one threshold (0.1) and one cap (100); other values exist only through the oracle's
        # err trace, so its result is the run of exactly n_exp iterations -- the count ITS trace dictates
        ref_score, _, _ = _oracle_pair(g, "rollout", force_iters=n_exp)
        if n_got == n_exp:
            assert rel_err(score.cpu(), ref_score).max() < SCORE_RTOL, (thresh, max_iter, n_got, n_exp)
        # the score-only kernel (evaluate path) takes the same decision
        eng.register(g.patches, g.centers, g.rollout, g.labels)
        idx = torch.arange(1, k + 1, dtype=torch.int32, device="cuda")[None, :]
        s2, n2 = eng.rerank_scores(idx, k, p, q_start=0, q_stride=1)
        assert int(n2[0]) == n_got
        if n_got == n_exp:
            assert rel_err(s2[0].cpu(), ref_score).max() < SCORE_RTOL, (thresh, max_iter, n_got, n_exp)


@pytest.mark.parametrize("case", CALC_CASES, ids=[c[0] for c in CALC_CASES])
def test_golden_calc_similarity(eng, golden_dir, case):
    """CUDA path against outputs of the REAL reference (tests/golden/calc_similarity.npz)."""
    name, seed, k, sigma, kw = case
    G = np.load(os.path.join(golden_dir, "calc_similarity.npz"))
    g = synth.make_gallery(k + 1, 128, 49, classes=2, seed=seed, sigma=sigma)
    mode = O.select_mode(kw.get("use_uniform", False), kw.get("use_inverse", False), kw.get("use_minus", False),
                         kw.get("use_soft", False))
    p = params(mode=mode, use_cls_token=kw.get("use_cls_token", False), temperature=kw.get("temperature", 1.0),
               ot_temp=kw.get("ot_temp", 0.05), ot_part=kw.get("ot_part", 1.0))
    score, uv, niter = eng.calc_similarity(g.patches[0], g.centers[0], g.patches[1:], g.centers[1:], p)
    n_ref = int(G[f"{name}_meta"][3])
    if int(niter) != n_ref:   # legal only at a borderline stop; the comparison below stays against the reference's output
        _, _, (n_o, errs) = O.structural_similarity(g.patches[0], g.centers[0], g.patches[1:], g.centers[1:], mode,
                                                    ot_temp=p.ot_temp, temperature=p.temperature,
                                                    use_cls_token=p.use_cls_token, ot_part=p.ot_part, trace=True)
        assert stop_ok(int(niter), n_o, errs), (int(niter), n_o, n_ref)
    assert rel_err(score.cpu(), G[f"{name}_score"]).max() < score_gate(niter, n_ref)
    np.testing.assert_allclose(uv[2].cpu(), G[f"{name}_T"], rtol=plan_tol(niter, n_ref), atol=1e-9)
    np.testing.assert_allclose(uv[3].cpu(), G[f"{name}_simr"], rtol=plan_tol(niter, n_ref), atol=1e-8)


@pytest.mark.parametrize("name", ["rollout", "rollout_iid", "rollout_part", "rollout_uniform"])
def test_golden_rollout(eng, golden_dir, name):
    G = np.load(os.path.join(golden_dir, "calc_similarity.npz"))
    seed, k, sigma, n_ref = [int(x) for x in G[f"{name}_meta"]]
    if sigma < 0:
        g = synth.make_gallery(k + 1, 128, 49, classes=2, seed=seed, structured=False)
    else:
        g = synth.make_gallery(k + 1, 128, 49, classes=2, seed=seed, sigma=sigma / 1000)
    p = params(mode="uniform" if name.endswith("uniform") else "rollout",
               ot_part=0.3 if name.endswith("part") else 1.0)
    score, uv, niter = eng.calc_similarity(g.patches[0], g.centers[0], g.patches[1:], g.centers[1:], p,
                                           q_rollout=g.rollout[0], c_rollout=g.rollout[1:])
    if int(niter) != n_ref:
        _, _, (n_o, errs) = O.structural_similarity(g.patches[0], g.centers[0], g.patches[1:], g.centers[1:], p.mode,
                                                    ot_part=p.ot_part, q_rollout=g.rollout[0],
                                                    c_rollout=g.rollout[1:], trace=True)
        assert stop_ok(int(niter), n_o, errs), (int(niter), n_o, n_ref)
    assert rel_err(score.cpu(), G[f"{name}_score"]).max() < score_gate(niter, n_ref)
    np.testing.assert_allclose(uv[2].cpu(), G[f"{name}_T"], rtol=plan_tol(niter, n_ref), atol=1e-9)


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_golden_sinkhorn(eng, golden_dir, name):
    G = np.load(os.path.join(golden_dir, "sinkhorn.npz"))
    seed, b, c, r, sigma, n_ref = [int(x) for x in G[f"{name}_meta"]]
    g = synth.make_gallery(b + 1, c, r, classes=2, seed=seed, sigma=sigma / 1000)
    K = O.gibbs(O.patch_similarity(g.patches[0], g.patches[1:]))
    u = g.rollout[1:] / (g.rollout[1:].sum(1, keepdim=True) + 1e-5)
    v = (g.rollout[0:1] / (g.rollout[0:1].sum(1, keepdim=True) + 1e-5)).expand(b, -1).contiguous()
    T, niter = eng.sinkhorn(K, u, v)
    # given K, u, v the CUDA Sinkhorn is bit-identical to torch on THIS host (test_sinkhorn_bit_exact_given_inputs); the
    # fixture was made on another host, whose exp() may differ in the last bit of K
    T_here, n_here, errs = O.sinkhorn(K, u, v, trace=True)
    assert int(niter) == n_here and torch.equal(T.cpu(), T_here)
    assert stop_ok(n_here, n_ref, errs)
    np.testing.assert_allclose(T.cpu(), G[f"{name}_T"], rtol=plan_tol(niter, n_ref), atol=1e-10)
    Ke, ue, ve = O.partial_extend(K, u, v, 0.5)
    Te, niter = eng.sinkhorn(Ke, ue, ve)
    n_refp = int(G[f"{name}_npartial"][0])
    _, n_here, errs = O.sinkhorn(Ke, ue, ve, trace=True)
    assert int(niter) == n_here and stop_ok(n_here, n_refp, errs)
    np.testing.assert_allclose(Te.cpu(), G[f"{name}_Tpartial"], rtol=plan_tol(niter, n_refp), atol=1e-10)


@pytest.mark.parametrize("c,r,k,mode,kw", [(32, 16, 9, "rollout", {}), (64, 36, 150, "rollout", {}),
                                           (128, 49, 130, "inverse", dict(temperature=0.1, use_cls_token=True)),
                                           (96, 196, 6, "minus", dict(use_cls_token=False, ot_part=0.5)),
                                           (128, 49, 12, "rollout", {})])
def test_generic_path(eng, c, r, k, mode, kw):
    """Shapes the fused kernel does not cover (and one it does, forced through here by K > 112 or
    another R/C) run the workspace-based path; same gates."""
    g = synth.make_gallery(k + 1, c, r, classes=3, seed=7 * k + r, sigma=0.6)
    ref_score, ref_uv, (n_ref, errs) = _oracle_pair(g, mode, **kw)
    score, uv, niter = eng.calc_similarity(g.patches[0], g.centers[0], g.patches[1:], g.centers[1:],
                                           params(mode=mode, **kw), q_rollout=g.rollout[0],
                                           c_rollout=g.rollout[1:])
    assert stop_ok(int(niter), n_ref, errs), f"n* {int(niter)} vs oracle {n_ref}, errs tail {errs[-3:]}"
    assert rel_err(score.cpu(), ref_score).max() < score_gate(niter, n_ref)
    np.testing.assert_allclose(uv[2].cpu(), ref_uv[2], rtol=plan_tol(niter, n_ref), atol=1e-9)


EVAL_CASES = [
    # n, classes, seed, sigma, truncs, flags
    (384, 12, 41, 0.6, [0, 100], dict(use_rollout=True, ot_part=1.0)),
    (256, 10, 42, 0.8, [0, 10, 50], dict(use_inverse=True, temperature=0.1, use_cls_token=True, ot_part=1.0)),
    (200, 8, 43, 0.6, [0, 20], dict(use_minus=True, ot_part=0.5)),
    (60, 4, 44, 0.6, [0, 100], dict(use_rollout=True, ot_part=1.0)),      # gallery smaller than K
    (300, 6, 45, 0.6, [0, 120], dict(use_rollout=True, ot_part=1.0)),     # K > 112: workspace path
]


@pytest.mark.parametrize("n,classes,seed,sigma,truncs,flags", EVAL_CASES)
def test_evaluate_matches_oracle(eng, n, classes, seed, sigma, truncs, flags):
    from vitrerank.engine import OTParams
    g = synth.make_gallery(n, 128, 49, classes=classes, seed=seed, sigma=sigma)
    ref = O.evaluate_banks(g.patches, g.centers, g.rollout, g.labels, trunc_nums=list(truncs), dump=True, **flags)
    p = OTParams.from_flags(**flags)
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    tal, nit = eng.evaluate(truncs, p, want_niter=True)
    n_ref = np.array([d["n_iter"] for d in ref["dump"]])
    flips = int((nit != n_ref).sum())
    for q, d in enumerate(ref["dump"]):
        assert stop_ok(int(nit[q]), d["n_iter"], d["errs"]), (q, int(nit[q]), d["n_iter"], d["errs"][-3:])
    scale = n / 100.0
    got = {"r1": tal[:, 0] / scale, "rp": tal[:, 1] / scale, "mapr": tal[:, 2] / scale}
    assert tal[0, 7] == n
    # un-forced: identical (==) when every query ran the oracle's iteration count; a query one iteration apart may
    # reorder near-equal totals, which moves a tally by at most 1 / scale per such query
    slack = flips / scale
    for key in ("r1", "rp", "mapr"):
        if flips == 0:
            assert (np.asarray(got[key]) == np.asarray(ref[key])).all(), (key, got[key], ref[key])
        else:
            np.testing.assert_allclose(got[key], ref[key], rtol=0, atol=slack + 1e-9)
    np.testing.assert_allclose(tal[:, 3:7] / scale, np.array(ref["recall_at_1_2_4_8"]), rtol=0, atol=slack + 1e-9)


@pytest.mark.parametrize("n,k,flags,nq", [
    (140, 113, dict(use_rollout=True, ot_part=1.0), 10),
    (300, 128, dict(use_rollout=True, ot_part=1.0), 10),
    (560, 500, dict(use_inverse=True, temperature=0.1, use_cls_token=True, ot_part=1.0), 6),
    (1040, 1024, dict(use_rollout=True, ot_part=1.0), 3),
    (1120, 1100, dict(use_rollout=True, ot_part=1.0), 2),      # beyond the 64-CTA groups: generic solver
])
def test_wide_shortlists_fused(eng, n, k, flags, nq):
    """113..1,024 candidates per query: ceil(K / 16) CTAs per query with the CTA-level exchange.  Per-pair scores,
    iteration counts and metrics of a strided query subset against the oracle."""
    from vitrerank.engine import OTParams
    g = synth.make_gallery(n, 128, 49, classes=max(3, n // 90), seed=n + k, sigma=0.6)
    stride = n // nq
    ids = list(range(0, n, stride))[:nq]
    p = OTParams.from_flags(**flags)
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    idx, approx = eng.stage0_topk(k, q_start=0, q_stride=stride, nq=nq)
    score, niter = eng.rerank_scores(idx, k, p, q_start=0, q_stride=stride)
    tal, _ = eng.finalize(idx, approx, score, k, [0, k], q_start=0, q_stride=stride)
    idx, score, niter = idx.cpu().numpy(), score.cpu().numpy(), niter.cpu().numpy()
    ref0 = O.evaluate_banks(g.patches, g.centers, g.rollout, g.labels, trunc_nums=[0, k], query_ids=ids, dump=True, **flags)
    flips = 0
    for q, d in enumerate(ref0["dump"]):
        assert stop_ok(int(niter[q]), d["n_iter"], d["errs"]), (q, int(niter[q]), d["n_iter"], d["errs"][-3:])
        flips += int(niter[q]) != d["n_iter"]
    for q, d in enumerate(ref0["dump"]):   # un-forced
        assert set(idx[q].tolist()) == set(d["top"].tolist())
        pos = {int(c): i for i, c in enumerate(idx[q])}
        mine = np.array([score[q, pos[int(c)]] for c in d["top"]])
        assert rel_err(mine, d["score"].numpy()).max() < score_gate(niter[q], d["n_iter"]), q
    scale = n / 100.0
    tal = tal.cpu().numpy()
    assert tal[0, 7] == nq
    for col, key in enumerate(("r1", "rp", "mapr")):
        np.testing.assert_allclose(tal[:, col] / scale, ref0[key], rtol=0, atol=flips / scale + 1e-9)


@pytest.mark.parametrize("n,k,flags,nq", [
    (400, 100, dict(use_rollout=True, ot_part=0.5), 12),
    (400, 100, dict(use_minus=True, ot_part=0.1), 12),
    (260, 37, dict(use_inverse=True, temperature=0.1, use_cls_token=True, ot_part=0.9), 13),
    (300, 112, dict(use_uniform=True, ot_part=0.6), 6),
])
def test_partial_ot_fused(eng, n, k, flags, nq):
    """Partial OT (one dummy point, diml.py:59-75) in the fused kernel -- the dummy row / column ride in the padding of
    strip 12 -- against the generic solver (same arithmetic order: iteration counts equal, scores equal to rounding of the
    tensor-core similarity) and against the oracle, un-forced."""
    from vitrerank.engine import OTParams
    g = synth.make_gallery(n, 128, 49, classes=max(3, n // 50), seed=n + k, sigma=0.6)
    stride = n // nq
    ids = list(range(0, n, stride))[:nq]
    p = OTParams.from_flags(**flags)
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    idx, approx = eng.stage0_topk(k, q_start=0, q_stride=stride, nq=nq)
    score, niter = eng.rerank_scores(idx, k, p, q_start=0, q_stride=stride)          # fused (scores only)
    eng.err_trace(nq)                                                                 # a diagnostics output: generic solver
    try:
        score_g, niter_g = eng.rerank_scores(idx, k, p, q_start=0, q_stride=stride)
    finally:
        eng.err_trace(0)
    idx, score, niter = idx.cpu().numpy(), score.cpu().numpy(), niter.cpu().numpy()
    score_g, niter_g = score_g.cpu().numpy(), niter_g.cpu().numpy()
    same = niter == niter_g
    assert same.sum() >= nq - 1, (niter, niter_g)
    assert rel_err(score[same], score_g[same]).max() < 2e-5
    ref0 = O.evaluate_banks(g.patches, g.centers, g.rollout, g.labels, trunc_nums=[0, k], query_ids=ids, dump=True, **flags)
    for q, d in enumerate(ref0["dump"]):
        assert stop_ok(int(niter[q]), d["n_iter"], d["errs"]), (q, int(niter[q]), d["n_iter"], d["errs"][-3:])
        pos = {int(c): i for i, c in enumerate(idx[q])}
        mine = np.array([score[q, pos[int(c)]] for c in d["top"]])
        assert rel_err(mine, d["score"].numpy()).max() < score_gate(niter[q], d["n_iter"]), q


@pytest.mark.parametrize("k", [1, 5, 8, 9, 24, 100, 105, 112])
def test_half_cta_packing_equals_whole_ctas(eng, k):
    """The score-only pair kernel packs a query's pairs into ceil(k / 8) warp groups that follow one another across CTA
    boundaries (a CTA may serve two queries); VR_PAIR_HALVES=0 keeps seven whole CTAs per query.  Same arithmetic, same
    exchange protocol: scores and iteration counts must be bit-identical -- for shortlists that fill warp groups exactly or
    not, an odd number of queries (the last CTA's second half has no query), full and partial OT."""
    from vitrerank.engine import OTParams
    g = synth.make_gallery(400, 128, 49, classes=8, seed=77 + k, sigma=0.6)
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    nq = 11
    idx, approx = eng.stage0_topk(max(k, 8), q_start=3, q_stride=7, nq=nq)
    for p in (OTParams(mode="rollout"), OTParams(mode="inverse", temperature=0.1, use_cls_token=True),
              OTParams(mode="rollout", ot_part=0.5)):
        os.environ.pop("VR_PAIR_HALVES", None)
        s1, n1 = eng.rerank_scores(idx, k, p, q_start=3, q_stride=7)
        os.environ["VR_PAIR_HALVES"] = "0"
        try:
            s0, n0 = eng.rerank_scores(idx, k, p, q_start=3, q_stride=7)
        finally:
            os.environ.pop("VR_PAIR_HALVES", None)
        assert torch.equal(n1, n0), (p, n1, n0)
        assert torch.equal(s1, s0), (p, (s1 - s0).abs().max())


@pytest.mark.parametrize("c,r", [(768, 196), (64, 36), (48, 130)])
def test_generic_operand_copy_equals_on_the_fly_split(eng, c, r, monkeypatch):
    """The generic path's S3 reads a registered bank from its pre-split operand copy (generic_repack: TMA + MMA only);
    VR_GENERIC_PACK=0 splits every pair's rows on the fly.  Same halves, same MMA order: bit-identical scores -- also after
    the bank is registered again with other contents (the copy must be rebuilt), one and two row tiles, ragged R.
    (VR_GENERIC_FUSED=0: the separate S3 / S4 kernels on both sides.)"""
    from vitrerank.engine import OTParams
    monkeypatch.setenv("VR_GENERIC_FUSED", "0")
    k, nq = 6, 5
    for seed in (3, 4):
        g = synth.make_gallery(40, c, r, classes=4, seed=seed, sigma=0.6)
        eng.register(g.patches, g.centers, g.rollout, g.labels)
        idx, _ = eng.stage0_topk(8, q_start=1, q_stride=3, nq=nq)
        for p in (OTParams(mode="rollout"), OTParams(mode="uniform", ot_part=0.6)):
            monkeypatch.delenv("VR_GENERIC_PACK", raising=False)
            s1, n1 = eng.rerank_scores(idx, k, p, q_start=1, q_stride=3)
            monkeypatch.setenv("VR_GENERIC_PACK", "0")
            s0, n0 = eng.rerank_scores(idx, k, p, q_start=1, q_stride=3)
            monkeypatch.delenv("VR_GENERIC_PACK", raising=False)
            assert torch.equal(n1, n0), (p, n1, n0)
            assert torch.equal(s1, s0), (p, (s1 - s0).abs().max())


@pytest.mark.parametrize("c,r,k,nq", [(768, 196, 10, 6), (64, 36, 7, 9), (48, 130, 5, 4), (32, 49, 200, 3)])
def test_generic_fused_equals_separate_kernels(eng, c, r, k, nq, monkeypatch):
    """generic_fused.cu (S3 + S4 in one kernel, a score per iteration, the stop decided afterwards) against the separate
    kernels of generic_ot.cu (VR_GENERIC_FUSED=0): the same K, the same FMA chains and the same stop arithmetic, so the
    iteration counts must be IDENTICAL and the scores equal up to the order of the final sum -- for loops that stop in
    the first pass of 8 iterations, in a later one (the work list), or never (max_iter not a multiple of 8), padded
    shortlist entries, more candidates per query than SMs, every marginal mode."""
    from vitrerank.engine import OTParams
    n = max(60, k + 20)
    g = synth.make_gallery(n, c, r, classes=5, seed=c + r, sigma=0.6)
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    idx, _ = eng.stage0_topk(k, q_start=2, q_stride=3, nq=nq)
    idx = idx.clone()
    idx[1, k - 2:] = -1                                   # padded entries take no part in the stop test and score 0
    longest = 0
    for p in (OTParams(mode="rollout"), OTParams(mode="uniform"), OTParams(mode="rollout", thresh=1e-6),
              OTParams(mode="rollout", thresh=1e-9, max_iter=19), OTParams(mode="uniform", ot_temp=0.1, thresh=1e-5, max_iter=40),
              # cross-correlation marginals, centres = patch means: generic_prepare_kernel writes them (marginals only) first
              OTParams(mode="relu"), OTParams(mode="soft", thresh=1e-3), OTParams(mode="inverse", temperature=0.1),
              # ... with cls centres: they come out of the MMA itself (the operand copy carries the centre as patch R), so the
              # marginals carry the split's 3e-7 instead of the fp32 chain's 1e-7: a stop test may flip in a borderline query
              OTParams(mode="inverse", temperature=0.1, use_cls_token=True), OTParams(mode="relu", use_cls_token=True),
              OTParams(mode="soft", use_cls_token=True, thresh=1e-3), OTParams(mode="minus", use_cls_token=True, thresh=1e-4)):
        monkeypatch.delenv("VR_GENERIC_FUSED", raising=False)
        s1, n1 = eng.rerank_scores(idx, k, p, q_start=2, q_stride=3)
        monkeypatch.setenv("VR_GENERIC_FUSED", "0")
        s0, n0 = eng.rerank_scores(idx, k, p, q_start=2, q_stride=3)
        monkeypatch.delenv("VR_GENERIC_FUSED", raising=False)
        in_mma = p.use_cls_token and p.mode in ("inverse", "relu", "soft", "minus")
        same = (n1 == n0)
        if in_mma:
            assert int((n1 - n0).abs().max()) <= 1 and same.float().mean() >= 0.75, (p, n1, n0)
        else:
            assert bool(same.all()), (p, n1, n0)
        assert (s1[1, k - 2:] == 0).all() and (s0[1, k - 2:] == 0).all()
        err = ((s1 - s0).abs() / s0.abs().clamp_min(1e-6))[same].max().item()
        assert err < (2e-5 if in_mma else 2e-6), (p, err)
        longest = max(longest, int(n1.max()))
    assert longest > 16                                   # (some case ran into the third pass)


def test_generic_rerank_runs_in_rounds_when_the_workspace_is_small(eng, monkeypatch):
    """generic_rerank with less workspace than all queries need at once runs as many queries per round as fit: the same
    scores and iteration counts as one round, for the separate kernels (2 R^2 floats per pair) and the fused kernel; a
    workspace below one query's need is refused (VR_E_WORKSPACE)."""
    from vitrerank.engine import OTParams
    from vitrerank._lib import VitRerankError, lib
    import ctypes as C
    c, r, k, nq = 64, 36, 6, 7
    g = synth.make_gallery(50, c, r, classes=4, seed=21, sigma=0.6)
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    idx, _ = eng.stage0_topk(8, q_start=1, q_stride=2, nq=nq)
    for fused, p in (("0", OTParams(mode="rollout")), ("0", OTParams(mode="inverse", temperature=0.1, ot_part=0.7)),
                     ("1", OTParams(mode="rollout"))):
        monkeypatch.setenv("VR_GENERIC_FUSED", fused)
        s_all, n_all = eng.rerank_scores(idx, k, p, q_start=1, q_stride=2)
        ps = p.struct()
        one = lib.vr_rerank_workspace_bytes(eng._h, 1, k, C.byref(ps))
        if fused == "1":                                  # (its figure covers the separate kernels' need for one query)
            one = 4096
        s_part, n_part = eng.rerank_scores(idx, k, p, q_start=1, q_stride=2, workspace_bytes=int(2.5 * one))
        assert torch.equal(n_part, n_all) and torch.equal(s_part, s_all), p
        with pytest.raises(VitRerankError, match="workspace"):
            eng.rerank_scores(idx, k, p, q_start=1, q_stride=2, workspace_bytes=one // 4)
    monkeypatch.delenv("VR_GENERIC_FUSED", raising=False)


def test_evaluate_stages_and_scores(eng):
    """Stage by stage on one gallery: shortlist sets, per-pair scores, reranked order."""
    from vitrerank.engine import OTParams
    n, k = 320, 100
    g = synth.make_gallery(n, 128, 49, classes=10, seed=77, sigma=0.6)
    ref0 = O.evaluate_banks(g.patches, g.centers, g.rollout, g.labels, trunc_nums=[0, k], use_rollout=True,
                            ot_part=1.0, dump=True)
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    idx, approx = eng.stage0_topk(k)
    score, niter = eng.rerank_scores(idx, k, OTParams(mode="rollout"))
    tal, rank = eng.finalize(idx, approx, score, k, [0, k], want_rank=True)
    idx, approx, score, niter, rank = [t.cpu().numpy() for t in (idx, approx, score, niter, rank)]
    for q, d in enumerate(ref0["dump"]):   # un-forced: the oracle keeps its own iteration counts
        assert stop_ok(int(niter[q]), d["n_iter"], d["errs"]), (q, int(niter[q]), d["n_iter"], d["errs"][-3:])
        assert set(idx[q].tolist()) == set(d["top"].tolist())
        pos = {int(c): i for i, c in enumerate(idx[q])}
        mine = np.array([score[q, pos[int(c)]] for c in d["top"]])
        gate = score_gate(niter[q], d["n_iter"])
        assert rel_err(mine, d["score"].numpy()).max() < gate, q
        ref_order = d["top"][d["rank"]].numpy()
        if not np.array_equal(rank[q], ref_order):
            tot = np.sort(d["total"].numpy())[::-1]
            assert np.min(np.abs(np.diff(tot))) < 10 * gate, f"query {q}: reranked order differs without a near tie"


def test_evaluate_host_equals_device(eng):
    from vitrerank.engine import OTParams
    g = synth.make_gallery(260, 128, 49, classes=9, seed=5, sigma=0.6)
    p = OTParams(mode="rollout")
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    a = eng.evaluate([0, 100], p)
    gp = g.pin()
    b = eng.evaluate_host(gp.patches, gp.centers, gp.rollout, gp.labels, [0, 100], p)
    np.testing.assert_array_equal(a, b)
    # query sharding: two interleaved shards add up to the whole pass
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    s0 = eng.evaluate([0, 100], p, q_start=0, q_stride=2)
    s1 = eng.evaluate([0, 100], p, q_start=1, q_stride=2)
    np.testing.assert_allclose(s0 + s1, a, rtol=1e-12)


def test_nan_propagation(eng):
    """A candidate whose whole marginal is zero makes the reference emit NaN and run all 100
    iterations; the CUDA path must do the same, not hide it."""
    g = synth.make_gallery(9, 128, 49, classes=2, seed=3, sigma=0.6)
    roll = g.rollout.clone()
    roll[0] = -1.0   # relu -> all zero query marginal
    ref_score, _, (n_ref, errs) = O.structural_similarity(g.patches[0], g.centers[0], g.patches[1:], g.centers[1:],
                                                          "rollout", q_rollout=roll[0], c_rollout=roll[1:],
                                                          trace=True)
    score, uv, niter = eng.calc_similarity(g.patches[0], g.centers[0], g.patches[1:], g.centers[1:],
                                           params(mode="rollout"), q_rollout=roll[0], c_rollout=roll[1:])
    assert n_ref == 100 and int(niter) == 100
    assert torch.isnan(ref_score).all() and torch.isnan(score).all()


@pytest.mark.parametrize("b,r,sigma", [(100, 49, 0.6), (37, 49, 1.0), (12, 50, 0.6), (9, 196, 0.6), (7, 16, 0.6), (5, 10, 1.0)])
def test_sinkhorn_bit_exact_given_inputs(eng, b, r, sigma):
    """Given the same K, u, v, the CUDA Sinkhorn reproduces torch's CPU result BIT FOR BIT: same
    sequential-FMA mat-vecs, IEEE division (utilities/diml.py:47-53), hence the same iteration
    count.  (End to end the inputs differ in the last bit -- torch's exp() is Intel VML -- see
    DESIGN.md "n* fragility".)"""
    g = synth.make_gallery(b + 1, 64, r, classes=2, seed=b * r, sigma=sigma)
    K = O.gibbs(O.patch_similarity(g.patches[0], g.patches[1:]))
    u = O._norm_sum(torch.relu(g.rollout[1:]))
    v = O._norm_sum(torch.relu(g.rollout[0].expand(b, -1)))
    T_ref, n_ref, errs = O.sinkhorn(K, u, v, trace=True)
    T, niter = eng.sinkhorn(K, u, v)
    assert int(niter) == n_ref
    assert torch.equal(T.cpu(), T_ref)


@pytest.mark.parametrize("mode,kw", [("rollout", {}), ("inverse", dict(temperature=0.1, use_cls_token=True)),
                                     ("minus", dict(use_cls_token=False))])
def test_packed_bank_path_equals_direct_path(eng, mode, kw, monkeypatch):
    """A registered bank is re-packed once into fp16 hi / lo operand planes and S3 is fed by TMA; direct calls with
    unregistered tensors split the fp32 rows on the fly.  Both feed the tensor cores the same bits, so scores and
    iteration counts must be identical, not just close.  (VR_PAIR_CC=fp32: with cls centres the registered path would
    otherwise take its cross-correlations out of the MMA, test_cross_correlations_out_of_the_mma.)"""
    monkeypatch.setenv("VR_PAIR_CC", "fp32")
    n, k = 150, 100
    g = synth.make_gallery(n, 128, 49, classes=5, seed=31, sigma=0.6)
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    idx, _ = eng.stage0_topk(k)
    p = params(mode=mode, **kw)
    score, niter = eng.rerank_scores(idx, k, p)
    idx_c = idx.cpu().long()
    for q in (0, 7, 149):
        s2, _, n2 = eng.calc_similarity(g.patches[q], g.centers[q], g.patches[idx_c[q]], g.centers[idx_c[q]], p,
                                        q_rollout=g.rollout[q], c_rollout=g.rollout[idx_c[q]], want_uv=False)
        assert int(n2) == int(niter[q])
        assert torch.equal(s2.cpu(), score[q].cpu())


@pytest.mark.parametrize("mode,kw", [("inverse", dict(temperature=0.1)), ("minus", {}), ("relu", {}), ("soft", {})])
def test_cross_correlations_out_of_the_mma(eng, mode, kw, monkeypatch):
    """With cls centres the operand copy of a registered bank carries every image's normalised centre as patch 50, so
    accumulator row / column 50 of the patch-similarity MMA ARE the cross-correlations <candidate centre, query patches>
    and <query centre, candidate patches> (utilities/diml.py:104-133): no second pass over the fp32 bank (VR_PAIR_CC=fp32).
    They carry the split's 3e-7 instead of the fp32 chain's 1e-7 -- the noise the Gibbs kernel already has -- so on a
    well-conditioned gallery (the Cars196-shaped one) most queries keep their iteration count, the others move by one, and
    where the count is kept every score agrees to 2e-5; K = 100 and the wide K = 300 groups, gallery queries and outside queries."""
    g = synth.make_named("cars196", seed=0, n=1560)
    sub = slice(0, 1500)
    eng.register(g.patches[sub], g.centers[sub], g.rollout[sub], g.labels[sub])
    p = params(mode=mode, use_cls_token=True, **kw)
    for k, nq in ((100, 300), (300, 40)):
        idx, _ = eng.stage0_topk(k, q_start=0, q_stride=5, nq=nq)
        monkeypatch.delenv("VR_PAIR_CC", raising=False)
        s1, n1 = eng.rerank_scores(idx, k, p, q_start=0, q_stride=5)
        monkeypatch.setenv("VR_PAIR_CC", "fp32")
        s0, n0 = eng.rerank_scores(idx, k, p, q_start=0, q_stride=5)
        same = n1 == n0
        assert same.float().mean() >= 0.75, (mode, k, same.float().mean().item())   # (relu: many marginal entries at 0, the touchiest)
        assert int((n1 - n0).abs().max()) <= 1, (mode, k, (n1 - n0).abs().max().item())
        rel = ((s1 - s0).abs() / s0.abs().clamp_min(1e-9))[same]
        assert rel.max().item() < 2e-5, (mode, k, rel.max().item())
    # queries that are not gallery items (training_tools/val.py:159-190): their operand copy is packed per call, centres included
    qs = slice(1500, 1560)
    idx, _ = eng.stage0_topk(100, q_centers=g.centers[qs])
    monkeypatch.delenv("VR_PAIR_CC", raising=False)
    s1, n1 = eng.rerank_scores_queries(g.patches[qs], g.centers[qs], idx, 100, p)
    monkeypatch.setenv("VR_PAIR_CC", "fp32")
    s0, n0 = eng.rerank_scores_queries(g.patches[qs], g.centers[qs], idx, 100, p)
    same = n1 == n0
    assert same.float().mean() >= 0.75 and int((n1 - n0).abs().max()) <= 1
    assert ((s1 - s0).abs() / s0.abs().clamp_min(1e-9))[same].max().item() < 2e-5


def test_exchange_transports_agree(eng, monkeypatch):
    """VR_PAIR_TRANSPORT=cluster (DSMEM, st.async) and the default global transport (tagged words in L2) run the same
    protocol: identical scores and iteration counts."""
    n, k = 120, 100
    g = synth.make_gallery(n, 128, 49, classes=4, seed=32, sigma=0.8)
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    idx, _ = eng.stage0_topk(k)
    p = params(mode="rollout")
    monkeypatch.delenv("VR_PAIR_TRANSPORT", raising=False)
    s_g, n_g = eng.rerank_scores(idx, k, p)
    torch.cuda.synchronize()
    monkeypatch.setenv("VR_PAIR_TRANSPORT", "cluster")
    s_c, n_c = eng.rerank_scores(idx, k, p)
    torch.cuda.synchronize()
    assert torch.equal(n_g.cpu(), n_c.cpu())
    assert torch.equal(s_g.cpu(), s_c.cpu())


def test_evaluate_host_sharded_single_rank(eng):
    """Without a process group the sharded host entry is evaluate_host on the whole query set."""
    from vitrerank import distributed as vd
    from vitrerank.engine import OTParams
    g = synth.make_gallery(260, 128, 49, classes=9, seed=5, sigma=0.6)
    p = OTParams(mode="rollout")
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    a = eng.evaluate([0, 100], p)
    gp = g.pin()
    b, h2d = vd.evaluate_host_sharded(eng, gp.patches, gp.centers, gp.rollout, gp.labels, [0, 100], p)
    np.testing.assert_array_equal(a, b)
    assert h2d == 260 * 128 * 49 * 4 + 260 * 128 * 4 + 260 * 49 * 4 + 260 * 8


NCCL_WORKER = '''
import os, sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, {pkg!r})
import numpy as np, torch, torch.distributed as dist
from vitrerank import synth, distributed as vd
from vitrerank.engine import OTParams, RerankEngine
rank = int(sys.argv[1])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2, device_id=dev)
g = synth.make_gallery(301, 128, 49, classes=9, seed=5, sigma=0.6)    # odd size: the last rank uploads one image less
eng = RerankEngine.get(dev)
gp = g.pin()
p = OTParams(mode="rollout")
for _ in range(2):   # the second pass reuses the staging buffer
    t, h2d = vd.evaluate_host_sharded(eng, gp.patches, gp.centers, gp.rollout, gp.labels, [0, 100], p)
eng.register(g.patches, g.centers, g.rollout, g.labels)
whole = eng.evaluate([0, 100], p)
if rank == 0:
    json.dump(dict(t=t.tolist(), whole=whole.tolist(), h2d=int(h2d)), open({out!r}, "w"))
dist.barrier(); dist.destroy_process_group()
'''


def test_evaluate_host_sharded_two_rank_nccl(tmp_path):
    """Each rank uploads half of the patch bank, NVLink all-gather, query shards, tally all-reduce:
    same tallies as one GPU over the whole gallery.  Needs two GPUs (skipped on a one-GPU box)."""
    import json
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "vit-reranking_b200")
    out = str(tmp_path / "t.json")
    script = tmp_path / "worker.py"
    script.write_text(NCCL_WORKER.format(root=root, pkg=pkg, port=29600 + (os.getpid() % 2000), out=out))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = json.load(open(out))
    np.testing.assert_allclose(np.array(got["t"]), np.array(got["whole"]), rtol=1e-12)
    # 8 pieces x 2 slices of 19 images: rank 0 uploads the even slices = 152 images
    assert got["h2d"] == 152 * 128 * 49 * 4 + 301 * 128 * 4 + 301 * 49 * 4 + 301 * 8


def test_full_size_properties_cars196(eng):
    """BASELINE configs[1] at full size (8,131 images, K = 100, rollout marginals): the oracle needs minutes for the
    whole pass, so the full-size run is checked through properties that do not depend on the size — sorted, duplicate-free,
    self-free shortlists; bit-identical repeat (every sum in the path has a fixed order or is an integer sum); query shards
    whose tallies add up to the whole pass — plus the oracle on a sample of queries."""
    from vitrerank.engine import OTParams
    g = synth.make_named("cars196", seed=0)
    n, k = g.patches.shape[0], 100
    p = OTParams(mode="rollout")
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    kp = max(k, eng.bank["max_num_pos"], 8)
    idx, approx = eng.stage0_topk(kp)
    score, niter = eng.rerank_scores(idx, k, p)
    tal, _ = eng.finalize(idx, approx, score, k, [0, k])
    idx_h, approx_h = idx.cpu().numpy(), approx.cpu().numpy()
    # shortlists: sorted by score, no duplicates, never the query itself
    assert (np.diff(approx_h, axis=1) <= 0).all()
    assert (np.sort(idx_h, axis=1)[:, 1:] != np.sort(idx_h, axis=1)[:, :-1]).all()
    assert (idx_h != np.arange(n)[:, None]).all() and idx_h.min() >= 0 and idx_h.max() < n
    assert torch.isfinite(score).all() and int(niter.min()) >= 1 and int(niter.max()) <= 100
    # repeat: bit-identical
    idx2, approx2 = eng.stage0_topk(kp)
    score2, niter2 = eng.rerank_scores(idx2, k, p)
    assert torch.equal(idx, idx2) and torch.equal(approx, approx2)
    assert torch.equal(score, score2) and torch.equal(niter, niter2)
    # whole pass = sum of three interleaved query shards; trunc 0 is the first stage alone
    whole = eng.evaluate([0, k], p)
    np.testing.assert_allclose(whole, tal.cpu().numpy(), rtol=1e-12)
    parts = sum(eng.evaluate([0, k], p, q_start=s, q_stride=3) for s in range(3))
    np.testing.assert_allclose(parts, whole, rtol=1e-12)
    assert whole[0, 7] == n and whole[1, 7] == n
    labels = g.labels.numpy()
    assert whole[0, 0] == float((labels[idx_h[:, 0]] == labels).sum())     # R@1 of the global ranking
    # oracle on a sample of queries: shortlist sets, iteration counts, per-pair scores
    ids = list(range(5, n, n // 12))[:12]
    ref0 = O.evaluate_banks(g.patches, g.centers, g.rollout, g.labels, trunc_nums=[0, k], use_rollout=True, ot_part=1.0,
                            query_ids=ids, dump=True)
    nit = niter.cpu().numpy()
    sc = score.cpu().numpy()
    for q, d in zip(ids, ref0["dump"]):    # un-forced (the whole pass is compared in tests/test_gpu_fullpass.py)
        assert stop_ok(int(nit[q]), d["n_iter"], d["errs"]), (q, int(nit[q]), d["n_iter"], d["errs"][-3:])
        assert set(idx_h[q, :k].tolist()) == set(d["top"].tolist())
        pos = {int(c): i for i, c in enumerate(idx_h[q, :k])}
        mine = np.array([sc[q, pos[int(c)]] for c in d["top"]])
        assert rel_err(mine, d["score"].numpy()).max() < score_gate(nit[q], d["n_iter"])
