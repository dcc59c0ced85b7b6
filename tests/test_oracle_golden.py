"""Pin oracle/rerank_oracle.py against fixtures produced by the REAL reference
(tests/golden/make_golden.py).  CPU only.

Tolerances: the fixtures were made with the same torch build, so on the same CPU the
match is bit-exact; another host CPU may pick another BLAS kernel (different fp32
summation order), hence rtol 2e-5 on plans/scores and exact equality on iteration
counts (cases were chosen with the stop test far from the 0.1 threshold).
"""
import os

import numpy as np
import pytest
import torch

from oracle import rerank_oracle as O
from vitrerank import synth

from golden.cases import CALC_CASES, LOOP_CASES

RTOL = 2e-5
ATOL = 1e-7


def close(a, b, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


@pytest.fixture(scope="module")
def G(golden_dir):
    return {k: np.load(os.path.join(golden_dir, k + ".npz")) for k in
            ("sinkhorn", "calc_similarity", "metrics", "loop")}


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_sinkhorn(G, name):
    g = G["sinkhorn"]
    seed, b, c, r, sigma, n_ref = [int(x) for x in g[f"{name}_meta"]]
    gal = synth.make_gallery(b + 1, c, r, classes=2, seed=seed, sigma=sigma / 1000)
    sim = O.patch_similarity(gal.patches[0], gal.patches[1:])
    K = O.gibbs(sim)
    u = gal.rollout[1:] / (gal.rollout[1:].sum(1, keepdim=True) + 1e-5)
    v = (gal.rollout[0:1] / (gal.rollout[0:1].sum(1, keepdim=True) + 1e-5)).expand(b, -1).contiguous()
    T, n_iter, errs = O.sinkhorn(K, u, v, trace=True)
    assert n_iter == n_ref
    close(T, g[f"{name}_T"])
    Te, n_p, _ = O.sinkhorn_partial(K, u, v, ot_part=0.5, trace=True)
    assert n_p == int(g[f"{name}_npartial"][0])
    assert Te.shape == (b, r + 1, r + 1)
    close(Te, g[f"{name}_Tpartial"])


@pytest.mark.parametrize("case", CALC_CASES, ids=[c[0] for c in CALC_CASES])
def test_calc_similarity(G, case):
    name, seed, k, sigma, kw = case
    g = G["calc_similarity"]
    assert [int(x) for x in g[f"{name}_meta"]][:3] == [seed, k, int(sigma * 1000)]
    gal = synth.make_gallery(k + 1, 128, 49, classes=2, seed=seed, sigma=sigma)
    mode = O.select_mode(kw.get("use_uniform", False), kw.get("use_inverse", False),
                         kw.get("use_minus", False), kw.get("use_soft", False))
    score, uv, (n_iter, errs) = O.structural_similarity(
        gal.patches[0], gal.centers[0], gal.patches[1:], gal.centers[1:], mode,
        ot_temp=kw.get("ot_temp", 0.05), temperature=kw.get("temperature", 1.0),
        use_cls_token=kw.get("use_cls_token", False), ot_part=kw.get("ot_part", 1.0), trace=True)
    assert n_iter == int(g[f"{name}_meta"][3])
    close(score, g[f"{name}_score"])
    close(uv[0], g[f"{name}_u"])
    close(uv[1], g[f"{name}_v"])
    close(uv[2], g[f"{name}_T"])
    close(uv[3], g[f"{name}_simr"])
    if f"{name}_cc" in g.files:
        close(uv[4], g[f"{name}_cc"], atol=1e-6)
    else:
        assert uv[4] is None


def test_stage0(G):
    g = G["calc_similarity"]
    seed, k = [int(x) for x in g["stage0_meta"]]
    gal = synth.make_gallery(k + 1, 128, 49, classes=2, seed=seed, sigma=0.6)
    close(O.global_similarity(gal.centers[0], gal.centers), g["stage0_sim"], atol=1e-6)


@pytest.mark.parametrize("name,kw", [("rollout", {}), ("rollout_iid", {}), ("rollout_part", dict(ot_part=0.3)),
                                     ("rollout_uniform", dict(use_uniform=True))])
def test_rollout(G, name, kw):
    g = G["calc_similarity"]
    seed, k, sigma, n_ref = [int(x) for x in g[f"{name}_meta"]]
    if sigma < 0:
        gal = synth.make_gallery(k + 1, 128, 49, classes=2, seed=seed, structured=False)
    else:
        gal = synth.make_gallery(k + 1, 128, 49, classes=2, seed=seed, sigma=sigma / 1000)
    mode = "uniform" if kw.get("use_uniform") else "rollout"
    score, uv, (n_iter, errs) = O.structural_similarity(
        gal.patches[0], gal.centers[0], gal.patches[1:], gal.centers[1:], mode, ot_temp=0.05,
        ot_part=kw.get("ot_part", 1.0), q_rollout=gal.rollout[0], c_rollout=gal.rollout[1:], trace=True)
    assert n_iter == n_ref
    close(score, g[f"{name}_score"])
    close(uv[0], g[f"{name}_u"])
    close(uv[1], g[f"{name}_v"])
    close(uv[2], g[f"{name}_T"])
    close(uv[3], g[f"{name}_simr"])


def test_metrics_rank(G):
    g = G["metrics"]
    labels = torch.from_numpy(g["labels"])
    for row, tops in zip(g["rows"], g["tops"]):
        q = int(row[0])
        r1, rp, mapr = O.metrics_rank(torch.from_numpy(tops), labels[q], labels)
        assert r1 == row[1]
        assert abs(rp - row[2]) < 1e-7
        assert abs(mapr - row[3]) < 1e-6


@pytest.mark.parametrize("case", LOOP_CASES, ids=[c[0] for c in LOOP_CASES])
def test_query_loop(G, case):
    name, n, classes, seed, sigma, truncs, use_rollout, flags = case
    g = G["loop"]
    gal = synth.make_gallery(n, 128, 49, classes=classes, seed=seed, sigma=sigma)
    out = O.evaluate_banks(gal.patches, gal.centers, gal.rollout, gal.labels, trunc_nums=list(truncs),
                           use_rollout=use_rollout, dump=True, **flags)
    ref = g[f"{name}_metrics"]
    close(out['r1'], ref[0], rtol=1e-9, atol=1e-9)
    close(out['rp'], ref[1], rtol=1e-6)
    close(out['mapr'], ref[2], rtol=1e-6)
    tops = np.stack([d["top"].numpy() for d in out["dump"]])
    assert np.array_equal(np.sort(tops, 1), np.sort(g[f"{name}_top"], 1))
    close(np.stack([d["score"].numpy() for d in out["dump"]]), g[f"{name}_score"])
    assert [d["n_iter"] for d in out["dump"]] == g[f"{name}_niter"].tolist()


def _torch_sum_order(x, W=8):
    """Python mirror of csrc/common.cuh::torch_sum_inner_w (ATen SumKernel.cpp order)."""
    f32 = np.float32
    n = x.shape[1]
    if n < 8:
        W = 1
    ILP, LEVELS = 4, 4
    vec_size = n // W
    size_ilp = vec_size // ILP
    acc = np.zeros((LEVELS, ILP, x.shape[0], W), dtype=f32)
    lg = 0
    while (1 << lg) < size_ilp:
        lg += 1
    level_power = max(4, lg // LEVELS)
    level_step = 1 << level_power
    level_mask = level_step - 1
    vec = lambda i: x[:, i * W:(i + 1) * W]
    i = 0
    while i + level_step <= size_ilp:
        for _ in range(level_step):
            for k in range(ILP):
                acc[0, k] = acc[0, k] + vec(i * ILP + k)
            i += 1
        for j in range(1, LEVELS):
            acc[j] = acc[j] + acc[j - 1]
            acc[j - 1] = 0
            if (i & (level_mask << (j * level_power))) != 0:
                break
    while i < size_ilp:
        for k in range(ILP):
            acc[0, k] = acc[0, k] + vec(i * ILP + k)
        i += 1
    for j in range(1, LEVELS):
        acc[0] = acc[0] + acc[j]
    for i in range(size_ilp * ILP, vec_size):
        acc[0, 0] = acc[0, 0] + vec(i)
    for k in range(1, ILP):
        acc[0, 0] = acc[0, 0] + acc[0, k]
    fin = np.zeros(x.shape[0], dtype=f32)
    for k in range(vec_size * W, n):
        fin = fin + x[:, k]
    for l in range(W):
        fin = fin + acc[0, 0][:, l]
    return fin


@pytest.mark.parametrize("n", [5, 16, 49, 50, 196, 197, 600, 1100])
def test_torch_sum_order(n):
    """The CUDA kernels add the marginal numerators in ATen's CPU order (common.cuh); this pins
    that order against the installed torch, bit for bit.  If a torch upgrade changes it, the
    Sinkhorn iteration counts of the CUDA path and of the reference drift apart."""
    rng = np.random.default_rng(n)
    x = (np.abs(rng.standard_normal((4000, n))) * rng.random((4000, 1))).astype(np.float32)
    ref = torch.from_numpy(x).sum(dim=1, keepdim=True).numpy()[:, 0]
    got = _torch_sum_order(x)
    assert np.array_equal(got.view(np.int32), ref.view(np.int32))


# ---- the restatement against the REAL reference functions, live (checkout here, oracle/_ref on the GPU box) ----
@pytest.mark.parametrize("flags", [dict(use_rollout=True, ot_part=1.0),
                                   dict(use_inverse=True, temperature=0.1, use_cls_token=True, ot_part=1.0),
                                   dict(use_minus=True, ot_part=0.5)])
def test_oracle_equals_real_reference_loop(flags):
    from oracle import ref_loader as RL
    if RL.root() is None:
        pytest.skip("neither /root/reference nor oracle/_ref present (python oracle/make_ref.py)")
    g = synth.make_gallery(150, 128, 49, classes=5, seed=17, sigma=0.6)
    ids = list(range(0, 150, 6))
    kw = dict(flags)
    use_rollout = kw.pop("use_rollout", False)
    ref = RL.reference_loop(g.patches, g.centers, g.rollout, g.labels, [0, 20, 100], use_rollout=use_rollout,
                            query_ids=ids, **kw)
    out = O.evaluate_banks(g.patches, g.centers, g.rollout, g.labels, trunc_nums=[0, 20, 100], query_ids=ids, dump=True,
                           **flags)
    for a, b in zip(ref["per_query"], out["dump"]):
        assert torch.equal(a["score"], b["score"])            # bit-identical per-pair scores
        assert a["metrics"] == b["metrics"]                   # and per-query r1 / rp / mapr
    for key in ("r1", "rp", "mapr"):
        assert ref[key] == out[key]


def test_parallel_runner_and_parity_counters():
    from oracle import parallel as OP
    from oracle import parity as PAR
    g = synth.make_gallery(120, 128, 49, classes=4, seed=23, sigma=0.6)
    ids = list(range(1, 120, 5))
    flags = dict(use_rollout=True, ot_part=1.0)
    serial = O.evaluate_banks(g.patches, g.centers, g.rollout, g.labels, trunc_nums=[0, 50], query_ids=ids, dump=True,
                              **flags)["dump"]
    par, _, procs = OP.run(g, ids, [0, 50], flags, procs=3, chunk=4)
    assert len(par) == len(ids) and [d["q"] for d in par] == ids
    for a, b in zip(serial, par):
        assert torch.equal(a["score"], b["score"]) and a["n_iter"] == b["n_iter"] and a["metrics"] == b["metrics"]
    # the oracle against itself: nothing to count
    idx = np.stack([d["top"].numpy() for d in par])
    score = np.stack([d["score"].numpy() for d in par])
    nit = np.array([d["n_iter"] for d in par])
    pq = np.array([[list(d["metrics"][t]) for t in (0, 50)] for d in par])
    c = PAR.compare(par, idx, score, nit, 50, trunc_nums=[0, 50], per_query=pq)
    assert c["niter_equal"] == len(ids) and c["pairs_over_1e-4"] == 0 and c["metric_mismatch_queries"] == 0
    assert c["stage0_set_mismatch"] == 0 and c["max_rel_err"] == 0.0 and c["pairs"] == 50 * len(ids)
    # and a doctored copy: one iteration count off, one score beyond the gate, one shortlist member swapped
    nit2, score2, idx2 = nit.copy(), score.copy(), idx.copy()
    nit2[0] += 1
    score2[1, 3] *= 1.001
    idx2[2, 0] = [i for i in range(120) if i not in set(idx[2].tolist())][0]
    c = PAR.compare(par, idx2, score2, nit2, 50, trunc_nums=[0, 50], per_query=pq)
    assert c["niter_off_by_one"] == 1 and c["pairs_over_1e-4"] == 1 and c["stage0_set_mismatch"] == 1
    assert c["pairs_over_1e-4_in_equal_niter_queries"] == 1 and c["queries_with_pairs_over_1e-4"] == 1
