"""Generate tests/golden/*.npz by running the REAL reference on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

It imports /root/reference/utilities/diml.py unmodified and loads
/root/reference/evaluation/metrics.py by file path (its package __init__ imports faiss),
feeds them seeded synthetic inputs and stores inputs' seeds + outputs.  The committed
fixtures pin oracle/rerank_oracle.py (tests/test_oracle_golden.py) and, through it, the
CUDA path.  Nothing here is imported at test time.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "vit-reranking_b200"))
sys.path.insert(0, REF)

import utilities.diml as ref_diml  # noqa: E402  (the reference's module)
from vitrerank import synth  # noqa: E402
sys.path.insert(0, HERE)
from cases import CALC_CASES, LOOP_CASES  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "evaluation", "metrics.py"))
ref_metrics = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_metrics)

torch.set_num_threads(1)


class MatmulCounter:
    """Sinkhorn issues two torch.matmul per iteration and one for T (diml.py:48-53)."""

    def __enter__(self):
        self.calls = 0
        self._orig = torch.matmul

        def counted(*a, **k):
            self.calls += 1
            return self._orig(*a, **k)
        torch.matmul = counted
        return self

    def __exit__(self, *exc):
        torch.matmul = self._orig

    @property
    def n_iter(self):
        return (self.calls - 1) // 2


def pair_inputs(seed, k, c=128, r=49, sigma=0.6):
    """A query (index 0) and k candidates from one class-structured gallery."""
    g = synth.make_gallery(k + 1, c, r, classes=2, seed=seed, sigma=sigma)
    return g


def gen_sinkhorn():
    out = {}
    for name, seed, b, r, sigma in [("a", 1, 7, 49, 0.6), ("b", 2, 5, 49, 1.0), ("c", 3, 3, 16, 0.3)]:
        g = pair_inputs(seed, b, 32, r, sigma)
        sim = torch.einsum('cm,ncs->nsm', g.patches[0], g.patches[1:]).contiguous()
        K = torch.exp(-(1.0 - sim) / 0.05)
        u = g.rollout[1:] / (g.rollout[1:].sum(1, keepdim=True) + 1e-5)
        v = (g.rollout[0:1] / (g.rollout[0:1].sum(1, keepdim=True) + 1e-5)).expand(b, -1).contiguous()
        with MatmulCounter() as mc:
            T = ref_diml.Sinkhorn(K, u, v)
        out[f"{name}_meta"] = np.array([seed, b, 32, r, int(sigma * 1000), mc.n_iter])
        out[f"{name}_T"] = T.numpy()
        with MatmulCounter() as mc:
            Te = ref_diml.Sinkhorn_partial(K, u, v, ot_part=0.5)
        out[f"{name}_Tpartial"] = Te.numpy()
        out[f"{name}_npartial"] = np.array([mc.n_iter])
    np.savez_compressed(os.path.join(HERE, "sinkhorn.npz"), **out)




def gen_calc_similarity():
    out = {}
    for name, seed, k, sigma, kw in CALC_CASES:
        g = pair_inputs(seed, k, sigma=sigma)
        with MatmulCounter() as mc:
            score, uv = ref_diml.calc_similarity(g.patches[0], g.centers[0], g.patches[1:], g.centers[1:],
                                                 stage=1, **kw)
        out[f"{name}_meta"] = np.array([seed, k, int(sigma * 1000), mc.n_iter])
        out[f"{name}_score"] = score.numpy()
        out[f"{name}_u"] = uv[0].numpy()
        out[f"{name}_v"] = uv[1].numpy()
        out[f"{name}_T"] = uv[2].numpy()
        out[f"{name}_simr"] = uv[3].numpy()
        if uv[4] is not None:
            out[f"{name}_cc"] = uv[4].numpy()
    # stage 0
    g = pair_inputs(19, 40)
    s0, none = ref_diml.calc_similarity(None, g.centers[0], None, g.centers, 0)
    assert none is None
    out["stage0_meta"] = np.array([19, 40])
    out["stage0_sim"] = s0.numpy()
    # rollout variant, full and partial
    for name, seed, k, sigma, kw in [("rollout", 21, 8, 0.6, {}), ("rollout_iid", 22, 8, -1.0, {}),
                                     ("rollout_part", 23, 6, 0.6, dict(ot_part=0.3)),
                                     ("rollout_uniform", 24, 4, 0.6, dict(use_uniform=True))]:
        if sigma < 0:
            g = synth.make_gallery(k + 1, 128, 49, classes=2, seed=seed, structured=False)
        else:
            g = pair_inputs(seed, k, sigma=sigma)
        with MatmulCounter() as mc:
            score, uv = ref_diml.calc_similarity_cvt_rollout(g.centers[0], g.patches[0], g.rollout[0],
                                                             g.centers[1:], g.patches[1:], g.rollout[1:],
                                                             stage=1, **kw)
        out[f"{name}_meta"] = np.array([seed, k, int(sigma * 1000), mc.n_iter])
        out[f"{name}_score"] = score.numpy()
        out[f"{name}_u"] = uv[0].numpy()
        out[f"{name}_v"] = uv[1].numpy()
        out[f"{name}_T"] = uv[2].numpy()
        out[f"{name}_simr"] = uv[3].numpy()
    np.savez_compressed(os.path.join(HERE, "calc_similarity.npz"), **out)


def gen_metrics():
    gen = torch.Generator().manual_seed(5)
    out = {}
    labels = synth.make_labels(300, 12, gen)
    rows = []
    tops_all = []
    for q in [0, 7, 150, 299]:
        tops = torch.randperm(300, generator=gen)
        r1, rp, mapr = ref_metrics.get_metrics_rank(tops, labels[q], labels)
        rows.append([q, r1, rp, mapr])
        tops_all.append(tops.numpy())
    out["labels"] = labels.numpy()
    out["tops"] = np.stack(tops_all)
    out["rows"] = np.array(rows, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)


def reference_loop(g, trunc_nums, use_rollout, **flags):
    """eval_cvt_diml.py:316-372,402-416 driven over banks, calling the reference's own
    calc_similarity / calc_similarity_cvt_rollout / get_metrics_rank."""
    n = g.patches.shape[0]
    sums = {t: [0.0, 0.0, 0.0] for t in trunc_nums}
    per_q = []
    for idx in range(n):
        anchor_center = g.centers[idx]
        anchor = g.patches[idx]
        approx_sim, _ = ref_diml.calc_similarity(None, anchor_center, None, g.centers, 0)
        approx_sim[idx] = -100
        approx_tops = torch.argsort(approx_sim, descending=True)
        top_inds = approx_tops[:max(trunc_nums)]
        with MatmulCounter() as mc:
            if not use_rollout:
                sim, uv = ref_diml.calc_similarity(anchor, anchor_center, g.patches[top_inds], g.centers[top_inds],
                                                   stage=1, use_uniform=flags.get("use_uniform", False),
                                                   use_inverse=flags.get("use_inverse", False),
                                                   temperature=flags.get("temperature", 1.0),
                                                   use_cls_token=flags.get("use_cls_token", False), ot_temp=0.05,
                                                   use_minus=flags.get("use_minus", False),
                                                   ot_part=flags.get("ot_part", 0.1))
            else:
                sim, uv = ref_diml.calc_similarity_cvt_rollout(anchor_center, anchor, g.rollout[idx],
                                                               g.centers[top_inds], g.patches[top_inds],
                                                               g.rollout[top_inds], stage=1,
                                                               use_uniform=flags.get("use_uniform", False),
                                                               use_ot=True, ot_part=flags.get("ot_part", 0.1))
        rank_in_tops = torch.argsort(sim + approx_sim[top_inds], descending=True)
        for t in trunc_nums:
            if t == 0:
                final_tops = approx_tops
            else:
                final_tops = torch.cat([top_inds[rank_in_tops][:t], approx_tops[t:]], dim=0)
            r1, rp, mapr = ref_metrics.get_metrics_rank(final_tops, g.labels[idx], g.labels)
            sums[t][0] += r1
            sums[t][1] += rp
            sums[t][2] += mapr
        per_q.append((top_inds.numpy(), sim.numpy(), approx_sim[top_inds].numpy(), mc.n_iter))
    res = {k: [sums[t][i] / float(n / 100) for t in trunc_nums] for i, k in enumerate(['r1', 'rp', 'mapr'])}
    return res, per_q




def gen_loop():
    out = {}
    for name, n, classes, seed, sigma, truncs, use_rollout, flags in LOOP_CASES:
        g = synth.make_gallery(n, 128, 49, classes=classes, seed=seed, sigma=sigma)
        res, per_q = reference_loop(g, truncs, use_rollout, **flags)
        out[f"{name}_meta"] = np.array([n, classes, seed, int(sigma * 1000)])
        out[f"{name}_truncs"] = np.array(truncs)
        out[f"{name}_metrics"] = np.array([res['r1'], res['rp'], res['mapr']], dtype=np.float64)
        out[f"{name}_top"] = np.stack([p[0] for p in per_q])
        out[f"{name}_score"] = np.stack([p[1] for p in per_q])
        out[f"{name}_approx"] = np.stack([p[2] for p in per_q])
        out[f"{name}_niter"] = np.array([p[3] for p in per_q])
        print(name, res, "n* min/mean/max", out[f"{name}_niter"].min(), out[f"{name}_niter"].mean(),
              out[f"{name}_niter"].max())
    np.savez_compressed(os.path.join(HERE, "loop.npz"), **out)


if __name__ == "__main__":
    gen_sinkhorn()
    gen_calc_similarity()
    gen_metrics()
    gen_loop()
    print("golden fixtures written to", HERE)
