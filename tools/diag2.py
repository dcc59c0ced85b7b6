"""Is the fused Sinkhorn loop itself exact?  Take K, u, v exactly as the kernel computed them
(max_iter = 0 call: T = K) and run torch's CPU Sinkhorn on them; compare err traces and T."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vit-reranking_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from oracle import rerank_oracle as O  # noqa: E402
from vitrerank import synth  # noqa: E402
from vitrerank.engine import OTParams, RerankEngine  # noqa: E402

n, k = 192, 100
g = synth.make_gallery(n, 128, 49, classes=8, seed=123, sigma=0.6)
eng = RerankEngine.get("cuda:0")
for q in [0, 18]:
    approx = O.global_similarity(g.centers[q], g.centers).clone()
    approx[q] = -100
    top = torch.argsort(approx, descending=True)[:k]
    fb, fc, fr = g.patches[top], g.centers[top], g.rollout[top]
    _, uv0, _ = eng.calc_similarity(g.patches[q], g.centers[q], fb, fc, OTParams(mode="rollout", max_iter=0),
                                    q_rollout=g.rollout[q], c_rollout=fr)
    Kg, ug, vg = uv0[2].cpu(), uv0[0].cpu(), uv0[1].cpu()
    T_cpu, n_cpu, errs = O.sinkhorn(Kg, ug, vg, trace=True)
    tr = eng.err_trace(1, 100)
    score, uv, nit = eng.calc_similarity(g.patches[q], g.centers[q], fb, fc, OTParams(mode="rollout"),
                                         q_rollout=g.rollout[q], c_rollout=fr)
    t = tr[0].cpu().numpy()
    eng.err_trace(0)
    print(f"q={q}: fused n* {int(nit)}; torch-CPU Sinkhorn on the kernel's own K,u,v: n* {n_cpu}")
    for i in [0, 1, 2, 3, 5, 10, 20, 30]:
        if i < len(errs) and not np.isnan(t[i]):
            print(f"   it {i:3d}: gpu {t[i]:.9g} cpu {errs[i]:.9g} rel {t[i]/errs[i]-1:+.2e}")
    for it in [1, 2]:
        _, uvi, _ = eng.calc_similarity(g.patches[q], g.centers[q], fb, fc, OTParams(mode="rollout", max_iter=it),
                                        q_rollout=g.rollout[q], c_rollout=fr)
        Ti = O.sinkhorn(Kg, ug, vg, force_iters=it)
        d = (uvi[2].cpu() - Ti).abs() / Ti.abs().clamp_min(1e-30)
        bad = (d > 0).float().mean().item()
        print(f"   T after {it} it: max rel diff {d.max().item():.2e}, fraction of differing entries {bad:.3f}")
        if it == 1:
            # which pairs / rows / cols differ?
            dd = (uvi[2].cpu() != Ti)
            print("      differing pairs:", dd.any(dim=(1, 2)).nonzero().flatten().tolist()[:20])
            print("      rows differing (pair 0):", dd[0].any(dim=1).nonzero().flatten().tolist())
            print("      cols differing (pair 0):", dd[0].any(dim=0).nonzero().flatten().tolist())
    # the same through the generic kernel (direct sinkhorn) for reference
    Tg, nit2 = eng.sinkhorn(Kg, ug, vg)
    print(f"   generic vr_sinkhorn on the same K,u,v: n* {int(nit2)}; T bit-exact vs torch: {bool((Tg.cpu() == T_cpu).all())}")
