import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vit-reranking_b200"))
import numpy as np, torch
from oracle import rerank_oracle as O
from vitrerank import synth
from vitrerank.engine import OTParams, RerankEngine
n = 384
g = synth.make_gallery(n, 128, 49, classes=12, seed=41, sigma=0.6)
eng = RerankEngine.get("cuda:0")
eng.register(g.patches, g.centers, g.rollout, g.labels)
kp = max(100, eng.bank["max_num_pos"])
idx, sc = eng.stage0_topk(kp)
tal, _ = eng.finalize(idx, sc, None, 0, [0])
idx = idx.cpu(); sc = sc.cpu()
print("max_num_pos", eng.bank["max_num_pos"], "kp", kp, "tallies", tal.cpu().numpy()[0, :3] / (n / 100))
tot = [0, 0, 0]; tot_o = [0, 0, 0]
for q in range(n):
    sim = O.global_similarity(g.centers[q], g.centers).clone(); sim[q] = -100
    order = torch.argsort(sim, descending=True)
    mo = O.metrics_rank(order, g.labels[q], g.labels)
    npos = int((g.labels == g.labels[q]).sum())
    mine = torch.cat([idx[q].long(), order[kp:]])
    mm = O.metrics_rank(mine, g.labels[q], g.labels)
    for i in range(3): tot[i] += mm[i]; tot_o[i] += mo[i]
    if abs(mm[1] - mo[1]) > 1e-9 or abs(mm[2] - mo[2]) > 1e-9:
        d = (idx[q, :npos].long() != order[:npos]).nonzero().flatten().tolist()
        print("query", q, "np", npos, "mine", mm, "oracle", mo, "first diffs at", d[:6])
        for j in d[:4]:
            a, b = int(idx[q, j]), int(order[j])
            print("   pos", j, "mine", a, float(sim[a]), "lab", int(g.labels[a]), "| oracle", b, float(sim[b]), "lab", int(g.labels[b]), "qlab", int(g.labels[q]))
print("python-metrics on my lists:", [t / (n / 100) for t in tot], "oracle:", [t / (n / 100) for t in tot_o])
