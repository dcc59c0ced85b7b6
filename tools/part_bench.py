import os, sys
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, ROOT+"/vit-reranking_b200"]
import torch
from vitrerank import synth
from vitrerank.engine import OTParams, RerankEngine
g = synth.make_gallery(8131, 128, 49, classes=98, seed=0, sigma=0.6)
eng = RerankEngine.get("cuda:0"); eng.register(g.patches, g.centers, g.rollout, g.labels)
idx, approx = eng.stage0_topk(100)
for part in (1.0, 0.5):
    p = OTParams(mode="rollout", ot_part=part)
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); score, niter = eng.rerank_scores(idx, 100, p); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"ot_part {part}: {ms:.2f} ms {813100/ms/1e3:.2f} M pairs/s niter {niter.float().mean().item():.2f}")
