"""Small fixed workload for ncu captures: python tools/prof.py [n] [k] [reps] [mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vit-reranking_b200"))
import torch  # noqa: E402
from vitrerank import synth  # noqa: E402
from vitrerank.engine import OTParams, RerankEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mode = sys.argv[4] if len(sys.argv) > 4 else "rollout"
g = synth.make_gallery(n, 128, 49, classes=max(2, n // 83), seed=0, sigma=0.6)
eng = RerankEngine.get("cuda:0")
eng.register(g.patches, g.centers, g.rollout, g.labels)
kp = max(k, eng.bank["max_num_pos"], 8)
p = OTParams(mode=mode, use_cls_token=True, temperature=0.1)
for _ in range(reps):
    idx, approx = eng.stage0_topk(kp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    score, niter = eng.rerank_scores(idx, k, p)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rerank_scores: {ms:.3f} ms, {n * k / ms / 1e3:.2f} M pairs/s, {ms * 1e3 / n:.2f} us/query")
    tal, _ = eng.finalize(idx, approx, score, k, [0, k])
torch.cuda.synchronize()
print("n", n, "k", k, "niter mean", niter.float().mean().item(), "tallies", (tal[:, :3] / (n / 100)).tolist())
