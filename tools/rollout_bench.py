"""Attention-rollout producer (csrc/rollout.cu) against the torch statements of eval_cvt_diml.py:54-146 run on the same GPU:
time per batch for a CvT-13-like stack (1 + 2 + 10 blocks) and for the largest single block.
usage: python tools/rollout_bench.py [batch]"""
import os
import sys
import time

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vit-reranking_b200"))
sys.path.insert(0, ROOT)
import evaluation.eval_cvt_diml as E  # noqa: E402
from vitrerank.engine import RerankEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = RerankEngine.get("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
sm = lambda *s: torch.softmax(torch.randn(*s, generator=g, device="cuda") * 1.5, dim=-1)
shapes = [(0, (B, 1, 3136, 784))] + [(1, (B, 3, 784, 196))] * 2 + [(2, (B, 6, 197, 50))] * 10
probs = [(si, sm(*s)) for si, s in shapes]
total_bytes = sum(p.numel() * 4 for _, p in probs)


def ours():
    mats = torch.stack([eng.rollout_block(p, drop_cls=(si == 2), grid=7) for si, p in probs])
    return eng.rollout_chain(mats)[-1].mean(1)


def torch_eager():
    resize = nn.AdaptiveAvgPool2d((7, 7))
    mats = torch.stack([E.resize_attn_map(E.filter_attention_map(p.clone(), 0.1, "min"), resize, si, 7) for si, p in probs])
    mats = mats + torch.eye(49, device="cuda")
    mats = mats / mats.sum(dim=-1).unsqueeze(-1)
    j = mats[0]
    for i in range(1, len(mats)):
        j = torch.bmm(mats[i], j)
    return j.mean(1)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, out


t1, o1 = timed(ours)
t2, o2 = timed(torch_eager)
print(f"batch {B}: attention of 13 blocks = {total_bytes / 1e9:.2f} GB")
print(f"  rollout.cu      {t1:8.2f} ms   ({total_bytes / t1 / 1e6:.0f} GB/s of attention read)")
print(f"  torch (same GPU){t2:8.2f} ms")
print(f"  marginals: max |diff| {float((o1 - o2).abs().max()):.2e} (the torch statements on the GPU use cuBLAS bmm and CUDA topk)")
si, p = probs[0]
t3, _ = timed(lambda: eng.rollout_block(p, drop_cls=False, grid=7))
print(f"  largest block {tuple(p.shape)}: {t3:.2f} ms = {p.numel() * 4 / t3 / 1e6:.0f} GB/s of its {p.numel() * 4 / 1e6:.0f} MB "
      f"(fuse + 3 select passes + mask + pool: 1 read of the attention, 1 write + 5 reads of the fused map)")
