"""Stage-0 timing: tensor-core path (stage0_mma.cu) against the fp32 SGEMM + row-select path, CUDA events, per shape.
    python tools/s0_bench.py [shape ...]      shapes: cars196 sop sop8 (the 1/8 query shard of SOP)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vit-reranking_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402
from vitrerank.engine import RerankEngine  # noqa: E402
from test_gpu_stage0_mma import centers_only, register_centers  # noqa: E402

SH = {"cars196": (8131, 98, {}), "sop": (60502, 11316, {}), "sop8": (60502, 11316, dict(q_start=0, q_stride=8)),
      "cub200": (5924, 100, {})}
eng = RerankEngine.get("cuda:0")
for name in (sys.argv[1:] or ["cars196", "sop", "sop8"]):
    n, classes, kw = SH[name]
    centers, labels = centers_only(n, classes=classes, seed=0)
    register_centers(eng, centers, labels)
    for path in ("mma", "sgemm"):
        if path == "sgemm":
            os.environ["VR_STAGE0"] = "sgemm"
        else:
            os.environ.pop("VR_STAGE0", None)
        for _ in range(2):
            eng.stage0_topk(100, **kw)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        for i in range(5):
            ev[i].record()
            eng.stage0_topk(100, **kw)
        ev[5].record()
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(5))
        print(f"{name:8s} {path:6s} median {ms[2]:8.3f} ms  min {ms[0]:8.3f}  stats {eng.stage0_stats()}", flush=True)
os.environ.pop("VR_STAGE0", None)
