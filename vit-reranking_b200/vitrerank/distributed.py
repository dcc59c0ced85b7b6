"""Multi-GPU: queries are independent units of the rerank path (each iteration of the reference's
loop at evaluation/eval_cvt_diml.py:316 only adds 3 x len(trunc_nums) scalars to the tallies
:370-372), so the path shards by QUERY with the gallery replicated on every GPU.  One process
per GPU (torchrun); rank r takes queries r, r + W, r + 2W, ... (interleaved, so that the
data-dependent Sinkhorn iteration counts balance); the only exchange is one all-reduce(sum) of
the [len(trunc_nums), 8] fp64 tallies (NCCL over NVLink on GPUs; gloo in the CPU tests).

When the banks start on the HOST (the reference keeps the patch bank in CPU memory,
eval_cvt_diml.py:278), replicating the gallery would push the same 204 MB through every GPU's PCIe
link.  `evaluate_host_sharded` sends every image over PCIe once per node instead — rank r uploads
images [r*m, (r+1)*m) — and the ranks all-gather the rest over NVLink (one NCCL all-gather, in
place), while stage 0, which needs only the centres, already runs.
"""
from __future__ import annotations

import numpy as np
import torch


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(n: int, rank: int, world_size: int):
    """Interleaved query shard of rank `rank`: (q_start, q_stride, nq)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    nq = (n - rank + world_size - 1) // world_size if n > rank else 0
    return rank, world_size, nq


def all_reduce_tallies(tallies: np.ndarray, device=None) -> np.ndarray:
    """Sum the per-rank tallies over the default group; identity without a group."""
    import torch.distributed as dist
    rank, w = world()
    if w == 1:
        return tallies
    backend = dist.get_backend()
    dev = torch.device(device) if (backend == "nccl" and device is not None) else torch.device("cpu")
    if backend == "nccl" and dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    t = torch.from_numpy(np.ascontiguousarray(tallies, dtype=np.float64)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def evaluate_sharded(engine, trunc_nums, params, gather_niter=False):
    """Run this rank's shard through `engine.evaluate` and all-reduce the tallies.  `engine` is a
    vitrerank.engine.RerankEngine (or anything with .bank['n'], .device and the same evaluate())."""
    rank, w = world()
    n = engine.bank["n"]
    q_start, q_stride, nq = shard(n, rank, w)
    if nq > 0:
        tallies, niter = engine.evaluate(trunc_nums, params, q_start=q_start, q_stride=q_stride, nq=nq,
                                         want_niter=True)
    else:
        tallies, niter = np.zeros((len(trunc_nums), 8)), np.zeros(0, dtype=np.int32)
    tallies = all_reduce_tallies(tallies, getattr(engine, "device", None))
    return tallies, niter


def evaluate_host_sharded(engine, patches, centers, rollout, labels, trunc_nums, params):
    """Host banks (CPU tensors, pinned for full PCIe speed) -> all-reduced tallies [len(trunc_nums), 8].
    Returns (tallies, h2d_bytes_of_this_rank).  Without an NCCL group of more than one rank this is
    engine.evaluate_host on the rank's query shard plus the tally all-reduce."""
    import torch.distributed as dist
    rank, w = world()
    n, c, r = patches.shape
    q_start, q_stride, nq = shard(n, rank, w)
    k = max(int(t) for t in trunc_nums)
    small = centers.numel() * 4 + (0 if rollout is None else rollout.numel() * 4) + labels.numel() * 8
    if w == 1 or dist.get_backend() != "nccl" or k == 0 or n < 2 * w:
        tallies = engine.evaluate_host(patches, centers, rollout, labels, trunc_nums, params, q_start=q_start,
                                       q_stride=q_stride, nq=nq)
        return all_reduce_tallies(tallies, engine.device), patches.numel() * 4 + small
    dev = engine.device
    m = (n + w - 1) // w                       # images uploaded per rank
    lo, hi = rank * m, min(n, (rank + 1) * m)
    st = getattr(engine, "_host_shard_state", None)
    if st is None or st["shape"] != (w * m, c, r):
        st = dict(shape=(w * m, c, r), buf=torch.empty(w * m, c, r, dtype=torch.float32, device=dev),
                  stream=torch.cuda.Stream(dev), event=torch.cuda.Event())
        engine._host_shard_state = st
    buf, copy_stream, ev = st["buf"], st["stream"], st["event"]
    cur = torch.cuda.current_stream(dev)
    copy_stream.wait_stream(cur)               # the previous pass is done with the buffer
    with torch.cuda.stream(copy_stream):
        if hi > lo:
            buf[lo:hi].copy_(patches[lo:hi], non_blocking=True)
        ev.record(copy_stream)
    centers_d = centers.to(dev, non_blocking=True)
    rollout_d = None if rollout is None else rollout.to(dev, non_blocking=True)
    labels_d = labels.to(dev, non_blocking=True)
    engine.register(buf[:n], centers_d, rollout_d, labels_d)     # pointers only; the patches are still in flight
    kp = max(k, engine.bank["max_num_pos"], 8)
    idx, approx = engine.stage0_topk(kp, q_start=q_start, q_stride=q_stride, nq=nq)
    cur.wait_event(ev)
    dist.all_gather_into_tensor(buf.view(-1), buf[lo:lo + m].view(-1))   # in place, NVLink
    score, _ = engine.rerank_scores(idx, k, params, q_start=q_start, q_stride=q_stride)
    t_dev = torch.zeros(len(trunc_nums), 8, dtype=torch.float64, device=dev)
    engine.finalize(idx, approx, score, k, trunc_nums, q_start=q_start, q_stride=q_stride, tallies=t_dev)
    dist.all_reduce(t_dev)
    return t_dev.cpu().numpy(), (hi - lo) * c * r * 4 + small
