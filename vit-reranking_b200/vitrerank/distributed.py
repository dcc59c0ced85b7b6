"""Multi-GPU: queries are independent units of the rerank path (each iteration of the reference's
loop at evaluation/eval_cvt_diml.py:316 only adds 3 x len(trunc_nums) scalars to the tallies
:370-372), so the path shards by QUERY with the gallery replicated on every GPU.  One process
per GPU (torchrun); rank r takes queries r, r + W, r + 2W, ... (interleaved, so that the
data-dependent Sinkhorn iteration counts balance); the only exchange is one all-reduce(sum) of
the [len(trunc_nums), 8] fp64 tallies (NCCL over NVLink on GPUs; gloo in the CPU tests).
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
