"""Multi-GPU: queries are independent units of the rerank path (each iteration of the reference's
loop at evaluation/eval_cvt_diml.py:316 only adds 3 x len(trunc_nums) scalars to the tallies
:370-372), so the path shards by QUERY with the gallery replicated on every GPU.  One process
per GPU (torchrun); rank r takes queries r, r + W, r + 2W, ... (interleaved, so that the
data-dependent Sinkhorn iteration counts balance); the only exchange is one all-reduce(sum) of
the [len(trunc_nums), 8] fp64 tallies (NCCL over NVLink on GPUs; gloo in the CPU tests).

When the banks start on the HOST (the reference keeps the patch bank in CPU memory,
eval_cvt_diml.py:278), replicating the gallery would push the same 204 MB through every GPU's PCIe
link.  `evaluate_host_sharded` sends every image over PCIe once per node instead — the gallery is cut
into pieces of W slices, rank r uploads slice r of every piece — and the ranks all-gather each piece
over NVLink (NCCL, in place) as soon as it has landed, while the next piece is still on PCIe and
while stage 0, which needs only the centres, already runs.
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


import os as _os

HOST_PIECES = int(_os.environ.get("VR_HOST_PIECES", "8"))   # upload / all-gather / re-pack pipeline depth of evaluate_host_sharded


def evaluate_host_sharded(engine, patches, centers, rollout, labels, trunc_nums, params):
    """Host banks (CPU tensors, pinned for full PCIe speed) -> all-reduced tallies [len(trunc_nums), 8].
    Returns (tallies, h2d_bytes_of_this_rank).  Without an NCCL group of more than one rank this is
    engine.evaluate_host (the C-ABI host entry) on the rank's query shard plus the tally all-reduce.

    With W ranks every image crosses PCIe once per node: the gallery is cut into HOST_PIECES pieces of W slices; rank r
    uploads slice r of every piece on a copy stream, each piece is all-gathered in place over NVLink as soon as its
    slices have landed (while the next piece is still in flight on PCIe) and re-packed into the operand layout of the
    fused kernel (vr_bank_prepare) on a third stream, under the all-gather of the next piece.  Stage 0 needs only the centres and runs meanwhile on the
    main stream; the rerank waits for the last piece."""
    import torch.distributed as dist
    rank, w = world()
    n, c, r = patches.shape
    q_start, q_stride, nq = shard(n, rank, w)
    k = max(int(t) for t in trunc_nums)
    small = centers.numel() * 4 + (0 if rollout is None else rollout.numel() * 4) + labels.numel() * 8
    if w == 1 or dist.get_backend() != "nccl" or k == 0 or n < 2 * w * HOST_PIECES:
        tallies = engine.evaluate_host(patches, centers, rollout, labels, trunc_nums, params, q_start=q_start,
                                       q_stride=q_stride, nq=nq)
        return all_reduce_tallies(tallies, engine.device), patches.numel() * 4 + small
    dev = engine.device
    pieces = HOST_PIECES
    mp = (n + w * pieces - 1) // (w * pieces)       # images per (piece, rank) slice
    total = w * pieces * mp                          # >= n; the tail is never referenced
    st = getattr(engine, "_host_shard_state", None)
    if st is None or st["shape"] != (total, c, r):
        st = dict(shape=(total, c, r), buf=torch.empty(total, c, r, dtype=torch.float32, device=dev),
                  copy=torch.cuda.Stream(dev), side=torch.cuda.Stream(dev), prep=torch.cuda.Stream(dev),
                  landed=[torch.cuda.Event() for _ in range(pieces)], gathered=[torch.cuda.Event() for _ in range(pieces)],
                  ready=torch.cuda.Event())
        engine._host_shard_state = st
    buf, copy_stream, side, prep = st["buf"], st["copy"], st["side"], st["prep"]
    cur = torch.cuda.current_stream(dev)
    copy_stream.wait_stream(cur)               # the previous pass is done with the buffer
    side.wait_stream(cur)
    prep.wait_stream(cur)
    # the small banks (centres, rollout, labels: 43 MB at SOP scale) go up whole on every rank.  Sharding them too
    # (VR_SHARD_SMALL=1: 1/W each + three all-gathers before stage 0) saves PCIe bytes -- with 8 ranks pulling from one host the
    # aggregate is ~190 GB/s -- but measured 5 ms SLOWER per pass on 2 GPUs (78.1 -> 83.0 ms): the collectives put a rank
    # rendezvous in front of stage 0.  Off by default.
    ms = (n + w - 1) // w
    sm_ = st.get("small")
    if sm_ is None or sm_["n"] != n or (sm_["r"] is None) != (rollout is None):
        sm_ = dict(n=n, c=torch.empty(w * ms, c, dtype=torch.float32, device=dev),
                   r=None if rollout is None else torch.empty(w * ms, r, dtype=torch.float32, device=dev),
                   l=torch.zeros(w * ms, dtype=torch.int64, device=dev))
        st["small"] = sm_
    slo, shi = rank * ms, min(n, (rank + 1) * ms)
    h2d = 0
    shard_small = _os.environ.get("VR_SHARD_SMALL", "0") == "1"
    for dst, src in ((sm_["c"], centers), (sm_["r"], rollout), (sm_["l"], labels)):
        if dst is None:
            continue
        if not shard_small:
            dst[:n].copy_(src, non_blocking=True)
            h2d += src.numel() * src.element_size()
            continue
        if shi > slo:
            dst[slo:shi].copy_(src[slo:shi], non_blocking=True)
            h2d += (shi - slo) * src[0].numel() * src.element_size()
        dist.all_gather_into_tensor(dst.view(-1), dst[slo:slo + ms].reshape(-1))
    centers_d, labels_d = sm_["c"][:n], sm_["l"][:n]
    rollout_d = None if rollout is None else sm_["r"][:n]
    with torch.cuda.stream(copy_stream):
        for p in range(pieces):
            lo = (p * w + rank) * mp
            hi = min(n, lo + mp)
            if hi > lo:
                buf[lo:hi].copy_(patches[lo:hi], non_blocking=True)
                h2d += (hi - lo) * c * r * 4
            st["landed"][p].record(copy_stream)
    engine.register(buf[:n], centers_d, rollout_d, labels_d)     # pointers only; the patches are still in flight
    # three streams, one per resource: PCIe (copy), NVLink (side: the all-gathers), HBM (prep: the re-pack of a gathered piece
    # runs under the all-gather of the next one)
    with torch.cuda.stream(side):
        for p in range(pieces):
            side.wait_event(st["landed"][p])
            piece = buf[p * w * mp:(p + 1) * w * mp]
            dist.all_gather_into_tensor(piece.view(-1), piece[rank * mp:(rank + 1) * mp].view(-1))   # in place, NVLink
            st["gathered"][p].record(side)
    prep.wait_stream(cur)                      # the re-pack stores every image's centre in its operand copy: centres first
    with torch.cuda.stream(prep):
        for p in range(pieces):
            prep.wait_event(st["gathered"][p])
            lo = p * w * mp
            hi = min(n, lo + w * mp)
            if hi > lo:
                engine.prepare_bank(lo, hi - lo, stream=prep)
        st["ready"].record(prep)
    kp = max(k, engine.bank["max_num_pos"], 8)
    idx, approx = engine.stage0_topk(kp, q_start=q_start, q_stride=q_stride, nq=nq)
    cur.wait_event(st["ready"])
    score, _ = engine.rerank_scores(idx, k, params, q_start=q_start, q_stride=q_stride)
    t_dev = torch.zeros(len(trunc_nums), 8, dtype=torch.float64, device=dev)
    engine.finalize(idx, approx, score, k, trunc_nums, q_start=q_start, q_stride=q_stride, tallies=t_dev)
    dist.all_reduce(t_dev)
    return t_dev.cpu().numpy(), h2d
