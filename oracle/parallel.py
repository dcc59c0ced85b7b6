"""Run the oracle loop (or the REAL reference loop) over many queries on all host cores (TEST INFRASTRUCTURE ONLY).

The reference's query loop (evaluation/eval_cvt_diml.py:316-372) is serial Python; its iterations are independent, so
the fair "all the host threads it can use" baseline is P worker processes, each running the unmodified loop
single-threaded on its share of the queries (intra-op threading of 49 x 49 matrices scales far worse: SURVEY.md
section 3.5 measured 8.2 k pairs/s on one thread against 18.9 k on eight).  Workers are forked, so they share the
banks copy-on-write and never touch CUDA.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import torch

_G = {}


def _work(args):
    ids, impl = args
    torch.set_num_threads(1)
    g, flags, truncs = _G["gal"], _G["flags"], _G["truncs"]
    t0 = time.perf_counter()
    if impl == "reference":
        from . import ref_loader
        kw = dict(flags)
        use_rollout = kw.pop("use_rollout", False)
        out = ref_loader.reference_loop(g.patches, g.centers, g.rollout, g.labels, truncs, use_rollout=use_rollout,
                                        query_ids=ids, **kw)
        recs = out["per_query"]
    else:
        from . import rerank_oracle as O
        out = O.evaluate_banks(g.patches, g.centers, g.rollout, g.labels, trunc_nums=list(truncs), query_ids=ids,
                               dump=True, **flags)
        recs = out["dump"]
    return ids, recs, time.perf_counter() - t0


def _to_numpy(rec):
    # torch tensors travel between processes as shared-memory handles that die with the sender: ship plain arrays
    return {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in rec.items()}


def _to_torch(rec):
    import numpy as np
    return {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in rec.items()}


def _worker(conn, jobs, deadline):
    out = []
    try:
        for j in jobs:
            if deadline is not None and time.perf_counter() > deadline:
                break
            ids, recs, dt = _work(j)
            out.append((ids, [_to_numpy(r) for r in recs], dt))
        conn.send(out)
    except BaseException as e:      # the parent re-raises
        conn.send(e)
    finally:
        conn.close()


def run(gal, query_ids, truncs, flags, procs=None, impl="port", chunk=8, budget_s=None):
    """-> (records, seconds, procs).  Without budget_s: one record per query, in the order of query_ids.  With budget_s
    every worker stops taking new chunks when the budget is spent, and the records of the chunks that were finished come
    back (each record carries its query id in rec["q"]).
    impl = "port": rerank_oracle.evaluate_banks records (top, approx, score, n_iter, errs, gap, metrics);
    impl = "reference": ref_loader.reference_loop records (top, score, metrics) from the reference's own functions."""
    procs = max(1, min(procs or os.cpu_count() or 1, 64))
    ids = [int(q) for q in query_ids]
    _G.update(gal=gal, flags=dict(flags), truncs=list(truncs))
    jobs = [(ids[i:i + chunk], impl) for i in range(0, len(ids), chunk)]
    procs = min(procs, max(1, len(jobs)))
    t0 = time.perf_counter()
    deadline = None if budget_s is None else t0 + budget_s
    done = {}
    if procs == 1:
        for j in jobs:
            if deadline is not None and time.perf_counter() > deadline:
                break
            r = _work(j)
            done[tuple(r[0])] = r[1]
    else:
        old = torch.get_num_threads()
        torch.set_num_threads(1)
        ctx = mp.get_context("fork")     # workers share the banks copy-on-write and never touch CUDA
        workers = []
        for w in range(procs):           # static interleaved assignment: chunks cost about the same
            recv, send = ctx.Pipe(duplex=False)
            pr = ctx.Process(target=_worker, args=(send, jobs[w::procs], deadline), daemon=True)
            pr.start()
            send.close()
            workers.append((pr, recv))
        err = None
        for pr, recv in workers:
            try:
                msg = recv.recv()
            except EOFError:
                msg = RuntimeError("oracle worker died")
            if isinstance(msg, BaseException):
                err = msg
            else:
                for r in msg:
                    done[tuple(r[0])] = [_to_torch(x) for x in r[1]]
            pr.join()
        torch.set_num_threads(old)
        if err is not None:
            raise err
    dt = time.perf_counter() - t0
    recs = []
    for j in jobs:
        if tuple(j[0]) in done:
            recs.extend(done[tuple(j[0])])
    return recs, dt, procs
