#!/usr/bin/env python
"""Benchmark of the DIML rerank hot path (BASELINE.json metric: reranked query-candidate
pairs/s at K=100, 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one full pass of the hot path: every gallery image is a query, the first stage
shortlists K'=max(K, max num_pos) candidates, K of them are reranked with Sinkhorn OT, blended,
re-sorted and tallied into r1 / RP / MAP@R.  Default workload = BASELINE.json configs[2], the SOP
test shape (60,502 images, 7x7 patches, embed_dim 128, K = 100, rollout marginals): it is the gallery
north_star puts the scaling target on, it fits one GPU (1.52 GB), and every N uses it, so BENCH, SCALE
and the efficiency computed from them share one config.  --workload cars196 | cub200 | sop_k1000 |
vitb16 select the other BASELINE configs.  With N GPUs the queries of the SAME pass are sharded
(interleaved) over the ranks, the gallery is replicated and only the tallies are all-reduced (NCCL)
-> strong scaling.

JSON line: `value` is measured with the banks resident in HBM (CUDA events on the launching stream);
`e2e` through the host-buffer entry (pinned host banks -> device, tallies -> host inside the timed
region); `roofline` (HBM) and `roofline_fp32` (useful Sinkhorn FMAs against the
FP32 pipe, the roof that actually binds) are for the dominant kernel (pair_fused_kernel) from CUDA
events around its launches inside the timed steps; `cpu_baseline` is the reference's own functions
(oracle/_ref, kind "reference"; else the oracle port) on this box's host cores on a bounded sample;
`parity` are the un-forced counters of that same sample against this run's CUDA results.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "vit-reranking_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

BYTES_PER_PAIR = {(128, 49): 25796, (768, 196): 605968}   # C*R*4 + C*4 + R*4
WORKLOADS = {
    # name: (synth shape name, default n (None = the shape's), K, flags, description)
    "sop": ("sop", None, 100, dict(use_rollout=True, ot_part=1.0),
            "SOP test shape: 60502 images, C=128, R=49 (7x7), top-100 OT rerank, rollout marginals"),
    "cars196": ("cars196", None, 100, dict(use_rollout=True, ot_part=1.0),
                "Cars196 test shape: 8131 images, C=128, R=49 (7x7), top-100 OT rerank, rollout marginals"),
    "cub200": ("cub200", None, 100, dict(use_rollout=True, ot_part=1.0),
               "CUB-200 test shape: 5924 images, C=128, R=49, top-100 OT rerank, rollout marginals"),
    "sop_k1000": ("sop", None, 1000, dict(use_inverse=True, temperature=0.1, use_cls_token=True, ot_part=1.0),
                  "SOP shape, top-1000, calc_similarity + use_inverse T=0.1"),
    "vitb16": ("sop_vitb16", 4096, 100, dict(use_rollout=True, ot_part=1.0),
               "ViT-B/16 shape: C=768, R=196 (14x14), top-100 OT rerank, rollout marginals, 4096-image synthetic gallery"),
}
WORKLOADS["sop_vitb16"] = WORKLOADS["vitb16"]   # (alias: BASELINE.json configs[4])
WORKLOADS["vitb16_cc"] = ("sop_vitb16", 4096, 100, dict(use_inverse=True, temperature=0.1, use_cls_token=True, ot_part=1.0),
                          "ViT-B/16 shape with the cross-correlation marginals of eval_attn_diml.py's calc_similarity call "
                          "(use_inverse T=0.1, cls centres), 4096-image synthetic gallery")
METRIC = "reranked query-candidate pairs/sec at K=100"
FP32_LANES_PER_SM = 128


def static_config(workload, n, c, r, k, truncs, mode):
    """The keys both arms print, identical by construction."""
    return {"workload": WORKLOADS[workload][4], "name": workload, "n": n, "c": c, "r": r, "k": k,
            "trunc_nums": list(truncs), "marginals": mode, "queries_per_step": n,
            "l2": "inputs larger than L2 (patch bank %.0f MB, re-packed operand bank %.0f MB)" %
                  (n * c * r * 4 / 1e6, n * 65536 / 1e6 if (c, r) == (128, 49) else
                   (n * c * ((r + 15) // 16 * 16) * 4 / 1e6 if c % 16 == 0 and r <= 224 else 0.0))}


def static_traffic(workload, pairs_per_launch):
    """DRAM bytes of one pair-kernel launch from the committed ncu --set full capture (profiles/r2_traffic.json), when it
    was taken on this workload and launch size: a STATIC figure (ncu cannot run inside the timed bench); None otherwise."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            rows = t if isinstance(t, list) else [t]
            for row in rows:
                if row["workload"] == workload and int(row["pairs_per_launch"]) == int(pairs_per_launch):
                    return float(row["dram_bytes_read"]) + float(row["dram_bytes_write"]), name
        except Exception:
            pass
    return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", 1965.0


def tensor_peak():
    """Dense bf16 / fp16 tensor peak in TFLOP/s: the burst figure of MEASURED_PEAKS.json (a kernel timed alone), else nominal."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    except Exception:
        return 2250.0, "fallback (nominal dense bf16)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        # NVML in-process when available (a sample takes well under a millisecond, so short timed regions still get tens of
        # samples); the nvidia-smi query of the profiling recipe otherwise (~100 ms per sample)
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            bits = (("hw_slowdown", N.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", N.nvmlClocksEventReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", N.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", N.nvmlClocksEventReasonSwPowerCap))
            while not self._stop.is_set():
                sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                try:
                    reasons = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    reasons = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if reasons & b else "Not Active" for _, b in bits] + ["nvml"])
                self._stop.wait(0.005)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [x.strip() for x in out.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '', 1).isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '', 1).isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows),
                "source": "nvml" if self.rows and self.rows[0][-1] == "nvml" else "nvidia-smi"}


def cpu_sample(gal, k, flags, budget_s, procs, seed=0, max_queries=None):
    """The reference's loop on uniformly sampled queries of the SAME workload until budget_s is spent: the reference's own
    functions (oracle/_ref or the checkout) when present, else the oracle port.  Returns the sampled ids, the per-query
    records and the timing."""
    from oracle import parallel as OP
    from oracle import ref_loader as RL
    n = gal.patches.shape[0]
    rng = np.random.default_rng(seed)
    order = rng.permutation(n)
    if max_queries:
        order = order[:max_queries]
    kind = "reference" if RL.root() is not None else "port"
    recs, dt, procs = OP.run(gal, order.tolist(), [0, k], flags, procs=procs, impl=kind, chunk=4, budget_s=budget_s)
    return {"ids": np.array([int(x["q"]) for x in recs], dtype=np.int64), "recs": recs, "seconds": dt,
            "queries": len(recs), "kind": kind,
            "pairs_per_s": len(recs) * min(k, n) / dt, "cores": procs}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores -- the unmodified
    utilities/diml.py + evaluation/metrics.py functions driven as evaluation/eval_cvt_diml.py:316-372 drives them
    (oracle/ref_loader.py over oracle/_ref, kind "reference"; the oracle port if that copy is missing, kind "port"),
    one single-threaded worker process per core, each step a bounded sample of the workload.  Rank 0 only."""
    if rank != 0:
        return
    from vitrerank import synth
    from vitrerank.engine import OTParams
    shape, n_default, k, flags, desc = WORKLOADS[args.workload]
    gal = synth.make_named(shape, seed=0, n=args.n or n_default)
    n, c, r = gal.shape
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    procs = os.cpu_count() or 1
    for i in range(args.warmup):
        cpu_sample(gal, k, flags, per_step_budget / 4, procs, seed=99 + i, max_queries=4 * procs)
    res = [cpu_sample(gal, k, flags, per_step_budget, procs, seed=i) for i in range(args.steps)]
    tot_q = sum(x["queries"] for x in res)
    tot_t = sum(x["seconds"] for x in res)
    value = tot_q * min(k, n) / tot_t
    sample = (f"{tot_q} uniformly sampled queries of {n} ({tot_q * min(k, n)} pairs) over {args.steps} steps, "
              f"{tot_t:.1f} s, {procs} single-threaded worker processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": static_config(args.workload, n, c, r, k, [0, k], OTParams.from_flags(**flags).mode),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": procs, "kind": res[0]["kind"], "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "detail": {"queries_per_s": tot_q / tot_t,
                   "note": "bounded sample per step; the reference's own functions on host cores, query-parallel"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sop", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=None, help="override gallery size (debug)")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of reference time for cpu_baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip cpu_baseline / parity (profiling runs)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from vitrerank import _lib, synth
    from vitrerank.engine import OTParams, RerankEngine

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    shape, n_default, k, flags, desc = WORKLOADS[args.workload]
    gal = synth.make_named(shape, seed=0, n=args.n or n_default)
    n, c, r = gal.shape
    params = OTParams.from_flags(**flags)
    truncs = [0, k]
    eng = RerankEngine.get(dev)
    eng.register(gal.patches, gal.centers, gal.rollout, gal.labels)
    kp = max(k, eng.bank["max_num_pos"], 8)
    nq = (n - rank + world - 1) // world           # interleaved shard: queries rank, rank+world, ...
    t_dev = torch.zeros(len(truncs), 8, dtype=torch.float64, device=dev)

    def step(timers=None, keep=False):
        """One pass over this rank's query shard, banks resident in HBM."""
        t_dev.zero_()
        if timers:
            timers[0].record()
        idx, approx = eng.stage0_topk(kp, q_start=rank, q_stride=world, nq=nq)
        if timers:
            timers[1].record()
        score, niter = eng.rerank_scores(idx, k, params, q_start=rank, q_stride=world)
        if timers:
            timers[2].record()
        res = eng.finalize(idx, approx, score, k, truncs, q_start=rank, q_stride=world, tallies=t_dev,
                           want_per_query=keep)
        if world > 1:
            dist.all_reduce(t_dev)                  # the path's only exchange: 16 doubles
        out = t_dev.cpu()                           # device -> host read of the tallies
        if timers:
            timers[3].record()
        if keep:
            return out, niter, (idx, score, res[2])
        return out, niter

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        tallies, niter = step()
    sync_all()
    niter_np = niter.cpu().numpy()
    _lib.take_launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        sync_all()
        t_begin.record()
        for s in range(args.steps):
            tallies, niter = step(evs[s])
        t_end.record()
        sync_all()
    launches = _lib.take_launch_count()
    s0_stats = eng.stage0_stats()          # of the last timed step's first stage (rows redone by the exact fp32 fallback)
    elapsed_ms = t_begin.elapsed_time(t_end)
    t_el = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_el, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t_el.item())
    pairs_per_step = n * min(k, n)
    value = pairs_per_step * args.steps / (elapsed_ms / 1e3)
    s0_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    pf_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    fin_ms = statistics.mean(e[2].elapsed_time(e[3]) for e in evs)

    # ---- end to end through the host-buffer entry ----
    e2e = None
    if not args.no_e2e:
        from vitrerank import distributed as vdist
        pin = gal.pin()
        d2h = len(truncs) * 8 * 8

        def e2e_step():
            # world == 1: vr_evaluate_host (the C-ABI host-buffer entry).  world > 1: every image crosses PCIe once per
            # node (this rank uploads its slices of the patch bank), pipelined NVLink all-gather, tallies all-reduced.
            return vdist.evaluate_host_sharded(eng, pin.patches, pin.centers, pin.rollout, pin.labels, truncs, params)

        for _ in range(2):
            tal_h, h2d = e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            tal_h, h2d = e2e_step()
        sync_all()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        dt = float(t_e.item())
        assert np.allclose(tal_h, tallies.numpy(), rtol=0, atol=1e-9), "end-to-end tallies differ from the resident pass"
        e2e = {"value": pairs_per_step * args.steps / dt, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * dt / args.steps,
               "note": ("vr_evaluate_host: pinned host banks -> HBM in pieces (re-packed as they land), S1..S5, tallies -> host"
                        if world == 1 else
                        "evaluate_host_sharded: each rank uploads 1/W of the patch bank (bytes are per rank), pipelined NVLink "
                        "all-gather + re-pack under stage 0, S1..S5 on its query shard, tallies all-reduced -> host")}
        launches_e2e = _lib.take_launch_count()
        eng.register(gal.patches, gal.centers, gal.rollout, gal.labels)
    else:
        launches_e2e = 0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src, sm_max_mhz = measured_peaks()
    bpp = BYTES_PER_PAIR.get((c, r), c * r * 4 + c * 4 + r * 4)
    pairs_per_launch = nq * min(k, n)
    achieved = pairs_per_launch * bpp / (pf_ms / 1e3) / 1e9
    # FP32 roof: the useful work of the Sinkhorn loop is 2 * R * R FMAs per pair and iteration (two mat-vecs)
    mean_it = float(niter_np.mean())
    fma_per_launch = float(pairs_per_launch) * mean_it * 2.0 * r * r
    fp32_peak = eng.sm_count * FP32_LANES_PER_SM * sm_max_mhz * 1e6 / 1e12      # TFMA/s
    fp32_achieved = fma_per_launch / (pf_ms / 1e3) / 1e12
    traffic, traffic_file = static_traffic(args.workload, pairs_per_launch)
    scale = n / 100.0
    if (c, r) == (128, 49) and k <= 1024:
        kname = "pair_fused_kernel"
    elif params.ot_part > 0.999 and c % 16 == 0 and 20 <= r <= 224:
        kname = "generic_fused_kernel"          # S3 + S4 in one kernel from the operand copy (generic_fused.cu)
    else:
        kname = "generic_sim_mma_kernel + generic_sk_chunk_kernel + generic_finish_kernel"
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": static_config(args.workload, n, c, r, k, truncs, params.mode),
        "detail": {"kp": kp, "parallelism": f"queries sharded x{world}, gallery replicated",
                   "queries_per_s": value / min(k, n),
                   "sinkhorn_iters": {"mean": mean_it, "min": int(niter_np.min()), "max": int(niter_np.max())},
                   "stage_ms": {"stage0_topk": s0_ms, "pair_kernel": pf_ms, "finalize_tally_d2h": fin_ms},
                   "metrics": {"r1": (tallies[:, 0] / scale).tolist(), "rp": (tallies[:, 1] / scale).tolist(),
                               "mapr": (tallies[:, 2] / scale).tolist()},
                   "stage0": dict(s0_stats, path="tcgen05 GEMM + fused select (stage0_mma.cu)" if os.environ.get("VR_STAGE0", "") != "sgemm"
                                  and c == 128 and nq >= 256 and kp <= 256 else "fp32 (stage0_topk.cu)"),
                   "sm_count": eng.sm_count, "pair_transport": os.environ.get("VR_PAIR_TRANSPORT", "global")},
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": (f"static: ncu --set full capture of this workload and launch size, profiles/{traffic_file}"
                                        if traffic is not None else None),
                     "traffic_unit": "bytes per launch (dram read + write)", "peak_source": peak_src,
                     "algorithmic_bytes_per_pair": bpp, "pairs_per_launch": pairs_per_launch,
                     "kernel_ms": pf_ms, "kernel_share_of_step": pf_ms / (elapsed_ms / args.steps),
                     "note": "the candidate gather is the only unavoidable HBM traffic (SURVEY 8d); the kernel is bound by "
                             "FP32 issue / latency in the Sinkhorn loop, see roofline_fp32"},
        "roofline_fp32": {"bound": "fp32", "kernel": kname, "achieved": fp32_achieved, "peak": fp32_peak,
                          "unit": "TFMA/s", "frac": fp32_achieved / fp32_peak,
                          "useful_fma_per_pair_iteration": 2 * r * r, "mean_iterations": mean_it,
                          "peak_source": f"{eng.sm_count} SMs x {FP32_LANES_PER_SM} FP32 lanes x {sm_max_mhz:.0f} MHz"},
        "clocks": clocks.summary(),
        "gpu_launches": int(launches),
    }
    if kname == "generic_fused_kernel":
        # S3 of the generic path on the tensor cores: 3 fp16 MMAs (hi.hi + hi.lo + lo.hi) over ceil(R / 128) x 128 rows and the
        # columns padded to 16 -- ISSUED flops, of which 2 R R C per pair are the fp32 product the reference computes
        tpk, tsrc = tensor_peak()
        mt, rp16 = (r + 127) // 128, (r + 15) // 16 * 16
        issued = 3.0 * 2.0 * mt * 128 * rp16 * c * pairs_per_launch
        line["roofline_tensor"] = {"bound": "tensor", "kernel": kname, "achieved": issued / (pf_ms / 1e3) / 1e12, "peak": tpk,
                                   "unit": "TFLOP/s", "frac": issued / (pf_ms / 1e3) / 1e12 / tpk, "peak_source": tsrc,
                                   "useful_tflops": 2.0 * r * r * c * pairs_per_launch / (pf_ms / 1e3) / 1e12,
                                   "note": "the tensor phase is ~35 % of the kernel's time (profiles/r2_ncu_generic_fused.md); the "
                                           "Sinkhorn chains (shared-memory load pipe) take ~42 %"}
    if line["detail"]["stage0"]["path"].startswith("tcgen05"):
        # first stage: two GEMM passes of nq x n x 128 with three fp16 MMAs per product term (hi.hi + lo.hi + hi.lo); the time is
        # the WHOLE stage (pack, both passes, threshold, final select), so this is a lower bound of the GEMM kernels' own rate
        tpk, tsrc = tensor_peak()
        tf = 2.0 * 3.0 * 2.0 * float(nq) * float(n) * float(c) / (s0_ms / 1e3) / 1e12
        line["roofline_stage0"] = {"bound": "tensor", "kernel": "stage0_mma_kernel (x2) + pack / thresh / final", "achieved": tf,
                                   "peak": tpk, "unit": "TFLOP/s", "frac": tf / tpk, "peak_source": tsrc, "stage_ms": s0_ms,
                                   "note": "issued fp16 flops of both passes over the time of the whole stage"}
    if e2e:
        line["e2e"] = e2e
        line["gpu_launches_e2e"] = int(launches_e2e)
    # ---- CPU baseline + parity counters: the reference on this box's host cores, bounded sample ----
    if world > 1:
        dist.destroy_process_group()
    if world == 1 and not args.no_cpu:
        from oracle import parity as PAR
        tallies, niter, (idx, score, per_q) = step(keep=True)
        torch.cuda.synchronize(dev)
        procs = os.cpu_count() or 1
        base = cpu_sample(gal, k, flags, args.cpu_budget, procs)
        one = cpu_sample(gal, k, flags, max(3.0, args.cpu_budget / 3), 1, seed=1)
        line["cpu_baseline"] = {
            "value": base["pairs_per_s"], "unit": "pairs/s", "cores": base["cores"], "kind": base["kind"],
            "sample": f"{base['queries']} uniformly sampled queries of {n} ({base['queries'] * min(k, n)} pairs), "
                      f"{base['seconds']:.1f} s, {base['cores']} single-threaded worker processes",
            "one_thread": {"value": one["pairs_per_s"], "unit": "pairs/s", "cores": 1,
                           "sample": f"{one['queries']} queries, {one['seconds']:.1f} s"}}
        # parity of the same sample, un-forced.  The reference's functions expose neither iteration counts nor err traces:
        # those counters come from the oracle port on the same queries (bit-identical scores, tests/test_oracle_golden.py)
        from oracle import parallel as OP
        ids = base["ids"][:min(len(base["ids"]), 256)]
        dumps, _, _ = OP.run(gal, ids.tolist(), truncs, flags, procs=procs, impl="port", chunk=4)
        ids_t = torch.as_tensor(ids, device=dev, dtype=torch.long)
        par = PAR.compare(dumps, idx[ids_t].cpu().numpy(), score[ids_t].cpu().numpy(), niter[ids_t].cpu().numpy(), k,
                          trunc_nums=truncs, per_query=per_q[ids_t].cpu().numpy())
        if base["kind"] == "reference":   # scores of the REAL functions against the port on the sample: must be identical
            par["reference_vs_port_score_mismatches"] = int(sum(
                0 if torch.equal(a["score"], b["score"]) else 1 for a, b in zip(base["recs"][:len(dumps)], dumps)))
        par["sample"] = f"{len(dumps)} of the cpu_baseline queries, compared un-forced (oracle keeps its iteration counts)"
        line["parity"] = par
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
