"""UN-FORCED full-pass parity on every BASELINE.json config shape: the CUDA path against the oracle run exactly as the
reference runs (its own stop test decides every iteration count; nothing is re-run at the CUDA counts).

Every case records its counters (oracle/parity.py) -- queries whose Sinkhorn count differs, pairs beyond the 1e-4
score gate, first-stage boundary near-ties, per-query metric mismatches, tally deltas -- prints them and appends them
to gpurun_out/parity_r2.jsonl; profiles/r2_parity.md is that file, committed.

What is asserted (north_star gates) and what is only counted:
  * first-stage top-K sets: identical, except at boundaries closer than 1e-6 (counted);
  * iteration counts: equal or one apart; every difference must sit where the oracle's err is within 2 % of the
    threshold at the decisive iteration (the stop test sits at the fp32 noise floor and the
    patch similarity comes from another summation order than MKL's); at most 4 % of the queries;
  * per-pair scores: within 1e-4 relative for every query with an equal count -- no exception; for the queries one
    iteration apart the un-forced difference is REPORTED (measured up to 2e-3: one more Sinkhorn iteration of the
    reference itself moves its scores that much) and only sanity-bounded by 1e-2;
  * per-query r1 / RP / MAP@R: bit-identical (==) for every query with an equal count, except where two first-stage
    scores within 1e-6 of each other swap places in the shortlist (counted); tallies reported.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np
import pytest
import torch

from oracle import parallel as OP
from oracle import parity as PAR
from oracle import ref_loader as RL
from oracle import rerank_oracle as O
from vitrerank import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vit-reranking_b200")


@pytest.fixture(scope="module")
def eng():
    from vitrerank.engine import RerankEngine
    return RerankEngine.get("cuda:0")


def record(name, counters, extra=None):
    row = dict(case=name, **counters)
    if extra:
        row.update(extra)
    print("PARITY", json.dumps(row))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_r2.jsonl"), "a") as f:
            f.write(json.dumps(row) + "\n")
    except OSError:
        pass


def check(c, max_flip_frac=0.04):
    assert c["stage0_set_mismatch_beyond_tie"] == 0, c
    assert c["niter_off_by_more"] == 0, c
    assert c["flips_outside_2pct_band"] == 0, c
    assert c["niter_off_by_one"] <= max(1, int(max_flip_frac * c["queries"])), c
    assert c["pairs_over_1e-4_in_equal_niter_queries"] == 0, c
    assert c["max_rel_err"] < 1e-2, c
    assert c["stage0_order_differs_beyond_tie"] == 0, c
    assert c["metric_mismatch_unexplained"] == 0, c


def run_case(eng, g, k, flags, ids, truncs=None):
    """CUDA pass over the strided query subset `ids` (a range) and the un-forced oracle on the same queries."""
    from vitrerank.engine import OTParams
    truncs = truncs or [0, k]
    n = g.patches.shape[0]
    p = OTParams.from_flags(**flags)
    eng.register(g.patches, g.centers, g.rollout, g.labels)
    kp = max(k, eng.bank["max_num_pos"], 8)
    q0, qs, nq = ids.start, ids.step, len(ids)
    idx, approx = eng.stage0_topk(kp, q_start=q0, q_stride=qs, nq=nq)
    score, niter = eng.rerank_scores(idx, k, p, q_start=q0, q_stride=qs)
    tal, _, pq = eng.finalize(idx, approx, score, k, truncs, q_start=q0, q_stride=qs, want_per_query=True)
    torch.cuda.synchronize()
    t0 = time.time()
    dumps, dt, procs = OP.run(g, list(ids), truncs, flags)
    assert len(dumps) == nq
    c = PAR.compare(dumps, idx.cpu().numpy(), score.cpu().numpy(), niter.cpu().numpy(), k, trunc_nums=truncs,
                    per_query=pq.cpu().numpy())
    scale = n / 100.0
    ref_tal = np.array([[sum(d["metrics"][t][j] for d in dumps) for j in range(3)] for t in truncs])
    got_tal = tal.cpu().numpy()[:, :3]
    c["abs_delta_r1_rp_mapr_percent"] = (np.abs(got_tal - ref_tal) / scale).tolist()
    c["tallies_identical"] = bool((got_tal == ref_tal).all())
    return c, dict(n=n, oracle_s=round(time.time() - t0, 1), oracle_procs=procs)


@pytest.mark.parametrize("name", ["cars196", "cub200"])
def test_full_pass_unforced(eng, name):
    """BASELINE configs[1] (Cars196, 8,131 images) and configs[0] (CUB-200, 5,924): EVERY query, K = 100, rollout."""
    g = synth.make_named(name, seed=0)
    n = g.patches.shape[0]
    c, extra = run_case(eng, g, 100, dict(use_rollout=True, ot_part=1.0), range(0, n, 1))
    record(f"{name}_full_k100_rollout", c, extra)
    check(c)


def test_sop_sample_unforced(eng):
    """BASELINE configs[2]: SOP shape (60,502 images), K = 100, 1,009 uniformly strided queries."""
    g = synth.make_named("sop", seed=0)
    n = g.patches.shape[0]
    ids = range(7, n, 60)[:1009]
    c, extra = run_case(eng, g, 100, dict(use_rollout=True, ot_part=1.0), ids)
    record("sop_sample_k100_rollout", c, extra)
    check(c)
    # BASELINE configs[3]: K = 1000, calc_similarity + use_inverse (T = 0.1), 101 strided queries of the same gallery
    ids = range(11, n, 599)[:101]
    c, extra = run_case(eng, g, 1000, dict(use_inverse=True, temperature=0.1, use_cls_token=True, ot_part=1.0), ids)
    record("sop_sample_k1000_inverse", c, extra)
    check(c, max_flip_frac=0.08)


def test_vitb16_shape_unforced(eng):
    """BASELINE configs[4]: ViT-B/16 shape, R = 196 patches, C = 768, on a small gallery (K = 24)."""
    g = synth.make_gallery(96, 768, 196, classes=6, seed=5, sigma=0.6)
    c, extra = run_case(eng, g, 24, dict(use_rollout=True, ot_part=1.0), range(0, 96, 1))
    record("vitb16_c768_r196_k24_rollout", c, extra)
    check(c, max_flip_frac=0.10)


def test_partial_ot_and_minus_unforced(eng):
    """The reference's live SOP script family: --use_minus without rollout, --ot_part sweeps
    (scripts/diml/test_diml_cvt_sop.sh:9-14, scripts/diml/test_diml_cvt.sh:34-71)."""
    g = synth.make_gallery(1200, 128, 49, classes=150, seed=9, sigma=0.6)
    for part in (0.3, 0.9):
        c, extra = run_case(eng, g, 100, dict(use_minus=True, ot_part=part), range(0, 1200, 4))
        record(f"n1200_k100_minus_part{part}", c, extra)
        check(c, max_flip_frac=0.08)
    c, extra = run_case(eng, g, 100, dict(use_rollout=True, ot_part=0.5), range(1, 1200, 4))
    record("n1200_k100_rollout_part0.5", c, extra)
    check(c, max_flip_frac=0.08)


@pytest.mark.skipif(RL.root() is None, reason="neither /root/reference nor oracle/_ref (python oracle/make_ref.py) present")
def test_reference_caller_unchanged_on_gpu(tmp_path):
    """test_diml_cvt.py of the reference, byte for byte, with the north-star flags, on the B200 through the drop-in
    packages (shims only supply what the checkout lacks: datasets/, a CvT stub, imp, matplotlib).  The banks it embeds
    are kept (bank file), and its printed / CSV results are checked against the REAL reference functions run on
    those banks on the host."""
    ref = RL.root()
    bank = str(tmp_path / "banks.vrbank")
    env = dict(os.environ)
    env["PYTHONSAFEPATH"] = "1"
    env["VITRERANK_SHIM_N"] = "192"
    env["VITRERANK_BANK_CACHE"] = bank
    env["PYTHONPATH"] = os.pathsep.join([PKG, os.path.join(PKG, "shims", "override"), ref,
                                         os.path.join(PKG, "shims", "fallback")])
    args = ["--dataset", "cub200", "--group", "t", "--arch", "cvt_13_normalize", "--embed_dim", "128", "--bs", "16",
            "--samples_per_class", "2", "--not_pretrained", "--use_ot", "--use_inverse", "--use_rollout",
            "--grid_size", "7", "--temperature", "0.1", "--ot_part", "1.0", "--source_path", "/tmp", "--kernels", "0"]
    r = subprocess.run([sys.executable, os.path.join(ref, "test_diml_cvt.py")] + args, cwd=str(tmp_path), env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "Now rank-1 acc=" in r.stdout
    import pandas as pd
    df = pd.read_csv(tmp_path / "test_results" / "test_diml_cub200.csv")
    assert list(df.columns[1:]) == ["method", "r1", "rp", "mapr"] and len(df) == 2      # test_diml_cvt.py:125-161
    from vitrerank import bankfile
    patches, centers, rollout, labels = [t.clone() if t is not None else None for t in bankfile.load(bank)]
    assert patches.shape == (192, 128, 49) and rollout is not None
    out = RL.reference_loop(patches, centers, rollout, labels, [0, 100], use_rollout=True, use_inverse=True,
                            temperature=0.1, ot_part=1.0)
    got = {k: df[k].to_numpy() for k in ("r1", "rp", "mapr")}
    row = dict(case="test_diml_cvt_unchanged_n192",
               delta={k: (got[k] - np.array(out[k])).tolist() for k in got}, reference=ref)
    record(row.pop("case"), row)
    assert (got["r1"] == np.array(out["r1"])).all()
    np.testing.assert_allclose(got["rp"], out["rp"], rtol=0, atol=0.6)       # one borderline stop may move one query
    np.testing.assert_allclose(got["mapr"], out["mapr"], rtol=0, atol=0.6)
