"""S1 on the tensor cores (csrc/stage0_mma.cu) against the fp32 FMA-chain kernels (csrc/stage0_topk.cu), whose lists the
oracle tests pin to the reference (tests/test_gpu_parity.py::test_stage0_topk): first-stage shortlists, their order and their
fp32 scores must be BIT-IDENTICAL (north_star: "first-stage top-K index sets are bit-exact"), on every shape, shard form and
on the inputs built to defeat the shortlist (exact duplicates, values the fp16 split cannot hold)."""
import os

import pytest
import torch

from vitrerank import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from vitrerank.engine import RerankEngine
    return RerankEngine.get("cuda:0")


def both_paths(eng, kp, **kw):
    """(idx, score, stats) of the tensor-core path and (idx, score) of the fp32 path on the same call."""
    os.environ.pop("VR_STAGE0", None)
    i1, s1 = eng.stage0_topk(kp, **kw)
    st = eng.stage0_stats()
    os.environ["VR_STAGE0"] = "sgemm"
    try:
        i0, s0 = eng.stage0_topk(kp, **kw)
        st0 = eng.stage0_stats()
    finally:
        os.environ.pop("VR_STAGE0", None)
    assert st0 == {"fallback_rows": 0, "overflow_rows": 0, "unsplittable": 0}
    return (i1, s1, st), (i0, s0)


def centers_only(n, c=128, classes=100, seed=0, sigma=0.6):
    """Class-structured unit-norm centres without the patch bank (stage 0 reads nothing else)."""
    gen = torch.Generator().manual_seed(seed)
    labels = synth.make_labels(n, classes, gen)
    proto = torch.randn(int(labels.max()) + 1, c, generator=gen)
    x = proto[labels] + sigma * torch.randn(n, c, generator=gen)
    return torch.nn.functional.normalize(x, dim=1), labels


def register_centers(eng, centers, labels):
    n = centers.shape[0]
    eng.register(torch.zeros(n, 128, 1), centers, None, labels)   # a placeholder patch bank: stage 0 never touches it


@pytest.mark.parametrize("n,kp,classes,sigma", [(3000, 100, 60, 0.6), (8131, 100, 98, 0.6), (5000, 256, 40, 0.3),
                                                (4096, 8, 500, 1.0), (2049, 33, 20, 0.6)])
def test_tensor_core_lists_equal_fp32_lists(eng, n, kp, classes, sigma):
    centers, labels = centers_only(n, classes=classes, seed=n + kp, sigma=sigma)
    register_centers(eng, centers, labels)
    (i1, s1, st), (i0, s0) = both_paths(eng, kp)
    assert torch.equal(i1, i0), f"shortlists differ in {(i1 != i0).any(dim=1).sum().item()} rows"
    assert torch.equal(s1, s0)
    assert st["unsplittable"] == 0 and st["overflow_rows"] <= n // 1000, st   # (class blocks can overflow a row's candidate lists)
    assert st["fallback_rows"] <= n // 100, st


def test_shards_and_explicit_queries(eng):
    n, kp = 6000, 100
    centers, labels = centers_only(n, classes=80, seed=5)
    register_centers(eng, centers, labels)
    (i1, s1, _), (i0, s0) = both_paths(eng, kp)
    assert torch.equal(i1, i0) and torch.equal(s1, s0)
    (j1, t1, _), (j0, t0) = both_paths(eng, kp, q_start=3, q_stride=8)            # 750 queries of an 8-way interleaved shard
    assert torch.equal(j1, j0) and torch.equal(t1, t0) and torch.equal(j1, i1[3::8])
    q = centers[1000:1900].clone()
    (k1, u1, _), (k0, u0) = both_paths(eng, kp, q_centers=q, self_idx=torch.arange(1000, 1900))
    assert torch.equal(k1, k0) and torch.equal(u1, u0) and torch.equal(k1, i1[1000:1900])
    (m1, v1, _), (m0, v0) = both_paths(eng, kp, q_centers=q)                       # no self mask: the query itself ranks first
    assert torch.equal(m1, m0) and torch.equal(v1, v0)
    assert (m1[:, 0].cpu() == torch.arange(1000, 1900)).all()


def test_exact_duplicates_take_the_fallback(eng):
    """Every image has 150 exact copies: a shortlist of 100 cuts through a 150-way exact tie, the acceptance test must fail
    and the exact fp32 fallback must return the fp32 path's lists (ties -> lower index first)."""
    n, kp = 4500, 100
    base, _ = centers_only(30, classes=5, seed=9)
    centers = base.repeat_interleave(150, dim=0).contiguous()
    register_centers(eng, centers, torch.arange(n) // 150)
    (i1, s1, st), (i0, s0) = both_paths(eng, kp)
    assert torch.equal(i1, i0) and torch.equal(s1, s0)
    assert st["fallback_rows"] == n, st


def test_values_the_split_cannot_hold(eng):
    n, kp = 3000, 50
    centers, labels = centers_only(n, classes=30, seed=11)
    big = centers * 30000.0                     # 64 x overflows fp16: flagged by the pack kernel, every row redone exactly
    register_centers(eng, big, labels)
    (i1, s1, st), (i0, s0) = both_paths(eng, kp)
    assert torch.equal(i1, i0) and torch.equal(s1, s0)
    assert st["unsplittable"] == 1 and st["fallback_rows"] == n
    scaled = centers * 3.0                      # fine for the split; the acceptance bound scales with the norms
    register_centers(eng, scaled, labels)
    (i1, s1, st), (i0, s0) = both_paths(eng, kp)
    assert torch.equal(i1, i0) and torch.equal(s1, s0)
    assert st["unsplittable"] == 0 and st["fallback_rows"] <= 30


def test_sop_shape_full(eng):
    """BASELINE configs[2] shape: all 60,502 queries, kp = 100."""
    n, kp = 60502, 100
    centers, labels = centers_only(n, classes=11316, seed=0)
    register_centers(eng, centers, labels)
    (i1, s1, st), (i0, s0) = both_paths(eng, kp)
    assert torch.equal(i1, i0) and torch.equal(s1, s0)
    assert st["overflow_rows"] == 0 and st["fallback_rows"] <= 60, st
    print("sop stage-0 stats", st)


def test_few_queries_against_a_large_gallery(eng):
    """300 queries against 60,502 images: three row blocks, so the gallery is cut into ~29 splits per block (1,856 segment
    maxima per row: the threshold kernel's long-row path) and every row's candidates arrive in ~116 sub-lists."""
    n, kp = 60502, 100
    centers, labels = centers_only(n, classes=11316, seed=1)
    register_centers(eng, centers, labels)
    (i1, s1, st), (i0, s0) = both_paths(eng, kp, q_start=5, q_stride=200, nq=300)
    assert torch.equal(i1, i0) and torch.equal(s1, s0)
    assert st["overflow_rows"] == 0 and st["fallback_rows"] <= 3, st
