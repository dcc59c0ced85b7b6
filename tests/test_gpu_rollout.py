"""The attention-rollout producer on the device (csrc/rollout.cu) against the torch statements of
evaluation/eval_cvt_diml.py:54-146 run on the CPU -- which tests/test_host_logic.py::test_attention_rollout_matches_reference
pins bit for bit to the reference's own functions.  filter + resize of a block must be BIT-IDENTICAL (head fusion, the
exact k smallest entries per image, the batch-union zeroing, AdaptiveAvgPool2d's arithmetic on both axes); the chain is
compared at 1e-6 (the CPU bmm's summation order is MKL's)."""
import sys
import os

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from vitrerank.engine import RerankEngine
    return RerankEngine.get("cuda:0")


def torch_block(probs, stage, grid, fusion="min"):
    import evaluation.eval_cvt_diml as E
    a = E.filter_attention_map(probs.clone(), discard_ratio=0.1, head_fusion=fusion)
    return E.resize_attn_map(a, nn.AdaptiveAvgPool2d((grid, grid)), stage, grid).contiguous()


@pytest.mark.parametrize("b,heads,ht,wt,stage,grid", [
    (3, 6, 197, 197, 2, 7),      # the ViT / CvT stage-2 form: cls row and column dropped, 14 x 14 -> 7 x 7 on both axes
    (2, 1, 784, 196, 0, 7),      # a convolutional-projection stage: 28 x 28 queries, 14 x 14 keys, one head
    (2, 3, 100, 64, 1, 7),       # 10 x 10 and 8 x 8 grids: AdaptiveAvgPool2d windows of unequal size
    (4, 2, 50, 50, 2, 7),        # already 7 x 7 after the cls drop: no pooling at all
    (2, 4, 256, 49, 1, 7),       # keys already 7 x 7, queries 16 x 16
    (5, 2, 65, 65, 2, 4),        # another target grid
])
def test_block_equals_the_torch_statements(eng, b, heads, ht, wt, stage, grid):
    g = torch.Generator().manual_seed(ht * 31 + wt)
    probs = torch.softmax(torch.randn(b, heads, ht, wt, generator=g) * 2.0, dim=-1)
    ref = torch_block(probs, stage, grid)
    got = eng.rollout_block(probs.cuda(), drop_cls=(stage == 2), grid=grid, discard_ratio=0.1, head_fusion="min").cpu()
    assert got.shape == ref.shape
    assert torch.equal(got, ref), f"max |diff| {(got - ref).abs().max().item():.3e}, {(got != ref).sum().item()} entries"
    got_max = eng.rollout_block(probs.cuda(), drop_cls=(stage == 2), grid=grid, discard_ratio=0.1, head_fusion="max").cpu()
    assert torch.equal(got_max, torch_block(probs, stage, grid, fusion="max"))


def test_batch_union_of_discarded_coordinates(eng):
    """A coordinate discarded in ANY image of the batch is zeroed in EVERY image (the reference's `new_attn[:, iy, ix] = 0`):
    the same image alone and inside a batch must differ exactly there."""
    g = torch.Generator().manual_seed(5)
    probs = torch.softmax(torch.randn(3, 2, 50, 50, generator=g), dim=-1)
    alone = eng.rollout_block(probs[:1].cuda(), drop_cls=True, grid=7).cpu()[0]      # 49 x 49: no pooling, entries visible
    batch = eng.rollout_block(probs.cuda(), drop_cls=True, grid=7).cpu()
    k = int(50 * 50 * 0.1)                                    # (chosen on the whole map; the cls row / column go afterwards)
    full = probs[0].min(dim=0).values.reshape(-1)
    inside = torch.zeros(50, 50, dtype=torch.bool)
    inside.view(-1)[full.topk(k, largest=False).indices] = True
    assert int((alone == 0).sum()) == int(inside[1:, 1:].sum())
    zeros = (batch == 0)
    assert torch.equal(zeros[0], zeros[1]) and torch.equal(zeros[0], zeros[2])
    assert int((alone == 0).sum()) < int(zeros[0].sum()) <= 3 * k
    assert bool(zeros[0][alone == 0].all())
    keep = ~zeros[0]
    assert torch.equal(batch[0][keep], alone[keep])


def test_ties_at_the_threshold_take_the_lowest_indices(eng):
    """Quantised attention: hundreds of entries equal the k-th smallest value.  Exactly k entries are discarded -- every
    entry below the threshold and, of those equal to it, the first in index order."""
    g = torch.Generator().manual_seed(9)
    probs = (torch.rand(1, 1, 50, 50, generator=g) * 20).floor() / 20 + 0.05
    probs = probs[:, :, :49, :49].contiguous()                # (no cls token: the 49 x 49 map is filtered and returned as it is)
    out = eng.rollout_block(probs.cuda(), drop_cls=False, grid=7).cpu()[0].reshape(-1)
    flat = probs[0, 0].reshape(-1)
    k = int(flat.numel() * 0.1)
    tau = flat.sort().values[k - 1]
    zero = out == 0
    assert int(zero.sum()) == k
    assert bool(zero[flat < tau].all()) and not bool(zero[flat > tau].any())
    eq = (flat == tau).nonzero().flatten()
    take = k - int((flat < tau).sum())
    assert 0 < take < eq.numel()
    assert bool(zero[eq[:take]].all()) and not bool(zero[eq[take:]].any())


@pytest.mark.parametrize("use_res", [True, False])
def test_chain(eng, use_res):
    g = torch.Generator().manual_seed(3)
    mats = torch.rand(13, 6, 49, 49, generator=g) * 0.05
    m = mats.clone()
    if use_res:
        m = m + torch.eye(49)
        m = m / m.sum(dim=-1).unsqueeze(-1)
    ref = [m[0]]
    for j in range(1, 13):
        ref.append(torch.bmm(m[j], ref[j - 1]))
    got = eng.rollout_chain(mats.cuda(), use_res=use_res).cpu()
    assert torch.equal(got[0], ref[0])                      # identity + normalisation: ATen's row sums, bit for bit
    for j in range(13):
        assert torch.allclose(got[j], ref[j], rtol=1e-5, atol=1e-9), j
    assert torch.allclose(got[-1].mean(1), ref[-1].mean(1), rtol=1e-5, atol=1e-9)


class _Blk:
    def __init__(self, probs):
        self._probs = [probs]


class _Stage:
    def __init__(self, blocks):
        self.blocks = blocks


class _Model:
    """What get_attention_rollout touches of a CvT: both_forward and stage{0,1,2}.blocks[i]._probs[0]."""

    def __init__(self, probs_by_stage, device):
        for si in range(3):
            setattr(self, f"stage{si}", _Stage([_Blk(p.to(device)) for p in probs_by_stage[si]]))

    def both_forward(self, x):
        return None, None


def test_get_attention_rollout_on_the_device():
    """The drop-in's get_attention_rollout with the attention of every block on the B200 (the library's kernels) against the
    same attention on the CPU (the reference's torch statements): a CvT-13-like stack of 1 + 2 + 10 blocks."""
    import evaluation.eval_cvt_diml as E
    from vitrerank import _lib
    g = torch.Generator().manual_seed(11)
    sm = lambda *shape: torch.softmax(torch.randn(*shape, generator=g) * 1.5, dim=-1)
    probs = {0: [sm(4, 1, 784, 196)], 1: [sm(4, 3, 196, 49) for _ in range(2)], 2: [sm(4, 6, 197, 50) for _ in range(10)]}
    ref = E.get_attention_rollout(_Model(probs, "cpu"), None)
    _lib.take_launch_count()
    got = E.get_attention_rollout(_Model(probs, "cuda"), None)
    assert _lib.take_launch_count() > 0                      # the library's kernels ran, not torch's
    assert len(got) == len(ref) == 13
    assert got[0].is_cuda and torch.equal(got[0].cpu(), ref[0])
    for a, b in zip(got, ref):
        assert torch.allclose(a.cpu(), b, rtol=1e-5, atol=1e-9)
    assert torch.allclose(got[-1].mean(1).cpu(), ref[-1].mean(1), rtol=1e-5, atol=1e-9)
