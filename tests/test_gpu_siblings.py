"""The sibling callers of the rerank loop through their drop-in modules: evaluation/eval_diml.py
(ResNet-50), eval_attn_diml.py (DeiT, 14 x 14 grid), eval_swin_diml.py, and the query != gallery core of
training_tools/val.py (MSLS).  Stub backbones with the reference's output contract -- model(img) -> (out, (enc, feat)),
model.model.head / last_linear -- feed the real embedding + loop code; the oracle loop on the very banks they embedded is
the check (un-forced: a query one Sinkhorn iteration apart may move a tally by at most 100 / N)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import rerank_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Stub(nn.Module):
    """Token backbone: conv patch embedding -> tokens [B, L, c_in]; out = head(mean token)."""

    def __init__(self, side, c_in=48, c=128, maps=False, head=True, seed=0):
        super().__init__()
        torch.manual_seed(seed)
        self.patch = nn.Conv2d(3, c_in, 2, stride=2)
        self.model = nn.Module()
        if head:
            self.model.head = nn.Linear(c_in, c)
        else:
            self.model.last_linear = nn.Linear(c_in, c)
        self.maps, self.side = maps, side

        class P:
            dataset, arch, not_pretrained = "stub", "stub", True
        self.pars = P()

    def forward(self, img):
        f = torch.tanh(self.patch(img))                       # [B, c_in, side, side]
        tok = f.flatten(2).permute(0, 2, 1)                   # [B, L, c_in]
        proj = self.model.head if hasattr(self.model, "head") else self.model.last_linear
        out = proj(tok.mean(1))
        return out, (tok.mean(1), f if self.maps else tok)


class _DS:
    def __init__(self, n, side, classes, seed):
        g = torch.Generator().manual_seed(seed)
        self.labels = torch.arange(n) % classes
        proto = torch.randn(classes, 3, 2 * side, 2 * side, generator=g)
        self.img = proto[self.labels] + 0.7 * torch.randn(n, 3, 2 * side, 2 * side, generator=g)

    def __len__(self):
        return len(self.labels)


class _Loader:
    def __init__(self, ds, bs):
        self.ds, self.bs = ds, bs

    def __iter__(self):
        for lo in range(0, len(self.ds), self.bs):
            hi = min(len(self.ds), lo + self.bs)
            yield self.ds.labels[lo:hi], self.ds.img[lo:hi], torch.arange(lo, hi)


def _check(data, truncs, flags, n):
    from vitrerank.engine import RerankEngine
    b = RerankEngine.get(DEV).bank
    cpu = [b[k].cpu() for k in ("patches", "centers")] + [None, b["labels"].cpu()]
    ref = O.evaluate_banks(*cpu, trunc_nums=truncs, dump=True, **flags)
    # the drop-in already ran; its iteration counts are not returned by evaluate(): allow one borderline query per 100
    slack = max(1, n // 100) * 100.0 / n
    for key in ("r1", "rp", "mapr"):
        np.testing.assert_allclose(data[key], ref[key], rtol=0, atol=slack + 1e-9)
    assert data["r1"][-1] > 5.0, "degenerate stub data"
    return ref


@pytest.mark.parametrize("grid,flags", [(14, dict(use_soft=True, ot_part=1.0, temperature=1.0)),
                                        (7, dict(use_minus=True, ot_part=0.6, use_cls_token=True)),
                                        (7, dict(use_inverse=True, temperature=0.1, ot_part=1.0))])
def test_eval_attn_diml(grid, flags, capsys):
    """DeiT form: 14 x 14 tokens.  grid 14 keeps R = 196 (generic solver); grid 7 pools in the ingest kernel (fused kernel)."""
    import evaluation.eval_attn_diml as E
    n = 120
    ds, model = _DS(n, 14, 10, 1), _Stub(14).to(DEV)
    data = E.evaluate(model, ds, _Loader(ds, 32), False, [0, 20], grid_size=grid, **flags)
    from vitrerank.engine import RerankEngine
    assert RerankEngine.get(DEV).bank["r"] == grid * grid
    assert "Now rank-1 acc=" in capsys.readouterr().out
    _check(data, [0, 20], flags, n)


def test_eval_swin_diml():
    import evaluation.eval_swin_diml as E
    n = 150
    ds, model = _DS(n, 7, 12, 2), _Stub(7).to(DEV)
    flags = dict(use_uniform=False, use_inverse=True, temperature=0.5, use_cls_token=True, ot_part=1.0)
    data = E.evaluate(model, ds, _Loader(ds, 64), False, [0, 10, 50], grid_size=7, **flags)
    _check(data, [0, 10, 50], flags, n)


@pytest.mark.parametrize("head", [True, False])
def test_eval_diml_resnet_form(head):
    """ResNet form: [B, C, H, W] maps through last_linear (no head) or tokens through head; positional flags of :170."""
    import evaluation.eval_diml as E
    n = 140
    ds = _DS(n, 7, 10, 3)
    model = _Stub(7, maps=not head, head=head, seed=4).to(DEV)
    data = E.evaluate(model, ds, _Loader(ds, 50), True, [0, 30], False, 7, True, 0.1, True)
    _check(data, [0, 30], dict(use_uniform=False, use_inverse=True, temperature=0.1, use_cls_token=True, ot_part=1.0), n)


def test_msls_rerank_city_vs_reference_loop():
    """training_tools/val.py:159-199 for one city: 60 queries against 400 database images."""
    from training_tools.val import rerank_city
    from vitrerank import synth
    g = synth.make_gallery(460, 128, 49, classes=23, seed=5, sigma=0.6)
    q_dense, db_dense = g.patches[:60], g.patches[60:]
    # val() keeps the model's global embeddings as they are (NetVLAD-style, unit norm here); un-normalised also works
    q_glob, db_glob = g.centers[:60] * 1.0, g.centers[60:] * 1.0
    T = 30
    preds, final, score = rerank_city(q_glob, db_glob, q_dense, db_dense, trunc_nums=[0, T], device=DEV, return_scores=True)
    preds, final, score = preds.cpu(), final.cpu(), score.cpu()
    assert preds.shape == (60, 100) and final.shape == (60, 100)
    flips = 0
    for i in range(60):
        approx = torch.einsum('c,mc->m', q_glob[i], db_glob)                       # [:173]
        order = torch.argsort(approx, descending=True, stable=True)[:100]           # [:174]
        assert set(preds[i].tolist()) == set(order.tolist())
        top = preds[i, :T]
        sim, _, (n_iter, errs) = O.structural_similarity(q_dense[i], q_glob[i], db_dense[top], db_glob[top], "minus", ot_temp=0.05,
                                                         temperature=0.1, use_cls_token=True, ot_part=1.0, trace=True)
        rel = (score[i] - sim).abs() / sim.abs().clamp_min(1e-12)
        if rel.max() < 1e-4:
            rank = torch.argsort(sim + approx[top], descending=True, stable=True)   # [:197]
            want = torch.cat([top[rank], preds[i, T:]])                             # [:198-199]
            total = (sim + approx[top])[rank]
            if (total[:-1] - total[1:]).min() > 1e-6:                               # no near tie: the order is decided
                assert torch.equal(final[i], want), i
        else:
            # The stop moved.  Legitimate only where the oracle's err at its decisive iteration is within 2 % of the threshold
            # (there the test sits at the fp32 noise floor, and once a stop is missed the err hovers around
            # the threshold, so the count can move by MORE than one -- measured 47 against 15 on query 51, scores 2.4 % apart).
            flips += 1
            assert abs(float(errs[n_iter - 1]) - 0.1) <= 0.02 * 0.1, (i, n_iter, errs[-3:])
            assert rel.max() < 5e-2
    assert flips <= 3
