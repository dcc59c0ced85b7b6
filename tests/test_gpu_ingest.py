"""Bank ingest (vr_bank_ingest: csrc/pair_fused.cu bank_ingest_kernel) against the reference's own bank construction
(evaluation/eval_cvt_diml.py:269-278,304-305) run with torch on the CPU: head-projected tokens -> permute -> AdaptiveAvgPool2d ->
F.normalize(dim=1), F.normalize of the global embeddings.  fp32 elementwise work: BIT-EXACT; and the operand planes the
kernel writes on the side must be the ones vr_bank_prepare derives from the finished bank (scores bit-identical)."""
import pytest
import torch

from vitrerank import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from vitrerank.engine import RerankEngine
    return RerankEngine.get("cuda:0")


def reference_banks(tokens, centers_raw, grid):
    """eval_cvt_diml.py:269-278 + :304-305 on CPU tensors (tokens = model.model.head(no_avg_feat): [B, L, C])."""
    x = tokens.permute(0, 2, 1)
    side = int(x.size(-1) ** 0.5)
    x = x.reshape(x.size(0), -1, side, side).contiguous()     # NCHW order pins the pooling's row-major block sum
    if x.size(-1) != grid:
        x = torch.nn.AdaptiveAvgPool2d(grid)(x)
    x = x.reshape(x.size(0), x.size(1), -1)
    # torch.cat at :299 leaves the bank contiguous [N, C, R]: the norm over dim=1 is then ATen's strided (sequential) reduction;
    # on the permuted view the batches still are, dim=1 would be the unit-stride axis and take the vectorised order instead
    x = torch.cat([x.contiguous()], dim=0)
    return (torch.nn.functional.normalize(x, p=2, dim=1), torch.nn.functional.normalize(centers_raw, p=2, dim=1))


@pytest.mark.parametrize("n,c,side,grid", [(70, 128, 14, 7), (33, 128, 7, 7), (9, 768, 14, 14), (20, 64, 8, 4), (5, 100, 12, 4)])
def test_ingest_bit_exact(eng, n, c, side, grid):
    gen = torch.Generator().manual_seed(n + c)
    tokens = torch.randn(n, side * side, c, generator=gen) * 0.7 + 0.1
    craw = torch.randn(n, c, generator=gen)
    ref_p, ref_c = reference_banks(tokens, craw, grid)
    eng.new_bank(n, c, grid)
    for lo in range(0, n, 16):                                 # batches, as the embedding loop delivers them
        eng.ingest(tokens[lo:lo + 16], craw[lo:lo + 16], lo)
    b = eng.bank
    # block means: ATen's row-major window sum, divided by kh and then by kw -- exact for every block size (3 x 3 too)
    assert torch.equal(b["patches"].cpu(), ref_p), (b["patches"].cpu() - ref_p).abs().max()
    if c % 8 == 0:                               # ATen's contiguous-row norm: 8 vector lanes, then the lanes in order
        assert torch.equal(b["centers"].cpu(), ref_c), (b["centers"].cpu() - ref_c).abs().max()
    else:                                        # widths with a vector tail (no reference config): to rounding
        torch.testing.assert_close(b["centers"].cpu(), ref_c, rtol=3e-7, atol=1e-8)


def test_channel_major_maps(eng):
    """The trained-model branch keeps [B, C, H, W] maps (eval_cvt_diml.py:286-289): same kernel, other strides."""
    n, c, grid = 12, 128, 7
    maps = torch.randn(n, c, grid * grid, generator=torch.Generator().manual_seed(3))
    craw = torch.randn(n, c, generator=torch.Generator().manual_seed(4))
    eng.new_bank(n, c, grid)
    eng.ingest(maps, craw, 0, channel_major=True)
    assert torch.equal(eng.bank["patches"].cpu(), torch.nn.functional.normalize(maps, p=2, dim=1))
    assert torch.equal(eng.bank["centers"].cpu(), torch.nn.functional.normalize(craw, p=2, dim=1))


def test_ingested_operand_planes_equal_the_repack(eng):
    """A pass over banks built by ingest (operand planes written by the ingest kernel) against the same pass over the same
    fp32 banks registered the ordinary way (planes derived by vr_bank_prepare): scores, iteration counts, tallies identical."""
    from vitrerank.engine import OTParams
    n, k = 600, 100
    gen = torch.Generator().manual_seed(7)
    labels = synth.make_labels(n, 12, gen)
    proto = torch.randn(int(labels.max()) + 1, 14 * 14, 128, generator=gen)
    tokens = proto[labels] + 0.6 * torch.randn(n, 14 * 14, 128, generator=gen)
    craw = tokens.mean(1) + 0.05 * torch.randn(n, 128, generator=gen)
    roll = torch.softmax(torch.randn(n, 49, generator=gen), -1)
    p = OTParams(mode="rollout")
    eng.new_bank(n, 128, 7, with_rollout=True)
    for lo in range(0, n, 128):
        eng.ingest(tokens[lo:lo + 128], craw[lo:lo + 128], lo)
    eng.bank["rollout"].copy_(roll)
    eng.register_labels(labels)
    kp = max(k, eng.bank["max_num_pos"], 8)
    idx, approx = eng.stage0_topk(kp)
    score, niter = eng.rerank_scores(idx, k, p)
    tal, _ = eng.finalize(idx, approx, score, k, [0, k])
    patches, centers = eng.bank["patches"].clone(), eng.bank["centers"].clone()
    eng.register(patches, centers, roll, labels)               # ordinary registration: lazily re-packed by the first rerank
    idx2, approx2 = eng.stage0_topk(kp)
    score2, niter2 = eng.rerank_scores(idx2, k, p)
    tal2, _ = eng.finalize(idx2, approx2, score2, k, [0, k])
    assert torch.equal(idx, idx2) and torch.equal(score, score2) and torch.equal(niter, niter2) and torch.equal(tal, tal2)


def test_rejects_what_it_cannot_pool(eng):
    from vitrerank._lib import VitRerankError
    eng.new_bank(4, 128, 7)
    with pytest.raises(VitRerankError):
        eng.ingest(torch.randn(4, 100, 128), None, 0)          # 10 x 10 does not pool to 7 x 7 in whole blocks
    with pytest.raises(VitRerankError):
        eng.ingest(torch.randn(4, 196, 128), None, 2)          # range runs past the bank
