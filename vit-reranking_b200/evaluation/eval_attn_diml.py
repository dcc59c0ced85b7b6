"""Drop-in for the reference's evaluation/eval_attn_diml.py::evaluate (DeiT / ViT, 14 x 14 token grid: R = 196 at
--grid_size 14; the caller test_diml_attn.py).  Embedding [:150-206]: head projection of the patch tokens -> [B, C, 14, 14]
-> AdaptiveAvgPool2d(grid_size) when larger -> normalise; query loop [:219-273]: calc_similarity with every marginal flag
(use_uniform / use_inverse / use_minus / use_soft, temperature, use_cls_token, ot_part), ot_temp 0.05.  Only the
`use_featvit` branch the reference hard-codes [:110] exists."""
from __future__ import annotations

from evaluation import _common


def evaluate(model, dataset, dataloader, training=False, trunc_nums=None, use_uniform=False, grid_size=4, use_inverse=False,
             temperature=1.0, use_cls_token=False, attn_blk_ind=0, use_ot=True, ot_part=0.1, to_submit=False, use_minus=False,
             use_rollout=False, use_soft=False):
    model.eval()

    def project(model, out, aux):
        if training:                                                            # [:184-185]
            return out.reshape(out.size(0), out.size(1), -1), out.reshape(out.size(0), out.size(1), -1).mean(2), True
        _, feat = aux
        return model.model.head(feat), out, False                               # [:169-176]: bs x L x C, pooled by the ingest

    n_total = len(dataset) if hasattr(dataset, "__len__") else None
    patches, centers, labels = _common.embed(model, dataloader, project, grid_size, n_total=n_total)
    return _common.run(patches, centers, labels, trunc_nums, use_uniform=use_uniform, use_inverse=use_inverse,
                       temperature=temperature, use_cls_token=use_cls_token, ot_part=ot_part, use_minus=use_minus,
                       use_soft=use_soft)
