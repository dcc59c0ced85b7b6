"""Drop-in for the reference's evaluation/eval_diml.py::evaluate (ResNet-50 DIML, the caller test_diml.py): same
signature and returned dict.  Embedding [:99-149]: head (or last_linear) projection -> AdaptiveAvgPool2d(grid_size) when
the map is larger than the grid -> per-location / per-centre L2 normalisation, through the ingest kernel; query loop
[:163-194] = calc_similarity with the positional flags (use_uniform, use_inverse, temperature, use_cls_token), ot_temp
0.05, full OT -- one batched pass of the rerank engine."""
from __future__ import annotations

from evaluation import _common


def evaluate(model, dataset, dataloader, no_training=True, trunc_nums=None, use_uniform=False, grid_size=4, use_inverse=False,
             temperature=0.1, use_cls_token=False, to_submit=False, plot_topk=False):
    model.eval()
    has_head = any('head' in name for name, _ in model.named_modules())        # [:56-61]

    def project(model, out, aux):
        if not no_training:                                                     # [:139-140] the output is the feature map
            return out, out.reshape(out.size(0), out.size(1), -1).mean(2), True          # [:147] centre = mean over locations
        _, feat = aux
        if has_head:                                                            # [:124-127]
            tok = model.model.head(feat)                                        # bs x L x C
            return tok, out, False
        feat = model.model.last_linear(feat.transpose(1, 3)).transpose(1, 3)    # [:128-131]  bs x C x H x W
        return feat, out, True

    n_total = len(dataset) if hasattr(dataset, "__len__") else None
    patches, centers, labels = _common.embed(model, dataloader, project, grid_size, n_total=n_total, pool_plain=(7 % grid_size == 0) or has_head)
    return _common.run(patches, centers, labels, trunc_nums, use_uniform=use_uniform, use_inverse=use_inverse,
                       temperature=temperature, use_cls_token=use_cls_token, ot_part=1.0)
