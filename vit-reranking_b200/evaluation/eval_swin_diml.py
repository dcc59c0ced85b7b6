"""Drop-in for the reference's evaluation/eval_swin_diml.py::evaluate (Swin, 7 x 7 final token grid; the caller
test_diml_swin.py).  Embedding [:160-197]: head projection -> [B, C, 7, 7] -> resize when the side differs from the grid ->
normalise; query loop [:241-271]: calc_similarity with the marginal flags.  Only the `use_featvit` branch the reference
hard-codes [:122] exists."""
from __future__ import annotations

from evaluation import _common


def evaluate(model, dataset, dataloader, training=False, trunc_nums=None, use_uniform=False, grid_size=4, blk_ind=0,
             use_cls_token=False, use_inverse=False, temperature=1.0, use_ot=True, ot_part=1.0, to_submit=False, use_minus=False,
             use_rollout=False, use_soft=False):
    model.eval()

    def project(model, out, aux):
        if training:
            return out.reshape(out.size(0), out.size(1), -1), out.reshape(out.size(0), out.size(1), -1).mean(2), True
        _, feat = aux
        return model.model.head(feat), out, False                               # [:178-181]

    n_total = len(dataset) if hasattr(dataset, "__len__") else None
    patches, centers, labels = _common.embed(model, dataloader, project, grid_size, n_total=n_total, resize_smaller=True)
    return _common.run(patches, centers, labels, trunc_nums, use_uniform=use_uniform, use_inverse=use_inverse,
                       temperature=temperature, use_cls_token=use_cls_token, ot_part=ot_part, use_minus=use_minus,
                       use_soft=use_soft)
