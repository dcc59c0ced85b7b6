"""Shared body of the sibling drivers of the DIML rerank loop: evaluation/eval_diml.py
(ResNet-50), eval_attn_diml.py (DeiT, 14 x 14 grid), eval_swin_diml.py.  Their query loops are the loop of
eval_cvt_diml.py with calc_similarity's cross-correlation marginals (reference eval_diml.py:163-194,
eval_attn_diml.py:219-273, eval_swin_diml.py:241-271); what differs is how the backbone's output becomes the token
map, which each module states in its `project` callback.  Embedding goes through the ingest kernel (vr_bank_ingest),
the loop through RerankEngine (one batched GPU pass)."""
from __future__ import annotations

import torch
import torch.nn as nn

from vitrerank.engine import RerankEngine

try:
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(x, **k):
        return x


def make_resize(grid_size, always_pool=False):
    """The reference's `resize` module (eval_diml.py:89-97, eval_attn_diml.py:137-144, eval_swin_diml.py:131-138)."""
    if always_pool or 7 % grid_size == 0:
        return nn.AdaptiveAvgPool2d(grid_size)
    return nn.Sequential(nn.Upsample(grid_size * 4, mode='bilinear', align_corners=True), nn.AdaptiveAvgPool2d(grid_size))


def embed(model, dataloader, project, grid_size, n_total=None, device=None, pool_plain=None, resize_smaller=False):
    """project(model, out, aux) -> (x, centres [B, C], channel_major) for one batch, x = tokens [B, L, C] or maps [B, C, H, W].
    A square map larger than the grid is pooled: by the ingest kernel when the reference's `resize` is a plain
    AdaptiveAvgPool2d (pool_plain; default 7 % grid_size == 0) and the side is a whole multiple of the grid, else by the
    reference's torch module first.  Returns the engine's registered banks (patches, centers, labels)."""
    device = device or torch.device('cuda')
    eng = RerankEngine.get(device)
    pool_plain = (7 % grid_size == 0) if pool_plain is None else pool_plain
    resize = make_resize(grid_size, always_pool=pool_plain)
    labels, staged = [], []
    state = {"ready": False, "lo": 0, "n": n_total}

    def put(x, craw, channel_major, side, grid):
        if not state["ready"]:
            if state["n"] is None:
                staged.append((x, craw, channel_major, side, grid))
                return
            c = x.shape[1] if channel_major else x.shape[2]
            eng.new_bank(int(state["n"]), int(c), grid)
            state["ready"] = True
        eng.ingest(x, craw, state["lo"], h=side, w=side, channel_major=channel_major)
        state["lo"] += x.shape[0]

    with torch.no_grad():
        for inp in tqdm(dataloader, desc='Embedding Data...'):
            img, target = inp[1].to(device), inp[0]
            out = model(img)
            aux = None
            if isinstance(out, tuple):
                out, aux = out
            x, craw, channel_major = project(model, out, aux)
            if channel_major:
                x = x.reshape(x.size(0), x.size(1), -1)
                side = int(round(x.size(-1) ** 0.5))
            else:
                side = int(round(x.size(1) ** 0.5))
            grid = side
            if side > grid_size or (side < grid_size and resize_smaller):
                if side > grid_size and pool_plain and side % grid_size == 0:
                    grid = grid_size                                   # pooled inside the ingest kernel
                else:                                                  # the reference's torch module (bilinear upsample + pool)
                    maps = x if channel_major else x.permute(0, 2, 1)
                    maps = resize(maps.reshape(maps.size(0), maps.size(1), side, side))
                    x, channel_major = maps.reshape(maps.size(0), maps.size(1), -1), True
                    side = grid = maps.size(-1)
            put(x.detach().float().contiguous(), craw.detach().float(), channel_major, side, grid)
            labels.append(torch.as_tensor(target).reshape(-1))
    labels = torch.cat(labels, 0).long()
    if not state["ready"]:
        state["n"] = int(labels.numel())
        for item in staged:
            put(*item)
    assert state["lo"] == eng.bank["n"], f"embedded {state['lo']} images into banks sized for {eng.bank['n']}"
    eng.register_labels(labels)
    b = eng.bank
    return b["patches"], b["centers"], b["labels"]


def run(patches, centers, labels, trunc_nums, report=True, **flags):
    """The query loop over finished banks -> the reference's dict {'r1', 'rp', 'mapr'} (percentages)."""
    from evaluation.eval_cvt_diml import evaluate_banks
    trunc_nums = trunc_nums or [0, 5, 10, 50, 100, 500, 1000]
    data = evaluate_banks(patches, centers, None, labels, trunc_nums=trunc_nums, **flags)
    if report:
        for i, t in enumerate(trunc_nums):
            print(f"trunc_num: {t}")
            print('###########')
            print('Now rank-1 acc=%f, RP=%f, MAP@R=%f' % (data['r1'][i], data['rp'][i], data['mapr'][i]))
    return data
