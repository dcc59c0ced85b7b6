"""Drop-in for the reference's evaluation/eval_cvt_diml.py::evaluate (the entry point of
test_diml_cvt.py).  Same signature, same printed lines, same returned dict
{'r1': [...], 'rp': [...], 'mapr': [...]} ordered as trunc_nums.

What changed underneath (reference file:line in brackets):
  * PHASE A (embedding, [:225-305]) still runs the caller's model batch by batch, but the three
    banks stay in HBM instead of being parked on the host [:278-279,:256].
  * PHASE B (the per-query Python loop [:316-372]) is ONE batched pass through
    libvitrerank.so: first-stage top-K' select, fused gather + patch-sim + Sinkhorn + score,
    blend / re-sort / metric tallies on the GPU (vitrerank.engine.RerankEngine.evaluate).  With
    torch.distributed initialised, the queries are sharded over the ranks and the tallies are
    all-reduced (vitrerank.distributed).
  * The matplotlib / cv2 heat-map hook [:375-397] is not part of this build (it needs the
    dataset's images and matplotlib; with --plot_topk 1 it raises in the reference, SURVEY.md
    section 8b).  Pass `visual_hook=callable(idx, uv)` to receive the `uv` tuple of the queries
    the reference would have drawn (idx < 1000 and idx % 10 == 0).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from vitrerank import distributed as vdist
from vitrerank.engine import OTParams, RerankEngine

try:  # progress bars as in the reference, optional
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(x, **k):
        return x


# ---- attention rollout (input production for --use_rollout; reference [:54-146]) -------------------
def resize_attn_map(attn, resize, stage, grid, blk_id=0):
    """[:54-71]: drop the cls row/column in stage 2, pool both token axes to grid x grid."""
    if stage == 2:
        attn = attn[:, 1:, 1:]
    b, h, w = attn.shape
    side = int(w ** .5)
    attn = attn.reshape(b, h, side, side)
    if attn.size(-1) > grid:
        attn = resize(attn)
    attn = attn.reshape(b, h, -1).permute(0, 2, 1)
    side = int(h ** .5)
    attn = attn.reshape(b, -1, side, side)
    if attn.size(-1) > grid:
        attn = resize(attn)
    return attn.reshape(b, grid * grid, grid * grid).permute(0, 2, 1)


def filter_attention_map(raw_attn, discard_ratio, head_fusion, show_fig=False):
    """[:74-108]: fuse heads, zero the `discard_ratio` smallest entries.  As in the reference the
    fancy-index assignment zeroes, in EVERY image of the batch, the union of the coordinates
    discarded in any image."""
    h, w = raw_attn.size(-2), raw_attn.size(-1)
    if head_fusion == 'mean':
        fused = raw_attn.mean(dim=1)
    elif head_fusion == 'max':
        fused = raw_attn.max(dim=1)[0]
    elif head_fusion == 'min':
        fused = raw_attn.min(dim=1)[0]
    else:
        raise ValueError("head fusion type not supported")
    flat = fused.reshape(fused.size(0), -1)
    _, idx = flat.topk(int(flat.size(-1) * discard_ratio), -1, False)
    iy = (idx / w).long()
    ix = (idx % w).long()
    out = flat.reshape(flat.size(0), h, w)
    out[:, iy, ix] = 0
    return out


def _rollout_on_device(blocks, grid, use_res):
    """The producer on the B200 (csrc/rollout.cu through vr_rollout_block / vr_rollout_chain) when every block's attention
    is an fp32 CUDA tensor with square token grids; None otherwise (the torch statements below then run)."""
    probs = [blk._probs[0] for _, blk in blocks]
    if not probs or not all(torch.is_tensor(p) and p.is_cuda and p.dtype == torch.float32 and p.dim() == 4 for p in probs):
        return None
    for (si, _), p in zip(blocks, probs):
        d = 1 if si == 2 else 0
        for t in (p.size(-2) - d, p.size(-1) - d):
            side = int(round(t ** .5))
            if side * side != t or side < grid:
                return None
    if grid * grid > 128:
        return None
    eng = RerankEngine.get(probs[0].device)
    mats = torch.stack([eng.rollout_block(p, drop_cls=(si == 2), grid=grid, discard_ratio=0.1, head_fusion='min')
                        for (si, _), p in zip(blocks, probs)])
    return list(eng.rollout_chain(mats, use_res=use_res))


def get_attention_rollout(model, input, grid=7, use_res=True, display_map=False):
    """[:111-146]: per-block attention (blk._probs[0]) -> min-fused, filtered, pooled to
    grid^2 x grid^2, identity added and row-normalised, then chained with bmm.  Returns the list
    of joint attentions (one per block); evaluate keeps joint[-1].mean(1).  Attention on a CUDA device is
    processed by the library's kernels (rollout.cu), on the CPU by the reference's torch statements."""
    with torch.no_grad():
        model.both_forward(input)
        blocks = [(si, blk) for si in range(3) for blk in getattr(model, f'stage{si}').blocks]
        joint = _rollout_on_device(blocks, grid, use_res)
        if joint is not None:
            return joint
        resize = nn.AdaptiveAvgPool2d((grid, grid))
        mats = []
        for si, blk in blocks:
            a = filter_attention_map(blk._probs[0], discard_ratio=0.1, head_fusion='min')
            mats.append(resize_attn_map(a, resize, si, grid).detach())
        mats = torch.stack(mats)
        if use_res:
            eye = torch.eye(mats.size(2), device=mats.device, dtype=mats.dtype)
            mats = mats + eye
            mats = mats / mats.sum(dim=-1).unsqueeze(-1)
    joint = [mats[0]]
    for j in range(1, len(mats)):
        joint.append(torch.bmm(mats[j], joint[j - 1]))
    return joint


def evaluate_patch_similarity(model, dataset, dataloader):
    raise NotImplementedError("evaluate_patch_similarity ([:168-194], ViT block self-similarity statistics) is "
                              "off the rerank path; not part of the B200 build")


# ---- PHASE A: embed the test set into three HBM-resident banks [:225-305] ----------------------------
def embed_banks(model, dataloader, training=False, grid_size=4, use_rollout=False, device=None, n_total=None):
    """The backbone (and, for --use_rollout, the attention rollout) runs in torch; everything after the head projection --
    permute, AdaptiveAvgPool2d(grid_size), per-patch / per-centre L2 normalisation [:271-278,:304-305] -- is the ingest
    kernel (vr_bank_ingest), which writes the fp32 banks and the operand copy of the fused rerank kernel batch by batch.
    n_total (= len(dataset)) lets the banks be allocated up front; without it the batches are staged and ingested at the end.
    Returns (patches [N,C,R], centers [N,C], rollout [N,R] or None, labels [N]) -- the engine's registered banks."""
    device = device or torch.device('cuda')
    eng = RerankEngine.get(device)
    no_training = not training
    resize = None
    if no_training and 7 % grid_size != 0:   # [:228-234] the bilinear path stays in torch, the kernel then only normalises
        resize = nn.Sequential(nn.Upsample(grid_size * 4, mode='bilinear', align_corners=True),
                               nn.AdaptiveAvgPool2d(grid_size))
    labels, staged = [], []
    state = {"ready": False, "lo": 0}

    def put(tokens, craw, roll, channel_major):
        if not state["ready"]:
            if n_total is None:
                staged.append((tokens, craw, roll, channel_major))
                return
            c = tokens.shape[1] if channel_major else tokens.shape[2]
            eng.new_bank(int(n_total), int(c), grid_size, with_rollout=use_rollout)
            state["ready"] = True
        lo = state["lo"]
        eng.ingest(tokens, craw, lo, channel_major=channel_major)
        if roll is not None:
            eng.bank["rollout"][lo:lo + tokens.shape[0]].copy_(roll)
        state["lo"] = lo + tokens.shape[0]

    with torch.no_grad():
        for inp in tqdm(dataloader, desc='Embedding Data...'):
            img, target = inp[1].to(device), inp[0]
            out = model(img)
            roll = None
            if use_rollout:
                rollout = get_attention_rollout(model.model, img, display_map=False)
                roll = rollout[-1].mean(1).detach().to(device).float()
            aux = None
            if isinstance(out, tuple):
                out, aux = out
            if no_training:
                _, tokens = aux
                tokens = model.model.head(tokens)                            # bs x L x C  [:269]
                side = int(tokens.size(1) ** 0.5)
                if resize is not None and side != grid_size:
                    t = tokens.permute(0, 2, 1).reshape(tokens.size(0), -1, side, side)
                    put(resize(t).reshape(t.size(0), t.size(1), -1).detach().float(), out.detach().float(), roll, True)
                else:
                    put(tokens.detach().float(), out.detach().float(), roll, False)
            else:
                put(out.reshape(out.size(0), out.size(1), -1).detach().float(), aux[0].detach().float(), roll, True)   # [:286-289]
            labels.append(torch.as_tensor(target).reshape(-1))
    labels = torch.cat(labels, 0).long()
    if not state["ready"]:
        n_total = int(labels.numel())
        for item in staged:
            put(*item)
    assert state["lo"] == eng.bank["n"], f"embedded {state['lo']} images into banks sized for {eng.bank['n']}"
    eng.register_labels(labels)
    b = eng.bank
    return b["patches"], b["centers"], b["rollout"], b["labels"]


def evaluate_banks(patches, centers, rollout, labels, trunc_nums=None, use_uniform=False, use_inverse=False,
                   temperature=1.0, use_cls_token=False, ot_part=0.1, use_minus=False, use_rollout=False,
                   device=None, return_extra=False, use_soft=False):
    """PHASE B [:308-372,:402-416] over pre-built banks: the reference's query loop as one batched
    GPU pass.  Returns the reference's dict (percentages over the N gallery images)."""
    trunc_nums = trunc_nums or [0, 5, 10, 50, 100, 500, 1000]                   # [:309]
    if use_rollout and rollout is None and not use_uniform:
        raise ValueError("use_rollout needs the rollout bank")
    eng = RerankEngine.get(device if device is not None else (patches.device if patches.is_cuda else None))
    b = eng._bank
    if b is not None and b["patches"] is patches and b["centers"] is centers and b["rollout"] is rollout:
        # the banks embed_banks just ingested: already registered, operand copy written by the ingest kernel
        if b["labels"] is not labels:
            eng.register_labels(labels)
    else:
        eng.register(patches, centers, rollout, labels)
    params = OTParams.from_flags(use_rollout=use_rollout, use_uniform=use_uniform, use_inverse=use_inverse,
                                 use_minus=use_minus, use_soft=use_soft, temperature=temperature,
                                 use_cls_token=use_cls_token, ot_part=ot_part, ot_temp=0.05)   # [:341], diml.py:325
    n = patches.shape[0]
    tallies, niter = vdist.evaluate_sharded(eng, trunc_nums, params)
    scale = float(n / 100)                                                       # [:403-405]
    data = {
        'r1': [float(tallies[i, 0] / scale) for i in range(len(trunc_nums))],
        'rp': [float(tallies[i, 1] / scale) for i in range(len(trunc_nums))],
        'mapr': [float(tallies[i, 2] / scale) for i in range(len(trunc_nums))],
    }
    if return_extra:
        data['recall_at_1_2_4_8'] = [[float(x / scale) for x in tallies[i, 3:7]] for i in range(len(trunc_nums))]
        data['sinkhorn_iters'] = niter
    return data


def evaluate(model, dataset, dataloader, training=False, trunc_nums=None, use_uniform=False, grid_size=4,
             use_inverse=False, temperature=1.0, use_cls_token=False, attn_blk_ind=0, use_ot=True, ot_part=0.1,
             to_submit=False, use_minus=False, use_rollout=False, plot_topk=1, visual_hook=None, bank_cache=None):
    """evaluation/eval_cvt_diml.py:196-416.  bank_cache (or $VITRERANK_BANK_CACHE): path of an on-disk bank
    (vitrerank/bankfile.py) -- read instead of embedding when it exists and has the rollout bank the flags need,
    written after embedding otherwise; the working form of the cache the reference keeps disabled
    (evaluation/eval_diml.py:80-85,151-153)."""
    device = torch.device('cuda')
    model.eval()
    bank_cache = bank_cache or os.environ.get("VITRERANK_BANK_CACHE")
    banks = None
    if bank_cache and os.path.exists(bank_cache):
        from vitrerank import bankfile
        banks = bankfile.load(bank_cache)
        if banks[3] is None or (use_rollout and not use_uniform and banks[2] is None):
            banks = None    # written by a run without labels / rollout: embed again
    if banks is None:
        n_total = len(dataset) if hasattr(dataset, "__len__") else None
        banks = embed_banks(model, dataloader, training=training, grid_size=grid_size, use_rollout=use_rollout,
                            device=device, n_total=n_total)
        if bank_cache:
            from vitrerank import bankfile
            bankfile.save(bank_cache, *banks)
    patches, centers, rollout, labels = banks
    trunc_nums = trunc_nums or [0, 5, 10, 50, 100, 500, 1000]
    data = evaluate_banks(patches, centers, rollout, labels, trunc_nums=trunc_nums, use_uniform=use_uniform,
                          use_inverse=use_inverse, temperature=temperature, use_cls_token=use_cls_token,
                          ot_part=ot_part, use_minus=use_minus, use_rollout=use_rollout, device=device)
    if visual_hook is not None and max(trunc_nums) > 0:
        _run_visual_hook(visual_hook, patches, centers, rollout, max(trunc_nums), use_rollout, use_uniform,
                         use_inverse, use_minus, temperature, use_cls_token, ot_part)
    for i, trunc_num in enumerate(trunc_nums):                                   # [:407-409]
        print(f"trunc_num: {trunc_num}, ot part: {ot_part}")
        print('###########')
        print('Now rank-1 acc=%f, RP=%f, MAP@R=%f' % (data['r1'][i], data['rp'][i], data['mapr'][i]))
    return data


def _run_visual_hook(hook, patches, centers, rollout, k, use_rollout, use_uniform, use_inverse, use_minus,
                     temperature, use_cls_token, ot_part):
    """Materialise the reference's `uv` tuple (u, v, T, sim_r, cc) only for the queries it would
    have visualised [:390]: idx < 1000 and idx % 10 == 0."""
    from utilities.diml import calc_similarity, calc_similarity_cvt_rollout
    eng = RerankEngine.get(patches.device)
    n = patches.shape[0]
    for idx in range(0, min(n, 1000), 10):
        top, _ = eng.stage0_topk(min(k, n), q_start=idx, q_stride=1, nq=1)
        top = top[0][top[0] >= 0].long()
        if use_rollout:
            _, uv = calc_similarity_cvt_rollout(centers[idx], patches[idx], rollout[idx], centers[top], patches[top],
                                                rollout[top], stage=1, use_uniform=use_uniform, ot_part=ot_part)
        else:
            _, uv = calc_similarity(patches[idx], centers[idx], patches[top], centers[top], stage=1,
                                    use_uniform=use_uniform, use_inverse=use_inverse, temperature=temperature,
                                    use_cls_token=use_cls_token, ot_temp=0.05, use_minus=use_minus, ot_part=ot_part)
        hook(idx, uv)
