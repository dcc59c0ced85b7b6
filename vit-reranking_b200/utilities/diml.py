"""Drop-in for the reference's utilities/diml.py: same names, same positional / keyword
signatures, same return conventions -- computed by libvitrerank.so on a B200.

    from utilities.diml import Sinkhorn, calc_similarity, calc_similarity_cvt_rollout, ...

Tensors follow the reference's convention: the device / dtype are taken from
`anchor_center` (utilities/diml.py:93-94,333-336) and everything else is moved onto it.
That device must be CUDA: there is no CPU or PyTorch fallback (a CPU tensor raises).
Functions of the reference that are off the rerank path (other backbones' variants,
matplotlib helpers) import fine, answer stage 0 (which is the same global cosine for all of
them) and raise NotImplementedError for their stage 1.
"""
from __future__ import annotations

import numpy as np
import torch

from vitrerank.engine import OTParams, RerankEngine
from vitrerank._lib import VitRerankError


def _engine_for(t: torch.Tensor) -> RerankEngine:
    if t is None or not torch.is_tensor(t) or t.device.type != "cuda":
        raise VitRerankError("utilities.diml (B200 drop-in) needs CUDA tensors; there is no CPU path")
    return RerankEngine.get(t.device)


def Sinkhorn(K, u, v, iter=100):
    """utilities/diml.py:42-54.  K [b, m, n], u [b, m], v [b, n] -> T [b, m, n]; the stop test
    (batch mean |r - r_prev| < 0.1) is taken on the device, without host round trips."""
    eng = _engine_for(K)
    T, _ = eng.sinkhorn(K, u, v, max_iter=iter)
    return T.to(K.dtype)


def Sinkhorn_partial(K, u, v, ot_part=0.1):
    """utilities/diml.py:59-75: one dummy row / column of mass 1 - ot_part; returns T_extended
    [b, m+1, n+1]."""
    assert ot_part < 1 and ot_part >= 0
    b, m, n = K.shape
    pad = K.new_tensor(1 - ot_part)
    K_ext = K.new_zeros(b, m + 1, n + 1)
    K_ext[:, :m, :n] = K
    K_ext[:, :m, n] = pad
    K_ext[:, m, :n] = pad
    u_ext = torch.cat([u, pad.expand(b, 1)], -1)
    v_ext = torch.cat([v, pad.expand(b, 1)], -1)
    return Sinkhorn(K_ext, u_ext, v_ext)


def _stage0(anchor_center, fb_center):
    eng = _engine_for(anchor_center)
    return eng.global_similarity(anchor_center, fb_center.to(anchor_center)).to(anchor_center.dtype), None


def calc_similarity(anchor, anchor_center, fb, fb_center, stage, use_uniform=False, use_inverse=False,
                    temperature=1.0, use_cls_token=False, ot_temp=0.05, use_minus=False, ot_part=1.0,
                    use_soft=False):
    """utilities/diml.py:77-147.  stage 0: (sim [N], None).  stage 1: (score [N],
    (u, v, T or T_ext, sim_r, cc)) with the reference's marginal modes."""
    if stage == 0:
        return _stage0(anchor_center, fb_center)
    ref = anchor_center if torch.is_tensor(anchor_center) else anchor
    eng = _engine_for(ref if ref is not None else anchor)
    if use_cls_token:
        assert anchor_center.ndim == 1
    p = OTParams.from_flags(use_rollout=False, use_uniform=use_uniform, use_inverse=use_inverse,
                            use_minus=use_minus, use_soft=use_soft, temperature=temperature,
                            use_cls_token=use_cls_token, ot_part=ot_part, ot_temp=ot_temp)
    score, uv, _ = eng.calc_similarity(anchor, anchor_center if use_cls_token else None, fb,
                                       fb_center if use_cls_token else None, p)
    return score, uv


def calc_similarity_cvt_rollout(anchor_center, anchor, anchor_query, fb_center, fb, fb_key, stage,
                                use_uniform=False, ot_temp=0.05, use_ot=True, ot_part=1.0,
                                device=None):
    """utilities/diml.py:323-366: attention-rollout marginals.  `use_ot` and `device` are
    accepted and ignored, as in the reference."""
    if stage == 0:
        return _stage0(anchor_center, fb_center)
    eng = _engine_for(anchor_center)
    p = OTParams.from_flags(use_rollout=True, use_uniform=use_uniform, ot_part=ot_part, ot_temp=ot_temp)
    n, _, r = fb.shape
    score, uv, _ = eng.calc_similarity(anchor, None, fb, None, p,
                                       q_rollout=anchor_query.reshape(-1)[:r] if not use_uniform else None,
                                       c_rollout=fb_key.reshape(n, r) if not use_uniform else None)
    return score, uv


def _stage1_off_path(name):
    raise NotImplementedError(
        f"utilities.diml.{name}(stage=1) is off the DIML CvT rerank path this B200 build covers "
        f"(SURVEY.md section 8b); stage 0 and calc_similarity / calc_similarity_cvt_rollout are available")


def calc_distance(anchor, anchor_center, fb, fb_center, stage, use_uniform=False, use_exp=True, temperature=1.0,
                  use_cls_token=False):
    if stage == 0:   # utilities/diml.py:150-153: a one-line torch expression, kept as such (off the rerank path)
        dist = torch.sqrt(torch.sum(torch.pow(anchor_center - fb_center, 2), dim=1) + 1e-6).view(fb_center.size(0))
        return dist, None
    _stage1_off_path("calc_distance")


def calc_similarity_vit(anchor_center, anchor_feat, anchor_query, fb_center, fb_feat, fb_keyt, stage,
                        use_uniform=False, use_exp=False, temperature=1.):
    if stage == 0:
        return _stage0(anchor_center, fb_center)
    _stage1_off_path("calc_similarity_vit")


def calc_similarity_cvt(anchor_center, anchor, anchor_query, fb_center, fb, fb_key, stage, use_uniform=False,
                        use_ot=False):
    if stage == 0:
        return _stage0(anchor_center, fb_center)
    _stage1_off_path("calc_similarity_cvt")


def calc_similarity_featvit(anchor_feat, fb_feat, stage, use_uniform=False, use_self=False, use_cam=False,
                            anchor_cam=None, fb_cam=None):
    if stage == 0:
        return _stage0(anchor_feat[:, 0].contiguous(), fb_feat[:, :, 0].contiguous())
    _stage1_off_path("calc_similarity_featvit")


def calc_similarity_mhvit(anchor_feat, fb_feat, stage, use_uniform=False):
    if stage == 0:
        return _stage0(anchor_feat[:, 0].contiguous(), fb_feat[:, :, 0].contiguous())
    _stage1_off_path("calc_similarity_mhvit")


def visual_cross_correlation(cc):
    raise NotImplementedError("matplotlib figure helper of the reference (utilities/diml.py:8-40); not part of the "
                              "rerank path")


def input_inv_transform(x):
    """utilities/diml.py:475-486: undo the ImageNet mean/std normalisation of a [3, H, W] array and
    return uint8 HWC."""
    assert x.ndim == 3 and x.shape[0] == 3
    std = np.asarray([0.229, 0.224, 0.225], dtype=np.float64).reshape(3, 1, 1)
    mean = np.asarray([0.485, 0.456, 0.406], dtype=np.float64).reshape(3, 1, 1)
    y = np.asarray(x) * std + mean
    return np.transpose(np.uint8(y * 255), (1, 2, 0))
