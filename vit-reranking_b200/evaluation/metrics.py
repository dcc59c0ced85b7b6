"""Drop-in for the reference's evaluation/metrics.py: r1 / R-precision / MAP@R of one ranked
list, computed by libvitrerank.so (vr_metrics_rank).  The batched query loop in
evaluation.eval_cvt_diml.evaluate does not call these per query (it tallies on the GPU);
they remain for external callers."""
import ctypes as C

import torch

from vitrerank._lib import check, lib
from vitrerank.engine import require_cuda


def get_metrics_rank(tops, query_label, gallery_label):
    """evaluation/metrics.py:26-47.  tops: ranked gallery indices (any device), query_label: scalar,
    gallery_label [N].  Returns python floats (r1, rp, mapr)."""
    dev = tops.device if tops.device.type == "cuda" else require_cuda()
    tops = tops.to(device=dev, dtype=torch.int64).contiguous()
    labels = gallery_label.to(device=dev, dtype=torch.int64).contiguous()
    out = torch.zeros(3, dtype=torch.float64, device=dev)
    check(lib.vr_metrics_rank(C.c_void_p(tops.data_ptr()), tops.numel(), int(query_label),
                              C.c_void_p(labels.data_ptr()), labels.numel(), C.c_void_p(out.data_ptr()),
                              C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "vr_metrics_rank")
    r1, rp, mapr = out.cpu().tolist()
    return r1, rp, mapr


def get_metrics(sim, query_label, gallery_label):
    """evaluation/metrics.py:3-24: the same metrics from a similarity vector (argsort first)."""
    dev = sim.device if sim.device.type == "cuda" else require_cuda()
    tops = torch.argsort(sim.to(dev), descending=True, stable=True)
    return get_metrics_rank(tops, query_label, gallery_label)
