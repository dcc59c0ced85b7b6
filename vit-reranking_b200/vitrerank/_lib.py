"""ctypes binding of libvitrerank.so (C ABI declared in include/vitrerank.h).

There is no fallback: if the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C vit-reranking_b200/csrc`)
importing this module raises, and every compute entry needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvitrerank.so")

VR_MODE = {"rollout": 0, "uniform": 1, "inverse": 2, "minus": 3, "soft": 4, "relu": 5}


class OTParamsStruct(C.Structure):
    _fields_ = [("mode", C.c_int32), ("use_cls_token", C.c_int32), ("ot_temp", C.c_float),
                ("temperature", C.c_float), ("ot_part", C.c_float), ("max_iter", C.c_int32),
                ("thresh", C.c_float)]


class VitRerankError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first (make -C vit-reranking_b200/csrc, or "
            f"__graft_entry__.build()).  vitrerank has no CPU or PyTorch fallback path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_float
    P = C.POINTER
    sig = {
        "vr_abi_version": (C.c_int, []),
        "vr_last_error": (C.c_char_p, []),
        "vr_partial_ot_pad": (C.c_float, [C.c_float]),
        "vr_create": (C.c_int, [C.c_int, P(vp)]),
        "vr_destroy": (C.c_int, [vp]),
        "vr_device_info": (C.c_int, [vp, P(i32), P(i32)]),
        "vr_bank_register": (C.c_int, [vp, vp, vp, vp, vp, vp, i64, i32, i32]),
        "vr_bank_prepare": (C.c_int, [vp, i64, i64, vp]),
        "vr_bank_labels": (C.c_int, [vp, vp, vp]),
        "vr_bank_ingest": (C.c_int, [vp, vp, vp, i32, i64, i64, i32, i32, vp]),
        "vr_num_pos": (C.c_int, [vp, vp, i64, vp, P(i32), vp]),
        "vr_stage0_workspace_bytes": (sz, [vp, i64, i32]),
        "vr_stage0_topk": (C.c_int, [vp, vp, vp, i64, i64, i64, i32, vp, vp, vp, sz, vp]),
        "vr_stage0_stats": (C.c_int, [vp, P(C.c_uint32), vp]),
        "vr_rerank_workspace_bytes": (sz, [vp, i64, i32, P(OTParamsStruct)]),
        "vr_rerank_scores": (C.c_int, [vp, i64, i64, i64, i32, vp, i32, P(OTParamsStruct), vp, vp, vp, sz, vp]),
        "vr_rerank_scores_queries": (C.c_int, [vp, vp, vp, vp, i64, i32, vp, i32, P(OTParamsStruct), vp, vp, vp, sz, vp]),
        "vr_finalize_workspace_bytes": (sz, [vp, i64, i32]),
        "vr_finalize": (C.c_int, [vp, i64, i64, i64, i32, i32, vp, vp, vp, P(i32), i32, vp, vp, vp, sz, vp]),
        "vr_blend_rank": (C.c_int, [vp, i64, i32, i32, vp, vp, vp, vp, vp]),
        "vr_rollout_block_workspace_bytes": (sz, [i64, i32, i32, i32]),
        "vr_rollout_block": (C.c_int, [vp, vp, i64, i32, i32, i32, i32, i32, i64, i32, vp, vp, sz, vp]),
        "vr_rollout_chain": (C.c_int, [vp, vp, i32, i64, i32, i32, vp, vp]),
        "vr_sinkhorn_workspace_bytes": (sz, [i64, i32, i32]),
        "vr_sinkhorn": (C.c_int, [vp, vp, vp, i64, i32, i32, i32, f32, vp, vp, vp, sz, vp]),
        "vr_calc_similarity_workspace_bytes": (sz, [i64, i32, i32, P(OTParamsStruct)]),
        "vr_calc_similarity": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, P(OTParamsStruct),
                                         vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
        "vr_global_similarity": (C.c_int, [vp, vp, i64, i32, vp, vp]),
        "vr_evaluate_host": (C.c_int, [vp, vp, vp, vp, vp, i64, i32, i32, i64, i64, i64, P(i32), i32,
                                       P(OTParamsStruct), P(C.c_double), vp]),
        "vr_evaluate_registered": (C.c_int, [vp, i64, i64, i64, P(i32), i32, i32, P(OTParamsStruct),
                                             P(C.c_double), vp, vp]),
        "vr_metrics_rank": (C.c_int, [vp, i64, i64, vp, i64, vp, vp]),
        "vr_debug_err_trace": (C.c_int, [vp, vp]),
        "vr_take_launch_count": (i64, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTS = _load()


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib.vr_last_error().decode("utf-8", "replace")
        raise VitRerankError(f"{what or 'libvitrerank'} failed ({rc}): {msg}")


def take_launch_count() -> int:
    return int(lib.vr_take_launch_count())
