"""ctypes binding of libmindrec.so (include/mindrec.h).  Fails loudly: there is no CPU or
PyTorch fallback anywhere in this package -- if the library is missing, not built for this
GPU, or a call returns an error code, a RuntimeError is raised."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MINDREC_LIB") or os.path.join(HERE, "libmindrec.so")   # MINDREC_LIB: A/B of two builds (scripts/)

MR_F32, MR_BF16 = 0, 1
MR_RNN_LSTM, MR_RNN_GRU = 0, 1


class CnnShape(Structure):
    _fields_ = [("N", c_int64), ("L", c_int64), ("E", c_int64), ("H", c_int64), ("V", c_int64), ("precision", c_int)]


class RnnShape(Structure):
    _fields_ = [("B", c_int64), ("S", c_int64), ("H", c_int64), ("kind", c_int), ("reverse", c_int), ("precision", c_int)]


P = c_void_p
I64 = c_int64
_SIGNATURES = {
    "mr_version": (c_int, []),
    "mr_last_error": (c_char_p, []),
    "mr_device_check": (c_int, [c_int]),
    "mr_launch_count": (I64, []),
    "mr_embed_gather_f32": (c_int, [P, c_int, P, P, I64, I64, I64, P]),
    "mr_gather_titles": (c_int, [P, P, I64, I64, P, I64, P, I64, c_int, P, P, P]),
    "mr_embed_grad_workspace_bytes": (I64, [I64, I64, I64]),
    "mr_embed_grad_segreduce": (c_int, [P, c_int, P, c_int, I64, P, I64, I64, I64, I64, P, I64, P]),
    "mr_news_cnn_set_hot_tokens": (c_int, [POINTER(c_int64), c_int]),
    "mr_news_cnn_set_hot_replicas": (c_int, [c_int]),
    "mr_news_cnn_workspace_bytes": (I64, [POINTER(CnnShape), c_int]),
    "mr_news_cnn_fwd": (c_int, [POINTER(CnnShape), P, c_int, P, P, c_int, P, P, P, P, P, P, P, P, P, P, P, I64, P]),
    "mr_news_cnn_bwd": (c_int, [POINTER(CnnShape), P, c_int, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, I64, P]),
    "mr_news_cnn_bwd_table_workspace_bytes": (I64, [POINTER(CnnShape)]),
    "mr_token_group_plan_bytes": (I64, [I64, I64]),
    "mr_token_group_plan": (c_int, [P, c_int, I64, I64, P, I64, P]),
    "mr_news_cnn_bwd_table": (c_int, [POINTER(CnnShape), P, c_int, P, I64, P, P, P, P, P, P, P, P, P, P, P, P, P, I64, P, I64, P, P, I64, P]),
    "mr_rnn_workspace_bytes": (I64, [POINTER(RnnShape), c_int]),
    "mr_rnn_user_fwd": (c_int, [POINTER(RnnShape), P, P, P, P, P, P, P, P, P, P, P, P, I64, P]),
    "mr_rnn_user_bwd": (c_int, [POINTER(RnnShape), P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, I64, P]),
    "mr_attnpool_fwd": (c_int, [P, P, P, P, P, I64, I64, I64, P]),
    "mr_attnpool_bwd": (c_int, [P, P, P, P, P, P, I64, I64, I64, P]),
    "mr_avgpool_fwd": (c_int, [P, P, I64, I64, I64, P]),
    "mr_avgpool_bwd": (c_int, [P, P, I64, I64, I64, P]),
    "mr_mha_core_fwd": (c_int, [P, P, P, P, P, I64, I64, I64, I64, I64, P]),
    "mr_mha_core_bwd": (c_int, [P, P, P, P, P, P, I64, I64, I64, I64, I64, P]),
    "mr_mha_attn_fwd": (c_int, [P, I64, P, I64, P, P, P, I64, I64, I64, I64, I64, P]),
    "mr_mha_attn_bwd": (c_int, [P, I64, P, I64, P, P, P, I64, P, I64, I64, I64, I64, I64, I64, P]),
    "mr_linear_workspace_bytes": (I64, [I64, I64, I64]),
    "mr_linear_fwd": (c_int, [P, P, P, P, I64, I64, I64, c_int, c_int, P]),
    "mr_linear_bwd": (c_int, [P, P, P, P, P, P, I64, I64, I64, c_int, P, I64, P]),
    "mr_linear_tc_workspace_bytes": (I64, [I64, I64, I64, I64, c_int]),
    "mr_linear_tc_fwd": (c_int, [P, P, c_int, P, I64, I64, P, P, P, I64, I64, I64, I64, P, I64, P]),
    "mr_linear_tc_bwd": (c_int, [P, P, c_int, P, I64, I64, I64, I64, P, P, I64, P, P, P, P, I64, I64, I64, P, I64, P]),
    "mr_layernorm_fwd": (c_int, [P, P, P, P, c_float, P, P, P, I64, I64, P]),
    "mr_layernorm_bwd": (c_int, [P, P, P, c_float, P, P, P, P, P, P, I64, I64, I64, P]),
    "mr_score_logsoftmax_fwd": (c_int, [P, P, P, P, c_int, P, I64, I64, I64, P]),
    "mr_score_logsoftmax_bwd": (c_int, [P, P, P, P, P, P, I64, I64, I64, P]),
    "mr_score_sigmoid_fwd": (c_int, [P, P, P, I64, I64, I64, c_int, P]),
    "mr_score_sigmoid_gather_fwd": (c_int, [P, P, c_int, P, P, P, I64, I64, I64, I64, P]),
    "mr_rank_metrics": (c_int, [P, P, P, P, P, I64, I64, P]),
    "mr_adam_step": (c_int, [P, P, P, P, I64, I64, c_double, c_double, c_double, c_double, c_double, P, I64, I64, P]),
    "mr_adam_step_multi": (c_int, [c_int, P, P, P, P, P, P, I64, c_double, c_double, c_double, c_double, c_int, P, I64, I64, P]),
    "mr_adam_step_multi_dyn": (c_int, [c_int, P, P, P, P, P, P, I64, c_double, c_double, c_double, c_double, c_int, P, I64, I64, P, P]),
    "mr_nll_loss_fwd": (c_int, [P, P, c_int, P, I64, I64, P]),
    "mr_nll_loss_bwd": (c_int, [P, c_int, P, P, I64, I64, P]),
    "mr_cast_pad_bf16": (c_int, [P, P, I64, I64, I64, P]),
    "mr_tc_selftest": (c_int, [P, I64, I64, P, I64, I64, P, c_int, c_int, I64, I64, I64, I64, c_int, c_int, c_int, c_int, P]),
}

_lib = None


def exported_symbols():
    """Names include/mindrec.h declares (used by the symbol-coverage test)."""
    return sorted(_SIGNATURES)


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libmindrec.so is not built (%s missing). Run `python -m news_recommendation_mind_b200.build`; "
            "this package has no CPU / PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise RuntimeError("libmindrec.so does not export %s (stale build?)" % name) from exc
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mr_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t):
    """device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libmindrec needs CUDA tensors (got %s); there is no CPU fallback" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("libmindrec needs contiguous tensors")
    return c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def workspace(nbytes: int, device) -> torch.Tensor:
    if nbytes < 0:
        raise RuntimeError("workspace query failed")
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def index_flag(t: torch.Tensor) -> int:
    if t.dtype == torch.int64:
        return 1
    if t.dtype == torch.int32:
        return 0
    raise RuntimeError("index tensors must be int32 or int64, got %s" % t.dtype)
