"""ctypes binding of libdcnr_sm100a.so (include/dcnr.h).

The library is the product; this module only marshals pointers.  It fails loudly when the
shared object is missing -- there is no CPU or eager-PyTorch fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_uint64, c_void_p

import torch

MAX_CAT, MAX_RES, MAX_CROSS, PAD = 8, 8, 8, 32
OK, ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_INDEX = 0, -1, -2, -3, -4
PRECISIONS = {"fp32": 0, "tf32x3": 1, "tf32": 2, "bf16": 3, "fp16x3": 4}
PRECISION_NAMES = {v: k for k, v in PRECISIONS.items()}
ABI_VERSION = 4

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdcnr_sm100a.so")


class Dims(Structure):
    _fields_ = [("emb_dim", c_int32), ("n_cat", c_int32), ("n_num", c_int32), ("hidden", c_int32),
                ("n_cross", c_int32), ("n_res", c_int32), ("in_dim", c_int32), ("in_dim_pad", c_int32),
                ("n_users", c_int64), ("n_items", c_int64), ("cat_rows", c_int64 * MAX_CAT),
                ("cat_width", c_int32 * MAX_CAT), ("dropout_p", c_float), ("bn_eps", c_float),
                ("bn_momentum", c_float), ("precision", c_int32), ("dp_sparse_tables", c_int32), ("dp_batch_cap", c_int64), ("dropout_step", c_void_p),
                ("eval_flags", c_void_p), ("tower_pack", c_void_p), ("comm", c_void_p)]


def mark_mutated(tensors):
    """Tell torch that kernels wrote these tensors through raw pointers (bumps ``Tensor._version``): autograd's saved-tensor
    checks and the model's derived-weight caches key on it."""
    ts = [t for t in tensors if t is not None]
    if ts:
        torch.autograd.graph.increment_version(ts)


def _ptr_fields(spec):
    out = []
    for name, n in spec:
        out.append((name, c_void_p if n == 0 else c_void_p * n))
    return out


_PARAM_SPEC = [("user_table", 0), ("item_table", 0), ("cat_table", MAX_CAT), ("w0", 0), ("b0", 0),
               ("res_w1", MAX_RES), ("res_b1", MAX_RES), ("res_g1", MAX_RES), ("res_be1", MAX_RES),
               ("res_rm1", MAX_RES), ("res_rv1", MAX_RES), ("res_nbt1", MAX_RES),
               ("res_w2", MAX_RES), ("res_b2", MAX_RES), ("res_g2", MAX_RES), ("res_be2", MAX_RES),
               ("res_rm2", MAX_RES), ("res_rv2", MAX_RES), ("res_nbt2", MAX_RES),
               ("cross_w", MAX_CROSS), ("cross_b", MAX_CROSS), ("wf", 0), ("bf", 0)]
_GRAD_SPEC = [("user_table", 0), ("item_table", 0), ("cat_table", MAX_CAT), ("w0", 0), ("b0", 0),
              ("res_w1", MAX_RES), ("res_b1", MAX_RES), ("res_g1", MAX_RES), ("res_be1", MAX_RES),
              ("res_w2", MAX_RES), ("res_b2", MAX_RES), ("res_g2", MAX_RES), ("res_be2", MAX_RES),
              ("cross_w", MAX_CROSS), ("cross_b", MAX_CROSS), ("wf", 0), ("bf", 0)]


class Params(Structure):
    _fields_ = _ptr_fields(_PARAM_SPEC)


class Grads(Structure):
    _fields_ = _ptr_fields(_GRAD_SPEC)


class Batch(Structure):
    _fields_ = [("user_ids", c_void_p), ("item_ids", c_void_p), ("cat_features", c_void_p),
                ("num_features", c_void_p), ("batch", c_int64)]


_SIGS = {
    "dcnr_abi_version": (c_int, []),
    "dcnr_last_error_string": (c_char_p, []),
    "dcnr_launch_count": (c_int64, [c_int]),
    "dcnr_gemm_timing_begin": (c_int, []),
    "dcnr_gemm_timing_end": (c_int, [POINTER(c_double), POINTER(c_int64), POINTER(c_double)]),
    "dcnr_workspace_bytes": (c_int64, [POINTER(Dims), c_int64, c_int]),
    "dcnr_forward_eval": (c_int, [POINTER(Dims), POINTER(Params), POINTER(Batch), c_void_p, c_void_p, c_int64, c_void_p]),
    "dcnr_forward_train": (c_int, [POINTER(Dims), POINTER(Params), POINTER(Batch), c_uint64, c_void_p, c_void_p,
                                   c_void_p, c_int64, c_void_p]),
    "dcnr_backward": (c_int, [POINTER(Dims), POINTER(Params), POINTER(Batch), c_void_p, c_void_p, c_int64,
                              POINTER(Grads), c_void_p, c_int64, c_void_p]),
    "dcnr_bce_with_logits": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dcnr_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_double,
                               c_double, c_int, c_int64, c_void_p]),
    "dcnr_check_ids": (c_int, [POINTER(Dims), POINTER(Batch), c_void_p, c_void_p]),
    "dcnr_embed_concat_fwd": (c_int, [POINTER(Dims), POINTER(Params), POINTER(Batch), c_void_p, c_int64, c_void_p]),
    "dcnr_embed_scatter_bwd": (c_int, [POINTER(Dims), POINTER(Batch), c_void_p, c_int64, POINTER(Grads), c_void_p,
                                       c_int64, c_void_p]),
    "dcnr_cross_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_int32, POINTER(c_void_p), POINTER(c_void_p),
                               c_void_p, c_int64, c_void_p]),
    "dcnr_cross_bwd_scratch_bytes": (c_int64, [c_int64, c_int32, c_int32]),
    "dcnr_cross_bwd": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_int32, POINTER(c_void_p), POINTER(c_void_p),
                               c_void_p, c_int64, c_void_p, c_int64, POINTER(c_void_p), POINTER(c_void_p), c_void_p,
                               c_int64, c_void_p]),
    "dcnr_linear_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                c_void_p, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "dcnr_linear_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "dcnr_cross_v2_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                  c_int64, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "dcnr_cross_v2_bwd_prep": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                       c_void_p, c_int64, c_int, c_int64, c_int32, c_void_p]),
    "dcnr_linear_dgrad": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                  c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "dcnr_linear_wgrad_scratch_bytes": (c_int64, [c_int64, c_int32, c_int32]),
    "dcnr_linear_wgrad": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32,
                                  c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "dcnr_bn_scratch_bytes": (c_int64, [c_int64, c_int32]),
    "dcnr_bn_stats": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_float, c_float, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "dcnr_bn_act_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                c_float, c_uint64, c_uint32, c_void_p, c_int64, c_int64, c_int32, c_void_p]),
    "dcnr_bn_act_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                c_float, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                c_int32, c_void_p, c_int64, c_void_p]),
    "dcnr_tower_eval_supported": (c_int, [POINTER(Dims)]),
    "dcnr_tower_pack_bytes": (c_int64, [POINTER(Dims)]),
    "dcnr_tower_prepare": (c_int, [POINTER(Dims), POINTER(Params), c_int32, c_void_p, c_int64, c_void_p]),
    "dcnr_tower_eval_workspace_bytes": (c_int64, [POINTER(Dims)]),
    "dcnr_tower_eval": (c_int, [POINTER(Dims), POINTER(Params), c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                c_void_p, c_void_p, c_int64, c_void_p]),
    "dcnr_rowdot_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "dcnr_knn_normalize": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "dcnr_knn_scratch_bytes": (c_int64, [c_int64, c_int32, c_int32, c_int32]),
    "dcnr_knn_topk": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_int64, c_void_p, c_void_p,
                              c_void_p, c_int64, c_void_p]),
    "dcnr_knn_tc_supported": (c_int, [c_int64, c_int32, c_int32, c_int32]),
    "dcnr_knn_tc_scratch_bytes": (c_int64, [c_int64, c_int32, c_int32, c_int32]),
    "dcnr_knn_topk_tc": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_int32, c_int32, c_int64, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_void_p, c_void_p]),
    "dcnr_knn_merge": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "dcnr_mmr_rerank": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_float, c_int32, c_int32,
                                c_void_p, c_void_p, c_void_p]),
    "dcnr_comm_unique_id": (c_int, [c_void_p]),
    "dcnr_comm_create": (c_int, [c_void_p, c_int32, c_int32, POINTER(c_void_p)]),
    "dcnr_comm_destroy": (c_int, [c_void_p]),
    "dcnr_comm_info": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int32)]),
    "dcnr_comm_uses_peer_memory": (c_int, [c_void_p]),
    "dcnr_comm_set_peer_memory": (c_int, [c_void_p, c_int32]),
    "dcnr_comm_allreduce_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "dcnr_comm_allgather": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "dcnr_comm_alltoallv": (c_int, [c_void_p, c_void_p, POINTER(c_int64), POINTER(c_int64), c_void_p, POINTER(c_int64),
                                    POINTER(c_int64), c_void_p]),
    "dcnr_gather_rows": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p, c_void_p]),
    "dcnr_scatter_rows": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                  c_void_p]),
}

EXPORTS = tuple(_SIGS)
_lib = None


def lib():
    """The loaded shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C <package>/csrc).  This package has no CPU / eager fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        if L.dcnr_abi_version() != ABI_VERSION:
            raise RuntimeError("libdcnr_sm100a.so ABI version mismatch")
        _lib = L
    return _lib


def last_error() -> str:
    return (lib().dcnr_last_error_string() or b"").decode()


def check(status: int) -> None:
    """Turn a dcnr_status into the exception the reference's callers would have seen."""
    if status == OK:
        return
    msg = last_error()
    if status == ERR_INDEX:
        raise IndexError(msg)
    if "more than 1 value per channel" in msg:
        raise ValueError(msg)
    raise RuntimeError(f"libdcnr_sm100a: {msg} (status {status})")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dcnr_b200 runs on CUDA tensors only (sm_100a kernels, no CPU fallback); "
                               "move the model and its inputs to a B200 with .cuda()")


def launch_count(reset: bool = False) -> int:
    return int(lib().dcnr_launch_count(1 if reset else 0))


def gemm_timing_begin() -> None:
    check(lib().dcnr_gemm_timing_begin())


def gemm_timing_end():
    """(summed milliseconds, launches, algorithmic flops) of the GEMM launches since gemm_timing_begin()."""
    ms, n, fl = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
    check(lib().dcnr_gemm_timing_end(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl)))
    return ms.value, n.value, fl.value


def pad_dim(d: int) -> int:
    return (d + PAD - 1) // PAD * PAD
