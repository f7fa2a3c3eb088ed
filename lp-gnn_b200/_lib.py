"""ctypes binding of ``liblpgnn.so`` (C ABI: ``include/lpgnn.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C lp-gnn_b200/csrc`` and
loaded lazily on first use.  There is NO fallback: if the shared object is missing, or no sm_100
device is present, every op raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart.so.12 into the process before liblpgnn.so is opened)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblpgnn.so")

F32, BF16 = 0, 1
BWD_TAIL, BWD_REST = 1, 2
F16 = 2      # IEEE half storage (the reference's --fp16 mode): inference entry points only
EPI_NONE, EPI_RELU = 0, 1
COO_SORTED = 1
GRAPH_MEAN = 4
STATUS_LP_MAJOR = 8
WS_X2 = 16

_p = C.c_void_p
_i32, _i64, _sz, _int = C.c_int32, C.c_int64, C.c_size_t, C.c_int

MAX_HIDDEN_LAYERS = 8


class GcnFcWeights(C.Structure):
    """Mirror of ``lpgnn_gcn_fc_weights`` (include/lpgnn.h)."""
    _fields_ = ([(k, _i32) for k in ("p", "q", "hids", "depth", "precision", "reserved")]
                + [(k, _p) for k in ("c1_l2r_wrel", "c1_l2r_b", "c1_l2r_wroot", "c1_r2l_wrel", "c1_r2l_b", "c1_r2l_wroot",
                                     "c1_l2r_wcat", "c1_r2l_wcat")]
                + [(k, _p * MAX_HIDDEN_LAYERS) for k in ("l2r_wrel", "l2r_wroot", "l2r_b", "r2l_wrel", "r2l_wroot", "r2l_b")]
                + [(k, _p) for k in ("head_left_w", "head_left_b", "head_right_w", "head_right_b")]
                + [(f"{d}_{k}", _p * MAX_HIDDEN_LAYERS) for d in ("l2r", "r2l")
                   for k in ("wrel_hi", "wrel_lo", "wroot_hi", "wroot_lo", "wscale")])


class EpilogueArgs(C.Structure):
    """Mirror of ``lpgnn_epilogue_args`` (include/lpgnn.h)."""
    _fields_ = [("epilogue", _i32), ("dropout_p", C.c_float), ("dropout_seed", C.c_uint64), ("mask_act", _p),
                ("mask_scale", C.c_float)]


class GcnFcGrads(C.Structure):
    """Mirror of ``lpgnn_gcn_fc_grads`` (include/lpgnn.h)."""
    _fields_ = ([(k, _p) for k in ("c1_l2r_wrel", "c1_l2r_b", "c1_l2r_wroot", "c1_r2l_wrel", "c1_r2l_b", "c1_r2l_wroot")]
                + [(k, _p * MAX_HIDDEN_LAYERS) for k in ("l2r_wrel", "l2r_wroot", "l2r_b", "r2l_wrel", "r2l_wroot", "r2l_b")]
                + [(k, _p) for k in ("head_left_w", "head_left_b", "head_right_w", "head_right_b")])


# name -> (restype, argtypes); must list every symbol include/lpgnn.h declares
SIGNATURES = {
    "lpgnn_version": (_int, []),
    "lpgnn_last_error": (C.c_char_p, []),
    "lpgnn_launch_count": (C.c_uint64, []),
    "lpgnn_device_info": (_int, [C.POINTER(_int), C.POINTER(_int), C.POINTER(_int)]),
    "lpgnn_graph_build_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "lpgnn_graph_build": (_int, [_p, _p, _int, _p, _i64, _i32, _i32, _int, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "lpgnn_set_graph_fused": (_int, [_int]),
    "lpgnn_set_graph_compact": (_int, [_int]),
    "lpgnn_copy_many_h2d": (_int, [_p, _p, _p, _i32, _p]),
    "lpgnn_pack_scatter": (_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i64, _p, _p, _p, _p, _p, _p]),
    "lpgnn_pack_offsets": (_int, [_p, _p, _i64, _p, _p, _p, _i32, _p]),
    "lpgnn_spmm": (_int, [_p, _p, _p, _i32, _p, _p, _i32, _int, _p]),
    "lpgnn_spmm_ex": (_int, [_p, _p, _p, _i32, _p, _p, _i32, _int, _int, _int, _p]),
    "lpgnn_conv_in_zcat_width": (_i32, [_i32, _i32]),
    "lpgnn_gather_cat": (_int, [_p, _p, _p, _i32, _p, _i32, _p, _i32, _p, _p, _p]),
    "lpgnn_conv_in_fused": (_int, [_p, _p, _p, _i32, _p, _i32, _p, _i32, _p, _p, _p, _i32, _p, _int, _int, _p, _p]),
    "lpgnn_conv_in_16": (_int, [_p, _p, _p, _i32, _p, _p, _p, _p, _p, _i32, _p, _int, _int, _p, _p]),
    "lpgnn_set_conv_in_regb": (_int, [_int]),
    "lpgnn_conv_in_16_pair": (_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _i32, _p, _p, _int, _int,
                                     _p, _p, _p]),
    "lpgnn_node_transform": (_int, [_p, _i32, _p, _p, _i32, _p, _p, _i32, _i32, _p, _int, _int, _int, _p]),
    "lpgnn_node_transform_ex": (_int, [_p, _i32, _p, _p, _i32, _p, _p, _i32, _i32, _p, _int, C.POINTER(EpilogueArgs), _p]),
    "lpgnn_split_bf16": (_int, [_p, _i64, _int, _p, _p]),
    "lpgnn_gather_cat_ex": (_int, [_p, _p, _p, _i32, _p, _i32, _p, _i32, _p, _p, _int, _p]),
    "lpgnn_node_transform_head_ex": (_int, [_p, _i32, _p, _p, _i32, _p, _p, _i32, _i32, _p, _int, _int, _p, _p, _p]),
    "lpgnn_node_transform_split": (_int, [_int, _p, _i32, _p, _p, _i32, _p, _p, _i32, _i32, _p, _int, _p]),
    "lpgnn_split_x2": (_int, [_p, _i32, _p, _i32, _i64, _p, _p, _p, _p, _p, _p]),
    "lpgnn_node_transform_x2": (_int, [_p, _p, _i32, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _i32, _i32, _p, _int, _p, _p, _p]),
    "lpgnn_set_x2_chunk": (_int, [_int]),
    "lpgnn_conv_in_fused_x2": (_int, [_p, _p, _p, _i32, _p, _i32, _p, _i32, _p, _p, _p, _i32, _p, _int, _p, _p, _p, _p, _p, _p]),
    "lpgnn_spmm_x2": (_int, [_p, _p, _p, _i32, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "lpgnn_set_spmm_pair": (_int, [_int]),
    "lpgnn_spmm_pair": (_int, [_p, _p, _p, _i32, _p, _p, _p, _i32, _i64, _p, _p, _p, _p, _i32, _int, _p]),
    "lpgnn_spmm_x2_pair": (_int, [_p, _p, _p, _i32, _p, _p, _p, _i32, _i64, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "lpgnn_node_transform_head_parts": (_i32, [_i32]),
    "lpgnn_node_transform_head": (_int, [_p, _i32, _p, _p, _i32, _p, _p, _i32, _i32, _p, _int, _p, _p, _p]),
    "lpgnn_head_finish": (_int, [_p, _i32, _i32, _p, _p, _i32, _p, _p]),
    "lpgnn_head_finish_ex": (_int, [_p, _i32, _i32, _p, _p, _i32, _p, _p, _p]),
    "lpgnn_node_transform_head_train": (_int, [_p, _i32, _p, _p, _i32, _p, _p, _i32, _i32, _p, C.POINTER(EpilogueArgs), _p, _p, _p]),
    "lpgnn_head_mask": (_int, [_p, _int, _i32, _i32, _p, _p, _p, _i32, _p, _p, _p]),
    "lpgnn_add_knowledge": (_int, [_p, _i32, _p, _i32, _p, _p]),
    "lpgnn_predict_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32, _i32, _i32, _int]),
    "lpgnn_predict_basis": (_int, [C.POINTER(GcnFcWeights), _p, _p, _p, _i64, _i32, _i32, _int, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "lpgnn_predict_basis_packed": (_int, [C.POINTER(GcnFcWeights), _p, _p, _p, _i64, _i32, _i32, _int, _p, _p, _p, _p, _i32, _p,
                                          _p, _p, _p, _sz, _p]),
    "lpgnn_set_predict_fork": (_int, [_int]),
    "lpgnn_basis_select_segmented_ex": (_int, [_p, _p, _p, _p, _i32, _i32, _i32, _p, _int, _int, _p, _sz, _p]),
    "lpgnn_basis_select_segmented": (_int, [_p, _p, _p, _p, _i32, _i32, _i32, _p, _int, _p, _sz, _p]),
    "lpgnn_gemm_tn_splits": (_i32, [_i32, _i32, _i32]),
    "lpgnn_gemm_tn_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "lpgnn_gemm_tn": (_int, [_p, _p, _i32, _i32, _i32, _p, _p, _sz, _p]),
    "lpgnn_wgrad_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "lpgnn_wgrad": (_int, [_p, _p, _i64, _i32, _i32, _p, _p, _sz, _p]),
    "lpgnn_head_mask_bwd": (_int, [_p, _p, _p, _int, _i32, _i32, _p, C.c_float, _p, _p, _p, _p]),
    "lpgnn_head_mask_bwd_colsum_workspace_bytes": (_sz, [_i32, _i32]),
    "lpgnn_head_mask_bwd_colsum": (_int, [_p, _p, _p, _int, _i32, _i32, _p, C.c_float, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "lpgnn_relu_bwd": (_int, [_p, _p, _p, _i64, _int, C.c_float, _p, _p]),
    "lpgnn_dropout": (_int, [_p, _i64, _int, C.c_float, C.c_uint64, _p]),
    "lpgnn_transpose": (_int, [_p, _int, _i64, _i64, _p, _i64, _p]),
    "lpgnn_colsum_workspace_bytes": (_sz, [_i64, _i32]),
    "lpgnn_colsum": (_int, [_p, _int, _i64, _i32, _p, _p, _sz, _p]),
    "lpgnn_small_wgrad_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "lpgnn_small_wgrad": (_int, [_p, _int, _p, _i32, _i32, _i64, _i32, _p, _p, _p, _sz, _p]),
    "lpgnn_train_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32, _i32, _int]),
    "lpgnn_train_forward": (_int, [C.POINTER(GcnFcWeights), _p, _p, _p, _p, _p, _p, _i32, _i32, _p, _p, C.c_float, C.c_uint64,
                                   _p, _p, _p, _sz, _p]),
    "lpgnn_train_backward": (_int, [C.POINTER(GcnFcWeights), _p, _p, _p, _p, _p, _p, _i32, _i32, C.c_float, _p, _p,
                                    C.POINTER(GcnFcGrads), _p, _sz, _p]),
    "lpgnn_train_backward_ex": (_int, [C.POINTER(GcnFcWeights), _p, _p, _p, _p, _p, _p, _i32, _i32, C.c_float, _p, _p,
                                       C.POINTER(GcnFcGrads), _int, _p, _sz, _p]),
    "lpgnn_set_gemm_cluster": (_int, [_int]),
    "lpgnn_sample_mark": (_int, [_p, _p, _p, _i32, _i32, C.c_uint64, _p, _p]),
    "lpgnn_induced_count": (_int, [_p, _p, _p, _i32, _p, _p, _p]),
    "lpgnn_induced_fill": (_int, [_p, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "lpgnn_sample_sizes_len": (_i32, []),
    "lpgnn_sample_nodes_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "lpgnn_sample_nodes": (_int, [_p, _p, _p, _p, _i32, _i32, _p, _i32, _p, _i32, C.c_uint64, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "lpgnn_induced_offsets": (_int, [_p, _p, _p, _i32, _p, _p, _p, _p, _sz, _p]),
    "lpgnn_induced_fill_sorted": (_int, [_p, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "lpgnn_sample_gather": (_int, [_p, _p, _p, _p, _p, _i32, _p, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "lpgnn_lp_features_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "lpgnn_lp_features": (_int, [_p] * 11 + [_i64, _i32, _i32] + [_p] * 10 + [_p, _sz, _p]),
    "lpgnn_balanced_ce_workspace_bytes": (_sz, [_i32, _i32]),
    "lpgnn_balanced_ce": (_int, [_p, _p, _i32, _p, _p, _i32, _int, _p, _p, _p, _p, _sz, _p]),
    "lpgnn_balanced_ce_segmented_workspace_bytes": (_sz, [_i32]),
    "lpgnn_balanced_ce_segmented": (_int, [_p, _p, _p, _p, _p, _p, _i32, _int, _p, _p, _p, _p, _sz, _p]),
    "lpgnn_flat_ce_workspace_bytes": (_sz, [_i32, _i32]),
    "lpgnn_flat_ce": (_int, [_p, _p, _i32, _p, _p, _i32, _int, C.c_float, _p, _p, _p, _p, _p, _sz, _p]),
    "lpgnn_basis_metrics": (_int, [_p, _int, _p, _i32, _p, _i32, _p, _p]),
    "lpgnn_set_select_fused": (_int, [_int]),
    "lpgnn_basis_select_workspace_bytes": (_sz, [_i64]),
    "lpgnn_basis_select": (_int, [_p, _i32, _p, _i32, _i32, _p, _int, _p, _p, _sz, _p]),
}

_lib = None


def load() -> C.CDLL:
    """Opens liblpgnn.so and binds every entry point.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"liblpgnn.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C lp-gnn_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().lpgnn_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    if dt == torch.float16:
        return F16
    raise TypeError(f"lpgnn kernels take float32, bfloat16 or float16 features, got {dt}")


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("lpgnn ops run on a CUDA (sm_100) device only; got a CPU tensor. "
                               "There is no CPU fallback.")
