"""ctypes binding of the C ABI declared in include/gmp_b200.h.

The product path has NO fallback: if libgmp_b200.so is missing or a call fails, an exception is
raised.  (The oracle under oracle/ is test infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libgmp_b200.so")

FP32_STRICT, BF16_TC = 0, 1

_lib = None
launches = 0  # number of C-ABI compute calls issued
_kernels = 0  # number of CUDA kernels those calls launched (bench.py's gpu_launches)
# kernels launched per entry point (default 1); memsets are not counted
_KERNELS_PER_CALL = {"gmp_exclusive_scan_i32": 3, "gmp_csr_fill": 3, "gmp_cells_build": 3, "gmp_tp_tc_contract": 2, "gmp_schnet_cfconv_fwd_tc2": 2, "gmp_schnet_cfconv_fwd_tc2_keep": 2, "gmp_egnn_tc2_edge_fwd": 2}


def kernel_launches() -> int:
    return _kernels

P, I32, I64, F32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class SchnetFilter(C.Structure):
    _fields_ = [("w1", P), ("b1", P), ("w2", P), ("b2", P), ("num_gaussians", I32), ("num_filters", I32),
                ("cutoff", F32), ("gauss_offset", P), ("gauss_coeff", F32)]


class NodeStage(C.Structure):
    _fields_ = [("w_img", P), ("bias", P), ("ln_g", P), ("ln_b", P), ("ln_eps", F32), ("act", I32), ("mul_aux", P), ("mul_mode", I32),
                ("add_res", P), ("out_f32", P), ("out_bf16", P), ("out_pre", P)]


class PeerRows(C.Structure):
    _fields_ = [("left", P), ("right", P), ("n_left", I64), ("n_own", I64)]


class EgnnParams(C.Structure):
    _fields_ = [(k, P) for k in ("wd", "ln1_g", "ln1_b", "w1", "b1", "ln2_g", "ln2_b", "w2", "b2", "ln3_g", "ln3_b",
                                 "w3", "b3")] + [("d", I32), ("act", I32), ("ln_eps", F32), ("aggr_mean", I32)]


_SIGS = {
    "gmp_radius_graph_count": [P, P, I64, I64, F32, I32, I32, P, P],
    "gmp_radius_graph_fill": [P, P, I64, I64, F32, I32, I32, P, P, P, P],
    "gmp_cells_build": [P, I64, F32, P, P, P, P, P, P, P],
    "gmp_radius_cells_count": [P, I64, F32, F32, P, P, P, P, I32, I32, P, P],
    "gmp_radius_cells_fill": [P, I64, F32, F32, P, P, P, P, I32, I32, P, P, P, P],
    "gmp_exclusive_scan_i32": [P, I64, P, P, P],
    "gmp_csr_count": [P, I64, I64, P, P],
    "gmp_csr_fill": [P, I64, I64, P, P, P, P, P],
    "gmp_gather_i64_to_i32": [P, P, I64, P, P],
    "gmp_index_is_sorted": [P, I64, P, P],
    "gmp_mark_unique_pairs": [P, P, P, I64, P, P],
    "gmp_compact_pairs": [P, P, P, P, P, I64, P, P, P],
    "gmp_segment_reduce_f32": [P, P, P, P, I64, I32, I32, P],
    "gmp_gather_mul_segsum_f32": [P, P, P, P, P, P, I64, I32, P],
    "gmp_gather_mul_segsum_wbf16": [P, P, P, P, I32, P, P, I64, I32, P],
    "gmp_gather_rows_f32": [P, P, P, I64, I32, P],
    "gmp_reduce_partials_f32": [P, I32, I64, P, P],
    "gmp_reduce_partials_batch_f32": [P, P, P, P, I32, P],
    "gmp_edge_length_fwd": [P, P, P, I64, P, P],
    "gmp_edge_geometry_fwd": [P, P, P, I64, I32, F32, I32, F32, P, P, P],
    "gmp_edge_length_bwd": [P, P, P, P, P, P, P, P, I64, P, P],
    "gmp_schnet_cfconv_fwd": [P, P, P, I64, I64, P, P, P, P, P, I32, P],
    "gmp_schnet_cfconv_bwd": [P, P, P, I64, I64, P, P, P, P, P, P, P, P, I32, P],
    "gmp_umma_selftest": [P, P, P, I32, P],
    "gmp_umma_selftest_mn": [P, P, P, I32, P],
    "gmp_egnn_edge_fwd": [P, P, I64, I64, P, P, P, P, P, P, I32, P],
    "gmp_egnn_edge_bwd": [P, P, P, I64, I64, P, P, P, P, P, P, I32, P, P, P, I32, P],
    "gmp_schnet_cfconv_fwd_tc2": [P, P, P, P, I64, I64, P, P, P, P, P, P],
    "gmp_schnet_cfconv_fwd_tc2_keep": [P, P, P, P, I64, I64, P, P, P, P, P, P, P, P],
    "gmp_schnet_cfconv_bwd_tc2": [P, P, P, P, I64, I64, P, P, P, P, P, I32, P],
    "gmp_linear_wgrad_tc": [P, P, I64, I32, I32, P, P],
    "gmp_egnn_tc_edge_fwd": [P, P, P, I64, I64, P, P, P, P, P, P, P],
    "gmp_egnn_tc2_edge_fwd": [P, P, P, I64, I64, P, P, P, P, P, P, P, P, P],
    "gmp_egnn_tc_edge_bwd_fused": [P, P, P, P, I64, I64, P, P, P, P, P, P, P, P, P, P, P, P, P],
    "gmp_segment_sum_bf16_f32": [P, P, P, P, I64, I32, P],
    "gmp_egnn_tc_edge_bwd": [P, P, P, P, I64, I64, P, P, P, P, P, P, I32, P, P, P, P],
    "gmp_tp_contract": [P, P, P, I64, I64, P, I32, P, I32, P, I32, P, I32, P, P, P, P, I32, P, P, I32, I32, P, I32, P],
    "gmp_symcontract_fwd": [P, P, P, P, I64, I32, I32, I32, I32, P, I32, P],
    "gmp_symcontract_bwd": [P, P, P, P, I64, I32, I32, I32, I32, P, I32, P, P, P],
    "gmp_tp_tc_pack_hid": [P, I64, P, I32, P, P, I32, P, P],
    "gmp_tp_tc_pack_w2": [P, I32, P, I32, P, P],
    "gmp_tp_ysum": [P, P, P, I64, I64, P, I32, P, I32, P, I32, I32, P, I32, P, P, I32, P],
    "gmp_tp_tc_contract": [P, P, P, I64, I64, P, I32, P, I32, P, P, I32, P, P, P, I32, I32, I32, P, P],
    "gmp_tp_tc_dhid": [P, P, P, P, I64, I64, P, I32, P, I32, P, I32, P, I32, P, P, P, P, I32, I32, I32, P, P, P],
    "gmp_tp_tc_dw2": [P, P, P, P, I64, I64, P, I32, P, I32, P, I32, P, P, I32, I32, P, I64, I32, P, P],
    "gmp_uvu_conv_fwd": [P, P, P, I64, I64, P, P, P, I32, P, P],
    "gmp_uvu_conv_dx": [P, P, P, I64, I64, P, P, P, I32, P, P],
    "gmp_uvu_conv_dw": [P, P, I64, P, P, P, I32, P, P],
    "gmp_halo_pull": [P, P, P, I32, I32, P],
    "gmp_gate_fwd": [P, P, I64, I32, I32, I32, F32, F32, P, P],
    "gmp_gate_bwd": [P, P, P, P, P, I64, I32, I32, I32, F32, F32, P, P],
    "gmp_node_pack_w": [P, I32, I32, I32, P, P],
    "gmp_node_chain_tc": [P, P, I64, I32, P, P],
    "gmp_node_pack_w_batch": [P, P, P, I32, P],
    "gmp_ln_act_bwd": [P, P, P, P, F32, I32, I64, P, P, P, P],
    "gmp_tp_wgrad": [P, P, P, I64, I64, P, I32, P, I32, P, I32, P, I32, P, P, P, I32, P, I32, P, P, P, P, I32, P],
}
_PLAIN = {"gmp_version": (I32, []), "gmp_last_error": (C.c_char_p, []),
          "gmp_schnet_bwd_num_parts": (I32, [I64]), "gmp_schnet_bwd_part_len": (I64, [I32, I32]),
          "gmp_egnn_bwd_num_parts": (I32, [I64]), "gmp_egnn_tc_bwd_num_parts": (I32, [I64]), "gmp_egnn_tc2_num_chunks": (I32, [I64]), "gmp_linear_wgrad_num_parts": (I32, [I64]), "gmp_schnet_tc2_num_chunks": (I32, [I64]), "gmp_egnn_bwd_part_len": (I64, [I32]),
          "gmp_tp_contract_smem_bytes": (I64, [I32, I32]), "gmp_symcontract_bwd_num_parts": (I32, [I64]), "gmp_symcontract_fast_path": (I32, [I32, I32, I32, I32, I32]), "gmp_tp_wgrad_part_len": (I64, [I32]),
          "gmp_node_w_image_bytes": (I64, [I32, I32, I32]), "gmp_ln_act_bwd_num_parts": (I32, [I64]), "gmp_tp_tc_num_chunks": (I32, [I64]), "gmp_tp_tc_hid_bytes": (I64, [I64, I32]), "gmp_tp_tc_w2_bytes": (I64, [I32, I32])}


def exported_symbols():
    """Every symbol include/gmp_b200.h declares (checked by tests/test_cabi_symbols.py)."""
    return sorted(list(_SIGS) + list(_PLAIN))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "gmp_b200 has no CPU or PyTorch fallback.")
        _lib = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = I32, args
        for name, (res, args) in _PLAIN.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


class GmpError(RuntimeError):
    pass


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Tensors must be contiguous CUDA tensors."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GmpError("gmp_b200 kernels need CUDA tensors; there is no CPU fallback")
    if not t.is_contiguous():
        raise GmpError("gmp_b200 kernels need contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    """Invoke a status-returning entry point on the current torch stream; raise on failure."""
    global launches, _kernels
    l = lib()
    rc = getattr(l, name)(*args, stream())
    launches += 1
    _kernels += _KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise GmpError(f"{name} failed ({rc}): {l.gmp_last_error().decode()}")
