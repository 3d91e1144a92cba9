"""ctypes binding of libmstcn_b200.so (include/mstcn_b200.h).

There is deliberately no fallback: if the CUDA library is missing this raises, so a GPU
test can never pass on a silent PyTorch/CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# MSTCN_B200_LIB points at another build of the same library (A/B timing of kernel variants); default: the in-tree build
LIB_PATH = os.environ.get("MSTCN_B200_LIB") or os.path.join(_HERE, "libmstcn_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mstcn_b200.h")


class MstcnDims(C.Structure):
    _fields_ = [("dim", C.c_int32), ("num_stages", C.c_int32), ("num_layers", C.c_int32),
                ("num_f_maps", C.c_int32), ("n_class", C.c_int32), ("flags", C.c_int32)]


FLAG_TENSOR_CORES = 1
FLAG_FFMA_BACKWARD = 2
FLAG_PACK_TC_ONLY = 4


class MstcnDropout(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("_pad", C.c_int32), ("seed", C.c_uint64), ("offset", C.c_uint64),
                ("offset_dev", C.c_void_p)]


class MstcnError(RuntimeError):
    pass


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_F = C.c_float
_DP = C.POINTER(MstcnDims)
_RP = C.POINTER(MstcnDropout)

# name -> (restype, argtypes); must list every function the header declares
_SIGNATURES = {
    "mstcn_abi_version": (C.c_int, []),
    "mstcn_last_error": (C.c_char_p, []),
    "mstcn_sm_count": (C.c_int, []),
    "mstcn_param_count": (_I64, [_DP]),
    "mstcn_packed_count": (_I64, [_DP]),
    "mstcn_param_offset": (_I64, [_DP, _I32]),
    "mstcn_param_tensors": (_I32, [_DP]),
    "mstcn_packed_offset": (_I64, [_DP, _I32, _I32, _I32]),
    "mstcn_pack_params": (C.c_int, [_DP, _P, _P, _P]),
    "mstcn_workspace_floats": (_I64, [_DP, _I32, _I32, _I32]),
    "mstcn_workspace_offset": (_I64, [_DP, _I32, _I32, _I32, _I32, _I32, _I32]),
    "mstcn_forward": (C.c_int, [_DP, _P, _P, _P, _P, _I32, _I32, _I32, _RP, _I32, _P, _P, _P, _P]),
    "mstcn_backward": (C.c_int, [_DP, _P, _P, _P, _P, _I32, _I32, _I32, _RP, _P, _P, _P, _P, _P, _I32, _P]),
    "mstcn_backward_stage": (C.c_int, [_DP, _P, _P, _P, _P, _I32, _I32, _I32, _RP, _P, _P, _P, _P, _P, _I32, _I32, _P]),
    "mstcn_bucket_boundary": (_I64, [_DP, _I32]),
    "mstcn_proj_fwd": (C.c_int, [_P, _I64, _I32, _P, _P, _P, _P]),
    "mstcn_proj_fwd_tc": (C.c_int, [_P, _I64, _I32, _P, _P, _P, _I32, _P, _P]),
    "mstcn_proj_bwd_scratch_floats": (_I64, [_I32]),
    "mstcn_proj_bwd": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P, _I32, _P]),
    "mstcn_proj_wgrad_tc_scratch_floats": (_I64, [_I32]),
    "mstcn_proj_wgrad_tc": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P, _P, _P, _I32, _P]),
    "mstcn_layer_fwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P, _RP, _I32, _P]),
    "mstcn_layer_fwd_tc": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _P, _P, _P, _RP, _I32, _P]),
    "mstcn_stage_fwd_tc": (C.c_int, [_DP, _P, _I32, _P, _P, _P, _I32, _I32, _RP, _P, _P]),
    "mstcn_layer_bwd_gx_tc": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _P, _P]),
    "mstcn_layer_bwd_scratch_floats": (_I64, []),
    "mstcn_layer_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P, _P, _RP, _I32,
                                  _P, _P, _P, _P, _P, _I32, _P]),
    "mstcn_tail_fwd": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mstcn_tail_bwd_scratch_floats": (_I64, []),
    "mstcn_tail_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P,
                                 _P, _P, _P, _P, _P, _P, _I32, _P]),
    "mstcn_ce_scratch_floats": (_I64, [_I64]),
    "mstcn_ce_loss": (C.c_int, [_P, _P, _I64, _I32, _I64, _P, _P, _P, _P]),
    "mstcn_loss_head": (C.c_int, [_DP, _P, _I32, _I32, _P, _I64, _P, _P, _P, _P, _P]),
    "mstcn_paper_loss_scratch_floats": (_I64, [_I32, _I64]),
    "mstcn_paper_loss": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I64, C.c_float, C.c_float, _P, _P, _P, _P]),
    "mstcn_pad_batch": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P]),
    "mstcn_frame_argmax": (C.c_int, [_P, _I64, _I32, _P, _P, _P]),
    "mstcn_segment_vote": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "mstcn_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _I32, _P]),
    "mstcn_adam_step_dev": (C.c_int, [_P, _P, _P, _P, _I64, _P, _F, _F, _F, _P, _P]),
    "mstcn_dp_allreduce": (C.c_int, [_P, _P, _P, _I64, _I64, _I32, _I32, _I32, _P]),
    "mstcn_dp_flag_words": (_I64, []),
    "mstcn_debug_tc_timing": (C.c_int, [_P]),
    "mstcn_debug_chain_trace": (C.c_int, [_P]),
    "mstcn_debug_trap_report": (C.c_int, [_P]),
    "mstcn_debug_backward_timing": (C.c_int, [_I32]),
    "mstcn_debug_backward_times": (C.c_int, [C.POINTER(C.c_float), _I32]),
    "mstcn_dropout_scale": (C.c_int, [_RP, _I32, _I64, _P, _P]),
}

_lib = None


def header_symbols():
    """Function names declared in include/mstcn_b200.h (used by the symbol-export test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mstcn_[a-z0-9_]+)\s*\(", text)))


def lib():
    """The loaded library; raises MstcnError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MstcnError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU or PyTorch fallback for the MS-TCN hot path.")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    if handle.mstcn_abi_version() != 2:
        raise MstcnError("libmstcn_b200.so ABI version mismatch; rebuild")
    _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().mstcn_last_error()
        raise MstcnError(msg.decode() if msg else f"mstcn call failed with code {rc}")


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
