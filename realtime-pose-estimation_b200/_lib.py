"""ctypes binding of csrc/libbrtpe.so (the C ABI declared in include/brtpe.h).

There is NO fallback: if the shared library has not been built, or CUDA is not
available, every product entry point raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C <pkg>/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# BRTPE_LIB: another build of the same library (A/B measurements of two builds on one GPU box)
LIB_PATH = os.environ.get("BRTPE_LIB") or os.path.join(_HERE, "csrc", "libbrtpe.so")

MAX_TAG_DIMS = 4
MAX_TOPK = 64
MAX_GROUP_K = 32
MAX_JOINTS = 32

DT_F32, DT_BF16, DT_BF16X2 = 0, 1, 2
ENGINE_AUTO, ENGINE_FFMA, ENGINE_UMMA, ENGINE_UMMA_HALO = 0, 1, 2, 3


class BrtpeError(RuntimeError):
    pass


class DecodeParams(C.Structure):
    """struct brtpe_decode_params (include/brtpe.h)."""
    _fields_ = [("num_joints", C.c_int32), ("max_num_people", C.c_int32),
                ("detection_threshold", C.c_double), ("tag_threshold", C.c_double),
                ("use_detection_val", C.c_int32), ("ignore_too_much", C.c_int32),
                ("tag_per_joint", C.c_int32), ("nms_ksize", C.c_int32),
                ("nms_padding", C.c_int32), ("munkres_start_rule", C.c_int32)]


class ConvDesc(C.Structure):
    """struct brtpe_conv_desc (include/brtpe.h)."""
    _fields_ = [("dtype", C.c_int32), ("engine", C.c_int32),
                ("N", C.c_int32), ("Hin", C.c_int32), ("Win", C.c_int32),
                ("Cin", C.c_int32), ("in_ld", C.c_int32), ("in_coff", C.c_int32),
                ("Hm", C.c_int32), ("Wm", C.c_int32), ("in_stride", C.c_int32),
                ("ntaps", C.c_int32), ("tap_dy", C.c_int32 * 9), ("tap_dx", C.c_int32 * 9),
                ("Hout", C.c_int32), ("Wout", C.c_int32), ("out_scale", C.c_int32),
                ("out_oy", C.c_int32), ("out_ox", C.c_int32), ("Cout", C.c_int32),
                ("out_ld", C.c_int32), ("out_coff", C.c_int32),
                ("res_ld", C.c_int32), ("res_coff", C.c_int32),
                ("relu", C.c_int32), ("Cout_store", C.c_int32),
                ("n_add", C.c_int32), ("add_ld", C.c_int32 * 3), ("add_shift", C.c_int32 * 3),
                ("out2_ld", C.c_int32), ("reverse_order", C.c_int32)]


class PrepackDesc(C.Structure):
    """struct brtpe_prepack_desc (include/brtpe.h)."""
    _fields_ = [("w_dtype", C.c_int32), ("transposed", C.c_int32),
                ("Cout", C.c_int32), ("Cin", C.c_int32), ("KH", C.c_int32), ("KW", C.c_int32),
                ("ntaps", C.c_int32), ("tap_kh", C.c_int32 * 9), ("tap_kw", C.c_int32 * 9),
                ("im2col", C.c_int32), ("Cin_store", C.c_int32), ("layout", C.c_int32),
                ("Cout_pack", C.c_int32), ("cin_pad", C.c_int32), ("cout_pad", C.c_int32),
                ("round_bf16", C.c_int32), ("bn_eps", C.c_float), ("split", C.c_int32)]


WT_F32, WT_BF16, WT_F16 = 0, 1, 2
PACK_CIN_COUT_F32, PACK_KMAJOR_BF16 = 0, 1

_lib = None

_P = C.c_void_p
_I = C.c_int
_SIGS = {
    # name: (restype, argtypes)
    "brtpe_last_error": (C.c_char_p, []),
    "brtpe_version": (_I, []),
    "brtpe_has_umma": (_I, []),
    "brtpe_nms": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "brtpe_topk_workspace_bytes": (C.c_size_t, [_I] * 5),
    "brtpe_nms_topk_gather": (_I, [_P, _P] + [_I] * 9 + [_P, _P, _P, _P, _P, C.c_size_t, _P]),
    "brtpe_group_workspace_bytes": (C.c_size_t, [_I] * 5),
    "brtpe_group_ae": (_I, [_P, _P, _P, _I, _I, _I, C.POINTER(DecodeParams), _P, _P, _P, _I, _P,
                            C.c_size_t, _P]),
    "brtpe_adjust": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "brtpe_scores": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "brtpe_refine_workspace_bytes": (C.c_size_t, [_I] * 4),
    "brtpe_refine": (_I, [_P, _P, _P, _P] + [_I] * 7 + [_P, C.c_size_t, _P]),
    "brtpe_bilinear_resize": (_I, [_P, C.c_longlong, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P]),
    "brtpe_preprocess_warp_normalize": (_I, [_P, _I, _I, _I, C.POINTER(C.c_double), _I, _I,
                                            C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P, _P]),
    "brtpe_aggregate_scale": (_I, [_P, _P, _P, _P] + [_I] * 9 + [C.POINTER(C.c_int32), _I,
                                                                 C.c_float, _P, _P, _P]),
    "brtpe_conv_run": (_I, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P]),
    "brtpe_conv_run_fused": (_I, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, C.POINTER(_P), _P, _P]),
    "brtpe_conv_chain_run": (_I, [C.POINTER(ConvDesc), _P, _P, _P, _P, C.POINTER(ConvDesc), _P, _P, _P, _P, _P]),
    "brtpe_plan_set_conv_fuse": (_I, [_P, C.POINTER(_P), _P]),
    "brtpe_conv_select_engine": (_I, [C.POINTER(ConvDesc)]),
    "brtpe_umma_weight_dims": (_I, [_I, _I, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "brtpe_prepack_weights": (_I, [C.POINTER(PrepackDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "brtpe_stem_conv1": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _P, _I, _P]),
    "brtpe_fuse_sum": (_I, [_I, _I, C.POINTER(_P), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                            _I, _I, _I, _I, _P, _I, _I, _P]),
    "brtpe_nhwc_to_nchw": (_I, [_I, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P]),
    "brtpe_plan_create": (_P, []),
    "brtpe_plan_destroy": (None, [_P]),
    "brtpe_plan_add_conv": (_I, [_P, C.POINTER(ConvDesc), _P, _P, _P, _P, _P]),
    "brtpe_plan_add_stem": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _P, _I]),
    "brtpe_plan_add_fuse": (_I, [_P, _I, _I, C.POINTER(_P), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32), _I, _I, _I, _I, _P, _I, _I]),
    "brtpe_plan_add_nhwc_to_nchw": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _P, _I]),
    "brtpe_plan_add_stem_im2col": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "brtpe_plan_set_sched": (_I, [_P, _I, C.POINTER(C.c_int32), _I]),
    "brtpe_stem_im2col": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "brtpe_plan_num_ops": (_I, [_P]),
    "brtpe_plan_conv_flops": (C.c_double, [_P]),
    "brtpe_plan_run": (_I, [_P, _P]),
    "brtpe_plan_graph_launch": (_I, [_P, _P]),
    "brtpe_debug_halo_prof": (_I, [_P, _I]),
    "brtpe_aux_run": (_I, [_I, _P, _P, _P, _P, C.POINTER(C.c_int32), _I, _P]),
    "brtpe_plan_add_aux": (_I, [_P, _I, _P, _P, _P, _P, C.POINTER(C.c_int32), _I]),
    "brtpe_plan_profile": (_I, [_P, _P, C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                C.POINTER(C.c_double)]),
}


def exported_symbols():
    """Names of every entry point include/brtpe.h declares."""
    return sorted(_SIGS)


def load(require_cuda: bool = True):
    """Load libbrtpe.so (once).  Raises BrtpeError if it is missing -- no fallback."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise BrtpeError(
                "libbrtpe.so is not built (%s). Run __graft_entry__.build() or "
                "`make -C realtime-pose-estimation_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)       # AttributeError if the ABI and header diverge
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    if require_cuda and not torch.cuda.is_available():
        raise BrtpeError("rtpe_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = _lib.brtpe_last_error().decode("utf-8", "replace") if _lib is not None else ""
        raise BrtpeError("%s failed (code %d): %s" % (what or "libbrtpe call", rc, msg))


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())
