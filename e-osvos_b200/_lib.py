"""ctypes binding of libeosvos_b200.so (C ABI: include/eosvos_b200.h).

The library is the product: if it is missing or the device is not sm_100, every call fails
loudly -- there is no CPU or eager-PyTorch fallback (BASELINE.json north_star).
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeosvos_b200.so")

FLAG_RELU = 1
FLAG_OUT_FP32 = 2
FLAG_RES_HALF = 4

_P = c_void_p
_I = c_int
_F = c_float
_L = c_longlong

# name -> argtypes (return type is always int unless listed in _RESTYPES)
SIGNATURES = {
    "eosvos_last_error": [],
    "eosvos_version": [],
    "eosvos_device_check": [_I],
    "eosvos_launch_count": [],
    "eosvos_act_dtype": [],
    "eosvos_conv2d_fprop": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "eosvos_conv2d_dgrad": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "eosvos_conv2d_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _I, _I, _P],
    "eosvos_gemm_wgrad": [_P, _P, _P, _L, _I, _I, _L, _I, _L, _L, _F, _I, _I, _I, _P],
    "eosvos_deconv2x2_fprop": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "eosvos_deconv2x2_dgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "eosvos_deconv2x2_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _F, _I, _I, _P],
    "eosvos_gn_stats": [_P, _P, _I, _I, _I, _P],
    "eosvos_gn_apply": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _P],
    "eosvos_gn_backward": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _F, _P],
    "eosvos_roi_align_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "eosvos_roi_align_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "eosvos_mask_targets": [_P, _P, _P, _I, _I, _I, _I, _P],
    "eosvos_mask_loss_lovasz": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "eosvos_mask_loss_bce": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "eosvos_mask_paste_threshold": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "eosvos_mask_to_bbox": [_P, _P, _I, _I, _I, _I, _P],
    "eosvos_jf_counts": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "eosvos_nms_scratch_bytes": [_I, _I],
    "eosvos_nms_segments": [_P, _P, _I, _I, _F, _P, _P, _P],
    "eosvos_rpn_scratch_bytes": [_I, _P, _I, _I],
    "eosvos_rpn_scratch_zero_bytes": [_I, _I],
    "eosvos_rpn_select": [_P, _P, _I, _I, _I, _P, _P, _I, _F, _F, _F, _P, _P, _P, _P, _P],
    "eosvos_rpn_postnms": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "eosvos_extend_boxes": [_P, _P, _P, _I, _I, _I, _F, _F, _F, _F, _F, _I, _I, _P, _P],
    "eosvos_det_top1": [_P, _P, _I, _I, _I, _P, _F, _F, _F, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P],
    "eosvos_roi_match": [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P, _P],
    "eosvos_rpn_anchor_match": [_P, _I, _P, _P, _I, _F, _F, _P, _P, _P, _P, _P],
    "eosvos_rpn_loss": [_P, _P, _P, _I, _I, _P, _I, _P, _P, _P, _P, _P, _F, _I, _P, _P, _P, _P],
    "eosvos_rpn_sparse_head": [_P, _P, _P, _P, _P, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _F, _P, _P, _P,
                               _P, _P, _P, _P, _P],
    "eosvos_rpn_sparse_scatter": [_P, _P, _P, _I, _I, _P, _I, _P, _P],
    "eosvos_roi_sample_scratch_bytes": [_I, _I],
    "eosvos_roi_sample": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P],
    "eosvos_roi_encode": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "eosvos_meta_update_chunk_elems": [],
    "eosvos_meta_update": [_P, _P, _I, _I, _P, _P],
    "eosvos_lr_grad": [_P, _P, _I, _I, _P],
    "eosvos_radam_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _F, _F, _F, _F, _F, _I, _F, _F, _I, _P],
    "eosvos_permute_cast": [_P, _P, _P, _P, _P, _I, _I, _P],
    "eosvos_permute_cast_multi_chunk_elems": [],
    "eosvos_permute_cast_multi": [_P, _P, _I, _P],
    "eosvos_weight_prep_tile_elems": [],
    "eosvos_weight_prep_multi": [_P, _P, _I, _P],
    "eosvos_affine_warp_cubic": [_P, _P, _P, _P, _I, _I, _I, _P],
    "eosvos_label_warp_nearest": [_P, _P, _P, _P, _I, _I, _I, _P],
    "eosvos_transform": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "eosvos_mask_resize_nearest": [_P, _P, _I, _I, _I, _I, _I, _P],
    "eosvos_im2col_stem": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "eosvos_maxpool_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "eosvos_maxpool_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "eosvos_subsample2": [_P, _P, _I, _I, _I, _I, _I, _P],
    "eosvos_sum2x2": [_P, _P, _I, _I, _I, _I, _P],
    "eosvos_relu_bwd": [_P, _P, _P, _L, _P],
    "eosvos_colsum": [_P, _P, _L, _I, _F, _P],
}
_RESTYPES = {"eosvos_nms_scratch_bytes": c_longlong, "eosvos_rpn_scratch_bytes": c_longlong,
             "eosvos_rpn_scratch_zero_bytes": c_longlong, "eosvos_roi_sample_scratch_bytes": c_longlong, "eosvos_last_error": c_char_p, "eosvos_launch_count": ctypes.c_ulonglong}


class EosvosError(RuntimeError):
    pass


_lib = None


def load():
    """Loads the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EosvosError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no fallback path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        _lib = lib
    return _lib


def last_error():
    return load().eosvos_last_error().decode("utf-8", "replace")


def call(name, *args):
    """Calls an int-returning entry point and raises EosvosError on a non-zero code."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise EosvosError(f"{name} failed ({rc}): {last_error()}")
    return rc


_replayed = 0


def add_replayed_launches(n):
    """Kernels of this library executed through a CUDA-graph replay (not visible to the C-side counter)."""
    global _replayed
    _replayed += int(n)


def launch_count():
    """Kernel launches issued by the library so far: direct launches + kernels replayed inside CUDA graphs."""
    return int(load().eosvos_launch_count()) + _replayed


_checked_devices = set()


def require_device(index):
    """Fails loudly unless cuda:<index> is an sm_100 device."""
    if index not in _checked_devices:
        call("eosvos_device_check", int(index))
        _checked_devices.add(index)
