"""ctypes binding of libfi_b200.so (include/fi_b200.h).

There is no CPU fallback anywhere in this package: if the library is missing it is built with nvcc, and if
that fails -- or a kernel is asked to run on a non-CUDA tensor -- the call raises.
"""
import ctypes as C
import os

import torch

from . import build as _build

_LIB = None

_P = C.c_void_p
_I = C.c_int
_F = C.c_float

# name -> (restype, argtypes); one row per declaration in include/fi_b200.h
SIGNATURES = {
    "fi_abi_version": (_I, []),
    "fi_last_error": (C.c_char_p, []),
    "fi_last_status": (_I, []),
    "fi_kernel_launches": (C.c_ulonglong, []),
    "fi_set_option": (_I, [_I, _I]),
    "fi_get_option": (_I, [_I]),
    "CropAndResizeLaucher": (None, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _P]),
    "CropAndResizeBackpropImageLaucher": (None, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "ROIPoolForwardLaucher": (_I, [_P, _F, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "ROIPoolBackwardLaucher": (_I, [_P, _F, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "_nms": (None, [_I, _P, _P, _F]),
    "fi_crop_and_resize_forward": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _I, _P]),
    "fi_crop_and_resize_forward_dual": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P]),
    "fi_crop_and_resize_backward": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    "fi_crop_and_resize_backward_multi": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    "fi_set_deterministic": (_I, [_I]),
    "fi_get_deterministic": (_I, []),
    "fi_crop_sets_forward": (_I, [_P, _I, _P]),
    "fi_crop_sets_backward": (_I, [_P, _I, _I, _P]),
    "fi_crop_sets_backward_workspace": (C.c_size_t, [_P, _I, _I, C.c_long]),
    "fi_crop_sets_backward_plan": (_I, [_P, _I, _I, C.c_long, _P, C.c_size_t, _P, _P]),
    "fi_crop_sets_backward_run": (_I, [_P, _P, _I, _I, _P]),
    "fi_crop_sets_backward_overflow": (_I, [_P, _P]),
    "fi_crop_taps": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "fi_roi_level": (_I, [_P, _I, _F, _F, _P, _P]),
    "fi_split_levels": (_I, [_P, _I, _P, _P, _P, _P, _P, _P]),
    "fi_split_levels_gather": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fi_segment_mean_forward": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "fi_segment_mean_backward": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "fi_segment_mean_forward_batch": (_I, [_P, _I, _I, _I, _P]),
    "fi_segment_mean_backward_batch": (_I, [_P, _I, _I, _I, _P]),
    "fi_spatial_order": (_I, [_P, _I, _I, _I, _P, _P]),
    "fi_segment_mean_forward_n": (_I, [_P, _P, _I, _P, _I, _I, _P, _P, _P]),
    "fi_segment_mean_backward_n": (_I, [_P, _P, _P, _I, _P, _I, _I, _P, _P]),
    "fi_sinkhorn": (_I, [_P, _P, _I, _I, _I, _F, _I, _P, _P, _P, _P]),
    "fi_sinkhorn_workspace": (C.c_size_t, [_I, _I, _I, _I]),
    "fi_sinkhorn_ws": (_I, [_P, _P, _I, _I, _I, _F, _I, _P, _P, _P, _P, C.c_size_t, _P]),
    "fi_buffer_update": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "fi_merge_stats": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "fi_merge_stats_backward": (_I, [_P, _P, _I, _I, _I, _F, _P, _P]),
    "fi_ot_head_prep": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "fi_ot_head_combine": (_I, [_P, _P, _I, _P, _P]),
    "fi_ot_head_dcritic": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P]),
    "fi_relu_mask": (_I, [_P, _P, C.c_long, _P]),
    "fi_col_sum": (_I, [_P, _I, _I, _P, _P]),
    "fi_ot_head_dsum": (_I, [_P, _P, _I, _I, _P, _P]),
    "fi_centre_tap_embed": (_I, [_P, C.c_long, _P, _P]),
    "fi_nms_batched": (_I, [_P, _I, _I, _F, _P, _P, _P, _P]),
    "fi_nms_batched_topk": (_I, [_P, _I, _I, _F, _I, _P, _P, _P, _P]),
    "fi_proposal_decode": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _F, _F, _P, _P, _P]),
    "fi_proposal_gather": (_I, [_P, _P, _P, _I, _I, _I, _F, _F, _P, _P, _P]),
    "fi_mask_targets": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "fi_detection_decode": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _F, _F, _F, _P, _P, _P, _P, _P]),
    "fi_roi_pool_forward": (_I, [_P, _F, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "fi_roi_pool_backward": (_I, [_P, _F, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "fi_roi_pool_forward_nhwc": (_I, [_P, _F, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "fi_roi_pool_backward_nhwc": (_I, [_P, _F, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "fi_peer_region_bytes": (C.c_size_t, [C.c_size_t]),
    "fi_peer_region_alloc": (_I, [C.c_size_t, _P]),
    "fi_peer_region_free": (_I, [_P]),
    "fi_peer_region_export": (_I, [_P, _P]),
    "fi_peer_region_import": (_I, [_P, _P]),
    "fi_peer_region_release": (_I, [_P]),
    "fi_peer_error": (_I, [_P, _P]),
    "fi_peer_allreduce_sum": (_I, [_P, _P, C.c_size_t, _I, _I, _P, C.c_size_t, _P]),
}


class CropSet(C.Structure):
    """struct fi_crop_set (include/fi_b200.h)."""
    _fields_ = [("grads", _P), ("grads2", _P), ("boxes", _P), ("box_ind", _P), ("src_row", _P),
                ("num_boxes", _I), ("crop_height", _I), ("crop_width", _I)]


class FwdSet(C.Structure):
    """struct fi_fwd_set."""
    _fields_ = [("image", _P), ("boxes", _P), ("box_ind", _P), ("dst_row", _P), ("crops", _P), ("crops_compact", _P),
                ("batch", _I), ("image_height", _I), ("image_width", _I), ("depth", _I), ("num_boxes", _I),
                ("crop_height", _I), ("crop_width", _I), ("extrapolation_value", _F), ("num_boxes_dev", _P)]


class BwdSet(C.Structure):
    """struct fi_bwd_set."""
    _fields_ = [("grads_image", _P), ("grads", _P), ("grads2", _P), ("boxes", _P), ("box_ind", _P), ("src_row", _P),
                ("batch", _I), ("image_height", _I), ("image_width", _I), ("depth", _I), ("num_boxes", _I),
                ("crop_height", _I), ("crop_width", _I), ("num_boxes_dev", _P)]


class SegSet(C.Structure):
    """struct fi_seg_set."""
    _fields_ = [("gt", _P), ("feat", _P), ("k", _I), ("k_dev", _P), ("mean", _P), ("cnt", _P), ("grad_mean", _P), ("grad_feat", _P)]


class BwdPlan(C.Structure):
    """struct fi_bwd_plan (opaque)."""
    _fields_ = [("opaque", C.c_ulonglong * 640)]


def lib():
    """Load (building first if stale/missing) libfi_b200.so.  Raises if it cannot be had."""
    global _LIB
    if _LIB is None:
        path = _build.build_library()
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here == ABI drift; fail loudly
            fn.restype, fn.argtypes = res, args
        _LIB = handle
    return _LIB


def library_path():
    return _build.LIB_PATH


class FiError(RuntimeError):
    pass


def check(status):
    if status != 0:
        raise FiError("libfi_b200 status %d: %s" % (status, lib().fi_last_error().decode()))


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Refuses anything that is not CUDA memory."""
    if t is None:
        return None
    if not t.is_cuda:
        raise FiError("libfi_b200 kernels need CUDA tensors; got a %s tensor (there is no CPU fallback)" % t.device)
    return t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise FiError("libfi_b200 kernels need CUDA tensors; got a %s tensor (there is no CPU fallback)" % t.device)


FI_LAYOUT_NCHW = 0
FI_LAYOUT_NHWC = 1

# fi_set_option keys (include/fi_b200.h)
OPTIONS = {"bwd_form": 0, "tile_shape": 1, "fwd_form": 2, "sinkhorn_generic": 3, "pix_cfg": 4, "pix_group": 5, "fwd_chunk": 6, "fwd_pair": 7, "fwd_sched": 8}
BWD_FORMS = {"pix": 0, "smem": 1, "fused": 2, "red": 3}


def set_option(name, value):
    """Process-wide kernel-selection switch (experiments / tests); returns the previous value."""
    if isinstance(value, str):
        value = BWD_FORMS[value]
    old = lib().fi_set_option(OPTIONS[name], int(value))
    if old < 0:
        raise FiError("fi_set_option(%s, %s): %s" % (name, value, lib().fi_last_error().decode()))
    return old


def get_option(name):
    return lib().fi_get_option(OPTIONS[name])


def layout_of(t):
    """(layout flag, tensor laid out that way) for a logical NCHW 4-D tensor."""
    if t.dim() != 4:
        raise FiError("expected a 4-D [N,C,H,W] tensor, got %s" % (tuple(t.shape),))
    if t.is_contiguous():
        # NB: a tensor with C == 1 or H == W == 1 is contiguous in BOTH formats; NCHW wins (same bytes)
        return FI_LAYOUT_NCHW, t
    if t.is_contiguous(memory_format=torch.channels_last):
        return FI_LAYOUT_NHWC, t
    return FI_LAYOUT_NCHW, t.contiguous()
