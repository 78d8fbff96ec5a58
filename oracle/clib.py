"""TEST INFRASTRUCTURE ONLY -- ctypes access to the CPU oracles.

Two families, both numpy in / numpy out:

* ``oracle_*``  our C restatement, ``oracle/fi_oracle.c`` -> ``oracle/_build/liboracle.so``
* ``ref_*``     the reference's OWN C sources compiled unmodified (``oracle/Makefile`` target ``ref``)
                -> ``oracle/_ref/libref_roi_align.so`` (lib/roi_align/src/crop_and_resize.c) and
                ``oracle/_ref/libref_nms.so`` (lib/nms/src/nms.c), driven through the TH shim.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module.  The product package (``feature_intertwiner_b200``) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_build", "liboracle.so")
_REF_DIR = os.path.join(_HERE, "_ref")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(ref=None):
    """Compile the oracle (always) and oracle/_ref (only where /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    if ref is None:
        ref = os.path.isdir("/root/reference")
    if ref:
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_oracle = None


def lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(_ORACLE_SO):
            build(ref=False)
        L = C.CDLL(_ORACLE_SO)
        L.fi_oracle_crop_and_resize_fwd.restype = C.c_int
        L.fi_oracle_crop_and_resize_fwd.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _i32p, C.c_int,
                                                    C.c_int, C.c_int, C.c_float, _f32p]
        L.fi_oracle_crop_and_resize_bwd.restype = C.c_int
        L.fi_oracle_crop_and_resize_bwd.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _i32p, C.c_int,
                                                    C.c_int, C.c_int, _f32p]
        L.fi_oracle_crop_taps.restype = None
        L.fi_oracle_crop_taps.argtypes = [C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _i32p]
        L.fi_oracle_crop_unique_pixels.restype = C.c_long
        L.fi_oracle_crop_unique_pixels.argtypes = [C.c_int, C.c_int, C.c_int, _f32p, _i32p, C.c_int, C.c_int, C.c_int, _u8p]
        L.fi_oracle_roi_level.restype = None
        L.fi_oracle_roi_level.argtypes = [_f32p, C.c_int, C.c_float, C.c_float, _i32p, C.c_void_p]
        L.fi_oracle_segment_mean.restype = None
        L.fi_oracle_segment_mean.argtypes = [_i32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p]
        L.fi_oracle_sinkhorn.restype = C.c_double
        L.fi_oracle_sinkhorn.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        L.fi_oracle_nms.restype = C.c_int
        L.fi_oracle_nms.argtypes = [_f32p, C.c_int, C.c_float, C.c_int, _i32p]
        L.fi_oracle_roi_pool_fwd.restype = None
        L.fi_oracle_roi_pool_fwd.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int,
                                             C.c_float, _f32p, _i32p]
        L.fi_oracle_roi_pool_bwd.restype = None
        L.fi_oracle_roi_pool_bwd.argtypes = [_f32p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int,
                                             C.c_int, C.c_float, _f32p]
        _oracle = L
    return _oracle


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ----------------------------------------------------------------------------- our restatement
def oracle_crop_and_resize_fwd(image, boxes, box_ind, ph, pw, extrapolation=0.0):
    image, boxes, box_ind = _c(image, np.float32), _c(boxes, np.float32).reshape(-1, 4), _c(box_ind, np.int32)
    B, Cc, H, W = image.shape
    R = boxes.shape[0]
    out = np.empty((R, Cc, ph, pw), np.float32)
    lib().fi_oracle_crop_and_resize_fwd(image, B, Cc, H, W, boxes, box_ind, R, ph, pw, extrapolation, out)
    return out


def oracle_crop_and_resize_bwd(grads, boxes, box_ind, im_size):
    grads, boxes, box_ind = _c(grads, np.float32), _c(boxes, np.float32).reshape(-1, 4), _c(box_ind, np.int32)
    B, Cc, H, W = im_size
    R, _, ph, pw = grads.shape
    out = np.empty((B, Cc, H, W), np.float32)
    lib().fi_oracle_crop_and_resize_bwd(grads, B, Cc, H, W, boxes, box_ind, R, ph, pw, out)
    return out


def oracle_crop_taps(H, W, boxes, ph, pw):
    boxes = _c(boxes, np.float32).reshape(-1, 4)
    taps = np.empty((boxes.shape[0], ph, pw, 5), np.int32)
    lib().fi_oracle_crop_taps(H, W, boxes, boxes.shape[0], ph, pw, taps)
    return taps


def oracle_unique_pixels(B, H, W, boxes, box_ind, ph, pw):
    boxes, box_ind = _c(boxes, np.float32).reshape(-1, 4), _c(box_ind, np.int32)
    scratch = np.empty(B * H * W, np.uint8)
    return int(lib().fi_oracle_crop_unique_pixels(B, H, W, boxes, box_ind, boxes.shape[0], ph, pw, scratch))


def oracle_roi_level(rois, image_area, base=224.0):
    rois = _c(rois, np.float32).reshape(-1, 4)
    level = np.empty(rois.shape[0], np.int32)
    pre = np.empty(rois.shape[0], np.float32)
    lib().fi_oracle_roi_level(rois, rois.shape[0], image_area, base, level, pre.ctypes.data)
    return level, pre


def oracle_segment_mean(gt, feat, ncls=81):
    gt, feat = _c(gt, np.int32), _c(feat, np.float32)
    k, F = feat.shape
    out = np.empty((F, ncls), np.float32)
    cnt = np.empty((1, ncls), np.float32)
    lib().fi_oracle_segment_mean(gt, feat, k, F, ncls, out, cnt)
    return out, cnt


def oracle_sinkhorn(x, y, inv_eps=1.0, L=5, wide=False, want_plan=False, want_grad=False):
    x, y = _c(x, np.float32), _c(y, np.float32)
    N, D = x.shape
    P = np.empty((N, N), np.float32) if want_plan else None
    gx = np.empty((N, D), np.float32) if want_grad else None
    gy = np.empty((N, D), np.float32) if want_grad else None
    loss = lib().fi_oracle_sinkhorn(x, y, N, D, inv_eps, L, int(wide),
                                    P.ctypes.data if want_plan else None,
                                    gx.ctypes.data if want_grad else None,
                                    gy.ctypes.data if want_grad else None)
    return loss, P, gx, gy


def oracle_nms(boxes_xyxys, thresh, strict):
    """boxes[n,5]=(x1,y1,x2,y2,score) sorted by descending score.  strict=True: GPU rule (>)."""
    b = _c(boxes_xyxys, np.float32).reshape(-1, 5)
    keep = np.empty(max(b.shape[0], 1), np.int32)
    n = lib().fi_oracle_nms(b, b.shape[0], thresh, int(strict), keep)
    return keep[:n].copy()


def oracle_roi_pool_fwd(feat, rois, ph, pw, scale):
    feat, rois = _c(feat, np.float32), _c(rois, np.float32).reshape(-1, 5)
    B, Cc, H, W = feat.shape
    R = rois.shape[0]
    top = np.empty((R, Cc, ph, pw), np.float32)
    arg = np.empty((R, Cc, ph, pw), np.int32)
    lib().fi_oracle_roi_pool_fwd(feat, B, Cc, H, W, rois, R, ph, pw, scale, top, arg)
    return top, arg


def oracle_roi_pool_bwd(top_diff, argmax, rois, feat_size, scale):
    top_diff, argmax, rois = _c(top_diff, np.float32), _c(argmax, np.int32), _c(rois, np.float32).reshape(-1, 5)
    B, Cc, H, W = feat_size
    R, _, ph, pw = top_diff.shape
    out = np.empty((B, Cc, H, W), np.float32)
    lib().fi_oracle_roi_pool_bwd(top_diff, argmax, B, Cc, H, W, rois, R, ph, pw, scale, out)
    return out


# ----------------------------------------------------------------------------- the reference itself
class _THTensor(C.Structure):
    _fields_ = [("size", C.c_long * 4), ("data", C.c_void_p), ("itemsize", C.c_long), ("capacity", C.c_long)]


def _th(a):
    t = _THTensor()
    for i in range(4):
        t.size[i] = a.shape[i] if i < a.ndim else 0
    t.data = a.ctypes.data
    t.itemsize = a.itemsize
    t.capacity = a.size
    return t


def have_ref():
    return all(os.path.exists(os.path.join(_REF_DIR, n)) for n in ("libref_roi_align.so", "libref_nms.so"))


_ref_ra = None
_ref_nms = None


def _ref_roi_align():
    global _ref_ra
    if _ref_ra is None:
        _ref_ra = C.CDLL(os.path.join(_REF_DIR, "libref_roi_align.so"))
        _ref_ra.crop_and_resize_forward.restype = None
        _ref_ra.crop_and_resize_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]
        _ref_ra.crop_and_resize_backward.restype = None
        _ref_ra.crop_and_resize_backward.argtypes = [C.c_void_p] * 4
    return _ref_ra


def ref_crop_and_resize_fwd(image, boxes, box_ind, ph, pw, extrapolation=0.0):
    """lib/roi_align/src/crop_and_resize.c:115-154 (crop_and_resize_forward), unmodified, OpenMP over boxes."""
    image, boxes, box_ind = _c(image, np.float32), _c(boxes, np.float32).reshape(-1, 4), _c(box_ind, np.int32)
    out = np.empty((boxes.shape[0], image.shape[1], ph, pw), np.float32)
    ti, tb, tx, to = _th(image), _th(boxes), _th(box_ind), _th(out)
    _ref_roi_align().crop_and_resize_forward(C.byref(ti), C.byref(tb), C.byref(tx), extrapolation, ph, pw, C.byref(to))
    return out


def ref_crop_and_resize_bwd(grads, boxes, box_ind, im_size):
    """lib/roi_align/src/crop_and_resize.c:157-252 (crop_and_resize_backward), unmodified, serial."""
    grads, boxes, box_ind = _c(grads, np.float32), _c(boxes, np.float32).reshape(-1, 4), _c(box_ind, np.int32)
    out = np.empty(tuple(im_size), np.float32)
    tg, tb, tx, to = _th(grads), _th(boxes), _th(box_ind), _th(out)
    _ref_roi_align().crop_and_resize_backward(C.byref(tg), C.byref(tb), C.byref(tx), C.byref(to))
    return out


def ref_cpu_nms(dets_yxyxs, thresh):
    """lib/nms/pth_nms.py:7-20 (CPU branch) around lib/nms/src/nms.c:4-69 (cpu_nms), unmodified.

    dets[n,5] = (y1,x1,y2,x2,score) as the reference's callers pass them.  The reference hands
    cpu_nms the un-reordered ``dets`` (pth_nms.py:19), whose first four columns cpu_nms reads as
    (x1,y1,x2,y2) -- IoU is symmetric under that swap, so the result is unaffected.
    """
    global _ref_nms
    if _ref_nms is None:
        _ref_nms = C.CDLL(os.path.join(_REF_DIR, "libref_nms.so"))
        _ref_nms.cpu_nms.restype = C.c_int
        _ref_nms.cpu_nms.argtypes = [C.c_void_p] * 5 + [C.c_float]
    d = _c(dets_yxyxs, np.float32).reshape(-1, 5)
    n = d.shape[0]
    areas = ((d[:, 3] - d[:, 1] + 1) * (d[:, 2] - d[:, 0] + 1)).astype(np.float32)
    order = np.argsort(-d[:, 4], kind="stable").astype(np.int64)
    keep = np.zeros(max(n, 1), np.int64)
    num = np.zeros(1, np.int64)
    tk, tn, td, to, ta = _th(keep), _th(num), _th(d), _th(order), _th(areas)
    _ref_nms.cpu_nms(C.byref(tk), C.byref(tn), C.byref(td), C.byref(to), C.byref(ta), thresh)
    return keep[: int(num[0])].copy()


def ref_cuda():
    """The reference's own CUDA kernels (lib/roi_align/src/cuda/crop_and_resize_kernel.cu, lib/roi_pooling/src/roi_pooling_kernel.cu,
    lib/nms/src/cuda/nms_kernel.cu) compiled unmodified for sm_100a (oracle/_ref/libref_cuda.so): a second oracle on the GPU box.
    None when the library did not travel."""
    p = os.path.join(_REF_DIR, "libref_cuda.so")
    if not os.path.exists(p):
        return None
    L = C.CDLL(p)
    P, I, F = C.c_void_p, C.c_int, C.c_float
    L.CropAndResizeLaucher.argtypes = [P, P, P, I, I, I, I, I, I, I, F, P, P]
    L.CropAndResizeLaucher.restype = None
    L.CropAndResizeBackpropImageLaucher.argtypes = [P, P, P, I, I, I, I, I, I, I, P, P]
    L.CropAndResizeBackpropImageLaucher.restype = None
    L.ROIPoolForwardLaucher.argtypes = [P, F, I, I, I, I, I, I, P, P, P, P]          # roi_pooling_kernel.h:8-12
    L.ROIPoolForwardLaucher.restype = I
    L.ROIPoolBackwardLaucher.argtypes = [P, F, I, I, I, I, I, I, I, P, P, P, P]      # roi_pooling_kernel.h:14-18
    L.ROIPoolBackwardLaucher.restype = I
    return L
