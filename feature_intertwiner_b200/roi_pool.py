"""RoIPool operator layer -- lib/roi_pooling/functions/roi_pool.py:6-38 and modules/roi_pool.py."""
import torch
from torch import nn

from . import _lib


class _RoIPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, rois, ph, pw, scale):
        _lib.require_cuda(features, rois)
        # channels_last maps with C % 128 == 0 stay in their layout (warp-per-bin kernel, coalesced 128-bit loads); anything else is
        # pooled in NCHW like the reference.  The flat argmax indices are NCHW offsets either way (roi_pooling_kernel.cu:78-91).
        nhwc = features.dim() == 4 and features.size(1) % 128 == 0 and not features.is_contiguous() \
            and features.is_contiguous(memory_format=torch.channels_last) and features.dtype == torch.float32
        fmt = torch.channels_last if nhwc else torch.contiguous_format
        features = features.contiguous(memory_format=fmt)
        rois = rois.detach().float().contiguous()
        if rois.dim() != 2 or rois.size(1) != 5:
            raise _lib.FiError("rois must be [R,5] = (batch_index, x1, y1, x2, y2)")   # roi_pooling_cuda.c:20-23
        B, Cc, H, W = features.shape
        R = rois.size(0)
        out = torch.empty((R, Cc, ph, pw), device=features.device, dtype=torch.float32, memory_format=fmt)
        argmax = torch.empty((R, Cc, ph, pw), device=features.device, dtype=torch.int32, memory_format=fmt)
        L = _lib.lib()
        fn = L.fi_roi_pool_forward_nhwc if nhwc else L.fi_roi_pool_forward
        with torch.cuda.device(features.device):
            _lib.check(fn(_lib.ptr(features), float(scale), B, R, H, W, Cc, ph, pw, _lib.ptr(rois), _lib.ptr(out), _lib.ptr(argmax),
                          _lib.stream_ptr(features.device)))
        ctx.save_for_backward(rois, argmax)
        ctx.meta = (B, Cc, H, W, ph, pw, float(scale), nhwc)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        rois, argmax = ctx.saved_tensors
        B, Cc, H, W, ph, pw, scale, nhwc = ctx.meta
        fmt = torch.channels_last if nhwc else torch.contiguous_format
        grad_out = grad_out.contiguous(memory_format=fmt)
        grad_in = torch.empty((B, Cc, H, W), device=grad_out.device, dtype=torch.float32, memory_format=fmt)
        L = _lib.lib()
        fn = L.fi_roi_pool_backward_nhwc if nhwc else L.fi_roi_pool_backward
        with torch.cuda.device(grad_out.device):
            _lib.check(fn(_lib.ptr(grad_out), scale, B, rois.size(0), H, W, Cc, ph, pw, _lib.ptr(rois), _lib.ptr(grad_in), _lib.ptr(argmax),
                          _lib.stream_ptr(grad_out.device)))
        return grad_in, None, None, None, None


class RoIPoolFunction(object):
    """``RoIPoolFunction(ph, pw, spatial_scale)(features, rois)`` -- the reference's old-style call shape."""

    def __init__(self, pooled_height, pooled_width, spatial_scale):
        self.pooled_height, self.pooled_width, self.spatial_scale = int(pooled_height), int(pooled_width), float(spatial_scale)

    def __call__(self, features, rois):
        return _RoIPool.apply(features, rois, self.pooled_height, self.pooled_width, self.spatial_scale)


class _RoIPooling(nn.Module):
    """lib/roi_pooling/modules/roi_pool.py:5."""

    def __init__(self, pooled_height, pooled_width, spatial_scale):
        super().__init__()
        self.fn = RoIPoolFunction(pooled_height, pooled_width, spatial_scale)

    def forward(self, features, rois):
        return self.fn(features, rois)
