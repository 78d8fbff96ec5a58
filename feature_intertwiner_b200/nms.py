"""NMS operator layer -- lib/nms/nms_wrapper.py:14-34 (``nms``) and lib/nms/pth_nms.py:5-46 (``pth_nms``).

Everything up to the final ``.cpu().numpy()`` the reference's API demands stays on the device: the
suppression mask AND the greedy reduce run as kernels (csrc/nms.cu), batched over the images, so the
4.5 MB-per-image D2H copy and the host loop of lib/nms/src/nms_cuda.c:33-58 are gone.

Rule: IoU > thresh suppresses (the reference's GPU branch, nms_kernel.cu:63; its CPU branch uses >=, nms.c:59).
Unsorted input: the reference's GPU branch feeds the *unsorted* boxes to the kernel and maps the result through
the sort order (pth_nms.py:28-46), which is only meaningful for pre-sorted input -- as all
its callers provide (lib/layers.py:103,690).  Here boxes are sorted first, then suppressed, i.e. the intended
behaviour; on pre-sorted input the two coincide.
"""
import numpy as np
import torch

from . import _lib


def nms_batched(dets, thresh):
    """dets[bs,N,5] = (y1,x1,y2,x2,score) on the GPU -> (keep[bs,N] int64 indices into the original order, padded with
    -1, num_keep[bs] int32), no host synchronisation."""
    if dets.dim() != 3 or dets.size(2) != 5:
        raise _lib.FiError("dets must be [bs,N,5], got %s" % (tuple(dets.shape),))
    _lib.require_cuda(dets)
    bs, n, _ = dets.shape
    dets = dets.detach().float()
    order = torch.sort(dets[:, :, 4], dim=1, descending=True, stable=True)[1]            # pth_nms.py:37
    srt = torch.gather(dets, 1, order.unsqueeze(2).expand(bs, n, 5))
    xyxy = srt[:, :, [1, 0, 3, 2, 4]].contiguous()                                       # pth_nms.py:28-33
    words = (n + 63) // 64
    mask = torch.empty((bs, n, max(words, 1)), device=dets.device, dtype=torch.int64)
    keep = torch.empty((bs, n), device=dets.device, dtype=torch.int32)
    num = torch.empty((bs,), device=dets.device, dtype=torch.int32)
    with torch.cuda.device(dets.device):
        _lib.check(_lib.lib().fi_nms_batched(_lib.ptr(xyxy), bs, n, float(thresh), _lib.ptr(mask), _lib.ptr(keep),
                                             _lib.ptr(num), _lib.stream_ptr(dets.device)))
    valid = keep >= 0
    keep_orig = torch.gather(order, 1, keep.clamp(min=0).long())                          # pth_nms.py:46
    return torch.where(valid, keep_orig, torch.full_like(keep_orig, -1)), num


def pth_nms(dets, thresh):
    """dets[N,5] -> LongTensor of kept indices, descending score (lib/nms/pth_nms.py:5-46)."""
    keep, num = nms_batched(dets.unsqueeze(0), thresh)
    return keep[0, : int(num[0].item())].contiguous()


def nms(dets, thresh):
    """dets[bs,N,5] -> numpy int32 [bs, min_keep] (lib/nms/nms_wrapper.py:14-34)."""
    keep, num = nms_batched(dets, thresh)
    m = int(num.min().item())            # the one host sync: the reference's API returns a numpy array
    return keep[:, :m].to(torch.int32).cpu().numpy().astype(np.int32)
