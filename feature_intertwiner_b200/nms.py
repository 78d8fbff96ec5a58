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


def nms_presorted(dets_xyxys, thresh, max_keep=None):
    """dets[bs,N,5] = (x1,y1,x2,y2,score), every image already in descending score order -> (keep[bs,N] int32 positions,
    kept ones first in score order, rest -1; num_keep[bs] int32).  No sort, no host synchronisation.  ``max_keep``: the caller
    uses only that many survivors per image -- the sweep stops there (same head of the list; num_keep >= max_keep means cut)."""
    _lib.require_cuda(dets_xyxys)
    bs, n, _ = dets_xyxys.shape
    words = (n + 63) // 64
    mask = torch.empty((bs, n, max(words, 1)), device=dets_xyxys.device, dtype=torch.int64)
    keep = torch.empty((bs, n), device=dets_xyxys.device, dtype=torch.int32)
    num = torch.empty((bs,), device=dets_xyxys.device, dtype=torch.int32)
    limit = n if max_keep is None else max(1, min(int(max_keep), n))
    with torch.cuda.device(dets_xyxys.device):
        _lib.check(_lib.lib().fi_nms_batched_topk(_lib.ptr(dets_xyxys), bs, n, float(thresh), max(limit, 1), _lib.ptr(mask), _lib.ptr(keep),
                                                  _lib.ptr(num), _lib.stream_ptr(dets_xyxys.device)))
    return keep, num


def proposal_decode(inputs, priors, config):
    """Front half of proposal_layer (lib/layers.py:87-122): the PRE_NMS_LIMIT best anchors per image by foreground score,
    refined by ``rpn_bbox * BBOX_STD_DEV`` and clipped to the image.  Returns (boxes[bs,K,4] = (y1,x1,y2,x2) in pixels,
    dets[bs,K,5] = (x1,y1,x2,y2,score), both in descending score order).  One sort + one launch (fi_proposal_decode)."""
    import ctypes as C
    probs, deltas = inputs[0], inputs[1]
    _lib.require_cuda(probs, deltas)
    dev = probs.device
    scores = probs.detach()[:, :, 1].float()
    deltas = deltas.detach().float().contiguous()
    anchors = priors.detach().to(device=dev, dtype=torch.float32).contiguous()
    bs, A = scores.shape
    if anchors.shape != (A, 4) or deltas.shape != (bs, A, 4):
        raise _lib.FiError("proposal_layer: rpn_probs [bs,A,2], rpn_bbox [bs,A,4] and priors [A,4] do not agree")
    pre = min(int(config.RPN.PRE_NMS_LIMIT), A)
    scores_s, order = torch.sort(scores, dim=1, descending=True, stable=True)       # layers.py:103 (stable: ties are defined)
    scores_s, order = scores_s[:, :pre].contiguous(), order[:, :pre].contiguous()
    boxes = torch.empty((bs, pre, 4), device=dev, dtype=torch.float32)
    dets = torch.empty((bs, pre, 5), device=dev, dtype=torch.float32)
    std = (C.c_float * 4)(*[float(v) for v in config.DATA.BBOX_STD_DEV])
    height, width = float(config.DATA.IMAGE_SHAPE[0]), float(config.DATA.IMAGE_SHAPE[1])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fi_proposal_decode(_lib.ptr(deltas), _lib.ptr(anchors), _lib.ptr(order), _lib.ptr(scores_s), bs, A, pre,
                                                 C.cast(std, C.c_void_p), height, width, _lib.ptr(boxes), _lib.ptr(dets),
                                                 _lib.stream_ptr(dev)))
    return boxes, dets


def proposal_layer(inputs, proposal_count, nms_threshold, priors, config=None, static=False):
    """lib/layers.py:71-139 with the same call shape: ``inputs = [rpn_probs[bs,A,2], rpn_bbox[bs,A,4]]``, ``priors[A,4]`` anchors
    in pixels (y1,x1,y2,x2) -> normalised proposals ``[bs, m, 4]``, m = min(proposal_count, smallest keep count of the batch)
    like lib/nms/nms_wrapper.py:24-33.

    The reference sorts on the device, gathers per image in a Python loop, runs five elementwise ops, concatenates for NMS,
    copies the 4.5 MB suppression mask of every image to the host, reduces it there and indexes again per image.  Here:
    one sort (torch), one decode launch (fi_proposal_decode: gather + deltas + clip + NMS layout), the batched on-device NMS,
    one gather launch (fi_proposal_gather: truncation to the batch's smallest keep count + gather + normalisation).

    ``static=False`` returns the reference's ``[bs, m, 4]`` -- reading ``m`` is the one host synchronisation, the API demands a
    tensor of that size.  ``static=True`` returns ``(rois[bs, proposal_count, 4], m)`` with the rows past ``m`` zero (what the
    reference's later layers pad with, lib/layers.py:413,427) and ``m`` a device int32: no host read at all, fixed shapes."""
    boxes, dets = proposal_decode(inputs, priors, config)
    bs, dev = boxes.size(0), boxes.device
    height, width = float(config.DATA.IMAGE_SHAPE[0]), float(config.DATA.IMAGE_SHAPE[1])
    keep, num = nms_presorted(dets, nms_threshold, max_keep=int(proposal_count))
    rois = torch.empty((bs, int(proposal_count), 4), device=dev, dtype=torch.float32)
    m_dev = torch.empty((1,), device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fi_proposal_gather(_lib.ptr(boxes), _lib.ptr(keep), _lib.ptr(num), bs, boxes.size(1), int(proposal_count), height, width,
                                                 _lib.ptr(rois), _lib.ptr(m_dev), _lib.stream_ptr(dev)))
    if static:
        return rois, m_dev
    return rois[:, : int(m_dev.item())]                                              # the one host sync
