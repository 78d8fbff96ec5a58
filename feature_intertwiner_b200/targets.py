"""The two remaining consumers of the RoIAlign / NMS kernels in the detector.

* ``mask_targets`` -- the mask-target crops of ``generate_roi`` (lib/layers.py:296-323): C = 1 RoIAlign with one "image" per
  box, the USE_MINI_MASK coordinate change and the rounding, in one launch straight from ``gt_masks`` (no gathered copy).
* ``detection_layer`` -- lib/layers.py:720-802 with ``conduct_nms`` (:664-717): same arguments, same output
  ``[bs, DET_MAX_INSTANCES, 6] = (y1, x1, y2, x2, class_id, score)``, zero rows after an image's detections.

  The reference loops over images and, inside, over the classes present, launching a sort, an NMS with its own device->host
  round trip and several index ops per class.  Here: one decode launch (fi_detection_decode), one per-image sort, ONE batched
  NMS launch for every class of every image at once -- boxes of class c are shifted by c * (extent + 2) pixels along both axes,
  so boxes of different classes never overlap while the IoU of two boxes of one class is unchanged (the coordinates are
  integers after torch.round: the shifted values and their differences are exact in fp32) -- and a gather.  No host read.
"""
import ctypes as C

import torch

from . import _lib
from .nms import nms_presorted


def mask_targets(pos_rois, gt_boxes, assignment, gt_masks, mask_shape, use_mini_mask=True):
    """pos_rois[n,4] normalised (y1,x1,y2,x2), gt_boxes[G,4], assignment[n] (index of each RoI's GT), gt_masks[G,mh,mw]
    -> targets[n, mask_shape[0], mask_shape[1]] in {0, 1} (lib/layers.py:296-323)."""
    _lib.require_cuda(pos_rois, gt_boxes, assignment, gt_masks)
    dev = pos_rois.device
    pos_rois = pos_rois.detach().float().contiguous()
    gt_boxes = gt_boxes.detach().float().contiguous()
    assignment = assignment.detach().to(torch.int32).contiguous()
    gt_masks = gt_masks.detach().float().contiguous()
    n, G = pos_rois.size(0), gt_masks.size(0)
    if gt_masks.dim() != 3 or pos_rois.shape != (n, 4) or gt_boxes.shape != (G, 4) or assignment.shape != (n,):
        raise _lib.FiError("mask_targets: pos_rois [n,4], gt_boxes [G,4], assignment [n], gt_masks [G,mh,mw]")
    MH, MW = int(mask_shape[0]), int(mask_shape[1])
    out = torch.empty((n, MH, MW), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fi_mask_targets(_lib.ptr(gt_masks), _lib.ptr(pos_rois), _lib.ptr(gt_boxes), _lib.ptr(assignment), n, G, gt_masks.size(1),
                                              gt_masks.size(2), MH, MW, 1 if use_mini_mask else 0, _lib.ptr(out), _lib.stream_ptr(dev)))
    return out


def detection_decode(rois, probs, deltas, windows, config):
    """Front half of detection_layer (lib/layers.py:738-766) -> (boxes[bs*R,4] pixels, rounded; scores; class_ids; keep)."""
    _lib.require_cuda(rois, probs, deltas, windows)
    dev = rois.device
    bs, R = rois.size(0), rois.size(1)
    rois = rois.detach().float().contiguous()
    probs = probs.detach().float().contiguous()
    deltas = deltas.detach().float().contiguous()
    windows = windows.detach().float().contiguous()
    ncls = probs.size(1)
    if probs.shape != (bs * R, ncls) or deltas.shape != (bs * R, ncls, 4) or windows.shape != (bs, 4):
        raise _lib.FiError("detection_layer: rois [bs,R,4], probs [bs*R,ncls], deltas [bs*R,ncls,4], windows [bs,4]")
    boxes = torch.empty((bs * R, 4), device=dev, dtype=torch.float32)
    scores = torch.empty((bs * R,), device=dev, dtype=torch.float32)
    class_ids = torch.empty((bs * R,), device=dev, dtype=torch.int32)
    keep = torch.empty((bs * R,), device=dev, dtype=torch.int32)
    std = (C.c_float * 4)(*[float(v) for v in config.DATA.BBOX_STD_DEV])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fi_detection_decode(_lib.ptr(rois), _lib.ptr(probs), _lib.ptr(deltas), _lib.ptr(windows), bs, R, ncls,
                                                  C.cast(std, C.c_void_p), float(config.DATA.IMAGE_SHAPE[0]), float(config.DATA.IMAGE_SHAPE[1]),
                                                  float(config.TEST.DET_MIN_CONFIDENCE), _lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(class_ids),
                                                  _lib.ptr(keep), _lib.stream_ptr(dev)))
    return boxes, scores, class_ids, keep


def detection_layer(rois, probs, deltas, windows, config, feature=None):
    """lib/layers.py:720-802.  Returns ``detections [bs, DET_MAX_INSTANCES, 6]`` (and the gathered ``feature`` rows when given,
    like the reference).  Per image the rows are the class-wise NMS survivors in descending score order, zero rows after."""
    bs, R = rois.size(0), rois.size(1)
    dev = rois.device
    K = int(config.TEST.DET_MAX_INSTANCES)
    boxes, scores, class_ids, keep = detection_decode(rois, probs, deltas, windows, config)
    boxes, scores, class_ids, keep = boxes.view(bs, R, 4), scores.view(bs, R), class_ids.view(bs, R), keep.view(bs, R)
    # candidates first, by descending score (conduct_nms sorts every class by score: layers.py:690); the rest behind them
    key = torch.where(keep > 0, scores, torch.full_like(scores, -1.0))
    key, order = torch.sort(key, dim=1, descending=True, stable=True)
    n_cand = keep.sum(dim=1, keepdim=True)                                   # [bs,1] on the device
    sb = torch.gather(boxes, 1, order.unsqueeze(2).expand(bs, R, 4))
    sc = torch.gather(class_ids, 1, order)
    # class-wise NMS as ONE batched launch: shift class c by c * (extent + 2) pixels -- integers, exact in fp32
    extent = float(max(int(config.DATA.IMAGE_SHAPE[0]), int(config.DATA.IMAGE_SHAPE[1])) + 2)
    shift = (sc.float() * extent).unsqueeze(2)
    dets = torch.cat([sb[:, :, [1, 0, 3, 2]] + shift, key.unsqueeze(2)], dim=2).contiguous()     # (x1,y1,x2,y2,score): pth_nms.py:28-33
    kept, num = nms_presorted(dets, float(config.TEST.DET_NMS_THRESHOLD), max_keep=K)   # positions in score order, survivors first
    pos = kept[:, :K] if kept.size(1) >= K else torch.cat([kept, kept.new_full((bs, K - kept.size(1)), -1)], dim=1)
    valid = (pos >= 0) & (pos < n_cand)                                      # a survivor that is no candidate sits behind all candidates
    idx = pos.clamp(min=0).long()
    out_boxes = torch.gather(sb, 1, idx.unsqueeze(2).expand(bs, K, 4))
    out = torch.cat([out_boxes, torch.gather(sc, 1, idx).unsqueeze(2).float(), torch.gather(key, 1, idx).unsqueeze(2)], dim=2)
    detections = out * valid.unsqueeze(2).to(out.dtype)
    if feature is None:
        return detections
    src = torch.gather(order, 1, idx) + (torch.arange(bs, device=dev) * R).unsqueeze(1)          # row of each detection in `feature`
    feat = feature[src.view(-1)].view(bs, K, -1) * valid.unsqueeze(2).to(feature.dtype)
    return detections, feat
