"""TEST INFRASTRUCTURE ONLY -- modern-torch restatement of the reference's Python layer of the hot path.

The reference's ``lib/sub_module.py::Dev``, ``lib/model.py::meta_loss``, ``lib/layers.py::pyramid_roi_align``
and the NMS wrappers cannot be imported on torch 2.x (``torch.utils.ffi`` is gone,
instance-style ``autograd.Function`` is rejected, uint8-mask arithmetic changed meaning ...).  This file
restates them with CPU tensors, calling the C oracle (``oracle/clib.py``) for every RoIAlign/NMS, so tests
can compare the CUDA product path against it.  ``lib/OT_module.py`` *is* importable; ``OptTransRef`` below is
checked against it output-for-output in tests/golden/make_golden.py (the frozen vectors live in
tests/golden/).  Parity status of everything else here: **unpinned by the reference** (it has no tests) --
pinned only transitively through the compiled-reference RoIAlign/NMS it calls.

Never imported by the product package.  All citations relative to /root/reference.
"""
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import clib

EPS = 1e-20  # lib/OT_module.py:4, lib/model.py (EPS)


# =============================================================================== config
def make_config(**over):
    """The subset of lib/config.py the path reads, with the reference's defaults (config.py:63,103-104,
    139-146,195-236) except STRUCTURE='beta' / SWITCH=True (the only runnable intertwiner, Appendix B.3)."""
    ns = types.SimpleNamespace
    cfg = ns(
        DEV=ns(SWITCH=True, STRUCTURE="beta", BASELINE=False, BUFFER_SIZE=1, LOSS_CHOICE="l2", OT_ONE_DIM_FORM="conv",
               LOSS_FAC=0.5, INST_LOSS=False, FEAT_BRANCH_POOL_SIZE=14, ASSIGN_BOX_ON_ALL_SCALE=False,
               BIG_FEAT_DETACH=True, UPSAMPLE_FAC=1.0, MULTI_UPSAMPLER=False, BIG_SUPERVISE=False, DIS_UPSAMPLER=False,
               INIT_BUFFER_WEIGHT="scratch"),
        ROIS=ns(METHOD="roi_align", ASSIGN_ANCHOR_BASE=224.0, TRAIN_ROIS_PER_IMAGE=200, ROI_POSITIVE_RATIO=0.33),
        MRCNN=ns(POOL_SIZE=7, MASK_POOL_SIZE=14),
        DATA=ns(IMAGE_SHAPE=np.array([1024, 1024, 3]), BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2])),
        DATASET=ns(NUM_CLASSES=81),
        RPN=ns(PRE_NMS_LIMIT=6000, NMS_THRESHOLD=0.7, POST_NMS_ROIS_TRAINING=2000, POST_NMS_ROIS_INFERENCE=1000),
    )
    for k, v in over.items():
        sec, key = k.split("__")
        setattr(getattr(cfg, sec), key, v)
    return cfg


# =============================================================================== RoIAlign on CPU
class _CropOracleFn(torch.autograd.Function):
    """crop_and_resize through the C oracle with the reference's autograd contract
    (lib/roi_align/crop_and_resize.py:21-54: grad only w.r.t. image)."""

    @staticmethod
    def forward(ctx, image, boxes, box_ind, ph, pw, extrap):
        out = clib.oracle_crop_and_resize_fwd(image.detach().numpy(), boxes.detach().numpy(), box_ind.numpy(), ph, pw, extrap)
        ctx.save_for_backward(boxes.detach(), box_ind)
        ctx.im_size = tuple(image.shape)
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, g):
        boxes, box_ind = ctx.saved_tensors
        gi = clib.oracle_crop_and_resize_bwd(g.contiguous().numpy(), boxes.numpy(), box_ind.numpy(), ctx.im_size)
        return torch.from_numpy(gi), None, None, None, None, None


def crop_and_resize_ref(image, boxes, box_ind, ph, pw, extrap=0.0):
    return _CropOracleFn.apply(image.contiguous(), boxes.contiguous(), box_ind.contiguous().int(), ph, pw, extrap)


# =============================================================================== level rule / split
def roi_level_ref(rois, image_shape, base=224.0):
    """lib/sub_module.py:397-410 (== lib/layers.py:168-181); log2 = tools/utils.py:50-55.  rois[bs,R,4]."""
    y1, x1, y2, x2 = rois.chunk(4, dim=2)
    h, w = y2 - y1, x2 - x1
    area = w * h
    image_area = torch.tensor([float(image_shape[0] * image_shape[1])], dtype=torch.float32)
    lvl = 4 + torch.log(torch.sqrt(area) / (base / torch.sqrt(image_area))) / torch.log(torch.tensor([2.0]))
    pre = lvl.squeeze(-1)
    # .round().int() of -inf is implementation-defined on CPU; clamp first in float (same result for finite values)
    lvl = lvl.round().clamp(2, 5).int().squeeze(-1)
    return lvl, pre


def big_mask_ref(level, roi_level):
    """lib/sub_module.py:367-378 (_find_big_box2): boxes assigned to coarser levels."""
    return roi_level > level if level < 5 else torch.zeros_like(roi_level, dtype=torch.bool)


def assign_feat2cls_ref(gt, feat, ncls):
    """lib/sub_module.py:664-684.  gt[k] ints, feat[k,F(,1,1)] -> feat[F,ncls], cnt[1,ncls]."""
    feat = feat.flatten(1)
    out = torch.zeros(feat.shape[1], ncls)
    cnt = torch.zeros(1, ncls)
    cols = []
    for c in torch.unique(gt).tolist():          # unique1d: sorted distinct values (tools/utils.py:30-41)
        if c == 0:
            continue
        idx = torch.nonzero(gt == c).squeeze(1)
        cnt[0, int(c)] = idx.numel()
        cols.append((int(c), feat[idx].mean(dim=0)))
    if cols:
        # functional scatter so autograd flows to `feat` as the reference's indexed assignment does
        out = _scatter_cols(out, cols)
    return out, cnt


def _scatter_cols(out, cols):
    idx = torch.tensor([c for c, _ in cols])
    vals = torch.stack([v for _, v in cols], dim=1)  # [F, m]
    return out.index_copy(1, idx, vals)


class DevRef(nn.Module):
    """lib/sub_module.py:286-692, structure 'beta', roi_align, ASSIGN_BOX_ON_ALL_SCALE=False."""

    def __init__(self, config, depth=256, feat_dim=1024):
        super().__init__()
        self.config, self.depth, self.feat_dim = config, depth, feat_dim
        self.pool_size, self.mask_pool_size = config.MRCNN.POOL_SIZE, config.MRCNN.MASK_POOL_SIZE
        self.feat_pool_size = config.DEV.FEAT_BRANCH_POOL_SIZE
        self.num_classes = config.DATASET.NUM_CLASSES
        self.image_shape = config.DATA.IMAGE_SHAPE
        n_up = 4 if config.DEV.MULTI_UPSAMPLER else 1
        conv = nn.Conv2d(depth, depth, 3, padding=1) if config.DEV.UPSAMPLE_FAC == 1.0 else \
            nn.ConvTranspose2d(depth, depth, 3, stride=2, padding=1, output_padding=1)
        # the reference shares ONE conv object across the upsamplers (sub_module.py:310-321, Appendix B.9)
        self.upsample = nn.ModuleList([nn.Sequential(conv, nn.BatchNorm2d(depth), nn.ReLU(inplace=True)) for _ in range(n_up)])
        k = self.feat_pool_size // 2
        self.feat_extract = nn.Sequential(
            nn.Conv2d(depth, feat_dim // 2, 3, padding=1, stride=2), nn.BatchNorm2d(feat_dim // 2), nn.ReLU(inplace=True),
            nn.Conv2d(feat_dim // 2, feat_dim, k), nn.BatchNorm2d(feat_dim), nn.ReLU(inplace=True),
            nn.Conv2d(feat_dim, feat_dim, 1), nn.BatchNorm2d(feat_dim), nn.ReLU(inplace=True))
        lc = config.DEV.LOSS_CHOICE
        self.last_op = nn.Sigmoid() if lc in ("l1", "l2") else (nn.Softmax(dim=1) if lc == "kl" else None)

    def forward(self, x, rois, roi_cls_gt=None):
        cfg = self.config
        train = roi_cls_gt is not None
        bs, R = rois.shape[:2]
        roi_level, _ = roi_level_ref(rois, self.image_shape, cfg.ROIS.ASSIGN_ANCHOR_BASE)
        pooled, mask, box_to_level = [], [], []
        big_feat, big_cnt, small_feat, small_cnt, big_loss = [], [], [], [], []
        small_output_all = torch.zeros(bs * R, self.feat_dim)
        small_gt_all = torch.zeros(bs * R)
        filled = 0
        zf = lambda: torch.zeros(self.feat_dim, self.num_classes)
        zc = lambda: torch.zeros(1, self.num_classes)
        for i, level in enumerate(range(2, 6)):
            fmap = x[i]
            use_meta = level in (2, 3, 4)
            small_ix = roi_level == level
            if not small_ix.any():                                            # sub_module.py:456-467
                if use_meta and train:
                    small_feat.append(zf()); small_cnt.append(zc()); big_feat.append(zf()); big_cnt.append(zc())
                    big_loss.append(torch.zeros(1))
                continue
            if train:                                                         # :472-536
                big_ix = big_mask_ref(level, roi_level)
                if not big_ix.any():
                    if use_meta:
                        big_feat.append(zf()); big_cnt.append(zc()); big_loss.append(torch.zeros(1))
                else:
                    bidx = torch.nonzero(big_ix)
                    bboxes = rois[bidx[:, 0], bidx[:, 1], :]
                    bgt = roi_cls_gt[bidx[:, 0], bidx[:, 1]]
                    bp = crop_and_resize_ref(fmap, bboxes, bidx[:, 0].int(), self.feat_pool_size, self.feat_pool_size)
                    bo = self.feat_extract(bp)
                    if self.last_op is not None:
                        bo = self.last_op(bo)
                    f, c = assign_feat2cls_ref(bgt, bo, self.num_classes)
                    big_feat.append(f); big_cnt.append(c); big_loss.append(torch.zeros(1))
            sidx = torch.nonzero(small_ix)                                    # :539-600
            box_to_level.append(sidx)
            sboxes = rois[sidx[:, 0], sidx[:, 1], :]
            sind = sidx[:, 0].int()
            fm = self.upsample[i if cfg.DEV.MULTI_UPSAMPLER else 0](fmap)
            pooled.append(crop_and_resize_ref(fm, sboxes, sind, self.pool_size, self.pool_size))
            mf = crop_and_resize_ref(fm, sboxes, sind, self.mask_pool_size, self.mask_pool_size)
            mask.append(mf)
            if use_meta:
                so = self.feat_extract(mf)
                if self.last_op is not None:
                    so = self.last_op(so)
                n = sidx.shape[0]
                small_output_all = small_output_all.clone()
                small_output_all[filled:filled + n] = so.reshape(n, -1)
                if train:
                    sgt = roi_cls_gt[sidx[:, 0], sidx[:, 1]]
                    f, c = assign_feat2cls_ref(sgt, so, self.num_classes)
                    small_feat.append(f); small_cnt.append(c)
                    small_gt_all[filled:filled + n] = sgt.float()
                else:
                    small_gt_all[filled:filled + n] = 1
                filled += n
        pooled_out, mask_out = reshape_result_ref(pooled, mask, box_to_level, (bs, R))
        if train:
            bf = torch.stack(big_feat).unsqueeze(0)
            if cfg.DEV.BIG_FEAT_DETACH:
                bf = bf.detach()
            feat_out = [bf, torch.stack(big_cnt).unsqueeze(0), torch.stack(small_feat).unsqueeze(0),
                        torch.stack(small_cnt).unsqueeze(0), torch.stack(big_loss).unsqueeze(0),
                        small_output_all, small_gt_all]
        else:
            feat_out = [small_output_all, small_gt_all]
        return pooled_out, mask_out, feat_out


def reshape_result_ref(pooled, mask, box_to_level, rois_size):
    """lib/sub_module.py:645-662: scatter per-level rows back to (image, roi) order."""
    pooled, mask, b2l = torch.cat(pooled, 0), torch.cat(mask, 0), torch.cat(box_to_level, 0)
    outs = []
    for t in (pooled, mask):
        o = torch.zeros(rois_size[0], rois_size[1], *t.shape[1:])
        o = o.index_put((b2l[:, 0], b2l[:, 1]), t)
        outs.append(o.view(-1, *t.shape[1:]))
    return outs[0], outs[1]


def pyramid_roi_align_ref(boxes, feature_maps, pool_size, image_shape, base=224.0):
    """lib/layers.py:145-218 (intertwiner disabled)."""
    roi_level, _ = roi_level_ref(boxes, image_shape, base)
    pooled, b2l = [], []
    for i, level in enumerate(range(2, 6)):
        ix = roi_level == level
        if not ix.any():
            continue
        index = torch.nonzero(ix)
        b2l.append(index)
        pooled.append(crop_and_resize_ref(feature_maps[i], boxes[index[:, 0], index[:, 1], :], index[:, 0].int(), pool_size, pool_size))
    pooled, b2l = torch.cat(pooled, 0), torch.cat(b2l, 0)
    out = torch.zeros(boxes.shape[0], boxes.shape[1], *pooled.shape[1:])
    out = out.index_put((b2l[:, 0], b2l[:, 1]), pooled)
    return out.view(-1, *pooled.shape[1:])


# =============================================================================== meta loss
def merge_feat_vec_ref(feat, cnt):
    """lib/model.py:217-224.  feat[G,S,F,ncls], cnt[G,S,1,ncls]."""
    s = (feat * cnt).sum(0).sum(0)
    n = cnt.sum(0).sum(0)
    return s / (n + EPS), n


class MetaLossRef:
    """lib/model.py:106-111 (buffer) + :143-215 (meta_loss), l1/l2/kl/ot, class- and instance-level."""

    def __init__(self, config, feat_dim=1024, ot_loss=None):
        self.config = config
        B, ncls = config.DEV.BUFFER_SIZE, config.DATASET.NUM_CLASSES
        self.buffer = torch.zeros(B, feat_dim, ncls)
        self.buffer_cnt = torch.zeros(B, 1, ncls)
        self.device = torch.device('cpu')
        self.ot_loss = ot_loss

    def __call__(self, feat_input):
        big_feat, big_cnt, small_feat, small_cnt, small_output_all, small_gt_all = feat_input
        bf, bc = merge_feat_vec_ref(big_feat.detach(), big_cnt.detach())
        if self.buffer.shape[0] == 1:                                            # model.py:153-158
            fsum = self.buffer * self.buffer_cnt + bf.unsqueeze(0) * bc.unsqueeze(0)
            self.buffer_cnt = self.buffer_cnt + bc.unsqueeze(0)
            self.buffer = fsum / (self.buffer_cnt + EPS)
            final_big = self.buffer[0]
        else:                                                                    # :159-166
            self.buffer = torch.cat([self.buffer[1:], bf.unsqueeze(0)], 0)
            self.buffer_cnt = torch.cat([self.buffer_cnt[1:], bc.unsqueeze(0)], 0)
            final_big = (self.buffer * self.buffer_cnt).sum(0) / (self.buffer_cnt.sum(0) + EPS)
        in_buffer = self.buffer_cnt.sum(0).squeeze(0) > 0
        if self.config.DEV.INST_LOSS:                                            # :168-174
            gt = small_gt_all.long()
            idx = torch.nonzero((gt != 0) & in_buffer[gt]).squeeze(1)
        else:                                                                    # :175-181
            fs, fc = merge_feat_vec_ref(small_feat, small_cnt)
            fc = fc.clone(); fc[0, 0] = 0
            idx = torch.nonzero((fc.squeeze(0) > 0) & in_buffer).squeeze(1)
        self.last_idx = idx
        if idx.numel() == 0:
            return torch.zeros(1, device=self.buffer.device)
        if self.config.DEV.INST_LOSS:
            SMALL = small_output_all[idx]
            BIG = final_big[:, small_gt_all[idx].long()].t()
        else:
            SMALL = fs[:, idx].t()
            BIG = final_big[:, idx].t()
        lc = self.config.DEV.LOSS_CHOICE
        if lc == "l2":
            return F.mse_loss(SMALL, BIG)
        if lc == "l1":
            return F.l1_loss(SMALL, BIG)
        if lc == "kl":
            # reference calls F.kl_div(log SMALL, BIG) with torch-0.3 defaults (size_average=True) == 'mean'
            return F.kl_div(torch.log(SMALL), BIG, reduction="mean")
        if lc == "ot":
            return self.ot_loss(SMALL.unsqueeze(-1), BIG.unsqueeze(-1).contiguous())
        raise ValueError(lc)


# =============================================================================== OptTrans
def sinkhorn_iterate_ref(x, y, inv_eps=1.0, L=5, detach_plan=True):
    """lib/OT_module.py:104-135, cosine cost, normalisation done OUT of place (the reference's in-place
    `x /= ...` breaks modern autograd); forward values are identical."""
    n = x.size(0)
    x = x / (torch.norm(x, p=2, dim=1, keepdim=True) + EPS)
    y = y / (torch.norm(y, p=2, dim=1, keepdim=True) + EPS)
    C = 1 - torch.mm(x, y.permute(1, 0))
    K = torch.exp(-inv_eps * C)
    b = torch.ones(n, 1, dtype=x.dtype, device=x.device) * (1.0 / n)
    const = torch.ones(n, 1, dtype=x.dtype, device=x.device) * (1.0 / n)
    a = const
    for _ in range(L):
        a = const / (torch.mm(K, b) + EPS)
        b = const / (torch.mm(K.permute(1, 0), a) + EPS)
    P = a * K * b.permute(1, 0)
    if detach_plan:
        P = P.detach()
    return torch.dot(P.reshape(-1), C.reshape(-1))


class OptTransRef(nn.Module):
    """lib/OT_module.py:7-102 with identical parameter names (G_net.*, critic.*) so a reference
    state_dict loads unchanged."""

    def __init__(self, config=None, ch_x=1024, spatial_x=-1, ch_y=-1, spatial_y=-1, epsilon=1.0, L=5, remove_bias=False,
                 no_bp_P_L=True):
        super().__init__()
        self.inv_eps, self.L, self.remove_bias, self.no_bp_P_L = 1.0 / epsilon, L, remove_bias, no_bp_P_L
        two_dim = spatial_x > 1
        ch_y = ch_x if ch_y == -1 else ch_y
        spatial_y = spatial_x if spatial_y == -1 else spatial_y
        if two_dim:
            stride, out_pad = (2, 1) if spatial_x != spatial_y else (1, 0)
            self.G_net = nn.Sequential(nn.ConvTranspose2d(ch_x, ch_y, 3, padding=1, stride=stride, output_padding=out_pad),
                                       nn.BatchNorm2d(ch_y), nn.ReLU())
            self.critic = nn.Sequential(nn.Conv2d(ch_y, ch_y // 2, 3, padding=1, stride=2), nn.BatchNorm2d(ch_y // 2), nn.ReLU(),
                                        nn.Conv2d(ch_y // 2, ch_y // 4, 3, padding=1, stride=2), nn.BatchNorm2d(ch_y // 4), nn.ReLU())
        else:
            self.G_net = nn.Sequential(nn.Conv1d(ch_x, ch_y, 3, padding=1), nn.ReLU())
            self.critic = nn.Sequential(nn.Conv1d(ch_y, ch_y // 4, 3, padding=1), nn.ReLU())

    def _w(self, x, y):
        bs = x.size(0)
        cx = self.critic(x); cx = cx.view(bs, cx.size(1), -1)
        cy = self.critic(y); cy = cy.view(bs, cy.size(1), -1)
        return torch.stack([sinkhorn_iterate_ref(cx[i], cy[i], self.inv_eps, self.L, self.no_bp_P_L) for i in range(bs)])

    def forward(self, x, y):
        xu = self.G_net(x)
        if self.remove_bias:
            return self._w(xu, y)
        return 2 * self._w(xu, y) - self._w(xu, xu) - self._w(y, y)


# =============================================================================== NMS wrappers
def pth_nms_ref(dets, thresh, strict=True):
    """lib/nms/pth_nms.py:5-46.  dets[N,5]=(y1,x1,y2,x2,score).  strict=True reproduces the GPU branch
    (IoU > thr, nms_kernel.cu:63; the mask kernel sees boxes in INPUT order, Appendix B.6),
    strict=False the CPU branch (IoU >= thr, nms.c:59; boxes visited in score order)."""
    d = dets.detach().cpu().numpy().astype(np.float32)
    # the reference's sort is unstable (pth_nms.py:37): with tied scores its result is implementation-defined;
    # a stable sort is used here and in the product so the two agree
    order = torch.sort(dets[:, 4], dim=0, descending=True, stable=True)[1].cpu().numpy()
    xyxy = d[:, [1, 0, 3, 2, 4]]
    if strict:
        keep = clib.oracle_nms(xyxy, thresh, True)          # un-reordered dets_temp (pth_nms.py:28-44)
    else:
        keep = clib.oracle_nms(xyxy[order], thresh, False)
    return torch.from_numpy(order[keep].astype(np.int64))


def nms_ref(dets, thresh, strict=True):
    """lib/nms/nms_wrapper.py:14-34: per image, truncated to the minimum keep count, numpy int32."""
    keeps = [pth_nms_ref(dets[i], thresh, strict) for i in range(dets.size(0))]
    m = min(len(k) for k in keeps)
    out = np.zeros((dets.size(0), m), dtype=np.int32)
    for i, k in enumerate(keeps):
        out[i] = k[:m].numpy()
    return out


# =============================================================================== proposal layer
def apply_box_deltas_ref(boxes, deltas):
    """Box refinement of tools/box_utils.py:7-29, restated: corners -> (centre, size), centre shifted by delta * size, size
    scaled by exp(delta), back to corners.  The order of the fp32 operations is the reference's (one rounding per torch op)."""
    y1, x1, y2, x2 = boxes.unbind(dim=2)
    dy, dx, dh, dw = deltas.unbind(dim=2)
    h, w = y2 - y1, x2 - x1                                  # :14-15
    cy = (y1 + 0.5 * h) + dy * h                             # :16,19
    cx = (x1 + 0.5 * w) + dx * w                             # :17,20
    h, w = h * torch.exp(dh), w * torch.exp(dw)              # :21-22
    top, left = cy - 0.5 * h, cx - 0.5 * w                   # :24-25
    return torch.stack([top, left, top + h, left + w], dim=2)   # :26-28


def proposal_layer_ref(inputs, proposal_count, nms_threshold, priors, config):
    """lib/layers.py:71-139 on CPU tensors: top PRE_NMS_LIMIT anchors by foreground score, deltas * BBOX_STD_DEV applied,
    clipped to the image, NMS per image (GPU rule, via the C oracle), truncated to the smallest keep count of the batch
    (lib/nms/nms_wrapper.py:24-33) and to proposal_count, normalised.  A stable sort stands in for the reference's
    unstable one (ties are implementation-defined there).  Returns (normalised boxes [bs, m, 4], keep [bs, m])."""
    scores = inputs[0][:, :, 1].detach().float().cpu()
    deltas = inputs[1].detach().float().cpu() * torch.from_numpy(np.reshape(config.DATA.BBOX_STD_DEV, [1, 1, 4])).float()
    anchors = priors.detach().float().cpu()
    bs, prior_num = scores.size(0), anchors.size(0)
    pre = min(config.RPN.PRE_NMS_LIMIT, prior_num)
    scores, order = scores.sort(dim=1, descending=True, stable=True)
    scores, order = scores[:, :pre], order[:, :pre]
    deltas_trim = torch.stack([deltas[i][order[i]] for i in range(bs)])
    anchors_trim = torch.stack([anchors[order[i]] for i in range(bs)])
    boxes = apply_box_deltas_ref(anchors_trim, deltas_trim)
    height, width = float(config.DATA.IMAGE_SHAPE[0]), float(config.DATA.IMAGE_SHAPE[1])
    boxes = torch.stack([boxes[:, :, 0].clamp(0.0, height), boxes[:, :, 1].clamp(0.0, width),
                         boxes[:, :, 2].clamp(0.0, height), boxes[:, :, 3].clamp(0.0, width)], 2)
    keep = nms_ref(torch.cat((boxes, scores.unsqueeze(2)), 2), nms_threshold, strict=True)      # numpy int32 [bs, min_keep]
    keep = torch.from_numpy(keep[:, :proposal_count].astype(np.int64))
    boxes_keep = torch.stack([boxes[i][keep[i]] for i in range(bs)])
    norm = torch.tensor([height, width, height, width])
    return boxes_keep / norm, keep


# =============================================================================== mask targets / detection layer
def mask_targets_ref(pos_rois, gt_boxes, assignment, gt_masks, mask_shape, use_mini_mask=True):
    """lib/layers.py:296-323: RoI rewritten into its GT box's frame, C = 1 crop of the assigned GT mask (C oracle), rounded."""
    a = assignment.long()
    roi_gt = gt_boxes[a]
    boxes = pos_rois
    if use_mini_mask:                                                                    # :304-313
        y1, x1, y2, x2 = pos_rois.chunk(4, dim=1)
        gy1, gx1, gy2, gx2 = roi_gt.chunk(4, dim=1)
        gh, gw = gy2 - gy1, gx2 - gx1
        boxes = torch.cat([(y1 - gy1) / gh, (x1 - gx1) / gw, (y2 - gy1) / gh, (x2 - gx1) / gw], dim=1)
    crops = clib.oracle_crop_and_resize_fwd(gt_masks[a].unsqueeze(1).contiguous().numpy(), boxes.contiguous().numpy(),
                                            np.arange(a.numel(), dtype=np.int32), int(mask_shape[0]), int(mask_shape[1]), 0.0)
    return torch.round(torch.from_numpy(crops).squeeze(1))                               # :323


def detection_layer_ref(rois, probs, deltas, windows, config):
    """lib/layers.py:720-802 + conduct_nms (:664-717): arg-max class, class-specific refinement, clip to the window, round, then per
    image and per class a greedy NMS in score order (C oracle, CPU rule IoU >= thr) and the DET_MAX_INSTANCES best survivors."""
    bs, R = rois.size(0), rois.size(1)
    K = int(config.TEST.DET_MAX_INSTANCES)
    scores, cls = probs.max(dim=1)
    d = deltas[torch.arange(bs * R), cls] * torch.from_numpy(np.reshape(config.DATA.BBOX_STD_DEV, [1, 4])).float()
    refined = apply_box_deltas_ref(rois.reshape(1, -1, 4), d.unsqueeze(0))[0]
    H, W = float(config.DATA.IMAGE_SHAPE[0]), float(config.DATA.IMAGE_SHAPE[1])
    refined = refined * torch.tensor([H, W, H, W])
    win = windows.repeat_interleave(R, dim=0)
    refined = torch.stack([torch.minimum(torch.maximum(refined[:, 0], win[:, 0]), win[:, 2]), torch.minimum(torch.maximum(refined[:, 1], win[:, 1]), win[:, 3]),
                           torch.minimum(torch.maximum(refined[:, 2], win[:, 0]), win[:, 2]), torch.minimum(torch.maximum(refined[:, 3], win[:, 1]), win[:, 3])], 1)
    refined = torch.round(refined)
    area = (refined[:, 0] - refined[:, 2]) * (refined[:, 1] - refined[:, 3])
    keep = (cls > 0) & (scores >= float(config.TEST.DET_MIN_CONFIDENCE)) & (area > 0)
    out = torch.zeros(bs, K, 6)
    for b in range(bs):
        sl = slice(b * R, (b + 1) * R)
        idx = torch.nonzero(keep[sl]).squeeze(1)
        if idx.numel() == 0:
            continue
        survivors = []
        for c in torch.unique(cls[sl][idx]).tolist():
            ix = idx[cls[sl][idx] == c]
            sc, order = scores[sl][ix].sort(descending=True, stable=True)
            bx = refined[sl][ix][order]
            k = clib.oracle_nms(torch.cat([bx[:, [1, 0, 3, 2]], sc.unsqueeze(1)], 1).numpy(), float(config.TEST.DET_NMS_THRESHOLD), False)
            survivors.append(ix[order[torch.from_numpy(k.astype(np.int64))]])
        surv = torch.cat(survivors)
        top = surv[scores[sl][surv].sort(descending=True, stable=True)[1][:K]]
        out[b, : top.numel()] = torch.cat([refined[sl][top], cls[sl][top].unsqueeze(1).float(), scores[sl][top].unsqueeze(1)], 1)
    return out
