"""The intertwiner proper: ``Dev`` (lib/sub_module.py:286-692, structure 'beta') and the meta loss with its
historical class buffer (lib/model.py:106-111,143-224), on the kernels of libfi_b200.

What changes relative to the reference, and what does not:

* same constructor arguments, sub-module names (``upsample``, ``feat_extract`` -> identical state_dict keys),
  same forward signature and return structure (``pooled_out, mask_out, feat_out``);
* the level rule, the four small / four big index lists and every count come from TWO kernel launches and ONE
  8-int device->host read per forward, instead of ~10 pointwise launches, 8 ``nonzero`` and 8 ``.any()`` syncs;
* 7x7 crops are written straight into their final row of ``pooled_out`` (the scatter of
  lib/sub_module.py:645-662 is fused into the RoIAlign kernel); so are the 14x14 crops of level 5;
* ``_assign_feat2cls`` is one deterministic segment-mean kernel instead of a python loop over classes;
* the buffer is a ring for BUFFER_SIZE > 1 (no 332 MB shift per iteration) and the class statistics can be
  all-reduced across ranks (one process per GPU) before the identical update on every rank.

The make-up layer (``upsample``) and the critic (``feat_extract``) are dense convolutions and stay stock
PyTorch / cuDNN.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .roi_align import crop_and_resize, crop_pair, crop_sets
from .roi_pool import RoIPoolFunction
from .dist import merged_class_sums, merged_class_sums_pair
from .ot import OptTrans

EPS = 1e-20


# ----------------------------------------------------------------------------------------------- level rule / split
def roi_level(rois, image_shape, base=224.0):
    """rois[bs,R,4] normalised -> int32 [bs,R] pyramid level in 2..5 (lib/sub_module.py:397-410)."""
    _lib.require_cuda(rois)
    flat = rois.detach().float().contiguous().view(-1, 4)
    level = torch.empty((flat.size(0),), device=flat.device, dtype=torch.int32)
    with torch.cuda.device(flat.device):
        _lib.check(_lib.lib().fi_roi_level(_lib.ptr(flat), flat.size(0), float(image_shape[0] * image_shape[1]), float(base),
                                           _lib.ptr(level), _lib.stream_ptr(flat.device)))
    return level.view(rois.shape[:-1])


class LevelSplit(object):
    """Per-level index lists of flat RoI ids (torch.nonzero order) for levels 2..5, optionally with the gathered boxes,
    image indices and class ids of every list (one kernel, one 8-int host read for all of it)."""

    def __init__(self, small_idx, small_cnt, big_idx, big_cnt, slot, gathered=None, img_cnt=None, sync=True):
        self.small_idx, self.big_idx, self.slot = small_idx, big_idx, slot
        self.g = gathered
        self.small_cnt_dev, self.big_cnt_dev = small_cnt, big_cnt      # int32 [4] each, on the device
        self.img_offsets = None
        if not sync:
            # No host read at all: every list keeps its full capacity (n) and its length stays on the device
            # (small_count(i) / big_count(i)); crop_sets / assign_feat2cls take those (fixed shapes, CUDA-graph capturable).
            n = small_idx.size(1)
            self.small_cnt, self.big_cnt = [n] * 4, [n] * 4
            return
        parts = [small_cnt, big_cnt] + ([img_cnt.view(-1)] if img_cnt is not None else [])
        counts = torch.cat(parts).tolist()                       # the only host sync of the split
        self.small_cnt, self.big_cnt = counts[:4], counts[4:8]
        if img_cnt is not None:                                  # per list: first member of every image (lists are image-major)
            nb = img_cnt.size(1)
            self.img_offsets = []
            for k in range(8):
                off, acc = [0], 0
                for v in counts[8 + k * nb: 8 + (k + 1) * nb]:
                    acc += v
                    off.append(acc)
                self.img_offsets.append(off)

    def small_count(self, i):
        """Length of list i as a device int32 [1] tensor."""
        return self.small_cnt_dev[i:i + 1]

    def big_count(self, i):
        return self.big_cnt_dev[i:i + 1]

    def small_img_offsets(self, i):
        return None if self.img_offsets is None else self.img_offsets[i]

    def big_img_offsets(self, i):
        return None if self.img_offsets is None else self.img_offsets[4 + i]

    def small(self, i):
        return self.small_idx[i, : self.small_cnt[i]]

    def big(self, i):
        return self.big_idx[i, : self.big_cnt[i]]

    def small_boxes(self, i):
        return self.g["small_boxes"][i, : self.small_cnt[i]]

    def small_ind(self, i):
        return self.g["small_ind"][i, : self.small_cnt[i]]

    def small_gt(self, i):
        return self.g["small_gt"][i, : self.small_cnt[i]]

    def big_boxes(self, i):
        return self.g["big_boxes"][i, : self.big_cnt[i]]

    def big_ind(self, i):
        return self.g["big_ind"][i, : self.big_cnt[i]]

    def big_gt(self, i):
        return self.g["big_gt"][i, : self.big_cnt[i]]


def spatial_order(rois, grid=None):
    """Visiting order that walks every image tile by tile (grid x grid coarse tiles by box centre): RoIs that overlap are
    then processed close together in time, so their taps / gradient lines are still in L2 (measured on C2: forward DRAM
    reads and backward read-modify-write traffic drop).  Results do not depend on the order (forward is
    bit-identical; backward sums in another order)."""
    if grid is None:
        grid = int(os.environ.get("FI_SORT_GRID", "8"))
    bs, R = rois.shape[0], rois.shape[1]
    if rois.is_cuda and rois.dim() == 3 and R <= 4096 and grid <= 128:
        flat = rois.detach().float().contiguous()
        order = torch.empty((bs * R,), device=rois.device, dtype=torch.int32)
        with torch.cuda.device(rois.device):
            _lib.check(_lib.lib().fi_spatial_order(_lib.ptr(flat), bs, R, int(grid), _lib.ptr(order), _lib.stream_ptr(rois.device)))
        return order
    # the same order with torch ops (CPU tensors in tests, more than 4096 RoIs per image)
    cy = ((rois[..., 0] + rois[..., 2]) * (0.5 * grid)).clamp(0, grid - 1).floor()
    cx = ((rois[..., 1] + rois[..., 3]) * (0.5 * grid)).clamp(0, grid - 1).floor()
    snake = torch.where(cy.long() % 2 == 0, cx, grid - 1 - cx)            # boustrophedon: neighbouring tiles stay neighbours
    key = (torch.arange(bs, device=rois.device).view(bs, 1) * (grid * grid) + cy * grid + snake).view(-1)
    return torch.sort(key, stable=True)[1].int()


def split_levels(level, rois=None, gt=None, order=None, sync=True):
    """level[...] int32 -> LevelSplit: small(l) = {level == l}, big(l) = {level > l} (lib/sub_module.py:442,367-378).
    With ``rois`` ([bs,R,4]) the same launch also gathers boxes / image index / class id (``gt`` [bs,R]) of every list;
    ``order`` (a permutation, e.g. ``spatial_order(rois)``) replaces torch.nonzero order by that visiting order.
    ``sync=False``: the list lengths are NOT read back; the lists keep their capacity and ``small_count(i)`` / ``big_count(i)``
    hand the device-side lengths to crop_sets / assign_feat2cls."""
    _lib.require_cuda(level)
    flat = level.contiguous().view(-1)
    n = flat.numel()
    dev = flat.device
    m = max(n, 1)
    small_idx = torch.empty((4, m), device=dev, dtype=torch.int32)
    big_idx = torch.empty((4, m), device=dev, dtype=torch.int32)
    small_cnt = torch.empty((4,), device=dev, dtype=torch.int32)
    big_cnt = torch.empty((4,), device=dev, dtype=torch.int32)
    slot = torch.empty((m,), device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        if rois is None:
            _lib.check(_lib.lib().fi_split_levels(_lib.ptr(flat), n, _lib.ptr(small_idx), _lib.ptr(small_cnt), _lib.ptr(big_idx),
                                                  _lib.ptr(big_cnt), _lib.ptr(slot), _lib.stream_ptr(dev)))
            return LevelSplit(small_idx, small_cnt, big_idx, big_cnt, slot, sync=sync)
        rois_flat = rois.detach().float().contiguous().view(-1, 4)
        gt_flat = None if gt is None else gt.detach().to(torch.int32).contiguous().view(-1)
        g = dict(small_boxes=torch.empty((4, m, 4), device=dev), big_boxes=torch.empty((4, m, 4), device=dev),
                 small_ind=torch.empty((4, m), device=dev, dtype=torch.int32), big_ind=torch.empty((4, m), device=dev, dtype=torch.int32))
        if gt_flat is not None:
            g["small_gt"] = torch.empty((4, m), device=dev, dtype=torch.int32)
            g["big_gt"] = torch.empty((4, m), device=dev, dtype=torch.int32)
        order = None if order is None else order.to(device=dev, dtype=torch.int32).contiguous()
        img_cnt = torch.empty((8, max(int(rois.size(0)), 1)), device=dev, dtype=torch.int32) if (rois.dim() == 3 and sync) else None
        _lib.check(_lib.lib().fi_split_levels_gather(
            _lib.ptr(flat), _lib.ptr(rois_flat), _lib.ptr(gt_flat), _lib.ptr(order), n, int(rois.size(-2)), _lib.ptr(small_idx), _lib.ptr(small_cnt),
            _lib.ptr(big_idx), _lib.ptr(big_cnt), _lib.ptr(slot), _lib.ptr(g["small_boxes"]), _lib.ptr(g["small_ind"]), _lib.ptr(g.get("small_gt")),
            _lib.ptr(g["big_boxes"]), _lib.ptr(g["big_ind"]), _lib.ptr(g.get("big_gt")), _lib.ptr(img_cnt), _lib.stream_ptr(dev)))
    # per-image extents are only meaningful when the visiting order is image-major (index order and spatial_order are)
    return LevelSplit(small_idx, small_cnt, big_idx, big_cnt, slot, gathered=g, img_cnt=img_cnt, sync=sync)


# ----------------------------------------------------------------------------------------------- segment mean
class _SegmentMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gt, feat, ncls, count):
        _lib.require_cuda(feat, gt)
        feat2 = feat.flatten(1).float().contiguous()
        gt = gt.detach().to(torch.int32).contiguous()
        k, Fd = feat2.shape
        if count is not None:
            count = count.detach().to(device=feat2.device, dtype=torch.int32).reshape(-1)[:1].contiguous()
        mean = torch.empty((Fd, ncls), device=feat2.device, dtype=torch.float32)
        cnt = torch.empty((1, ncls), device=feat2.device, dtype=torch.float32)
        with torch.cuda.device(feat2.device):
            _lib.check(_lib.lib().fi_segment_mean_forward_n(_lib.ptr(gt), _lib.ptr(feat2), k, _lib.ptr(count), Fd, ncls, _lib.ptr(mean),
                                                            _lib.ptr(cnt), _lib.stream_ptr(feat2.device)))
        ctx.save_for_backward(gt, cnt, count if count is not None else torch.empty(0))
        ctx.has_count = count is not None
        ctx.shape = tuple(feat.shape)
        ctx.dims = (k, Fd, ncls)
        ctx.mark_non_differentiable(cnt)
        return mean, cnt

    @staticmethod
    def backward(ctx, gmean, _gcnt):
        gt, cnt, count = ctx.saved_tensors
        k, Fd, ncls = ctx.dims
        gmean = gmean.contiguous()
        gfeat = torch.empty((k, Fd), device=gmean.device, dtype=torch.float32)
        with torch.cuda.device(gmean.device):
            _lib.check(_lib.lib().fi_segment_mean_backward_n(_lib.ptr(gt), _lib.ptr(gmean), _lib.ptr(cnt), k, _lib.ptr(count) if ctx.has_count else None,
                                                             Fd, ncls, _lib.ptr(gfeat), _lib.stream_ptr(gmean.device)))
        return None, gfeat.view(ctx.shape), None, None


class _SegmentMeanMulti(torch.autograd.Function):
    """Class means of several (gt, feat) lists in ONE launch each way.  apply(ncls, n, *gts, *counts, *feats) -> means..., cnts..."""

    @staticmethod
    def forward(ctx, ncls, n, *args):
        gts, counts, feats = args[:n], args[n:2 * n], args[2 * n:3 * n]
        dev = feats[0].device
        _lib.require_cuda(*feats)
        feat2 = [f.flatten(1).float().contiguous() for f in feats]
        gts = [g.detach().to(torch.int32).contiguous() for g in gts]
        counts = [None if c is None else c.detach().to(device=dev, dtype=torch.int32).reshape(-1)[:1].contiguous() for c in counts]
        Fd = feat2[0].size(1)
        if any(f.size(1) != Fd for f in feat2):
            raise _lib.FiError("assign_feat2cls_multi: all lists must share the feature width")
        means = [torch.empty((Fd, ncls), device=dev, dtype=torch.float32) for _ in range(n)]
        cnts = [torch.empty((1, ncls), device=dev, dtype=torch.float32) for _ in range(n)]
        arr = (_lib.SegSet * n)(*[_lib.SegSet(_lib.ptr(gts[i]), _lib.ptr(feat2[i]), feat2[i].size(0), _lib.ptr(counts[i]), _lib.ptr(means[i]), _lib.ptr(cnts[i]),
                                               None, None) for i in range(n)])
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fi_segment_mean_forward_batch(arr, n, Fd, ncls, _lib.stream_ptr(dev)))
        ctx.lists = (gts, counts, cnts)
        ctx.shapes = [tuple(f.shape) for f in feats]
        ctx.dims = (n, Fd, ncls)
        ctx.mark_non_differentiable(*cnts)
        ctx.set_materialize_grads(False)
        return tuple(means) + tuple(cnts)

    @staticmethod
    def backward(ctx, *grads):
        n, Fd, ncls = ctx.dims
        gts, counts, cnts = ctx.lists
        dev = cnts[0].device
        live = [i for i in range(n) if grads[i] is not None and ctx.needs_input_grad[2 + 2 * n + i]]
        out = [None] * n
        if live:
            gm = {i: grads[i].contiguous() for i in live}
            for i in live:
                out[i] = torch.empty((gts[i].numel(), Fd), device=dev, dtype=torch.float32)
            arr = (_lib.SegSet * len(live))(*[_lib.SegSet(_lib.ptr(gts[i]), None, gts[i].numel(), _lib.ptr(counts[i]), None, _lib.ptr(cnts[i]), _lib.ptr(gm[i]),
                                                           _lib.ptr(out[i])) for i in live])
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().fi_segment_mean_backward_batch(arr, len(live), Fd, ncls, _lib.stream_ptr(dev)))
        return (None, None) + (None,) * (2 * n) + tuple(None if o is None else o.view(ctx.shapes[i]) for i, o in enumerate(out))


def assign_feat2cls_multi(lists, num_classes):
    """``lists`` = [(box_gt, input_feat[, count]), ...] (at most 8) -> [(feat[F,ncls], cnt[1,ncls]), ...]: assign_feat2cls of every
    list in one launch forward and one backward."""
    n = len(lists)
    gts = [l[0] for l in lists]
    feats = [l[1] for l in lists]
    counts = [l[2] if len(l) > 2 else None for l in lists]
    res = _SegmentMeanMulti.apply(int(num_classes), n, *gts, *counts, *feats)
    return [(res[i], res[n + i]) for i in range(n)]


def assign_feat2cls(box_gt, input_feat, num_classes, count=None):
    """lib/sub_module.py:664-684: per-class mean of instance features -> (feat[F,ncls], cnt[1,ncls]); background skipped.
    ``count`` (a device int32): only the first ``count`` rows are in use (list length kept on the device)."""
    return _SegmentMean.apply(box_gt, input_feat, int(num_classes), count)


# ----------------------------------------------------------------------------------------------- Dev
class Dev(nn.Module):
    def __init__(self, config, depth):
        super().__init__()
        self.depth = depth
        self.use_dev = config.DEV.SWITCH
        self.pool_size = config.MRCNN.POOL_SIZE
        self.mask_pool_size = config.MRCNN.MASK_POOL_SIZE
        self.image_shape = config.DATA.IMAGE_SHAPE
        self.num_classs = config.DATASET.NUM_CLASSES
        self.config = config
        self.dis_upsample = config.DEV.DIS_UPSAMPLER
        self.structure = config.DEV.STRUCTURE
        self.roi_type = config.ROIS.METHOD
        self.roi_spatial_scale = [1. / 4, 1. / 8, 1. / 16, 1. / 32]
        # spatial_sort: visit RoIs tile by tile inside each image (L2 reuse) instead of in index order.  Off by default: the
        # rows of small_output_all / small_gt_all are then level-major in THAT order rather than torch.nonzero order
        # (lib/sub_module.py:586-600); every other output is unchanged.
        self.spatial_sort = False
        if self.use_dev:
            self.feat_pool_size = config.DEV.FEAT_BRANCH_POOL_SIZE
            self.upsample_fac = config.DEV.UPSAMPLE_FAC
            assert self.feat_pool_size % 2 == 0, 'pool size of feature branch has to be even'
            if not self.dis_upsample:
                if config.DEV.UPSAMPLE_FAC == 1.:
                    conv_opt = nn.Conv2d(depth, depth, kernel_size=3, padding=1)
                elif config.DEV.UPSAMPLE_FAC == 2.:
                    conv_opt = nn.ConvTranspose2d(depth, depth, kernel_size=3, stride=2, padding=1, output_padding=1)
                upsample_num = 4 if config.DEV.MULTI_UPSAMPLER else 1
                # one conv object shared by all make-up layers, as in the reference (sub_module.py:310-321)
                self.upsample = nn.ModuleList([nn.Sequential(conv_opt, nn.BatchNorm2d(depth), nn.ReLU(inplace=True))
                                               for _ in range(upsample_num)])
            if not config.DEV.BASELINE:
                ksize = int(self.feat_pool_size / 2)
                self.feat_extract = nn.Sequential(
                    nn.Conv2d(depth, 512, kernel_size=3, padding=1, stride=2), nn.BatchNorm2d(512), nn.ReLU(inplace=True),
                    nn.Conv2d(512, 1024, kernel_size=ksize, stride=1), nn.BatchNorm2d(1024), nn.ReLU(inplace=True),
                    nn.Conv2d(1024, 1024, kernel_size=1, stride=1), nn.BatchNorm2d(1024), nn.ReLU(inplace=True))
                if config.DEV.LOSS_CHOICE in ('l2', 'l1'):
                    self.last_op = nn.Sigmoid()
                elif config.DEV.LOSS_CHOICE == 'kl':
                    self.last_op = nn.Softmax(dim=1)
                if config.DEV.BIG_SUPERVISE:
                    self.big_fc_layer = nn.Linear(1024, self.num_classs)

    # -- RoI feature op in either flavour (lib/sub_module.py:500-507) ------------------------------------
    def _roi_op(self, level_i, size, fmap, boxes, box_ind, out=None, dst_row=None):
        if self.roi_type == 'roi_align':
            return crop_and_resize(fmap, boxes, box_ind, size, size, 0.0, out=out, dst_row=dst_row)
        pooled = RoIPoolFunction(size, size, self.roi_spatial_scale[level_i])(fmap, self._make_roi_pool_box_input(boxes, box_ind))
        if out is not None:
            out.index_copy_(0, dst_row.long(), pooled)
            return out
        return pooled

    def _critic(self, pooled):
        out = self.feat_extract(pooled)
        if self.config.DEV.LOSS_CHOICE != 'ot':
            out = self.last_op(out)
        return out

    def forward(self, x, rois, roi_cls_gt=None):
        """Same contract as the reference (lib/sub_module.py:380-642).  With channels_last maps and roi_align every crop of
        the pass is taken by ONE level-batched launch (crop_sets); otherwise level by level like the reference."""
        cfg = self.config
        if self.use_dev and self.structure == 'beta' and not cfg.DEV.ASSIGN_BOX_ON_ALL_SCALE and self.roi_type == 'roi_align' \
                and all(m.dim() == 4 and m.size(1) % 128 == 0 and m.is_contiguous(memory_format=torch.channels_last) and not m.is_contiguous()
                        for m in x[:4]):
            return self._forward_batched(x, rois, roi_cls_gt)
        return self._forward_per_level(x, rois, roi_cls_gt)

    def _forward_batched(self, x, rois, roi_cls_gt=None):
        cfg = self.config
        train_phase = roi_cls_gt is not None
        use_stats = train_phase and not cfg.DEV.BASELINE
        bs, R = rois.size(0), rois.size(1)
        total_box = bs * R
        dev = rois.device
        cl = torch.channels_last
        rois_flat = rois.detach().float().contiguous().view(total_box, 4)
        # one launch: level lists + rois[idx], idx // R, gt[idx] of every list (sub_module.py:489-493,541-548)
        split = split_levels(roi_level(rois, self.image_shape, cfg.ROIS.ASSIGN_ANCHOR_BASE), rois=rois, gt=roi_cls_gt if train_phase else None,
                             order=spatial_order(rois) if self.spatial_sort else None)
        # every RoI is assigned to exactly one level, so every row below is written by a crop: no zero fill (sub_module.py:650,656)
        pooled_out = torch.empty((total_box, self.depth, self.pool_size, self.pool_size), device=dev, memory_format=cl)
        mask_out = torch.empty((total_box, self.depth, self.mask_pool_size, self.mask_pool_size), device=dev, memory_format=cl)
        zf = lambda: torch.zeros(1024, self.num_classs, device=dev)
        zc = lambda: torch.zeros(1, self.num_classs, device=dev)

        # ---- pass 1: plan every crop of every level (sub_module.py:489-502,541-572)
        specs, plan = [], []
        for i, level in enumerate(range(2, 6)):
            use_meta = level in (2, 3, 4)
            want_critic = use_meta and not cfg.DEV.BASELINE
            n_small, n_big = split.small_cnt[i], split.big_cnt[i]
            info = dict(i=i, use_meta=use_meta, want_critic=want_critic, n_small=n_small, n_big=n_big, big=None, s7=None, s14=None)
            if n_small > 0:
                if use_stats and n_big > 0:
                    info["big"] = len(specs)
                    specs.append(dict(image=x[i], boxes=split.big_boxes(i), box_ind=split.big_ind(i), size=self.feat_pool_size,
                                      img_offsets=split.big_img_offsets(i)))
                s32 = split.small(i)
                feat_maps = self.upsample[i if cfg.DEV.MULTI_UPSAMPLER else 0](x[i]).contiguous(memory_format=cl)   # make-up layer
                boxes, ind = split.small_boxes(i), split.small_ind(i)
                info["s7"] = len(specs)
                specs.append(dict(image=feat_maps, boxes=boxes, box_ind=ind, size=self.pool_size, out=pooled_out, dst_row=s32,
                                  img_offsets=split.small_img_offsets(i)))
                info["s14"] = len(specs)
                specs.append(dict(image=feat_maps, boxes=boxes, box_ind=ind, size=self.mask_pool_size, out=mask_out, dst_row=s32,
                                  compact=want_critic, img_offsets=split.small_img_offsets(i)))
            plan.append(info)
        outs, comps = crop_sets(specs) if specs else ([], [])
        for info in plan:                          # the tensors after the (single) in-place node
            if info["s7"] is not None:
                pooled_out, mask_out = outs[info["s7"]], outs[info["s14"]]

        # ---- pass 2: critic, class statistics, in the reference's per-level order
        big_feat, big_cnt, small_feat, small_cnt, big_loss = [], [], [], [], []
        small_output_all = torch.zeros(total_box, 1024, device=dev)
        small_gt_all = torch.zeros(total_box, device=dev)
        small_out_cnt = 0
        for info in plan:
            use_meta = info["use_meta"]
            if info["n_small"] == 0:                                            # sub_module.py:456-467
                if use_meta and use_stats:
                    small_feat.append(zf()); small_cnt.append(zc()); big_feat.append(zf()); big_cnt.append(zc())
                    big_loss.append(torch.zeros(1, device=dev))
                continue
            if use_stats:                                                       # sub_module.py:472-536
                if info["n_big"] == 0:
                    if use_meta:
                        big_feat.append(zf()); big_cnt.append(zc()); big_loss.append(torch.zeros(1, device=dev))
                else:
                    big_box_gt = split.big_gt(info["i"])
                    big_before_last = self.feat_extract(comps[info["big"]])
                    big_output = big_before_last if cfg.DEV.LOSS_CHOICE == 'ot' else self.last_op(big_before_last)
                    b_feat, b_cnt = assign_feat2cls(big_box_gt, big_output, self.num_classs)
                    big_feat.append(b_feat); big_cnt.append(b_cnt)
                    if cfg.DEV.BIG_SUPERVISE:
                        digits = self.big_fc_layer(big_before_last.view(-1, 1024))
                        big_loss.append(F.cross_entropy(digits, big_box_gt.long()).view(1))
                    else:
                        big_loss.append(torch.zeros(1, device=dev))
            if info["want_critic"]:                                             # sub_module.py:580-600
                small_output = self._critic(comps[info["s14"]])
                n = info["n_small"]
                small_output_all[small_out_cnt:small_out_cnt + n, :] = small_output.view(n, -1)
                if train_phase:
                    small_box_gt = split.small_gt(info["i"])
                    s_feat, s_cnt = assign_feat2cls(small_box_gt, small_output, self.num_classs)
                    small_feat.append(s_feat); small_cnt.append(s_cnt)
                    small_gt_all[small_out_cnt:small_out_cnt + n] = small_box_gt.float()
                else:
                    small_gt_all[small_out_cnt:small_out_cnt + n] = 1
                small_out_cnt += n
        if use_stats:
            bf = torch.stack(big_feat).unsqueeze(dim=0)
            if cfg.DEV.BIG_FEAT_DETACH:
                bf = bf.detach()
            feat_out = [bf, torch.stack(big_cnt).unsqueeze(dim=0), torch.stack(small_feat).unsqueeze(dim=0),
                        torch.stack(small_cnt).unsqueeze(dim=0), torch.stack(big_loss).unsqueeze(dim=0),
                        small_output_all, small_gt_all]
        elif not train_phase:
            feat_out = [small_output_all, small_gt_all]
        else:
            feat_out = []
        return pooled_out, mask_out, feat_out

    def _forward_per_level(self, x, rois, roi_cls_gt=None):
        cfg = self.config
        base = cfg.ROIS.ASSIGN_ANCHOR_BASE
        if not self.use_dev:
            pooled_out = pyramid_roi_align([rois] + list(x), self.pool_size, self.image_shape, base=base)
            mask_out = pyramid_roi_align([rois] + list(x), self.mask_pool_size, self.image_shape, base=base)
            return pooled_out, mask_out, None
        if self.structure != 'beta' or cfg.DEV.ASSIGN_BOX_ON_ALL_SCALE:
            # the reference itself only implements 'beta' (sub_module.py:385-391,642)
            raise _lib.FiError("Dev: only STRUCTURE='beta' with ASSIGN_BOX_ON_ALL_SCALE=False is built")
        train_phase = roi_cls_gt is not None
        use_stats = train_phase and not cfg.DEV.BASELINE
        bs, R = rois.size(0), rois.size(1)
        total_box = bs * R
        dev = rois.device
        rois_flat = rois.detach().float().contiguous().view(total_box, 4)
        gt_flat = roi_cls_gt.contiguous().view(total_box) if train_phase else None

        split = split_levels(roi_level(rois, self.image_shape, base))          # steps 1+2a: no per-level syncs below
        fmt = torch.channels_last if x[0].is_contiguous(memory_format=torch.channels_last) and not x[0].is_contiguous() \
            else torch.contiguous_format
        # every RoI is assigned to exactly one level, so every row below is written by a crop: no zero fill (sub_module.py:650,656)
        pooled_out = torch.empty((total_box, self.depth, self.pool_size, self.pool_size), device=dev, memory_format=fmt)
        mask_out = torch.empty((total_box, self.depth, self.mask_pool_size, self.mask_pool_size), device=dev, memory_format=fmt)
        big_feat, big_cnt, small_feat, small_cnt, big_loss = [], [], [], [], []
        small_output_all = torch.zeros(total_box, 1024, device=dev)
        small_gt_all = torch.zeros(total_box, device=dev)
        small_out_cnt = 0
        zf = lambda: torch.zeros(1024, self.num_classs, device=dev)
        zc = lambda: torch.zeros(1, self.num_classs, device=dev)

        for i, level in enumerate(range(2, 6)):
            curr_feat_maps = x[i]
            use_meta = level in (2, 3, 4)
            n_small, n_big = split.small_cnt[i], split.big_cnt[i]
            if n_small == 0:                                                    # sub_module.py:456-467
                if use_meta and use_stats:
                    small_feat.append(zf()); small_cnt.append(zc()); big_feat.append(zf()); big_cnt.append(zc())
                    big_loss.append(torch.zeros(1, device=dev))
                continue
            if use_stats:                                                       # sub_module.py:472-536 (reliable set)
                if n_big == 0:
                    if use_meta:
                        big_feat.append(zf()); big_cnt.append(zc()); big_loss.append(torch.zeros(1, device=dev))
                else:
                    bidx = split.big(i).long()
                    big_boxes = rois_flat[bidx]
                    big_box_gt = gt_flat[bidx]
                    big_box_ind = (bidx // R).int()
                    big_pooled = self._roi_op(i, self.feat_pool_size, curr_feat_maps, big_boxes, big_box_ind)
                    big_before_last = self.feat_extract(big_pooled)
                    big_output = big_before_last if cfg.DEV.LOSS_CHOICE == 'ot' else self.last_op(big_before_last)
                    b_feat, b_cnt = assign_feat2cls(big_box_gt, big_output, self.num_classs)
                    big_feat.append(b_feat); big_cnt.append(b_cnt)
                    if cfg.DEV.BIG_SUPERVISE:
                        digits = self.big_fc_layer(big_before_last.view(-1, 1024))
                        big_loss.append(F.cross_entropy(digits, big_box_gt.long()).view(1))
                    else:
                        big_loss.append(torch.zeros(1, device=dev))
            # less-reliable set: RoIs assigned to this level (sub_module.py:539-600)
            sidx32 = split.small(i)
            sidx = sidx32.long()
            small_boxes = rois_flat[sidx]
            box_ind = (sidx // R).int()
            feat_maps = self.upsample[i if cfg.DEV.MULTI_UPSAMPLER else 0](curr_feat_maps).contiguous(memory_format=fmt)
            want_critic = use_meta and not cfg.DEV.BASELINE
            fused = self.roi_type == 'roi_align' and fmt == torch.channels_last and feat_maps.size(1) % 128 == 0
            if fused:
                # both crops of this level in one autograd node: 7x7 and 14x14 land directly in their final (image, roi)
                # rows (fused _reshape_result), the 14x14 one is also emitted compact for the critic, ONE backward pass
                res = crop_pair(feat_maps, small_boxes, box_ind, sidx32, pooled_out, self.pool_size, mask_out, self.mask_pool_size,
                                compact_b=want_critic)
                pooled_out, mask_out = res[0], res[1]
                mask_and_feat = res[2] if want_critic else None
            else:
                pooled_out = self._roi_op(i, self.pool_size, feat_maps, small_boxes, box_ind, out=pooled_out, dst_row=sidx32)
                if want_critic:
                    mask_and_feat = self._roi_op(i, self.mask_pool_size, feat_maps, small_boxes, box_ind)
                    mask_out.index_copy_(0, sidx, mask_and_feat)
                else:
                    mask_out = self._roi_op(i, self.mask_pool_size, feat_maps, small_boxes, box_ind, out=mask_out, dst_row=sidx32)
            if want_critic:
                small_output = self._critic(mask_and_feat)
                n = n_small
                small_output_all[small_out_cnt:small_out_cnt + n, :] = small_output.view(n, -1)
                if train_phase:
                    small_box_gt = gt_flat[sidx]
                    s_feat, s_cnt = assign_feat2cls(small_box_gt, small_output, self.num_classs)
                    small_feat.append(s_feat); small_cnt.append(s_cnt)
                    small_gt_all[small_out_cnt:small_out_cnt + n] = small_box_gt.float()
                else:
                    small_gt_all[small_out_cnt:small_out_cnt + n] = 1
                small_out_cnt += n

        if use_stats:
            bf = torch.stack(big_feat).unsqueeze(dim=0)
            if cfg.DEV.BIG_FEAT_DETACH:
                bf = bf.detach()
            feat_out = [bf, torch.stack(big_cnt).unsqueeze(dim=0), torch.stack(small_feat).unsqueeze(dim=0),
                        torch.stack(small_cnt).unsqueeze(dim=0), torch.stack(big_loss).unsqueeze(dim=0),
                        small_output_all, small_gt_all]
        elif not train_phase:
            feat_out = [small_output_all, small_gt_all]
        else:
            feat_out = []
        return pooled_out, mask_out, feat_out

    def _make_roi_pool_box_input(self, boxes, box_ind):
        """lib/sub_module.py:686-692: (b, x1, y1, x2, y2) in pixels; both axes scaled by image_shape[0] as the reference does."""
        b = boxes * float(self.image_shape[0])
        return torch.stack([box_ind.float(), b[:, 1], b[:, 0], b[:, 3], b[:, 2]], dim=1)


def pyramid_roi_align(inputs, pool_size, image_shape, base=224.):
    """lib/layers.py:145-218: vanilla FPN RoIAlign (intertwiner off); crops written straight to their final rows."""
    boxes, feature_maps = inputs[0], inputs[1:]
    bs, R = boxes.size(0), boxes.size(1)
    split = split_levels(roi_level(boxes, image_shape, base))
    flat = boxes.detach().float().contiguous().view(bs * R, 4)
    fm0 = feature_maps[0]
    fmt = torch.channels_last if fm0.is_contiguous(memory_format=torch.channels_last) and not fm0.is_contiguous() else torch.contiguous_format
    out = torch.zeros((bs * R, fm0.size(1), pool_size, pool_size), device=boxes.device).contiguous(memory_format=fmt)
    for i in range(4):
        if split.small_cnt[i] == 0:
            continue
        idx32 = split.small(i)
        idx = idx32.long()
        out = crop_and_resize(feature_maps[i], flat[idx], (idx // R).int(), pool_size, pool_size, 0.0, out=out, dst_row=idx32)
    return out


# ----------------------------------------------------------------------------------------------- meta loss
def _linear_relu(bias, x, w, out):
    """out = relu(x @ w^T + bias): one cuBLASLt call with the bias + ReLU epilogue where torch offers it, else addmm + relu_."""
    fused = getattr(torch, "_addmm_activation", None)
    if fused is not None:
        fused(bias, x, w.t(), out=out)
    else:
        torch.addmm(bias, x, w.t(), out=out)
        out.relu_()
    return out


class _ClassOTHead(torch.autograd.Function):
    """The class-level OT loss head after the buffer update, 1-D OptTrans (lib/model.py:176-207 + lib/OT_module.py:67-102), with
    every pointwise step a kernel of csrc/loss_head.cu and the dense products library GEMMs:

        X = (small_sum / (small_n + EPS))^T[1:], Y = final_big^T[1:], mask            fi_ot_head_prep
        H = relu(X Wg^T + bg);  [cx; cy] = relu([H; Y] Wc^T + bc)                     two addmm (+ relu_), W = Conv1d weight[:, :, 1]
        w = Sinkhorn((cx,cy), (cx,cx), (cy,cy))                                       fi_sinkhorn_ws (one launch for the 3n problems)
        loss = (2 w1 - w2 - w3) * mask                                                fi_ot_head_combine

    and the same chain backwards by hand (fi_ot_head_dcritic, fi_relu_mask, fi_ot_head_dsum, fi_centre_tap_embed + five GEMMs).
    Same arithmetic as IntertwinerLoss._head -> OptTrans.forward in torch ops (tests/test_ops_gpu.py::test_fused_loss_head_matches_torch_ops):
    ~27 launches per iteration instead of ~80."""

    @staticmethod
    def forward(ctx, mod, final_big, s_sum, s_n, Wg, bg, Wc, bc):
        from .ot import sinkhorn_raw
        L_ = _lib.lib()
        dev = s_sum.device
        Fd, ncls = s_sum.shape
        n, N = ncls - 1, Wc.size(0)
        final_big, s_sum, s_n = final_big.contiguous(), s_sum.detach().contiguous(), s_n.detach().contiguous()
        buffer_cnt = mod.buffer_cnt.contiguous()
        X = torch.empty((n, Fd), device=dev, dtype=torch.float32)
        Z = torch.empty((2 * n, Fd), device=dev, dtype=torch.float32)          # [H; Y]: what the critic sees
        mask = torch.empty((n,), device=dev, dtype=torch.float32)
        st = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            _lib.check(L_.fi_ot_head_prep(_lib.ptr(final_big), _lib.ptr(s_sum), _lib.ptr(s_n), _lib.ptr(buffer_cnt), buffer_cnt.size(0), Fd, ncls,
                                          _lib.ptr(X), _lib.ptr(Z[n:]), _lib.ptr(mask), st))
        Wg1, Wc1 = Wg.detach()[:, :, 1].contiguous(), Wc.detach()[:, :, 1].contiguous()    # Conv1d(k=3, pad=1) on length 1: the centre tap
        Cc = torch.empty((2 * n, N), device=dev, dtype=torch.float32)
        _linear_relu(bg.detach(), X, Wg1, Z[:n])
        _linear_relu(bc.detach(), Z, Wc1, Cc)
        cx, cy = Cc[:n], Cc[n:]
        need = any(ctx.needs_input_grad)
        w, gx, gy = sinkhorn_raw(torch.cat([cx, cx, cy]).unsqueeze(2), torch.cat([cy, cx, cy]).unsqueeze(2), mod.ot_loss.epsilon, mod.ot_loss.L, need)
        loss = torch.empty((n,), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(L_.fi_ot_head_combine(_lib.ptr(w), _lib.ptr(mask), n, _lib.ptr(loss), st))
        if need:
            ctx.save_for_backward(X, Z, Cc, gx, gy, mask, Wg1, Wc1, s_n)
        ctx.dims = (Fd, ncls, n, N)
        ctx.wshapes = (tuple(Wg.shape), tuple(Wc.shape))
        mod.last_idx, mod.last_mask = None, mask
        return loss

    @staticmethod
    def backward(ctx, g):
        X, Z, Cc, gx, gy, mask, Wg1, Wc1, s_n = ctx.saved_tensors
        Fd, ncls, n, N = ctx.dims
        L_ = _lib.lib()
        dev = X.device
        st = _lib.stream_ptr(dev)
        need = ctx.needs_input_grad           # mod, final_big, s_sum, s_n, Wg, bg, Wc, bc
        g = g.contiguous()
        dC = torch.empty((2 * n, N), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(L_.fi_ot_head_dcritic(_lib.ptr(gx), _lib.ptr(gy), _lib.ptr(g), _lib.ptr(mask), _lib.ptr(Cc), n, N, _lib.ptr(dC), st))
            dbc = None
            if need[7]:
                dbc = torch.empty((N,), device=dev, dtype=torch.float32)
                _lib.check(L_.fi_col_sum(_lib.ptr(dC), 2 * n, N, _lib.ptr(dbc), st))
            dWc = None
            if need[6]:
                dWc = torch.empty(ctx.wshapes[1], device=dev, dtype=torch.float32)
                dWc1 = dC.t().mm(Z)
                _lib.check(L_.fi_centre_tap_embed(_lib.ptr(dWc1), dWc1.numel(), _lib.ptr(dWc), st))
            dss = dWg = dbg = None
            if need[2] or need[4] or need[5]:
                dH = dC[:n].mm(Wc1)
                _lib.check(L_.fi_relu_mask(_lib.ptr(dH), _lib.ptr(Z), dH.numel(), st))          # Z[:n] = H
                if need[5]:
                    dbg = torch.empty((Fd,), device=dev, dtype=torch.float32)
                    _lib.check(L_.fi_col_sum(_lib.ptr(dH), n, Fd, _lib.ptr(dbg), st))
                if need[4]:
                    dWg = torch.empty(ctx.wshapes[0], device=dev, dtype=torch.float32)
                    dWg1 = dH.t().mm(X)
                    _lib.check(L_.fi_centre_tap_embed(_lib.ptr(dWg1), dWg1.numel(), _lib.ptr(dWg), st))
                if need[2]:
                    dX = dH.mm(Wg1)
                    dss = torch.empty((Fd, ncls), device=dev, dtype=torch.float32)
                    _lib.check(L_.fi_ot_head_dsum(_lib.ptr(dX), _lib.ptr(s_n), Fd, ncls, _lib.ptr(dss), st))
        return None, None, dss, None, dWg, dbg, dWc, dbc


class IntertwinerLoss(nn.Module):
    """``MaskRCNN.initialize_buffer`` + ``meta_loss`` + ``_merge_feat_vec`` (lib/model.py:106-111,143-224) as a module.

    ``forward(feat_input)`` takes exactly what the reference's ``meta_loss`` takes:
    ``[big_feat, big_cnt, small_feat, small_cnt, small_output_all, small_gt_all]`` with the leading
    ``[gpu_num, scale_num]`` axes of lib/model.py:394-402.  With ``process_group`` set (one process per GPU) the
    un-normalised class statistics are all-reduced first, replacing the gather-to-GPU-0 of nn.DataParallel.
    """

    def __init__(self, config, ot_loss=None, feat_dim=1024, process_group=None, distributed=False, ddp_compensate=True,
                 ot_padded=False):
        super().__init__()
        self.config = config
        self.feat_dim = feat_dim
        self.distributed = distributed
        self.process_group = process_group
        self.ddp_compensate = ddp_compensate
        # ot_padded: class-level OT loss over ALL foreground classes with the absent ones masked to 0 -> fixed shapes, no
        # `nonzero` host sync; returns [ncls-1] instead of the reference's [n] (same sum, same gradients)
        self.ot_padded = ot_padded
        # fused_head: the padded class-level OT head as ~27 launches (csrc/loss_head.cu + library GEMMs) instead of ~80 torch ops
        self.fused_head = True
        B, ncls = config.DEV.BUFFER_SIZE, config.DATASET.NUM_CLASSES
        # persistent state of the hot path; round-trips through checkpoints like tools/utils.py:374-389,575-585
        self.register_buffer('buffer', torch.zeros(B, feat_dim, ncls))
        self.register_buffer('buffer_cnt', torch.zeros(B, 1, ncls))
        self.register_buffer('ring_pos', torch.zeros((), dtype=torch.long))   # next slot to overwrite (B > 1)
        fg = torch.ones(ncls); fg[0] = 0
        self.register_buffer('fg_mask', fg, persistent=False)                  # background excluded (model.py:178)
        if ot_loss is not None:
            self.ot_loss = ot_loss
        self.last_idx = None

    def initialize_buffer(self, log_file=None):
        self.buffer.zero_(); self.buffer_cnt.zero_(); self.ring_pos.zero_()

    def fifo_buffer(self):
        """(buffer, buffer_cnt) in the reference's oldest-first order, e.g. for a reference-format checkpoint."""
        B = self.buffer.size(0)
        if B == 1:
            return self.buffer, self.buffer_cnt
        order = (torch.arange(B, device=self.buffer.device) + self.ring_pos) % B
        return self.buffer[order], self.buffer_cnt[order]

    def _sums(self, feat, cnt, differentiable):
        """_merge_feat_vec's numerator and denominator (lib/model.py:219-222), all-reduced when distributed."""
        return merged_class_sums(feat, cnt, self.process_group if self.distributed else None, self.distributed,
                                 differentiable, self.ddp_compensate)

    # ---- optional CUDA-graph capture of the loss head ------------------------------------------------------------
    def enable_cuda_graph(self, feat_input):
        """Capture forward AND backward of the loss head (buffer update -> match -> OptTrans / Sinkhorn) into CUDA graphs:
        ~60 small launches per iteration become two graph launches.  Possible when every shape is fixed: class-level loss,
        l1 / l2 or padded OT, BUFFER_SIZE == 1.  The statistics merge (and, when distributed, its all-reduce) stays eager in
        front of the graph, so the captured part contains no collective and the same graph serves 1 and N processes.
        ``feat_input`` is a representative input (shapes / requires_grad as in training).  Returns True when the graphed path
        is active; falls back to eager on any failure.  The buffer is snapshotted around the warm-up replays, so capture does
        not disturb the running means."""
        cfg = self.config
        ok = self.buffer.size(0) == 1 and not cfg.DEV.INST_LOSS and \
            (cfg.DEV.LOSS_CHOICE in ('l1', 'l2') or (cfg.DEV.LOSS_CHOICE == 'ot' and self.ot_padded))
        if not ok:
            return False
        snap = (self.buffer.clone(), self.buffer_cnt.clone())
        try:
            with torch.no_grad():
                sums = self._stage_sums([t.detach() for t in feat_input[:4]])
            head = _LossHead(self)
            sample = (sums[0].clone(), sums[1].clone(), sums[2].clone().requires_grad_(), sums[3].clone())
            graphed = torch.cuda.make_graphed_callables(head, sample)
            self._graphed = graphed
            self._graph_shapes = tuple(tuple(t.shape) for t in sample)
        except Exception as exc:           # noqa: BLE001 -- capture is an optimisation, never a requirement
            self._graphed = None
            self._graph_error = repr(exc)
        finally:
            self.buffer.copy_(snap[0]); self.buffer_cnt.copy_(snap[1])
        return self._graphed is not None

    def forward(self, feat_input):
        g = getattr(self, '_graphed', None)
        if g is not None and torch.is_grad_enabled() and feat_input[2].requires_grad:
            big_sum, big_n, s_sum, s_n = self._stage_sums(feat_input)
            if tuple(tuple(t.shape) for t in (big_sum, big_n, s_sum, s_n)) == self._graph_shapes:
                return g(big_sum.detach(), big_n.detach(), s_sum, s_n.detach())
            return self._head(big_sum, big_n, s_sum, s_n, feat_input[4], feat_input[5])
        return self._forward_eager(feat_input)

    def _stage_sums(self, feat_input):
        """The exchange step: un-normalised class statistics of the reliable and of the less-reliable set, summed over
        (gpu, scale) and -- when distributed -- all-reduced over the ranks in ONE collective (lib/model.py:151,177,217-224)."""
        big_feat, big_cnt, small_feat, small_cnt = feat_input[:4]
        if self.config.DEV.INST_LOSS:
            big_sum, big_n = self._sums(big_feat.detach(), big_cnt.detach(), differentiable=False)
            return big_sum.contiguous(), big_n.contiguous(), None, None
        big_sum, big_n, s_sum, s_n = merged_class_sums_pair(big_feat, big_cnt, small_feat, small_cnt, self.process_group if self.distributed else None,
                                                            self.distributed, self.ddp_compensate)
        return big_sum.contiguous(), big_n.contiguous(), s_sum, s_n

    def _forward_eager(self, feat_input):
        big_sum, big_n, s_sum, s_n = self._stage_sums(feat_input)
        return self._head(big_sum, big_n, s_sum, s_n, feat_input[4], feat_input[5])

    def _fused_head_ok(self, s_sum, s_n):
        """The kernel-fused head (_ClassOTHead) covers the configuration every runnable reference config uses: 1-D OptTrans with the
        Conv1d critic (OT_ONE_DIM_FORM 'conv'), debiased loss, features of the buffer's width in and out of G_net."""
        ot = getattr(self, 'ot_loss', None)
        if not (self.fused_head and isinstance(ot, OptTrans) and s_sum is not None and s_sum.is_cuda and s_sum.dtype == torch.float32):
            return False
        if ot.two_dim or ot.remove_bias or ot.skip_critic or not isinstance(getattr(ot, 'critic', None), nn.Sequential):
            return False
        g0, c0 = ot.G_net[0], ot.critic[0]
        Fd = self.feat_dim
        return isinstance(g0, nn.Conv1d) and isinstance(c0, nn.Conv1d) and tuple(g0.weight.shape) == (Fd, Fd, 3) and \
            tuple(c0.weight.shape[1:]) == (Fd, 3) and g0.bias is not None and c0.bias is not None and \
            tuple(s_sum.shape) == (Fd, self.config.DATASET.NUM_CLASSES) and s_n.numel() == s_sum.size(1)

    def _head(self, big_sum, big_n, s_sum, s_n, small_output_all, small_gt_all):
        cfg = self.config
        Fd, ncls = self.feat_dim, cfg.DATASET.NUM_CLASSES
        B = self.buffer.size(0)
        dev = self.buffer.device
        # ---- reliable-set statistics -> historical buffer (model.py:148-166), one kernel
        big_sum, big_n = big_sum.contiguous(), big_n.contiguous()
        final_big = torch.empty((Fd, ncls), device=dev, dtype=torch.float32)
        slot = int(self.ring_pos.item()) if B > 1 else 0
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fi_buffer_update(_lib.ptr(big_sum), _lib.ptr(big_n), B, slot, Fd, ncls, _lib.ptr(self.buffer),
                                                   _lib.ptr(self.buffer_cnt), _lib.ptr(final_big), _lib.stream_ptr(dev)))
        if B > 1:
            self.ring_pos.fill_((slot + 1) % B)
        lc = cfg.DEV.LOSS_CHOICE
        if lc == 'ot' and self.ot_padded and not cfg.DEV.INST_LOSS and self._fused_head_ok(s_sum, s_n):
            ot = self.ot_loss
            return _ClassOTHead.apply(self, final_big, s_sum, s_n, ot.G_net[0].weight, ot.G_net[0].bias, ot.critic[0].weight, ot.critic[0].bias)
        in_buffer = self.buffer_cnt.sum(dim=0).view(-1) > 0
        # ---- comparison set (model.py:168-190)
        if cfg.DEV.INST_LOSS:
            gt = small_gt_all.long()
            mask = (gt != 0) & in_buffer[gt]
            SMALL_all, BIG_all = small_output_all, final_big.t()[gt]
        else:
            final_small = s_sum / (s_n + EPS)
            s_n = s_n * self.fg_mask                                            # background excluded (model.py:178); no host scalar
            mask = (s_n > 0) & in_buffer
            SMALL_all, BIG_all = final_small.t(), final_big.t()
        if lc == 'ot' and self.ot_padded and not cfg.DEV.INST_LOSS:
            self.last_idx, self.last_mask = None, mask
            w = self.ot_loss(SMALL_all[1:].unsqueeze(dim=-1).contiguous(), BIG_all[1:].unsqueeze(dim=-1).contiguous())
            return w * mask[1:].to(w.dtype)
        if lc == 'ot' or lc == 'kl':
            idx = torch.nonzero(mask).squeeze(1)                               # host-visible size: the reference's API returns [n]
            self.last_idx = idx
            if idx.numel() == 0:
                return torch.zeros(1, device=dev)
            SMALL, BIG = SMALL_all[idx], BIG_all[idx]
            if lc == 'kl':
                return F.kl_div(torch.log(SMALL), BIG, reduction='mean')
            return self.ot_loss(SMALL.unsqueeze(dim=-1), BIG.unsqueeze(dim=-1).contiguous())
        # l1 / l2: masked mean == F.mse_loss(SMALL[idx], BIG[idx]) without materialising idx (no host sync);
        # an empty comparison set gives 0 like model.py:208-209
        self.last_idx = None
        self.last_mask = mask
        diff = SMALL_all - BIG_all
        per = diff * diff if lc == 'l2' else diff.abs()
        m = mask.to(per.dtype).unsqueeze(1)
        err, cnt = (per * m).sum(), m.sum() * per.size(1)
        if cfg.DEV.INST_LOSS and self.distributed:
            # instance level under sharding: every rank scores its own rows of small_output_all against the
            # replicated buffer; (sum of errors, element count) are all-reduced so that every rank holds the mean over ALL instances
            from .dist import _AllReduceSum
            pair = _AllReduceSum.apply(torch.stack([err, cnt.to(err.dtype)]), self.process_group, self.ddp_compensate)
            err, cnt = pair[0], pair[1].detach()
        return err / cnt.clamp(min=1.0)


class _LossHead(nn.Module):
    """The fixed-shape part of IntertwinerLoss (everything after the statistics merge / all-reduce) as a 4-tensor callable for
    torch.cuda.make_graphed_callables.  Shares the
    OptTrans module (so its parameters receive gradients) and reaches the buffers through the parent."""

    def __init__(self, parent):
        super().__init__()
        object.__setattr__(self, '_parent', parent)
        if hasattr(parent, 'ot_loss'):
            self.ot_loss = parent.ot_loss

    def forward(self, big_sum, big_n, s_sum, s_n):
        return self._parent._head(big_sum, big_n, s_sum, s_n, None, None)


def meta_loss_module(config, ot_loss=None, **kw):
    return IntertwinerLoss(config, ot_loss=ot_loss, **kw)
