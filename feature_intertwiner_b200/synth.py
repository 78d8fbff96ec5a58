"""Seeded synthetic inputs of the hot path; shared by bench.py and the tests.

Everything is generated on the CPU with a seeded ``torch.Generator`` (seed 2000, the reference's own
MISC.SEED, lib/config.py:264) so the GPU arm and the CPU-reference arm see identical inputs.
"""
import math

import torch

# BASELINE.json configs -> shapes.  (H, W) is the padded image; FPN strides 4,8,16,32 (lib/config.py:318).
WORKLOADS = {
    # C1: the reference's own CPU-runnable case (configs/105/meta_105_quick_1 shapes)
    "c1": dict(batch=4, image=(256, 256), rois_per_image=64, sinkhorn_iters=5),
    # C2: the configuration the metric is quoted on
    "c2": dict(batch=8, image=(832, 1344), rois_per_image=512, sinkhorn_iters=50),
    # C3 (= C4 per GPU): the RoIs come out of the proposal layer (decode + NMS) inside the step
    "c3": dict(batch=4, image=(832, 1344), rois_per_image=1000, sinkhorn_iters=50, proposals=True),
    "c5": dict(batch=4, image=(832, 1344), rois_per_image=2000, sinkhorn_iters=100),
}
STRIDES = (4, 8, 16, 32)


def level_shapes(image_hw):
    return [(int(math.ceil(image_hw[0] / s)), int(math.ceil(image_hw[1] / s))) for s in STRIDES]


def make_rois(batch, rois_per_image, image_hw, gen, zero_frac=0.05, straddle_frac=0.01):
    """[batch, R, 4] normalised (y1,x1,y2,x2): side log-uniform in [16,512] px, aspect log-uniform in [1/2,2],
    boxes inside the image; the last ``zero_frac`` rows all-zero (the reference zero-pads, lib/layers.py:413,427);
    ``straddle_frac`` of the boxes pushed across the border (exercises the extrapolation branch)."""
    H, W = image_hw
    R = rois_per_image
    side = torch.exp(torch.empty(batch, R).uniform_(math.log(16.0), math.log(512.0), generator=gen))
    aspect = torch.exp(torch.empty(batch, R).uniform_(math.log(0.5), math.log(2.0), generator=gen))
    h = (side * aspect.sqrt()).clamp(max=H - 1.0)
    w = (side / aspect.sqrt()).clamp(max=W - 1.0)
    cy = torch.rand(batch, R, generator=gen) * (H - 1 - h) + h / 2
    cx = torch.rand(batch, R, generator=gen) * (W - 1 - w) + w / 2
    rois = torch.stack([(cy - h / 2) / H, (cx - w / 2) / W, (cy + h / 2) / H, (cx + w / 2) / W], dim=2)
    n_straddle = int(round(straddle_frac * R))
    if n_straddle:
        rois[:, :n_straddle, 1] -= 0.05   # x1 < 0
        rois[:, :n_straddle, 2] += 0.05   # y2 > 1
    n_zero = int(round(zero_frac * R))
    if n_zero:
        rois[:, R - n_zero:, :] = 0.0
    return rois.float().contiguous()


def make_class_ids(batch, rois_per_image, gen, num_classes=81, positive_ratio=0.33):
    """[batch, R] int: the first 33 % of each image's RoIs positive (ROIS.ROI_POSITIVE_RATIO, lib/config.py:142),
    uniform in 1..num_classes-1; the rest background."""
    ids = torch.zeros(batch, rois_per_image, dtype=torch.int32)
    n_pos = int(positive_ratio * rois_per_image)
    ids[:, :n_pos] = torch.randint(1, num_classes, (batch, n_pos), generator=gen, dtype=torch.int32)
    return ids


def make_feature_maps(batch, image_hw, channels, gen, channels_last=True):
    """FPN P2..P5, fp32 N(0,1) ``[batch, channels, H/s, W/s]`` (logical NCHW; channels_last memory by default)."""
    maps = []
    for (h, w) in level_shapes(image_hw):
        t = torch.randn(batch, channels, h, w, generator=gen)
        maps.append(t.contiguous(memory_format=torch.channels_last) if channels_last else t)
    return maps


def make_nms_boxes(n_images, n, gen, extent=1024.0):
    """[n_images, n, 5] (y1,x1,y2,x2,score) pixel boxes, each image sorted by descending score
    (callers pre-sort, lib/layers.py:103)."""
    ctr = torch.rand(n_images, n, 2, generator=gen) * extent
    size = torch.exp(torch.empty(n_images, n, 2).uniform_(math.log(16.0), math.log(400.0), generator=gen))
    # strictly decreasing scores: fp32 rand collides at these sizes and the reference's sort is unstable on ties
    score = torch.linspace(1.0, 0.0, n).expand(n_images, n).contiguous()
    lo = (ctr - size / 2).clamp(0, extent)
    hi = (ctr + size / 2).clamp(0, extent)
    return torch.cat([lo, hi, score.unsqueeze(2)], dim=2).float().contiguous()
