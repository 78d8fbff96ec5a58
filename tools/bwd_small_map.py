"""One RoIAlign backward call on a SMALL map (C2's level-5 and level-4 'small' sets, 14x14) through the drop-in launcher entry,
a few times -- for `ncu --metrics gpu__time_duration.sum` (which of the launch's kernels carries the time when the map has fewer
tiles than the persistent grid has CTAs)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import feature_intertwiner_b200 as fi  # noqa: E402
from feature_intertwiner_b200 import _lib, synth  # noqa: E402

wl = synth.WORKLOADS["c2"]
g = torch.Generator().manual_seed(2000)
B, R, hw = wl["batch"], wl["rois_per_image"], wl["image"]
rois = synth.make_rois(B, R, hw, g).cuda()
maps = [m.cuda() for m in synth.make_feature_maps(B, hw, 256, g, channels_last=True)]
split = fi.split_levels(fi.roi_level(rois, (hw[0], hw[1], 3)))
flat = rois.view(-1, 4)
s = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for lvl in (3, 2):
    idx = split.small(lvl).long()
    boxes, ind = flat[idx].contiguous(), (idx // R).int().contiguous()
    n, (Hh, Ww) = boxes.size(0), maps[lvl].shape[2:]
    grads = torch.randn(n, 256, 14, 14, device="cuda").contiguous(memory_format=torch.channels_last)
    gimg = torch.empty_like(maps[lvl])
    for it in range(3):
        flush.add_(1.0)
        _lib.check(_lib.lib().fi_crop_and_resize_backward(grads.data_ptr(), _lib.FI_LAYOUT_NHWC, boxes.data_ptr(), ind.data_ptr(), None, n, B, Hh, Ww,
                                                          14, 14, 256, gimg.data_ptr(), _lib.FI_LAYOUT_NHWC, 0, s))
    torch.cuda.synchronize()
    print("level", lvl + 2, "boxes", n, "map", Hh, Ww, flush=True)
