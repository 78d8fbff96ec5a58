"""Run ONE RoIAlign call of a workload a few times (for `ncu --set full -k regex:crop_`):

    python tools/profile_case.py --level 2 --kind big --P 14 --dir fwd [--fmt nhwc]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import feature_intertwiner_b200 as fi  # noqa: E402
from feature_intertwiner_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--level", type=int, default=2)
ap.add_argument("--kind", default="big")
ap.add_argument("--P", type=int, default=14)
ap.add_argument("--dir", default="both")
ap.add_argument("--fmt", default="nhwc")
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
wl = synth.WORKLOADS[a.workload]
g = torch.Generator().manual_seed(2000)
B, R, hw = wl["batch"], wl["rois_per_image"], wl["image"]
rois = synth.make_rois(B, R, hw, g).cuda()
maps = synth.make_feature_maps(B, hw, 256, g, channels_last=True)
split = fi.split_levels(fi.roi_level(rois, (hw[0], hw[1], 3)))
i = a.level - 2
idx = (split.big(i) if a.kind == "big" else split.small(i)).long()
boxes, ind = rois.view(-1, 4)[idx].contiguous(), (idx // R).int()
img = maps[i].cuda()
if a.fmt == "nchw":
    img = img.contiguous()
img.requires_grad_()
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for _ in range(a.iters):
    flush.add_(1.0)
    out = fi.crop_and_resize(img, boxes, ind, a.P, a.P)
    if a.dir in ("both", "bwd"):
        flush.add_(1.0)
        out.backward(torch.ones_like(out))
        img.grad = None
torch.cuda.synchronize()
print("done", tuple(out.shape))
