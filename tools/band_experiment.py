"""Does the reduction backward get faster when the dense gradient map fits in L2?  Same boxes, same kernel, the map of
ONE image (71 MB at P2) at a time vs all 8 images (572 MB) at once.  (Experiment behind DESIGN.md section 8 item 1.)"""
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
split = fi.split_levels(fi.roi_level(rois, (hw[0], hw[1], 3)))
flat = rois.view(-1, 4)
H, W = synth.level_shapes(hw)[0]
flush = torch.empty(128 * 1024 * 1024, device="cuda")
s = torch.cuda.current_stream().cuda_stream
L = _lib.lib()


def timed(fn, n=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for kind, P in (("big", 14), ("small", 14), ("small", 7)):
    idx = (split.big(0) if kind == "big" else split.small(0)).long()
    boxes, ind = flat[idx].contiguous(), (idx // R).int().contiguous()
    n = boxes.size(0)
    grads = torch.randn(n, 256, P, P, device="cuda").contiguous(memory_format=torch.channels_last)
    gimg = torch.empty(B, 256, H, W, device="cuda").contiguous(memory_format=torch.channels_last)

    def whole():
        _lib.check(L.fi_crop_and_resize_backward(grads.data_ptr(), 1, boxes.data_ptr(), ind.data_ptr(), None, n, B, H, W, P, P, 256, gimg.data_ptr(), 1, 0, s))
    # per image: boxes are grouped by image (nonzero order)
    counts = torch.bincount(ind.long(), minlength=B).tolist()
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    zero_ind = torch.zeros(n, dtype=torch.int32, device="cuda")

    def banded():
        for b in range(B):
            lo, hi = offs[b], offs[b + 1]
            _lib.check(L.fi_crop_and_resize_backward(grads[lo:hi].data_ptr(), 1, boxes[lo:hi].data_ptr(), zero_ind[lo:hi].data_ptr(), None, hi - lo, 1, H, W, P, P,
                                                     256, gimg[b:b + 1].data_ptr(), 1, 0, s))
    t_w, t_b = timed(whole), timed(banded)
    alg = 4 * 256 * n * P * P + 4 * 256 * B * H * W
    print("P2 %s %dx%d boxes=%d: whole map %.3f ms (%.0f GB/s)  |  image by image %.3f ms (%.0f GB/s, 16 launches)"
          % (kind, P, P, n, t_w, alg / t_w / 1e6, t_b, alg / t_b / 1e6), flush=True)
