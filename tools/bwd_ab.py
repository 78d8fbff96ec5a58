"""A/B harness for the RoIAlign backward at a BASELINE.json workload: every crop set of one Dev.forward pass through
fi.crop_sets, backward timed with CUDA events (L2 flushed between iterations), plus an order-independent fingerprint of
the bits of every dense gradient map so that two kernel variants (separate processes: the variant switches are read from
the environment once per process) can be compared bit for bit.

    python tools/bwd_ab.py --workload c2 [--exact] [--iters 20] --out gpurun_out/ab_x.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import feature_intertwiner_b200 as fi  # noqa: E402
from feature_intertwiner_b200 import synth  # noqa: E402


def build_specs(wl, dev, seed=2000, spatial=True):
    g = torch.Generator().manual_seed(seed)
    B, R, hw = wl["batch"], wl["rois_per_image"], wl["image"]
    rois = synth.make_rois(B, R, hw, g).to(dev)
    gt = synth.make_class_ids(B, R, g, 81).to(dev)
    raw = [m.to(dev).requires_grad_() for m in synth.make_feature_maps(B, hw, 256, g, channels_last=True)]
    madeup = [m.to(dev).requires_grad_() for m in synth.make_feature_maps(B, hw, 256, g, channels_last=True)]
    split = fi.split_levels(fi.roi_level(rois, (hw[0], hw[1], 3)), rois=rois, gt=gt, order=fi.spatial_order(rois) if spatial else None)
    total = B * R
    cl = torch.channels_last
    pooled = torch.empty((total, 256, 7, 7), device=dev, memory_format=cl)
    mask = torch.empty((total, 256, 14, 14), device=dev, memory_format=cl)
    specs = []
    for i in range(4):
        if split.small_cnt[i] == 0:
            continue
        if i < 3 and split.big_cnt[i]:
            specs.append(dict(image=raw[i], boxes=split.big_boxes(i), box_ind=split.big_ind(i), size=14))
        specs.append(dict(image=madeup[i], boxes=split.small_boxes(i), box_ind=split.small_ind(i), size=7, out=pooled, dst_row=split.small(i)))
        specs.append(dict(image=madeup[i], boxes=split.small_boxes(i), box_ind=split.small_ind(i), size=14, out=mask, dst_row=split.small(i),
                          compact=(i < 3)))
    return raw, madeup, specs, split


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--exact", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--plan-in-backward", action="store_true", help="build the sample lists inside backward (timed) instead of at forward time")
    args = ap.parse_args()
    if args.plan_in_backward:
        fi.roi_align.plan_at_forward(False)
    dev = torch.device("cuda", 0)
    wl = synth.WORKLOADS[args.workload]
    raw, madeup, specs, split = build_specs(wl, dev)
    if args.exact:
        fi.set_deterministic(True)
    outs, comps = fi.crop_sets(specs)
    heads = []
    seen = set()
    for o in outs:
        if o is not None and id(o) not in seen:
            seen.add(id(o)); heads.append(o)
    heads += [c for c in comps if c is not None]
    g = torch.Generator(device=dev).manual_seed(7)
    grads = [torch.randn(h.shape, device=dev, generator=g).contiguous(memory_format=torch.channels_last) for h in heads]
    leaves = raw + madeup
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    times = []
    res = None
    for it in range(args.iters + 3):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = torch.autograd.grad(heads, leaves, grads, retain_graph=True, allow_unused=True)
        b.record()
        torch.cuda.synchronize()
        if it >= 3:
            times.append(a.elapsed_time(b))
    times.sort()
    prints = []
    for r in res:
        if r is None:
            prints.append(None)
            continue
        bits = r.contiguous(memory_format=torch.channels_last).view(torch.int32).long()
        flat = bits.permute(0, 2, 3, 1).reshape(-1)
        idx = torch.arange(flat.numel(), device=dev, dtype=torch.long)
        prints.append([int(flat.sum().item()), int(((flat * ((idx % 65521) + 1)) % 2147483647).sum().item()), float(r.abs().sum().item())])
    out = dict(workload=args.workload, exact=args.exact, env={k: v for k, v in os.environ.items() if k.startswith("FI_")},
               plan_in_backward=args.plan_in_backward, bwd_ms_median=times[len(times) // 2], bwd_ms_min=times[0], bwd_ms_all=[round(t, 4) for t in times],
               fingerprints=prints, small=split.small_cnt, big=split.big_cnt)
    s = json.dumps(out)
    print(s)
    if args.out:
        open(args.out, "w").write(s)


if __name__ == "__main__":
    main()
