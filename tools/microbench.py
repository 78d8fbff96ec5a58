"""Per-kernel timings on the B200 box (CUDA events, L2 flushed between iterations): every RoIAlign call of
Dev.forward at a BASELINE.json workload, forward and backward, in both memory formats, next to the
reference's own CUDA kernels recompiled for sm_100a (oracle/_ref/libref_cuda.so, context row only) and a plain
copy of the same bytes; the Sinkhorn kernel at the class- and instance-level problem counts.

    python tools/microbench.py --workload c2 --out gpurun_out/microbench_c2.json
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import feature_intertwiner_b200 as fi  # noqa: E402
from feature_intertwiner_b200 import _lib, synth  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(512 * 1024 * 1024 // 4, device="cuda")     # 512 MB >> 126 MB L2
    _flush.add_(1.0)


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def ref_cuda():
    p = os.path.join(ROOT, "oracle", "_ref", "libref_cuda.so")
    if not os.path.exists(p):
        return None
    L = ctypes.CDLL(p)
    P, I, Fl = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    L.CropAndResizeLaucher.argtypes = [P, P, P, I, I, I, I, I, I, I, Fl, P, P]
    L.CropAndResizeLaucher.restype = None
    L.CropAndResizeBackpropImageLaucher.argtypes = [P, P, P, I, I, I, I, I, I, I, P, P]
    L.CropAndResizeBackpropImageLaucher.restype = None
    L.ROIPoolForwardLaucher.argtypes = [P, Fl, I, I, I, I, I, I, P, P, P, P]           # roi_pooling_kernel.h:8-12
    L.ROIPoolForwardLaucher.restype = I
    L.ROIPoolBackwardLaucher.argtypes = [P, Fl, I, I, I, I, I, I, I, P, P, P, P]       # roi_pooling_kernel.h:14-18
    L.ROIPoolBackwardLaucher.restype = I
    return L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "microbench.json"))
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    wl = synth.WORKLOADS[args.workload]
    peak, peak_kind = peaks()
    g = torch.Generator().manual_seed(2000)
    B, R, hw = wl["batch"], wl["rois_per_image"], wl["image"]
    rois = synth.make_rois(B, R, hw, g).cuda()
    maps = [m.cuda() for m in synth.make_feature_maps(B, hw, 256, g, channels_last=True)]
    level = fi.roi_level(rois, (hw[0], hw[1], 3))
    split = fi.split_levels(level)
    flat = rois.view(-1, 4)
    refL = ref_cuda()
    from oracle import clib
    rows = []
    s = torch.cuda.current_stream().cuda_stream
    for i in range(4):
        calls = []
        if split.small_cnt[i]:
            calls += [("small", 7, split.small(i)), ("small", 14, split.small(i))]
        if i < 3 and split.big_cnt[i]:
            calls += [("big", 14, split.big(i))]
        for kind, P, idx in calls:
            idx = idx.long()
            boxes = flat[idx].contiguous()
            ind = (idx // R).int().contiguous()
            n = boxes.size(0)
            Hh, Ww = maps[i].shape[2:]
            U = clib.oracle_unique_pixels(B, Hh, Ww, boxes.cpu().numpy(), ind.cpu().numpy(), P, P)
            fwd_bytes = 4 * 256 * n * P * P + 4 * 256 * U + 20 * n
            bwd_bytes = 4 * 256 * n * P * P + 4 * 256 * B * Hh * Ww + 20 * n
            row = dict(level=i + 2, kind=kind, P=P, boxes=n, map=[Hh, Ww], unique_px=U, fwd_alg_bytes=fwd_bytes, bwd_alg_bytes=bwd_bytes)
            for fmt in ("nhwc", "nchw"):
                img = maps[i] if fmt == "nhwc" else maps[i].contiguous()
                crops = fi.crop_and_resize(img, boxes, ind, P, P)
                t = timeit(lambda: fi.crop_and_resize(img, boxes, ind, P, P), args.iters)
                grads = torch.randn_like(crops)
                gimg = torch.empty_like(img)
                lay = _lib.FI_LAYOUT_NHWC if fmt == "nhwc" else _lib.FI_LAYOUT_NCHW

                def bwd():
                    _lib.check(_lib.lib().fi_crop_and_resize_backward(grads.data_ptr(), lay, boxes.data_ptr(), ind.data_ptr(), None, n, B, Hh, Ww,
                                                                      P, P, 256, gimg.data_ptr(), lay, 0, s))
                tb = timeit(bwd, args.iters)
                row[fmt] = dict(fwd_ms=t, bwd_ms=tb, fwd_gbs=fwd_bytes / t / 1e6, bwd_gbs=bwd_bytes / tb / 1e6,
                                fwd_frac=fwd_bytes / t / 1e6 / peak, bwd_frac=bwd_bytes / tb / 1e6 / peak)
                del crops, grads, gimg
            if refL is not None:       # the reference's kernels as they are (NCHW, memset + kernel as crop_and_resize_gpu.c does)
                img = maps[i].contiguous()
                crops = torch.empty(n, 256, P, P, device="cuda")

                def rf():
                    crops.zero_()
                    refL.CropAndResizeLaucher(img.data_ptr(), boxes.data_ptr(), ind.data_ptr(), n, B, Hh, Ww, P, P, 256, 0.0, crops.data_ptr(), s)
                t = timeit(rf, args.iters)
                grads = torch.randn_like(crops)
                gimg = torch.empty_like(img)

                def rb():
                    gimg.zero_()
                    refL.CropAndResizeBackpropImageLaucher(grads.data_ptr(), boxes.data_ptr(), ind.data_ptr(), n, B, Hh, Ww, P, P, 256, gimg.data_ptr(), s)
                tb = timeit(rb, args.iters)
                row["reference_cuda_sm100a"] = dict(fwd_ms=t, bwd_ms=tb, fwd_gbs=fwd_bytes / t / 1e6, bwd_gbs=bwd_bytes / tb / 1e6)
                del crops, grads, gimg, img
            print(json.dumps(row), flush=True)
            rows.append(row)
    # plain copy of 1 GB as the roofline sanity row
    a = torch.empty(256 * 1024 * 1024, device="cuda")
    b = torch.empty_like(a)
    t = timeit(lambda: b.copy_(a), args.iters)
    copy_row = dict(kind="copy_1GiB", ms=t, gbs=2 * a.numel() * 4 / t / 1e6)
    print(json.dumps(copy_row), flush=True)
    t = timeit(lambda: b.zero_(), args.iters)
    zero_row = dict(kind="memset_1GiB", ms=t, gbs=a.numel() * 4 / t / 1e6)
    print(json.dumps(zero_row), flush=True)
    del a, b
    # RoIPool (lib/roi_pooling): every RoI of the workload on the P3 map at 7x7, both layouts, next to the reference's kernel
    pool_rows = []
    g = torch.Generator().manual_seed(5)
    Hh, Ww = maps[1].shape[2], maps[1].shape[3]
    R = wl["batch"] * wl["rois_per_image"]
    ctr = torch.rand(R, 2, generator=g) * torch.tensor([Ww * 8.0, Hh * 8.0])
    half = 16 + 100 * torch.rand(R, 2, generator=g)
    rois5 = torch.cat([torch.randint(0, wl["batch"], (R, 1), generator=g).float(), (ctr - half).clamp(min=0), ctr + half], dim=1).cuda()
    for fmt in ("nhwc", "nchw"):
        img = maps[1] if fmt == "nhwc" else maps[1].contiguous()
        op = fi.RoIPoolFunction(7, 7, 1.0 / 8)
        x = img.detach().requires_grad_()
        out = op(x, rois5)
        gy = torch.randn_like(out)
        t = timeit(lambda: op(img, rois5), args.iters)
        tb = timeit(lambda: torch.autograd.grad(out, x, gy, retain_graph=True), args.iters)
        pool_rows.append(dict(kind="roi_pool", layout=fmt, rois=R, map=[Hh, Ww], fwd_ms=t, bwd_ms=tb))
        del out, gy, x
    if refL is not None:
        img = maps[1].contiguous()
        top = torch.empty(R, 256, 7, 7, device="cuda")
        arg = torch.empty(R, 256, 7, 7, device="cuda", dtype=torch.int32)
        t = timeit(lambda: refL.ROIPoolForwardLaucher(img.data_ptr(), 1.0 / 8, R, Hh, Ww, 256, 7, 7, rois5.data_ptr(), top.data_ptr(), arg.data_ptr(), s),
                   args.iters)
        gimg = torch.empty_like(img)
        gy = torch.randn_like(top)

        def rpb():
            gimg.zero_()                                                   # roi_pooling_cuda.c zero-fills before the atomics kernel
            refL.ROIPoolBackwardLaucher(gy.data_ptr(), 1.0 / 8, wl["batch"], R, Hh, Ww, 256, 7, 7, rois5.data_ptr(), gimg.data_ptr(), arg.data_ptr(), s)
        tb = timeit(rpb, args.iters)
        pool_rows.append(dict(kind="roi_pool", layout="reference_cuda_sm100a (nchw)", rois=R, map=[Hh, Ww], fwd_ms=t, bwd_ms=tb))
        del top, arg, gimg, gy
    for r in pool_rows:
        print(json.dumps(r), flush=True)
    sk = []
    for (P, N, D, L) in [(240, 256, 1, 50), (240, 256, 1, wl["sinkhorn_iters"]), (4050, 256, 1, 50), (12288, 256, 1, 50), (24, 64, 256, 5), (24, 64, 4096, 5), (240, 128, 1, 50), (240, 200, 1, 50)]:
        x = torch.randn(P, N, D, device="cuda").abs()
        y = torch.randn(P, N, D, device="cuda").abs()
        t = timeit(lambda: fi.sinkhorn_loss(x, y, 1.0, L), args.iters)
        flops = P * (2.0 * N * N * D + N * N + 4.0 * N * N * L + 3.0 * N * N)
        r = dict(kind="sinkhorn", problems=P, N=N, D=D, L=L, ms=t, tflops=flops / t / 1e9, us_per_problem=1e3 * t / P)
        xg, yg = x.clone().requires_grad_(), y.clone().requires_grad_()
        r["ms_with_grad"] = timeit(lambda: fi.sinkhorn_loss(xg, yg, 1.0, L), args.iters)
        print(json.dumps(r), flush=True)
        sk.append(r)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(dict(workload=args.workload, peak_gbs=peak, peak_kind=peak_kind, roi_align=rows, roi_pool=pool_rows, copy=copy_row, memset=zero_row, sinkhorn=sk),
              open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
