"""Generate tests/golden/*.npz from the REFERENCE ITSELF.  Run in the build container only
(needs /root/reference and oracle/_ref):

    python tests/golden/make_golden.py

* roi_align.npz   outputs of lib/roi_align/src/crop_and_resize.c compiled unmodified (oracle/_ref)
* nms.npz         outputs of lib/nms/src/nms.c compiled unmodified (oracle/_ref)
* sinkhorn.npz    outputs of lib/OT_module.py::OptTrans._sinkhorn_iterate, module imported as-is
* opttrans.npz    outputs of lib/OT_module.py::OptTrans.forward (1-D and 2-D), weights included
* targets.npz     mask targets (lib/layers.py:297-323, inline code of generate_roi, with CropAndResizeFunction bound to the compiled
                  crop_and_resize.c) and detection_layer + conduct_nms (lib/layers.py:664-802, with `nms` bound to the compiled
                  cpu_nms), same ast / PyTorch-0.3-shim mechanism as intertwiner.npz
* intertwiner.npz outputs of the reference's own Python bodies on the path, cut out of the source with `ast` and
                  exec'd unmodified under a PyTorch-0.3 behaviour shim (tests/golden/ref_exec.py): the level rule
                  (lib/sub_module.py:397-410 and its twin lib/layers.py:168-181), tools/utils.py unique1d / log2,
                  Dev._assign_feat2cls (lib/sub_module.py:664-684), MaskRCNN._merge_feat_vec / meta_loss /
                  _assign_from_buffer (lib/model.py:143-224; buffer sizes 1 and 3, l2 / l1 / ot, class- and
                  instance-level, three consecutive iterations), tools/box_utils.py apply_box_deltas / clip_boxes

The only liberty taken with lib/OT_module.py is neutralising ``Tensor.cuda`` (hard-coded at
OT_module.py:118-119) because this container has no GPU; inputs are cloned before the call because the
reference normalises them in place (OT_module.py:111-112).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import clib  # noqa: E402

REF = "/root/reference"


def roi_align_cases():
    rng = np.random.default_rng(2000)
    B, C, H, W = 2, 4, 12, 10
    image = rng.standard_normal((B, C, H, W)).astype(np.float32)
    R = 24
    ctr = rng.uniform(0.1, 0.9, (R, 2))
    size = rng.uniform(0.05, 0.7, (R, 2))
    boxes = np.concatenate([ctr - size / 2, ctr + size / 2], 1).astype(np.float32)
    boxes[0] = [0, 0, 1, 1]                 # whole image
    boxes[1] = [0, 0, 0, 0]                 # zero-padded RoI
    boxes[2] = [-0.2, 0.1, 0.5, 1.3]        # straddles the border -> extrapolation
    boxes[3] = [0.25, 0.25, 0.25, 0.25]     # degenerate point
    boxes[4] = [3 / 11, 2 / 9, 8 / 11, 7 / 9]   # lands on integer pixels
    box_ind = rng.integers(0, B, R).astype(np.int32)
    out = {"image": image, "boxes": boxes, "box_ind": box_ind}
    for P in (1, 2, 7, 14):
        crops = clib.ref_crop_and_resize_fwd(image, boxes, box_ind, P, P, 0.25)
        grads = rng.standard_normal(crops.shape).astype(np.float32)
        out[f"crops_{P}"] = crops
        out[f"grads_{P}"] = grads
        out[f"grad_image_{P}"] = clib.ref_crop_and_resize_bwd(grads, boxes, box_ind, image.shape)
    crops = clib.ref_crop_and_resize_fwd(image, boxes, box_ind, 3, 5, 0.0)   # non-square crop
    out["crops_3x5"] = crops
    np.savez_compressed(os.path.join(HERE, "roi_align.npz"), **out)


def nms_cases():
    rng = np.random.default_rng(2001)
    n = 200
    ctr = rng.uniform(20, 200, (n, 2))
    size = rng.uniform(10, 80, (n, 2))
    d = np.concatenate([ctr - size / 2, ctr + size / 2, rng.uniform(0, 1, (n, 1))], 1).astype(np.float32)
    d = d[np.argsort(-d[:, 4], kind="stable")]          # callers pre-sort (layers.py:103)
    out = {"dets": d}
    for thr in (0.3, 0.5, 0.7):
        out[f"keep_cpu_{thr}"] = clib.ref_cpu_nms(d, thr)
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **out)


def load_reference_ot():
    torch.Tensor.cuda = lambda self, *a, **k: self        # no GPU here (OT_module.py:118-119)
    spec = importlib.util.spec_from_file_location("ref_OT_module", os.path.join(REF, "lib", "OT_module.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sinkhorn_cases(mod):
    cfg = types.SimpleNamespace(DEV=types.SimpleNamespace(OT_ONE_DIM_FORM="conv"))
    g = torch.Generator().manual_seed(2000)
    out = {}
    for name, (N, D) in {"n256_d1": (256, 1), "n64_d256": (64, 256), "n16_d3": (16, 3)}.items():
        x = torch.randn(N, D, generator=g).abs()      # the critic ends in ReLU: inputs are >= 0
        y = torch.randn(N, D, generator=g).abs()
        x[::7] = 0                                    # ReLU zeros -> zero rows
        out[f"{name}_x"], out[f"{name}_y"] = x.numpy().copy(), y.numpy().copy()
        for L in (1, 5, 50):
            for eps in (1.0, 0.1):
                m = mod.OptTrans(cfg, ch_x=16, epsilon=eps, L=L)
                with torch.no_grad():
                    out[f"{name}_L{L}_eps{eps}"] = np.float32(m._sinkhorn_iterate(x.clone(), y.clone()).item())
    np.savez_compressed(os.path.join(HERE, "sinkhorn.npz"), **out)


def opttrans_cases(mod):
    cfg = types.SimpleNamespace(DEV=types.SimpleNamespace(OT_ONE_DIM_FORM="conv"))
    out = {}
    torch.manual_seed(2000)
    m = mod.OptTrans(cfg, ch_x=64, L=5).eval()
    x, y = torch.randn(6, 64, 1), torch.randn(6, 64, 1).abs()
    with torch.no_grad():
        out["d1_loss"] = m(x.clone(), y.clone()).numpy()
    out["d1_x"], out["d1_y"] = x.numpy(), y.numpy()
    for k, v in m.state_dict().items():
        out["d1_sd_" + k] = v.numpy()
    m2 = mod.OptTrans(cfg, ch_x=16, spatial_x=8, spatial_y=16, L=5).eval()
    x2, y2 = torch.randn(3, 16, 8, 8), torch.randn(3, 16, 16, 16)
    with torch.no_grad():
        out["d2_loss"] = m2(x2.clone(), y2.clone()).numpy()
    out["d2_x"], out["d2_y"] = x2.numpy(), y2.numpy()
    for k, v in m2.state_dict().items():
        out["d2_sd_" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "opttrans.npz"), **out)


def intertwiner_cases(ot_mod):
    """Everything below runs the reference's own source text (see ref_exec.py); only the inputs are ours."""
    sys.path.insert(0, HERE)
    import ref_exec as rx
    out, cited = {}, {}
    g = torch.Generator().manual_seed(2000)
    ns = rx.load_functions([("tools/utils.py", "unique1d", None), ("tools/utils.py", "log2", None),
                            ("lib/sub_module.py", "_assign_feat2cls", "Dev"),
                            ("lib/model.py", "_merge_feat_vec", "MaskRCNN"), ("lib/model.py", "_assign_from_buffer", "MaskRCNN"),
                            ("lib/model.py", "meta_loss", "MaskRCNN"),
                            ("tools/box_utils.py", "apply_box_deltas", None), ("tools/box_utils.py", "clip_boxes", None)])
    cited.update(ns["__cited__"])

    # ---- level rule: inline code of Dev.forward (lib/sub_module.py:396-410) and of pyramid_roi_align (lib/layers.py:167-181)
    from feature_intertwiner_b200 import synth
    level_src = {"dev": ("lib/sub_module.py", 397, 410), "pyramid": ("lib/layers.py", 168, 181)}
    for tag, hw, bs, R in (("c1", (256, 256), 4, 64), ("c2", (832, 1344), 8, 512), ("sq", (1024, 1024), 2, 300)):
        rois = synth.make_rois(bs, R, hw, g)
        # level boundaries: squares whose side is 224 * 2^(k - 4 +- 1/2) * (1 +- tiny) pixels, k = 2..5
        edge = []
        for k in (2.5, 3.5, 4.5):
            for rel in (-3e-7, -1e-7, 0.0, 1e-7, 3e-7):
                s = 224.0 * 2.0 ** (k - 4) * (1 + rel)
                edge.append([0.1, 0.1, 0.1 + s / hw[0], 0.1 + s / hw[1]])
        rois[0, :len(edge)] = torch.tensor(edge)
        out["level_%s_rois" % tag] = rois.numpy().copy()
        out["level_%s_image_shape" % tag] = np.array([hw[0], hw[1], 3])
        for which, (rel, a, b) in level_src.items():
            with rx.torch03():
                cfgobj = types.SimpleNamespace(DEV=types.SimpleNamespace(ASSIGN_BOX_ON_ALL_SCALE=False))
                self_ = types.SimpleNamespace(config=cfgobj, image_shape=np.array([hw[0], hw[1], 3]))
                loc = dict(ns, rois=rois.clone(), boxes=rois.clone(), self=self_, base=224.0, image_shape=np.array([hw[0], hw[1], 3]),
                           inputs=[rois.clone()])
                exec(compile(rx.extract_lines(rel, a, b), rel, "exec"), loc)
                out["level_%s_%s" % (tag, which)] = loc["roi_level"].numpy().astype(np.int32)
            cited["level_rule_" + which] = "%s:%d-%d" % (rel, a, b)
        assert np.array_equal(out["level_%s_dev" % tag], out["level_%s_pyramid" % tag])

    # ---- unique1d / log2
    with rx.torch03():
        v = torch.randint(0, 81, (300,), generator=g)
        out["unique_in"], out["unique_out"] = v.numpy(), ns["unique1d"](v.clone()).numpy()
        x = torch.rand(1000, generator=g) * 8 + 1e-3
        out["log2_in"], out["log2_out"] = x.numpy(), ns["log2"](x.clone()).numpy()

    # ---- _assign_feat2cls (feat given as [k,1024]: the [k,1024,1,1] form of the caller cannot be slice-assigned on torch 2.x)
    self_ = types.SimpleNamespace(num_classs=81)
    for tag, k in (("a", 1), ("b", 37), ("c", 150)):
        gt = torch.randint(0, 81, (k,), generator=g)
        gt[: k // 3] = 0
        feat = torch.rand(k, 1024, generator=g)
        with rx.torch03():
            f, c = ns["_assign_feat2cls"](self_, [gt.clone(), feat.clone()])
        out["seg_%s_gt" % tag], out["seg_%s_feat" % tag] = gt.numpy().astype(np.int32), feat.numpy()
        out["seg_%s_mean" % tag], out["seg_%s_cnt" % tag] = f.numpy(), c.numpy()

    # ---- _merge_feat_vec + meta_loss over three iterations
    Fd = 64                                   # meta_loss takes the feature width from the buffer; 64 keeps the fixture small
    torch.manual_seed(2000)
    cfg_ot = types.SimpleNamespace(DEV=types.SimpleNamespace(OT_ONE_DIM_FORM="conv"))
    ot = ot_mod.OptTrans(cfg_ot, ch_x=Fd, L=5).eval()
    for k, v in ot.state_dict().items():
        out["meta_ot_sd_" + k] = v.numpy()

    def stats(G, S, density):
        cnt = (torch.rand(G, S, 1, 81, generator=g) < density).float() * torch.randint(1, 9, (G, S, 1, 81), generator=g).float()
        feat = torch.rand(G, S, Fd, 81, generator=g) * (cnt > 0).float()
        return feat, cnt

    f, c = stats(2, 3, 0.5)
    with rx.torch03():
        mf, mc = ns["_merge_feat_vec"](f.clone(), c.clone())
    out["merge_feat"], out["merge_cnt"], out["merge_out_feat"], out["merge_out_cnt"] = f.numpy(), c.numpy(), mf.numpy(), mc.numpy()

    for B in (1, 3):
        for lc in ("l2", "l1", "ot"):
            for inst in (False, True):
                if inst and lc == "ot":
                    continue            # 3 x n Sinkhorn problems per instance: covered at class level
                if B > 1 and not inst:
                    continue            # the reference's class-level match is shape-invalid for BUFFER_SIZE > 1 (lib/model.py:180,
                                        # `buffer_cnt.squeeze()` is [B,81]): it raises, there is nothing to pin
                tag = "meta_B%d_%s_%s" % (B, lc, "inst" if inst else "cls")
                model = types.SimpleNamespace(
                    config=types.SimpleNamespace(DEV=types.SimpleNamespace(INST_LOSS=inst, LOSS_CHOICE=lc)),
                    buffer=torch.zeros(B, Fd, 81), buffer_cnt=torch.zeros(B, 1, 81), ot_loss=ot)
                model._merge_feat_vec = ns["_merge_feat_vec"]
                model._assign_from_buffer = ns["_assign_from_buffer"]
                for it in range(3):
                    bf, bc = stats(1, 3, 0.3 if it == 0 else 0.6)
                    sf, sc = stats(1, 3, 0.4)
                    n_inst = 40
                    so = torch.rand(n_inst, Fd, generator=g)
                    sg = torch.randint(0, 81, (n_inst,), generator=g)
                    sg[::3] = 0
                    with rx.torch03(), torch.no_grad():
                        loss = ns["meta_loss"](model, [bf.clone(), bc.clone(), sf.clone(), sc.clone(), so.clone(), sg.clone()])
                    pre = "%s_it%d_" % (tag, it)
                    for name, t in (("big_feat", bf), ("big_cnt", bc), ("small_feat", sf), ("small_cnt", sc), ("small_out", so), ("small_gt", sg)):
                        out[pre + name] = t.numpy()
                    out[pre + "loss"] = loss.detach().numpy().reshape(-1)
                    out[pre + "buffer"], out[pre + "buffer_cnt"] = model.buffer.numpy().copy(), model.buffer_cnt.numpy().copy()
    # empty comparison set -> zeros(1) (lib/model.py:208-209) is a known answer; torch 0.3's "empty tensor has no size" has no 2.x analogue

    # ---- apply_box_deltas / clip_boxes (tools/box_utils.py, imported functions exec'd as they are)
    priors = torch.rand(400, 4, generator=g) * 600
    priors[:, 2:] = priors[:, :2] + torch.rand(400, 2, generator=g) * 300 + 1
    anchors = priors.unsqueeze(0).expand(2, 400, 4).contiguous()            # lib/layers.py:96
    deltas = torch.randn(2, 400, 4, generator=g) * 0.3
    window = torch.tensor([0.0, 0.0, 832.0, 1344.0])
    with rx.torch03():
        boxes = ns["apply_box_deltas"](anchors.clone(), deltas.clone())
        clipped = ns["clip_boxes"](boxes.clone(), window.clone())
    out["box_anchors"], out["box_deltas"], out["box_window"] = anchors.numpy(), deltas.numpy(), window.numpy()
    out["box_applied"], out["box_clipped"] = boxes.numpy(), clipped.numpy()
    out["cited"] = np.array(sorted("%s=%s" % kv for kv in cited.items()))
    np.savez_compressed(os.path.join(HERE, "intertwiner.npz"), **out)


def targets_cases():
    """Mask targets and the detection layer from the reference's own source text (ref_exec.py)."""
    sys.path.insert(0, HERE)
    import ref_exec as rx
    out, cited = {}, {}
    g = torch.Generator().manual_seed(2002)

    class CropAndResizeFunction(object):                     # lib/roi_align/crop_and_resize.py:14-37 over the compiled reference C
        def __init__(self, h, w, extrapolation_value=0):
            self.h, self.w, self.e = h, w, extrapolation_value

        def __call__(self, image, boxes, box_ind):
            o = clib.ref_crop_and_resize_fwd(image.numpy(), boxes.numpy(), box_ind.numpy(), self.h, self.w, float(self.e))
            return torch.from_numpy(o)

    def nms(dets, thresh):                                   # lib/nms/nms_wrapper.py:14-34 over the compiled cpu_nms (lib/nms/src/nms.c)
        keeps = [clib.ref_cpu_nms(dets[i].numpy(), thresh) for i in range(dets.size(0))]
        m = min(len(k) for k in keeps)
        return np.stack([k[:m] for k in keeps]).astype(np.int32)

    ns = rx.load_functions([("tools/utils.py", "unique1d", None), ("tools/utils.py", "intersect1d", None),
                            ("tools/box_utils.py", "apply_box_deltas", None), ("tools/box_utils.py", "clip_boxes", None),
                            ("lib/layers.py", "conduct_nms", None), ("lib/layers.py", "detection_layer", None)],
                           CropAndResizeFunction=CropAndResizeFunction, nms=nms)
    cited.update(ns["__cited__"])

    # ---- mask targets: lib/layers.py:297-323
    G, n = 9, 60
    for mini in (True, False):
        mh, mw = (56, 56) if mini else (120, 168)
        gt_masks = (torch.rand(G, mh, mw, generator=g) < 0.5).float()
        for k in range(G):                                   # blobs rather than noise: a disc per mask
            yy, xx = torch.meshgrid(torch.arange(mh).float(), torch.arange(mw).float(), indexing="ij")
            cy, cx, rad = mh * (0.3 + 0.4 * torch.rand(1, generator=g)), mw * (0.3 + 0.4 * torch.rand(1, generator=g)), min(mh, mw) * (0.15 + 0.2 * torch.rand(1, generator=g))
            gt_masks[k] = (((yy - cy) ** 2 + (xx - cx) ** 2) < rad ** 2).float()
        gt_boxes = torch.rand(G, 4, generator=g) * 0.5
        gt_boxes[:, 2:] = gt_boxes[:, :2] + 0.1 + torch.rand(G, 2, generator=g) * 0.4
        assign = torch.randint(0, G, (n,), generator=g)
        roi_gt = gt_boxes[assign]
        POS_ROIS = (roi_gt + (torch.rand(n, 4, generator=g) - 0.5) * 0.12).clamp(0, 1)
        cfgobj = types.SimpleNamespace(MRCNN=types.SimpleNamespace(USE_MINI_MASK=mini, MASK_SHAPE=[28, 28]))
        loc = dict(ns, gt_masks=gt_masks.clone(), roi_gt_box_assignment=assign.clone(), POS_ROIS=POS_ROIS.clone(), roi_gt_boxes=roi_gt.clone(), config=cfgobj)
        with rx.torch03():
            exec(compile(rx.extract_lines("lib/layers.py", 297, 323), "lib/layers.py", "exec"), loc)
        tag = "mask_mini" if mini else "mask_full"
        out[tag + "_gt_masks"], out[tag + "_gt_boxes"], out[tag + "_assign"] = gt_masks.numpy(), gt_boxes.numpy(), assign.numpy().astype(np.int32)
        out[tag + "_pos_rois"], out[tag + "_targets"] = POS_ROIS.numpy(), loc["MASKS"].numpy()
    cited["mask_targets"] = "lib/layers.py:297-323"

    # ---- detection_layer + conduct_nms: lib/layers.py:664-802
    bs, R, ncls, hw = 2, 300, 81, (832, 1344)
    cfgobj = types.SimpleNamespace(TEST=types.SimpleNamespace(DET_MAX_INSTANCES=100, DET_MIN_CONFIDENCE=0.3, DET_NMS_THRESHOLD=0.3),
                                   DATA=types.SimpleNamespace(BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2]), IMAGE_SHAPE=np.array([hw[0], hw[1], 3])),
                                   MISC=types.SimpleNamespace(GPU_COUNT=0))
    ctr = torch.rand(bs, R, 2, generator=g) * 0.8 + 0.1
    size = torch.rand(bs, R, 2, generator=g) * 0.25 + 0.03
    rois = torch.cat([ctr - size / 2, ctr + size / 2], dim=2).clamp(0, 1)
    rois[:, 200:] = rois[:, :100] + (torch.rand(bs, 100, 4, generator=g) - 0.5) * 0.02      # near-duplicates: NMS has work to do
    logits = torch.randn(bs * R, ncls, generator=g) * 2.0
    logits[:, 0] += 1.0
    cls = torch.randint(1, 12, (bs * R,), generator=g)                                      # few classes: several boxes per class
    logits[torch.arange(bs * R), cls] += 6.0 * torch.rand(bs * R, generator=g)
    probs = torch.softmax(logits, dim=1)
    deltas = torch.randn(bs * R, ncls, 4, generator=g) * 0.5
    windows = torch.tensor([[0.0, 0.0, 832.0, 1344.0], [16.0, 40.0, 800.0, 1300.0]])
    with rx.torch03(), torch.no_grad():
        det, _ = ns["detection_layer"](rois.clone(), probs.clone(), deltas.clone(), windows.clone(), cfgobj, feature=torch.zeros(bs * R, 4))
    out["det_rois"], out["det_probs"], out["det_deltas"], out["det_windows"] = rois.numpy(), probs.numpy(), deltas.numpy(), windows.numpy()
    out["det_out"] = det.numpy()
    cfgobj.TEST.DET_MIN_CONFIDENCE = 0.93                        # fewer candidates than DET_MAX_INSTANCES: zero rows after the detections
    with rx.torch03(), torch.no_grad():
        det2, _ = ns["detection_layer"](rois.clone(), probs.clone(), deltas.clone(), windows.clone(), cfgobj, feature=torch.zeros(bs * R, 4))
    out["det_out_conf093"] = det2.numpy()
    out["cited"] = np.array(sorted("%s=%s" % kv for kv in cited.items()))
    np.savez_compressed(os.path.join(HERE, "targets.npz"), **out)


if __name__ == "__main__":
    assert os.path.isdir(REF), "run in the build container (needs /root/reference)"
    clib.build(ref=True)
    roi_align_cases()
    nms_cases()
    ot = load_reference_ot()
    sinkhorn_cases(ot)
    opttrans_cases(ot)
    intertwiner_cases(ot)
    targets_cases()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
