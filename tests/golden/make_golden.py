"""Generate tests/golden/*.npz from the REFERENCE ITSELF.  Run in the build container only
(needs /root/reference and oracle/_ref):

    python tests/golden/make_golden.py

* roi_align.npz   outputs of lib/roi_align/src/crop_and_resize.c compiled unmodified (oracle/_ref)
* nms.npz         outputs of lib/nms/src/nms.c compiled unmodified (oracle/_ref)
* sinkhorn.npz    outputs of lib/OT_module.py::OptTrans._sinkhorn_iterate, module imported as-is
* opttrans.npz    outputs of lib/OT_module.py::OptTrans.forward (1-D and 2-D), weights included

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


if __name__ == "__main__":
    assert os.path.isdir(REF), "run in the build container (needs /root/reference)"
    clib.build(ref=True)
    roi_align_cases()
    nms_cases()
    ot = load_reference_ot()
    sinkhorn_cases(ot)
    opttrans_cases(ot)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
