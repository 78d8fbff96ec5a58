"""The Python half of the hot path pinned to the REFERENCE'S OWN CODE.

tests/golden/intertwiner.npz holds outputs of the reference's own function bodies -- the level rule (lib/sub_module.py:397-410,
lib/layers.py:168-181), tools/utils.py unique1d / log2, Dev._assign_feat2cls (lib/sub_module.py:664-684),
MaskRCNN._merge_feat_vec / meta_loss / _assign_from_buffer (lib/model.py:143-224), tools/box_utils.py apply_box_deltas /
clip_boxes -- cut out of the source with `ast` and executed unmodified (tests/golden/make_golden.py + ref_exec.py; the
`cited` array in the fixture lists file:line of every body that ran).

* CPU tests: the oracle restatements (oracle/pyref.py, oracle/fi_oracle.c) reproduce those outputs.
* GPU tests (`-m gpu`): the CUDA path, through the C ABI, reproduces them.
"""
import types

import numpy as np
import pytest
import torch

from oracle import clib, pyref

LOSS_TOL = 1e-4   # north_star: "loss within 1e-4 fp32 of the reference"


@pytest.fixture(scope="module")
def Z(golden_dir):
    return np.load(golden_dir + "/intertwiner.npz")


def _tie(pre, tol=2e-6):
    """RoIs whose un-rounded level sits within rounding of x.5: log/sqrt of another libm may land on the other side."""
    return np.abs(pre - np.floor(pre) - 0.5) < tol


def _meta_cases(Z):
    tags = sorted({k[: k.index("_it")] for k in Z.files if k.startswith("meta_B")})
    assert len(tags) == 7, tags          # B1: l2/l1 x cls/inst + ot cls; B3: l2/l1 inst
    return tags


def _meta_cfg(tag):
    _, b, lc, kind = tag.split("_")
    return pyref.make_config(DEV__BUFFER_SIZE=int(b[1:]), DEV__LOSS_CHOICE=lc, DEV__INST_LOSS=(kind == "inst"))


def _ot_state(Z):
    return {k[len("meta_ot_sd_"):]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith("meta_ot_sd_")}


# ======================================================================================================= CPU: the oracle
def test_fixture_names_the_reference_code(Z):
    cited = set(Z["cited"].tolist())
    for need in ("level_rule_dev=lib/sub_module.py:397-410", "level_rule_pyramid=lib/layers.py:168-181", "unique1d=tools/utils.py:30-41",
                 "log2=tools/utils.py:50-55", "_assign_feat2cls=lib/sub_module.py:664-684", "_merge_feat_vec=lib/model.py:218-224",
                 "meta_loss=lib/model.py:143-210", "apply_box_deltas=tools/box_utils.py:7-29", "clip_boxes=tools/box_utils.py:32-60"):
        assert need in cited, need


def test_level_rule_oracle_vs_reference(Z):
    for tag in ("c1", "c2", "sq"):
        rois = torch.from_numpy(Z["level_%s_rois" % tag])
        shape = tuple(int(v) for v in Z["level_%s_image_shape" % tag])
        want = Z["level_%s_dev" % tag]
        np.testing.assert_array_equal(want, Z["level_%s_pyramid" % tag])
        got, pre = pyref.roi_level_ref(rois, shape)                        # same torch ops on the same CPU: identical
        np.testing.assert_array_equal(got.numpy(), want)
        got_c, pre_c = clib.oracle_roi_level(rois.numpy().reshape(-1, 4), float(shape[0] * shape[1]), 224.0)
        bad = got_c != want.reshape(-1)
        assert not np.any(bad & ~_tie(pre_c)), "C restatement differs away from a rounding tie"
        assert set(np.unique(want)) == {2, 3, 4, 5}


def test_unique_and_log2_oracle_vs_reference(Z):
    np.testing.assert_array_equal(torch.unique(torch.from_numpy(Z["unique_in"])).numpy(), Z["unique_out"])   # what assign_feat2cls_ref uses
    x = torch.from_numpy(Z["log2_in"])
    np.testing.assert_array_equal((torch.log(x) / torch.log(torch.tensor([2.0]))).numpy(), Z["log2_out"])      # what roi_level_ref uses


def test_segment_mean_oracle_vs_reference(Z):
    for tag in "abc":
        gt, feat = Z["seg_%s_gt" % tag], Z["seg_%s_feat" % tag]
        m, c = pyref.assign_feat2cls_ref(torch.from_numpy(gt).long(), torch.from_numpy(feat), 81)
        np.testing.assert_array_equal(c.numpy(), Z["seg_%s_cnt" % tag])
        np.testing.assert_allclose(m.numpy(), Z["seg_%s_mean" % tag], rtol=1e-6, atol=1e-7)
        mc, cc = clib.oracle_segment_mean(gt, feat, 81)
        np.testing.assert_array_equal(cc, Z["seg_%s_cnt" % tag])
        np.testing.assert_allclose(mc, Z["seg_%s_mean" % tag], rtol=1e-6, atol=1e-7)


def test_merge_feat_vec_oracle_vs_reference(Z):
    f, n = pyref.merge_feat_vec_ref(torch.from_numpy(Z["merge_feat"]), torch.from_numpy(Z["merge_cnt"]))
    np.testing.assert_array_equal(n.numpy(), Z["merge_out_cnt"])
    np.testing.assert_allclose(f.numpy(), Z["merge_out_feat"], rtol=1e-6, atol=1e-7)


def test_meta_loss_oracle_vs_reference(Z):
    """Three consecutive iterations per configuration: the buffer is state (lib/model.py:148-166)."""
    for tag in _meta_cases(Z):
        cfg = _meta_cfg(tag)
        ot = None
        if cfg.DEV.LOSS_CHOICE == "ot":
            ot = pyref.OptTransRef(ch_x=64, L=5)
            ot.load_state_dict(_ot_state(Z))
        ref = pyref.MetaLossRef(cfg, 64, ot_loss=ot)
        for it in range(3):
            pre = "%s_it%d_" % (tag, it)
            inp = [torch.from_numpy(Z[pre + k]) for k in ("big_feat", "big_cnt", "small_feat", "small_cnt", "small_out", "small_gt")]
            with torch.no_grad():
                loss = ref(inp)
            np.testing.assert_allclose(loss.numpy().reshape(-1), Z[pre + "loss"], atol=2e-6, rtol=1e-5, err_msg=pre)
            np.testing.assert_allclose(ref.buffer.numpy(), Z[pre + "buffer"], rtol=1e-6, atol=1e-7, err_msg=pre)
            np.testing.assert_array_equal(ref.buffer_cnt.numpy(), Z[pre + "buffer_cnt"])


def test_box_deltas_oracle_vs_reference(Z):
    got = pyref.apply_box_deltas_ref(torch.from_numpy(Z["box_anchors"]), torch.from_numpy(Z["box_deltas"]))
    np.testing.assert_array_equal(got.numpy(), Z["box_applied"])              # same op order, same CPU: identical bits
    H, W = float(Z["box_window"][2]), float(Z["box_window"][3])
    clip = torch.stack([got[:, :, 0].clamp(0.0, H), got[:, :, 1].clamp(0.0, W), got[:, :, 2].clamp(0.0, H), got[:, :, 3].clamp(0.0, W)], 2)
    np.testing.assert_array_equal(clip.numpy(), Z["box_clipped"])


# ======================================================================================================= GPU: the CUDA path
@pytest.mark.gpu
def test_level_rule_cuda_vs_reference(Z):
    import feature_intertwiner_b200 as fi
    for tag in ("c1", "c2", "sq"):
        rois = torch.from_numpy(Z["level_%s_rois" % tag])
        shape = tuple(int(v) for v in Z["level_%s_image_shape" % tag])
        want = Z["level_%s_dev" % tag]
        got = fi.roi_level(rois.cuda(), shape, 224.0).cpu().numpy()
        _, pre = pyref.roi_level_ref(rois, shape)
        bad = got != want
        assert not np.any(bad & ~_tie(pre.numpy())), "device level differs from the reference away from a rounding tie"
        assert bad.sum() <= 15                                            # only the planted boundary squares may flip
        # and the split built on it reproduces nonzero order of the reference's masks (lib/sub_module.py:442,367-378)
        sp = fi.split_levels(torch.from_numpy(got).cuda())
        for i in range(4):
            np.testing.assert_array_equal(sp.small(i).cpu().numpy(), np.nonzero(got.reshape(-1) == i + 2)[0])
            np.testing.assert_array_equal(sp.big(i).cpu().numpy(), np.nonzero(got.reshape(-1) > i + 2)[0])


@pytest.mark.gpu
def test_segment_mean_cuda_vs_reference(Z):
    import feature_intertwiner_b200 as fi
    for tag in "abc":
        gt, feat = torch.from_numpy(Z["seg_%s_gt" % tag]), torch.from_numpy(Z["seg_%s_feat" % tag])
        m, c = fi.assign_feat2cls(gt.cuda(), feat.cuda().view(-1, 1024, 1, 1), 81)
        np.testing.assert_array_equal(c.cpu().numpy(), Z["seg_%s_cnt" % tag])
        np.testing.assert_allclose(m.cpu().numpy(), Z["seg_%s_mean" % tag], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_merge_and_meta_loss_cuda_vs_reference(Z):
    import feature_intertwiner_b200 as fi
    from feature_intertwiner_b200.dist import merged_class_sums
    s, n = merged_class_sums(torch.from_numpy(Z["merge_feat"]).cuda(), torch.from_numpy(Z["merge_cnt"]).cuda(), None, False, False, True)
    np.testing.assert_array_equal(n.cpu().numpy().reshape(1, 81), Z["merge_out_cnt"])
    np.testing.assert_allclose((s / (n + 1e-20)).cpu().numpy(), Z["merge_out_feat"], rtol=1e-5, atol=1e-6)
    for tag in _meta_cases(Z):
        cfg = _meta_cfg(tag)
        ot = None
        if cfg.DEV.LOSS_CHOICE == "ot":
            ot = fi.OptTrans(cfg, ch_x=64, L=5)
            ot.load_state_dict(_ot_state(Z))
        mod = fi.IntertwinerLoss(cfg, ot_loss=ot, feat_dim=64).cuda()
        for it in range(3):
            pre = "%s_it%d_" % (tag, it)
            inp = [torch.from_numpy(Z[pre + k]).cuda() for k in ("big_feat", "big_cnt", "small_feat", "small_cnt", "small_out", "small_gt")]
            with torch.no_grad():
                loss = mod(inp)
            np.testing.assert_allclose(loss.cpu().numpy().reshape(-1), Z[pre + "loss"], atol=LOSS_TOL, rtol=1e-5, err_msg=pre)
            fb, fbc = mod.fifo_buffer()
            np.testing.assert_allclose(fb.cpu().numpy(), Z[pre + "buffer"], rtol=1e-5, atol=1e-6, err_msg=pre)
            np.testing.assert_array_equal(fbc.cpu().numpy(), Z[pre + "buffer_cnt"])


@pytest.mark.gpu
def test_proposal_decode_cuda_vs_reference(Z):
    """fi_proposal_decode (gather + deltas + clip in one launch) against tools/box_utils.py's own apply_box_deltas / clip_boxes."""
    import feature_intertwiner_b200 as fi
    anchors, deltas = torch.from_numpy(Z["box_anchors"]), torch.from_numpy(Z["box_deltas"])
    bs, A = deltas.shape[:2]
    H, W = float(Z["box_window"][2]), float(Z["box_window"][3])
    cfg = pyref.make_config(DATA__IMAGE_SHAPE=np.array([int(H), int(W), 3]), RPN__PRE_NMS_LIMIT=A)
    cfg.DATA.BBOX_STD_DEV = np.array([1.0, 1.0, 1.0, 1.0])                   # the fixture's deltas are already scaled (lib/layers.py:94)
    fg = torch.linspace(1.0, 0.0, A).unsqueeze(0).expand(bs, A)              # strictly descending scores: the sort is the identity
    probs = torch.stack([1 - fg, fg], 2)
    boxes, dets = fi.proposal_decode([probs.cuda(), deltas.cuda()], anchors[0].cuda(), cfg)
    # expf on the device may differ from the host's in the last place: height/width scale by exp(delta)
    np.testing.assert_allclose(boxes.cpu().numpy(), Z["box_clipped"], rtol=3e-6, atol=2e-4)
    inside = (Z["box_applied"] == Z["box_clipped"])
    assert inside.mean() > 0.5 and (~inside).sum() > 0                      # both clipped and unclipped corners are exercised


# ======================================================================================= mask targets / detection layer (8 f3, f4)
@pytest.fixture(scope="module")
def T(golden_dir):
    return np.load(golden_dir + "/targets.npz")


def _det_cfg(min_conf):
    cfg = pyref.make_config(DATA__IMAGE_SHAPE=np.array([832, 1344, 3]))
    cfg.TEST = types.SimpleNamespace(DET_MAX_INSTANCES=100, DET_MIN_CONFIDENCE=min_conf, DET_NMS_THRESHOLD=0.3)
    cfg.DATA.BBOX_STD_DEV = np.array([0.1, 0.1, 0.2, 0.2])
    return cfg


def _same_detections(got, want):
    """Rows agree; scores that tie may come in either order (the reference's sort is unstable)."""
    assert got.shape == want.shape
    n_got, n_want = (got[:, :, 5] > 0).sum(1), (want[:, :, 5] > 0).sum(1)
    np.testing.assert_array_equal(n_got, n_want)
    for b in range(got.shape[0]):
        g = got[b][np.lexsort((got[b][:, 0], got[b][:, 1], -got[b][:, 5]))]
        w = want[b][np.lexsort((want[b][:, 0], want[b][:, 1], -want[b][:, 5]))]
        np.testing.assert_allclose(g[:, :5], w[:, :5], rtol=0, atol=0)
        np.testing.assert_allclose(g[:, 5], w[:, 5], rtol=1e-6, atol=0)


def test_targets_fixture_names_the_reference_code(T):
    cited = set(T["cited"].tolist())
    for need in ("mask_targets=lib/layers.py:297-323", "conduct_nms=lib/layers.py:664-717", "detection_layer=lib/layers.py:720-802"):
        assert need in cited, need


def test_mask_targets_oracle_vs_reference(T):
    for tag, mini in (("mask_mini", True), ("mask_full", False)):
        got = pyref.mask_targets_ref(torch.from_numpy(T[tag + "_pos_rois"]), torch.from_numpy(T[tag + "_gt_boxes"]), torch.from_numpy(T[tag + "_assign"]),
                                     torch.from_numpy(T[tag + "_gt_masks"]), (28, 28), mini)
        np.testing.assert_array_equal(got.numpy(), T[tag + "_targets"])
        assert 0.05 < T[tag + "_targets"].mean() < 0.95 and set(np.unique(T[tag + "_targets"])) == {0.0, 1.0}


def test_detection_layer_oracle_vs_reference(T):
    for key, conf in (("det_out", 0.3), ("det_out_conf093", 0.93)):
        got = pyref.detection_layer_ref(torch.from_numpy(T["det_rois"]), torch.from_numpy(T["det_probs"]), torch.from_numpy(T["det_deltas"]),
                                        torch.from_numpy(T["det_windows"]), _det_cfg(conf))
        _same_detections(got.numpy(), T[key])
    assert (T["det_out_conf093"][:, :, 5] > 0).sum(1).max() < 100 and (T["det_out"][:, :, 5] > 0).sum(1).min() == 100


@pytest.mark.gpu
def test_mask_targets_cuda_vs_reference(T):
    import feature_intertwiner_b200 as fi
    for tag, mini in (("mask_mini", True), ("mask_full", False)):
        got = fi.mask_targets(torch.from_numpy(T[tag + "_pos_rois"]).cuda(), torch.from_numpy(T[tag + "_gt_boxes"]).cuda(),
                              torch.from_numpy(T[tag + "_assign"]).cuda(), torch.from_numpy(T[tag + "_gt_masks"]).cuda(), (28, 28), mini)
        np.testing.assert_array_equal(got.cpu().numpy(), T[tag + "_targets"])
        # and the operator the reference itself calls (CropAndResizeFunction with C = 1, one image per box) gives the same targets
        a = torch.from_numpy(T[tag + "_assign"]).long()
        boxes = torch.from_numpy(T[tag + "_pos_rois"])
        if mini:
            gb = torch.from_numpy(T[tag + "_gt_boxes"])[a]
            gh, gw = gb[:, 2:3] - gb[:, 0:1], gb[:, 3:4] - gb[:, 1:2]
            boxes = torch.cat([(boxes[:, 0:1] - gb[:, 0:1]) / gh, (boxes[:, 1:2] - gb[:, 1:2]) / gw, (boxes[:, 2:3] - gb[:, 0:1]) / gh,
                               (boxes[:, 3:4] - gb[:, 1:2]) / gw], 1)
        crops = fi.CropAndResizeFunction(28, 28)(torch.from_numpy(T[tag + "_gt_masks"])[a].unsqueeze(1).cuda(), boxes.cuda(),
                                                 torch.arange(a.numel(), dtype=torch.int32).cuda())
        np.testing.assert_array_equal(torch.round(crops.squeeze(1)).cpu().numpy(), T[tag + "_targets"])


@pytest.mark.gpu
def test_detection_layer_cuda_vs_reference(T):
    import feature_intertwiner_b200 as fi
    for key, conf in (("det_out", 0.3), ("det_out_conf093", 0.93)):
        feat = torch.arange(600, dtype=torch.float32).view(600, 1).cuda()
        got, gfeat = fi.detection_layer(torch.from_numpy(T["det_rois"]).cuda(), torch.from_numpy(T["det_probs"]).cuda(),
                                        torch.from_numpy(T["det_deltas"]).cuda(), torch.from_numpy(T["det_windows"]).cuda(), _det_cfg(conf), feature=feat)
        got = got.cpu().numpy()
        # expf on the device may differ from the host's in the last place; after rounding to pixels the boxes agree except on exact .5 ties
        want = T[key]
        n_got, n_want = (got[:, :, 5] > 0).sum(1), (want[:, :, 5] > 0).sum(1)
        np.testing.assert_array_equal(n_got, n_want)
        for b in range(2):
            g = got[b][np.lexsort((got[b][:, 0], got[b][:, 1], -got[b][:, 5]))]
            w = want[b][np.lexsort((want[b][:, 0], want[b][:, 1], -want[b][:, 5]))]
            np.testing.assert_allclose(g[:, 5], w[:, 5], rtol=1e-6)
            np.testing.assert_array_equal(g[:, 4], w[:, 4])
            assert np.abs(g[:, :4] - w[:, :4]).max() <= 1.0 and (g[:, :4] != w[:, :4]).mean() < 0.01
        # the gathered feature rows name the RoI every detection came from
        src = gfeat.cpu().numpy()[:, :, 0].astype(np.int64)
        probs = T["det_probs"]
        for b in range(2):
            k = int(n_got[b])
            np.testing.assert_allclose(probs[src[b, :k]].max(1), got[b, :k, 5], rtol=1e-6)
