"""Sinkhorn / OptTrans / level rule / split / segment mean / buffer / NMS / RoIPool: CUDA path (through the
C ABI) vs the oracle and the golden vectors generated from the reference itself."""
import numpy as np
import pytest
import torch

from oracle import clib, pyref

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-4   # north_star: "loss within 1e-4 fp32 of the reference"


def _fi():
    import feature_intertwiner_b200 as fi
    return fi


# ------------------------------------------------------------------------------------------------ Sinkhorn
def test_sinkhorn_golden(golden_dir):
    """Values produced by the reference's own OptTrans._sinkhorn_iterate (lib/OT_module.py:104-135)."""
    fi = _fi()
    z = np.load(golden_dir + "/sinkhorn.npz")
    for name in ("n256_d1", "n64_d256", "n16_d3"):
        x = torch.from_numpy(z[name + "_x"]).cuda()[None]
        y = torch.from_numpy(z[name + "_y"]).cuda()[None]
        for L in (1, 5, 50):
            for eps in (1.0, 0.1):
                got = fi.sinkhorn_loss(x, y, epsilon=eps, L=L).item()
                want = float(z[f"{name}_L{L}_eps{eps}"])
                assert abs(got - want) < LOSS_TOL, (name, L, eps, got, want)


@pytest.mark.parametrize("N,D", [(256, 1), (256, 5), (64, 256), (64, 1024), (100, 7), (224, 2), (1, 1), (33, 40)])
def test_sinkhorn_vs_oracle_values_and_grads(N, D):
    fi = _fi()
    g = torch.Generator().manual_seed(N * 1000 + D)
    P = 5
    x = torch.randn(P, N, D, generator=g).abs()
    y = torch.randn(P, N, D, generator=g).abs()
    if N > 8:
        x[:, ::9] = 0
    xc, yc = x.cuda().requires_grad_(), y.cuda().requires_grad_()
    loss = fi.sinkhorn_loss(xc, yc, epsilon=0.5, L=20)
    w = torch.arange(1, P + 1, dtype=torch.float32).cuda()
    (loss * w).sum().backward()
    for p in range(P):
        want, _, gx, gy = clib.oracle_sinkhorn(x[p].numpy(), y[p].numpy(), inv_eps=2.0, L=20, wide=True, want_grad=True)
        assert abs(loss[p].item() - want) < 2e-5, (p, loss[p].item(), want)
        rows = (x[p].norm(dim=1) > 0).numpy()      # d/dx of x/(|x|+1e-20) at x == 0 is 1e20: compare the well-posed rows
        np.testing.assert_allclose(xc.grad[p].cpu().numpy()[rows] / w[p].item(), gx[rows], rtol=2e-3, atol=2e-6)
        np.testing.assert_allclose(yc.grad[p].cpu().numpy() / w[p].item(), gy, rtol=2e-3, atol=2e-6)


@pytest.mark.parametrize("N", [256, 100, 1])
def test_sinkhorn_d1_class_kernel_vs_dense_kernels_and_oracle(N):
    """D = 1: problems whose normalised entries are exactly -1 / 0 / +1 are solved by classes (sinkhorn_d1_classes_kernel, O(L) per
    problem); problems with an entry within a few orders of 1e-20 fall through to the dense kernels (masked launch).  Both
    against the fp64-accumulated oracle and against the dense kernels alone (fi_set_option: K in registers / generic)."""
    fi = _fi()
    g = torch.Generator().manual_seed(40 + N)
    P = 12
    x = torch.randn(P, N, 1, generator=g).abs()
    y = torch.randn(P, N, 1, generator=g).abs()
    x[:, ::5] = 0                                           # ReLU zeros
    y[2:, ::7] = 0
    x[3] = -x[3]                                            # a problem with -1 entries (never produced by the critic; API is general)
    if N > 1:
        y[4, 1::2] *= -1
        x[5, 3] = 2.5e-20                                   # x^ = 0.71...: problem 5 must go to the dense kernel
        y[6, 0] = 7.0e-21
        x[7] = 0                                            # all-zero rows: C = 1 everywhere
    res = {}
    for mode in (0, 1, 2):
        old = fi.set_option("sinkhorn_generic", mode)
        try:
            xc, yc = x.cuda().requires_grad_(), y.cuda().requires_grad_()
            loss = fi.sinkhorn_loss(xc, yc, epsilon=0.7, L=30)
            loss.sum().backward()
            res[mode] = (loss.detach().cpu(), xc.grad.cpu(), yc.grad.cpu())
        finally:
            fi.set_option("sinkhorn_generic", old)
    for p in range(P):
        want, _, gx, gy = clib.oracle_sinkhorn(x[p].numpy(), y[p].numpy(), inv_eps=1.0 / 0.7, L=30, wide=True, want_grad=True)
        for mode in (0, 1, 2):
            assert abs(res[mode][0][p].item() - want) < 2e-5, (N, p, mode, res[mode][0][p].item(), want)
        well = (x[p].abs().view(-1) > 1e-10).numpy()        # d/dx of x / (|x| + 1e-20) is 1e20 at 0: compare the well-posed rows
        welly = (y[p].abs().view(-1) > 1e-10).numpy()
        np.testing.assert_allclose(res[0][1][p].numpy()[well], gx[well], rtol=2e-3, atol=2e-6)
        np.testing.assert_allclose(res[0][2][p].numpy()[welly], gy[welly], rtol=2e-3, atol=2e-6)
    # the class path reproduces the dense kernels to rounding, ill-posed rows included
    torch.testing.assert_close(res[0][0], res[1][0], rtol=2e-6, atol=2e-7)
    torch.testing.assert_close(res[0][0], res[2][0], rtol=2e-6, atol=2e-7)
    for k in (1, 2):
        torch.testing.assert_close(res[0][k], res[1][k], rtol=1e-4, atol=1e-7)


def test_sinkhorn_large_d_workspace_form_vs_one_cta_form():
    """Large D (FPN-level loss shapes): with a workspace the cost matrix is summed over slices of D by (problem, slice, tile) CTAs and
    the gradient runs as its own (problem, D chunk) launches; without one everything stays in one CTA per problem.  Same
    arithmetic, different summation order over D: equal to rounding, and the workspace form is deterministic."""
    from feature_intertwiner_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(9)
    s = torch.cuda.current_stream().cuda_stream
    for (P, N, D) in ((6, 64, 256), (3, 64, 1500), (2, 128, 128), (4, 100, 300), (24, 64, 4096)):
        x = torch.randn(P, N, D, generator=g).abs().cuda()
        y = torch.randn(P, N, D, generator=g).abs().cuda()
        x[:, 3] = 0
        res = []
        for use_ws in (False, True, True):
            loss = torch.empty(P, device="cuda")
            gx, gy = torch.full_like(x, 7.0), torch.full_like(y, 7.0)
            nbytes = L.fi_sinkhorn_workspace(P, N, D, 1) if use_ws else 0
            assert (not use_ws) or (nbytes >= P * (2 * N * N + 2 * N) * 4 and nbytes == L.fi_sinkhorn_workspace(P, N, D, 0))
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device="cuda")
            _lib.check(L.fi_sinkhorn_ws(x.data_ptr(), y.data_ptr(), P, N, D, 2.0, 5, loss.data_ptr(), gx.data_ptr(), gy.data_ptr(),
                                        ws.data_ptr() if use_ws else None, nbytes, s))
            res.append((loss, gx, gy))
        for a, b in zip(res[1], res[2]):
            assert torch.equal(a, b)                                       # deterministic
        torch.testing.assert_close(res[0][0], res[1][0], rtol=1e-5, atol=1e-7)
        well = x.norm(dim=2) > 0                                           # d/dx of x / (|x| + 1e-20) at x == 0 is 1e20
        torch.testing.assert_close(res[0][1][well], res[1][1][well], rtol=1e-3, atol=1e-7)
        torch.testing.assert_close(res[0][2], res[1][2], rtol=1e-3, atol=1e-7)
    assert L.fi_sinkhorn_workspace(240, 256, 1, 1) == 0 and L.fi_sinkhorn_workspace(24, 64, 64, 1) == 0


def test_sinkhorn_properties_full_size():
    """Size-independent properties at BASELINE sizes: x == y => debiased loss 0; batch order invariance."""
    fi = _fi()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(240, 256, 1, generator=g).abs().cuda()
    y = torch.randn(240, 256, 1, generator=g).abs().cuda()
    wxy = fi.sinkhorn_loss(x, y, 1.0, 50)
    wxx = fi.sinkhorn_loss(x, x, 1.0, 50)
    wyx = fi.sinkhorn_loss(y, x, 1.0, 50)
    assert torch.all(torch.isfinite(wxy))
    # W(x,x) with x^ in {0,1}: identical arguments -> 2W(x,x) - W(x,x) - W(x,x) == 0 exactly
    assert torch.equal(2 * wxx - wxx - wxx, torch.zeros_like(wxx))
    perm = torch.randperm(240, generator=g).cuda()
    assert torch.equal(fi.sinkhorn_loss(x[perm], y[perm], 1.0, 50), wxy[perm])     # problems are independent, deterministic
    assert wyx.shape == wxy.shape


def test_opttrans_golden(golden_dir):
    """OptTrans.forward of the reference (1-D and 2-D), same weights."""
    fi = _fi()
    z = np.load(golden_dir + "/opttrans.npz")
    cfg = pyref.make_config()
    m = fi.OptTrans(cfg, ch_x=64, L=5).eval()
    m.load_state_dict({k[len("d1_sd_"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("d1_sd_")})
    m.cuda()
    got = m(torch.from_numpy(z["d1_x"]).cuda(), torch.from_numpy(z["d1_y"]).cuda())
    np.testing.assert_allclose(got.detach().cpu().numpy(), z["d1_loss"], atol=LOSS_TOL, rtol=0)
    m2 = fi.OptTrans(cfg, ch_x=16, spatial_x=8, spatial_y=16, L=5).eval()
    m2.load_state_dict({k[len("d2_sd_"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("d2_sd_")})
    m2.cuda()
    got2 = m2(torch.from_numpy(z["d2_x"]).cuda(), torch.from_numpy(z["d2_y"]).cuda())
    np.testing.assert_allclose(got2.detach().cpu().numpy(), z["d2_loss"], atol=LOSS_TOL, rtol=0)


def test_opttrans_backward_matches_restatement():
    fi = _fi()
    torch.manual_seed(0)
    ref = pyref.OptTransRef(ch_x=32, L=7)
    m = fi.OptTrans(pyref.make_config(), ch_x=32, L=7)
    m.load_state_dict(ref.state_dict())
    m.cuda()
    x, y = torch.randn(4, 32, 1), torch.randn(4, 32, 1).abs()
    xr = x.clone().requires_grad_()
    lr = ref(xr, y)
    lr.sum().backward()
    xc = x.cuda().requires_grad_()
    lc = m(xc, y.cuda())
    lc.sum().backward()
    np.testing.assert_allclose(lc.detach().cpu().numpy(), lr.detach().numpy(), atol=LOSS_TOL, rtol=0)
    np.testing.assert_allclose(xc.grad.cpu().numpy(), xr.grad.numpy(), atol=1e-5, rtol=1e-3)
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        np.testing.assert_allclose(p.grad.cpu().numpy(), q.grad.numpy(), atol=1e-5, rtol=1e-3, err_msg=n)


# ------------------------------------------------------------------------------------------------ level / split
def test_roi_level_bit_exact_vs_torch_and_oracle():
    fi = _fi()
    from feature_intertwiner_b200 import synth
    g = torch.Generator().manual_seed(2000)
    rois = synth.make_rois(8, 2000, (832, 1344), g)
    shape = (832, 1344, 3)
    got = fi.roi_level(rois.cuda(), shape, 224.0).cpu()
    # (1) the reference's op sequence executed by torch on the SAME device (what the reference itself would do)
    r = rois.cuda()
    y1, x1, y2, x2 = r.chunk(4, dim=2)
    area = (x2 - x1) * (y2 - y1)
    ia = torch.tensor([float(shape[0] * shape[1])]).cuda()
    ln2 = torch.log(torch.tensor([2.0])).cuda()
    lvl = (4 + torch.log(torch.sqrt(area) / (224.0 / torch.sqrt(ia))) / ln2).round().clamp(2, 5).int().squeeze(-1).cpu()
    assert torch.equal(got, lvl)
    # (2) the C oracle; a mismatch is only tolerated on an exact .5 tie of the pre-round value (none expected)
    want, pre = clib.oracle_roi_level(rois.numpy().reshape(-1, 4), float(shape[0] * shape[1]), 224.0)
    diff = got.numpy().reshape(-1) != want
    assert not np.any(diff & (np.abs(pre - np.floor(pre) - 0.5) > 1e-5))
    assert set(np.unique(want)) == {2, 3, 4, 5}
    assert np.all(got.numpy()[:, -100:] == 2)        # zero-padded RoIs: log(0) = -inf -> level 2


@pytest.mark.parametrize("n", [0, 1, 257, 4096, 16000])
def test_split_matches_nonzero(n):
    fi = _fi()
    g = torch.Generator().manual_seed(n)
    level = torch.randint(2, 6, (n,), generator=g, dtype=torch.int32).cuda()
    if n == 257:
        level[:] = 3                                  # empty lists
    sp = fi.split_levels(level)
    for i, l in enumerate(range(2, 6)):
        small = torch.nonzero(level == l).squeeze(1).int()
        big = torch.nonzero(level > l).squeeze(1).int()
        assert sp.small_cnt[i] == small.numel() and sp.big_cnt[i] == big.numel()
        assert torch.equal(sp.small(i), small) and torch.equal(sp.big(i), big)
        if small.numel():
            assert torch.equal(sp.slot[small.long()], torch.arange(small.numel(), dtype=torch.int32).cuda())
    if n >= 257:
        # same launch + gathers: rois[idx], idx // R, gt[idx] (lib/sub_module.py:489-493,541-548)
        R = 257 if n == 257 else 16
        g2 = torch.Generator().manual_seed(n + 1)
        rois = torch.rand(n // R, R, 4, generator=g2).cuda()
        gt = torch.randint(0, 81, (n // R, R), generator=g2, dtype=torch.int32).cuda()
        lv = level[: (n // R) * R].view(n // R, R)
        sg = fi.split_levels(lv, rois=rois, gt=gt)
        flat_r, flat_g = rois.view(-1, 4), gt.view(-1)
        for i, l in enumerate(range(2, 6)):
            for idx, boxes, ind, cls in ((sg.small(i), sg.small_boxes(i), sg.small_ind(i), sg.small_gt(i)),
                                         (sg.big(i), sg.big_boxes(i), sg.big_ind(i), sg.big_gt(i))):
                assert torch.equal(boxes, flat_r[idx.long()]) and torch.equal(ind, (idx // R).int()) and torch.equal(cls, flat_g[idx.long()])
        # a visiting order: same sets, listed in that order
        order = fi.spatial_order(rois)
        assert torch.equal(torch.sort(order.long())[0], torch.arange(order.numel()).cuda())
        so = fi.split_levels(lv, rois=rois, gt=gt, order=order)
        for i, l in enumerate(range(2, 6)):
            want = order[(lv.view(-1)[order.long()] == l)]
            assert torch.equal(so.small(i), want) and torch.equal(so.small_boxes(i), flat_r[want.long()])
            assert torch.equal(torch.sort(so.big(i))[0], torch.sort(sg.big(i))[0])


# ------------------------------------------------------------------------------------------------ segment mean / buffer
@pytest.mark.parametrize("k", [0, 1, 500, 3000])
def test_segment_mean_fwd_bwd(k):
    fi = _fi()
    g = torch.Generator().manual_seed(k)
    gt = torch.randint(0, 81, (k,), generator=g, dtype=torch.int32)
    if k > 10:
        gt[k // 2:] = 0
        gt[:5] = 7
    feat = torch.randn(k, 1024, 1, 1, generator=g)
    want_f, want_c = clib.oracle_segment_mean(gt.numpy(), feat.view(k, 1024).numpy(), 81)
    fc = feat.cuda().requires_grad_()
    mean, cnt = fi.assign_feat2cls(gt.cuda(), fc, 81)
    np.testing.assert_array_equal(cnt.cpu().numpy(), want_c)
    np.testing.assert_allclose(mean.detach().cpu().numpy(), want_f, rtol=1e-5, atol=1e-6)
    if k == 0:
        assert float(mean.abs().sum()) == 0
        return
    w = torch.randn(1024, 81, generator=g)
    (mean * w.cuda()).sum().backward()
    fr = feat.clone().requires_grad_()
    mr, _ = pyref.assign_feat2cls_ref(gt.long(), fr, 81)
    (mr * w).sum().backward()
    np.testing.assert_allclose(fc.grad.cpu().numpy(), fr.grad.numpy(), rtol=1e-5, atol=1e-7)


def test_segment_mean_multi_and_spatial_order_match_the_single_forms():
    """assign_feat2cls_multi (one launch for several lists, device-side lengths) == assign_feat2cls list by list, forward and
    backward; fi_spatial_order == the torch formulation of the same visiting order."""
    fi = _fi()
    from feature_intertwiner_b200 import synth
    g = torch.Generator().manual_seed(77)
    lists, singles = [], []
    for k, live in ((300, 300), (1, 1), (50, 20), (0, 0), (700, 650)):
        gt = torch.randint(0, 81, (k,), generator=g, dtype=torch.int32).cuda()
        f = torch.randn(k, 1024, generator=g).cuda()
        cnt = torch.tensor([live], dtype=torch.int32, device="cuda") if live != k else None
        a, b = f.clone().requires_grad_(), f.clone().requires_grad_()
        lists.append((gt, a, cnt)); singles.append((gt, b, cnt))
    multi = fi.assign_feat2cls_multi(lists, 81)
    w = [torch.randn(1024, 81, generator=g).cuda() for _ in lists]
    sum((m * wi).sum() for (m, _), wi in zip(multi, w)).backward()
    for (gt, b, cnt), (m, c), wi, (_, a, _) in zip(singles, multi, w, lists):
        ms, cs = fi.assign_feat2cls(gt, b, 81, count=cnt)
        assert torch.equal(ms, m) and torch.equal(cs, c)
        if b.numel():
            (ms * wi).sum().backward()
            assert torch.equal(a.grad, b.grad)
    for (bs, R, hw) in ((8, 512, (832, 1344)), (4, 2000, (832, 1344)), (2, 1, (256, 256))):
        rois = synth.make_rois(bs, R, hw, g).cuda()
        got = fi.spatial_order(rois)
        grid = 8
        cy = ((rois[..., 0] + rois[..., 2]) * (0.5 * grid)).clamp(0, grid - 1).floor()
        cx = ((rois[..., 1] + rois[..., 3]) * (0.5 * grid)).clamp(0, grid - 1).floor()
        snake = torch.where(cy.long() % 2 == 0, cx, grid - 1 - cx)
        key = (torch.arange(bs, device="cuda").view(bs, 1) * (grid * grid) + cy * grid + snake).view(-1)
        want = torch.sort(key, stable=True)[1].int()
        assert torch.equal(got, want)


@pytest.mark.parametrize("B,loss", [(1, "l2"), (1, "l1"), (3, "l2"), (1, "ot")])
@pytest.mark.parametrize("inst", [False, True])
def test_intertwiner_loss_vs_restatement(B, loss, inst):
    fi = _fi()
    torch.manual_seed(B * 10 + inst)
    cfg = pyref.make_config(DEV__BUFFER_SIZE=B, DEV__LOSS_CHOICE=loss, DEV__INST_LOSS=inst)
    Fd, ncls, S, n_inst = 1024, 81, 3, 60
    # The restatement runs on the SAME device: with D=1 the cosine normalisation is a sign function of the critic
    # output, so a 1e-7 difference between a CPU and a cuDNN convolution can flip x^ between
    # 0 and 1.  Same device => identical stock-conv outputs => only the Sinkhorn kernel itself is under test.
    ot_ref = pyref.OptTransRef(ch_x=Fd, L=5).cuda() if loss == "ot" else None
    ref = pyref.MetaLossRef(cfg, Fd, ot_loss=ot_ref)
    ref.buffer, ref.buffer_cnt = ref.buffer.cuda(), ref.buffer_cnt.cuda()
    ot = None
    if loss == "ot":
        ot = fi.OptTrans(cfg, ch_x=Fd, L=5)
        ot.load_state_dict(ot_ref.state_dict())
    mod = fi.IntertwinerLoss(cfg, ot_loss=ot, feat_dim=Fd).cuda()
    for step in range(4):                                    # several iterations: the buffer is state
        def stats():
            cnt = torch.randint(0, 4, (1, S, 1, ncls)).float()
            feat = torch.rand(1, S, Fd, ncls) * (cnt > 0)
            return feat, cnt
        bf, bc = stats()
        sf, sc = stats()
        so = torch.rand(n_inst, Fd)
        sg = torch.randint(0, ncls, (n_inst,)).float()
        sf_r, so_r = sf.cuda().requires_grad_(), so.cuda().requires_grad_()
        want = ref([bf.cuda(), bc.cuda(), sf_r, sc.cuda(), so_r, sg.cuda()])
        sf_c, so_c = sf.cuda().requires_grad_(), so.cuda().requires_grad_()
        got = mod([bf.cuda(), bc.cuda(), sf_c, sc.cuda(), so_c, sg.cuda()])
        np.testing.assert_allclose(got.detach().cpu().numpy().reshape(-1), want.detach().cpu().numpy().reshape(-1), atol=LOSS_TOL, rtol=1e-5)
        fb, fbc = mod.fifo_buffer()
        np.testing.assert_allclose(fb.cpu().numpy(), ref.buffer.cpu().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(fbc.cpu().numpy(), ref.buffer_cnt.cpu().numpy())
        if want.requires_grad:
            want.sum().backward(); got.sum().backward()
            a, b = (so_c, so_r) if inst else (sf_c, sf_r)
            # OT with D=1: d/dx of x/(|x|+1e-20) underflows to exactly 0 in the kernel; autograd leaves ~1e-7 rounding noise
            np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.cpu().numpy(), rtol=1e-3, atol=1e-6 if loss == "ot" else 1e-7)


def test_intertwiner_ot_padded_equals_compact():
    """ot_padded=True (fixed shapes, no nonzero sync) gives the same total loss and gradients as the reference-shaped [n] vector."""
    fi = _fi()
    torch.manual_seed(3)
    cfg = pyref.make_config(DEV__LOSS_CHOICE="ot")
    ot = fi.OptTrans(cfg, ch_x=1024, L=5).cuda()
    a = fi.IntertwinerLoss(cfg, ot_loss=ot).cuda()
    b = fi.IntertwinerLoss(cfg, ot_loss=ot, ot_padded=True).cuda()
    cnt = torch.randint(0, 3, (1, 3, 1, 81)).float().cuda()
    bf = torch.rand(1, 3, 1024, 81).cuda() * (cnt > 0)
    scnt = torch.randint(0, 3, (1, 3, 1, 81)).float().cuda()
    sf = torch.rand(1, 3, 1024, 81).cuda() * (scnt > 0)
    s1, s2 = sf.clone().requires_grad_(), sf.clone().requires_grad_()
    la = a([bf, cnt, s1, scnt, None, None])
    lb = b([bf, cnt, s2, scnt, None, None])
    assert lb.numel() == 80 and la.numel() == int(a.last_idx.numel())
    torch.testing.assert_close(la.sum(), lb.sum(), rtol=1e-5, atol=1e-6)
    la.sum().backward(); lb.sum().backward()
    torch.testing.assert_close(s1.grad, s2.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("surrogate", [False, True])
def test_fused_loss_head_matches_torch_ops(surrogate, monkeypatch):
    """The kernel-fused class-level OT head (dist._MergeStats + intertwiner._ClassOTHead: csrc/loss_head.cu between library GEMMs)
    against the same head written in torch ops (fused_head = False, FUSED_MERGE = False): loss vector, buffers over two
    iterations, and every gradient -- class means and the four OptTrans parameters.  With one feature position per row the real
    Sinkhorn gradient is exactly zero behind the critic's ReLU, so the backward chain is ALSO run with a smooth stand-in for the
    Sinkhorn launch (same call signature, patched under both heads) that gives dense, non-zero gradients."""
    fi = _fi()
    from feature_intertwiner_b200 import dist as fdist, ot as fot
    if surrogate:
        def fake(x, y, inv_eps, L, need_grad):
            x, y = x.detach().float(), y.detach().float()
            loss = (x * y).sum(dim=(1, 2)) * 1e-2 + 0.5e-2 * (x * x).sum(dim=(1, 2))
            return loss, ((y + x) * 1e-2 if need_grad else None), (x * 1e-2 if need_grad else None)
        monkeypatch.setattr(fot, "sinkhorn_raw", fake)
    torch.manual_seed(11)
    cfg = pyref.make_config(DEV__LOSS_CHOICE="ot")
    ot_a = fi.OptTrans(cfg, ch_x=1024, L=5).cuda()
    ot_b = fi.OptTrans(cfg, ch_x=1024, L=5).cuda()
    ot_b.load_state_dict(ot_a.state_dict())
    a = fi.IntertwinerLoss(cfg, ot_loss=ot_a, ot_padded=True).cuda()          # fused (default)
    b = fi.IntertwinerLoss(cfg, ot_loss=ot_b, ot_padded=True).cuda()
    b.fused_head = False
    assert a.fused_head
    for it in range(2):
        cnt = torch.randint(0, 3, (1, 3, 1, 81)).float().cuda()
        bf = torch.rand(1, 3, 1024, 81).cuda() * (cnt > 0)
        scnt = torch.randint(0, 3, (1, 3, 1, 81)).float().cuda()
        sf = (torch.rand(1, 3, 1024, 81).cuda() - 0.3) * (scnt > 0)
        s1, s2 = sf.clone().requires_grad_(), sf.clone().requires_grad_()
        up = torch.rand(80).cuda() + 0.5
        n0 = fi.lib().fi_kernel_launches()
        la = a([bf, cnt, s1, scnt, None, None])
        assert fi.lib().fi_kernel_launches() - n0 >= 4             # merge, buffer update, prep, combine are this library's kernels
        monkeypatch.setattr(fdist, "FUSED_MERGE", False)
        lb = b([bf, cnt, s2, scnt, None, None])
        monkeypatch.setattr(fdist, "FUSED_MERGE", True)
        assert la.shape == lb.shape == (80,)
        torch.testing.assert_close(la, lb, rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(a.buffer, b.buffer, rtol=1e-6, atol=1e-7)
        assert torch.equal(a.buffer_cnt, b.buffer_cnt)
        (la * up).sum().backward()
        (lb * up).sum().backward()
        torch.testing.assert_close(s1.grad, s2.grad, rtol=2e-4, atol=1e-7)
        if surrogate:
            assert float(s1.grad.abs().max()) > 0
        for (name, pa), (_, pb) in zip(ot_a.named_parameters(), ot_b.named_parameters()):
            assert pa.grad is not None and pb.grad is not None, name
            torch.testing.assert_close(pa.grad, pb.grad, rtol=2e-4, atol=1e-6, msg=lambda m, name=name: "%s: %s" % (name, m))
            if surrogate:
                assert float(pa.grad.abs().max()) > 0, name
            pa.grad = None
            pb.grad = None


@pytest.mark.parametrize("loss", ["l2", "ot"])
def test_intertwiner_loss_cuda_graph_matches_eager(loss):
    """enable_cuda_graph(): graphed forward+backward of the loss head == eager, including the buffer's evolution."""
    fi = _fi()
    torch.manual_seed(4)
    cfg = pyref.make_config(DEV__LOSS_CHOICE=loss)
    ot_a = fi.OptTrans(cfg, ch_x=1024, L=5).cuda() if loss == "ot" else None
    ot_b = fi.OptTrans(cfg, ch_x=1024, L=5).cuda() if loss == "ot" else None
    if loss == "ot":
        ot_b.load_state_dict(ot_a.state_dict())
    a = fi.IntertwinerLoss(cfg, ot_loss=ot_a, ot_padded=True).cuda()
    b = fi.IntertwinerLoss(cfg, ot_loss=ot_b, ot_padded=True).cuda()

    def batch(seed):
        g = torch.Generator().manual_seed(seed)
        cnt = torch.randint(0, 3, (1, 3, 1, 81), generator=g).float().cuda()
        bf = torch.rand(1, 3, 1024, 81, generator=g).cuda() * (cnt > 0)
        scnt = torch.randint(0, 3, (1, 3, 1, 81), generator=g).float().cuda()
        sf = torch.rand(1, 3, 1024, 81, generator=g).cuda() * (scnt > 0)
        return bf, cnt, sf, scnt
    bf, cnt, sf, scnt = batch(0)
    assert b.enable_cuda_graph([bf, cnt, sf.clone().requires_grad_(), scnt]), getattr(b, "_graph_error", "")
    assert float(b.buffer_cnt.sum()) == 0          # capture left the buffer untouched
    for step in range(3):
        bf, cnt, sf, scnt = batch(step + 1)
        s1, s2 = sf.clone().requires_grad_(), sf.clone().requires_grad_()
        la = a([bf, cnt, s1, scnt, None, None]).sum()
        lb = b([bf, cnt, s2, scnt, None, None]).sum()
        la.backward(); lb.backward()
        torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(s1.grad, s2.grad, rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(a.buffer, b.buffer, rtol=1e-6, atol=1e-7)
        if loss == "ot":
            for p, q in zip(ot_a.parameters(), ot_b.parameters()):
                torch.testing.assert_close(p.grad, q.grad, rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------------------------ NMS
def test_nms_golden_and_oracle(golden_dir):
    fi = _fi()
    z = np.load(golden_dir + "/nms.npz")
    d = torch.from_numpy(z["dets"]).cuda()
    for thr in (0.3, 0.5, 0.7):
        got = fi.pth_nms(d, thr).cpu().numpy()
        strict = clib.oracle_nms(z["dets"][:, [1, 0, 3, 2, 4]], thr, True)
        np.testing.assert_array_equal(got, strict)
        # reference's compiled cpu_nms uses >=; identical unless an IoU sits exactly on the threshold
        np.testing.assert_array_equal(got, z[f"keep_cpu_{thr}"])


@pytest.mark.parametrize("n", [1, 63, 64, 65, 1000, 6000])
def test_nms_batched_vs_oracle(n):
    fi = _fi()
    from feature_intertwiner_b200 import synth
    g = torch.Generator().manual_seed(n)
    dets = synth.make_nms_boxes(3, n, g)
    keep = fi.nms(dets.cuda(), 0.7)
    want = pyref.nms_ref(dets, 0.7, strict=True)
    assert keep.dtype == np.int32
    np.testing.assert_array_equal(keep, want)


@pytest.mark.parametrize("n,limit", [(6000, 1000), (6000, 1), (1000, 100), (200, 500), (65, 64), (9000, 2000)])
def test_nms_early_exit_keeps_the_same_head(n, limit):
    """fi_nms_batched_topk: the sweep stops once `limit` survivors are known; the first `limit` entries equal the full sweep's."""
    from feature_intertwiner_b200 import synth
    from feature_intertwiner_b200.nms import nms_presorted
    g = torch.Generator().manual_seed(n + limit)
    dets = synth.make_nms_boxes(3, n, g)
    order = torch.sort(dets[:, :, 4], dim=1, descending=True, stable=True)[1]
    srt = torch.gather(dets, 1, order.unsqueeze(2).expand(3, n, 5))[:, :, [1, 0, 3, 2, 4]].contiguous().cuda()
    full, num_full = nms_presorted(srt, 0.7)
    cut, num_cut = nms_presorted(srt, 0.7, max_keep=limit)
    for b in range(3):
        m = min(limit, int(num_full[b]))
        assert int(num_cut[b]) >= m and int(num_cut[b]) <= int(num_full[b])
        assert torch.equal(cut[b, :m], full[b, :m])
        k = int(num_cut[b])
        assert torch.equal(cut[b, :k], full[b, :k]) and bool((cut[b, k:] == -1).all())


def test_nms_known_answers_and_threshold_rule():
    fi = _fi()
    # disjoint boxes keep all; identical boxes keep the first
    d = torch.tensor([[0, 0, 10, 10, .9], [20, 20, 30, 30, .8], [0, 0, 10, 10, .7]], dtype=torch.float32).cuda()
    assert fi.pth_nms(d, 0.5).tolist() == [0, 1]
    # IoU exactly at the threshold: the GPU rule (>) keeps, the CPU rule (>=) would suppress (Appendix B.7)
    a = torch.tensor([[0, 0, 9, 9, .9], [0, 0, 9, 4, .8]], dtype=torch.float32).cuda()     # inter 50, union 100 -> 0.5
    assert fi.pth_nms(a, 0.5).tolist() == [0, 1]
    assert clib.oracle_nms(a.cpu().numpy()[:, [1, 0, 3, 2, 4]], 0.5, False).tolist() == [0]
    # mask words of the reference-named _nms launcher
    from feature_intertwiner_b200 import _lib
    b = synth_boxes = torch.tensor([[0, 0, 10, 10, .9], [1, 1, 10, 10, .8], [50, 50, 60, 60, .7]], dtype=torch.float32).cuda()
    mask = torch.zeros(3, 1, dtype=torch.int64).cuda()
    _lib.lib()._nms(3, b.data_ptr(), mask.data_ptr(), 0.5)
    torch.cuda.synchronize()
    assert mask.view(-1).tolist() == [2, 0, 0]


# ------------------------------------------------------------------------------------------------ RoIPool
@pytest.mark.parametrize("scale", [0.25, 0.0625])
def test_roi_pool_fwd_bwd(scale):
    fi = _fi()
    g = torch.Generator().manual_seed(1)
    B, C, H, W, R = 2, 16, 30, 40, 50
    feat = torch.randn(B, C, H, W, generator=g)
    xy = torch.rand(R, 2, generator=g) * torch.tensor([W / scale, H / scale])
    wh = torch.rand(R, 2, generator=g) * 60 + 1
    rois = torch.cat([torch.randint(0, B, (R, 1), generator=g).float(), xy, xy + wh], 1)
    rois[0, 1:] = torch.tensor([40., 40., 20., 20.])          # malformed (end < start): forced 1x1, no gradient
    rois[1, 1:] = torch.tensor([1e4, 1e4, 1e4 + 5, 1e4 + 5])  # outside the map: empty bins -> 0 / argmax -1
    top, arg = clib.oracle_roi_pool_fwd(feat.numpy(), rois.numpy(), 7, 7, scale)
    fc = feat.cuda().requires_grad_()
    out = fi.RoIPoolFunction(7, 7, scale)(fc, rois.cuda())
    np.testing.assert_array_equal(out.detach().cpu().numpy(), top)
    gy = torch.randn(out.shape, generator=g)
    out.backward(gy.cuda())
    want = clib.oracle_roi_pool_bwd(gy.numpy(), arg, rois.numpy(), feat.shape, scale)
    np.testing.assert_allclose(fc.grad.cpu().numpy(), want, rtol=1e-5, atol=1e-5)   # atomics: summation order
    assert np.all(top[1] == 0) and np.all(arg[1] == -1)


# ------------------------------------------------------------------------------------------------ Dev end to end
@pytest.mark.parametrize("fmt", ["nchw", "nhwc"])
@pytest.mark.parametrize("loss", ["l2", "ot"])
def test_dev_forward_backward_vs_restatement(fmt, loss):
    """Dev.forward + meta loss vs the CPU restatement (lib/sub_module.py:380-642, lib/model.py:143-210), same weights."""
    fi = _fi()
    from feature_intertwiner_b200 import synth
    torch.manual_seed(5)
    g = torch.Generator().manual_seed(5)
    shape = (256, 256, 3)                       # BASELINE.json config 1: 4 x 256x256 images, 64 RoIs/img, P2 64^2 .. P5 8^2
    cfg = pyref.make_config(DATA__IMAGE_SHAPE=np.array(shape), DEV__LOSS_CHOICE=loss)
    depth, bs, R = 256, 4, 64
    ref = pyref.DevRef(cfg, depth=depth, feat_dim=1024).eval()
    dev = fi.Dev(cfg, depth).eval()
    dev.load_state_dict(ref.state_dict())
    dev.cuda()
    maps = [torch.randn(bs, depth, 64 >> i, 64 >> i, generator=g) for i in range(4)]
    rois = synth.make_rois(bs, R, (256, 256), g, zero_frac=0.1)
    rois[:, :, :] = rois[:, :, :] * 1.0
    # shrink the side range so every level 2..5 is populated at this image size
    gt = synth.make_class_ids(bs, R, g)
    maps_r = [m.clone().requires_grad_() for m in maps]
    po_r, mo_r, fo_r = ref(maps_r, rois, gt.long())
    maps_c = [m.cuda().requires_grad_() for m in maps]
    xin = [m.contiguous(memory_format=torch.channels_last) if fmt == "nhwc" else m for m in maps_c]
    po, mo, fo = dev(xin, rois.cuda(), gt.cuda())
    np.testing.assert_allclose(po.detach().cpu().numpy(), po_r.detach().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(mo.detach().cpu().numpy(), mo_r.detach().numpy(), rtol=1e-4, atol=1e-4)
    for a, b in zip(fo, fo_r):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().numpy(), rtol=1e-3, atol=1e-4)
    ot_ref = pyref.OptTransRef(ch_x=1024, L=5) if loss == "ot" else None
    ml_r = pyref.MetaLossRef(cfg, 1024, ot_loss=ot_ref)
    ot = None
    if loss == "ot":
        ot = fi.OptTrans(cfg, ch_x=1024, L=5)
        ot.load_state_dict(ot_ref.state_dict())
    ml = fi.IntertwinerLoss(cfg, ot_loss=ot).cuda()
    l_r = ml_r([fo_r[0], fo_r[1], fo_r[2], fo_r[3], fo_r[5], fo_r[6]])
    l_c = ml([fo[0], fo[1], fo[2], fo[3], fo[5], fo[6]])
    np.testing.assert_allclose(l_c.detach().cpu().numpy().reshape(-1), l_r.detach().numpy().reshape(-1), atol=LOSS_TOL, rtol=1e-4)
    (l_r.sum() + 1e-3 * po_r.sum() + 1e-3 * mo_r.pow(2).sum()).backward()
    (l_c.sum() + 1e-3 * po.sum() + 1e-3 * mo.pow(2).sum()).backward()
    for a, b in zip(maps_c, maps_r):
        if b.grad is None:
            assert a.grad is None or float(a.grad.abs().max()) == 0
            continue
        np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.numpy(), rtol=1e-3, atol=2e-5)


def test_dev_inference_and_disabled_paths():
    """Dev.forward without class ids (test phase, lib/sub_module.py:593-600,636-638) and with the intertwiner switched off
    (pyramid_roi_align, lib/layers.py:145-218), channels_last and NCHW, vs the restatements."""
    fi = _fi()
    from feature_intertwiner_b200 import synth
    torch.manual_seed(6)
    g = torch.Generator().manual_seed(6)
    shape = (256, 256, 3)
    bs, R, depth = 2, 40, 256
    maps = [torch.randn(bs, depth, 64 >> i, 64 >> i, generator=g) for i in range(4)]
    rois = synth.make_rois(bs, R, (256, 256), g, zero_frac=0.1)
    cfg = pyref.make_config(DATA__IMAGE_SHAPE=np.array(shape))
    ref = pyref.DevRef(cfg, depth=depth).eval()
    dev = fi.Dev(cfg, depth).eval()
    dev.load_state_dict(ref.state_dict())
    dev.cuda()
    with torch.no_grad():
        po_r, mo_r, fo_r = ref(maps, rois, None)
        for fmt in ("nhwc", "nchw"):
            xin = [m.cuda().contiguous(memory_format=torch.channels_last) if fmt == "nhwc" else m.cuda() for m in maps]
            po, mo, fo = dev(xin, rois.cuda(), None)
            assert len(fo) == 2
            np.testing.assert_allclose(po.cpu().numpy(), po_r.numpy(), rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(mo.cpu().numpy(), mo_r.numpy(), rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(fo[0].cpu().numpy(), fo_r[0].numpy(), rtol=1e-3, atol=1e-4)
            np.testing.assert_array_equal(fo[1].cpu().numpy(), fo_r[1].numpy())
        # intertwiner off -> plain pyramid RoIAlign, bit-exact against the C oracle through the restatement
        cfg_off = pyref.make_config(DATA__IMAGE_SHAPE=np.array(shape), DEV__SWITCH=False)
        plain = fi.Dev(cfg_off, depth).cuda()
        for fmt in ("nhwc", "nchw"):
            xin = [m.cuda().contiguous(memory_format=torch.channels_last) if fmt == "nhwc" else m.cuda() for m in maps]
            po, mo, fo = plain(xin, rois.cuda())
            assert fo is None
            np.testing.assert_array_equal(po.cpu().numpy(), pyref.pyramid_roi_align_ref(rois, maps, 7, shape).numpy())
            np.testing.assert_array_equal(mo.cpu().numpy(), pyref.pyramid_roi_align_ref(rois, maps, 14, shape).numpy())


def test_dev_spatial_sort_changes_only_row_order():
    """spatial_sort=True: pooled / mask outputs and the class statistics are unchanged (forward is order independent)."""
    fi = _fi()
    from feature_intertwiner_b200 import synth
    torch.manual_seed(7)
    g = torch.Generator().manual_seed(7)
    cfg = pyref.make_config(DATA__IMAGE_SHAPE=np.array((256, 256, 3)))
    dev = fi.Dev(cfg, 256).eval().cuda()
    maps = [torch.randn(3, 256, 64 >> i, 64 >> i, generator=g).cuda().contiguous(memory_format=torch.channels_last) for i in range(4)]
    rois = synth.make_rois(3, 64, (256, 256), g).cuda()
    gt = synth.make_class_ids(3, 64, g).cuda()
    with torch.no_grad():
        po, mo, fo = dev(maps, rois, gt)
        dev.spatial_sort = True
        po2, mo2, fo2 = dev(maps, rois, gt)
    assert torch.equal(po, po2) and torch.equal(mo, mo2)
    for a, b in zip(fo[:5], fo2[:5]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    # the same (feature row, class id) pairs, level-major in another order
    ka = torch.sort(fo[5].sum(1) * 1000 + fo[6])[0]
    kb = torch.sort(fo2[5].sum(1) * 1000 + fo2[6])[0]
    torch.testing.assert_close(ka, kb, rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------------------------------------ proposal layer
def _proposal_case(seed, bs, A, hw):
    g = torch.Generator().manual_seed(seed)
    H, W = hw
    cy, cx = torch.rand(A, generator=g) * H, torch.rand(A, generator=g) * W
    h = torch.exp(torch.rand(A, generator=g) * 3.0 + 2.5)
    w = h * torch.exp(torch.rand(A, generator=g) * 1.4 - 0.7)
    anchors = torch.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1)
    fg = torch.rand(bs, A, generator=g)
    probs = torch.stack([1 - fg, fg], 2)
    deltas = torch.randn(bs, A, 4, generator=g) * torch.tensor([1.5, 1.5, 2.0, 2.0])
    return probs, deltas, anchors


@pytest.mark.parametrize("bs,A,pre,count", [(2, 3000, 1000, 300), (3, 9000, 6000, 2000), (1, 70, 6000, 50)])
def test_proposal_layer_vs_restatement(bs, A, pre, count):
    """fi.proposal_layer (lib/layers.py:71-139): decoded boxes vs the torch restatement of apply_box_deltas / clip_boxes; the
    final proposals are exactly what the CPU NMS oracle keeps of OUR decoded boxes (so a 1-ulp difference in exp() cannot flip
    a suppression between the two sides), with the reference's truncation to the smallest keep count of the batch."""
    fi = _fi()
    hw = (832, 1344)
    cfg = pyref.make_config(DATA__IMAGE_SHAPE=np.array([hw[0], hw[1], 3]), RPN__PRE_NMS_LIMIT=pre)
    probs, deltas, anchors = _proposal_case(100 + A, bs, A, hw)
    boxes, dets = fi.proposal_decode([probs.cuda(), deltas.cuda()], anchors.cuda(), cfg)
    K = min(pre, A)
    assert tuple(boxes.shape) == (bs, K, 4) and tuple(dets.shape) == (bs, K, 5)
    # decode parity
    scores_s, order = probs[:, :, 1].sort(dim=1, descending=True, stable=True)
    scores_s, order = scores_s[:, :K], order[:, :K]
    std = torch.from_numpy(np.reshape(cfg.DATA.BBOX_STD_DEV, [1, 1, 4])).float()
    d_trim = torch.stack([(deltas * std)[i][order[i]] for i in range(bs)])
    a_trim = torch.stack([anchors[order[i]] for i in range(bs)])
    want = pyref.apply_box_deltas_ref(a_trim, d_trim)
    want = torch.stack([want[:, :, 0].clamp(0.0, hw[0]), want[:, :, 1].clamp(0.0, hw[1]), want[:, :, 2].clamp(0.0, hw[0]), want[:, :, 3].clamp(0.0, hw[1])], 2)
    np.testing.assert_allclose(boxes.cpu().numpy(), want.numpy(), rtol=2e-6, atol=1e-3)
    b = boxes.cpu()
    assert torch.equal(dets.cpu(), torch.cat((b[:, :, [1, 0, 3, 2]], scores_s.unsqueeze(2)), 2))
    # end to end
    got = fi.proposal_layer([probs.cuda(), deltas.cuda()], count, 0.7, anchors.cuda(), cfg).cpu()
    keep = pyref.nms_ref(torch.cat((b, scores_s.unsqueeze(2)), 2), 0.7, strict=True)[:, :count].astype(np.int64)
    exp = torch.stack([b[i][torch.from_numpy(keep[i])] for i in range(bs)]) / torch.tensor([hw[0], hw[1], hw[0], hw[1]], dtype=torch.float32)
    assert tuple(got.shape) == tuple(exp.shape)
    assert torch.equal(got, exp)
    # and the whole restatement agrees whenever no suppression decision sits within rounding of the threshold
    ref, _ = pyref.proposal_layer_ref([probs, deltas], count, 0.7, anchors, cfg)
    if tuple(ref.shape) == tuple(got.shape):
        assert (got - ref).abs().max() < 1e-5 or (got - ref).abs().median() < 1e-6


@pytest.mark.parametrize("scale,hw", [(0.25, (52, 84)), (0.0625, (26, 42)), (0.125, (104, 168))])
def test_roi_pool_vs_reference_cuda_kernel(scale, hw):
    """RoIPool forward (values + argmax) and backward against the REFERENCE'S OWN CUDA kernels (lib/roi_pooling/src/roi_pooling_kernel.cu
    compiled unmodified for sm_100a, oracle/_ref/libref_cuda.so), on level-sized maps, through the reference-named launchers of both
    libraries; the reference's backward is a gather over the RoIs (fixed order), ours sums the same terms: compared to summation tolerance."""
    ref = clib.ref_cuda()
    if ref is None:
        pytest.skip("oracle/_ref/libref_cuda.so did not travel")
    from feature_intertwiner_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(17)
    B, Cc, (H, W), R = 3, 64, hw, 400
    feat = torch.randn(B, Cc, H, W, generator=g).cuda()
    xy = torch.rand(R, 2, generator=g) * torch.tensor([W / scale, H / scale])
    wh = torch.exp(torch.rand(R, 2, generator=g) * 3.5 + 2.0)
    rois = torch.cat([torch.randint(0, B, (R, 1), generator=g).float(), xy - wh / 2, xy + wh / 2], 1)
    rois[0, 1:] = torch.tensor([40., 40., 20., 20.])                    # end < start
    rois[1, 1:] = torch.tensor([1e5, 1e5, 1e5 + 5, 1e5 + 5])            # outside the map: empty bins
    rois = rois.cuda().contiguous()
    s = torch.cuda.current_stream().cuda_stream
    out = {}
    for name, lib in (("ref", ref), ("ours", L)):
        top = torch.full((R, Cc, 7, 7), 3.0, device="cuda")
        arg = torch.full((R, Cc, 7, 7), -7, device="cuda", dtype=torch.int32)
        assert lib.ROIPoolForwardLaucher(feat.data_ptr(), scale, R, H, W, Cc, 7, 7, rois.data_ptr(), top.data_ptr(), arg.data_ptr(), s) == 1
        out[name] = (top, arg)
    torch.cuda.synchronize()
    assert torch.equal(out["ours"][0], out["ref"][0]) and torch.equal(out["ours"][1], out["ref"][1])
    gy = torch.randn(R, Cc, 7, 7, generator=g).cuda()
    grads = {}
    for name, lib in (("ref", ref), ("ours", L)):
        gi = torch.zeros(B, Cc, H, W, device="cuda")
        assert lib.ROIPoolBackwardLaucher(gy.data_ptr(), scale, B, R, H, W, Cc, 7, 7, rois.data_ptr(), gi.data_ptr(), out["ref"][1].data_ptr(), s) == 1
        grads[name] = gi
    torch.cuda.synchronize()
    torch.testing.assert_close(grads["ours"], grads["ref"], rtol=1e-5, atol=1e-5)
    # and the operator layer on top (autograd) gives the same numbers
    fi = _fi()
    fc = feat.clone().requires_grad_()
    o = fi.RoIPoolFunction(7, 7, scale)(fc, rois)
    assert torch.equal(o, out["ref"][0])
    o.backward(gy)
    torch.testing.assert_close(fc.grad, grads["ref"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("scale,hw", [(0.25, (52, 84)), (0.0625, (26, 42))])
def test_roi_pool_nhwc_kernel_equals_nchw_and_reference(scale, hw):
    """channels_last RoIPool (warp per bin, 128-bit loads): values and arg-max identical to the NCHW kernel and to the reference's own
    CUDA kernel; backward equal to the NCHW one to summation tolerance."""
    fi = _fi()
    from feature_intertwiner_b200 import _lib
    g = torch.Generator().manual_seed(23)
    B, Cc, (H, W), R = 2, 256, hw, 300
    feat = torch.randn(B, Cc, H, W, generator=g).cuda()
    xy = torch.rand(R, 2, generator=g) * torch.tensor([W / scale, H / scale])
    wh = torch.exp(torch.rand(R, 2, generator=g) * 3.5 + 2.0)
    rois = torch.cat([torch.randint(0, B, (R, 1), generator=g).float(), xy - wh / 2, xy + wh / 2], 1)
    rois[0, 1:] = torch.tensor([40., 40., 20., 20.])
    rois[1, 1:] = torch.tensor([1e5, 1e5, 1e5 + 5, 1e5 + 5])
    rois = rois.cuda()
    a = feat.clone().requires_grad_()
    b = feat.clone().contiguous(memory_format=torch.channels_last).requires_grad_()
    oa = fi.RoIPoolFunction(7, 7, scale)(a, rois)
    ob = fi.RoIPoolFunction(7, 7, scale)(b, rois)
    assert ob.is_contiguous(memory_format=torch.channels_last) and torch.equal(oa, ob)
    arg_nhwc = ob.grad_fn.saved_tensors[1].contiguous()
    assert torch.equal(oa.grad_fn.saved_tensors[1], arg_nhwc)                                      # arg-max: the same flat NCHW offsets
    gy = torch.randn(oa.shape, generator=g).cuda()
    oa.backward(gy); ob.backward(gy)
    torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-5)
    ref = clib.ref_cuda()
    if ref is not None:
        top = torch.empty(R, Cc, 7, 7, device="cuda"); arg = torch.empty(R, Cc, 7, 7, device="cuda", dtype=torch.int32)
        assert ref.ROIPoolForwardLaucher(feat.data_ptr(), scale, R, H, W, Cc, 7, 7, rois.data_ptr(), top.data_ptr(), arg.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream) == 1
        torch.cuda.synchronize()
        assert torch.equal(ob, top) and torch.equal(arg_nhwc, arg)


def test_proposal_layer_static_has_no_host_read_and_matches():
    """static=True: [bs, proposal_count, 4] zero-padded past the batch's smallest keep count, which stays on the device."""
    fi = _fi()
    hw = (832, 1344)
    cfg = pyref.make_config(DATA__IMAGE_SHAPE=np.array([hw[0], hw[1], 3]), RPN__PRE_NMS_LIMIT=2000)
    probs, deltas, anchors = _proposal_case(321, 3, 5000, hw)
    want = fi.proposal_layer([probs.cuda(), deltas.cuda()], 600, 0.7, anchors.cuda(), cfg)
    rois, m = fi.proposal_layer([probs.cuda(), deltas.cuda()], 600, 0.7, anchors.cuda(), cfg, static=True)
    assert tuple(rois.shape) == (3, 600, 4) and m.dtype == torch.int32 and m.is_cuda
    k = int(m.item())
    assert k == want.size(1) and torch.equal(rois[:, :k], want) and float(rois[:, k:].abs().sum()) == 0.0
