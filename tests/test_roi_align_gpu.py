"""RoIAlign CUDA path vs the oracle (bit-exact values forward, bit-exact taps; backward to fp32 summation-order
tolerance), through the C ABI, in both memory formats.  Edge cases follow SURVEY.md 8(c)."""
import numpy as np
import pytest
import torch

from oracle import clib

pytestmark = pytest.mark.gpu

BWD_RTOL, BWD_ATOL = 1e-5, 1e-5   # fp32 scatter-add: order of summation differs from the serial CPU loop


def assert_bwd_close(got, grads, rois, box_ind, im_size, want):
    """|got - want| <= 8 ulp-ish of the sum of |contributions| at that pixel: the only legitimate difference between
    two correct fp32 scatter-adds is the order of summation."""
    mag = clib.oracle_crop_and_resize_bwd(np.abs(grads), rois, box_ind, im_size)
    err = np.abs(got - want)
    bound = 1e-6 * mag + 1e-7
    assert np.all(err <= bound), "max err/bound = %g" % float((err / bound).max())


def _fi():
    import feature_intertwiner_b200 as fi
    return fi


def _case(seed, B, C, H, W, R, zero_rows=2):
    from feature_intertwiner_b200 import synth
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(B, C, H, W, generator=g)
    rois = synth.make_rois(1, R, (H * 4, W * 4), g, zero_frac=zero_rows / R, straddle_frac=2.0 / R)[0]
    box_ind = torch.randint(0, B, (R,), generator=g, dtype=torch.int32)
    return image, rois, box_ind


@pytest.mark.parametrize("fmt", ["nchw", "nhwc"])
@pytest.mark.parametrize("P", [1, 7, 14, (3, 5)])
@pytest.mark.parametrize("C", [256, 6])
def test_forward_bit_exact(fmt, P, C):
    fi = _fi()
    ph, pw = (P, P) if isinstance(P, int) else P
    image, rois, box_ind = _case(1, 3, C, 40, 52, 97)
    want = clib.oracle_crop_and_resize_fwd(image.numpy(), rois.numpy(), box_ind.numpy(), ph, pw, 0.5)
    img = image.cuda()
    if fmt == "nhwc":
        img = img.contiguous(memory_format=torch.channels_last)
    got = fi.CropAndResizeFunction(ph, pw, 0.5)(img, rois.cuda(), box_ind.cuda())
    assert got.shape == want.shape
    if fmt == "nhwc" and ph * pw > 1 and C > 1:
        assert got.is_contiguous(memory_format=torch.channels_last)
    np.testing.assert_array_equal(got.cpu().numpy(), want)      # bit-exact


@pytest.mark.parametrize("P", [1, 2, 7, 14])
def test_taps_bit_exact(P):
    fi = _fi()
    _, rois, _ = _case(2, 1, 1, 208, 336, 512, zero_rows=20)
    want = clib.oracle_crop_taps(208, 336, rois.numpy(), P, P)
    got = fi.crop_taps(rois.cuda(), 208, 336, P, P).cpu().numpy()
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("fmt", ["nchw", "nhwc"])
@pytest.mark.parametrize("P", [1, 7, 14])
@pytest.mark.parametrize("C", [256, 6])
def test_backward_matches_oracle(fmt, P, C):
    fi = _fi()
    image, rois, box_ind = _case(3, 2, C, 26, 42, 150, zero_rows=10)
    g = torch.Generator().manual_seed(5)
    grads = torch.randn(150, C, P, P, generator=g)
    want = clib.oracle_crop_and_resize_bwd(grads.numpy(), rois.numpy(), box_ind.numpy(), tuple(image.shape))
    # NCHW with C % 128 == 0 goes through the same tile-owner kernels between two transposes (csrc/roi_align_nchw_bwd.cu)
    for deterministic in ((False, True) if C % 128 == 0 else (False,)):
        img = image.cuda().requires_grad_()
        x = img.contiguous(memory_format=torch.channels_last) if fmt == "nhwc" else img
        old = fi.set_deterministic(deterministic)
        try:
            out = fi.CropAndResizeFunction(P, P)(x, rois.cuda(), box_ind.cuda())
            out.backward(grads.cuda())
        finally:
            fi.set_deterministic(old)
        got = img.grad.cpu().numpy()
        if deterministic:
            # exact tile-owner backward (csrc/roi_align_bwd_tile.cu): same order, same un-fused arithmetic as the serial CPU loop
            np.testing.assert_array_equal(got, want)
        else:
            assert_bwd_close(got, grads.numpy(), rois.numpy(), box_ind.numpy(), tuple(image.shape), want)


def test_golden_vectors(golden_dir):
    """Outputs of the reference's own crop_and_resize.c (tests/golden/make_golden.py)."""
    fi = _fi()
    z = np.load(golden_dir + "/roi_align.npz")
    image, boxes, box_ind = (torch.from_numpy(z[k]).cuda() for k in ("image", "boxes", "box_ind"))
    for fmt in ("nchw", "nhwc"):
        img = image.contiguous(memory_format=torch.channels_last) if fmt == "nhwc" else image
        for P in (1, 2, 7, 14):
            x = img.clone().requires_grad_()
            out = fi.CropAndResizeFunction(P, P, 0.25)(x, boxes, box_ind)
            np.testing.assert_array_equal(out.detach().cpu().numpy(), z[f"crops_{P}"])
            out.backward(torch.from_numpy(z[f"grads_{P}"]).cuda())
            np.testing.assert_allclose(x.grad.cpu().numpy(), z[f"grad_image_{P}"], rtol=BWD_RTOL, atol=BWD_ATOL)
        out = fi.crop_and_resize(img, boxes, box_ind, 3, 5)
        np.testing.assert_array_equal(out.cpu().numpy(), z["crops_3x5"])


def test_backward_multi_sets_bit_exact_and_scatter_fallback(monkeypatch):
    """fi_crop_and_resize_backward_multi: 7x7 + 14x14 crops of one map (+ a second, compact gradient) in one pass."""
    import ctypes
    from feature_intertwiner_b200 import _lib
    image, rois, box_ind = _case(9, 2, 256, 26, 42, 120, zero_rows=6)
    g = torch.Generator().manual_seed(1)
    perm = torch.randperm(120, generator=g).int()
    g7 = torch.randn(120, 256, 7, 7, generator=g)
    g14 = torch.randn(120, 256, 14, 14, generator=g)
    g14b = torch.randn(120, 256, 14, 14, generator=g)
    want = clib.oracle_crop_and_resize_bwd(g7[perm.long()].numpy(), rois.numpy(), box_ind.numpy(), tuple(image.shape))
    # the kernel adds set 0 completely, then set 1, with (grads + grads2) formed first
    want14 = clib.oracle_crop_and_resize_bwd((g14[perm.long()] + g14b).numpy(), rois.numpy(), box_ind.numpy(), tuple(image.shape))
    cl = torch.channels_last
    d = dict(g7=g7.cuda().contiguous(memory_format=cl), g14=g14.cuda().contiguous(memory_format=cl),
             g14b=g14b.cuda().contiguous(memory_format=cl), boxes=rois.cuda(), ind=box_ind.cuda(), perm=perm.cuda())
    sets = (_lib.CropSet * 2)()
    sets[0] = _lib.CropSet(d["g7"].data_ptr(), None, d["boxes"].data_ptr(), d["ind"].data_ptr(), d["perm"].data_ptr(), 120, 7, 7)
    sets[1] = _lib.CropSet(d["g14"].data_ptr(), d["g14b"].data_ptr(), d["boxes"].data_ptr(), d["ind"].data_ptr(), d["perm"].data_ptr(), 120, 14, 14)
    out = torch.full((2, 256, 26, 42), 7.0, device="cuda").contiguous(memory_format=cl)      # garbage: must be overwritten
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().fi_crop_and_resize_backward_multi(sets, 2, 2, 26, 42, 256, out.data_ptr(), 0, 1, s))
    got = out.cpu().numpy()
    mag = clib.oracle_crop_and_resize_bwd(np.abs(g7.numpy()), rois.numpy(), box_ind.numpy(), tuple(image.shape)) + \
        clib.oracle_crop_and_resize_bwd(np.abs(g14.numpy()) + np.abs(g14b.numpy()), rois.numpy(), box_ind.numpy(), tuple(image.shape))
    assert np.all(np.abs(got - (want + want14)) <= 1e-6 * mag + 1e-7)
    # determinism: the same call twice gives the same bits
    out2 = torch.empty_like(out)
    _lib.check(_lib.lib().fi_crop_and_resize_backward_multi(sets, 2, 2, 26, 42, 256, out2.data_ptr(), 0, 1, s))
    assert torch.equal(out, out2)
    # accumulate=1 adds onto the existing map
    _lib.check(_lib.lib().fi_crop_and_resize_backward_multi(sets, 1, 2, 26, 42, 256, out2.data_ptr(), 1, 1, s))
    np.testing.assert_allclose(out2.cpu().numpy(), got + want, rtol=1e-5, atol=1e-5)
    # the default (vector reduction) mode computes the same sums in another order
    out3 = torch.full_like(out, 3.0)
    _lib.check(_lib.lib().fi_crop_and_resize_backward_multi(sets, 2, 2, 26, 42, 256, out3.data_ptr(), 0, 0, s))
    assert np.all(np.abs(out3.cpu().numpy() - (want + want14)) <= 1e-6 * mag + 1e-7)


def test_crop_pair_matches_separate_calls():
    """crop_pair == two crop_and_resize calls + index_copy, forward (bit-exact) and backward."""
    fi = _fi()
    image, rois, box_ind = _case(12, 2, 256, 30, 34, 90, zero_rows=4)
    g = torch.Generator().manual_seed(2)
    rows = torch.randperm(200, generator=g)[:90].int()
    cl = torch.channels_last
    img = image.cuda().contiguous(memory_format=cl)
    xa = img.clone().requires_grad_()
    oa = torch.zeros(200, 256, 7, 7, device="cuda").contiguous(memory_format=cl)
    ob = torch.zeros(200, 256, 14, 14, device="cuda").contiguous(memory_format=cl)
    pa, pb, comp = fi.crop_pair(xa, rois.cuda(), box_ind.cuda(), rows.cuda(), oa, 7, ob, 14, compact_b=True)
    xb = img.clone().requires_grad_()
    ra = fi.crop_and_resize(xb, rois.cuda(), box_ind.cuda(), 7, 7)
    rb = fi.crop_and_resize(xb, rois.cuda(), box_ind.cuda(), 14, 14)
    idx = rows.long().cuda()
    assert torch.equal(pa[idx], ra) and torch.equal(pb[idx], rb) and torch.equal(comp, rb)
    ga, gb, gc = torch.randn(pa.shape, generator=g).cuda(), torch.randn(pb.shape, generator=g).cuda(), torch.randn(comp.shape, generator=g).cuda()
    torch.autograd.backward([pa, pb, comp], [ga, gb, gc])
    torch.autograd.backward([ra, rb], [ga[idx], gb[idx] + gc])
    torch.testing.assert_close(xa.grad, xb.grad, rtol=1e-4, atol=1e-4)


def test_crop_sets_matches_separate_calls():
    """The level-batched launch (fi_crop_sets_forward/backward) == the same crops taken one call at a time."""
    fi = _fi()
    cl = torch.channels_last
    g = torch.Generator().manual_seed(21)
    maps = [torch.randn(2, 256, 30 >> k, 34 >> k, generator=g).cuda().contiguous(memory_format=cl) for k in range(2)]
    cases = []
    for k in range(2):
        _, rois, box_ind = _case(30 + k, 2, 1, 30 >> k, 34 >> k, 40 + 10 * k, zero_rows=3)
        cases.append((rois.cuda(), box_ind.cuda()))
    total = 40 + 50
    rows = torch.randperm(total, generator=g).int().cuda()
    dst = [rows[:40], rows[40:]]
    for deterministic in (False, True):
        old = fi.set_deterministic(deterministic)
        try:
            xa = [m.clone().requires_grad_() for m in maps]
            o7 = torch.empty(total, 256, 7, 7, device="cuda").contiguous(memory_format=cl)
            o14 = torch.empty(total, 256, 14, 14, device="cuda").contiguous(memory_format=cl)
            specs = [dict(image=xa[0], boxes=cases[1][0], box_ind=cases[1][1], size=14)]          # a compact-only "big" set
            for k in range(2):
                specs.append(dict(image=xa[k], boxes=cases[k][0], box_ind=cases[k][1], size=7, out=o7, dst_row=dst[k]))
                specs.append(dict(image=xa[k], boxes=cases[k][0], box_ind=cases[k][1], size=14, out=o14, dst_row=dst[k], compact=(k == 0)))
            outs, comps = fi.crop_sets(specs)
            o7n, o14n = outs[1], outs[2]
            xb = [m.clone().requires_grad_() for m in maps]
            big = fi.crop_and_resize(xb[0], cases[1][0], cases[1][1], 14, 14)
            r7 = [fi.crop_and_resize(xb[k], cases[k][0], cases[k][1], 7, 7) for k in range(2)]
            r14 = [fi.crop_and_resize(xb[k], cases[k][0], cases[k][1], 14, 14) for k in range(2)]
            assert torch.equal(comps[0], big) and torch.equal(comps[2], r14[0]) and comps[1] is None and comps[4] is None
            for k in range(2):
                assert torch.equal(o7n[dst[k].long()], r7[k]) and torch.equal(o14n[dst[k].long()], r14[k])
            gb, g7, g14, gc = (torch.randn(t.shape, generator=g).cuda() for t in (big.cpu(), o7n.cpu(), o14n.cpu(), r14[0].cpu()))
            torch.autograd.backward([comps[0], o7n, o14n, comps[2]], [gb, g7, g14, gc])
            torch.autograd.backward([big] + r7 + r14, [gb, g7[dst[0].long()], g7[dst[1].long()], g14[dst[0].long()] + gc, g14[dst[1].long()]])
            for a, b in zip(xa, xb):
                torch.testing.assert_close(a.grad, b.grad, rtol=1e-4, atol=1e-4)
        finally:
            fi.set_deterministic(old)


@pytest.mark.parametrize("form", [0, 1, 2, 3, 4, 5, 6])
def test_forward_forms_bit_identical(form):
    """Every NHWC forward formulation (fi_set_option fwd_form: 0 default, 1 round-1 unit, 2..6 the lean shapes with packed
    two-float lerps) gives the bits of the reference: single calls on 128 / 256 / 384 / 512 channels, crops 1..33 wide (more than
    one 32-sample pass), extrapolation, out-of-range image indices, and the level-batched launch with scattered rows + compact
    copies + device-side counts."""
    fi = _fi()
    cl = torch.channels_last
    old = fi.set_option("fwd_form", form)
    try:
        for C, (ph, pw), extrap in ((256, (7, 7), 0.0), (256, (14, 14), 0.5), (128, (3, 5), -1.0), (384, (14, 14), 0.0), (512, (2, 33), 0.25),
                                    (256, (1, 1), 0.0)):
            image, rois, box_ind = _case(40 + C + pw, 3, C, 26, 42, 83, zero_rows=5)
            box_ind[3] = -1
            box_ind[7] = 3
            want = clib.oracle_crop_and_resize_fwd(image.numpy(), rois.numpy(), np.clip(box_ind.numpy(), 0, 2), ph, pw, extrap)
            want[3] = 0
            want[7] = 0
            for sched in (1, 2):                                    # static chunks / tickets
                osd = fi.set_option("fwd_sched", sched)
                try:
                    got = fi.CropAndResizeFunction(ph, pw, extrap)(image.cuda().contiguous(memory_format=cl), rois.cuda(), box_ind.cuda())
                finally:
                    fi.set_option("fwd_sched", osd)
                np.testing.assert_array_equal(got.cpu().numpy(), want, err_msg="C=%d crop %dx%d sched %d" % (C, ph, pw, sched))
        # level-batched launch: two maps, 7x7 + 14x14 into shared outputs by dst_row, a compact copy, a device-side count; the 7x7 and
        # 14x14 sets of a map share their box tensors (walked box by box as a pair unless fwd_pair = 1), every chunk size
        g = torch.Generator().manual_seed(77)
        maps = [torch.randn(2, 256, 30 >> k, 34 >> k, generator=g) for k in range(2)]
        cases = [_case(50 + k, 2, 1, 30 >> k, 34 >> k, 40 + 10 * k, zero_rows=3)[1:] for k in range(2)]
        dev_cases = [(c[0].cuda(), c[1].cuda()) for c in cases]
        total = 90
        rows = torch.randperm(total, generator=g).int().cuda()
        dst = [rows[:40], rows[40:]]
        live = [40, 33]                                              # the second list is only partly live
        xs = [m.cuda().contiguous(memory_format=cl) for m in maps]
        cnts = [torch.tensor(live[k], dtype=torch.int32, device="cuda") for k in range(2)]
        for chunk, pair, sched in ((0, 0, 0), (1, 1, 1), (1, 2, 1), (3, 0, 1), (6, 1, 1), (6, 2, 1), (1, 2, 2), (3, 1, 2), (6, 0, 2)):
            oc, op, osd = fi.set_option("fwd_chunk", chunk), fi.set_option("fwd_pair", pair), fi.set_option("fwd_sched", sched)
            try:
                o7 = torch.full((total, 256, 7, 7), -7.0, device="cuda").contiguous(memory_format=cl)
                o14 = torch.full((total, 256, 14, 14), -7.0, device="cuda").contiguous(memory_format=cl)
                specs = []
                for k in range(2):
                    specs.append(dict(image=xs[k], boxes=dev_cases[k][0], box_ind=dev_cases[k][1], size=7, out=o7, dst_row=dst[k], count=cnts[k]))
                    specs.append(dict(image=xs[k], boxes=dev_cases[k][0], box_ind=dev_cases[k][1], size=14, out=o14, dst_row=dst[k],
                                      compact=(k == 0), count=cnts[k]))
                outs, comps = fi.crop_sets(specs)
                for k in range(2):
                    n = live[k]
                    for P, out in ((7, outs[0]), (14, outs[1])):
                        want = clib.oracle_crop_and_resize_fwd(maps[k].numpy(), cases[k][0].numpy()[:n], cases[k][1].numpy()[:n], P, P, 0.0)
                        np.testing.assert_array_equal(out[dst[k][:n].long()].cpu().numpy(), want, err_msg="chunk %d pair %d sched %d" % (chunk, pair, sched))
                        if n < len(dst[k]):                         # rows past the count are not written
                            assert bool((out[dst[k][n:].long()] == -7.0).all())
                want = clib.oracle_crop_and_resize_fwd(maps[0].numpy(), cases[0][0].numpy(), cases[0][1].numpy(), 14, 14, 0.0)
                np.testing.assert_array_equal(comps[1].cpu().numpy(), want)
            finally:
                fi.set_option("fwd_chunk", oc)
                fi.set_option("fwd_pair", op)
                fi.set_option("fwd_sched", osd)
        if form == 0:
            # the ticket counters reset themselves: more launches than there are counter slots, the last one still exact
            for _ in range(300):
                outs, comps = fi.crop_sets(specs)
            want = clib.oracle_crop_and_resize_fwd(maps[0].numpy(), cases[0][0].numpy(), cases[0][1].numpy(), 14, 14, 0.0)
            np.testing.assert_array_equal(comps[1].cpu().numpy(), want)
    finally:
        fi.set_option("fwd_form", old)


def test_known_answers():
    """SURVEY.md 8(c) i-vi: identity crop, constant image, linear ramp, extrapolation, P=1 centre, zero box."""
    fi = _fi()
    H, W = 9, 13
    g = torch.Generator().manual_seed(7)
    image = torch.randn(2, 4, H, W, generator=g).cuda()
    ind0 = torch.zeros(1, dtype=torch.int32).cuda()
    # (i) identity
    out = fi.crop_and_resize(image, torch.tensor([[0., 0., 1., 1.]]).cuda(), ind0, H, W)
    torch.testing.assert_close(out[0], image[0], rtol=0, atol=1e-6)
    # (ii) constant image -> constant crop
    const = torch.full((1, 3, H, W), 2.5).cuda()
    out = fi.crop_and_resize(const, torch.tensor([[0.1, 0.2, 0.7, 0.9]]).cuda(), ind0, 7, 7)
    assert torch.all(out == 2.5)
    # (iii) linear ramp f(y,x) = 2y + 3x -> exact analytic values
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    ramp = (2 * yy + 3 * xx)[None, None].cuda()
    y1, x1, y2, x2 = 0.125, 0.25, 0.75, 0.875
    out = fi.crop_and_resize(ramp, torch.tensor([[y1, x1, y2, x2]]).cuda(), ind0, 5, 5)[0, 0].cpu()
    ys = y1 * (H - 1) + torch.arange(5) * (y2 - y1) * (H - 1) / 4
    xs = x1 * (W - 1) + torch.arange(5) * (x2 - x1) * (W - 1) / 4
    torch.testing.assert_close(out, 2 * ys[:, None] + 3 * xs[None, :], rtol=1e-5, atol=1e-5)
    # (iv) partly outside -> extrapolation value exactly there
    out = fi.crop_and_resize(const, torch.tensor([[-0.5, 0.0, 0.5, 1.0]]).cuda(), ind0, 5, 3, extrapolation_value=-7.0)[0, 0].cpu()
    assert torch.all(out[:2] == -7.0) and torch.all(out[2:] == 2.5)     # rows at y=-4,-2 are outside; 0,2,4 inside
    # (v) P=1 -> centre sample
    out = fi.crop_and_resize(ramp, torch.tensor([[0.25, 0.25, 0.75, 0.75]]).cuda(), ind0, 1, 1)
    assert abs(out.item() - (2 * 0.5 * (H - 1) + 3 * 0.5 * (W - 1))) < 1e-5
    # (vi) all-zero padded box -> every sample is image[b,:,0,0]
    out = fi.crop_and_resize(image, torch.zeros(1, 4).cuda(), torch.ones(1, dtype=torch.int32).cuda(), 7, 7)
    assert torch.all(out[0] == image[1, :, 0, 0][:, None, None])


@pytest.mark.parametrize("fmt", ["nchw", "nhwc"])
def test_backward_is_transpose_of_forward(fmt):
    """SURVEY.md 8(c) vii at a BASELINE-sized level map: <fwd(x), g> == <x, bwd(g)>."""
    fi = _fi()
    from feature_intertwiner_b200 import synth
    g = torch.Generator().manual_seed(11)
    B, C, H, W, R = 2, 256, 104, 168, 1024
    x = torch.randn(B, C, H, W, generator=g).cuda()
    if fmt == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_()
    rois = synth.make_rois(1, R, (832, 1344), g)[0].cuda()
    ind = torch.randint(0, B, (R,), generator=g, dtype=torch.int32).cuda()
    out = fi.crop_and_resize(x, rois, ind, 7, 7)
    gy = torch.randn(out.shape, generator=g).cuda()
    out.backward(gy)
    lhs = (out.detach().double() * gy.double()).sum()
    rhs = (x.detach().double() * x.grad.double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5


def test_bad_box_ind_rows_are_zero_and_dst_row_scatter():
    fi = _fi()
    image, rois, box_ind = _case(4, 2, 8, 20, 20, 16)
    box_ind[3] = 5
    box_ind[7] = -1
    for fmt in ("nchw", "nhwc"):
        img = image.cuda()
        if fmt == "nhwc":
            img = img.contiguous(memory_format=torch.channels_last)
        out = fi.crop_and_resize(img, rois.cuda(), box_ind.cuda(), 7, 7)
        assert torch.all(out[3] == 0) and torch.all(out[7] == 0)       # crop_and_resize_kernel.cu:34-38
        keep = [i for i in range(16) if i not in (3, 7)]
        want = clib.oracle_crop_and_resize_fwd(image.numpy(), rois.numpy()[keep], box_ind.numpy()[keep], 7, 7)
        np.testing.assert_array_equal(out[keep].cpu().numpy(), want)
        # fused scatter-back (lib/sub_module.py:645-662)
        perm = torch.randperm(16, generator=torch.Generator().manual_seed(0)).int()
        dest = torch.zeros(16, 8, 7, 7, device="cuda").contiguous(memory_format=torch.channels_last if fmt == "nhwc" else torch.contiguous_format)
        fi.crop_and_resize(img, rois.cuda(), box_ind.cuda(), 7, 7, out=dest, dst_row=perm.cuda())
        torch.testing.assert_close(dest[perm.long().cuda()], out, rtol=0, atol=0)


def test_reference_named_launchers():
    """The exact C symbols of crop_and_resize_kernel.h:8-18, raw pointers, caller-zeroed grads."""
    from feature_intertwiner_b200 import _lib
    image, rois, box_ind = _case(6, 2, 16, 30, 34, 40)
    img, b, bi = image.cuda(), rois.cuda(), box_ind.cuda()
    crops = torch.empty(40, 16, 7, 7, device="cuda")
    L = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream
    L.CropAndResizeLaucher(img.data_ptr(), b.data_ptr(), bi.data_ptr(), 40, 2, 30, 34, 7, 7, 16, 0.0, crops.data_ptr(), s)
    assert L.fi_last_status() == 0
    want = clib.oracle_crop_and_resize_fwd(image.numpy(), rois.numpy(), box_ind.numpy(), 7, 7)
    np.testing.assert_array_equal(crops.cpu().numpy(), want)
    g = torch.randn(40, 16, 7, 7, device="cuda")
    gi = torch.zeros_like(img)
    L.CropAndResizeBackpropImageLaucher(g.data_ptr(), b.data_ptr(), bi.data_ptr(), 40, 2, 30, 34, 7, 7, 16, gi.data_ptr(), s)
    want = clib.oracle_crop_and_resize_bwd(g.cpu().numpy(), rois.numpy(), box_ind.numpy(), tuple(image.shape))
    np.testing.assert_allclose(gi.cpu().numpy(), want, rtol=BWD_RTOL, atol=BWD_ATOL)


def test_refuses_cpu_tensors():
    fi = _fi()
    with pytest.raises(fi.FiError):
        fi.crop_and_resize(torch.randn(1, 4, 8, 8), torch.zeros(1, 4), torch.zeros(1, dtype=torch.int32), 7, 7)


def test_full_size_properties_c5():
    """BASELINE.json config 5 sizes (batch 4, 2000 RoIs/img, P2 208x336): size-independent properties where the oracle would
    take minutes -- <fwd(x),g> == <x,bwd(g)>, constant map -> constant crop, determinism of the deterministic mode."""
    fi = _fi()
    from feature_intertwiner_b200 import synth
    g = torch.Generator().manual_seed(55)
    B, R = 4, 2000
    rois = synth.make_rois(B, R, (832, 1344), g)
    lvl = fi.roi_level(rois.cuda(), (832, 1344, 3))
    sp = fi.split_levels(lvl, rois=rois.cuda())
    boxes, ind = sp.small_boxes(0), sp.small_ind(0)
    x = torch.randn(B, 256, 208, 336, generator=g).cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    out = fi.crop_and_resize(x, boxes, ind, 14, 14)
    gy = torch.randn(out.shape, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    out.backward(gy)
    lhs = (out.detach().double() * gy.double()).sum()
    rhs = (x.detach().double() * x.grad.double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5
    const = torch.full((B, 256, 208, 336), 1.5, device="cuda").contiguous(memory_format=torch.channels_last)
    inside = (boxes.min(dim=1)[0] >= 0) & (boxes.max(dim=1)[0] <= 1)
    assert torch.all(fi.crop_and_resize(const, boxes, ind, 7, 7)[inside] == 1.5)
    old = fi.set_deterministic(True)
    try:
        grads = []
        for _ in range(2):
            x.grad = None
            fi.crop_and_resize(x, boxes[:300], ind[:300], 7, 7).backward(gy[:300, :, :7, :7].contiguous(memory_format=torch.channels_last))
            grads.append(x.grad.clone())
        assert torch.equal(grads[0], grads[1])
    finally:
        fi.set_deterministic(old)


@pytest.mark.parametrize("mode", ["pix", "pix_exact", "smem", "smem_exact", "fused", "fused_exact", "red", "pix16", "pix_g8_exact"])
@pytest.mark.parametrize("shape", [(2, 256, 26, 42, 150), (3, 128, 9, 11, 40), (1, 384, 33, 70, 300)])
def test_backward_formulations_agree(mode, shape):
    """Every formulation of the NHWC backward on the same maps -- tile-owner lists + the bulk-copy staged register-accumulating
    kernel (csrc/roi_align_bwd_pix.cu, default; also with 16-slot batches / 8-tile tickets), + the shared-memory accumulate
    kernel, the fused single kernel (csrc/roi_align_bwd_tile.cu), each with default and exact arithmetic; vector reductions --
    over ragged tile edges (H, W not multiples of 4 / 8), 128 / 256 / 384 channels, many boxes per tile (> one 64-entry list
    chunk), zero-padded / inverted / outside boxes.  The exact modes are bit-identical to the serial CPU reference."""
    fi = _fi()
    B, C, H, W, R = shape
    image, rois, box_ind = _case(77, B, C, H, W, R, zero_rows=8)
    rois[3] = torch.tensor([0.9, 0.8, 0.1, 0.2])          # inverted
    rois[4] = torch.tensor([-0.5, -0.5, 1.5, 1.5])        # mostly outside
    rois[5] = torch.tensor([0.5, 0.25, 0.5, 0.25])        # degenerate, on one (fractional) position
    g = torch.Generator().manual_seed(6)
    form = mode.split("_")[0]
    exact = mode.endswith("_exact")
    saved = {k: fi.get_option(k) for k in ("bwd_form", "pix_cfg", "pix_group")}
    for P in (7, 14, 16):
        grads = torch.randn(R, C, P, P, generator=g)
        want = clib.oracle_crop_and_resize_bwd(grads.numpy(), rois.numpy(), box_ind.numpy(), tuple(image.shape))
        img = image.cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
        fi.set_option("bwd_form", "pix" if form.startswith("pix") else form)
        fi.set_option("pix_cfg", 2 if form == "pix16" else 0)
        fi.set_option("pix_group", 7 if "g8" in mode else 1)
        old = fi.set_deterministic(exact)                   # the process-wide switch selects the exact arithmetic
        try:
            fi.crop_and_resize(img, rois.cuda(), box_ind.cuda(), P, P).backward(grads.cuda().contiguous(memory_format=torch.channels_last))
        finally:
            fi.set_deterministic(old)
            for k, v in saved.items():
                fi.set_option(k, v)
        got = img.grad.cpu().numpy()
        if exact:
            np.testing.assert_array_equal(got, want)
        else:
            assert_bwd_close(got, grads.numpy(), rois.numpy(), box_ind.numpy(), tuple(image.shape), want)


def test_tile_backward_full_size_c2_properties():
    """C2 sizes through the level-batched entry: transpose property per map and exact == default within summation tolerance,
    run-to-run identical bits in both modes (the tile-owner kernel has a fixed summation order)."""
    fi = _fi()
    from feature_intertwiner_b200 import synth
    g = torch.Generator().manual_seed(56)
    B, R = 8, 512
    hw = (832, 1344)
    rois = synth.make_rois(B, R, hw, g).cuda()
    sp = fi.split_levels(fi.roi_level(rois, (hw[0], hw[1], 3)), rois=rois)
    cl = torch.channels_last
    maps = [torch.randn(B, 256, hw[0] >> (2 + i), hw[1] >> (2 + i), generator=g).cuda().contiguous(memory_format=cl) for i in range(2)]
    res = {}
    for mode in ("tile", "tile2", "exact", "exact2"):
        xs = [m.clone().requires_grad_() for m in maps]
        specs = []
        for i in range(2):
            specs.append(dict(image=xs[i], boxes=sp.small_boxes(i), box_ind=sp.small_ind(i), size=7))
            specs.append(dict(image=xs[i], boxes=sp.small_boxes(i), box_ind=sp.small_ind(i), size=14))
            specs.append(dict(image=xs[i], boxes=sp.big_boxes(i), box_ind=sp.big_ind(i), size=14))
        _, comps = fi.crop_sets(specs)
        gg = torch.Generator().manual_seed(9)
        gys = [torch.randn(c.shape, generator=gg).cuda().contiguous(memory_format=cl) for c in comps]
        old = fi.set_deterministic(mode.startswith("exact"))
        try:
            torch.autograd.backward(comps, gys)
        finally:
            fi.set_deterministic(old)
        res[mode] = [x.grad for x in xs]
        if mode == "tile":
            for i in range(2):
                lhs = sum((comps[3 * i + k].detach().double() * gys[3 * i + k].double()).sum() for k in range(3))
                rhs = (xs[i].detach().double() * xs[i].grad.double()).sum()
                assert abs(lhs - rhs) / abs(lhs) < 1e-5
    for i in range(2):
        assert torch.equal(res["tile"][i], res["tile2"][i]) and torch.equal(res["exact"][i], res["exact2"][i])
        torch.testing.assert_close(res["tile"][i], res["exact"][i], rtol=1e-4, atol=1e-4)


def _sets_case(seed=31, B=2, R=180, hw=(26, 42)):
    from feature_intertwiner_b200 import synth
    g = torch.Generator().manual_seed(seed)
    cl = torch.channels_last
    maps = [torch.randn(B, 256, hw[0] * s, hw[1] * s, generator=g).cuda().contiguous(memory_format=cl) for s in (2, 1)]
    rois = synth.make_rois(1, R, (hw[0] * 32, hw[1] * 32), g, zero_frac=0.05, straddle_frac=0.03)[0].cuda()
    ind = torch.randint(0, B, (R,), generator=g, dtype=torch.int32).cuda()
    rows = torch.randperm(R + 20, generator=g)[:R].int().cuda()
    return maps, rois, ind, rows


def _run_sets(fi, maps, rois, ind, rows, n_live, capacity, use_counts, max_entries=0, grads=None, cnt=None):
    """Two maps, three sets (7x7 scattered, 14x14 scattered + compact = two-source, 14x14 compact on the second map)."""
    cl = torch.channels_last
    xs = [m.clone().requires_grad_() for m in maps]
    R = rois.size(0)
    pad = lambda t: torch.cat([t[:n_live], t.new_zeros((capacity - n_live,) + tuple(t.shape[1:]))]) if capacity > n_live else t[:n_live]
    boxes, bi, rw = pad(rois), pad(ind), pad(rows)
    if capacity > n_live:                               # garbage past the count must be ignored
        boxes[n_live:] = 0.37; bi[n_live:] = 99; rw[n_live:] = 0
    if use_counts and cnt is None:
        cnt = torch.tensor([n_live], dtype=torch.int32, device="cuda")
    o7 = torch.zeros(R + 20, 256, 7, 7, device="cuda").contiguous(memory_format=cl)
    o14 = torch.zeros(R + 20, 256, 14, 14, device="cuda").contiguous(memory_format=cl)
    specs = [dict(image=xs[0], boxes=boxes, box_ind=bi, size=7, out=o7, dst_row=rw, count=cnt),
             dict(image=xs[0], boxes=boxes, box_ind=bi, size=14, out=o14, dst_row=rw, compact=True, count=cnt),
             dict(image=xs[1], boxes=boxes, box_ind=bi, size=14, count=cnt)]
    outs, comps = fi.crop_sets(specs, max_entries=max_entries)
    heads = [outs[0], outs[1], comps[1], comps[2]]
    if grads is None:
        g = torch.Generator().manual_seed(5)
        grads = [torch.randn(h.shape, generator=g).cuda().contiguous(memory_format=cl) for h in heads]
        for gg in grads[2:]:
            gg[n_live:] = float("nan")                  # rows past the count are never read
    else:
        grads = [gr[: h.size(0)] if gr.size(0) >= h.size(0) else torch.cat([gr, gr.new_full((h.size(0) - gr.size(0),) + tuple(gr.shape[1:]), float("nan"))])
                 for gr, h in zip(grads, heads)]
        grads = [gr.contiguous(memory_format=cl) for gr in grads]
    return xs, heads, grads


@pytest.mark.parametrize("exact", [False, True])
def test_crop_sets_device_counts_plan_reuse_and_side_stream(exact):
    """Lists with a capacity and a device-side count give the same bits as exact-size lists (forward and backward); a plan built
    at forward time on the side stream == lists built inside backward; running the same plan twice gives the same bits."""
    fi = _fi()
    maps, rois, ind, rows = _sets_case()
    n = 150
    old = fi.set_deterministic(exact)
    try:
        xs_a, heads_a, grads = _run_sets(fi, maps, rois, ind, rows, n, n, use_counts=False)
        ga = torch.autograd.grad(heads_a, xs_a, grads, retain_graph=True)
        ga2 = torch.autograd.grad(heads_a, xs_a, grads)                       # the same plan, run again
        xs_b, heads_b, grads_b = _run_sets(fi, maps, rois, ind, rows, n, 180, use_counts=True, grads=grads)
        gb = torch.autograd.grad(heads_b, xs_b, grads_b)
        prev = fi.roi_align.plan_at_forward(False)
        try:
            xs_c, heads_c, _ = _run_sets(fi, maps, rois, ind, rows, n, n, use_counts=False)
            gc = torch.autograd.grad(heads_c, xs_c, grads)
        finally:
            fi.roi_align.plan_at_forward(prev["at_forward"], prev["side_stream"])
    finally:
        fi.set_deterministic(old)
    for a, b in zip(heads_a[2:], heads_b[2:]):
        assert torch.equal(a, b[:n])                                          # compact crops: live rows identical
    assert torch.equal(heads_a[0], heads_b[0]) and torch.equal(heads_a[1], heads_b[1])
    for a, a2, b, c in zip(ga, ga2, gb, gc):
        assert torch.equal(a, a2) and torch.equal(a, b) and torch.equal(a, c)
        assert torch.isfinite(a).all()


def test_crop_sets_backward_overflow_flag_and_graph_capture():
    """A caller's own (too small) entry bound is detected, never written out of bounds; the plan / run pair on a torch-allocated
    workspace can be captured into a CUDA graph and replays to the eager result."""
    fi = _fi()
    from feature_intertwiner_b200 import roi_align
    maps, rois, ind, rows = _sets_case(seed=32)
    xs, heads, grads = _run_sets(fi, maps, rois, ind, rows, 180, 180, use_counts=False, max_entries=4096)
    node = heads[0].grad_fn
    want_xs, want_heads, _ = _run_sets(fi, maps, rois, ind, rows, 180, 180, use_counts=False)
    want = torch.autograd.grad(want_heads, want_xs, grads)
    torch.autograd.grad(heads, xs, grads)                                      # runs, truncated lists, no crash
    torch.cuda.synchronize()
    assert node.bwd.overflowed(maps[0].device)
    assert not want_heads[0].grad_fn.bwd.overflowed(maps[0].device)
    # ---- capture forward (+ plan on the side stream) + backward, replay twice
    old = roi_align.plan_at_forward(True, side_stream=True)
    cnt = torch.tensor([180], dtype=torch.int32, device="cuda")                # made outside the capture (host -> device copy)
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):                                                 # warm-up on the capture stream (allocator, attributes)
                a, b, c = _run_sets(fi, maps, rois, ind, rows, 180, 180, use_counts=True, grads=grads, cnt=cnt)
                torch.autograd.grad(b, a, c)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            a, b, c = _run_sets(fi, maps, rois, ind, rows, 180, 180, use_counts=True, grads=grads, cnt=cnt)
            got = torch.autograd.grad(b, a, c)
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
    finally:
        roi_align.plan_at_forward(old["at_forward"], old["side_stream"])
    for g, w in zip(got, want):
        assert torch.equal(g, w)


@pytest.mark.parametrize("wl_name", ["c2", "c5"])
def test_full_size_vs_compiled_reference(wl_name):
    """BASELINE.json sizes (C2: batch 8 x 512 RoIs; C5: batch 4 x 2000 RoIs per GPU, 832x1344, P2 208x336 ... P5 26x42, 256
    channels): forward AND exact-mode backward of every crop set of a Dev.forward pass, through the level-batched entry
    (fi.crop_sets: heaviest-first tile order, collapse-free exact lists, > 64-entry chunk chains, two-source sets), against the
    REFERENCE'S OWN C (lib/roi_align/src/crop_and_resize.c compiled unmodified, oracle/_ref; the restatement that is pinned
    bit-identical to it where oracle/_ref did not travel).  Bit-exact both ways.  The exact backward equals the serial CPU loop
    for one set per map, so the sets are run family by family (big 14x14 / small 7x7 / small 14x14 with its two gradient
    sources pre-added as autograd does in the reference); the default-mode pass over all sets at once must agree with their
    sum within the summation-order bound."""
    fi = _fi()
    from feature_intertwiner_b200 import synth
    fwd = clib.ref_crop_and_resize_fwd if clib.have_ref() else clib.oracle_crop_and_resize_fwd
    bwd = clib.ref_crop_and_resize_bwd if clib.have_ref() else clib.oracle_crop_and_resize_bwd
    wl = synth.WORKLOADS[wl_name]
    B, R, hw = wl["batch"], wl["rois_per_image"], wl["image"]
    g = torch.Generator().manual_seed(2000)
    cl = torch.channels_last
    rois = synth.make_rois(B, R, hw, g).cuda()
    raw = [m.cuda() for m in synth.make_feature_maps(B, hw, 256, g, channels_last=True)]
    madeup = [m.cuda() for m in synth.make_feature_maps(B, hw, 256, g, channels_last=True)]
    sp = fi.split_levels(fi.roi_level(rois, (hw[0], hw[1], 3)), rois=rois, order=fi.spatial_order(rois))
    total = B * R
    nchw = lambda t: t.detach().contiguous().cpu().numpy()
    gg = torch.Generator(device="cuda").manual_seed(11)

    def family(name):
        """specs of one family on fresh leaves + (leaf, boxes, ind, rows, size, dual) per set"""
        xs_raw = [m.clone().requires_grad_() for m in raw]
        xs_mu = [m.clone().requires_grad_() for m in madeup]
        pooled = torch.zeros((total, 256, 7, 7), device="cuda").contiguous(memory_format=cl)
        mask = torch.zeros((total, 256, 14, 14), device="cuda").contiguous(memory_format=cl)
        specs, meta = [], []
        for i in range(4):
            if sp.small_cnt[i] == 0:
                continue
            if name in ("big", "all") and i < 3 and sp.big_cnt[i]:
                specs.append(dict(image=xs_raw[i], boxes=sp.big_boxes(i), box_ind=sp.big_ind(i), size=14))
                meta.append((xs_raw[i], sp.big_boxes(i), sp.big_ind(i), None, 14, False))
            if name in ("s7", "all"):
                specs.append(dict(image=xs_mu[i], boxes=sp.small_boxes(i), box_ind=sp.small_ind(i), size=7, out=pooled, dst_row=sp.small(i)))
                meta.append((xs_mu[i], sp.small_boxes(i), sp.small_ind(i), sp.small(i), 7, False))
            if name in ("s14", "all"):
                specs.append(dict(image=xs_mu[i], boxes=sp.small_boxes(i), box_ind=sp.small_ind(i), size=14, out=mask, dst_row=sp.small(i), compact=(i < 3)))
                meta.append((xs_mu[i], sp.small_boxes(i), sp.small_ind(i), sp.small(i), 14, i < 3))
        return specs, meta

    def run(name, exact, grads_by_set=None):
        specs, meta = family(name)
        outs, comps = fi.crop_sets(specs)
        heads, gr, per_set = [], [], []
        seen = {}
        for k, (leaf, boxes, ind, rows, P, dual) in enumerate(meta):
            n = boxes.size(0)
            if rows is None:
                h = comps[k]
                gk = (torch.randn((n, 256, P, P), device="cuda", generator=gg).contiguous(memory_format=cl),) if grads_by_set is None else grads_by_set[k]
                heads.append(h); gr.append(gk[0]); per_set.append(gk)
                continue
            o = outs[k]
            if id(o) not in seen:                         # one gradient tensor per shared output; rows of other levels stay zero here
                seen[id(o)] = torch.zeros(o.shape, device="cuda").contiguous(memory_format=cl)
                heads.append(o); gr.append(seen[id(o)])
            gfull = seen[id(o)]
            if grads_by_set is None:
                g1 = torch.randn((n, 256, P, P), device="cuda", generator=gg).contiguous(memory_format=cl)
                gk = (g1, torch.randn((n, 256, P, P), device="cuda", generator=gg).contiguous(memory_format=cl)) if dual else (g1,)
            else:
                gk = grads_by_set[k]
            gfull[rows.long()] = gk[0]
            if dual:
                heads.append(comps[k]); gr.append(gk[1])
            per_set.append(gk)
        old = fi.set_deterministic(exact)
        try:
            leaves = []
            for m in meta:
                if not any(m[0] is l for l in leaves):
                    leaves.append(m[0])
            res = torch.autograd.grad(heads, leaves, gr)
        finally:
            fi.set_deterministic(old)
        crops = [(outs[k][meta[k][3].long()] if meta[k][3] is not None else comps[k]) for k in range(len(meta))]
        return meta, crops, per_set, leaves, res

    sums, mags = {}, {}
    all_grads = {}
    for name in ("big", "s7", "s14"):
        meta, crops, per_set, leaves, res = run(name, exact=True)
        all_grads[name] = per_set
        for k, (leaf, boxes, ind, rows, P, dual) in enumerate(meta):
            img = nchw(leaf)
            want = fwd(img, boxes.cpu().numpy(), ind.cpu().numpy(), P, P, 0.0)
            np.testing.assert_array_equal(nchw(crops[k]), want)                               # forward: bit-exact
            gsum = per_set[k][0] + per_set[k][1] if dual else per_set[k][0]                   # autograd's (g1 + g2) in the reference
            want_g = bwd(nchw(gsum), boxes.cpu().numpy(), ind.cpu().numpy(), tuple(leaf.shape))
            li = [i for i, l in enumerate(leaves) if l is leaf][0]
            np.testing.assert_array_equal(nchw(res[li]), want_g)                              # exact backward: bit-exact
            key = (name != "big", [tuple(m.shape) for m in madeup].index(tuple(leaf.shape)) if name != "big" else [tuple(m.shape) for m in raw].index(tuple(leaf.shape)))
            sums[key] = sums.get(key, 0) + want_g.astype(np.float64)
            mags[key] = mags.get(key, 0) + bwd(np.abs(nchw(gsum)), boxes.cpu().numpy(), ind.cpu().numpy(), tuple(leaf.shape)).astype(np.float64)
    # ---- all sets in one pass, default arithmetic, same gradients
    specs_all, meta_all = family("all")
    order = []
    cursor = {"big": 0, "s7": 0, "s14": 0}
    for (leaf, boxes, ind, rows, P, dual) in meta_all:
        fam = "big" if rows is None else ("s7" if P == 7 else "s14")
        order.append(all_grads[fam][cursor[fam]]); cursor[fam] += 1
    meta, crops, per_set, leaves, res = run("all", exact=False, grads_by_set=order)
    shapes_raw, shapes_mu = [tuple(m.shape) for m in raw], [tuple(m.shape) for m in madeup]
    for leaf, got in zip(leaves, res):
        is_mu = any(leaf is m[0] and m[3] is not None for m in meta)
        key = (is_mu, (shapes_mu if is_mu else shapes_raw).index(tuple(leaf.shape)))
        err = np.abs(nchw(got).astype(np.float64) - sums[key])
        assert np.all(err <= 2e-6 * mags[key] + 1e-6), "default-mode pass: max err/bound = %g" % float((err / (2e-6 * mags[key] + 1e-6)).max())


def test_reference_named_backward_launcher_c256_accumulates_and_matches_reference_kernel():
    """CropAndResizeBackpropImageLaucher at the model's channel count (NCHW, C = 256: transposed through the tile-owner kernels):
    ADDS onto the caller's map like the reference kernel (crop_and_resize_kernel.cu:84-165), compared with that kernel itself
    (oracle/_ref/libref_cuda.so, atomics: summation-order tolerance) and with the CPU oracle."""
    from feature_intertwiner_b200 import _lib
    image, rois, box_ind = _case(8, 3, 256, 52, 84, 300, zero_rows=12)
    b, bi = rois.cuda(), box_ind.cuda()
    L = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(4)
    for P in (7, 14):
        grads = torch.randn(300, 256, P, P, generator=g)
        base = torch.randn(3, 256, 52, 84, generator=g)
        gi = base.clone().cuda()
        L.CropAndResizeBackpropImageLaucher(grads.cuda().data_ptr(), b.data_ptr(), bi.data_ptr(), 300, 3, 52, 84, P, P, 256, gi.data_ptr(), s)
        assert L.fi_last_status() == 0, L.fi_last_error()
        want = clib.oracle_crop_and_resize_bwd(grads.numpy(), rois.numpy(), box_ind.numpy(), (3, 256, 52, 84))
        # summation-order bound on the scatter sum + one rounding of (base + sum)
        mag = clib.oracle_crop_and_resize_bwd(np.abs(grads.numpy()), rois.numpy(), box_ind.numpy(), (3, 256, 52, 84))
        total = base.numpy() + want
        bound = 2e-6 * mag + 2e-6 * np.abs(total) + 1e-6
        assert np.all(np.abs(gi.cpu().numpy() - total) <= bound)
        ref = clib.ref_cuda()
        if ref is not None:
            gr = base.clone().cuda()
            ref.CropAndResizeBackpropImageLaucher(grads.cuda().data_ptr(), b.data_ptr(), bi.data_ptr(), 300, 3, 52, 84, P, P, 256, gr.data_ptr(), s)
            torch.cuda.synchronize()
            # the reference kernel adds every term onto (base + partial sum) with atomics, in arrival order: each of its hundreds of
            # additions rounds at the magnitude of the running value -- a sanity cross-check, the parity bound is the one above
            torch.testing.assert_close(gi, gr, rtol=1e-4, atol=2e-4)


def test_in_place_out_gradient_is_zero_on_overwritten_rows():
    """An `out` that carries a gradient: the rows a crop call overwrites no longer depend on its previous contents."""
    fi = _fi()
    image, rois, box_ind = _case(21, 2, 256, 26, 42, 40, zero_rows=2)
    cl = torch.channels_last
    img = image.cuda().contiguous(memory_format=cl).requires_grad_()
    base = torch.randn(60, 256, 7, 7, device="cuda").contiguous(memory_format=cl).requires_grad_()
    rows = torch.randperm(60)[:40].int().cuda()
    prev = base * 2.0                                       # has a grad_fn
    out = fi.crop_and_resize(img, rois.cuda(), box_ind.cuda(), 7, 7, out=prev, dst_row=rows)
    out.sum().backward()
    untouched = torch.ones(60, dtype=torch.bool)
    untouched[rows.cpu().long()] = False
    assert torch.all(base.grad[untouched] == 2.0) and torch.all(base.grad[~untouched] == 0.0)
    # the level-batched entry
    img2 = image.cuda().contiguous(memory_format=cl).requires_grad_()
    base2 = torch.randn(60, 256, 7, 7, device="cuda").contiguous(memory_format=cl).requires_grad_()
    outs, _ = fi.crop_sets([dict(image=img2, boxes=rois.cuda(), box_ind=box_ind.cuda(), size=7, out=base2 * 2.0, dst_row=rows)])
    outs[0].sum().backward()
    assert torch.all(base2.grad[untouched] == 2.0) and torch.all(base2.grad[~untouched] == 0.0)
    torch.testing.assert_close(img.grad, img2.grad)
