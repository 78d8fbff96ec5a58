"""CPU suite: the hand-written backward of the kernel-fused class-level loss head (feature_intertwiner_b200/intertwiner.py::
_ClassOTHead + dist.py::_MergeStats, kernels in csrc/loss_head.cu) restated step by step in torch and checked against autograd of
the reference composition (lib/model.py:176-207 -> lib/OT_module.py:67-102, 1-D branch, as oracle/pyref.py restates it).

What this pins is the ALGEBRA the kernels implement -- which product feeds which gradient, the transposes, the ReLU masks, the
background column, the all-reduce scale -- not the kernels themselves (tests/test_ops_gpu.py::test_fused_loss_head_matches_torch_ops
does that on the GPU).  With one feature position per row the real Sinkhorn gradient vanishes behind the critic's ReLU, so the
Sinkhorn launch is replaced by a smooth stand-in with dense gradients (same contract: loss[P], dloss/dx, dloss/dy, plan constant).
"""
import torch

EPS = 1e-20


def _surrogate(x, y):
    loss = (x * y).sum(dim=(1, 2)) * 1e-2 + 0.5e-2 * (x * x).sum(dim=(1, 2))
    return loss, (y + x) * 1e-2, x * 1e-2


class _Plan(torch.autograd.Function):
    """loss with the gradients handed back as computed at forward time (no_bp_P_L), like ot.py::_Sinkhorn"""

    @staticmethod
    def forward(ctx, x, y):
        loss, gx, gy = _surrogate(x.detach(), y.detach())
        ctx.save_for_backward(gx, gy)
        return loss

    @staticmethod
    def backward(ctx, g):
        gx, gy = ctx.saved_tensors
        g = g.view(-1, 1, 1)
        return gx * g, gy * g


def test_manual_backward_of_the_fused_head_equals_autograd():
    torch.manual_seed(0)
    G, S, Fd, ncls, N, world = 1, 3, 32, 9, 16, 4.0
    n = ncls - 1
    small_cnt = torch.randint(0, 3, (G, S, 1, ncls)).float()
    small_cnt[..., 3] = 0                                      # a class absent from the batch ...
    small_feat = ((torch.rand(G, S, Fd, ncls) - 0.3) * (small_cnt > 0)).requires_grad_()
    final_big = torch.rand(Fd, ncls)
    buffer_cnt = torch.randint(0, 2, (1, 1, ncls)).float()
    buffer_cnt[0, 0, 1:5] = 1
    buffer_cnt[0, 0, 6] = 0                                    # ... and one absent from the buffer: both masked out
    Wg, bg = torch.randn(Fd, Fd, 3, requires_grad=True), torch.randn(Fd, requires_grad=True)
    Wc, bc = torch.randn(N, Fd, 3, requires_grad=True), torch.randn(N, requires_grad=True)
    up = torch.rand(n) + 0.5

    # ---- the reference composition under autograd (all-reduce of `world` identical ranks' sums, gradient scale = world) ----
    s_sum = (small_feat * small_cnt).sum(dim=(0, 1)) * 1.0
    s_n = small_cnt.sum(dim=(0, 1)).reshape(-1)
    fs = s_sum / (s_n + EPS)
    fg = torch.ones(ncls)
    fg[0] = 0
    mask = ((s_n * fg) > 0) & (buffer_cnt.sum(dim=0).view(-1) > 0)
    X, Y = fs.t()[1:], final_big.t()[1:]
    lin = torch.nn.functional.linear
    H = torch.relu(lin(X, Wg[:, :, 1], bg))
    cx, cy = torch.relu(lin(H, Wc[:, :, 1], bc)), torch.relu(lin(Y, Wc[:, :, 1], bc))
    w = _Plan.apply(torch.cat([cx, cx, cy]).unsqueeze(2), torch.cat([cy, cx, cy]).unsqueeze(2))
    loss = (2 * w[:n] - w[n:2 * n] - w[2 * n:]) * mask[1:].float()
    (loss * up).sum().backward()
    want_small = small_feat.grad * world          # dist._AllReduceSum: backward = identity x world size

    # ---- the kernels' steps, one line each ----
    with torch.no_grad():
        sc = small_cnt.view(G * S, ncls)
        ss = (small_feat.view(G * S, Fd, ncls) * sc.view(G * S, 1, ncls)).sum(0)                       # merge_stats
        sn = sc.sum(0)
        Xm = (ss / (sn + EPS)).t()[1:].contiguous()                                                   # ot_head_prep
        Z = torch.empty(2 * n, Fd)
        Z[n:] = final_big.t()[1:]
        maskf = ((sn[1:] > 0) & (buffer_cnt.view(-1, ncls).sum(0)[1:] > 0)).float()
        Wg1, Wc1 = Wg[:, :, 1].contiguous(), Wc[:, :, 1].contiguous()
        Z[:n] = torch.relu(torch.addmm(bg, Xm, Wg1.t()))
        Cc = torch.relu(torch.addmm(bc, Z, Wc1.t()))
        wm, gx, gy = _surrogate(torch.cat([Cc[:n], Cc[:n], Cc[n:]]).unsqueeze(2), torch.cat([Cc[n:], Cc[:n], Cc[n:]]).unsqueeze(2))
        gx, gy = gx.squeeze(2), gy.squeeze(2)
        lossm = (2 * wm[:n] - wm[n:2 * n] - wm[2 * n:]) * maskf                                       # ot_head_combine
        gm = (up * maskf).view(-1, 1)
        dC = torch.cat([gx[:n] * 2 * gm - gx[n:2 * n] * gm - gy[n:2 * n] * gm,                           # ot_head_dcritic
                        gy[:n] * 2 * gm - gx[2 * n:] * gm - gy[2 * n:] * gm]) * (Cc > 0)
        dbc = dC.sum(0)                                                                                # col_sum
        dWc1 = dC.t().mm(Z)
        dH = dC[:n].mm(Wc1) * (Z[:n] > 0)                                                              # relu_mask
        dbg = dH.sum(0)
        dWg1 = dH.t().mm(Xm)
        dX = dH.mm(Wg1)
        dss = torch.zeros(Fd, ncls)                                                                    # ot_head_dsum
        dss[:, 1:] = (dX / (sn[1:] + EPS).view(-1, 1)).t()
        dsf = (dss * world).view(1, Fd, ncls) * sc.view(G * S, 1, ncls)                                # merge_stats_bwd
        dWg, dWc = torch.zeros_like(Wg), torch.zeros_like(Wc)                                          # centre_tap_embed
        dWg[:, :, 1], dWc[:, :, 1] = dWg1, dWc1

    assert int(mask[1:].sum()) >= 2 and int((~mask[1:]).sum()) >= 1            # both kinds of class are in the case
    torch.testing.assert_close(lossm, loss.detach(), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(dsf.view_as(small_feat), want_small, rtol=1e-5, atol=1e-7)
    assert float(want_small.abs().max()) > 0
    for got, p in ((dWg, Wg), (dbg, bg), (dWc, Wc), (dbc, bc)):
        torch.testing.assert_close(got, p.grad, rtol=1e-5, atol=1e-7)
        assert float(p.grad.abs().max()) > 0
