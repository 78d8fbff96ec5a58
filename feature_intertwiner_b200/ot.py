"""Optimal-transport (Sinkhorn) loss -- lib/OT_module.py.

``OptTrans`` keeps the reference's constructor, parameter names (``G_net.*``, ``critic.*`` -- a reference
state_dict loads unchanged) and return value (a per-sample loss VECTOR).  The python
double loop ``for i in range(bs): for _ in range(L):`` of OT_module.py:100-122 is one kernel launch for all
3*bs problems of the debiased loss 2W(x^,y) - W(x^,x^) - W(y,y) (csrc/sinkhorn.cu).
"""
import torch
import torch.nn as nn

from . import _lib


def sinkhorn_raw(x, y, inv_eps, L, need_grad):
    """One batched launch (plus the large-D gradient launches) of csrc/sinkhorn.cu on [P,N,D] problems: (loss[P], dloss/dx, dloss/dy)
    with the transport plan held constant (no_bp_P_L); the gradients are None unless ``need_grad``.  No autograd."""
    if x.shape != y.shape or x.dim() != 3:
        raise _lib.FiError("sinkhorn: x and y must both be [P,N,D]; got %s / %s" % (tuple(x.shape), tuple(y.shape)))
    _lib.require_cuda(x, y)
    x = x.detach().float().contiguous()
    y = y.detach().float().contiguous()
    P, N, D = x.shape
    loss = torch.empty((P,), device=x.device, dtype=torch.float32)
    gx = torch.empty_like(x) if need_grad else None
    gy = torch.empty_like(y) if need_grad else None
    L_ = _lib.lib()
    nbytes = L_.fi_sinkhorn_workspace(P, N, D, 1 if need_grad else 0)          # large D: the gradient runs as its own launches
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device) if nbytes else None
    with torch.cuda.device(x.device):
        _lib.check(L_.fi_sinkhorn_ws(_lib.ptr(x), _lib.ptr(y), P, N, D, float(inv_eps), int(L), _lib.ptr(loss),
                                     _lib.ptr(gx), _lib.ptr(gy), _lib.ptr(ws), nbytes, _lib.stream_ptr(x.device)))
    return loss, gx, gy


class _Sinkhorn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, inv_eps, L):
        need_grad = any(ctx.needs_input_grad[:2])
        loss, gx, gy = sinkhorn_raw(x, y, inv_eps, L, need_grad)
        if need_grad:
            ctx.save_for_backward(gx, gy)
        return loss

    @staticmethod
    def backward(ctx, g):
        gx, gy = ctx.saved_tensors
        g = g.view(-1, 1, 1)
        return (gx * g if ctx.needs_input_grad[0] else None), (gy * g if ctx.needs_input_grad[1] else None), None, None


def sinkhorn_loss(x, y, epsilon=1.0, L=5):
    """x,y [P,N,D]: P independent problems, rows = samples (critic channels), cols = feature positions.
    Returns loss[P] = <P*,C> of OT_module.py:104-135 (cosine cost, transport plan detached)."""
    return _Sinkhorn.apply(x, y, 1.0 / float(epsilon), int(L))


class OptTrans(nn.Module):
    def __init__(self, config, ch_x, spatial_x=-1, ch_y=-1, spatial_y=-1, epsilon=1., L=5, remove_bias=False,
                 C_form='cosine', no_bp_P_L=True, skip_critic=False):
        super().__init__()
        if C_form != 'cosine':
            raise _lib.FiError("OptTrans: only C_form='cosine' is built (no caller of the reference selects 'l2', OT_module.py:106-109)")
        if not no_bp_P_L:
            raise _lib.FiError("OptTrans: no_bp_P_L=False (back-prop through the Sinkhorn iterations) is not built")
        self.config = config
        self.epsilon = 1. / epsilon                 # sic: the reference stores the reciprocal (OT_module.py:13)
        self.L = L
        self.remove_bias = remove_bias
        self.no_bp_P_L = no_bp_P_L
        self.C_form = C_form
        self.skip_critic = skip_critic
        self.two_dim = spatial_x > 1
        ch_y = ch_x if ch_y == -1 else ch_y
        spatial_y = spatial_x if spatial_y == -1 else spatial_y
        if self.two_dim:
            stride, out_pad = (2, 1) if spatial_x != spatial_y else (1, 0)
            self.G_net = nn.Sequential(
                nn.ConvTranspose2d(ch_x, ch_y, kernel_size=3, padding=1, stride=stride, output_padding=out_pad),
                nn.BatchNorm2d(ch_y), nn.ReLU())
        else:
            self.G_net = nn.Sequential(nn.Conv1d(ch_x, ch_y, kernel_size=3, padding=1, stride=1), nn.ReLU())
        if not self.skip_critic:
            if self.two_dim:
                self.critic = nn.Sequential(
                    nn.Conv2d(ch_y, int(ch_y / 2), kernel_size=3, padding=1, stride=2), nn.BatchNorm2d(int(ch_y / 2)), nn.ReLU(),
                    nn.Conv2d(int(ch_y / 2), int(ch_y / 4), kernel_size=3, padding=1, stride=2), nn.BatchNorm2d(int(ch_y / 4)), nn.ReLU())
            else:
                form = getattr(getattr(config, 'DEV', None), 'OT_ONE_DIM_FORM', 'conv')
                if form == 'conv':
                    self.critic = nn.Sequential(nn.Conv1d(ch_y, int(ch_y / 4), kernel_size=3, padding=1, stride=1), nn.ReLU())
                elif form == 'fc':
                    self.critic = nn.Linear(ch_y, int(ch_y / 8))

    @staticmethod
    def _apply_1d(seq, t):
        """The 1-D branch runs Conv1d(k=3, padding=1) on length-1 sequences (x is [n, ch, 1], lib/model.py:207): only the
        centre tap ever meets data, so the layer is the GEMM  t @ W[:, :, 1]^T + b  -- same numbers, one cuBLAS call instead
        of cuDNN's NCHW<->NHWC round trip."""
        if t.dim() == 3 and t.size(2) == 1 and isinstance(seq, nn.Sequential) and isinstance(seq[0], nn.Conv1d) \
                and seq[0].kernel_size == (3,) and seq[0].padding == (1,) and seq[0].stride == (1,):
            out = torch.nn.functional.linear(t.squeeze(2), seq[0].weight[:, :, 1], seq[0].bias).unsqueeze(2)
            for layer in list(seq)[1:]:
                out = layer(out)
            return out
        return seq(t)

    def _critic_rows(self, t):
        c = self._apply_1d(self.critic, t)
        return c.view(c.size(0), c.size(1), -1)       # bs, channels (= samples N), positions (= D)

    def forward(self, x, y):
        x_up = self._apply_1d(self.G_net, x)
        cx, cy = self._critic_rows(x_up), self._critic_rows(y)
        bs = cx.size(0)
        if self.remove_bias:
            return sinkhorn_loss(cx, cy, 1. / self.epsilon, self.L)
        # all three terms in one launch: W(x^,y), W(x^,x^), W(y,y)        (OT_module.py:78-80)
        w = sinkhorn_loss(torch.cat([cx, cx, cy], 0), torch.cat([cy, cx, cy], 0), 1. / self.epsilon, self.L)
        return 2 * w[:bs] - w[bs:2 * bs] - w[2 * bs:]
