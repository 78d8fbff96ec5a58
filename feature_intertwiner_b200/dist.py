"""The one exchange step of the path when it is sharded by image batch (one process per GPU).

The reference runs nn.DataParallel: every replica returns ``feat[1,S,F,ncls]`` / ``cnt[1,S,1,ncls]``, gather concatenates
them on GPU 0 along dim 0 and ``_merge_feat_vec`` reduces over (gpu, scale) (lib/model.py:217-224,394-402).  Here every
rank reduces its own (scale) axis and ONE all-reduce(SUM) of the packed un-normalised sums replaces the gather; every
rank then holds the same totals and performs the identical buffer update.  Pure torch: runs on NCCL and on gloo.
"""
import torch


class _AllReduceSum(torch.autograd.Function):
    """Differentiable all-reduce(SUM).  Every rank's contribution enters the total with weight 1, so the local gradient is
    the incoming one -- times world_size when ``compensate`` is set, because DDP later AVERAGES parameter gradients over
    ranks whereas the reference's DataParallel SUMS the replicas' gradients."""

    @staticmethod
    def forward(ctx, t, group, compensate):
        import torch.distributed as dist
        ctx.scale = float(dist.get_world_size(group)) if compensate else 1.0
        out = t.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        return g * ctx.scale, None, None


def merged_class_sums(feat, cnt, group=None, distributed=False, differentiable=True, compensate=True):
    """feat[G,S,F,ncls], cnt[G,S,1,ncls] -> (sum_{g,s} feat*cnt [F,ncls], sum_{g,s} cnt [ncls]) over ALL ranks."""
    s = (feat * cnt).sum(dim=(0, 1))
    n = cnt.sum(dim=(0, 1)).reshape(-1)
    if distributed:
        import torch.distributed as dist
        packed = torch.cat([s.reshape(-1), n])
        if differentiable and packed.requires_grad:
            packed = _AllReduceSum.apply(packed, group, compensate)
        else:
            packed = packed.detach().clone()
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        s, n = packed[: s.numel()].view_as(s), packed[s.numel():]
    return s, n


def merged_class_sums_pair(big_feat, big_cnt, small_feat, small_cnt, group=None, distributed=False, compensate=True):
    """Both exchanges of an iteration -- reliable set (no gradient) and less-reliable set (differentiable) -- in ONE all-reduce of
    the packed sums: (big_sum [F,ncls], big_n [ncls], small_sum [F,ncls], small_n [ncls]).  At 1.3 MB the cost of an all-reduce on
    NVSwitch is its launch + latency, not bandwidth, so one call instead of two halves it."""
    bs = (big_feat.detach() * big_cnt.detach()).sum(dim=(0, 1))
    bn = big_cnt.detach().sum(dim=(0, 1)).reshape(-1)
    ss = (small_feat * small_cnt).sum(dim=(0, 1))
    sn = small_cnt.sum(dim=(0, 1)).reshape(-1)
    if distributed:
        import torch.distributed as dist
        packed = torch.cat([bs.reshape(-1), bn, ss.reshape(-1), sn])
        if packed.requires_grad:
            packed = _AllReduceSum.apply(packed, group, compensate)
        else:
            packed = packed.detach().clone()
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        a, b = bs.numel(), bn.numel()
        bs, bn = packed[:a].view_as(bs).detach(), packed[a:a + b].detach()
        ss, sn = packed[a + b:2 * a + b].view_as(ss), packed[2 * a + b:]
    return bs, bn, ss, sn


class GradAllReduce(object):
    """All-reduce(SUM) of a module's parameter gradients in one flat bucket, started the moment the LAST of them has been
    accumulated (post-accumulate-grad hooks, like DDP's buckets) so that it overlaps with the rest of the backward pass -- here the
    RoIAlign backward, which runs after the loss head's.  ``finish()`` waits and writes the reduced values back."""

    def __init__(self, params, group=None):
        import torch.distributed as dist
        self.params = [p for p in params if p.requires_grad]
        self.group, self.dist = group, dist
        self.pending, self.work, self.flat = len(self.params), None, None
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    def _hook(self, _param):
        self.pending -= 1
        if self.pending == 0:
            self.flat = torch.cat([p.grad.reshape(-1) for p in self.params])
            self.work = self.dist.all_reduce(self.flat, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        if self.work is not None:
            self.work.wait()
            off = 0
            for p in self.params:
                n = p.numel()
                p.grad.copy_(self.flat[off:off + n].view_as(p.grad))
                off += n
        self.pending, self.work = len(self.params), None
        return self.flat

    def remove(self):
        for h in self.handles:
            h.remove()


def shard_batch(n_items, rank, world):
    """Contiguous shard [lo, hi) of a batch of n_items images for `rank` of `world` (DataParallel's scatter along dim 0)."""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)
