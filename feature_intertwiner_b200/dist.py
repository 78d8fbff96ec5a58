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


def shard_batch(n_items, rank, world):
    """Contiguous shard [lo, hi) of a batch of n_items images for `rank` of `world` (DataParallel's scatter along dim 0)."""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)
