"""Diagnostic (2 ranks): where do the OptTrans gradients of two ranks differ, and by how much?"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import feature_intertwiner_b200 as fi
    from oracle import pyref
    from test_dist_gpu import _inputs, _stats
    cfg = pyref.make_config(DEV__LOSS_CHOICE="ot")
    torch.manual_seed(3)
    ot = fi.OptTrans(cfg, ch_x=1024, L=5).to(dev)
    shards = _inputs(world)
    mod = fi.IntertwinerLoss(cfg, ot_loss=ot, feat_dim=1024, distributed=True, ot_padded=True).to(dev)

    def gather_equal(t, name):
        every = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(every, t.contiguous())
        d = float((every[0] - every[1]).abs().max())
        if rank == 0:
            print("%-40s max|r0-r1| = %.3e   scale %.3e" % (name, d, float(every[0].abs().max())), flush=True)

    for it in range(2):
        leaves = []
        feat_in = list(_stats(fi, shards[rank], dev, leaves)) + [None, None]
        loss = mod(feat_in)
        gather_equal(loss.detach(), "it%d loss" % it)
        for p in ot.parameters():
            p.grad = None
        loss.sum().backward()
        for n, p in ot.named_parameters():
            gather_equal(p.grad, "it%d grad %s" % (it, n))
    # the head alone on identical inputs, twice on the same rank and across ranks
    torch.manual_seed(11)
    x = torch.rand(80, 1024, 1, device=dev).requires_grad_()
    y = torch.rand(80, 1024, 1, device=dev)
    dist.broadcast(x.data, 0); dist.broadcast(y, 0)
    runs = []
    for k in range(2):
        for p in ot.parameters():
            p.grad = None
        l = ot(x, y)
        l.sum().backward()
        runs.append([l.detach().clone()] + [p.grad.clone() for p in ot.parameters()])
    if rank == 0:
        print("same rank, two runs: max diff", max(float((a - b).abs().max()) for a, b in zip(runs[0], runs[1])), flush=True)
    gather_equal(runs[0][0], "head alone: loss")
    for (n, _), g in zip(ot.named_parameters(), runs[0][1:]):
        gather_equal(g, "head alone: grad %s" % n)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
