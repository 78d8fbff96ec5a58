"""CPU suite, part 3: the sharded path's host logic on 2 gloo ranks (no GPU): the all-reduced class statistics equal the
reference's gather-then-merge (lib/model.py:217-224) and the differentiable all-reduce back-propagates like DataParallel."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyref


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from feature_intertwiner_b200.dist import merged_class_sums, shard_batch
    g = torch.Generator().manual_seed(123)
    G, S, Fd, ncls = world, 3, 16, 81
    cnt = torch.randint(0, 4, (G, S, 1, ncls), generator=g).float()
    feat = torch.rand(G, S, Fd, ncls, generator=g) * (cnt > 0)
    lo, hi = shard_batch(G, rank, world)
    mine_f = feat[lo:hi].clone().requires_grad_()
    s, n = merged_class_sums(mine_f, cnt[lo:hi], None, True, differentiable=True, compensate=True)
    mean = s / (n + 1e-20)
    w = torch.rand(Fd, ncls, generator=g)
    (mean * w).sum().backward()
    # reference: gather every replica on one device, then _merge_feat_vec
    full_f = feat.clone().requires_grad_()
    want_mean, want_n = pyref.merge_feat_vec_ref(full_f, cnt)
    (want_mean * w).sum().backward()
    ok = torch.allclose(mean, want_mean, rtol=1e-6, atol=1e-7) and torch.equal(n, want_n.reshape(-1))
    # DataParallel sums replica gradients; DDP will average ours over `world`, hence the compensation factor
    ok = ok and torch.allclose(mine_f.grad / world, full_f.grad[lo:hi], rtol=1e-5, atol=1e-8)
    s2, n2 = merged_class_sums(feat[lo:hi], cnt[lo:hi], None, True, differentiable=False)
    ok = ok and torch.allclose(s2, s.detach()) and not s2.requires_grad
    # every rank must hold bit-identical totals (the replicated buffer update depends on it)
    gathered = [torch.empty_like(s2) for _ in range(world)]
    dist.all_gather(gathered, s2)
    ok = ok and all(torch.equal(gathered[0], t) for t in gathered)
    # both exchanges of an iteration in ONE collective (merged_class_sums_pair) == the two separate ones
    from feature_intertwiner_b200.dist import GradAllReduce, merged_class_sums_pair
    cnt_s = torch.randint(0, 3, (G, S, 1, ncls), generator=g).float()
    feat_s = torch.rand(G, S, Fd, ncls, generator=g) * (cnt_s > 0)
    fs = feat_s[lo:hi].clone().requires_grad_()
    bs, bn, ss, sn = merged_class_sums_pair(feat[lo:hi], cnt[lo:hi], fs, cnt_s[lo:hi], None, True, True)
    ss_ref, sn_ref = merged_class_sums(feat_s[lo:hi].clone(), cnt_s[lo:hi], None, True, differentiable=False)
    ok = ok and torch.equal(bs, s2) and torch.equal(bn, n2) and torch.allclose(ss, ss_ref) and torch.equal(sn, sn_ref) and not bs.requires_grad
    (ss * w).sum().backward()
    fr = feat_s.clone().requires_grad_()
    ((fr * cnt_s).sum(dim=(0, 1)) * w).sum().backward()
    ok = ok and torch.allclose(fs.grad / world, fr.grad[lo:hi], rtol=1e-5, atol=1e-8)
    # gradient bucket: hooks fire on the last accumulated gradient, finish() leaves the SUM over ranks in .grad
    lin = torch.nn.Linear(4, 3)
    with torch.no_grad():
        lin.weight.fill_(0.5); lin.bias.fill_(0.1)
    bucket = GradAllReduce(lin.parameters())
    x = torch.full((2, 4), float(rank + 1))
    lin(x).sum().backward()
    bucket.finish()
    want_w = sum(2.0 * (r + 1) for r in range(world))
    ok = ok and torch.allclose(lin.weight.grad, torch.full((3, 4), want_w)) and torch.allclose(lin.bias.grad, torch.full((3,), 2.0 * world))
    bucket.remove()
    # instance-level loss under sharding: all-reduced (sum of errors, count) == the loss over the concatenated instances
    from feature_intertwiner_b200.dist import _AllReduceSum
    inst = torch.rand(world, 5 + 3, Fd, generator=g)
    tgt = torch.rand(world, 5 + 3, Fd, generator=g)
    msk = (torch.rand(world, 5 + 3, generator=g) < 0.6).float()
    mine_i = inst[rank].clone().requires_grad_()
    per = (mine_i - tgt[rank]) ** 2 * msk[rank].unsqueeze(1)
    pair = _AllReduceSum.apply(torch.stack([per.sum(), msk[rank].sum() * Fd]), None, True)
    loss_sh = pair[0] / pair[1].detach().clamp(min=1.0)
    loss_sh.backward()
    full_i = inst.clone().requires_grad_()
    per_f = (full_i - tgt) ** 2 * msk.unsqueeze(2)
    loss_full = per_f.sum() / (msk.sum() * Fd).clamp(min=1.0)
    loss_full.backward()
    ok = ok and torch.allclose(loss_sh, loss_full, rtol=1e-6) and torch.allclose(mine_i.grad / world, full_i.grad[rank], rtol=1e-5, atol=1e-9)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_class_statistics_allreduce_matches_gather_merge():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_shard_batch_covers_everything_once():
    from feature_intertwiner_b200.dist import shard_batch
    for n in (1, 7, 8, 9):
        for world in (1, 2, 4, 8):
            spans = [shard_batch(n, r, world) for r in range(world)]
            covered = [i for lo, hi in spans for i in range(lo, hi)]
            assert covered == list(range(n))
