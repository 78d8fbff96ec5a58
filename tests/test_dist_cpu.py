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
