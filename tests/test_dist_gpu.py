"""GPU multi-rank parity (NCCL, 2 ranks; skipped on a 1-GPU box): the step sharded by image batch equals the single-rank step
on the concatenated batch -- what the reference's nn.DataParallel gather + _merge_feat_vec computes (lib/model.py:217-224,
394-402, tools/utils.py:645-654) -- the historical buffers stay bit-identical across ranks, and the CUDA-graphed loss head behind
the eager all-reduce gives the eager result."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(world, seed=5):
    """Per-rank shard r: features / class ids of the RoIs of its images on 3 levels, for the reliable and the less reliable set."""
    g = torch.Generator().manual_seed(seed)
    shards = []
    for r in range(world):
        lv = []
        for s in range(3):
            nb, ns = 40 + 7 * s + r, 55 + 3 * s + 2 * r
            lv.append(dict(big_gt=torch.randint(0, 81, (nb,), generator=g, dtype=torch.int32), big_f=torch.rand(nb, 1024, generator=g),
                           small_gt=torch.randint(0, 81, (ns,), generator=g, dtype=torch.int32), small_f=torch.rand(ns, 1024, generator=g)))
        shards.append(lv)
    return shards


def _stats(fi, lv, dev, leaves):
    bf, bc, sf, sc = [], [], [], []
    for d in lv:
        f, c = fi.assign_feat2cls(d["big_gt"].to(dev), d["big_f"].to(dev), 81)
        bf.append(f); bc.append(c)
        x = d["small_f"].to(dev).requires_grad_()
        leaves.append(x)
        f, c = fi.assign_feat2cls(d["small_gt"].to(dev), x, 81)
        sf.append(f); sc.append(c)
    return torch.stack(bf)[None], torch.stack(bc)[None], torch.stack(sf)[None], torch.stack(sc)[None]


def _worker(rank, world, port, loss_choice, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import feature_intertwiner_b200 as fi
    from oracle import pyref
    cfg = pyref.make_config(DEV__LOSS_CHOICE=loss_choice)
    torch.manual_seed(3)
    ot = fi.OptTrans(cfg, ch_x=1024, L=5).to(dev) if loss_choice == "ot" else None
    shards = _inputs(world)
    res = {}
    # ---- sharded: every rank its own images, class statistics all-reduced (twice: the buffer is state)
    mod = fi.IntertwinerLoss(cfg, ot_loss=ot, feat_dim=1024, distributed=True, ot_padded=(loss_choice == "ot")).to(dev)
    losses, grads = [], []
    for it in range(2):
        leaves = []
        feat_in = list(_stats(fi, shards[rank], dev, leaves)) + [None, None]
        loss = mod(feat_in)
        loss.sum().backward()
        losses.append(loss.detach().clone()); grads.append([x.grad.clone() for x in leaves])
    # buffers: bit-identical on every rank
    bufs = [torch.empty_like(mod.buffer) for _ in range(world)]
    dist.all_gather(bufs, mod.buffer.contiguous())
    cnts = [torch.empty_like(mod.buffer_cnt) for _ in range(world)]
    dist.all_gather(cnts, mod.buffer_cnt.contiguous())
    res["buffers_identical"] = all(torch.equal(b, bufs[0]) for b in bufs) and all(torch.equal(c, cnts[0]) for c in cnts)
    # ---- graphed loss head behind the eager all-reduce == eager
    mod_g = fi.IntertwinerLoss(cfg, ot_loss=ot, feat_dim=1024, distributed=True, ot_padded=(loss_choice == "ot")).to(dev)
    leaves = []
    feat_in = list(_stats(fi, shards[rank], dev, leaves)) + [None, None]
    res["graph_on"] = bool(mod_g.enable_cuda_graph([feat_in[0], feat_in[1], feat_in[2].detach().requires_grad_(), feat_in[3]]))
    glosses = []
    for it in range(2):
        leaves = []
        feat_in = list(_stats(fi, shards[rank], dev, leaves)) + [None, None]
        loss = mod_g(feat_in)
        loss.sum().backward()
        glosses.append(loss.detach().clone())
        res["graph_grad_%d" % it] = max(float((a.grad - b).abs().max()) for a, b in zip(leaves, grads[it]))
    res["graph_loss"] = max(float((a - b).abs().max()) for a, b in zip(glosses, losses))
    # ---- single rank on the concatenated batch: [G = world, S, F, ncls] exactly as DataParallel gathers it
    if rank == 0:
        single = fi.IntertwinerLoss(cfg, ot_loss=ot, feat_dim=1024, distributed=False, ot_padded=(loss_choice == "ot")).to(dev)
        for it in range(2):
            leaves, parts = [], []
            for r in range(world):
                parts.append(_stats(fi, shards[r], dev, leaves))
            feat_in = [torch.cat([p[k] for p in parts], dim=0) for k in range(4)] + [None, None]
            loss = single(feat_in)
            loss.sum().backward()
            res["loss_%d" % it] = float((loss.detach() - losses[it]).abs().max())
            mine = leaves[: len(grads[it])]                    # rank 0's leaves come first
            # DataParallel SUMS replica gradients; the sharded path compensates DDP's later averaging by x world
            res["grad_%d" % it] = max(float((a.grad * world - b).abs().max()) for a, b in zip(mine, grads[it]))
            res["grad_scale_%d" % it] = max(float(a.grad.abs().max()) for a in mine)
        res["buffer_vs_single"] = float((single.buffer - mod.buffer).abs().max())
        res["cnt_vs_single"] = bool(torch.equal(single.buffer_cnt, mod.buffer_cnt))
        out.update(res)
    else:
        out["rank%d" % rank] = {k: v for k, v in res.items() if k in ("buffers_identical", "graph_on", "graph_loss")}
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("loss_choice", ["l2", "ot"])
def test_two_rank_nccl_step_equals_single_rank(loss_choice):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), loss_choice, out), nprocs=world, join=True)
    out = dict(out)
    assert out["buffers_identical"] and out["rank1"]["buffers_identical"]
    assert out["cnt_vs_single"] and out["buffer_vs_single"] < 1e-6
    for it in range(2):
        assert out["loss_%d" % it] < LOSS_TOL, out
        assert out["grad_%d" % it] <= 1e-4 * max(out["grad_scale_%d" % it], 1e-6) + 1e-7, out
    if out["graph_on"]:
        assert out["graph_loss"] < 1e-6 and out["graph_grad_0"] < 1e-6 and out["graph_grad_1"] < 1e-6, out
