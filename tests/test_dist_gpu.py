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
    # OptTrans gradients: every rank evaluated the same loss head on the same totals -> the same bits, nothing to exchange
    # (the reference computes meta_loss on GPU 0 only: lib/model.py:143-144)
    res["ot_grads_identical"] = True
    if ot is not None:
        flat = torch.cat([p.grad.reshape(-1) for p in ot.parameters() if p.grad is not None])
        every = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(every, flat)
        # (in this 1-D form they are exactly zero besides: the critic's rows have ONE position, their cosine cost is a sign and
        # has no slope -- OT_module.py:106-109 with x [n, ch, 1], lib/model.py:207)
        res["ot_grads_identical"] = all(torch.equal(e, every[0]) for e in every)
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
        out["rank%d" % rank] = {k: v for k, v in res.items() if k in ("buffers_identical", "graph_on", "graph_loss", "ot_grads_identical")}
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
    assert out["ot_grads_identical"] and out["rank1"]["ot_grads_identical"]
    assert out["cnt_vs_single"] and out["buffer_vs_single"] < 1e-6
    for it in range(2):
        assert out["loss_%d" % it] < LOSS_TOL, out
        assert out["grad_%d" % it] <= 1e-4 * max(out["grad_scale_%d" % it], 1e-6) + 1e-7, out
    if out["graph_on"]:
        assert out["graph_loss"] < 1e-6 and out["graph_grad_0"] < 1e-6 and out["graph_grad_1"] < 1e-6, out


def _peer_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from feature_intertwiner_b200 import dist as fdist
    res = {}
    cap = 200003
    comm = fdist.PeerAllReduce(cap, None, dev)
    res["ok"] = comm.ok
    res["why"] = comm.why
    if comm.ok:
        g = torch.Generator().manual_seed(100 + rank)
        worst = 0.0
        # lengths: the tail path (n % 4 != 0), a single value, the full capacity, and a short call after a long one (slot reuse)
        for n in (cap, 1, 7, 4096, 166082, 5, cap, 1023):
            x = torch.randn(n, generator=g).to(dev)
            ref = x.clone()
            dist.all_reduce(ref)
            y = torch.empty_like(x)
            comm(x, out=y)                       # out of place
            comm(x)                              # in place
            worst = max(worst, float((y - ref).abs().max()), float((x - ref).abs().max()))
        res["eager_max_diff"] = worst            # two ranks: a + b in either order is the same fp32 value
        # inside a CUDA graph: the call counter lives on the device, every replay is a new call
        x = torch.zeros(166082, device=dev)
        y = torch.empty_like(x)
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            comm(x, out=y)
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            comm(x, out=y)
        gd = 0.0
        for it in range(5):
            x.copy_(torch.randn(166082, generator=g).to(dev))
            ref = x.clone()
            dist.all_reduce(ref)
            graph.replay()
            gd = max(gd, float((y - ref).abs().max()))
        res["graph_max_diff"] = gd
        # the loss module through the installed kernel == through NCCL (same inputs, fresh modules)
        import feature_intertwiner_b200 as fi
        from oracle import pyref
        cfg = pyref.make_config(DEV__LOSS_CHOICE="l2")
        shards = _inputs(world)
        losses = {}
        for mode in ("nccl", "peer"):
            if mode == "peer":
                inst = fdist.install_peer_allreduce(2 * (1024 * 81 + 81), None, dev)
                res["installed"] = inst.ok
            mod = fi.IntertwinerLoss(cfg, ot_loss=None, feat_dim=1024, distributed=True).to(dev)
            leaves = []
            loss = mod(list(_stats(fi, shards[rank], dev, leaves)) + [None, None])
            loss.sum().backward()
            losses[mode] = (loss.detach().clone(), [v.grad.clone() for v in leaves], mod.buffer.clone())
        res["loss_diff"] = float((losses["nccl"][0] - losses["peer"][0]).abs().max())
        res["grad_diff"] = max(float((a - b).abs().max()) for a, b in zip(losses["nccl"][1], losses["peer"][1]))
        res["buffer_diff"] = float((losses["nccl"][2] - losses["peer"][2]).abs().max())
        res["timeouts"] = comm.error() or inst.error()
        fdist.uninstall_peer_allreduce(None)
        comm.close()
    out["rank%d" % rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_peer_memory_allreduce_equals_nccl():
    """csrc/peer_allreduce.cu on 2 ranks: eager (every length class, in and out of place), replayed from a CUDA graph, and under
    the loss module -- bit-equal to NCCL's all-reduce (two addends commute)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_peer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    out = dict(out)
    for r in range(world):
        res = out["rank%d" % r]
        assert res["ok"], res
        assert res["eager_max_diff"] == 0.0 and res["graph_max_diff"] == 0.0, res
        assert res["installed"] and res["loss_diff"] == 0.0 and res["grad_diff"] == 0.0 and res["buffer_diff"] == 0.0, res
        assert not res["timeouts"], res
