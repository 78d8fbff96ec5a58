"""The one exchange step of the path when it is sharded by image batch (one process per GPU).

The reference runs nn.DataParallel: every replica returns ``feat[1,S,F,ncls]`` / ``cnt[1,S,1,ncls]``, gather concatenates
them on GPU 0 along dim 0 and ``_merge_feat_vec`` reduces over (gpu, scale) (lib/model.py:217-224,394-402).  Here every
rank reduces its own (scale) axis and ONE all-reduce(SUM) of the packed un-normalised sums replaces the gather; every
rank then holds the same totals and performs the identical buffer update.  The collective is torch.distributed's (NCCL or
gloo) unless a ``PeerAllReduce`` is installed for the group: then it is ONE kernel of this library over NVLink / NVSwitch peer
memory (csrc/peer_allreduce.cu) -- a few microseconds instead of NCCL's 0.1 ms at this size, bit-identical totals on every
rank, and capturable inside the whole-step CUDA graph.
"""
import ctypes as C

import torch

_PEER = {}        # group (None = default) -> PeerAllReduce used by all_reduce_sum_ for CUDA fp32 tensors that fit
FUSED_MERGE = True   # merged_class_sums_pair on CUDA tensors: one kernel each way (False: the torch op sequence, kept for A/B tests)


class PeerAllReduce(object):
    """All-reduce(SUM) of fp32 CUDA tensors of up to ``capacity`` values through peer-mapped regions (fi_peer_* in
    include/fi_b200.h).  Collective constructor: every rank of ``group`` calls it; the IPC handles travel through
    torch.distributed.  Every rank must then make the same sequence of calls.  ``ok`` is False (on every rank) when the regions
    could not be mapped -- e.g. no peer access between two of the GPUs -- and the caller keeps using NCCL."""

    def __init__(self, capacity, group=None, device=None):
        import torch.distributed as dist
        from . import _lib
        self._lib, self.group = _lib, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.capacity = int(capacity)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.own, self.imported, self.regions = None, [], None
        L = _lib.lib()
        ok, why = True, ""
        handle = torch.zeros(L_HANDLE, dtype=torch.uint8)
        try:
            if self.world > 16:
                raise _lib.FiError("at most 16 ranks")
            with torch.cuda.device(self.device):
                own = C.c_void_p()
                _lib.check(L.fi_peer_region_alloc(self.capacity, C.byref(own)))
                self.own = own.value
                buf = (C.c_ubyte * L_HANDLE)()
                _lib.check(L.fi_peer_region_export(self.own, C.cast(buf, C.c_void_p)))
                handle = torch.tensor(list(buf), dtype=torch.uint8)
        except Exception as exc:                        # noqa: BLE001 -- reported through `ok` on every rank
            ok, why = False, repr(exc)[:200]
        # handles of every rank (a CUDA tensor on NCCL groups, a CPU one on gloo)
        on_dev = dist.get_backend(group) == "nccl"
        mine = torch.cat([handle, torch.tensor([1 if ok else 0], dtype=torch.uint8)])
        mine = mine.to(self.device) if on_dev else mine
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=group)
        allh = [h.cpu() for h in allh]
        ok = ok and all(int(h[-1]) == 1 for h in allh)
        ptrs = []
        if ok:
            try:
                with torch.cuda.device(self.device):
                    for r in range(self.world):
                        if r == self.rank:
                            ptrs.append(self.own)
                            continue
                        buf = (C.c_ubyte * L_HANDLE)(*[int(v) for v in allh[r][:L_HANDLE]])
                        p = C.c_void_p()
                        _lib.check(L.fi_peer_region_import(C.cast(buf, C.c_void_p), C.byref(p)))
                        self.imported.append(p.value)
                        ptrs.append(p.value)
            except Exception as exc:                    # noqa: BLE001
                ok, why = False, repr(exc)[:200]
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32)
        flag = flag.to(self.device) if on_dev else flag
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.ok, self.why = bool(int(flag.item()) == 1), why
        if self.ok:
            self.regions = (C.c_void_p * self.world)(*ptrs)
        else:
            self.close()

    def __call__(self, t, out=None):
        """out = sum over ranks of t (``out`` defaults to ``t``: in place).  Enqueued on the current stream of t's device."""
        _lib = self._lib
        if not self.ok:
            raise _lib.FiError("PeerAllReduce: regions are not mapped (%s)" % (self.why,))
        out = t if out is None else out
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and out.is_contiguous() and out.dtype == torch.float32
                and out.numel() == t.numel() and t.numel() <= self.capacity):
            raise _lib.FiError("PeerAllReduce: contiguous fp32 CUDA tensors of at most %d values" % self.capacity)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().fi_peer_allreduce_sum(_lib.ptr(t), _lib.ptr(out), t.numel(), self.rank, self.world, C.cast(self.regions, C.c_void_p),
                                                        self.capacity, _lib.stream_ptr(t.device)))
        return out

    def fits(self, t):
        return self.ok and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() <= self.capacity \
            and t.data_ptr() % 16 == 0

    def error(self):
        """True when a wait inside one of the kernels timed out (a peer did not make the matching call).  Synchronises."""
        v = C.c_int(0)
        self._lib.check(self._lib.lib().fi_peer_error(self.own, C.byref(v)))
        return v.value != 0

    def close(self):
        L = self._lib.lib()
        for p in self.imported:
            L.fi_peer_region_release(p)
        self.imported = []
        if self.own is not None:
            L.fi_peer_region_free(self.own)
            self.own = None
        self.ok = False


L_HANDLE = 64     # FI_PEER_HANDLE_BYTES


def install_peer_allreduce(capacity, group=None, device=None):
    """Collective: route this module's all-reduces of ``group`` (tensors of up to ``capacity`` fp32 values) through a
    PeerAllReduce.  Returns it (``.ok`` False -> nothing installed, NCCL stays)."""
    comm = PeerAllReduce(capacity, group, device)
    if comm.ok:
        _PEER[group] = comm
    return comm


def uninstall_peer_allreduce(group=None):
    comm = _PEER.pop(group, None)
    if comm is not None:
        comm.close()


def all_reduce_sum_(t, group=None):
    """In-place all-reduce(SUM): the peer-memory kernel when one is installed for the group and the tensor fits, else
    torch.distributed's."""
    comm = _PEER.get(group)
    if comm is not None and comm.fits(t):
        return comm(t)
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class _AllReduceSum(torch.autograd.Function):
    """Differentiable all-reduce(SUM).  Every rank's contribution enters the total with weight 1, so the local gradient is
    the incoming one -- times world_size when ``compensate`` is set, because DDP later AVERAGES parameter gradients over
    ranks whereas the reference's DataParallel SUMS the replicas' gradients."""

    @staticmethod
    def forward(ctx, t, group, compensate):
        import torch.distributed as dist
        ctx.scale = float(dist.get_world_size(group)) if compensate else 1.0
        return all_reduce_sum_(t.clone(), group)

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
            packed = all_reduce_sum_(packed.detach().clone(), group)
        s, n = packed[: s.numel()].view_as(s), packed[s.numel():]
    return s, n


class _MergeStats(torch.autograd.Function):
    """_merge_feat_vec of both sets + the exchange, CUDA path: ONE kernel writes the un-normalised sums of the reliable and the
    less-reliable set straight into the packed buffer the all-reduce works on (csrc/loss_head.cu::merge_stats_kernel), the
    collective runs in place, and the four results are views of that buffer.  Backward is one kernel as well (through the
    all-reduce's scale and the count weighting).  Replaces ~12 pointwise / reduction / cat launches each way."""

    @staticmethod
    def forward(ctx, big_feat, big_cnt, small_feat, small_cnt, group, distributed, compensate):
        from . import _lib
        G, S, Fd, ncls = small_feat.shape
        bf, bc = big_feat.detach().contiguous(), big_cnt.detach().contiguous()
        sf, sc = small_feat.detach().contiguous(), small_cnt.detach().contiguous()
        a = Fd * ncls
        packed = torch.empty((2 * (a + ncls),), device=sf.device, dtype=torch.float32)
        with torch.cuda.device(sf.device):
            _lib.check(_lib.lib().fi_merge_stats(_lib.ptr(bf), _lib.ptr(bc), _lib.ptr(sf), _lib.ptr(sc), G * S, Fd, ncls, _lib.ptr(packed),
                                                 _lib.stream_ptr(sf.device)))
        ctx.scale = 1.0
        if distributed:
            import torch.distributed as dist
            all_reduce_sum_(packed, group)
            ctx.scale = float(dist.get_world_size(group)) if compensate else 1.0
        ctx.save_for_backward(sc)
        ctx.dims = (G, S, Fd, ncls)
        bs, bn = packed[:a].view(Fd, ncls), packed[a:a + ncls]
        ss, sn = packed[a + ncls:2 * a + ncls].view(Fd, ncls), packed[2 * a + ncls:]
        ctx.mark_non_differentiable(bs, bn, sn)
        ctx.set_materialize_grads(False)          # no zero-filled gradients for the three outputs nothing differentiates
        return bs, bn, ss, sn

    @staticmethod
    def backward(ctx, _dbs, _dbn, dss, _dsn):
        from . import _lib
        if dss is None or not ctx.needs_input_grad[2]:
            return None, None, None, None, None, None, None
        (sc,) = ctx.saved_tensors
        G, S, Fd, ncls = ctx.dims
        dss = dss.contiguous()
        dsf = torch.empty((G, S, Fd, ncls), device=dss.device, dtype=torch.float32)
        with torch.cuda.device(dss.device):
            _lib.check(_lib.lib().fi_merge_stats_backward(_lib.ptr(dss), _lib.ptr(sc), G * S, Fd, ncls, ctx.scale, _lib.ptr(dsf),
                                                          _lib.stream_ptr(dss.device)))
        return None, None, dsf, None, None, None, None


def _merge_kernel_ok(big_feat, big_cnt, small_feat, small_cnt):
    ts = (big_feat, big_cnt, small_feat, small_cnt)
    return all(t.is_cuda and t.dtype == torch.float32 for t in ts) and small_feat.dim() == 4 and big_feat.shape == small_feat.shape \
        and big_cnt.numel() == small_cnt.numel() == small_feat.size(0) * small_feat.size(1) * small_feat.size(3) \
        and not small_cnt.requires_grad


def merged_class_sums_pair(big_feat, big_cnt, small_feat, small_cnt, group=None, distributed=False, compensate=True):
    """Both exchanges of an iteration -- reliable set (no gradient) and less-reliable set (differentiable) -- in ONE all-reduce of
    the packed sums: (big_sum [F,ncls], big_n [ncls], small_sum [F,ncls], small_n [ncls]).  At 1.3 MB the cost of an all-reduce on
    NVSwitch is its launch + latency, not bandwidth, so one call instead of two halves it."""
    if FUSED_MERGE and _merge_kernel_ok(big_feat, big_cnt, small_feat, small_cnt):
        return _MergeStats.apply(big_feat, big_cnt, small_feat, small_cnt, group, bool(distributed), bool(compensate))
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
            packed = all_reduce_sum_(packed.detach().clone(), group)
        a, b = bs.numel(), bn.numel()
        bs, bn = packed[:a].view_as(bs).detach(), packed[a:a + b].detach()
        ss, sn = packed[a + b:2 * a + b].view_as(ss), packed[2 * a + b:]
    return bs, bn, ss, sn


class GradAllReduce(object):
    """All-reduce(SUM) of a module's parameter gradients in one flat bucket, started the moment the LAST of them has been
    accumulated (post-accumulate-grad hooks, like DDP's buckets) so that it overlaps with the rest of the backward pass -- here the
    RoIAlign backward, which runs after the loss head's.  ``finish()`` waits and writes the reduced values back.

    With ``peer`` (a PeerAllReduce of at least the bucket's size) the collective is this library's kernel on a side stream:
    fork after the last gradient, join in ``finish()`` -- a pattern a CUDA graph capture records as two branches."""

    def __init__(self, params, group=None, peer=None):
        import torch.distributed as dist
        self.params = [p for p in params if p.requires_grad]
        self.group, self.dist, self.peer = group, dist, peer
        self.pending, self.work, self.flat = len(self.params), None, None
        self.side = torch.cuda.Stream(device=self.params[0].device) if peer is not None else None
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    def _hook(self, _param):
        self.pending -= 1
        if self.pending == 0:
            self.flat = torch.cat([p.grad.reshape(-1) for p in self.params])
            if self.peer is not None:
                cur = torch.cuda.current_stream(self.flat.device)
                self.side.wait_stream(cur)
                with torch.cuda.stream(self.side):
                    self.peer(self.flat)
                self.flat.record_stream(self.side)
                self.work = self.side
            else:
                self.work = self.dist.all_reduce(self.flat, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        if self.work is not None:
            if self.peer is not None:
                torch.cuda.current_stream(self.flat.device).wait_stream(self.side)
            else:
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
