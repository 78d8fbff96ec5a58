"""RoIAlign operator layer -- same names and call shapes as the reference's lib/roi_align package.

* ``CropAndResizeFunction(crop_h, crop_w, extrapolation_value=0)(image, boxes, box_ind)``
  mirrors lib/roi_align/crop_and_resize.py:14-54 (construct with hyper-parameters, call on tensors;
  gradient w.r.t. ``image`` only).
* ``RoIAlign(crop_h, crop_w, extrapolation_value=0, transform_fpcoor=True)`` mirrors
  lib/roi_align/roi_align.py:6-48.

Both accept the image in either memory format: contiguous NCHW (the reference's) or
``torch.channels_last`` (the native one here -- see csrc/roi_align.cu); crops come back in the same format,
logical shape ``[R, C, crop_h, crop_w]`` either way.
"""
import ctypes as C
import os

import torch
from torch import nn

from . import _lib

# ---- optional per-launch timing (bench.py): CUDA events on the launching stream + what is needed to count the
# launch's algorithmic bytes afterwards.  Off by default; nothing is recorded and no event is created then.
_PROFILE = None
_EVENT_POOL = []


def enable_profiling(prealloc_events=0):
    """Record a (start, end) CUDA-event pair around every RoIAlign launch.  ``prealloc_events`` timing events are created AND
    recorded once up front: creating them inside a timed loop costs sporadic 5-100 ms host stalls in the driver (measured with
    bench.py: every third step or so), which would be charged to the step."""
    global _PROFILE
    _PROFILE = []
    for _ in range(prealloc_events):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()                                   # torch creates the CUDA event lazily, on its first record
        _EVENT_POOL.append(ev)
    return _PROFILE


def disable_profiling():
    global _PROFILE
    rec, _PROFILE = _PROFILE, None
    del _EVENT_POOL[:]
    return rec or []


def _event():
    return _EVENT_POOL.pop() if _EVENT_POOL else torch.cuda.Event(enable_timing=True)


_SETS_BY_SHAPE = {}


class _Timed(object):
    def __init__(self, kernel, stream=None, **meta):
        self.rec = None
        self.stream = stream                         # the events go on the stream the kernels are launched on
        if _PROFILE is not None:
            if "sets" in meta:
                # Keep the tensors of ONE launch per shape signature (they are only needed afterwards, to count the distinct tap
                # pixels); holding every step's box lists alive stops the caching allocator from recycling the split buffers,
                # and the fresh cudaMalloc segments it then needs showed up as 5-100 ms steps in bench.py.
                sig = (kernel,) + tuple((st["im_size"], st["crop"], int(st["boxes"].size(0)), bool(st.get("dual")), st.get("count") is not None)
                                        for st in meta["sets"])
                _SETS_BY_SHAPE.setdefault(sig, meta["sets"])
                meta = dict(like=sig)
            self.rec = dict(kernel=kernel, start=_event(), end=_event(), **meta)

    def __enter__(self):
        if self.rec is not None:
            self.rec["start"].record(self.stream)

    def __exit__(self, *exc):
        if self.rec is not None:
            self.rec["end"].record(self.stream)
            _PROFILE.append(self.rec)


_UNIQUE_CACHE = {}


def _distinct_pixels(boxes, box_ind, B, H, W, crops):
    """Number of distinct (image, y, x) feature pixels the crops of these boxes tap -- the UNION over the crop sizes in ``crops``
    -- counted on the device from the kernel's own tap table."""
    R = boxes.size(0)
    key = (boxes.data_ptr(), box_ind.data_ptr(), R, B, H, W, tuple(crops))
    if key not in _UNIQUE_CACHE:
        seen = torch.zeros(B * H * W, dtype=torch.bool, device=boxes.device)
        for ph, pw in crops:
            taps = crop_taps(boxes, H, W, ph, pw).view(R, -1, 5).long()
            b = box_ind.long().view(R, 1).expand(R, taps.size(1))
            ok = (taps[:, :, 4] == 1) & (b >= 0) & (b < B)
            for yy, xx in ((0, 2), (0, 3), (1, 2), (1, 3)):
                lin = (b * H + taps[:, :, yy]) * W + taps[:, :, xx]
                seen[lin[ok]] = True
        _UNIQUE_CACHE[key] = int(seen.sum().item())
    return _UNIQUE_CACHE[key]


def _sets_forward_bytes(sets):
    """Level-batched forward launch: every set writes its crops; the sets that crop the SAME boxes from the SAME map (Dev's 7x7 and
    14x14 crops of the small boxes) are walked box by box by the kernel, so a pixel they both tap is read once: U is the union."""
    total, groups = 0, {}
    for st in sets:
        B, Cc, H, W = st["im_size"]
        ph, pw = st["crop"]
        R = st["boxes"].size(0)
        if st.get("count") is not None:
            R = min(R, int(st["count"].item()))
        total += 4 * Cc * R * ph * pw * (2 if st.get("dual") else 1) + 20 * R
        key = (st.get("image"), st["boxes"].data_ptr(), st["box_ind"].data_ptr(), R, st["im_size"])
        groups.setdefault(key, dict(st=st, R=R, crops=[]))["crops"].append((ph, pw))
    for g in groups.values():
        st, R = g["st"], g["R"]
        B, Cc, H, W = st["im_size"]
        total += 4 * Cc * _distinct_pixels(st["boxes"][:R], st["box_ind"][:R], B, H, W, sorted(set(g["crops"])))
    return total


def algorithmic_bytes(rec):
    """Algorithmic bytes of one recorded launch:
    forward  = 4*C*R*P^2 (write crops) + 4*C*U (each touched feature pixel read once) + 20*R (boxes, box_ind)
    backward = 4*C*R*P^2 (read grads)  + 4*C*B*H*W (dense grad map written once)      + 20*R
    U = distinct (b,y,x) tap pixels, counted on the device from the kernel's own tap table; in a level-batched launch the union
    over the crop sizes taken from the same boxes on the same map (_sets_forward_bytes)."""
    if "alg_bytes" in rec:
        return rec["alg_bytes"]
    if "bwd" in rec:                                   # backward of a crop_sets pass: (C, P, two sources, capacity, device count) per set
        total = rec["bwd"]["map_bytes"]
        for Cc, P, dual, cap, count in rec["bwd"]["items"]:
            R = cap if count is None else min(cap, int(count.item()))
            total += 4 * Cc * R * P * P * (2 if dual else 1) + 20 * R
        rec["alg_bytes"] = total
        return total
    if "like" in rec:                                  # same shapes as the launch whose tensors were kept
        sets = _SETS_BY_SHAPE[rec["like"]]
        if isinstance(sets, list):
            _SETS_BY_SHAPE[rec["like"]] = _sets_forward_bytes(sets) if rec["kernel"].startswith("crop_fwd") else \
                sum(algorithmic_bytes(dict(kernel=rec["kernel"], **st)) for st in sets)
        return _SETS_BY_SHAPE[rec["like"]]
    if "sets" in rec:
        if rec["kernel"].startswith("crop_fwd"):
            return _sets_forward_bytes(rec["sets"])
        return sum(algorithmic_bytes(dict(kernel=rec["kernel"], **st)) for st in rec["sets"])
    B, Cc, H, W = rec["im_size"]
    ph, pw = rec["crop"]
    R = rec["boxes"].size(0)
    if rec.get("count") is not None:                   # device-side list length: only the live boxes move bytes
        R = min(R, int(rec["count"].item()))
        rec = dict(rec, boxes=rec["boxes"][:R], box_ind=rec["box_ind"][:R])
    base = 4 * Cc * R * ph * pw * (2 if rec.get("dual") else 1) + 20 * R
    if rec["kernel"].startswith("crop_bwd"):
        return base + 4 * Cc * B * H * W
    return base + 4 * Cc * _distinct_pixels(rec["boxes"], rec["box_ind"], B, H, W, [(ph, pw)])


def _mem_format(layout):
    return torch.channels_last if layout == _lib.FI_LAYOUT_NHWC else torch.contiguous_format


def _prep_boxes(boxes, box_ind, device):
    if boxes.dim() != 2 or boxes.size(1) != 4:
        raise _lib.FiError("boxes must be [R,4] (y1,x1,y2,x2), got %s" % (tuple(boxes.shape),))
    if box_ind.dim() != 1 or box_ind.size(0) != boxes.size(0):
        raise _lib.FiError("box_ind must be [R]")
    boxes = boxes.detach().to(device=device, dtype=torch.float32).contiguous()
    box_ind = box_ind.detach().to(device=device, dtype=torch.int32).contiguous()
    return boxes, box_ind


class _CropAndResize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, boxes, box_ind, crop_h, crop_w, extrap, out, dst_row):
        _lib.require_cuda(image)
        if image.dtype != torch.float32:
            raise _lib.FiError("crop_and_resize computes in fp32 like the reference; got %s" % image.dtype)
        layout, image = _lib.layout_of(image)
        boxes, box_ind = _prep_boxes(boxes, box_ind, image.device)
        B, Cc, H, W = image.shape
        R = boxes.size(0)
        if out is None:
            if dst_row is not None:
                raise _lib.FiError("dst_row needs a preallocated `out`")
            crops = torch.empty((R, Cc, crop_h, crop_w), device=image.device, dtype=torch.float32,
                                memory_format=_mem_format(layout))
        else:
            crops = out
            ok = crops.is_contiguous(memory_format=_mem_format(layout)) and crops.dtype == torch.float32 \
                and tuple(crops.shape[1:]) == (Cc, crop_h, crop_w)
            if not ok:
                raise _lib.FiError("`out` must be fp32 [rows,%d,%d,%d] in the image's memory format" % (Cc, crop_h, crop_w))
            ctx.mark_dirty(crops)
        ctx.has_out = out is not None
        if dst_row is not None:
            dst_row = dst_row.to(device=image.device, dtype=torch.int32).contiguous()
        tag = "crop_fwd_nhwc" if layout == _lib.FI_LAYOUT_NHWC else "crop_fwd_nchw"
        with torch.cuda.device(image.device), _Timed(tag, im_size=(B, Cc, H, W), crop=(crop_h, crop_w), boxes=boxes, box_ind=box_ind):
            _lib.check(_lib.lib().fi_crop_and_resize_forward(
                _lib.ptr(image), layout, _lib.ptr(boxes), _lib.ptr(box_ind), _lib.ptr(dst_row), R, B, H, W,
                crop_h, crop_w, Cc, float(extrap), _lib.ptr(crops), layout, _lib.stream_ptr(image.device)))
        ctx.save_for_backward(boxes, box_ind, dst_row if dst_row is not None else torch.empty(0))
        ctx.has_rows = dst_row is not None
        ctx.im_size = (B, Cc, H, W)
        ctx.layout = layout
        ctx.crop = (crop_h, crop_w)
        return crops

    @staticmethod
    def backward(ctx, grad_out):
        boxes, box_ind, rows = ctx.saved_tensors
        rows = rows if ctx.has_rows else None
        B, Cc, H, W = ctx.im_size
        layout = ctx.layout
        grad_out = grad_out.contiguous(memory_format=_mem_format(layout))
        grad_image = torch.empty(ctx.im_size, device=grad_out.device, dtype=torch.float32, memory_format=_mem_format(layout))
        tag = "crop_bwd_nhwc" if layout == _lib.FI_LAYOUT_NHWC else "crop_bwd_nchw"
        with torch.cuda.device(grad_out.device), _Timed(tag, im_size=ctx.im_size, crop=ctx.crop, boxes=boxes, box_ind=box_ind):
            _lib.check(_lib.lib().fi_crop_and_resize_backward(
                _lib.ptr(grad_out), layout, _lib.ptr(boxes), _lib.ptr(box_ind), _lib.ptr(rows), boxes.size(0), B, H, W,
                ctx.crop[0], ctx.crop[1], Cc, _lib.ptr(grad_image), layout, 0, _lib.stream_ptr(grad_out.device)))
        # `out` was written in place: rows this call did not touch pass their gradient through to whoever wrote them; the rows it
        # overwrote no longer depend on the previous contents (zero gradient)
        g_prev = None
        if ctx.has_out and ctx.needs_input_grad[6]:
            g_prev = grad_out.clone()
            if rows is not None:
                g_prev.index_fill_(0, rows.long(), 0.0)
            else:
                g_prev[: boxes.size(0)].zero_()
        return grad_image, None, None, None, None, None, g_prev, None


class _CropPair(torch.autograd.Function):
    """Both crops Dev.forward takes of one level's made-up map (lib/sub_module.py:555-572) as ONE autograd node:
    size_a x size_a into rows ``dst_row`` of ``out_a``; size_b x size_b into rows ``dst_row`` of ``out_b`` and, when
    ``compact_b``, also into a compact [R,...] tensor for the critic.  Backward is ONE pass over the map
    (fi_crop_and_resize_backward_multi): one zero fill instead of two, no grad-accumulation add, no index_copy."""

    @staticmethod
    def forward(ctx, image, boxes, box_ind, dst_row, out_a, out_b, size_a, size_b, compact_b):
        _lib.require_cuda(image, out_a, out_b)
        layout, image = _lib.layout_of(image)
        if layout != _lib.FI_LAYOUT_NHWC or image.dtype != torch.float32:
            raise _lib.FiError("crop_pair needs an fp32 channels_last feature map")
        boxes, box_ind = _prep_boxes(boxes, box_ind, image.device)
        dst_row = dst_row.to(device=image.device, dtype=torch.int32).contiguous()
        B, Cc, H, W = image.shape
        R = boxes.size(0)
        for o, sz in ((out_a, size_a), (out_b, size_b)):
            if not (o.is_contiguous(memory_format=torch.channels_last) and o.dtype == torch.float32 and tuple(o.shape[1:]) == (Cc, sz, sz)):
                raise _lib.FiError("crop_pair outputs must be fp32 channels_last [rows,%d,%d,%d]" % (Cc, sz, sz))
        compact = torch.empty((R, Cc, size_b, size_b), device=image.device, memory_format=torch.channels_last) if compact_b else None
        L, s = _lib.lib(), _lib.stream_ptr(image.device)
        with torch.cuda.device(image.device):
            with _Timed("crop_fwd_nhwc", im_size=(B, Cc, H, W), crop=(size_a, size_a), boxes=boxes, box_ind=box_ind):
                _lib.check(L.fi_crop_and_resize_forward(_lib.ptr(image), layout, _lib.ptr(boxes), _lib.ptr(box_ind), _lib.ptr(dst_row), R, B, H, W,
                                                        size_a, size_a, Cc, 0.0, _lib.ptr(out_a), layout, s))
            with _Timed("crop_fwd_nhwc", im_size=(B, Cc, H, W), crop=(size_b, size_b), boxes=boxes, box_ind=box_ind, dual=compact_b):
                if compact_b:
                    _lib.check(L.fi_crop_and_resize_forward_dual(_lib.ptr(image), _lib.ptr(boxes), _lib.ptr(box_ind), _lib.ptr(dst_row), R, B, H, W,
                                                                 size_b, size_b, Cc, 0.0, _lib.ptr(out_b), _lib.ptr(compact), s))
                else:
                    _lib.check(L.fi_crop_and_resize_forward(_lib.ptr(image), layout, _lib.ptr(boxes), _lib.ptr(box_ind), _lib.ptr(dst_row), R, B, H, W,
                                                            size_b, size_b, Cc, 0.0, _lib.ptr(out_b), layout, s))
        ctx.mark_dirty(out_a, out_b)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(boxes, box_ind, dst_row)
        ctx.meta = (B, Cc, H, W, size_a, size_b, compact_b)
        return (out_a, out_b, compact) if compact_b else (out_a, out_b)

    @staticmethod
    def backward(ctx, g_a, g_b, g_c=None):
        boxes, box_ind, dst_row = ctx.saved_tensors
        B, Cc, H, W, size_a, size_b, compact_b = ctx.meta
        R = boxes.size(0)
        cl = torch.channels_last
        g_a = None if g_a is None else g_a.contiguous(memory_format=cl)
        g_b = None if g_b is None else g_b.contiguous(memory_format=cl)
        g_c = None if g_c is None else g_c.contiguous(memory_format=cl)
        dev = boxes.device
        sets = []
        if g_a is not None:
            sets.append(_lib.CropSet(_lib.ptr(g_a), None, _lib.ptr(boxes), _lib.ptr(box_ind), _lib.ptr(dst_row), R, size_a, size_a))
        if g_b is not None:
            sets.append(_lib.CropSet(_lib.ptr(g_b), _lib.ptr(g_c), _lib.ptr(boxes), _lib.ptr(box_ind), _lib.ptr(dst_row), R, size_b, size_b))
        elif g_c is not None:
            sets.append(_lib.CropSet(_lib.ptr(g_c), None, _lib.ptr(boxes), _lib.ptr(box_ind), None, R, size_b, size_b))
        grad_image = torch.empty((B, Cc, H, W), device=dev, dtype=torch.float32, memory_format=cl)
        if not sets:
            grad_image.zero_()
        else:
            arr = (_lib.CropSet * len(sets))(*sets)
            nbytes = sum(4 * Cc * R * st.crop_height * st.crop_width * (2 if st.grads2 else 1) for st in sets) + 20 * R * len(sets) + 4 * Cc * B * H * W
            with torch.cuda.device(dev), _Timed("crop_bwd_nhwc", im_size=(B, Cc, H, W), crop=(size_b, size_b), boxes=boxes, box_ind=box_ind, alg_bytes=nbytes):
                _lib.check(_lib.lib().fi_crop_and_resize_backward_multi(arr, len(sets), B, H, W, Cc, _lib.ptr(grad_image), 0,
                                                                        _lib.lib().fi_get_deterministic(), _lib.stream_ptr(dev)))
        # gradient w.r.t. the previous contents of the in-place outputs: the rows this call overwrote get none
        prev = []
        for pos, gg in ((4, g_a), (5, g_b)):
            if gg is not None and ctx.needs_input_grad[pos]:
                gg = gg.clone()
                gg.index_fill_(0, dst_row.long(), 0.0)
                prev.append(gg)
            else:
                prev.append(None)
        return grad_image, None, None, None, prev[0], prev[1], None, None, None


def crop_pair(image, boxes, box_ind, dst_row, out_a, size_a, out_b, size_b, compact_b=False):
    """See _CropPair.  Returns (out_a, out_b[, compact_b])."""
    return _CropPair.apply(image, boxes, box_ind, dst_row, out_a, out_b, int(size_a), int(size_b), bool(compact_b))


# ---- backward planning at forward time ---------------------------------------------------------------------------------
# The per-tile sample lists of the backward (tile_prep + bin_enumerate, ~15 % of its time) depend on the boxes only.  They
# are built right after the forward launch, on a side stream, into a torch-allocated workspace that the autograd node keeps:
# the work overlaps with whatever runs between forward and backward (critic, loss head) and the backward proper is
# tile_collapse + accumulate.  FI_PLAN_AT_FORWARD=0 (or plan_at_forward(False)) builds the lists inside backward instead.
_PLAN = {"at_forward": os.environ.get("FI_PLAN_AT_FORWARD", "1") != "0", "side_stream": os.environ.get("FI_PLAN_STREAM", "side") == "side"}
_SIDE_STREAMS = {}


def plan_at_forward(on=True, side_stream=True):
    old = dict(_PLAN)
    _PLAN["at_forward"], _PLAN["side_stream"] = bool(on), bool(side_stream)
    return old


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


def _bwd_sets(keep, live, dual, g_images=None, grads=None):
    """(BwdSet array, indices of the planned sets).  Without gradients (plan time) maps and second sources are named by
    distinct aligned placeholders: only WHICH sets share a map / have two sources matters there."""
    sets = []
    for k, kp in enumerate(keep):
        if not live[k]:
            continue
        B, Cc, H, W = kp["im_size"]
        P = kp["crop"][0]
        if grads is None:
            gi, g1, g2 = 16 * (kp["image"] + 1), 16, (16 if dual[k] else None)
        else:
            gi, (g1, g2) = _lib.ptr(g_images[kp["image"]]), grads[k]
        rows = kp["dst_row"] if kp["scattered"] else None
        sets.append(_lib.BwdSet(gi, g1, g2, _lib.ptr(kp["boxes"]), _lib.ptr(kp["box_ind"]), _lib.ptr(rows), B, H, W, Cc, kp["cap"], P, P,
                                _lib.ptr(kp["count"])))
    return (_lib.BwdSet * len(sets))(*sets) if sets else None


class _BwdPlan(object):
    """Workspace + opaque fi_bwd_plan of one crop_sets pass, and the event that says the lists are built."""

    def __init__(self, keep, live, dual, exact, max_entries, dev):
        self.live, self.dual, self.exact, self.ok = list(live), list(dual), int(exact), False
        self.ws, self.event = None, None
        arr = _bwd_sets(keep, live, dual)
        if arr is None:
            return
        L = _lib.lib()
        nbytes = L.fi_crop_sets_backward_workspace(arr, len(arr), self.exact, int(max_entries))
        if nbytes == 0:
            return                                      # these sets take the reduction kernels: nothing to plan
        self.ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        self.plan = _lib.BwdPlan()
        cur = torch.cuda.current_stream(dev)
        side = _side_stream(dev) if _PLAN["side_stream"] else cur
        if side is not cur:
            side.wait_stream(cur)                       # boxes / counts come from kernels queued on the current stream
            self.ws.record_stream(side)
        with torch.cuda.device(dev), _Timed("crop_bwd_lists", stream=side, alg_bytes=0):
            rc = L.fi_crop_sets_backward_plan(arr, len(arr), self.exact, int(max_entries), _lib.ptr(self.ws), nbytes, C.byref(self.plan),
                                              side.cuda_stream)
        if rc == -3:
            return
        _lib.check(rc)
        if side is not cur:
            self.event = torch.cuda.Event()
            self.event.record(side)
        self.n = len(arr)
        self.ok = True

    def overflowed(self, dev):
        return _lib.lib().fi_crop_sets_backward_overflow(C.byref(self.plan), _lib.stream_ptr(dev)) == 1


class _CropSets(torch.autograd.Function):
    """Every RoIAlign of one Dev.forward pass as ONE autograd node and ONE launch each way (fi_crop_sets_forward /
    fi_crop_sets_backward_plan + _run): the per-level calls of lib/sub_module.py:429-600 are far too small to fill a B200 one
    at a time (a few hundred boxes on the 26x42 map of P5).  apply(plan, *images, *outs)."""

    @staticmethod
    def forward(ctx, plan, *tensors):
        n_img, n_out = plan["n_img"], plan["n_out"]
        images, outs = list(tensors[:n_img]), list(tensors[n_img:n_img + n_out])
        cl = torch.channels_last
        dev = images[0].device
        fixed = []
        for im in images:
            _lib.require_cuda(im)
            layout, im = _lib.layout_of(im)
            if layout != _lib.FI_LAYOUT_NHWC or im.dtype != torch.float32 or im.size(1) % 128 != 0:
                raise _lib.FiError("crop_sets needs fp32 channels_last feature maps with C % 128 == 0")
            fixed.append(im)
        images = fixed
        arr = (_lib.FwdSet * len(plan["sets"]))()
        compacts, keep = [], []
        for k, sp in enumerate(plan["sets"]):
            im = images[sp["image"]]
            B, Cc, H, W = im.shape
            boxes, box_ind = _prep_boxes(sp["boxes"], sp["box_ind"], dev)
            R, P = boxes.size(0), sp["size"]
            count = sp.get("count")
            if count is not None:
                count = count.detach().to(device=dev, dtype=torch.int32).reshape(-1)[:1].contiguous()
            dst_row = None if sp["dst_row"] is None else sp["dst_row"].to(device=dev, dtype=torch.int32).contiguous()
            out = outs[sp["out"]] if sp["out"] is not None else None
            if out is not None and not (out.is_contiguous(memory_format=cl) and tuple(out.shape[1:]) == (Cc, P, P) and out.dtype == torch.float32):
                raise _lib.FiError("crop_sets: `out` of set %d must be fp32 channels_last [rows,%d,%d,%d]" % (k, Cc, P, P))
            comp = torch.empty((R, Cc, P, P), device=dev, memory_format=cl) if (sp["compact"] or out is None) else None
            compacts.append(comp)
            primary, second = (out, comp) if out is not None else (comp, None)
            arr[k] = _lib.FwdSet(_lib.ptr(im), _lib.ptr(boxes), _lib.ptr(box_ind), _lib.ptr(dst_row) if out is not None else None,
                                 _lib.ptr(primary), _lib.ptr(second), B, H, W, Cc, R, P, P, float(sp.get("extrapolation", 0.0)), _lib.ptr(count))
            keep.append(dict(boxes=boxes, box_ind=box_ind, dst_row=dst_row, im_size=(B, Cc, H, W), crop=(P, P), dual=(out is not None and comp is not None),
                             img_offsets=sp.get("img_offsets"), image=sp["image"], cap=R, count=count, scattered=out is not None,
                             out=sp["out"]))
        with torch.cuda.device(dev), _Timed("crop_fwd_nhwc", sets=keep):
            _lib.check(_lib.lib().fi_crop_sets_forward(arr, len(plan["sets"]), _lib.stream_ptr(dev)))
        ctx.mark_dirty(*outs)
        ctx.set_materialize_grads(False)
        ctx.plan, ctx.keep = plan, keep
        ctx.comp_slots = [k for k, c in enumerate(compacts) if c is not None]
        ctx.bwd = None
        if _PLAN["at_forward"] and any(ctx.needs_input_grad[1:1 + n_img]) and _lib.get_option("bwd_form") != 3:
            # assume every output will receive a gradient (backward re-plans if not)
            live = [kp["cap"] > 0 for kp in keep]
            bp = _BwdPlan(keep, live, [kp["dual"] for kp in keep], _lib.lib().fi_get_deterministic(), plan.get("max_entries", 0), dev)
            ctx.bwd = bp if bp.ok else None
        return tuple(outs) + tuple(c for c in compacts if c is not None)

    @staticmethod
    def backward(ctx, *grads):
        plan, keep = ctx.plan, ctx.keep
        n_img, n_out = plan["n_img"], plan["n_out"]
        cl = torch.channels_last
        g_outs = [None if g is None else g.contiguous(memory_format=cl) for g in grads[:n_out]]
        g_comp = {k: (None if g is None else g.contiguous(memory_format=cl)) for k, g in zip(ctx.comp_slots, grads[n_out:])}
        dev = keep[0]["boxes"].device
        # one standalone tensor per dense gradient map (autograd can then hand it to the leaf without a copy); the kernels
        # write every pixel of every named map once (the reduction fallback zero-fills first)
        sizes = {kp["image"]: kp["im_size"] for kp in keep}
        g_images = {i: torch.empty(sizes[i], device=dev, dtype=torch.float32, memory_format=cl) for i in sorted(sizes)}
        live, dual, srcs, items = [], [], [], []
        for k, kp in enumerate(keep):
            gs = g_outs[kp["out"]] if kp["out"] is not None else None
            gc = g_comp.get(k)
            if kp["cap"] == 0 or (gs is None and gc is None):
                live.append(False); dual.append(False); srcs.append(None)
                continue
            live.append(True)
            if gs is not None:
                dual.append(gc is not None); srcs.append((_lib.ptr(gs), _lib.ptr(gc)))
            else:
                dual.append(False); srcs.append((_lib.ptr(gc), None))
                if kp["scattered"]:                      # only the compact copy received a gradient: its rows are in box order
                    kp = dict(kp, scattered=False)
                    keep = keep[:k] + [kp] + keep[k + 1:]
            items.append((kp["im_size"][1], kp["crop"][0], dual[-1], kp["cap"], kp["count"]))
        touched = {keep[k]["image"] for k in range(len(keep)) if live[k]}
        meta = dict(items=items, map_bytes=4 * sum(sizes[i][0] * sizes[i][1] * sizes[i][2] * sizes[i][3] for i in touched))
        if any(live):
            L = _lib.lib()
            arr = _bwd_sets(keep, live, dual, g_images, srcs)
            exact = L.fi_get_deterministic()
            bp = ctx.bwd
            usable = bp is not None and bp.live == live and bp.dual == dual and bp.exact == exact and keep is ctx.keep \
                and _lib.get_option("bwd_form") != 3
            if not usable and _lib.get_option("bwd_form") != 3:
                old = plan_at_forward(True, side_stream=False)         # lists built here, on this stream
                try:
                    bp = _BwdPlan(keep, live, dual, exact, plan.get("max_entries", 0), dev)
                finally:
                    _PLAN.update(old)
                usable = bp.ok
            with torch.cuda.device(dev), _Timed("crop_bwd_nhwc", bwd=meta):
                if usable:
                    if bp.event is not None:
                        torch.cuda.current_stream(dev).wait_event(bp.event)
                    _lib.check(L.fi_crop_sets_backward_run(C.byref(bp.plan), arr, len(arr), 1, _lib.stream_ptr(dev)))
                else:                                                  # reduction kernels (shapes the tile-owner form does not take)
                    _lib.check(L.fi_crop_sets_backward(arr, len(arr), 1, _lib.stream_ptr(dev)))
        for i, gi in g_images.items():
            if i not in touched:
                gi.zero_()                       # a map none of whose crops received a gradient
        # gradient w.r.t. the previous contents of the in-place outputs (only when somebody asks for it: Dev hands in fresh
        # tensors): the rows the sets overwrote no longer depend on them
        g_prev = []
        for o in range(n_out):
            if g_outs[o] is None or not ctx.needs_input_grad[1 + n_img + o]:
                g_prev.append(None)
                continue
            gp = g_outs[o].clone()
            for kp in ctx.keep:
                if kp["out"] == o and kp["dst_row"] is not None and kp["cap"] > 0:
                    if kp["count"] is None:
                        gp.index_fill_(0, kp["dst_row"].long(), 0.0)
                    else:                                # device-side length: only the live rows were written
                        livemask = torch.arange(kp["cap"], device=dev) < kp["count"].to(dev)
                        gp.index_put_((kp["dst_row"].long()[livemask],), torch.zeros((), device=dev))
            g_prev.append(gp)
        return (None,) + tuple(g_images.get(i) for i in range(n_img)) + tuple(g_prev)


def crop_sets(specs, max_entries=0):
    """specs: list of dict(image=Tensor, boxes=[R,4], box_ind=[R], size=int, out=Tensor|None, dst_row=[R]|None, compact=bool).
    A set with ``out`` writes crop r into row ``dst_row[r]`` of ``out`` (several sets may share one ``out``) and, with
    ``compact``, also returns the compact [R,C,size,size] crop; a set without ``out`` returns the compact crop only.
    ``count`` (optional, a device int32): the actual number of boxes of the set, ``boxes.size(0)`` then being the capacity of the
    lists -- no host ever needs to know the list lengths (fixed shapes, CUDA-graph capturable); rows past the count are not written.
    ``max_entries``: see fi_crop_sets_backward_plan (a caller's tighter bound on the backward's sample lists).
    Returns (outs_by_spec, compacts_by_spec): per spec the (updated) ``out`` or None, and the compact crop or None."""
    images, outs, sets = [], [], []

    def slot(lst, t):
        for i, u in enumerate(lst):
            if u is t:
                return i
        lst.append(t)
        return len(lst) - 1
    for sp in specs:
        sets.append(dict(image=slot(images, sp["image"]), out=(slot(outs, sp["out"]) if sp.get("out") is not None else None),
                         boxes=sp["boxes"], box_ind=sp["box_ind"], dst_row=sp.get("dst_row"), size=int(sp["size"]),
                         compact=bool(sp.get("compact", False)), extrapolation=float(sp.get("extrapolation", 0.0)),
                         img_offsets=sp.get("img_offsets"), count=sp.get("count")))
        if sets[-1]["out"] is not None and sets[-1]["dst_row"] is None:
            raise _lib.FiError("crop_sets: a set with `out` needs `dst_row`")
    plan = dict(n_img=len(images), n_out=len(outs), sets=sets, max_entries=int(max_entries))
    res = _CropSets.apply(plan, *images, *outs)
    new_outs, comps = res[:len(outs)], list(res[len(outs):])
    out_by_spec = [new_outs[st["out"]] if st["out"] is not None else None for st in sets]
    comp_by_spec = [comps.pop(0) if (st["compact"] or st["out"] is None) else None for st in sets]
    return out_by_spec, comp_by_spec


def set_deterministic(on=True):
    """Exact RoIAlign backward (channels_last, C % 128 == 0, crops <= 16x16): every pixel is summed by one warp in the order
    AND with the un-fused arithmetic of the reference's serial CPU loop -- bit-identical to crop_and_resize.c:157-252.
    The default mode of the tile-owner kernel sums in the same (run-to-run deterministic) order with packed FMAs.
    Returns the previous setting."""
    return bool(_lib.lib().fi_set_deterministic(1 if on else 0))


def crop_and_resize(image, boxes, box_ind, crop_h, crop_w, extrapolation_value=0.0, out=None, dst_row=None):
    """Functional form.  ``out``/``dst_row`` fuse the scatter-back of lib/sub_module.py:645-662: crop ``r`` is
    written into row ``dst_row[r]`` of the preallocated ``out`` (rows not named by ``dst_row`` are left untouched)."""
    return _CropAndResize.apply(image, boxes, box_ind, int(crop_h), int(crop_w), float(extrapolation_value), out, dst_row)


class CropAndResizeFunction(object):
    """Old-style call shape of lib/roi_align/crop_and_resize.py:14-54 on top of a modern static Function."""

    def __init__(self, crop_height, crop_width, extrapolation_value=0):
        self.crop_height = crop_height
        self.crop_width = crop_width
        self.extrapolation_value = extrapolation_value

    def __call__(self, image, boxes, box_ind):
        return crop_and_resize(image, boxes, box_ind, self.crop_height, self.crop_width, self.extrapolation_value)


class RoIAlign(nn.Module):
    """lib/roi_align/roi_align.py:6-48: boxes are (x1,y1,x2,y2) in pixels of the feature map."""

    def __init__(self, crop_height, crop_width, extrapolation_value=0, transform_fpcoor=True):
        super().__init__()
        self.crop_height = crop_height
        self.crop_width = crop_width
        self.extrapolation_value = extrapolation_value
        self.transform_fpcoor = transform_fpcoor

    def forward(self, featuremap, boxes, box_ind):
        x1, y1, x2, y2 = torch.split(boxes, 1, dim=1)
        image_height, image_width = featuremap.size()[2:4]
        if self.transform_fpcoor:
            spacing_w = (x2 - x1) / float(self.crop_width)
            spacing_h = (y2 - y1) / float(self.crop_height)
            nx0 = (x1 + spacing_w / 2 - 0.5) / float(image_width - 1)
            ny0 = (y1 + spacing_h / 2 - 0.5) / float(image_height - 1)
            nw = spacing_w * float(self.crop_width - 1) / float(image_width - 1)
            nh = spacing_h * float(self.crop_height - 1) / float(image_height - 1)
            boxes = torch.cat((ny0, nx0, ny0 + nh, nx0 + nw), 1)
        else:
            boxes = torch.cat((y1 / float(image_height - 1), x1 / float(image_width - 1),
                               y2 / float(image_height - 1), x2 / float(image_width - 1)), 1)
        return crop_and_resize(featuremap, boxes.detach().contiguous(), box_ind.detach(),
                               self.crop_height, self.crop_width, self.extrapolation_value)


def crop_taps(boxes, image_height, image_width, crop_h, crop_w):
    """Integer bilinear taps [R,crop_h,crop_w,5] = (y_lo,y_hi,x_lo,x_hi,inside) computed on the device."""
    _lib.require_cuda(boxes)
    boxes = boxes.detach().to(dtype=torch.float32).contiguous()
    taps = torch.empty((boxes.size(0), crop_h, crop_w, 5), device=boxes.device, dtype=torch.int32)
    with torch.cuda.device(boxes.device):
        _lib.check(_lib.lib().fi_crop_taps(_lib.ptr(boxes), boxes.size(0), image_height, image_width, crop_h, crop_w,
                                           _lib.ptr(taps), _lib.stream_ptr(boxes.device)))
    return taps
