#!/usr/bin/env python
"""Benchmark of the Feature Intertwiner hot path on B200 (BASELINE.json metric: RoIs/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c5] [--impl ours|reference]
                    [--mode graph|eager] [--with-critic] [--no-other-workloads]

One "step" = one pass of the hot path over one batch of synthetic input:

    [c3: proposal layer -- decode + on-device NMS producing the RoIs]  ->  level rule + reliable/less-reliable split
    ->  every RoIAlign call of Dev.forward (big 14x14 on the raw level maps, small 7x7 + 14x14 on the made-up maps;
    crops written straight to their final rows)  ->  per-class segment means of the critic features  ->  statistics merge
    (+ all-reduce)  ->  buffer update + class match  ->  OptTrans / Sinkhorn(L) loss  ->  backward of all of it
    (Sinkhorn gradient, segment-mean backward, RoIAlign backward of every crop).

Nothing in the step is read back by the host: list lengths stay on the device (fixed-capacity lists), so the WHOLE step
is captured in one CUDA graph (`--mode graph`, default) and replayed; `--mode eager` enqueues it launch by launch.

The make-up conv and the critic convs are stock cuDNN.  In the default step their outputs (made-up
maps, critic features) and the upstream crop gradients are synthetic inputs; `--with-critic` times the a5-inclusive variant
instead: fi.Dev.forward (make-up conv + all crops + critic) + fi.IntertwinerLoss + backward, called directly.

`value`  : RoIs/s with all inputs resident in HBM (device-timed, CUDA events, max over ranks).
`e2e`    : the same step starting from pinned HOST buffers (H2D of every input, D2H of the loss) inside the timed region.
`--impl reference`: the reference's own CPU implementation (oracle/_ref: lib/roi_align/src/crop_and_resize.c compiled
           unmodified, OpenMP forward + serial backward) on every crop of the step -- FULL steps, nothing sampled --
           plus the torch-CPU port of lib/OT_module.py for the loss (oracle/pyref.py).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FEAT = 1024
NCLS = 81
DEPTH = 256
METRIC = "RoIs/sec (RoIAlign fwd+bwd + split + class means + Sinkhorn intertwiner loss, fwd+bwd)"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(name, wl):
    extra = ", proposal layer + NMS in the step" if wl.get("proposals") else ""
    return "%s: batch %d/GPU, %dx%d, %d RoIs/img, FPN P2-P5 C=256, pools 7+14, Sinkhorn N=256 L=%d, class-level OT loss%s" % (
        name, wl["batch"], wl["image"][0], wl["image"][1], wl["rois_per_image"], wl["sinkhorn_iters"], extra)


def bind_to_gpu_numa(local):
    """Pin this process (and, by first touch, its pinned host buffers) to the CPUs next to its GPU -- NVML's ideal CPU affinity of
    the device, or the NUMA node sysfs names for its PCI function: with every rank's buffers on node 0 the e2e leg of ranks 4-7
    crossed the socket link (round 1: 77 -> 180 ms/step from 1 to 8 GPUs)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[local]) if visible and visible.split(",")[local].isdigit() else local
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        allowed_now = set(os.sched_getaffinity(0))
        cpus, how = [], None
        try:
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cpus = [64 * i + bit for i, wd in enumerate(words) for bit in range(64) if (int(wd) >> bit) & 1]
            how = "nvml cpu affinity"
        except Exception:
            cpus = []
        if not cpus or set(cpus) >= allowed_now:
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
            if len(bus.split(":")[0]) == 8:
                bus = bus[4:]                               # nvml: 00000000:1b:00.0, sysfs: 0000:1b:00.0
            node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
            if node < 0:
                return {"bound": False, "why": "no NUMA information for the GPU (nvml affinity = all CPUs, sysfs numa_node = -1)"}
            cpus = []
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus += list(range(int(lo), int(hi or lo) + 1))
            how = "sysfs numa_node %d" % node
        allowed = sorted(set(cpus) & allowed_now)
        if allowed and len(allowed) < len(allowed_now):
            os.sched_setaffinity(0, allowed)
            return {"bound": True, "how": how, "cpus": len(allowed)}
        return {"bound": False, "why": "the GPU's CPU set is every CPU this process may use"}
    except Exception as exc:        # noqa: BLE001
        return {"bound": False, "why": repr(exc)[:120]}


class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is queried in-process from
    a thread (every 10 ms; a K-step region lasts tens of ms, an `nvidia-smi -lms 100` loop would see 0-1 samples of it and
    its per-loop device enumeration stalls kernel launches); `nvidia-smi` is the fallback when pynvml is unavailable."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False
        self.mode = os.environ.get("FI_SAMPLER", "nvml")

    def start(self):
        if self.mode == "none":
            return
        if self.mode == "nvml":
            try:
                import pynvml
                pynvml.nvmlInit()
                visible = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
                h = pynvml.nvmlDeviceGetHandleByIndex(phys)
                self.nvml = (pynvml, h)
                self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
                self.thread = threading.Thread(target=self._poll, daemon=True)
                self.thread.start()
                return
            except Exception:
                self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        pynvml, h = self.nvml
        bits = [getattr(pynvml, n, 0) for n in ("nvmlClocksThrottleReasonHwSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown",
                                                "nvmlClocksThrottleReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwPowerCap")]
        while not self.stop_flag:
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(self.sm_max)] + ["Active" if (b and (r & b)) else "Not Active" for b in bits])
            except Exception:
                pass
            time.sleep(0.01)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
        elif self.proc is not None:
            self.proc.terminate()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = sorted({self.NAMES[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "how": "nvml thread, 10 ms" if self.nvml is not None else "nvidia-smi -lms 100"}


# =================================================================================================== inputs
def anchors_for(hw, gen):
    """RPN anchors of the five pyramid levels (3 ratios per cell, scales 32..512, lib/config.py:90-91,318) in pixels."""
    out = []
    for stride, scale in zip((4, 8, 16, 32, 64), (32, 64, 128, 256, 512)):
        h, w = -(-hw[0] // stride), -(-hw[1] // stride)
        ys = (torch.arange(h, dtype=torch.float32) * stride).view(h, 1, 1).expand(h, w, 3)
        xs = (torch.arange(w, dtype=torch.float32) * stride).view(1, w, 1).expand(h, w, 3)
        ratios = torch.tensor([0.5, 1.0, 2.0]).view(1, 1, 3).expand(h, w, 3)
        hh, ww = scale / ratios.sqrt(), scale * ratios.sqrt()
        out.append(torch.stack([ys - hh / 2, xs - ww / 2, ys + hh / 2, xs + ww / 2], dim=3).reshape(-1, 4))
    return torch.cat(out).contiguous()


def make_inputs(wl, seed):
    """Everything a step consumes, on the HOST, seeded.  The split itself is computed by the step."""
    from feature_intertwiner_b200 import synth
    g = torch.Generator().manual_seed(seed)
    B, R, hw = wl["batch"], wl["rois_per_image"], wl["image"]
    host = {
        "rois": synth.make_rois(B, R, hw, g),
        "gt": synth.make_class_ids(B, R, g, NCLS),
        "raw": synth.make_feature_maps(B, hw, DEPTH, g, channels_last=True),      # P2..P5
        "madeup": synth.make_feature_maps(B, hw, DEPTH, g, channels_last=True),   # upsample(P2..P5): stock conv output
    }
    if wl.get("proposals"):
        # RPN outputs whose proposals ARE the RoIs of the step: anchor a of image b regresses (with noise) onto one of the
        # boxes make_rois drew; the foreground score decays with the anchor's index so that the sort is well defined
        anchors = anchors_for(hw, g)
        A = anchors.size(0)
        tgt = host["rois"] * torch.tensor([hw[0], hw[1], hw[0], hw[1]], dtype=torch.float32)
        pick = torch.randint(0, R, (B, A), generator=g)
        t = torch.gather(tgt, 1, pick.unsqueeze(2).expand(B, A, 4))
        t = torch.where((t[:, :, 2:3] - t[:, :, 0:1]) > 1, t, anchors.unsqueeze(0).expand(B, A, 4))     # zero-padded RoIs: keep the anchor
        ah, aw = anchors[:, 2] - anchors[:, 0], anchors[:, 3] - anchors[:, 1]
        th, tw = (t[:, :, 2] - t[:, :, 0]).clamp(min=1.0), (t[:, :, 3] - t[:, :, 1]).clamp(min=1.0)
        acy, acx = anchors[:, 0] + 0.5 * ah, anchors[:, 1] + 0.5 * aw
        tcy, tcx = t[:, :, 0] + 0.5 * th, t[:, :, 1] + 0.5 * tw
        std = torch.tensor([0.1, 0.1, 0.2, 0.2])
        deltas = torch.stack([(tcy - acy) / ah, (tcx - acx) / aw, torch.log(th / ah), torch.log(tw / aw)], dim=2)
        deltas = (deltas.clamp(-4, 4) + 0.02 * torch.randn(B, A, 4, generator=g)) / std
        fg = torch.rand(B, A, generator=g)
        host["rpn_probs"] = torch.stack([1 - fg, fg], dim=2).contiguous()
        host["rpn_bbox"] = deltas.contiguous()
        host["anchors"] = anchors
    return host


def pin(t):
    return t.pin_memory() if torch.cuda.is_available() else t


def build_config(wl):
    import types
    import numpy as np
    ns = types.SimpleNamespace
    return ns(
        DEV=ns(SWITCH=True, STRUCTURE="beta", BASELINE=False, BUFFER_SIZE=1, LOSS_CHOICE="ot", OT_ONE_DIM_FORM="conv", LOSS_FAC=0.5,
               INST_LOSS=False, FEAT_BRANCH_POOL_SIZE=14, ASSIGN_BOX_ON_ALL_SCALE=False, BIG_FEAT_DETACH=True, UPSAMPLE_FAC=1.0,
               MULTI_UPSAMPLER=False, BIG_SUPERVISE=False, DIS_UPSAMPLER=False, INIT_BUFFER_WEIGHT="scratch"),
        ROIS=ns(METHOD="roi_align", ASSIGN_ANCHOR_BASE=224.0), MRCNN=ns(POOL_SIZE=7, MASK_POOL_SIZE=14),
        RPN=ns(PRE_NMS_LIMIT=6000), DATA=ns(IMAGE_SHAPE=np.array([wl["image"][0], wl["image"][1], 3]), BBOX_STD_DEV=np.array([0.1, 0.1, 0.2, 0.2])),
        DATASET=ns(NUM_CLASSES=NCLS))


PEER = {}      # "stats" / "grads": feature_intertwiner_b200.dist.PeerAllReduce of this process (several ranks only)


def setup_peer_allreduce(dev, world):
    """The two exchanges of a step as this library's peer-memory kernel (csrc/peer_allreduce.cu) instead of NCCL: the packed
    class statistics (0.66 MB, on the critical path) and the OptTrans gradients (15.7 MB, on a side stream under the RoIAlign
    backward).  Collective.  FI_PEER_ALLREDUCE=0 keeps NCCL (and the three-graph step)."""
    if world == 1 or os.environ.get("FI_PEER_ALLREDUCE", "1") == "0":
        return False
    import feature_intertwiner_b200 as fi
    from feature_intertwiner_b200 import dist as fdist
    cfg = build_config(dict(image=(64, 64)))
    n_grads = sum(p.numel() for p in fi.OptTrans(cfg, ch_x=FEAT, L=1).parameters())
    stats = fdist.install_peer_allreduce(2 * (FEAT * NCLS + NCLS), None, dev)
    if not stats.ok:
        PEER["why"] = stats.why
        return False
    grads = fdist.PeerAllReduce(n_grads, None, dev)
    if not grads.ok:
        PEER["why"] = grads.why
        fdist.uninstall_peer_allreduce(None)
        return False
    PEER.update(stats=stats, grads=grads)
    return True


class Step(object):
    """Device-side state + one pass of the hot path through the public API.  Every list keeps its capacity (batch * RoIs per
    image) with its length on the device, so the step has fixed shapes and no host read."""

    def __init__(self, wl, device, world, seed, ot_grad_allreduce=False):
        import feature_intertwiner_b200 as fi
        self.fi, self.wl, self.dev, self.world = fi, wl, device, world
        # The reference computes meta_loss ONCE, on GPU 0, from the gathered statistics (lib/model.py:143-144 "the loss is computed
        # in GPU 0"; lib/workflow.py:191): OptTrans lives on one device and its gradients are never exchanged.  Sharded, every rank
        # evaluates the same loss head on the same all-reduced totals, so every rank already holds the full, bit-identical OptTrans
        # gradient (tests/test_dist_gpu.py; with one position per row the cosine cost is a sign, so that gradient is in fact
        # exactly zero: tools/diag_ot_identity.py) and the path needs ONE exchange, the class statistics.  ot_grad_allreduce=True adds
        # what wrapping OptTrans in DistributedDataParallel would do anyway (a 15.7 MB all-reduce of identical values, overlapped
        # with the RoIAlign backward) -- reported as a variant, not needed for the result.
        self.ot_grad_allreduce = bool(ot_grad_allreduce) and world > 1
        self.spatial_sort = os.environ.get("FI_SPATIAL_SORT", "1") != "0"
        self.cfg = build_config(wl)
        torch.manual_seed(2000)
        self.ot = fi.OptTrans(self.cfg, ch_x=FEAT, L=wl["sinkhorn_iters"]).to(device)
        self.loss_mod = fi.IntertwinerLoss(self.cfg, ot_loss=self.ot, feat_dim=FEAT, distributed=world > 1, ot_padded=True).to(device)
        self.host = make_inputs(wl, seed)
        B, R = wl["batch"], wl["rois_per_image"]
        total = B * R
        self.total = total
        # the split of THIS input tells how many rows of the (capacity-sized) synthetic critic features / upstream gradients are live
        rois_d = self.host["rois"].to(device)
        if wl.get("proposals"):
            with torch.no_grad():
                rois_d, _ = fi.proposal_layer([self.host["rpn_probs"].to(device), self.host["rpn_bbox"].to(device)], R, 0.7,
                                              self.host["anchors"].to(device), self.cfg, static=True)
        split = fi.split_levels(fi.roi_level(rois_d, self.cfg.DATA.IMAGE_SHAPE))
        self.counts = (list(split.small_cnt), list(split.big_cnt))
        g = torch.Generator().manual_seed(seed + 1)
        self.host["small_feat"] = [torch.rand(split.small_cnt[i], FEAT, generator=g) for i in range(3)]
        self.host["big_feat"] = [torch.rand(split.big_cnt[i], FEAT, generator=g) for i in range(3)]
        self.host["g_pooled"] = torch.randn(total, DEPTH, 7, 7, generator=g).contiguous(memory_format=torch.channels_last)
        self.host["g_mask"] = torch.randn(total, DEPTH, 14, 14, generator=g).contiguous(memory_format=torch.channels_last)
        self.host["g_small"] = [torch.randn(split.small_cnt[i], DEPTH, 14, 14, generator=g).contiguous(memory_format=torch.channels_last)
                                for i in range(3)]           # gradient the critic sends back into the compact 14x14 crops
        self.host["g_big"] = [torch.randn(split.big_cnt[i], DEPTH, 14, 14, generator=g).contiguous(memory_format=torch.channels_last)
                              for i in range(3)]
        if wl.get("proposals"):
            del self.host["rois"]                            # the step makes its own RoIs
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self._flat(self.host))
        self.pinned = self._map(self.host, pin)
        # device buffers: per-level tensors at full capacity (only the live rows are ever read), the rest as they are
        cl = torch.channels_last
        self.resident = {}
        for k, v in self.host.items():
            if k in ("small_feat", "big_feat"):
                self.resident[k] = [torch.zeros(total, FEAT, device=device) for _ in v]
            elif k in ("g_small", "g_big"):
                self.resident[k] = [torch.zeros((total, DEPTH, 14, 14), device=device).contiguous(memory_format=cl) for _ in v]
            else:
                self.resident[k] = [t.to(device) for t in v] if isinstance(v, list) else v.to(device)
        self.upload(sync=True)
        # upper bound on the backward's sample-list entries for ANY split of `total` boxes: every RoI is small at exactly one
        # level (7x7 + 14x14, the latter with two gradient sources) and big at no more than three (14x14)
        self.max_entries = total * (4 * 49 + 8 * 196 + 12 * 196 + 32 * 11)
        self.grad_bucket = None
        if self.ot_grad_allreduce:
            from feature_intertwiner_b200.dist import GradAllReduce
            self.grad_bucket = GradAllReduce(self.ot.parameters())
        self.graph, self.graph_loss, self.graph_error = None, None, None
        self.launches_per_step = None

    @staticmethod
    def _map(d, f):
        return {k: ([f(t) for t in v] if isinstance(v, list) else f(v)) for k, v in d.items()}

    @staticmethod
    def _flat(d):
        for v in d.values():
            for t in (v if isinstance(v, list) else [v]):
                yield t

    def upload(self, sync=False):
        """H2D of every input of the step from pinned host memory into the step's device buffers (the e2e leg)."""
        for k, v in self.pinned.items():
            for src, dst in zip(v if isinstance(v, list) else [v], self.resident[k] if isinstance(v, list) else [self.resident[k]]):
                dst[: src.size(0)].copy_(src, non_blocking=not sync)

    def loss_only(self):
        """The intertwiner loss alone (BASELINE.json's second metric): statistics merge (+ all-reduce) -> buffer update ->
        class match -> OptTrans / Sinkhorn, forward and backward, on the class statistics of the last step."""
        bf, bc, sf, sc = self.last_feat_in
        loss = self.loss_mod([bf, bc, sf.requires_grad_(), sc, None, None]).sum()
        loss.backward()
        for p in self.ot.parameters():
            p.grad = None
        sf.grad = None
        return loss

    def capture_loss_only(self):
        """loss_only() -- forward and backward -- as ONE CUDA graph, so that the loss-only figure is device time like the step's
        (launch by launch it measures Python: ~40 launches at ~15 us of host time each).  Single process only (no collective in a
        capture here).  Returns the replay callable, or None when the capture failed."""
        if self.world > 1:
            return None
        try:
            cur = torch.cuda.current_stream(self.dev)
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                for _ in range(3):
                    self.loss_only()
            cur.wait_stream(s)
            torch.cuda.synchronize(self.dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.loss_only()
            torch.cuda.synchronize(self.dev)
            self.loss_graph = graph
            return graph.replay
        except Exception as exc:            # noqa: BLE001 -- fall back to the launch-by-launch timing
            self.loss_graph, self.loss_graph_error = None, repr(exc)[:300]
            torch.cuda.synchronize(self.dev)
            return None

    # ---- the step in three parts (one process: run() chains them; several: the two all-reduces sit between them) ----------
    def forward_part(self, inp):
        """[proposal layer] -> level rule + split -> every crop -> class means -> this rank's packed un-normalised statistics."""
        fi, cfg = self.fi, self.cfg
        B, R = self.wl["batch"], self.wl["rois_per_image"]
        total = B * R
        gt = inp["gt"]
        if self.wl.get("proposals"):
            # producer of the RoIs (lib/layers.py:71-139): decode + batched on-device NMS + gather, no host read
            with torch.no_grad():
                rois, _ = fi.proposal_layer([inp["rpn_probs"], inp["rpn_bbox"]], R, 0.7, inp["anchors"], cfg, static=True)
        else:
            rois = inp["rois"]
        # fresh leaves every step (a training loop clears .grad each iteration; re-using the leaf would time an extra
        # read-modify-write of every map in AccumulateGrad)
        raw = [m.detach().requires_grad_() for m in inp["raw"]]
        madeup = [m.detach().requires_grad_() for m in inp["madeup"]]
        small_f = [t.detach().requires_grad_() for t in inp["small_feat"]]
        big_f = inp["big_feat"]
        split = fi.split_levels(fi.roi_level(rois, cfg.DATA.IMAGE_SHAPE, cfg.ROIS.ASSIGN_ANCHOR_BASE), rois=rois, gt=gt,
                                order=fi.spatial_order(rois) if self.spatial_sort else None, sync=False)
        pooled_out = torch.empty((total, DEPTH, 7, 7), device=self.dev, memory_format=torch.channels_last)
        mask_out = torch.empty((total, DEPTH, 14, 14), device=self.dev, memory_format=torch.channels_last)
        # every crop of the pass in one level-batched launch (fi.crop_sets), like Dev.forward
        specs, where = [], {}
        for i in range(4):
            if i < 3:
                where[("big", i)] = len(specs)
                specs.append(dict(image=raw[i], boxes=split.big_boxes(i), box_ind=split.big_ind(i), size=14, count=split.big_count(i)))
            s32 = split.small(i)
            boxes, ind = split.small_boxes(i), split.small_ind(i)
            where[("small", i)] = len(specs)
            specs.append(dict(image=madeup[i], boxes=boxes, box_ind=ind, size=7, out=pooled_out, dst_row=s32, count=split.small_count(i)))
            specs.append(dict(image=madeup[i], boxes=boxes, box_ind=ind, size=14, out=mask_out, dst_row=s32, compact=(i < 3), count=split.small_count(i)))
        res_out, res_comp = fi.crop_sets(specs, max_entries=self.max_entries)
        outs, grads, lists = [], [], []
        for i in range(3):
            k = where[("big", i)]
            outs.append(res_comp[k]); grads.append(inp["g_big"][i])          # compact 14x14 crop -> critic (stock conv, not timed)
            k = where[("small", i)]
            outs.append(res_comp[k + 1]); grads.append(inp["g_small"][i])
        # class means of the 3 reliable + 3 less-reliable lists: one launch each way (lib/sub_module.py:664-684)
        for i in range(3):
            lists.append((split.big_gt(i), big_f[i], split.big_count(i)))
        for i in range(3):
            lists.append((split.small_gt(i), small_f[i], split.small_count(i)))
        stats = fi.assign_feat2cls_multi(lists, NCLS)
        bfeat, bcnt = [stats[i][0] for i in range(3)], [stats[i][1] for i in range(3)]
        sfeat, scnt = [stats[3 + i][0] for i in range(3)], [stats[3 + i][1] for i in range(3)]
        pooled_out, mask_out = res_out[where[("small", 3)]], res_out[where[("small", 3)] + 1]
        feat_in = [torch.stack(bfeat)[None].detach(), torch.stack(bcnt)[None], torch.stack(sfeat)[None], torch.stack(scnt)[None], None, None]
        self.last_feat_in = [t.detach() for t in feat_in[:4]]          # for the loss-head-only timing
        self.last_split = split
        return dict(feat_in=feat_in, heads=[pooled_out, mask_out] + outs, head_grads=[inp["g_pooled"], inp["g_mask"]] + grads,
                    leaves=raw + madeup + small_f)

    def run(self, inp=None):
        """One process: the whole step through the public API, autograd end to end."""
        inp = self.resident if inp is None else inp
        st = self.forward_part(inp)
        loss = self.loss_mod(st["feat_in"]).sum()
        torch.autograd.backward([loss] + st["heads"], [torch.ones_like(loss)] + st["head_grads"])
        if self.ot_grad_allreduce:
            # gradient all-reduce of the path's own parameters (OptTrans, 15.7 MB): started by a hook as soon as the loss head's
            # backward has produced them, i.e. overlapped with the RoIAlign backward that follows it on the compute stream
            if self.grad_bucket is not None:
                self.grad_bucket.finish()
            else:
                import torch.distributed as dist
                dist.all_reduce(torch.cat([p.grad.reshape(-1) for p in self.ot.parameters()]))
        for p in self.ot.parameters():
            p.grad = None
        return loss

    def capture(self):
        """The whole step (side-stream list building included) as ONE CUDA graph.  Returns True when the graph is live."""
        lib = self.fi.lib()
        try:
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                for _ in range(3):
                    self.run()
            torch.cuda.current_stream(self.dev).wait_stream(s)
            torch.cuda.synchronize(self.dev)
            graph = torch.cuda.CUDAGraph()
            n0 = lib.fi_kernel_launches()
            with torch.cuda.graph(graph):
                self.graph_loss = self.run()
            self.launches_per_step = int(lib.fi_kernel_launches() - n0)
            self.graph = graph
        except Exception as exc:            # noqa: BLE001 -- capture is an optimisation: fall back to launch-by-launch
            self.graph, self.graph_error = None, repr(exc)[:300]
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            # every rank replays or none does: the peer all-reduce kernels inside the graph must stay matched
            import torch.distributed as dist
            flag = torch.tensor([1 if self.graph is not None else 0], device=self.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self.graph = None
                self.graph_error = self.graph_error or "capture failed on another rank"
        if self.graph is not None:
            self.graph.replay()
            torch.cuda.synchronize(self.dev)
        return self.graph is not None

    # ---- several processes: three graphs, the two all-reduces eager between them (no collective inside a capture) ---------
    def _packed_local(self, st):
        from feature_intertwiner_b200.dist import merged_class_sums_pair
        bs, bn, ss, sn = merged_class_sums_pair(st["feat_in"][0], st["feat_in"][1], st["feat_in"][2], st["feat_in"][3], None, False)
        return torch.cat([bs.reshape(-1), bn, ss.reshape(-1), sn])

    def _head_from_packed(self, packed):
        a, b = FEAT * NCLS, NCLS
        big_sum, big_n = packed[:a].view(FEAT, NCLS).detach(), packed[a:a + b].detach()
        s_sum, s_n = packed[a + b:2 * a + b].view(FEAT, NCLS), packed[2 * a + b:].detach()
        return self.loss_mod._head(big_sum, big_n, s_sum, s_n, None, None).sum()

    def capture_segmented(self):
        """Graph A: forward part -> packed local statistics.  [all-reduce]  Graph B: loss head forward + backward -> gradient of
        the packed statistics, OptTrans parameter gradients (flat).  [async all-reduce of those, overlapping graph C]  Graph C:
        backward of the forward part (class-mean backward, RoIAlign backward of every crop)."""
        import torch.distributed as dist
        lib = self.fi.lib()
        ok = True
        try:
            if self.grad_bucket is not None:
                self.grad_bucket.remove()
                self.grad_bucket = None
            params = [p for p in self.ot.parameters()]
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                for _ in range(3):                                  # warm-up of every op on a side stream (allocator, lazy init)
                    st = self.forward_part(self.resident)
                    packed = self._packed_local(st)
                    pin_ = packed.detach().clone().requires_grad_()
                    loss = self._head_from_packed(pin_)
                    gs = torch.autograd.grad(loss, [pin_] + params)
                    torch.autograd.grad([packed] + st["heads"], st["leaves"], [gs[0]] + st["head_grads"], allow_unused=True)
            torch.cuda.current_stream(self.dev).wait_stream(s)
            torch.cuda.synchronize(self.dev)
            n0 = lib.fi_kernel_launches()
            pool = torch.cuda.graph_pool_handle()
            gA, gB, gC = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(gA, pool=pool):
                st = self.forward_part(self.resident)
                node = st["heads"][0].grad_fn                       # the crop_sets node: join its side-stream list building HERE
                if getattr(node, "bwd", None) is not None and node.bwd.event is not None:
                    torch.cuda.current_stream(self.dev).wait_event(node.bwd.event)
                    node.bwd.event = None
                packed = self._packed_local(st)
                self.seg_packed = packed.detach().clone()           # all-reduced in place between the graphs
            self.seg_in = self.seg_packed.clone().requires_grad_()
            with torch.cuda.graph(gB, pool=pool):
                self.seg_in.data.copy_(self.seg_packed)
                loss = self._head_from_packed(self.seg_in)
                gs = torch.autograd.grad(loss, [self.seg_in] + params)
                self.graph_loss = loss.detach()
                self.seg_gpacked = gs[0] * float(self.world)        # DataParallel sums replica gradients (dist.py::_AllReduceSum)
                self.seg_flat = torch.cat([g_.reshape(-1) for g_ in gs[1:]])
            with torch.cuda.graph(gC, pool=pool):
                self.seg_grads = torch.autograd.grad([packed] + st["heads"], st["leaves"], [self.seg_gpacked] + st["head_grads"], allow_unused=True)
            self.launches_per_step = int(lib.fi_kernel_launches() - n0)
            self.seg = (gA, gB, gC)
            self._seg_state = st                                     # keeps the autograd graph of A alive
        except Exception as exc:            # noqa: BLE001
            ok = False
            self.graph_error = repr(exc)[:300]
        torch.cuda.synchronize(self.dev)
        flag = torch.tensor([1 if ok else 0], device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)                 # every rank or none: the collectives must stay matched
        if int(flag.item()) == 0:
            self.seg = None
            if self.grad_bucket is None and self.ot_grad_allreduce:
                from feature_intertwiner_b200.dist import GradAllReduce
                self.grad_bucket = GradAllReduce(self.ot.parameters())
            return False
        self.step_segmented()
        torch.cuda.synchronize(self.dev)
        return True

    def step_segmented(self):
        import torch.distributed as dist
        from feature_intertwiner_b200.dist import all_reduce_sum_
        gA, gB, gC = self.seg
        gA.replay()
        all_reduce_sum_(self.seg_packed)                            # class statistics of both sets, one collective (0.66 MB)
        gB.replay()
        work = dist.all_reduce(self.seg_flat, async_op=True) if self.ot_grad_allreduce else None   # OptTrans gradients (15.7 MB): overlaps graph C
        gC.replay()
        if work is not None:
            work.wait()
        return self.graph_loss

    def step(self):
        if self.graph is not None:
            self.graph.replay()
            return self.graph_loss
        if getattr(self, "seg", None) is not None:
            return self.step_segmented()
        return self.run()


class CriticStep(object):
    """The a5-inclusive variant: fi.Dev.forward (make-up conv, level rule, split, every crop, critic convs, class means) +
    fi.IntertwinerLoss + backward, called directly, with stock cuDNN convolutions inside the timed region."""

    def __init__(self, wl, device, world, seed):
        import feature_intertwiner_b200 as fi
        self.fi, self.wl, self.dev, self.world = fi, wl, device, world
        self.cfg = build_config(wl)
        torch.manual_seed(2000)
        self.dev_mod = fi.Dev(self.cfg, DEPTH).to(device).to(memory_format=torch.channels_last)
        self.dev_mod.spatial_sort = True
        self.dev_mod.eval()                               # the reference always runs in eval()
        self.ot = fi.OptTrans(self.cfg, ch_x=FEAT, L=wl["sinkhorn_iters"]).to(device)
        self.loss_mod = fi.IntertwinerLoss(self.cfg, ot_loss=self.ot, feat_dim=FEAT, distributed=world > 1, ot_padded=True).to(device)
        host = make_inputs(dict(wl, proposals=False), seed)
        self.rois, self.gt = host["rois"].to(device), host["gt"].to(device)
        self.maps = [m.to(device) for m in host["raw"]]
        g = torch.Generator().manual_seed(seed + 1)
        total = wl["batch"] * wl["rois_per_image"]
        self.g_pooled = torch.randn(total, DEPTH, 7, 7, generator=g).to(device).contiguous(memory_format=torch.channels_last)
        self.g_mask = torch.randn(total, DEPTH, 14, 14, generator=g).to(device).contiguous(memory_format=torch.channels_last)

    def step(self):
        x = [m.detach().requires_grad_() for m in self.maps]
        pooled, mask, feat_out = self.dev_mod(x, self.rois, self.gt)
        loss = self.loss_mod([feat_out[0], feat_out[1], feat_out[2], feat_out[3], None, None]).sum()
        torch.autograd.backward([loss, pooled, mask], [torch.ones_like(loss), self.g_pooled, self.g_mask])
        for p in list(self.ot.parameters()) + list(self.dev_mod.parameters()):
            p.grad = None
        return loss


# =================================================================================================== ours
class Timer(object):
    def __init__(self, dev, world, lib, flush):
        self.dev, self.world, self.lib, self.flush = dev, world, lib, flush
        self.host_s, self.per_step, self.launches, self.by_rank = 0.0, [], 0, []

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def __call__(self, fn, steps, warmup):
        # the cyclic collector of a process with torch loaded walks millions of objects: a generation-2 pass landing inside a
        # 2 ms step shows up as a 6-60 ms step.  Collect now, keep it off for the K timed steps (training loops do the same with
        # gc.freeze / a manual collection between iterations).  Done BEFORE the warm-up steps, like the creation of the timing
        # events: the collection takes tens of ms, and a GPU left idle that long right before the timed region starts it cold.
        gc.collect()
        gc.disable()
        evs = []
        pool = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps)]     # created and first-recorded outside the timed steps
        for ev in pool:
            ev.record()
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        # barrier + synchronize LAST, right before the first timed step: the collection above takes tens of ms and a different
        # time on every rank -- with the barrier in front of it the first step of the early ranks timed their wait for the late ones
        self.barrier()
        if self.world > 1:
            # ... and a DEVICE-side rendezvous behind it: the ranks leave the host barrier up to a few hundred microseconds apart,
            # which the first timed step would otherwise spend waiting for its peers inside the statistics exchange.  The tiny
            # all-reduce completes on every GPU at the same moment, with the first step already queued behind it.
            import torch.distributed as dist
            if getattr(self, "_rdv", None) is None:
                self._rdv = torch.zeros(1, device=self.dev)
            dist.all_reduce(self._rdv)
        self.host_s = 0.0
        launch0 = self.lib.fi_kernel_launches()
        for _ in range(steps):
            self.flush.add_(1.0)
            a, b = pool.pop(), pool.pop()
            t0 = time.perf_counter()
            a.record(); fn(); b.record()
            self.host_s += time.perf_counter() - t0          # host time to ENQUEUE a step
            evs.append((a, b))
        self.barrier()
        gc.enable()
        self.launches = self.lib.fi_kernel_launches() - launch0
        self.per_step = [a.elapsed_time(b) for a, b in evs]
        ms = sum(self.per_step)
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        self.by_rank = [ms / steps]
        if self.world > 1:
            import torch.distributed as dist
            every = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(every, t)
            self.by_rank = [round(float(v.item()) / steps, 4) for v in every]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps


def kernel_families(fi, step, timer, steps, peak):
    """Per-launch CUDA events of the RoIAlign kernels (eager pass: events inside a captured graph cannot be timed)."""
    fi.roi_align.enable_profiling(prealloc_events=12 * (steps + 3) + 64)
    ms = timer(step.run, steps, 3)
    records = fi.roi_align.disable_profiling()
    fam = {}
    n_per = len(records) // (steps + 3)
    for rec in (records[3 * n_per:] if records else []):
        d = fam.setdefault(rec["kernel"], [0.0, 0.0, 0])
        d[0] += fi.roi_align.algorithmic_bytes(rec); d[1] += rec["start"].elapsed_time(rec["end"]); d[2] += 1
    out = {}
    for k, v in fam.items():
        out[k] = {"launches_per_step": v[2] // steps, "avg_ms": v[1] / v[2], "alg_bytes_per_launch": v[0] / v[2]}
        if v[0] > 0:
            out[k].update(gbs=v[0] / v[1] / 1e6, frac=v[0] / v[1] / 1e6 / peak)
    return out, ms


def measure_workload(name, wl, dev, rank, world, args, lib, flush, full, ot_grad_allreduce=None):
    """One workload: device-resident value (graph or eager), e2e, loss-only, kernel families.  `full`: all legs; else value only."""
    import feature_intertwiner_b200 as fi
    timer = Timer(dev, world, lib, flush)
    step = Step(wl, dev, world, seed=2000 + rank, ot_grad_allreduce=args.ddp_ot_grads if ot_grad_allreduce is None else ot_grad_allreduce)
    # several ranks: one graph when the exchange is the peer-memory kernel; NCCL (no peer mapping, or the DDP-style gradient
    # all-reduce variant, where NCCL's ring beats a one-shot read of 8 x 15.7 MB) stays outside captures: three graphs
    one_graph = world == 1 or ("stats" in PEER and not step.ot_grad_allreduce)
    graphed = args.mode == "graph" and (step.capture() if one_graph else step.capture_segmented())
    ms = timer(step.step, args.steps if full else max(3, args.steps // 2), args.warmup)
    res = {"ms_per_step": ms, "value": wl["batch"] * wl["rois_per_image"] * world / (ms / 1e3), "graphed": graphed,
           "host_enqueue_ms_per_step": 1e3 * timer.host_s / (args.steps if full else max(3, args.steps // 2)),
           "ms_each_step": [round(v, 3) for v in timer.per_step], "counts": step.counts, "graph_error": step.graph_error,
           "ms_per_step_by_rank": timer.by_rank}
    res["launches_in_region"] = int(timer.launches)
    res["launches_per_step"] = step.launches_per_step if graphed else timer.launches / max(1, len(timer.per_step))
    if not full:
        del step
        torch.cuda.empty_cache()
        return res, None
    res["step"] = step
    return res, timer


def run_ours(args):
    rank, world, local = dist_env()
    import feature_intertwiner_b200 as fi
    from feature_intertwiner_b200 import _lib, synth
    numa = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    peer_on = setup_peer_allreduce(dev, world)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)    # 256 MB > 126 MB L2 (inputs alone are > 3 GB anyway)
    wl = dict(synth.WORKLOADS[args.workload])
    peak, peak_kind = measured_peak()

    if args.with_critic:
        return run_with_critic(args, wl, dev, rank, world, lib, flush)

    # ---- device-resident value ------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    res, timer = measure_workload(args.workload, wl, dev, rank, world, args, lib, flush, full=True)
    clocks = sampler.stop() if rank == 0 else None
    step = res.pop("step")
    ms = res["ms_per_step"]
    h2d_bytes = step.h2d_bytes
    # ---- e2e: pinned host -> device -> step -> loss back on the host ------------------------------
    def e2e_step():
        step.upload()
        return float(step.step().item())
    ms_e2e = timer(e2e_step, max(2, args.steps // 2), 2)
    # ---- kernel families (eager, per-launch events) and the loss head alone ------------------------
    kernels, ms_eager = kernel_families(fi, step, timer, max(3, args.steps // 2), peak)
    host_eager = 1e3 * timer.host_s / max(3, args.steps // 2)
    if step.grad_bucket is not None:
        step.grad_bucket.remove()       # the loss-only leg: statistics all-reduce included, OptTrans gradient all-reduce not
    # intertwiner loss alone, fwd + bwd: one CUDA graph per iteration (device time); several processes: the module's own graphed
    # head behind the eager statistics all-reduce
    loss_replay = step.capture_loss_only()
    loss_graph_error = getattr(step, "loss_graph_error", None)
    if loss_replay is not None:
        loss_graphed = "one CUDA graph per iteration (forward + backward)"
        ms_loss = timer(loss_replay, args.steps, args.warmup)
    else:
        ok = step.loss_mod.enable_cuda_graph([step.last_feat_in[0], step.last_feat_in[1], step.last_feat_in[2].detach().requires_grad_(),
                                              step.last_feat_in[3]])
        loss_graphed = "loss head graphed behind the eager statistics exchange (host-bound)" if ok else False
        loss_graph_error = loss_graph_error or getattr(step.loss_mod, "_graph_error", None)
        ms_loss = timer(step.loss_only, args.steps, args.warmup)
    # all-reduce of the class statistics: bus bandwidth at this size (it is latency, not bandwidth, that matters at 1.3 MB)
    nvlink = None
    if world > 1:
        import torch.distributed as dist
        buf = torch.zeros(2 * (FEAT * NCLS + NCLS), device=dev)
        big = torch.zeros(64 * 1024 * 1024, device=dev)
        def ar_small():
            dist.all_reduce(buf)
        def ar_big():
            dist.all_reduce(big)
        t_small = timer(ar_small, 20, 5)
        t_big = timer(ar_big, 10, 3)
        f = 2.0 * (world - 1) / world
        nvlink = {"class_stats_allreduce_nccl": {"bytes": buf.numel() * 4, "ms": t_small, "bus_gbs": f * buf.numel() * 4 / t_small / 1e6},
                  "allreduce_256MB_nccl": {"bytes": big.numel() * 4, "ms": t_big, "bus_gbs": f * big.numel() * 4 / t_big / 1e6},
                  "nvlink_peak_gbs_per_direction": 900.0, "measured_reference_bus_gbs_8rank_1GiB": 725.0}
        del big
        if peer_on:
            # the same two exchanges through csrc/peer_allreduce.cu (one-shot: every rank READS (world-1) x bytes over NVLink)
            gbuf = torch.zeros(PEER["grads"].capacity, device=dev)
            t_ps = timer(lambda: PEER["stats"](buf), 20, 5)
            t_pg = timer(lambda: PEER["grads"](gbuf), 20, 5)
            nvlink["class_stats_allreduce_peer_kernel"] = {"bytes": buf.numel() * 4, "ms": t_ps, "nvlink_read_gbs_per_rank": (world - 1) * buf.numel() * 4 / t_ps / 1e6}
            nvlink["ot_gradients_allreduce_peer_kernel"] = {"bytes": gbuf.numel() * 4, "ms": t_pg, "nvlink_read_gbs_per_rank": (world - 1) * gbuf.numel() * 4 / t_pg / 1e6}
            nvlink["peer_kernel_timeouts"] = bool(PEER["stats"].error() or PEER["grads"].error())
            del gbuf
    # loss value of the last step against the CPU port on the same class statistics (north_star: loss within 1e-4)
    stats = [t.detach().cpu() for t in step.last_feat_in]
    ot_state = {k: v.detach().cpu() for k, v in step.ot.state_dict().items()}
    with torch.no_grad():
        step.loss_mod.initialize_buffer()
        gpu_loss = step.loss_mod([t.to(dev) for t in stats] + [None, None]).detach().cpu()

    others = {}
    if not args.no_other_workloads:
        del step
        torch.cuda.empty_cache()
        for name in ("c1", "c3", "c5"):
            if name == args.workload:
                continue
            r, _ = measure_workload(name, dict(synth.WORKLOADS[name]), dev, rank, world, args, lib, flush, full=False)
            others[name] = {"workload": workload_name(name, synth.WORKLOADS[name]), "value": r["value"], "unit": "RoIs/s", "ms_per_step": r["ms_per_step"],
                            "graphed": r["graphed"], "small_counts": r["counts"][0], "big_counts": r["counts"][1], "steps": len(r["ms_each_step"])}

    variants = {}
    if world > 1 and not args.ddp_ot_grads and not args.no_other_workloads:
        r, _ = measure_workload(args.workload, wl, dev, rank, world, args, lib, flush, full=False, ot_grad_allreduce=True)
        variants["with_ddp_style_ot_gradient_allreduce"] = {
            "ms_per_step": r["ms_per_step"], "value": r["value"], "unit": "RoIs/s", "graphed": r["graphed"],
            "what": "the same step as three graphs + an async NCCL all-reduce of the OptTrans gradients (15.7 MB) under the RoIAlign backward graph "
                    "(what wrapping OptTrans in DistributedDataParallel adds; the values are identical on every rank, the reference never exchanges them)"}
    peer_timeouts = bool(peer_on and (PEER["stats"].error() or PEER["grads"].error()))
    rois_per_step = wl["batch"] * wl["rois_per_image"] * world
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the dominant kernel family -----------------------------------------------------
    for k in kernels:
        kernels[k]["share_of_step"] = kernels[k]["avg_ms"] * kernels[k]["launches_per_step"] / ms
    if not kernels:
        kernels = {"none": {"share_of_step": 0.0, "gbs": 0.0, "frac": 0.0, "avg_ms": 0.0}}
    cand = [k for k in kernels if "frac" in kernels[k]]
    dom = max(cand, key=lambda k: kernels[k]["share_of_step"]) if cand else "none"
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath) and args.workload == "c2":
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get(dom), tj.get("source")
        for k in kernels:
            kernels[k]["dram_traffic_per_launch"] = tj.get(k)
    roof = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom].get("gbs"), "peak": peak, "unit": "GB/s", "frac": kernels[dom].get("frac"),
            "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind}
    if "crop_bwd_nhwc" in kernels and "crop_bwd_lists" in kernels and "frac" in kernels["crop_bwd_nhwc"]:
        b, l = kernels["crop_bwd_nhwc"], kernels["crop_bwd_lists"]
        roof["note"] = ("crop_bwd_nhwc = tile_collapse + accumulate, the part of the backward that needs the gradients; its per-tile sample lists "
                        "(crop_bwd_lists = tile_prep + bin_enumerate, boxes only) are built at forward time on a side stream and overlap with the "
                        "segment means / loss head; frac_with_lists charges them to the backward as if nothing overlapped")
        roof["frac_with_lists"] = b["alg_bytes_per_launch"] / (b["avg_ms"] + l["avg_ms"]) / 1e6 / peak
        if "crop_fwd_nhwc" in kernels:
            f = kernels["crop_fwd_nhwc"]
            tot_b = b["alg_bytes_per_launch"] + f["alg_bytes_per_launch"]
            roof["fwd_plus_bwd_frac"] = tot_b / (b["avg_ms"] + f["avg_ms"]) / 1e6 / peak
            roof["fwd_plus_bwd_frac_with_lists"] = tot_b / (b["avg_ms"] + f["avg_ms"] + l["avg_ms"]) / 1e6 / peak
    out = {
        "metric": METRIC, "value": rois_per_step / (ms / 1e3), "unit": "RoIs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, wl),
                   "layout": "channels_last maps/crops (logical NCHW)", "ot": "all 80 foreground classes, absent ones masked (fixed shapes, no host sync)",
                   "roi_order": "spatially sorted per image (L2 reuse)" if os.environ.get("FI_SPATIAL_SORT", "1") != "0" else "index order",
                   "step": (("ONE CUDA graph replay per step (whole step: split, crops, list building on a side stream, class means, loss head, backward)"
                             if world == 1 else "ONE CUDA graph replay per step on every rank; the exchange of the class statistics is this library's peer-memory all-reduce "
                             "kernel over NVLink (csrc/peer_allreduce.cu), inside the graph"
                             if peer_on else "three CUDA graph replays per step (forward part | loss head fwd+bwd | backward part) with the two NCCL "
                             "all-reduces (class statistics; OptTrans gradients, overlapped with the backward graph) eager between them")
                            if res["graphed"] else "eager, launch by launch" + (" (graph capture failed: %s)" % res["graph_error"] if res["graph_error"] else "")),
                   "list_lengths": "kept on the device (fixed-capacity lists): the step has no device->host read",
                   "critic_and_makeup_convs": "excluded (stock cuDNN; SURVEY.md 8 a5) -- see --with-critic for the inclusive variant",
                   "l2": "GB-class working set per step (> 126 MB L2) + 256 MB flush write between steps", "gc": "python cyclic GC collected before and disabled during the timed steps",
                   "small_counts": res["counts"][0], "big_counts": res["counts"][1], "parallelism": "dp%d by image batch" % world},
        "e2e": {"value": rois_per_step / (ms_e2e / 1e3), "unit": "RoIs/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4, "h2d_gbs_per_gpu": h2d_bytes / (ms_e2e - ms) / 1e6,
                "bound": "the host->device copies (PCIe Gen5 x16: ~55 GB/s per GPU in practice); the step itself is %.1f %% of the e2e time" % (100.0 * ms / ms_e2e),
                "host_buffers": "pinned, first-touched on the GPU's NUMA node: %s" % (numa,)},
        "intertwiner_loss": {"ms_per_iter": ms_loss, "graphed": loss_graphed, "graph_error": loss_graph_error,
                             "what": "statistics merge -> buffer update -> class match -> OptTrans / Sinkhorn(L=%d), "
                             "%d classes, forward + backward, device-timed alone (it is also inside every step above)" % (wl["sinkhorn_iters"], NCLS - 1),
                             "gpu_vs_cpu_port_abs_diff": None},
        "gpu_launches": int(round(res["launches_per_step"] * args.steps)), "gpu_launches_per_step": res["launches_per_step"],
        "gpu_launches_note": "libfi_b200 kernels per step (counted while the step was captured) x steps; they run from the replayed graph" if res["graphed"]
                             else "libfi_b200 kernels launched directly in the timed region",
        "host_enqueue_ms_per_step": res["host_enqueue_ms_per_step"], "ms_each_step": res["ms_each_step"],
        "eager": {"ms_per_step": ms_eager, "host_enqueue_ms_per_step": host_eager, "what": "the same step launch by launch with per-launch events (kernel families below)"},
        "roofline": roof, "kernels": kernels, "clocks": clocks, "other_workloads": others,
    }
    if nvlink:
        out["nvlink"] = nvlink
    if world > 1:
        out["ms_per_step_by_rank"] = res["ms_per_step_by_rank"]
        out["exchange"] = {"class_statistics": "peer-memory all-reduce kernel (csrc/peer_allreduce.cu), inside the step's graph" if peer_on
                           else "NCCL all-reduce between the graphs (%s)" % PEER.get("why", "FI_PEER_ALLREDUCE=0"),
                           "ot_gradients": "all-reduced every step (--ddp-ot-grads)" if args.ddp_ot_grads else
                           "not exchanged: every rank computes the same loss head on the same totals (the reference computes it on GPU 0 only, lib/model.py:143-144)",
                           "peer_kernel_timeouts": peer_timeouts}
        out["variants"] = variants
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference(wl, seed=2000, budget_s=args.cpu_budget, stats=stats, ot_state=ot_state)
        cpu_loss = out["cpu_baseline"].pop("loss_vector", None)
        if cpu_loss is not None:
            out["intertwiner_loss"]["gpu_vs_cpu_port_abs_diff"] = float((gpu_loss.view(-1) - torch.tensor(cpu_loss).view(-1)).abs().max())
            out["intertwiner_loss"]["loss_sum"] = float(gpu_loss.sum())
    print(json.dumps(out), flush=True)


def run_with_critic(args, wl, dev, rank, world, lib, flush):
    timer = Timer(dev, world, lib, flush)
    step = CriticStep(wl, dev, world, seed=2000 + rank)
    ms = timer(step.step, args.steps, args.warmup)
    host_ms = 1e3 * timer.host_s / args.steps
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    rois_per_step = wl["batch"] * wl["rois_per_image"] * world
    print(json.dumps({
        "metric": METRIC + " + make-up conv + critic convs (a5 inclusive)", "value": rois_per_step / (ms / 1e3), "unit": "RoIs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(args.workload, wl), "variant": "fi.Dev.forward + fi.IntertwinerLoss + backward called "
                                        "directly: make-up conv3x3+BN+ReLU on every level map and the 27.9 M-parameter critic (stock cuDNN, fp32, TF32 off) are inside "
                                        "the timed region; one host read per step (the list lengths size the critic's batches)"},
        "host_enqueue_ms_per_step": host_ms, "ms_each_step": [round(v, 3) for v in timer.per_step], "gpu_launches": int(timer.launches),
        "e2e": None, "roofline": None}), flush=True)


# =================================================================================================== reference (CPU)
def cpu_reference(wl, seed, budget_s=20.0, full=False, stats=None, ot_state=None):
    """The reference's CPU path: crop_and_resize.c (compiled unmodified, oracle/_ref) for every RoIAlign call of the step, forward
    (OpenMP, all cores) + backward (serial, as written), plus the whole class-level OT loss (torch-CPU port of lib/OT_module.py,
    all cores).  full=True: every crop of the step.  Otherwise a bounded sample -- 1/k of the boxes of every call, per-box work
    scaled by k, the per-call map allocation / zero fill counted once -- sized to cost about budget_s seconds."""
    import numpy as np
    from oracle import clib, pyref
    from feature_intertwiner_b200 import synth
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    torch.set_num_threads(cores)
    kind = "reference" if clib.have_ref() else "port"
    fwd = clib.ref_crop_and_resize_fwd if clib.have_ref() else clib.oracle_crop_and_resize_fwd
    bwd = clib.ref_crop_and_resize_bwd if clib.have_ref() else clib.oracle_crop_and_resize_bwd
    g = torch.Generator().manual_seed(seed)
    B, R, hw = wl["batch"], wl["rois_per_image"], wl["image"]
    rois = synth.make_rois(B, R, hw, g)
    level, _ = pyref.roi_level_ref(rois, (hw[0], hw[1], 3))
    level = level.view(-1).numpy()
    flat = rois.view(-1, 4).numpy()
    shapes = synth.level_shapes(hw)
    cache = cpu_reference.__dict__.setdefault("maps", {})
    key = (B, hw)
    if key not in cache:
        rng = np.random.default_rng(seed)
        cache.clear()
        cache[key] = [rng.standard_normal((B, DEPTH, h, w), dtype=np.float32) for (h, w) in shapes]
    maps = cache[key]
    calls = []
    for i in range(4):
        sm = np.nonzero(level == i + 2)[0]
        bg = np.nonzero(level > i + 2)[0]
        if len(sm) == 0:
            continue
        calls += [(i, 7, sm), (i, 14, sm)]
        if i < 3 and len(bg):
            calls += [(i, 14, bg)]
    total_crops = sum(len(c[2]) for c in calls)
    k = 1
    if not full:
        # one probe call sizes the sample so the whole baseline costs about budget_s
        t0 = time.perf_counter()
        probe = calls[0][2][:16]
        o = fwd(maps[0], flat[probe], (probe // R).astype(np.int32), 7, 7)
        bwd(o, flat[probe], (probe // R).astype(np.int32), maps[0].shape)
        per_crop = (time.perf_counter() - t0) / 16 * 2.5
        k = max(1, int(np.ceil(total_crops * per_crop / (0.6 * budget_s))))
    t_var, t_fixed, n_sample = 0.0, 0.0, 0
    empty_b, empty_i = np.zeros((0, 4), np.float32), np.zeros((0,), np.int32)
    for (i, P, idx) in calls:
        sub = idx[::k]
        ind = (sub // R).astype(np.int32)
        fixed = 0.0
        if k > 1:
            # per-call cost that does not depend on the number of boxes (allocation + memset of the dense gradient map, as the
            # reference does it): measured with an empty box list and counted ONCE per call, not scaled by the sampling factor
            t0 = time.perf_counter()
            bwd(np.zeros((0, DEPTH, P, P), np.float32), empty_b, empty_i, maps[i].shape)
            fixed = time.perf_counter() - t0
        t0 = time.perf_counter()
        o = fwd(maps[i], flat[sub], ind, P, P)
        bwd(o, flat[sub], ind, maps[i].shape)
        whole = time.perf_counter() - t0
        t_fixed += fixed
        t_var += max(0.0, whole - fixed)
        n_sample += len(sub)
    t_roi_full = t_fixed + t_var * k
    # ---- the loss: torch-CPU PORT of lib/OT_module.py + lib/model.py:meta_loss (oracle/pyref.py), all 80 foreground classes
    torch.manual_seed(seed)
    ot = pyref.OptTransRef(ch_x=FEAT, L=wl["sinkhorn_iters"])
    n_cls = NCLS - 1
    loss_vec = None
    if stats is not None:
        # the very class statistics of the GPU step, same OptTrans weights: merge -> buffer -> padded class match -> OT loss
        ot.load_state_dict(ot_state)
        bf, bc, sf, sc = stats
        s_mean, s_n = pyref.merge_feat_vec_ref(sf, sc)
        b_mean, b_n = pyref.merge_feat_vec_ref(bf, bc)
        x = s_mean.t()[1:].unsqueeze(-1).contiguous().requires_grad_()
        y = b_mean.t()[1:].unsqueeze(-1).contiguous()
        mask = ((s_n.view(-1) > 0) & (b_n.view(-1) > 0))[1:].float()
        t0 = time.perf_counter()
        w = ot(x, y)
        (w * mask).sum().backward()
        t_loss = time.perf_counter() - t0
        loss_vec = (w * mask).detach().tolist()
    else:
        x = torch.rand(n_cls, FEAT, 1).requires_grad_()
        y = torch.rand(n_cls, FEAT, 1)
        t0 = time.perf_counter()
        ot(x, y).sum().backward()
        t_loss = time.perf_counter() - t0
    t_full = t_roi_full + t_loss
    if full:
        sample = ("full step: every crop of every RoIAlign call (%d crops; fwd OpenMP x%d + bwd serial, %.2f s) + the full %d-class OT loss "
                  "fwd+bwd, L=%d (torch-CPU port, %.2f s)" % (total_crops, cores, t_roi_full, n_cls, wl["sinkhorn_iters"], t_loss))
    else:
        sample = ("every RoIAlign call of the step on 1/%d of its boxes (%d of %d crops; fwd OpenMP x%d + bwd serial: %.2f s per-box "
                  "work, scaled by %d, + %.2f s per-call map allocation / zero fill, counted once) + the full %d-class OT loss fwd+bwd, "
                  "L=%d (torch-CPU port, %.2f s); full step %.1f s; `--impl reference` runs full steps"
                  % (k, n_sample, total_crops, cores, t_var, k, t_fixed, n_cls, wl["sinkhorn_iters"], t_loss, t_full))
    return {"value": B * R / t_full, "unit": "RoIs/s", "cores": cores, "kind": kind,
            "kind_note": "RoIAlign: lib/roi_align/src/crop_and_resize.c compiled unmodified (reference); loss: torch-CPU port of lib/OT_module.py (oracle/pyref.py)",
            "sample": sample, "roialign_s_full_step": t_roi_full, "loss_s": t_loss, "loss_vector": loss_vec}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from feature_intertwiner_b200 import synth
    wl = synth.WORKLOADS[args.workload]
    # FULL steps (every crop of the step, ~3 s each at c2); only if W + K of them would take more than ~4 minutes are the steps
    # bounded samples instead
    probe = cpu_reference(wl, seed=2000, full=True)
    t_step = wl["batch"] * wl["rois_per_image"] / probe["value"]
    n = args.warmup + args.steps
    full = t_step * n <= 240.0
    vals = []
    for s in range(n):
        r = cpu_reference(wl, seed=2000, full=full, budget_s=max(3.0, 200.0 / n))
        if s >= args.warmup:
            vals.append(r)
    v = sum(r["value"] for r in vals) / len(vals)
    last = dict(vals[-1], value=v)
    last.pop("loss_vector", None)
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "RoIs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wl["batch"] * wl["rois_per_image"] / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(args.workload, wl)},
        "cpu_baseline": last, "e2e": {"value": v, "unit": "RoIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"])
    ap.add_argument("--with-critic", action="store_true", help="a5-inclusive variant: fi.Dev.forward + fi.IntertwinerLoss called directly")
    ap.add_argument("--ddp-ot-grads", action="store_true", help="several ranks: also all-reduce the (identical) OptTrans gradients every step, as "
                    "DistributedDataParallel would; default: measured as a variant next to the headline")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the short c1 / c3 / c5 measurements")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
        run_ours(args)


if __name__ == "__main__":
    main()
