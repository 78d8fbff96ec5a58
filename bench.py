#!/usr/bin/env python
"""Benchmark of the Feature Intertwiner hot path on B200 (BASELINE.json metric: RoIs/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input:

    level rule + reliable/less-reliable split  ->  every RoIAlign call of Dev.forward (big 14x14 on the raw
    level maps, small 7x7 + 14x14 on the made-up maps; 7x7 written straight to its final row)  ->  per-class
    segment means of the critic features  ->  buffer update + class match  ->  OptTrans / Sinkhorn(L) loss
    ->  backward of all of it (Sinkhorn gradient, segment-mean backward, RoIAlign backward of every crop).

The make-up conv and the critic convs are stock cuDNN and are NOT part of the path:
their outputs (made-up maps, critic features) and the upstream crop gradients are synthetic inputs.

`value`  : RoIs/s with all inputs resident in HBM (device-timed, CUDA events, max over ranks).
`e2e`    : the same step through the public API starting from pinned HOST buffers (H2D of every input,
           D2H of the loss) inside the timed region.
`--impl reference`: the reference's own CPU implementation (oracle/_ref: lib/roi_align/src/crop_and_resize.c
           compiled unmodified, OpenMP forward + serial backward) plus the torch-CPU restatement of
           lib/OT_module.py for the loss, on a bounded sample of the same workload.
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
def make_inputs(wl, seed):
    """Everything a step consumes, on the HOST (pinned), seeded.  The split itself is computed by the step."""
    from feature_intertwiner_b200 import synth
    g = torch.Generator().manual_seed(seed)
    B, R, hw = wl["batch"], wl["rois_per_image"], wl["image"]
    host = {
        "rois": synth.make_rois(B, R, hw, g),
        "gt": synth.make_class_ids(B, R, g, NCLS),
        "raw": synth.make_feature_maps(B, hw, DEPTH, g, channels_last=True),      # P2..P5
        "madeup": synth.make_feature_maps(B, hw, DEPTH, g, channels_last=True),   # upsample(P2..P5): stock conv output
    }
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
        DATA=ns(IMAGE_SHAPE=np.array([wl["image"][0], wl["image"][1], 3])), DATASET=ns(NUM_CLASSES=NCLS))


class Step(object):
    """Device-side state + one pass of the hot path through the public API."""

    def __init__(self, wl, device, world, seed):
        import feature_intertwiner_b200 as fi
        self.fi, self.wl, self.dev, self.world = fi, wl, device, world
        self.spatial_sort = os.environ.get("FI_SPATIAL_SORT", "1") != "0"
        self.use_graph = os.environ.get("FI_GRAPH", "1") != "0"     # CUDA-graph the fixed-shape loss head (no collective inside)
        self.graph_tried, self.graphed = False, False
        self.cfg = build_config(wl)
        torch.manual_seed(2000)
        self.ot = fi.OptTrans(self.cfg, ch_x=FEAT, L=wl["sinkhorn_iters"]).to(device)
        self.loss_mod = fi.IntertwinerLoss(self.cfg, ot_loss=self.ot, feat_dim=FEAT, distributed=world > 1, ot_padded=True).to(device)
        self.host = make_inputs(wl, seed)
        B, R = wl["batch"], wl["rois_per_image"]
        # the split of THIS input fixes the shapes of the synthetic critic features / upstream gradients
        rois_d = self.host["rois"].to(device)
        split = fi.split_levels(fi.roi_level(rois_d, self.cfg.DATA.IMAGE_SHAPE))
        self.counts = (list(split.small_cnt), list(split.big_cnt))
        g = torch.Generator().manual_seed(seed + 1)
        self.host["small_feat"] = [torch.rand(split.small_cnt[i], FEAT, generator=g) for i in range(3)]
        self.host["big_feat"] = [torch.rand(split.big_cnt[i], FEAT, generator=g) for i in range(3)]
        self.host["g_pooled"] = torch.randn(B * R, DEPTH, 7, 7, generator=g).contiguous(memory_format=torch.channels_last)
        self.host["g_mask"] = torch.randn(B * R, DEPTH, 14, 14, generator=g).contiguous(memory_format=torch.channels_last)
        self.host["g_small"] = [torch.randn(split.small_cnt[i], DEPTH, 14, 14, generator=g).contiguous(memory_format=torch.channels_last)
                                for i in range(3)]           # gradient the critic sends back into the compact 14x14 crops
        self.host["g_big"] = [torch.randn(split.big_cnt[i], DEPTH, 14, 14, generator=g).contiguous(memory_format=torch.channels_last)
                              for i in range(3)]
        self.h2d_bytes = 0
        self.pinned = self._map(self.host, pin)
        self.resident = self._map(self.host, lambda t: t.to(device))
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self._flat(self.host))
        self.algo = None

    @staticmethod
    def _map(d, f):
        return {k: ([f(t) for t in v] if isinstance(v, list) else f(v)) for k, v in d.items()}

    @staticmethod
    def _flat(d):
        for v in d.values():
            for t in (v if isinstance(v, list) else [v]):
                yield t

    def loss_only(self):
        """The intertwiner loss alone (BASELINE.json's second metric): statistics merge (+ all-reduce) -> buffer update ->
        class match -> OptTrans / Sinkhorn, forward and backward, on the class statistics of the last step."""
        bf, bc, sf, sc = self.last_feat_in
        loss = self.loss_mod([bf, bc, sf.requires_grad_(), sc, None, None]).sum()
        loss.backward()
        for p in self.ot.parameters():
            p.grad = None
        return loss

    def upload(self):
        """H2D of every input of the step from pinned host memory (the e2e leg)."""
        return self._map(self.pinned, lambda t: t.to(self.dev, non_blocking=True))

    def run(self, inp):
        fi, cfg = self.fi, self.cfg
        B, R = self.wl["batch"], self.wl["rois_per_image"]
        total = B * R
        rois, gt = inp["rois"], inp["gt"]
        # fresh leaves every step (a training loop clears .grad each iteration; re-using the leaf would time an extra
        # read-modify-write of every map in AccumulateGrad)
        raw = [m.detach().requires_grad_() for m in inp["raw"]]
        madeup = [m.detach().requires_grad_() for m in inp["madeup"]]
        small_f = [t.detach().requires_grad_() for t in inp["small_feat"]]
        big_f = inp["big_feat"]
        split = fi.split_levels(fi.roi_level(rois, cfg.DATA.IMAGE_SHAPE, cfg.ROIS.ASSIGN_ANCHOR_BASE), rois=rois, gt=gt,
                                order=fi.spatial_order(rois) if self.spatial_sort else None)
        pooled_out = torch.empty((total, DEPTH, 7, 7), device=self.dev, memory_format=torch.channels_last)
        mask_out = torch.empty((total, DEPTH, 14, 14), device=self.dev, memory_format=torch.channels_last)
        # every crop of the pass in one level-batched launch (fi.crop_sets), like Dev.forward
        specs, where = [], {}
        for i in range(4):
            if split.small_cnt[i] == 0:
                continue
            if i < 3 and split.big_cnt[i]:
                where[("big", i)] = len(specs)
                specs.append(dict(image=raw[i], boxes=split.big_boxes(i), box_ind=split.big_ind(i), size=14, img_offsets=split.big_img_offsets(i)))
            s32 = split.small(i)
            boxes, ind = split.small_boxes(i), split.small_ind(i)
            where[("small", i)] = len(specs)
            specs.append(dict(image=madeup[i], boxes=boxes, box_ind=ind, size=7, out=pooled_out, dst_row=s32, img_offsets=split.small_img_offsets(i)))
            specs.append(dict(image=madeup[i], boxes=boxes, box_ind=ind, size=14, out=mask_out, dst_row=s32, compact=(i < 3),
                              img_offsets=split.small_img_offsets(i)))
        res_out, res_comp = fi.crop_sets(specs)
        outs, grads = [], []
        bfeat, bcnt, sfeat, scnt = [], [], [], []
        for i in range(4):
            if ("small", i) not in where:
                continue
            if ("big", i) in where:
                k = where[("big", i)]
                outs.append(res_comp[k]); grads.append(inp["g_big"][i])      # compact 14x14 crop -> critic (stock conv, not timed)
                f, c = fi.assign_feat2cls(split.big_gt(i), big_f[i], NCLS)
                bfeat.append(f); bcnt.append(c)
            k = where[("small", i)]
            pooled_out, mask_out = res_out[k], res_out[k + 1]
            if i < 3:
                outs.append(res_comp[k + 1]); grads.append(inp["g_small"][i])
                f, c = fi.assign_feat2cls(split.small_gt(i), small_f[i], NCLS)
                sfeat.append(f); scnt.append(c)
        feat_in = [torch.stack(bfeat)[None].detach(), torch.stack(bcnt)[None], torch.stack(sfeat)[None], torch.stack(scnt)[None], None, None]
        self.last_feat_in = [t.detach() for t in feat_in[:4]]          # for the loss-head-only timing
        if self.use_graph and not self.graph_tried:
            self.graph_tried = True
            self.graphed = self.loss_mod.enable_cuda_graph([feat_in[0], feat_in[1], feat_in[2].detach().requires_grad_(), feat_in[3]])
        loss = self.loss_mod(feat_in).sum()
        torch.autograd.backward([loss, pooled_out, mask_out] + outs, [torch.ones_like(loss), inp["g_pooled"], inp["g_mask"]] + grads)
        if self.world > 1:
            import torch.distributed as dist
            flat = torch.cat([p.grad.reshape(-1) for p in self.ot.parameters()])
            dist.all_reduce(flat)                     # gradient all-reduce of the path's own parameters (OptTrans)
        for p in self.ot.parameters():
            p.grad = None
        return loss


# =================================================================================================== ours
def run_ours(args):
    rank, world, local = dist_env()
    import feature_intertwiner_b200 as fi
    from feature_intertwiner_b200 import _lib, synth
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    wl = synth.WORKLOADS[args.workload]
    step = Step(wl, dev, world, seed=2000 + rank)
    lib = _lib.lib()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)    # 256 MB > 126 MB L2 (inputs alone are > 3 GB anyway)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    host = {"s": 0.0}

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        # the cyclic collector of a process with torch loaded walks millions of objects: a generation-2 pass landing inside a
        # 4 ms step shows up as a 6-60 ms step.  Collect now, keep it off for the K timed steps (training loops do the same with
        # gc.freeze / a manual collection between iterations).
        gc.collect()
        gc.disable()
        evs = []
        pool = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps)]     # created and first-recorded outside the timed steps
        for ev in pool:
            ev.record()
        torch.cuda.synchronize()
        host["s"] = 0.0
        host["launch0"] = lib.fi_kernel_launches()
        for _ in range(steps):
            flush.add_(1.0)
            a, b = pool.pop(), pool.pop()
            t0 = time.perf_counter()
            a.record(); fn(); b.record()
            host["s"] += time.perf_counter() - t0          # host time to ENQUEUE a step (no sync inside except the split read)
            evs.append((a, b))
        barrier()
        gc.enable()
        host["launches"] = lib.fi_kernel_launches() - host["launch0"]
        per_step = [a.elapsed_time(b) for a, b in evs]
        host["per_step"] = per_step
        ms = sum(per_step)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    # ---- device-resident value ------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if os.environ.get("FI_BENCH_NOPROF") != "1":
        fi.roi_align.enable_profiling(prealloc_events=8 * (args.steps + args.warmup) + 64)     # 2 launches x 2 events per step + slack
    ms = timed(lambda: step.run(step.resident), args.steps, args.warmup)
    host_ms = 1e3 * host["s"] / args.steps
    per_step_ms = [round(v, 3) for v in host["per_step"]]
    launches = host["launches"]            # kernels of libfi_b200 enqueued directly inside the timed region
    records = fi.roi_align.disable_profiling()
    clocks = sampler.stop() if rank == 0 else None
    # ---- e2e: pinned host -> device -> step -> loss back on the host ------------------------------
    def e2e_step():
        loss = step.run(step.upload())
        return float(loss.item())
    ms_e2e = timed(e2e_step, max(2, args.steps // 2), 2)
    ms_loss = timed(step.loss_only, args.steps, args.warmup)          # intertwiner loss alone, fwd + bwd

    rois_per_step = wl["batch"] * wl["rois_per_image"] * world
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_kind = measured_peak()
    # ---- roofline of the dominant kernel family, from the per-launch events of the timed region ----
    fam = {}
    per_step = len(records) // (args.steps + args.warmup)
    for rec in (records[args.warmup * per_step:] if records else []):
        d = fam.setdefault(rec["kernel"], [0.0, 0.0, 0])
        d[0] += fi.roi_align.algorithmic_bytes(rec); d[1] += rec["start"].elapsed_time(rec["end"]); d[2] += 1
    kernels = {k: {"launches_per_step": v[2] // args.steps, "avg_ms": v[1] / v[2], "alg_bytes_per_launch": v[0] / v[2], "gbs": v[0] / v[1] / 1e6,
                   "frac": v[0] / v[1] / 1e6 / peak, "share_of_step": v[1] / args.steps / ms} for k, v in fam.items()}
    if not kernels:
        kernels = {"none": {"share_of_step": 0.0, "gbs": 0.0, "frac": 0.0}}
    dom = max(kernels, key=lambda k: kernels[k]["share_of_step"])
    # DRAM bytes per launch of the same kernels from the committed `ncu --set full` capture of this command (profiles/)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and args.workload == "c2":
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get(dom), tj.get("source")
        for k in kernels:
            kernels[k]["dram_traffic_per_launch"] = tj.get(k)
    out = {
        "metric": "RoIs/sec (RoIAlign fwd+bwd + split + class means + Sinkhorn intertwiner loss, fwd+bwd)",
        "value": rois_per_step / (ms / 1e3), "unit": "RoIs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: batch %d/GPU, %dx%d, %d RoIs/img, FPN P2-P5 C=256, pools 7+14, Sinkhorn N=256 L=%d, class-level OT loss"
                               % (args.workload, wl["batch"], wl["image"][0], wl["image"][1], wl["rois_per_image"], wl["sinkhorn_iters"]),
                   "layout": "channels_last maps/crops (logical NCHW)", "ot": "all 80 foreground classes, absent ones masked (fixed shapes, no host sync)",
                   "roi_order": "spatially sorted per image (L2 reuse)" if step.spatial_sort else "index order",
                   "loss_head": "CUDA graph (fwd+bwd)" if step.graphed else "eager", "critic_and_makeup_convs": "excluded (stock cuDNN; SURVEY.md 8 a5)",
                   "l2": "512 MB-class working set per step (> 126 MB L2) + 256 MB flush write between steps", "gc": "python cyclic GC collected before and disabled during the timed steps",
                   "small_counts": step.counts[0], "big_counts": step.counts[1], "parallelism": "dp%d by image batch" % world},
        "e2e": {"value": rois_per_step / (ms_e2e / 1e3), "unit": "RoIs/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": step.h2d_bytes, "d2h_bytes_per_step": 4},
        "intertwiner_loss": {"ms_per_iter": ms_loss, "what": "statistics merge -> buffer update -> class match -> OptTrans / Sinkhorn(L=%d), "
                             "%d classes, forward + backward, device-timed alone (it is also inside every step above)" % (wl["sinkhorn_iters"], NCLS - 1)},
        "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
        "gpu_launches_note": "libfi_b200 kernels launched directly in the timed region; with the loss head captured in CUDA graphs its 3 "
                             "libfi_b200 kernels per step (buffer update x2, Sinkhorn) replay from the graph and are not in this count",
        "host_enqueue_ms_per_step": host_ms, "ms_each_step": per_step_ms,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                     "frac": kernels[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind},
        "kernels": kernels,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference(wl, seed=2000, budget_s=args.cpu_budget)
    print(json.dumps(out), flush=True)


# =================================================================================================== reference (CPU)
def cpu_reference(wl, seed, budget_s=20.0):
    """The reference's CPU path on a bounded sample: crop_and_resize.c (compiled unmodified, oracle/_ref) for every
    RoIAlign call of the step on 1/k of the boxes, forward (OpenMP, all cores) + backward (serial, as written), plus
    the whole class-level OT loss (torch-CPU restatement of lib/OT_module.py, all cores).  Throughput is scaled to
    the full step: t_full = t_roialign_sample * k + t_loss."""
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
    rng = np.random.default_rng(seed)
    maps = [rng.standard_normal((B, DEPTH, h, w), dtype=np.float32) for (h, w) in shapes]
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
    torch.manual_seed(seed)
    ot = pyref.OptTransRef(ch_x=FEAT, L=wl["sinkhorn_iters"])
    n_cls = NCLS - 1
    x = torch.rand(n_cls, FEAT, 1).requires_grad_()
    y = torch.rand(n_cls, FEAT, 1)
    t0 = time.perf_counter()
    ot(x, y).sum().backward()
    t_loss = time.perf_counter() - t0
    t_full = t_roi_full + t_loss
    return {"value": B * R / t_full, "unit": "RoIs/s", "cores": cores, "kind": kind,
            "sample": "every RoIAlign call of the step on 1/%d of its boxes (%d of %d crops; fwd OpenMP x%d + bwd serial: %.2f s per-box "
                      "work, scaled by %d, + %.2f s per-call map allocation / zero fill, counted once) + the full %d-class OT loss fwd+bwd, "
                      "L=%d (%.2f s); full step %.1f s"
                      % (k, n_sample, total_crops, cores, t_var, k, t_fixed, n_cls, wl["sinkhorn_iters"], t_loss, t_full),
            "roialign_s_full_step": t_roi_full, "loss_s": t_loss}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from feature_intertwiner_b200 import synth
    wl = synth.WORKLOADS[args.workload]
    # every step is a bounded sample of the workload; its size is chosen so that the W + K steps end within a few minutes
    budget = max(3.0, min(args.cpu_budget, 150.0 / max(1, args.warmup + args.steps)))
    vals = []
    for s in range(args.warmup + args.steps):
        r = cpu_reference(wl, seed=2000, budget_s=budget)
        if s >= args.warmup:
            vals.append(r)
    v = sum(r["value"] for r in vals) / len(vals)
    out = {
        "impl": "reference",
        "metric": "RoIs/sec (RoIAlign fwd+bwd + split + class means + Sinkhorn intertwiner loss, fwd+bwd)",
        "value": v, "unit": "RoIs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wl["batch"] * wl["rois_per_image"] / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: batch %d/GPU, %dx%d, %d RoIs/img, FPN P2-P5 C=256, pools 7+14, Sinkhorn N=256 L=%d, class-level OT loss"
                               % (args.workload, wl["batch"], wl["image"][0], wl["image"][1], wl["rois_per_image"], wl["sinkhorn_iters"])},
        "cpu_baseline": dict(vals[-1], value=v),
        "e2e": {"value": v, "unit": "RoIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work per reference sample")
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
