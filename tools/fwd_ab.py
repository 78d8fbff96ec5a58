"""A/B harness for the NHWC RoIAlign forward at a BASELINE.json workload: every crop set of one Dev.forward pass through
fi.crop_sets (ONE launch), timed with CUDA events per formulation (fi_set_option fwd_form; L2 flushed between iterations),
each formulation's outputs compared bit for bit with the round-1 unit (form 1).

    python tools/fwd_ab.py --workload c2 --iters 20 --forms 1,0,3:1:1,3:2:2,4:3:2 --out gpurun_out/fwd_ab.json

A configuration is fwd_form[:fwd_chunk[:fwd_pair[:fwd_sched]]] (include/fi_b200.h FI_OPT_FWD_*).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import feature_intertwiner_b200 as fi  # noqa: E402
from feature_intertwiner_b200 import synth  # noqa: E402
from bwd_ab import build_specs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--forms", default="1,0,3,4,5,6")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    wl = synth.WORKLOADS[args.workload]
    raw, madeup, specs, split = build_specs(wl, dev)
    for m in raw + madeup:
        m.requires_grad_(False)
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    want = None
    rows = []
    for cfg in args.forms.split(","):
        form, chunk, pair, sched = ([int(v) for v in cfg.split(":")] + [0, 0, 0])[:4]
        fi.set_option("fwd_form", form)
        fi.set_option("fwd_chunk", chunk)
        fi.set_option("fwd_pair", pair)
        fi.set_option("fwd_sched", sched)
        times = []
        res = None
        for it in range(args.iters + 3):
            flush.add_(1.0)
            flush.add_(1.0)                     # the second pass lets the host run ahead of the device
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            outs, comps = fi.crop_sets(specs)
            b.record()
            torch.cuda.synchronize()
            if it >= 3:
                times.append(a.elapsed_time(b))
            res = [t for t in list(outs) + list(comps) if t is not None]
        times.sort()
        live = [int(c) for c in split.small_cnt]
        # bits of everything the launch wrote (rows of the shared outputs are all live at full lists)
        got = [t.clone() for t in res]
        same = None
        if want is None:
            want = got
        else:
            same = all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(got, want))
        rows.append(dict(form=form, chunk=chunk, pair=pair, sched=sched, fwd_ms_median=round(times[len(times) // 2], 4), fwd_ms_min=round(times[0], 4), identical_to_first=same))
        print(rows[-1], flush=True)
    for k in ("fwd_form", "fwd_chunk", "fwd_pair", "fwd_sched"):
        fi.set_option(k, 0)
    out = dict(workload=args.workload, rows=rows, small=split.small_cnt, big=split.big_cnt,
               note="time = one fi_crop_sets_forward launch + the Python call around it (a few us of host time, the launch is ~0.8 ms)")
    s = json.dumps(out)
    print(s)
    if args.out:
        open(args.out, "w").write(s)


if __name__ == "__main__":
    main()
