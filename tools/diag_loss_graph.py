"""Why does IntertwinerLoss.enable_cuda_graph fall back (or not) after the whole-step capture of bench.py?  Prints the capture
result, the stored error and the loss-only timing both ways."""
import argparse
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import feature_intertwiner_b200 as fi  # noqa: E402
from feature_intertwiner_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--no-step-graph", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
wl = dict(synth.WORKLOADS[a.workload])
step = bench.Step(wl, dev, 1, 2000)
step.run()
torch.cuda.synchronize()
if not a.no_step_graph:
    print("whole-step capture:", step.capture(), step.graph_error)
step.run()
torch.cuda.synchronize()


def time_loss(n=20):
    for _ in range(3):
        step.loss_only()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step.loss_only()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("loss-only eager ms/iter: %.4f" % time_loss())
mod = step.loss_mod
feat = [step.last_feat_in[0], step.last_feat_in[1], step.last_feat_in[2].detach().requires_grad_(), step.last_feat_in[3]]
ok = mod.enable_cuda_graph(feat)
print("enable_cuda_graph:", ok, getattr(mod, "_graph_error", None))
if not ok:
    # the same capture without the try / except, for the traceback
    try:
        from feature_intertwiner_b200.intertwiner import _LossHead
        with torch.no_grad():
            sums = mod._stage_sums([t.detach() for t in feat[:4]])
        head = _LossHead(mod)
        sample = (sums[0].clone(), sums[1].clone(), sums[2].clone().requires_grad_(), sums[3].clone())
        torch.cuda.make_graphed_callables(head, sample)
    except Exception:       # noqa: BLE001
        traceback.print_exc()
print("loss-only after enable ms/iter: %.4f" % time_loss())
