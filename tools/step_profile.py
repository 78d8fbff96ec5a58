"""Host-side view of one bench step (torch.profiler): which ops cost CPU time / where the GPU idles."""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from feature_intertwiner_b200 import synth  # noqa: E402

torch.cuda.set_device(0)
step = bench.Step(synth.WORKLOADS["c2"], torch.device("cuda", 0), 1, seed=2000)
for _ in range(3):
    step.run(step.resident)
torch.cuda.synchronize()
for k in range(3):
    t0 = time.perf_counter()
    step.run(step.resident)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("step %d: host enqueue %.2f ms, +sync %.2f ms" % (k, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step.run(step.resident)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
