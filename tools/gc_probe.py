"""Why do some bench steps take 10-70 ms?  Per step: host time, reserved / allocated CUDA memory, cudaMalloc count; then what the
cyclic GC finds after N steps with the collector off (reference cycles keep multi-GB step tensors alive -> the caching
allocator has to cudaMalloc fresh segments -> host stalls)."""
import collections
import gc
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from feature_intertwiner_b200 import synth  # noqa: E402

torch.cuda.set_device(0)
step = bench.Step(synth.WORKLOADS["c2"], torch.device("cuda", 0), 1, seed=2000)
for _ in range(5):
    step.run(step.resident)
torch.cuda.synchronize()
gc.collect()
gc.disable()
for k in range(12):
    t0 = time.perf_counter()
    step.run(step.resident)
    torch.cuda.synchronize()
    st = torch.cuda.memory_stats()
    print("step %2d  %.2f ms  reserved %.2f GB  allocated %.2f GB  segments %d  cudaMalloc retries %d  gc counts %s" % (
        k, 1e3 * (time.perf_counter() - t0), st["reserved_bytes.all.current"] / 2**30, st["allocated_bytes.all.current"] / 2**30,
        st["segment.all.current"], st["num_alloc_retries"], gc.get_count()), flush=True)
gc.set_debug(gc.DEBUG_SAVEALL)
n = gc.collect()
types = collections.Counter(type(o).__name__ for o in gc.garbage)
print("unreachable objects found by the collector after 12 steps:", n)
print(types.most_common(25))
tens = [o for o in gc.garbage if isinstance(o, torch.Tensor)]
print("tensors in cycles:", len(tens), sum(t.numel() * t.element_size() for t in tens if t.is_cuda) / 2**30, "GB")
for o in gc.garbage:
    if type(o).__name__.endswith("Backward") or "Function" in type(o).__name__:
        print("  node:", type(o).__name__)
        break
