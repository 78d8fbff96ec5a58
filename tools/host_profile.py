"""cProfile of the host side of the bench step (where do the ~4 ms of enqueue time go?)."""
import cProfile
import os
import pstats
import sys

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
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step.run(step.resident)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
