"""The FPN-level Sinkhorn shape (N = 64, D = 4096, 24 problems = batch 8 x 3 terms), forward + gradient, a few times -- for
`ncu --metrics gpu__time_duration.sum -k regex:sinkhorn` (per-kernel split of the microbench row)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_intertwiner_b200 as fi  # noqa: E402

x = torch.randn(24, 64, 4096, device="cuda").abs().requires_grad_()
y = torch.randn(24, 64, 4096, device="cuda").abs().requires_grad_()
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for _ in range(3):
    flush.add_(1.0)
    fi.sinkhorn_loss(x, y, 1.0, 5)
torch.cuda.synchronize()
