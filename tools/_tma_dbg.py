import sys, torch
sys.path.insert(0, '/root/repo')
import feature_intertwiner_b200 as fi
g = torch.Generator().manual_seed(0)
img = torch.randn(2, 128, 60, 72, generator=g).cuda()
boxes = torch.tensor([[0.1, 0.1, 0.3, 0.3], [0.2, 0.2, 0.25, 0.26]]).cuda()
ind = torch.tensor([0, 1], dtype=torch.int32).cuda()
out = fi.crop_and_resize(img, boxes, ind, 7, 7)
torch.cuda.synchronize()
print("ok", out.sum().item())
