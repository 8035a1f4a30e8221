"""Short launch sequence for `ncu --set full` of the selection kernels (never a bench value)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402,F401  (sets sys.path for the package)
import torch  # noqa: E402
from lgcnhs_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(1)
U, M = 148 * 128, 16384
xu = (torch.randn(U, 64) * 0.1).to(dev)
xi = (torch.randn(M, 64) * 0.1).to(dev)
su = torch.randint(U, (U * 50,)).to(dev)
si = torch.randint(M, (U * 50,)).to(dev)
seen = ops.seen_csr(su, si, U, M)
for _ in range(2):
    idx, val = ops.score_topk(xu, xi, 20, seen)
F = torch.rand(6040, 3708, device=dev)[:, :3706]
mask = ops.ExclusionMask.from_pairs(torch.randint(6040, (900000,)).to(dev), torch.randint(3706, (900000,)).to(dev), 6040, 3706)
for _ in range(2):
    i2, v2 = ops.topk_rows(F, 20, mask)
torch.cuda.synchronize()
print("ok", int(idx.sum()), int(i2.sum()))
