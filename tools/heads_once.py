"""Launch every head kernel family twice at its microbenchmark size (for one `ncu --set full` pass: tools/run_profile.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eoe_b200 import metrics, ops  # noqa: E402

dev = "cuda"
n = 1 << 20
y = torch.randint(0, 2, (n,), device=dev)
for dt, d in ((torch.float32, 256), (torch.bfloat16, 512)):
    z = (0.05 * torch.randn(n, d, device=dev)).to(dt)
    for _ in range(2):
        ops.hsc_fused(z, y, 0)
        ops.hsc_score(z)
x = torch.randn(1 << 24, 1, device=dev)
yb = torch.randint(0, 2, (1 << 24,), device=dev)
for _ in range(2):
    ops.bce_fused(x, yb, 0)
for dt in (torch.float32, torch.bfloat16):
    z = torch.randn(n, 512, device=dev).to(dt)
    c = torch.nn.functional.normalize(torch.randn(30, 512, device=dev), dim=-1)
    for _ in range(2):
        ops.clip_score(z, c)
        ops.clip_oe_fused(z, y, c, 0, True)
s = 1 - torch.exp(-torch.randn(1000000, device=dev).abs())
ya = (torch.rand(1000000, device=dev) < 0.5).long()
for _ in range(2):
    metrics.roc_auc_device(s, ya)
torch.cuda.synchronize()
print("ok")
