"""Launch the single-launch AUC kernel a few times at n = 10 000 (for an `ncu --set full --import-source on -k regex:auc_small` pass)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eoe_b200 import metrics  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
s = 1 - torch.exp(-torch.randn(n, device="cuda").abs())
y = (torch.rand(n, device="cuda") < 0.5).long()
ws = metrics.AucWorkspace()
for _ in range(3):
    metrics.roc_auc_device(s, y, workspace=ws)
torch.cuda.synchronize()
print("ok")
