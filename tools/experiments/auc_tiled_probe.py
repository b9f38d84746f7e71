"""Tiled AUC pipeline (n > 49 152): time per call at several sizes, L2 flushed between calls (as tools/microbench_heads.py),
and a sklearn bit-exactness check.  EOE_B200_LIB selects the library variant."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from eoe_b200 import _lib, metrics  # noqa: E402

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tag = os.environ.get("EOE_B200_LIB", "default")
SIZES = [int(a) for a in sys.argv[1:]] or [65536, 262144, 1000000, 1 << 20, 1 << 22, 1 << 24]
for n in SIZES:
    s = 1 - torch.exp(-torch.randn(n, device=dev).abs())
    y = (torch.rand(n, device=dev) < 0.5).long()
    ws = metrics.AucWorkspace()
    rows = {}
    for name, kw, sc in (("auc", {}, s), ("auc+ap", dict(with_prc=True), s), ("auc_f16ties", {}, s.half())):
        for _ in range(3):
            metrics.roc_auc_device(sc, y, workspace=ws, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); metrics.roc_auc_device(sc, y, workspace=ws, **kw); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        rows[name] = dict(ms=ts[len(ts) // 2], best_ms=ts[0], ms_per_1m=ts[len(ts) // 2] * 1e6 / n)
    exact = None
    if n <= (1 << 22):
        from sklearn.metrics import roc_auc_score
        got = metrics.roc_auc(s, y)
        exact = bool(got == roc_auc_score(y.cpu().numpy(), s.cpu().numpy()))
    print(json.dumps(dict(lib=tag, build=_lib.lib().eoe_build_id().decode(), n=n, bit_exact_vs_sklearn=exact, **rows)))
