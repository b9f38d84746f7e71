#!/usr/bin/env python
"""Timing of the warp-level attention kernel at the ViT-B/32 shape (B x 50 tokens x 12 heads) and the text tower's (causal, 77)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from eoe_b200 import encoder as E  # noqa: E402


def main():
    from eoe_b200 import _lib as L
    res = {}
    flag = int(sys.argv[1]) if len(sys.argv) > 1 else 0         # 128: warp-level kernel instead of the tcgen05 pair kernel
    L.lib().eoe_debug_set(flag)
    for B, Lq, H in ((1514, 50, 12), (512, 50, 12), (30, 64, 8)):
        W = H * 64
        g = torch.Generator(device="cuda").manual_seed(0)
        qkvs = [torch.randn(B * Lq, 3 * W, device="cuda", generator=g).to(torch.bfloat16) for _ in range(2)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            E.attention(qkvs[i & 1], B, Lq, H)
        torch.cuda.synchronize()
        e0.record()
        n = 50
        for i in range(n):
            E.attention(qkvs[i & 1], B, Lq, H)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        res[f"B{B}_L{Lq}"] = {"us": round(us, 1), "gbs": round((B * Lq * 4 * W * 2) / us / 1e3, 1)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
