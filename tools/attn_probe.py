#!/usr/bin/env python
"""Stand-alone timing + parity of the attention kernels at the encoder's shape (B x 197 tokens x 12 heads):
bandwidth counts the algorithmic bytes (qkv read once, output written once)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from eoe_b200 import _lib as L, encoder as E  # noqa: E402


def main():
    lib = L.lib()
    B, Lq, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 197, 12, 768
    g = torch.Generator(device="cuda").manual_seed(0)
    qkvs = [torch.randn(B * Lq, 3 * W, device="cuda", generator=g).to(torch.bfloat16) for _ in range(2)]
    res = {}
    outs = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, flag in (("cur", 0),):
        lib.eoe_debug_set(flag)
        for i in range(3):
            o = E.attention(qkvs[i & 1], B, Lq, H)
        torch.cuda.synchronize()
        e0.record()
        n = 20
        for i in range(n):
            o = E.attention(qkvs[i & 1], B, Lq, H)
        e1.record()
        torch.cuda.synchronize()
        outs[name] = E.attention(qkvs[0], B, Lq, H)
        us = e0.elapsed_time(e1) / n * 1e3
        res[name] = {"us": round(us, 1), "tflops": round(4.0 * Lq * Lq * 64 * H * B / us / 1e6, 1),
                     "gbs": round((B * Lq * 4 * W * 2) / us / 1e3, 1)}
    lib.eoe_debug_set(0)
    q, k, v = (t.reshape(B, Lq, H, 64).transpose(1, 2) for t in qkvs[0][: 8 * Lq].float().reshape(8, Lq, 3 * W).reshape(8 * Lq, 3 * W).split(W, dim=-1)) \
        if False else (None, None, None)
    x = qkvs[0][: 8 * Lq].float()
    q, k, v = (t.reshape(8, Lq, H, 64).transpose(1, 2) for t in x.split(W, dim=-1))
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v).transpose(1, 2).reshape(8 * Lq, W)
    for name in outs:
        got = outs[name][: 8 * Lq].float()
        res[name]["rel_err_vs_torch"] = float(((got - ref).norm() / ref.norm()).item())
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
