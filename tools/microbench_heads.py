"""Microbenchmarks of the HBM-bound kernels (heads, AUC): CUDA-event timing on the launching stream,
inputs larger than L2 (126 MB) or an explicit L2 flush between iterations."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eoe_b200 import metrics, ops  # noqa: E402

PEAK = 6546.6
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []
    for n, d, dt in [(1 << 20, 256, torch.float32), (1 << 22, 256, torch.float32), (1 << 20, 512, torch.bfloat16),
                     (1 << 21, 512, torch.float32), (256, 256, torch.float32)]:
        z = (0.05 * torch.randn(n, d, device=dev)).to(dt)
        y = torch.randint(0, 2, (n,), device=dev)
        med, best = timeit(lambda: ops.hsc_fused(z, y, 0), flush=flush if n * d < (1 << 26) else None)
        bytes_ = 2 * n * d * z.element_size() + 12 * n + 4
        out.append(dict(kernel="hsc_fwd_bwd", n=n, d=d, dtype=str(dt), ms=med, best_ms=best, gbs=bytes_ / med / 1e6,
                        frac=bytes_ / med / 1e6 / PEAK))
        med, best = timeit(lambda: ops.hsc_score(z), flush=flush if n * d < (1 << 26) else None)
        bytes_ = n * d * z.element_size() + 4 * n
        out.append(dict(kernel="hsc_score", n=n, d=d, dtype=str(dt), ms=med, best_ms=best, gbs=bytes_ / med / 1e6,
                        frac=bytes_ / med / 1e6 / PEAK))
        del z, y
    for n in [1 << 24, 1 << 26, 256]:
        x = torch.randn(n, 1, device=dev)
        y = torch.randint(0, 2, (n,), device=dev)
        med, best = timeit(lambda: ops.bce_fused(x, y, 0), flush=flush if n < (1 << 25) else None)
        out.append(dict(kernel="bce_fwd_bwd", n=n, ms=med, best_ms=best, gbs=20 * n / med / 1e6, frac=20 * n / med / 1e6 / PEAK))
        del x, y
    for n, K, dt in [(1 << 20, 2, torch.float32), (1 << 20, 10, torch.float32), (1 << 20, 30, torch.float32),
                     (1 << 20, 30, torch.bfloat16), (10000, 10, torch.float32)]:
        z = torch.randn(n, 512, device=dev).to(dt)
        c = torch.nn.functional.normalize(torch.randn(K, 512, device=dev), dim=-1)
        med, best = timeit(lambda: ops.clip_score(z, c), flush=flush if n * 512 < (1 << 26) else None)
        bytes_ = n * 512 * z.element_size() + 4 * n + K * 2048
        out.append(dict(kernel="clip_score", n=n, K=K, dtype=str(dt), ms=med, best_ms=best, gbs=bytes_ / med / 1e6,
                        frac=bytes_ / med / 1e6 / PEAK))
        y = torch.randint(0, 2, (n,), device=dev)
        med, best = timeit(lambda: ops.clip_oe_fused(z, y, c, 0, True), flush=flush if n * 512 < (1 << 26) else None)
        bytes_ = 2 * n * 512 * z.element_size() + 8 * n + K * 2048
        out.append(dict(kernel="clip_oe_loss", n=n, K=K, dtype=str(dt), ms=med, best_ms=best, gbs=bytes_ / med / 1e6,
                        frac=bytes_ / med / 1e6 / PEAK))
        del z
    for n in [10000, 1 << 20, 1000000, 1 << 24]:
        s = 1 - torch.exp(-torch.randn(n, device=dev).abs())
        y = (torch.rand(n, device=dev) < 0.5).long()
        ws = metrics.AucWorkspace()
        med, best = timeit(lambda: metrics.roc_auc_device(s, y, workspace=ws), flush=flush)
        out.append(dict(kernel="auc", n=n, ms=med, best_ms=best, ms_per_1m=med * 1e6 / n, gbs=12 * n / med / 1e6))
        med, best = timeit(lambda: metrics.roc_auc_device(s, y, workspace=ws, with_prc=True), flush=flush)
        out.append(dict(kernel="auc+ap", n=n, ms=med, best_ms=best, ms_per_1m=med * 1e6 / n))
        s16 = s.half()
        med, best = timeit(lambda: metrics.roc_auc_device(s16, y, workspace=ws), flush=flush)
        out.append(dict(kernel="auc_f16ties", n=n, ms=med, best_ms=best, ms_per_1m=med * 1e6 / n))
    # CLIP text tower (prepare_metric, once per class): K prompts of 77 tokens through 12 blocks of width 512
    from eoe_b200.text_encoder import ClipTextEncoder
    g = torch.Generator().manual_seed(0)
    sd = {"token_embedding.weight": torch.randn(49408, 512, generator=g) * 0.02, "positional_embedding": torch.randn(77, 512, generator=g) * 0.01,
          "ln_final.weight": torch.ones(512), "ln_final.bias": torch.zeros(512), "text_projection": torch.randn(512, 512, generator=g) * 512 ** -0.5}
    for i in range(12):
        p = f"transformer.resblocks.{i}."
        for nm, shp, std in (("attn.in_proj_weight", (1536, 512), 512 ** -0.5), ("attn.out_proj.weight", (512, 512), 0.01),
                             ("mlp.c_fc.weight", (2048, 512), 0.03), ("mlp.c_proj.weight", (512, 2048), 0.01)):
            sd[p + nm] = torch.randn(*shp, generator=g) * std
        for nm, k in (("attn.in_proj_bias", 1536), ("attn.out_proj.bias", 512), ("mlp.c_fc.bias", 2048), ("mlp.c_proj.bias", 512),
                      ("ln_1.bias", 512), ("ln_2.bias", 512)):
            sd[p + nm] = torch.zeros(k)
        sd[p + "ln_1.weight"], sd[p + "ln_2.weight"] = torch.ones(512), torch.ones(512)
    enc = ClipTextEncoder(sd, device=dev)
    for K in (2, 10, 30):
        tok = torch.zeros(K, 77, dtype=torch.int64)
        tok[:, 0], tok[:, 1:6], tok[:, 6] = 49406, 1000, 49407
        tok = tok.to(dev)
        med, best = timeit(lambda: enc(tok))
        out.append(dict(kernel="text_tower", prompts=K, ms=med, best_ms=best, launches=2 + 12 * 7))
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
