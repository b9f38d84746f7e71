"""Latency at the sizes the REFERENCE actually runs: heads at n = 256 rows
(datasets/bases.py:591-597: 128 normal || 128 OE), the AUC at 3 000 - 10 000 scores per class and epoch
(training/ad_trainer.py:452-455,516-522), and one whole training step of a CIFAR-sized CNN eager vs CUDA graph.

CUDA-event timing on the launching stream, median of `iters` after warm-up, L2 NOT flushed (these calls follow the
kernels that produced their inputs, so warm L2 is the operating condition).  sklearn / torch-CPU on the host beside it.
One JSON object per line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eoe_b200 import _lib, metrics, ops  # noqa: E402


def gpu_us(fn, iters=50, warmup=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def gpu_us_b2b(fn, calls=20, iters=10, warmup=3):
    """device time per call with `calls` launches in flight between the events: what the GPU needs, not what Python needs"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(calls):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / calls)
    ts.sort()
    return ts[len(ts) // 2]


def graph_us(fn, iters=20):
    """the same call replayed from a CUDA graph: device time without any host launch cost"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / 10)
    ts.sort()
    return ts[len(ts) // 2]


def host_us(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e6)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    dev = "cuda"
    out = []
    lib = _lib.lib()
    # ---- AUC at the reference's sizes
    from sklearn.metrics import auc, roc_curve
    for n in (1000, 3000, 10000, 16384, 32768, 65536, 131072):
        rng = np.random.default_rng(n)
        s_np = (1 - np.exp(-np.abs(rng.standard_normal(n)))).astype(np.float32)
        y_np = (rng.random(n) < 0.5).astype(np.int64)
        s, y = torch.from_numpy(s_np).to(dev), torch.from_numpy(y_np).to(dev)
        ws = metrics.AucWorkspace()
        l0 = lib.eoe_launch_count()
        metrics.roc_auc_device(s, y, workspace=ws)
        launches = lib.eoe_launch_count() - l0
        med, best = gpu_us(lambda: metrics.roc_auc_device(s, y, workspace=ws))
        row = dict(kernel="auc", n=n, us=med, best_us=best, launches=launches,
                   device_us_back_to_back=gpu_us_b2b(lambda: metrics.roc_auc_device(s, y, workspace=ws)),
                   device_us_graph_replay=graph_us(lambda: metrics.roc_auc_device(s, y, workspace=ws)))
        if n <= _lib.EOE_AUC_SINGLE_LAUNCH_MAX:
            info = metrics.roc_auc_device(s, y, workspace=ws)[1].cpu()
            row["phase_clocks_keys_sort__scans__terms_sum"] = [int(v) for v in info[5:8]]
            med_t, best_t = gpu_us(lambda: metrics.roc_auc_device(s, y, workspace=ws, force_tiled=True))
            row.update(tiled_us=med_t, tiled_best_us=best_t,
                       tiled_device_us_graph_replay=graph_us(lambda: metrics.roc_auc_device(s, y, workspace=ws, force_tiled=True)))
        if n <= 131072:
            row["cluster_device_us_graph_replay"] = graph_us(lambda: metrics.roc_auc_device(s, y, workspace=ws, force_cluster=True))
        if n <= 16384:
            row["one_cta_device_us_graph_replay"] = graph_us(lambda: metrics.roc_auc_device(s, y, workspace=ws, force_single_cta=True))
        med_prc, _ = gpu_us(lambda: metrics.roc_auc_device(s, y, workspace=ws, with_prc=True))
        row["with_ap_us"] = med_prc
        med_h, best_h = host_us(lambda: auc(*roc_curve(y_np, s_np)[:2]))
        row.update(sklearn_host_us=med_h, sklearn_host_best_us=best_h)
        # end to end as the trainer uses it: device scores -> python float (one 80-byte D2H)
        t_e2e, _ = host_us(lambda: metrics.roc_auc(s, y), iters=50, warmup=10)
        row["roc_auc_to_host_us"] = t_e2e
        assert metrics.roc_auc(s, y) == auc(*roc_curve(y_np, s_np)[:2])
        out.append(row)
    # ---- heads at n = 256
    for name, d, dt in (("hsc", 256, torch.float32), ("hsc", 512, torch.float16)):
        z = (0.05 * torch.randn(256, d, device=dev)).to(dt)
        y = (torch.arange(256, device=dev) >= 128).long()
        med, best = gpu_us(lambda: ops.hsc_fused(z, y, 0))
        out.append(dict(kernel="hsc_fwd_bwd_score", n=256, d=d, dtype=str(dt), us=med, best_us=best, launches=1,
                        device_us_graph_replay=graph_us(lambda: ops.hsc_fused(z, y, 0))))
    x = torch.randn(256, 1, device=dev)
    y = (torch.arange(256, device=dev) >= 128).long()
    med, best = gpu_us(lambda: ops.bce_fused(x, y, 0))
    out.append(dict(kernel="bce_fwd_bwd_score", n=256, us=med, best_us=best, launches=1,
                    device_us_graph_replay=graph_us(lambda: ops.bce_fused(x, y, 0))))
    for n, K in ((128, 10), (128, 30), (10000, 10), (3000, 30), (128, 100), (128, 200)):
        z = torch.randn(n, 512, device=dev)
        c = torch.nn.functional.normalize(torch.randn(K, 512, device=dev), dim=-1)
        yy = torch.randint(0, 2, (n,), device=dev)
        med, best = gpu_us(lambda: ops.clip_score(z, c))
        med_l, best_l = gpu_us(lambda: ops.clip_oe_fused(z, yy, c, 0, True))
        out.append(dict(kernel="clip_score", n=n, K=K, us=med, best_us=best, clip_oe_loss_us=med_l))
    # ---- the same objectives through torch ops (what the reference launches: 17 fwd + 28 bwd aten ops for HSC)
    z = (0.05 * torch.randn(256, 256, device=dev)).requires_grad_(True)
    y = (torch.arange(256, device=dev) >= 128).long()

    def torch_hsc():
        z.grad = None
        d = torch.sqrt(torch.norm(z, p=2, dim=1) ** 2 + 1) - 1
        sc = 1 - torch.exp(-d)
        torch.where(y == 0, d, -torch.log(sc + 1e-9)).mean().backward()
        return 1 - torch.exp(-(torch.sqrt(torch.norm(z.detach(), p=2, dim=1) ** 2 + 1) - 1))
    med, best = gpu_us(torch_hsc)
    out.append(dict(kernel="hsc_torch_ops_gpu (reference's op sequence)", n=256, d=256, us=med, best_us=best))

    # ---- one training step of a CIFAR-sized CNN (cfg1 shapes): eager vs captured graph
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from test_gpu_trainers import _cnn32_like
    from eoe_b200.training import TRAINER
    g = torch.Generator().manual_seed(0)
    loader = [(torch.randn(256, 3, 32, 32, generator=g).to(dev), (torch.arange(256) >= 128).long().to(dev), None) for _ in range(40)]
    for graph in (False, True):
        model = _cnn32_like()
        tr = TRAINER["hsc"](model, epochs=1, lr=1e-3, device=dev, graph_step=graph)
        tr.train_cls(model, loader[:8], nominal_label=0)            # warm-up (cudnn autotune)
        torch.cuda.synchronize()

        def run(epochs):
            t0 = time.perf_counter()
            tr.train_cls(model, loader, nominal_label=0, epochs=epochs)
            torch.cuda.synchronize()
            return time.perf_counter() - t0
        t1 = min(run(1) for _ in range(2))
        t4 = min(run(4) for _ in range(2))
        out.append(dict(kernel="train_cls step, CNN32-shaped model, batch 128+128, HSC" + (" [CUDA graph]" if graph else " [eager]"),
                        us_per_step=(t4 - t1) / (3 * len(loader)) * 1e6, first_epoch_us_per_step=t1 / len(loader) * 1e6,
                        note="wall clock per step of epochs 2-4 (40 batches each, incl. the epoch-end AUC); the first epoch "
                             "carries the one-off capture"))
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
