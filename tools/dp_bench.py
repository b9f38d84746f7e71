#!/usr/bin/env python
"""Data-parallel training at MODEL scale (BASELINE configs 4 and 5):

  cfg4  BCE + outlier exposure, ResNet-18-sized feature model (11.2 M parameters, 44.7 MB of fp32 gradients; the reference
        trains `WideResNet(clf=True)`, models/resnet.py:25-109, with Adam lr 1e-3, main/train_imagenet.py:16,46),
        per-rank batch 128 normal || 128 OE at 3 x 224 x 224, loss / backward w.r.t. the logits by eoe_bce_fwd_bwd
  cfg5  HSC fine-tuning of a ViT-B/16-sized image tower (86 M parameters, 345 MB of fp32 gradients; the reference
        fine-tunes CLIP's visual tower with SGD + Nesterov momentum, training/ad_trainer.py:380-381), per-rank batch
        32 || 32 at 224 x 224 under bf16 autocast, loss / backward w.r.t. the features by eoe_hsc_fwd_bwd

The feature models are plain torch restatements of the two SHAPES (the reference's models stay out of the product: only
the head kernels, the gradient buckets and the collectives are ours).  Per configuration, on N ranks (torchrun):

  step_ms          data-parallel step: forward, fused head kernel, backward with bucketed all-reduce launched from
                   gradient hooks (eoe_b200.dist.GradBuckets), optimiser step     -- CUDA events, max over ranks
  step_local_ms    the same step without any collective (what N = 1 costs on this box, same process)
  allreduce_ms     the bucket all-reduces alone, back to back, nothing to overlap with
  overlap          1 - (step_ms - step_local_ms) / allreduce_ms   (1 = fully hidden behind backward)
  busbw_gbs        2 (N-1)/N x gradient bytes / allreduce time

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dp_bench.py [--bucket-mb 32]
    python tools/dp_bench.py            # N = 1: step_local_ms only
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as tdist  # noqa: E402
from torch import nn  # noqa: E402

from eoe_b200 import dist as edist, ops  # noqa: E402


class BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.c1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.b1 = nn.BatchNorm2d(cout)
        self.c2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.b2 = nn.BatchNorm2d(cout)
        self.down = None
        if stride != 1 or cin != cout:
            self.down = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = torch.relu(self.b1(self.c1(x)))
        y = self.b2(self.c2(y))
        return torch.relu(y + (x if self.down is None else self.down(x)))


def resnet18_sized(out_dim=1):
    """ResNet-18 layout (7x7 stem, 4 stages of 2 basic blocks, 64..512 channels) with the reference's 1-logit classifier
    head (`clf=True`, models/resnet.py:51-53,108)."""
    layers = [nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(), nn.MaxPool2d(3, 2, 1)]
    cin = 64
    for cout, stride in ((64, 1), (128, 2), (256, 2), (512, 2)):
        layers += [BasicBlock(cin, cout, stride), BasicBlock(cout, cout, 1)]
        cin = cout
    layers += [nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(512, 256), nn.Linear(256, out_dim)]
    return nn.Sequential(*layers)


class VitB16Sized(nn.Module):
    """12 x (768 wide, 12 heads, 3072 MLP) pre-norm encoder over 196 + 1 tokens, 512-dimensional projection."""

    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(3, 768, 16, 16, bias=False)
        self.cls = nn.Parameter(torch.zeros(1, 1, 768))
        self.pos = nn.Parameter(torch.randn(1, 197, 768) * 0.02)
        self.blocks = nn.TransformerEncoder(
            nn.TransformerEncoderLayer(768, 12, 3072, dropout=0.0, activation="gelu", batch_first=True, norm_first=True), 12)
        self.ln = nn.LayerNorm(768)
        self.proj = nn.Linear(768, 512, bias=False)

    def forward(self, x):
        x = self.conv(x).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls.expand(x.shape[0], -1, -1), x], 1) + self.pos
        return self.proj(self.ln(self.blocks(x)[:, 0]))


def bench_cfg(name, model, make_batch, loss_fn, opt_fn, steps, warmup, bucket_mb, autocast, dev, ws):
    params = [p for p in model.parameters() if p.requires_grad]
    grad_bytes = sum(p.numel() * 4 for p in params)
    opt = opt_fn(params)
    buckets = edist.GradBuckets(params, bucket_bytes=bucket_mb << 20)
    imgs, lbls = make_batch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step(collective):
        buckets.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            f = model(imgs)
        loss = loss_fn(f.float(), lbls)
        buckets.enabled = collective                             # False: the same step with purely local gradients
        loss.backward()
        buckets.finish()
        opt.step()
        return loss

    def timed(collective, n):
        if ws > 1:
            tdist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            step(collective)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
        if ws > 1:
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t)

    for _ in range(warmup):
        step(ws > 1)
    res = {"config": name, "world_size": ws, "params_m": sum(p.numel() for p in params) / 1e6, "grad_mb": grad_bytes / 1e6,
           "buckets": len(buckets.buckets), "bucket_mb": bucket_mb, "batch_per_rank": int(imgs.shape[0]),
           "autocast_bf16": autocast}
    res["step_local_ms"] = timed(False, steps)
    if ws > 1:
        res["step_ms"] = timed(True, steps)
        # the collectives alone
        tdist.barrier()
        torch.cuda.synchronize()
        for _ in range(2):
            for b in buckets.buckets:
                tdist.all_reduce(b)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            for b in buckets.buckets:
                tdist.all_reduce(b)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        res["allreduce_ms"] = float(t)
        res["overlap"] = 1.0 - (res["step_ms"] - res["step_local_ms"]) / res["allreduce_ms"]
        res["busbw_gbs"] = 2.0 * (ws - 1) / ws * grad_bytes / (res["allreduce_ms"] * 1e-3) / 1e9
        res["images_per_s"] = ws * imgs.shape[0] / (res["step_ms"] * 1e-3)
        res["scaling_efficiency_vs_local_step"] = res["step_local_ms"] / res["step_ms"]
    else:
        res["images_per_s"] = imgs.shape[0] / (res["step_local_ms"] * 1e-3)
    del opt, buckets
    return res


def run(configs=("cfg4", "cfg5"), steps=8, warmup=3, bucket_mb=32):
    rank, local, ws = edist.init_from_env()
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    out = []
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    if "cfg4" in configs:
        model = resnet18_sized(1).to(dev).to(memory_format=torch.channels_last).train()
        out.append(bench_cfg(
            "cfg4: BCE + OE, ResNet-18-sized, 128||128 x 3x224x224 per rank, Adam 1e-3, eoe_bce_fwd_bwd head", model,
            lambda: (torch.randn(256, 3, 224, 224, device=dev, generator=g).contiguous(memory_format=torch.channels_last),
                     (torch.arange(256, device=dev) >= 128).long()),
            lambda f, y: ops.bce_loss(f, y, 0)[0], lambda ps: torch.optim.Adam(ps, lr=1e-3), steps, warmup, bucket_mb, False, dev, ws))
        del model
        torch.cuda.empty_cache()
    if "cfg5" in configs:
        model = VitB16Sized().to(dev).train()
        out.append(bench_cfg(
            "cfg5: HSC + OE, ViT-B/16-sized tower, 32||32 x 3x224x224 per rank, SGD nesterov, bf16 autocast, eoe_hsc_fwd_bwd head",
            model, lambda: (torch.randn(64, 3, 224, 224, device=dev, generator=g), (torch.arange(64, device=dev) >= 32).long()),
            lambda f, y: ops.hsc_loss(f, y, 0)[0],
            lambda ps: torch.optim.SGD(ps, lr=1e-3, momentum=0.9, nesterov=True), steps, warmup, bucket_mb, True, dev, ws))
        del model
        torch.cuda.empty_cache()
    return rank, ws, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bucket-mb", type=int, default=32)
    ap.add_argument("--configs", default="cfg4,cfg5")
    a = ap.parse_args()
    rank, ws, out = run(tuple(a.configs.split(",")), a.steps, a.warmup, a.bucket_mb)
    if rank == 0:
        for r in out:
            print(json.dumps(r), flush=True)
    if ws > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
