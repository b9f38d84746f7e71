#!/usr/bin/env python
"""Kernel timeline of the zero-shot scoring step (CUPTI through torch.profiler): per-kernel busy time, launch gaps and
share of the step, measured with the kernels running back to back (unlike ncu, which serialises and cold-caches them).

    python tools/timeline.py --batch 512 --steps 3 [--patch 16] > gpurun_out/timeline.md
"""
import argparse
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("eoe::", "")
    return name[:60]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--patch", type=int, default=16)
    ap.add_argument("--prompts", type=int, default=30)
    ap.add_argument("--seq", action="store_true", help="also print the launch sequence of one step")
    ap.add_argument("--no-fold", action="store_true")
    a = ap.parse_args()
    from eoe_b200.encoder import ClipImageEncoder
    from eoe_b200.synth import random_vit_state_dict
    dev = torch.device("cuda", 0)
    enc = ClipImageEncoder(random_vit_state_dict(a.patch, seed=0), device=dev, max_batch=a.batch,
                           fold_layernorm=not a.no_fold)
    imgs = [torch.randn(a.batch, 3, 224, 224, device=dev) for _ in range(2)]
    text = torch.nn.functional.normalize(torch.randn(a.prompts, 512, device=dev), dim=-1)
    out = torch.empty(a.batch, device=dev)
    for k in range(3):
        enc.score(imgs[k & 1], text, out=out)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for k in range(a.steps):
            enc.score(imgs[k & 1], text, out=out)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    evs = sorted(evs, key=lambda e: e.time_range.start)
    evs = [e for e in evs if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    if not evs:
        print("no kernel events captured")
        return
    t0, t1 = evs[0].time_range.start, evs[-1].time_range.end
    span = (t1 - t0) / a.steps
    busy = collections.defaultdict(float)
    cnt = collections.Counter()
    gap_after = collections.defaultdict(float)
    gaps = 0.0
    for i, e in enumerate(evs):
        d = e.time_range.end - e.time_range.start
        busy[short(e.name)] += d
        cnt[short(e.name)] += 1
        if i + 1 < len(evs):
            g = max(0.0, evs[i + 1].time_range.start - e.time_range.end)
            gaps += g
            gap_after[short(e.name)] += g
    print(f"# Timeline, ViT-B/{a.patch}, {a.batch} images/step, {a.steps} steps (torch.profiler / CUPTI, kernels back to back)\n")
    print(f"step span {span / 1e3:.3f} ms; kernel-busy {sum(busy.values()) / a.steps / 1e3:.3f} ms; "
          f"gaps {gaps / a.steps / 1e3:.3f} ms ({len(evs) // a.steps} launches/step)\n")
    print("| kernel | launches/step | us/launch | ms/step | share of span | gap after, us/launch |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, v in sorted(busy.items(), key=lambda kv: -kv[1]):
        print(f"| `{k}` | {cnt[k] / a.steps:.1f} | {v / cnt[k]:.1f} | {v / a.steps / 1e3:.3f} | {100 * v / a.steps / span:.1f}% |"
              f" {gap_after[k] / cnt[k]:.1f} |")
    if a.seq:
        n = len(evs) // a.steps
        print("\nLaunch sequence of the last step (start us, duration us, kernel):\n")
        base = evs[-n].time_range.start
        for e in evs[-n:]:
            print(f"    {e.time_range.start - base:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {short(e.name)}")


if __name__ == "__main__":
    main()
