#!/usr/bin/env python
"""Interleaved A/B timing of encoder variants in ONE process (same box, same power / thermal state).

Boxes differ by several percent in their power-capped clocks, so two separate bench runs cannot resolve a 2 % change.
Here every variant is built once and timed in alternating blocks (A B A B ...) of `--block` steps for `--rounds`
rounds; the first round is discarded.  Prints ms/step and images/s per variant plus the per-GEMM-kind TFLOP/s of a
final instrumented pass.

    python tools/ab_encoder.py --variants fold,nofold --batch 512
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="fold,nofold")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--block", type=int, default=10)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--patch", type=int, default=16)
    ap.add_argument("--prompts", type=int, default=30)
    ap.add_argument("--debug-flags", default="", help="name=flags,... : variants that only differ by eoe_debug_set flags (e.g. pdl=0,nopdl=32)")
    a = ap.parse_args()
    from eoe_b200.encoder import ClipImageEncoder
    from eoe_b200.synth import random_vit_state_dict
    dev = torch.device("cuda", 0)
    sd = random_vit_state_dict(a.patch, seed=0)
    from eoe_b200 import _lib
    dbg = {}
    if a.debug_flags:
        for kv in a.debug_flags.split(","):
            k, v = kv.split("=")
            dbg[k] = int(v)
        names = list(dbg)
        base = ClipImageEncoder(sd, device=dev, max_batch=a.batch)
        encs = {n: base for n in names}
    else:
        names = a.variants.split(",")
        # variant names: fold / nofold (LayerNorm fold on / off, fp16) or an operand dtype: f16 / bf16 / f16x2 (precise mode)
        ops = {"f16": torch.float16, "bf16": torch.bfloat16, "f16x2": "f16x2"}
        encs = {n: ClipImageEncoder(sd, device=dev, max_batch=a.batch, fold_layernorm=(n != "nofold"),
                                    operand_dtype=ops.get(n, torch.float16)) for n in names}
    imgs = [torch.randn(a.batch, 3, 224, 224, device=dev) for _ in range(2)]
    text = torch.nn.functional.normalize(torch.randn(a.prompts, 512, device=dev), dim=-1)
    out = torch.empty(a.batch, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = {n: [] for n in names}
    for r in range(a.rounds + 1):
        for n in names:
            enc = encs[n]
            if dbg:
                _lib.lib().eoe_debug_set(dbg[n])
            e0.record()
            for k in range(a.block):
                enc.score(imgs[k & 1], text, out=out)
            e1.record()
            torch.cuda.synchronize()
            if r > 0:
                ms[n].append(e0.elapsed_time(e1) / a.block)
    res = {}
    for n in names:
        m = sum(ms[n]) / len(ms[n])
        res[n] = {"ms_per_step": round(m, 4), "images_per_s": round(a.batch / m * 1e3, 1),
                  "blocks": [round(x, 3) for x in ms[n]]}
    for n in names:
        enc = encs[n]
        if dbg:
            _lib.lib().eoe_debug_set(dbg[n])
        enc.profile(True)
        for k in range(a.block):
            enc.score(imgs[k & 1], text, out=out)
        torch.cuda.synchronize()
        prof = enc.profile_read()
        enc.profile(False)
        res[n]["gemm_tflops"] = {k: round(v[2] / (v[0] * 1e-3) / 1e12) for k, v in prof.items() if v[0] > 0}
        res[n]["gemm_us"] = {k: round(1e3 * v[0] / v[1], 1) for k, v in prof.items() if v[1] > 0}
    if len(names) == 2:
        res["ratio_%s_over_%s" % (names[0], names[1])] = round(res[names[0]]["images_per_s"] / res[names[1]]["images_per_s"], 4)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
