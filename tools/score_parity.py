#!/usr/bin/env python
"""End-to-end SCORE parity of the CLIP zero-shot path, on the CPU (test infrastructure; imports oracle/).

north_star states 1e-3 relative on scores.  Scores are softmax(100 cos), so a score's relative error is the error of a
difference of logits = 100 x the error of a cosine: 1e-3 on the score needs ~1e-5 on the cosine.  This tool measures, for
BASELINE configs 2 (ViT-B/32, K = 10) and 3 (ViT-B/16, K = 30) on the seeded images of oracle.golden_inputs:

  * the precision-matched oracle of our CUDA path (oracle.vit.encode_image operand_dtype = bf16 / f16, LayerNorm fold)
    against the fp32 oracle: feature rel-L2, score relative error (median / p90 / max / fraction within 1e-3), AUC delta;
  * the same with ONE rounding point active at a time (oracle.vit.ROUND_POINTS): which stored 16-bit tensor the error
    comes from;
  * the precise mode's emulation (oracle.vit.F16X2: every stored 16-bit tensor an fp16 (hi, lo) pair) against the same;
  * the reference's OWN GPU precision (fp16 weights and fp16 residual stream, model.py:371-392), both as the emulation
    oracle.vit.encode_image_ref_fp16 and, where /root/reference is mounted, as the live reference run in half.

    python tools/score_parity.py [--images 64] [--out profiles/r2_score_parity_attribution.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import auc as oauc  # noqa: E402
from oracle import golden_inputs as gi  # noqa: E402
from oracle import heads as oh  # noqa: E402
from oracle import vit as ovit  # noqa: E402


def stats(scores, want, feats=None, feats_want=None, labels=None):
    rel = np.abs(scores.astype(np.float64) - want.astype(np.float64)) / np.abs(want.astype(np.float64))
    out = {"score_rel_median": float(np.median(rel)), "score_rel_p90": float(np.quantile(rel, 0.9)),
           "score_rel_max": float(rel.max()), "frac_within_1e-3": float((rel <= 1e-3).mean())}
    if feats is not None:
        out["feat_rel_l2"] = float(np.linalg.norm(feats - feats_want) / np.linalg.norm(feats_want))
        c = (feats * feats_want).sum(1) / np.linalg.norm(feats, axis=1) / np.linalg.norm(feats_want, axis=1)
        out["one_minus_cos_max"] = float((1 - c).max())
    if labels is not None:
        out["auc"] = oauc.roc_auc(labels, scores)
        out["auc_minus_fp32"] = out["auc"] - oauc.roc_auc(labels, want)
    return out


def run_cfg(patch, K, n_img, live):
    sd = ovit.synth_state_dict(patch, seed=gi.VIT_WEIGHT_SEED)
    imgs, text, labels = gi.score_parity_inputs(K, n_img)
    t0 = time.time()
    f32 = ovit.encode_image(sd, imgs).numpy()
    s32 = oh.clip_score(f32, text)
    res = {"patch": patch, "K": K, "images": n_img, "fp32_oracle_s": time.time() - t0,
           "score_range": [float(s32.min()), float(s32.max())]}
    for name, dt in (("f16", torch.float16), ("bf16", torch.bfloat16)):
        f = ovit.encode_image(sd, imgs, operand_dtype=dt, fold_layernorm=True).numpy()
        res[f"ours_{name}"] = stats(oh.clip_score(f, text), s32, f, f32, labels)
        per = {}
        for pt in ovit.ROUND_POINTS:
            f = ovit.encode_image(sd, imgs, operand_dtype=dt, fold_layernorm=True, points=(pt,)).numpy()
            st = stats(oh.clip_score(f, text), s32, f, f32)
            per[pt] = {"feat_rel_l2": st["feat_rel_l2"], "score_rel_median": st["score_rel_median"]}
        res[f"ours_{name}_one_point_at_a_time"] = per
        print(patch, name, res[f"ours_{name}"], flush=True)
    # the precise mode (operand dtype EOE_F16X2: fp16 (hi, lo) pairs at every stored 16-bit tensor), all points and all but one
    f = ovit.encode_image(sd, imgs, operand_dtype=ovit.F16X2, fold_layernorm=True).numpy()
    res["ours_f16x2"] = stats(oh.clip_score(f, text), s32, f, f32, labels)
    print(patch, "f16x2", res["ours_f16x2"], flush=True)
    f = ovit.encode_image_ref_fp16(sd, imgs).numpy()
    res["reference_fp16_path_emulated"] = stats(oh.clip_score(f, text), s32, f, f32, labels)
    print(patch, "ref fp16 emu", res["reference_fp16_path_emulated"], flush=True)
    if live:
        from oracle import _ref_import
        from oracle.make_golden import live_half_features
        f = live_half_features(_ref_import.hooks(), patch, sd, imgs)
        res["reference_fp16_path_live_cpu_half"] = stats(oh.clip_score(f, text), s32, f, f32, labels)
        print(patch, "ref fp16 live", res["reference_fp16_path_live_cpu_half"], flush=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_score_parity_attribution.json"))
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import _ref_import
    live = _ref_import.available()
    out = {"what": "CPU study: score error of our rounding points vs the fp32 oracle, and of the reference's own fp16 path",
           "cfg2_vitb32_K10": run_cfg(32, 10, args.images, live), "cfg3_vitb16_K30": run_cfg(16, 30, args.images, live)}
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print("written", args.out)


if __name__ == "__main__":
    main()
