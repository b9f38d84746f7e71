#!/usr/bin/env python
"""Anchor of bench.py's CPU arm: the arm times the oracle PORT (oracle/vit.py + oracle/heads.py)
because the reference cannot travel to the GPU box.  Here, in the build container where /root/reference is mounted, the
LIVE reference -- CLIP.encode_image (clip_official/clip/model.py:336-337) + ADClipTrainer.compute_anomaly_score
(training/clip.py:66-79) -- and the port run the same 64 ViT-B/16 images, same weights, same torch threads, interleaved.

    python tools/anchor_cpu_arm.py > profiles/r2_cpu_arm_anchor.json
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import _ref_import, heads as oh, vit as ovit  # noqa: E402


class _Self:
    ad_mode = "leave_one_out"


def main():
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    h = _ref_import.hooks()
    out = {"threads": threads, "images": 64, "chunk": 32}
    for patch, K in ((16, 30), (32, 10)):
        sd = ovit.synth_state_dict(patch, seed=0)
        m = h["CLIP"](512, 224, 12, 768, patch, 77, 49408, 512, 8, 12).eval()
        m.load_state_dict(sd, strict=False)
        g = torch.Generator().manual_seed(1)
        imgs = torch.randn(64, 3, 224, 224, generator=g)
        text = torch.nn.functional.normalize(torch.randn(K, 512, generator=g), dim=-1)

        def live():
            with torch.no_grad():
                for s in range(0, 64, 32):
                    f = m.encode_image(imgs[s:s + 32])
                    h["ADClipTrainer"].compute_anomaly_score(_Self(), f, text)

        def port():
            for s in range(0, 64, 32):
                f = ovit.encode_image(sd, imgs[s:s + 32])
                oh.clip_score(f.numpy(), text.numpy())
        live(); port()
        tl, tp = [], []
        for _ in range(3):
            t0 = time.perf_counter(); live(); tl.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); port(); tp.append(time.perf_counter() - t0)
        out[f"vitb{patch}_K{K}"] = {"live_reference_img_per_s": 64 / min(tl), "oracle_port_img_per_s": 64 / min(tp),
                                    "port_over_live": min(tl) / min(tp)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
