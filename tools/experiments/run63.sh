#!/bin/bash
# ViT-B/32 (cfg2): wave-exact batch (1514 images = 296 m-tiles = 4 x 74 CTA pairs) vs 512, plus a launch list
timeout 300 python bench.py --patch 32 --prompts 10 --no-cpu-baseline --no-side > gpurun_out/b32_512.json 2> gpurun_out/b32.err
timeout 300 python bench.py --patch 32 --prompts 10 --batch 757 --no-cpu-baseline --no-side > gpurun_out/b32_757.json 2>> gpurun_out/b32.err
timeout 300 python bench.py --patch 32 --prompts 10 --batch 1514 --no-cpu-baseline --no-side > gpurun_out/b32_1514.json 2>> gpurun_out/b32.err
python - <<'PY'
import json
for b in (512, 757, 1514):
    d = json.load(open(f"gpurun_out/b32_{b}.json"))
    print(b, round(d["value"]), round(d["e2e_u8"]["value"]) if "e2e_u8" in d else None, d["roofline"]["per_kind_tflops"], d["clocks"])
PY
CMD="python bench.py --patch 32 --prompts 10 --batch 1514 --steps 2 --warmup 3 --no-cpu-baseline --no-side"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 140 --csv --log-file gpurun_out/launches_b32_1514.csv $CMD > gpurun_out/ncu_b32.log 2>&1
echo launchlist rc=$?
