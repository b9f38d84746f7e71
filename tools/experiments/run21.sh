#!/bin/bash
# batch sweep + CUPTI timeline
set -x
python tools/timeline.py --batch 512 --steps 3 --seq > gpurun_out/timeline_b512.md 2> gpurun_out/timeline_b512.err
for b in 96 192 288 384 512 768; do
  python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline --no-side > gpurun_out/sweep_b$b.json 2> gpurun_out/sweep_b$b.err
done
python tools/timeline.py --batch 96 --steps 3 > gpurun_out/timeline_b96.md 2> gpurun_out/timeline_b96.err
