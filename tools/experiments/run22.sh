#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -15
python tools/timeline.py --batch 512 --steps 3 > gpurun_out/timeline_fold_b512.md 2> gpurun_out/timeline_fold.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side > gpurun_out/ab_fold.json 2> gpurun_out/ab_fold.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side --no-fold > gpurun_out/ab_nofold.json 2> gpurun_out/ab_nofold.err
