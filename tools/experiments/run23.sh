#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side > gpurun_out/ab2_fold.json 2> gpurun_out/ab2_fold.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side --no-fold > gpurun_out/ab2_nofold.json 2> gpurun_out/ab2_nofold.err
