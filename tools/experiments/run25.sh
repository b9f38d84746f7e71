#!/bin/bash
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side > gpurun_out/plain25.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -s 70 -c 5 -o gpurun_out/prof_fold_a python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side > gpurun_out/ncu25.log 2>&1
echo rc=$?
