#!/bin/bash
mkdir -p gpurun_out
python tools/auc_small_once.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"auc_small" -s 1 -c 1 -o gpurun_out/prof_auc_small python tools/auc_small_once.py > gpurun_out/ncu_auc_small.log 2>&1
echo rc=$?; tail -3 gpurun_out/ncu_auc_small.log
