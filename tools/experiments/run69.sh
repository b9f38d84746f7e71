#!/bin/bash
# final build (v12): launch list + microbench lines (the ncu --set full captures of this build's kernels are r1_v10_* / r1_v12_*)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side"
timeout 300 $CMD > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 140 --csv --log-file gpurun_out/launches_v12.csv $CMD > gpurun_out/ncu1_v12.log 2>&1
echo launchlist rc=$?
timeout 200 python tools/microbench_heads.py > gpurun_out/microbench_heads_v12.jsonl 2> gpurun_out/mb_v12.err; echo mb rc=$?
