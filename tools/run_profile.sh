#!/bin/bash
# Round profile of the current build: launch list + ncu --set full of the dominant kernels + default bench lines.
# usage: tools/run_profile.sh <tag>
tag=${1:-r1}
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side"
timeout 300 $CMD > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 140 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu1_$tag.log 2>&1
echo launchlist rc=$?
$CMD > gpurun_out/plain2_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attention_tc" -s 50 -c 6 -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu2_$tag.log 2>&1
echo full rc=$?
python bench.py > gpurun_out/bench_${tag}_b16.json 2> gpurun_out/bench_${tag}.err; tail -2 gpurun_out/bench_${tag}.err
python bench.py --patch 32 --prompts 10 --no-cpu-baseline --no-side > gpurun_out/bench_${tag}_b32.json 2>> gpurun_out/bench_${tag}.err
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_${tag}_reference.json 2>> gpurun_out/bench_${tag}.err
python tools/microbench_heads.py > gpurun_out/microbench_heads_${tag}.jsonl 2>> gpurun_out/bench_${tag}.err
# head kernels: one ncu --set full pass over every head kernel family at its microbenchmark size (second launch of each)
python tools/heads_once.py > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"clip_score_mma|clip_oe_loss_mma|hsc_rows_kernel|bce_kernel|auc_sort_pass" -c 28 -o gpurun_out/prof_heads_$tag python tools/heads_once.py > gpurun_out/ncu3_$tag.log 2>&1
echo heads rc=$?
