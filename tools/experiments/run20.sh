set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/launches_r1_final.csv $CMD > gpurun_out/ncu1.log 2>&1
echo launchlist rc=$?
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attention_tc" -s 60 -c 6 -o gpurun_out/prof_final_r1 $CMD > gpurun_out/ncu2.log 2>&1
echo full rc=$?
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cat gpurun_out/bench_r1_final.json
python bench.py --patch 32 --prompts 10 --no-cpu-baseline --no-side > gpurun_out/bench_r1_final_b32.json 2>> gpurun_out/bench_final.err; cat gpurun_out/bench_r1_final_b32.json
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_r1_reference.json; cat gpurun_out/bench_r1_reference.json
