set -x
CMD="python bench.py --steps 2 --warmup 3 --batch 256 --no-cpu-baseline --no-side"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_r1_v2.csv $CMD > gpurun_out/ncu1.log 2>&1
echo launchlist rc=$?
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 60 -c 5 -o gpurun_out/prof_gemm_r1 $CMD > gpurun_out/ncu2.log 2>&1
echo full rc=$?
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out/
