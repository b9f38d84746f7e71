set -x
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; echo rc=$?; tail -3 gpurun_out/bench_r1_a.err; cat gpurun_out/bench_r1_a.json
python bench.py --steps 8 --warmup 3 --patch 32 --prompts 10 --no-cpu-baseline --no-side > gpurun_out/bench_r1_b32.json 2>> gpurun_out/bench_r1_a.err; cat gpurun_out/bench_r1_b32.json
python bench.py --steps 2 --warmup 3 --batch 256 --no-cpu-baseline --no-side > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --batch 256 --no-cpu-baseline --no-side > gpurun_out/ncu.log 2>&1
echo ncu rc=$?
tail -3 gpurun_out/ncu.log
