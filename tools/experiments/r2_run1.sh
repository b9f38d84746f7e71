#!/bin/bash
# round 2, call 1: bf16 vs f16 bench lines on the round-1 build (is f16 "within noise" on images/s?)
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dtype bf16 > gpurun_out/r2_a_bench_bf16.json 2> gpurun_out/r2_a_bench_bf16.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side --dtype f16 > gpurun_out/r2_a_bench_f16.json 2> gpurun_out/r2_a_bench_f16.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side --dtype bf16 > gpurun_out/r2_a_bench_bf16_2.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side --dtype f16 > gpurun_out/r2_a_bench_f16_2.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side --dtype f16 --patch 32 --prompts 10 > gpurun_out/r2_a_bench_f16_b32.json 2>/dev/null
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side --dtype bf16 --patch 32 --prompts 10 > gpurun_out/r2_a_bench_bf16_b32.json 2>/dev/null
nvidia-smi -L; lscpu | head -20; numactl -H 2>/dev/null | head; nvidia-smi topo -m
