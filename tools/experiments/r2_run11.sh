#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_j_pytest_all.log 2>&1; echo "all gpu tests rc=$?"; tail -8 gpurun_out/r2_j_pytest_all.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2_j_bench_default.json 2> gpurun_out/r2_j_bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_j_bench_default.json; tail -3 gpurun_out/r2_j_bench_default.err | cut -c1-300
timeout 900 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r2_j_bench_reference.json 2>> gpurun_out/r2_j_bench_default.err; cat gpurun_out/r2_j_bench_reference.json | cut -c1-600
