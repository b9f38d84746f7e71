#!/bin/bash
# round 2, call 3: full GPU test suite on the new build (single-launch AUC, large prompt sets, graph-captured step),
# latency microbench at the reference's sizes, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_c_pytest_all.log 2>&1; echo "all gpu tests rc=$?"; tail -15 gpurun_out/r2_c_pytest_all.log
timeout 600 python tools/microbench_latency.py > gpurun_out/r2_c_latency.jsonl 2> gpurun_out/r2_c_latency.err; echo "latency rc=$?"; cat gpurun_out/r2_c_latency.jsonl; tail -5 gpurun_out/r2_c_latency.err
timeout 900 python bench.py > gpurun_out/r2_c_bench_default.json 2> gpurun_out/r2_c_bench_default.err; echo "bench rc=$?"; cat gpurun_out/r2_c_bench_default.json; tail -3 gpurun_out/r2_c_bench_default.err
